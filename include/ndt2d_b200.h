/*
 * ndt2d_b200.h -- C ABI of libndt2d_b200.so: the B200 (sm_100a) backend for
 * ndt_2d's scan-matching hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain host pointers
 * and sizes (doubles, caller-owned), returns an int status and never throws.
 * There is no CPU fallback: if no CUDA device is usable, create() fails with
 * NDT2D_ERR_NO_DEVICE and nothing else can be called.
 *
 * Each function cites the reference interface it replaces; file:line are
 * relative to the reference tree (mikeferguson/ndt_2d).  INTEGRATION.md shows
 * the C++ plugin class (ndt_2d::ScanMatcherNDT / ndt_2d::ParticleFilter) that
 * binds these into the reference's pluginlib seam.
 *
 * Conventions
 *   pose3      : {x, y, theta}                       (pose_2d.hpp:35-55)
 *   pts_xy     : interleaved {x0, y0, x1, y1, ...} in the sensor frame,
 *                exactly the layout of std::vector<ndt_2d::Point>
 *                (point.hpp:35-51), so `&points[0].x` can be passed as is
 *   cov9       : 3x3 row-major
 *   A handle serialises its own calls (internal mutex + one CUDA stream);
 *   distinct handles are independent.
 */
#ifndef NDT2D_B200_H_
#define NDT2D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDT2D_API __attribute__((visibility("default")))

typedef enum ndt2d_status
{
  NDT2D_OK = 0,
  NDT2D_ERR_INVALID = 1,    /* bad argument (null pointer, non-positive resolution ...) */
  NDT2D_ERR_NO_DEVICE = 2,  /* no usable CUDA device: the product has no CPU path */
  NDT2D_ERR_CUDA = 3,       /* a CUDA runtime call failed; see ndt2d_last_error() */
  NDT2D_ERR_NO_MAP = 4,     /* search/score entry called before add_scans (reference
                               returns 0.0 and leaves outputs untouched:
                               scan_matcher_ndt.cpp:80,159) */
  NDT2D_ERR_SIZE = 5,       /* grid or scan larger than the device layout supports */
  NDT2D_ERR_STATE = 6       /* call sequence error (e.g. fetch before stage) */
} ndt2d_status;

/* Parameters of one ScanMatcherNDT instance: the six ROS parameters the
 * reference declares under "<name>." (scan_matcher_ndt.cpp:37-44, same
 * defaults) plus range_max (:46) and the CUDA placement. */
#define NDT2D_MAX_DEVICES 16

typedef struct ndt2d_params
{
  double ndt_resolution;             /* default 0.25   */
  double search_angular_resolution;  /* default 0.0025 */
  double search_angular_size;        /* default 0.1    */
  double search_linear_resolution;   /* default 0.005  */
  double search_linear_size;         /* default 0.05   */
  int laser_max_beams;               /* default 100    */
  double range_max;                  /* initialize(..., range_max) */
  int device;                        /* CUDA device ordinal, -1 = current device */
  void * stream;                     /* cudaStream_t to run on; NULL = handle-owned stream */
  int kernel_variant;                /* 0 = auto: warp-per-region kernel; for small searches
                                        the window kernel (thread per candidate, windows a
                                        few cells wide) or the dense warp-per-candidate
                                        kernel (3 forces dense, 4 forces region, 5 forces
                                        window where eligible, else dense);
                                        1 = plain per-candidate kernel (reference arithmetic
                                        per evaluation; the on-device cross-check) */
  int n_devices;                     /* > 1: ONE handle drives the GPUs devices[0 .. n_devices-1] of
                                        this process (device / stream are ignored): the model and the
                                        query scan are replicated, a large matchScan scores theta slices
                                        r, r + N, ... on device r, and the 128-byte partial records are
                                        exchanged by peer stores from each search's last kernel (peer
                                        access enabled between the devices; host combine otherwise).
                                        Everything else runs on devices[0].  0 / 1 = single device. */
  int devices[NDT2D_MAX_DEVICES];
} ndt2d_params;

typedef struct ndt2d_matcher ndt2d_matcher;
typedef struct ndt2d_filter ndt2d_filter;

/* Number of doubles in one partial search record (see ndt2d_matcher_search_staged). */
#define NDT2D_PARTIAL_DOUBLES 16

NDT2D_API const char * ndt2d_version(void);
/* Message of the last failing CUDA call on this thread ("" if none). */
NDT2D_API const char * ndt2d_last_error(void);
/* Number of CUDA devices visible (0 => every create() will fail). */
NDT2D_API int ndt2d_device_count(void);

/* Fills the reference's defaults (scan_matcher_ndt.cpp:37-44); range_max = 0,
 * device = -1, stream = NULL. */
NDT2D_API void ndt2d_default_params(ndt2d_params * p);

/* ------------------------------------------------------------------------
 * ScanMatcherNDT  (include/ndt_2d/scan_matcher.hpp:42-91,
 *                  src/scan_matcher_ndt.cpp:35-183)
 * ---------------------------------------------------------------------- */

/* replaces ScanMatcherNDT::initialize (scan_matcher_ndt.cpp:35-47) */
NDT2D_API int ndt2d_matcher_create(const ndt2d_params * params, ndt2d_matcher ** out);
NDT2D_API int ndt2d_matcher_destroy(ndt2d_matcher * m);

/* replaces ScanMatcherNDT::reset (scan_matcher_ndt.cpp:180-183) */
NDT2D_API int ndt2d_matcher_reset(ndt2d_matcher * m);

/* replaces ScanMatcherNDT::addScans (scan_matcher_ndt.cpp:49-74) and through
 * it NDT::NDT / addScan / compute (ndt_model.cpp:118-160).  Always builds a
 * fresh model.  poses: 3 doubles per scan; pt_offsets: n_scans+1 entries, in
 * points; pts_xy: concatenated sensor-frame points (fewer than 2^30 of them: NDT2D_ERR_SIZE
 * otherwise).  A model of up to 1 MB of points (every rolling window) is built asynchronously:
 * the call returns once one copy from the handle's pinned arena and one launch are enqueued, and
 * the next value-returning call waits for it. */
NDT2D_API int ndt2d_matcher_add_scans(
  ndt2d_matcher * m, size_t n_scans, const double * poses, const uint64_t * pt_offsets,
  const double * pts_xy);

/* replaces ScanMatcherNDT::matchScan (scan_matcher_ndt.cpp:76-149).
 *   out_delta3     written only if some candidate scores < 0 (:128-134);
 *                  *delta_written says whether it was
 *   out_cov9       (1/s) k + (1/s^2) u u^T (:146), always written
 *   *out_score     best_score / n (:148)
 * Without a model: returns NDT2D_ERR_NO_MAP, *out_score = 0.0, nothing else
 * touched (:80). */
NDT2D_API int ndt2d_matcher_match_scan(
  ndt2d_matcher * m, const double * pose3, const double * pts_xy, size_t npts,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score);

/* replaces ScanMatcherNDT::scorePoints (scan_matcher_ndt.cpp:156-178);
 * scoreScan (:151-154) is this with the scan's own pose. */
NDT2D_API int ndt2d_matcher_score_points(
  ndt2d_matcher * m, const double * pts_xy, size_t npts, const double * pose3,
  double * out_score);

/* Batched scorePoints: one score per pose for the same points -- the body of
 * ParticleFilter::measure's loop (particle_filter.cpp:81-87) in one launch. */
NDT2D_API int ndt2d_matcher_score_poses(
  ndt2d_matcher * m, const double * pts_xy, size_t npts, const double * poses3, size_t n_poses,
  double * out_scores);

/* replaces NDT::likelihood(const ScanPtr&) (ndt_model.cpp:189-201): all
 * points, positive sign, not normalised. */
NDT2D_API int ndt2d_matcher_likelihood_scan(
  ndt2d_matcher * m, const double * pose3, const double * pts_xy, size_t npts,
  double * out_likelihood);

/* Loop-closure batch (ndt_mapper.cpp:619-671): for each job j build a fresh
 * model from its own scans (reset + addScans, :628-635) and run matchScan
 * (:638-643) of the job's query scan against it, all jobs in one submission.
 * Map scans of all jobs are concatenated: job j owns scans
 * [job_scan_offsets[j], job_scan_offsets[j+1]); query scans likewise through
 * query_pt_offsets.  Outputs are per job, laid out as in match_scan.
 * The handle's model is left empty afterwards. */
NDT2D_API int ndt2d_matcher_match_scan_batch(
  ndt2d_matcher * m, size_t n_jobs,
  const uint64_t * job_scan_offsets, const double * map_poses, const uint64_t * map_pt_offsets,
  const double * map_pts_xy,
  const double * query_poses, const uint64_t * query_pt_offsets, const double * query_pts_xy,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score);

/* replaces the inner loop of Mapper::loopClosureThread (ndt_mapper.cpp:619-671) for ONE
 * new scan, with its sequential semantics: candidates (indices into the graph's scans, in
 * Graph::findNearest order) are taken in order; a candidate whose scan has no points is
 * skipped without counting (:625); candidate i is matched against a fresh model of the
 * scans [i-1 (i if i == 0), i+1 if i < rolling else i) (:628-635); a finite score below
 * typical_response (:645) is accepted and moves the query scan's pose by the correction
 * (:652-655) BEFORE the next candidate is matched; at most search_limit candidates are
 * processed (:671) -- search_limit == 0 means NO limit, as in the reference, whose size_t
 * countdown `--num_scans_to_check == 0` wraps.  A candidate whose window is empty (i == 0 with
 * rolling == 0) has no map: it scores 0.0, like matchScan without a model
 * (scan_matcher_ndt.cpp:80), is not accepted and counts as processed.  Internally all remaining candidates are matched speculatively in one
 * batch and the remainder is re-issued after every acceptance, so the outputs equal the
 * sequential loop's.
 *   scan_poses / scan_pt_offsets / scan_pts_xy : the graph's scans (n_scans)
 *   query_pose3   in: pose of the new scan; out: its pose after all accepted corrections
 *   out_*         one entry per processed candidate (room for min(search_limit,
 *                 n_candidates), n_candidates if search_limit == 0): candidate index, score, accepted flag, scan pose after
 *                 this candidate, covariance of the match
 *   *n_batches    number of batch submissions it took (1 + number of acceptances that
 *                 were not the last processed candidate) */
NDT2D_API int ndt2d_matcher_close_loop(
  ndt2d_matcher * m, size_t n_scans, const double * scan_poses, const uint64_t * scan_pt_offsets,
  const double * scan_pts_xy, const uint64_t * candidates, size_t n_candidates, size_t rolling,
  size_t search_limit, double typical_response, double * query_pose3, const double * query_pts_xy,
  size_t query_npts, uint64_t * out_candidate, double * out_score, int * out_accepted,
  double * out_pose3, double * out_cov9, size_t * n_processed, size_t * n_batches);

/* ---- staged / partial search: device-resident inputs, theta-sliced ----- */

/* Candidate lattice of this handle: the reference's accumulated-double loops
 * (scan_matcher_ndt.cpp:103,117,119) replayed on the host. */
NDT2D_API int ndt2d_matcher_search_shape(
  const ndt2d_matcher * m, uint64_t * n_angular, uint64_t * n_linear);
/* Copies the replayed loop values (dth: n_angular, dlin: n_linear). */
NDT2D_API int ndt2d_matcher_search_values(const ndt2d_matcher * m, double * dth, double * dlin);

/* Uploads one query scan (subsampled as scan_matcher_ndt.cpp:95-96,110) and
 * its per-theta cos/sin (:106-107, host libm) to the device. */
NDT2D_API int ndt2d_matcher_stage_scan(
  ndt2d_matcher * m, const double * pose3, const double * pts_xy, size_t npts);

/* Launches the search over theta indices [theta_begin, theta_end) of the
 * staged scan, asynchronously on the handle's stream.  The 16-double partial
 * record is left on the device: at d_partial if non-NULL (a device pointer,
 * e.g. a slot of an all-gather buffer), else in the handle.
 *   [0] best score (sum, <= 0)      [1] best global candidate index
 *   [2..7] k upper triangle (xx,xy,xt,yy,yt,tt)   [8..10] u   [11] s
 *   [12] candidates evaluated       [13] points used (n)   [14],[15] reserved
 * Global candidate index = (itheta * n_lin + ix) * n_lin + iy, i.e. the
 * reference's loop order, so "lowest index wins" == "first wins" (:128). */
NDT2D_API int ndt2d_matcher_search_staged(
  ndt2d_matcher * m, uint64_t theta_begin, uint64_t theta_end, void * d_partial);

/* The same over theta indices theta_begin, theta_begin + stride, ... (< theta_end).
 * Interleaving the slices of N ranks (begin = rank, stride = N) balances the work:
 * how much of the scan overlaps the map varies smoothly with theta. */
NDT2D_API int ndt2d_matcher_search_staged_strided(
  ndt2d_matcher * m, uint64_t theta_begin, uint64_t theta_end, uint64_t theta_stride,
  void * d_partial);

/* ---- fused cross-GPU exchange (one process per GPU, one node) ------------------
 * Instead of handing the partial records to a collective library, the last kernel of the
 * search itself stores this rank's 128-byte record into a mailbox in every rank's device
 * memory (peer stores over NVLink, mapped with CUDA IPC), waits -- bounded -- for the other
 * ranks' records in its own mailbox and reduces them exactly like ndt2d_combine_partials.
 *   exchange_init     allocates this rank's mailbox; handle64 = its cudaIpcMemHandle_t
 *   exchange_connect  handles = world * 64 bytes, rank-ordered (the caller all-gathers them
 *                     once, e.g. with torch.distributed); maps the peers' mailboxes
 *   search_exchange   search_staged_strided + publish / wait / reduce; seq > 0 must be the
 *                     same on every rank and grow by 1 per search
 *   fetch_result      the combined result (every rank holds the same);
 *                     NDT2D_ERR_STATE if a rank did not arrive within 10 s */
NDT2D_API int ndt2d_matcher_exchange_init(
  ndt2d_matcher * m, uint32_t world, uint32_t rank, unsigned char * handle64);
NDT2D_API int ndt2d_matcher_exchange_connect(ndt2d_matcher * m, const unsigned char * handles);
NDT2D_API int ndt2d_matcher_search_exchange(
  ndt2d_matcher * m, uint64_t theta_begin, uint64_t theta_end, uint64_t theta_stride, uint64_t seq);
NDT2D_API int ndt2d_matcher_fetch_result(
  ndt2d_matcher * m, double * out_delta3, int * delta_written, double * out_cov9,
  double * out_score);

/* Multi-device handle (ndt2d_params.n_devices > 1): info4 = number of devices, 1 if the fused
 * peer-to-peer exchange is in use (0: host combine), matchScans that ran on all devices so far,
 * sequence number of the last exchange. */
NDT2D_API int ndt2d_matcher_group_info(ndt2d_matcher * m, uint64_t * info4);
/* A matchScan of a multi-device handle is spread over the devices when it has at least
 * min_pairs (candidate, scan point) pairs (default 1e10, ~0.4 ms of one B200: smaller searches
 * are latency-bound and stay on devices[0]). */
NDT2D_API int ndt2d_matcher_set_group_threshold(ndt2d_matcher * m, double min_pairs);
/* The large-search kernel can tally the work it did -- (candidate, point) evaluations that
 * reached an occupied cell, (point, region) items -- for ndt2d_matcher_search_stats /
 * _group_search_stats.  The bookkeeping costs about 2 % of the kernel, so it is off by default
 * (the tallies then read 0); on != 0 turns it on for the searches that follow. */
NDT2D_API int ndt2d_matcher_set_tallies(ndt2d_matcher * m, int on);
/* Small searches and small model builds (the per-scan local match: tens of microseconds) skip
 * the CUDA event records that feed search_stats / build_stats kernel times unless on != 0
 * (default off; large searches and builds are always timed). */
NDT2D_API int ndt2d_matcher_set_timing(ndt2d_matcher * m, int on);
/* Per-device duration (ms, CUDA events on each device's own stream) of the search kernels of the
 * last matchScan of a multi-device handle, and its tallies summed over the devices
 * (totals3 = useful evaluations, (point, region) items, 0). */
NDT2D_API int ndt2d_matcher_group_search_stats(
  ndt2d_matcher * m, double * out_ms, size_t cap, uint64_t * totals3);

/* Synchronises the stream and copies the handle-held partial record out. */
NDT2D_API int ndt2d_matcher_fetch_partial(ndt2d_matcher * m, double * partial16);

/* Reduces n partial records (host memory) exactly like one sequential search
 * would: lexicographic min on (score, index), sums of k/u/s, then the
 * covariance formula (:146) and best/n (:148).  Pure arithmetic on n*16
 * doubles; used after the single cross-GPU exchange. */
NDT2D_API int ndt2d_combine_partials(
  const ndt2d_matcher * m, const double * partials, size_t n_partials,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score);

/* The same reduction without a handle (pure host arithmetic, usable on a machine
 * without a device): dth / dlin are the replayed lattices (ndt2d_search_lattice). */
NDT2D_API int ndt2d_combine_partials_host(
  const double * dth, size_t n_ang, const double * dlin, size_t n_lin, const double * partials,
  size_t n_partials, double * out_delta3, int * delta_written, double * out_cov9,
  double * out_score);
/* Replays `for (v = -size; v < size; v += resolution)` (scan_matcher_ndt.cpp:103,117,
 * 119) on the host: *n values; copied to out if out != NULL (cap entries). */
NDT2D_API int ndt2d_search_lattice(double size, double resolution, double * out, size_t cap,
  size_t * n);

/* Same reduction on the device: d_partials points at n records in device
 * memory (e.g. the all-gather output); result fetched to the host. */
NDT2D_API int ndt2d_matcher_combine_device(
  ndt2d_matcher * m, const void * d_partials, size_t n_partials,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score);

/* ---- parity / introspection ------------------------------------------- */

/* info[5] = size_x, size_y, origin_x, origin_y, cell_size (ndt_model.cpp:118-126) */
NDT2D_API int ndt2d_matcher_grid_info(ndt2d_matcher * m, double * info5);
/* Dense dump, 16 doubles per cell: valid, n, mean[2], covariance[4] (row-major),
 * correlation[4], information[4] -- the members of ndt_2d::Cell
 * (ndt_model.hpp:56-64).  out must hold size_x*size_y*16 doubles. */
NDT2D_API int ndt2d_matcher_dump_cells(ndt2d_matcher * m, double * out);
/* Cell index (NDT::getIndex, ndt_model.cpp:203-218; -1 = outside) of every
 * map point in add_scans order, as computed on the device. */
NDT2D_API int ndt2d_matcher_dump_keys(ndt2d_matcher * m, int32_t * out, size_t n_points);
/* Scores of every candidate of the last search, in loop order (needs
 * n_ang*n_lin^2 doubles); the search is re-run with score capture on. */
NDT2D_API int ndt2d_matcher_dump_scores(
  ndt2d_matcher * m, const double * pose3, const double * pts_xy, size_t npts, double * out,
  size_t n_out);
/* Counters: [0] kernels launched by this handle so far, [1] H2D bytes,
 * [2] D2H bytes, [3] valid (n>=5) cells in the current model. */
NDT2D_API int ndt2d_matcher_counters(ndt2d_matcher * m, uint64_t * out4);
/* Work statistics of the last search launch of the production kernel:
 * [0] (candidate, point) evaluations that reached an occupied cell (the Gaussians
 *     the reference evaluates with a non-zero result), [1] (scan point, candidate
 *     region) pairs that passed the dilated-occupancy test, [2] job counter at
 *     exit, [3] duration of the search kernel alone in nanoseconds (CUDA events
 *     recorded around it on the handle's stream). */
NDT2D_API int ndt2d_matcher_search_stats(ndt2d_matcher * m, uint64_t * out4);
/* The last model build: [0] duration of its kernels (K1..K3, H2D copies excluded) in
 * nanoseconds, [1] map points, [2] occupied (n >= 5) cells, [3] cells of the grid. */
NDT2D_API int ndt2d_matcher_build_stats(ndt2d_matcher * m, uint64_t * out4);
/* cudaStream_t the handle runs on. */
NDT2D_API void * ndt2d_matcher_stream(ndt2d_matcher * m);

/* ------------------------------------------------------------------------
 * ParticleFilter measurement update + resampling
 * (include/ndt_2d/particle_filter.hpp:45-115, src/particle_filter.cpp)
 * The particle set lives on the device between calls.
 * ---------------------------------------------------------------------- */

/* replaces ParticleFilter::ParticleFilter (particle_filter.cpp:36-51):
 * min_particles particles at (0,0,0), uniform weights, statistics updated. */
NDT2D_API int ndt2d_filter_create(
  size_t min_particles, size_t max_particles, int device, void * stream, ndt2d_filter ** out);
NDT2D_API int ndt2d_filter_destroy(ndt2d_filter * f);

/* Direct access to the particle set (particles_: 3 doubles each; weights_). */
NDT2D_API int ndt2d_filter_set_particles(
  ndt2d_filter * f, const double * particles3, const double * weights, size_t n);
NDT2D_API int ndt2d_filter_size(ndt2d_filter * f, size_t * n);
NDT2D_API int ndt2d_filter_get_particles(ndt2d_filter * f, double * particles3, double * weights);

/* replaces ParticleFilter::init (particle_filter.cpp:53-69).  The reference
 * draws from std::normal_distribution<float> on a random_device-seeded
 * mt19937; here a counter-based generator seeded with `seed` (statistical
 * parity only). */
NDT2D_API int ndt2d_filter_init(
  ndt2d_filter * f, double x, double y, double theta, double sigma_x, double sigma_y,
  double sigma_theta, uint64_t seed);

/* replaces ParticleFilter::update -> MotionModel::sample
 * (particle_filter.cpp:71-76, motion_model.cpp:45-83); alphas5 = odom_alpha1..5
 * (statistical parity only, see init). */
NDT2D_API int ndt2d_filter_update(
  ndt2d_filter * f, double dx, double dy, double dth, const double * alphas5, uint64_t seed);

/* replaces ParticleFilter::measure (particle_filter.cpp:78-89): one
 * scorePoints per particle against the matcher's model, then updateStatistics.
 * The scan's own pose is ignored, as in the reference (:85). */
NDT2D_API int ndt2d_filter_measure(
  ndt2d_filter * f, ndt2d_matcher * m, const double * pts_xy, size_t npts);

/* replaces ParticleFilter::resample (particle_filter.cpp:91-137) incl. the
 * KD-tree bin count (kd_tree.hpp:97-189) and the trailing updateStatistics.
 * uniforms: the variates std::discrete_distribution would consume, one per
 * draw, at least max_particles of them; NULL = generate on the device from
 * `seed`. */
NDT2D_API int ndt2d_filter_resample(
  ndt2d_filter * f, double kld_err, double kld_z, const double * uniforms, size_t n_uniforms,
  uint64_t seed);

/* replaces getMean / getCovariance (particle_filter.cpp:139-147).  Note the
 * reference accumulates cov(2,2) across calls (`+=`, :216); so does this. */
NDT2D_API int ndt2d_filter_stats(ndt2d_filter * f, double * mean3, double * cov9);
/* Test hook: overwrite the retained covariance (cov_). */
NDT2D_API int ndt2d_filter_set_cov(ndt2d_filter * f, const double * cov9);
/* Indices drawn by the last resample (out must hold size() entries). */
NDT2D_API int ndt2d_filter_last_draws(ndt2d_filter * f, uint64_t * out);

/* ------------------------------------------------------------------------
 * The step before the path
 * ---------------------------------------------------------------------- */

/* replaces the LaserScan -> ndt_2d::Scan conversion of Mapper::laserCallback
 * (ndt_mapper.cpp:385-453): beams that are NaN or beyond range_max are dropped, the rest
 * are projected in the laser frame (angle = angle_min + i * angle_increment, float
 * arithmetic as in the reference), moved to the robot frame by laser_tf3 = {x, y, theta}
 * and de-skewed with the motion during the scan: translation3 = end-of-scan odometry
 * pose minus start-of-scan pose, spread linearly over the n beams.  laser_inverted
 * walks the beams backwards with negated angles (index 0 is never visited, :411).
 * out_pts_xy needs room for n points; *n_out = points kept, in the reference's order.
 * Which beams are kept is exact; coordinates agree with the reference to ~1e-15
 * relative (device cos / sin). */
NDT2D_API int ndt2d_laser_to_points(
  int device, const float * ranges, size_t n, float angle_min, float angle_increment,
  double range_max, const double * laser_tf3, const double * translation3, int laser_inverted,
  double * out_pts_xy, size_t * n_out);

/* replaces Graph::findNearest (graph.cpp:167-189), the candidate selection of the loop
 * closure (ndt_mapper.cpp:612-616): the indices of the scans whose position lies within
 * the radius of the query position, nearest first.  The reference builds a nanoflann
 * KD-tree (un-vendored dependency, L2_Simple_Adaptor, 2-D) over the first `limit` scans
 * on every call and runs radiusSearch, which takes a SQUARED radius (so the node's
 * global_search_size 0.2 means sqrt(0.2) m), keeps dist < radius (strict) and returns the
 * matches by ascending distance.  Here: one thread per scan, distance accumulated like
 * evalMetric (dx*dx then + dy*dy), order (distance, index) -- nanoflann's order among
 * exactly equal distances is unspecified.
 *   scan_xy            2 * n_scans doubles: Scan::getPose() (or getBarycenterPose(),
 *                      graph.cpp:173,178) x, y of every scan of the graph
 *   limit_scan_index   > 0: only scans [0, limit) are searched; <= 0: all (:171)
 *   out_indices / out_dist_sq (optional)  room for `capacity` entries
 *   *n_found           number of matches (may exceed capacity; the nearest are written) */
NDT2D_API int ndt2d_find_nearest(
  int device, const double * scan_xy, size_t n_scans, int64_t limit_scan_index,
  const double * query_xy, double radius_sq, uint64_t * out_indices, double * out_dist_sq,
  size_t capacity, size_t * n_found);

/* ------------------------------------------------------------------------
 * Occupancy-grid export: ndt_2d::OccupancyGrid
 * (include/ndt_2d/occupancy_grid.hpp:40-68, src/occupancy_grid.cpp)
 * ---------------------------------------------------------------------- */
typedef struct ndt2d_occupancy ndt2d_occupancy;

/* replaces OccupancyGrid::OccupancyGrid(resolution, occ_thresh) (occupancy_grid.cpp:35-45) */
NDT2D_API int ndt2d_occupancy_create(
  double resolution, double occ_thresh, int device, ndt2d_occupancy ** out);
NDT2D_API int ndt2d_occupancy_destroy(ndt2d_occupancy * g);
/* replaces OccupancyGrid::getMsg (:47-152) incl. updateBounds (:155-185): the bounds
 * persist across calls and grow with the scans not seen before (only when the number of
 * scans changed, as in the reference); every scan is ray-traced into hit / empty counters;
 * info5 = {width, height, origin_x, origin_y, resolution} (grid.info).  The cell data
 * stays on the device until fetched. */
NDT2D_API int ndt2d_occupancy_render(
  ndt2d_occupancy * g, size_t n_scans, const double * poses, const uint64_t * pt_offsets,
  const double * pts_xy, double * info5);
/* grid.data of the last render: width * height int8 (-1 unknown, 0 free, 100 occupied),
 * row-major, index = x + y * width (:112). */
NDT2D_API int ndt2d_occupancy_fetch(ndt2d_occupancy * g, int8_t * data, size_t capacity);

/* ------------------------------------------------------------------------
 * Roofline probes (bench.py): measured on the device the bench runs on.
 * ---------------------------------------------------------------------- */

/* GB/s of random 32-byte record reads from a table of table_bytes (L1-, L2- or
 * HBM-resident depending on its size): the gather roofline of SURVEY.md 8(d). */
NDT2D_API int ndt2d_probe_gather(int device, size_t table_bytes, double * out_gbps);
/* Gaussian evaluations per second of the search kernel's own evaluation recipe run back to
 * back with no bookkeeping (2 packed FMAs + one ex2 on the SFU + one packed add per
 * evaluation, register accumulators, the search kernel's launch shape): the measured
 * arithmetic bound of the correlative search on this device (bench.py roofline.peak). */
NDT2D_API int ndt2d_probe_ex2(int device, double * out_evals_per_s);
/* GB/s (read + write) of a plain device-to-device copy of `bytes`. */
NDT2D_API int ndt2d_probe_copy(int device, size_t bytes, double * out_gbps);
/* Self-check of the build's divide-by-point-count (csrc/build_common.cuh: a correctly rounded
 * quotient from the count's reciprocal, 3 dependent operations instead of the IEEE divide's ~10
 * on the critical path of Cell::addPoint's recurrence, ndt_model.cpp:50-63): `trials`
 * pseudo-random (numerator, count <= 2^20) pairs against __ddiv_rn; *out_mismatches must be 0. */
NDT2D_API int ndt2d_probe_div_by_count(int device, uint64_t seed, uint64_t trials,
  uint64_t * out_mismatches);
/* Latency of the node's per-scan calls as a C / C++ caller sees them (no binding overhead):
 * `calls` repetitions, each timed with the host's steady clock, microseconds into out_us[calls].
 *   what = 0  ndt2d_matcher_match_scan(pose3, pts)                      (ndt_mapper.cpp:513)
 *   what = 1  reset + add_scans + score_points + match_scan             (ndt_mapper.cpp:508-515)
 * The map arguments are only read for what = 1. */
NDT2D_API int ndt2d_probe_call_latency(
  ndt2d_matcher * m, int what, size_t n_scans, const double * map_poses,
  const uint64_t * map_pt_offsets, const double * map_pts_xy, const double * pose3,
  const double * pts_xy, size_t npts, size_t calls, double * out_us);

#ifdef __cplusplus
}
#endif

#endif  /* NDT2D_B200_H_ */
