// ndt2d_internal.h -- shared declarations of the libndt2d_b200 translation units.
#ifndef NDT2D_INTERNAL_H_
#define NDT2D_INTERNAL_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include "ndt2d_b200.h"

// Description of one NDT grid, passed by value to kernels.
//
// The reference's grid (ndt_model.cpp:118-126) is size_x * size_y cells with
// origin (origin_x, origin_y).  On the device every lookup goes through a
// PADDED grid of (size_x + 2) * (size_y + 2) cells: padded coordinate
// ex = 0 means "x < origin_x" and ex = size_x + 1 means "grid_x >= size_x",
// the two ways NDT::getIndex (ndt_model.cpp:203-218) returns -1.  Border cells
// are never occupied, so out-of-map points need no special case.
struct GridDesc
{
  double origin_x, origin_y, cell_size;
  uint32_t size_x, size_y;  // reference cells per axis
  uint32_t pitch;           // size_x + 2
  uint32_t n_cells;         // size_x * size_y
  uint32_t n_padded;        // (size_x + 2) * (size_y + 2)
  uint32_t n_words;         // ceil(n_padded / 32)
  // search parameters the build needs (they are fixed per matcher handle):
  uint32_t dil_x;           // the dilated bitmap ORs cell columns c .. c + dil_x (1 or 2), rows c, c + 1
  double lin_res;           // search_linear_resolution (row step of the vertex-form records)
};

// Device-resident model of one matcher.
//   occ[w]   .x = occupancy bits of padded cells 32w..32w+31 (cell has n >= 5,
//                 the only cells Cell::score does not return 0 for,
//                 ndt_model.cpp:107-111)
//            .y = number of occupied cells before word w  (rank prefix)
//   rec[r*6] = mean_x, mean_y, -0.5*I(0,0), -0.5*I(1,0), -0.5*I(0,1), -0.5*I(1,1)
//              of the r-th occupied cell: what Cell::score (ndt_model.cpp:113-115)
//              needs, 48 bytes
//   rec_fast[r*6] = mean_x, mean_y, A, B, D, stiff  with
//              A = -0.5 log2(e) I(0,0), B = -0.5 log2(e) (I(0,1) + I(1,0)),
//              D = -0.5 log2(e) I(1,1): log2 of the likelihood is
//              qx (A qx + B qy) + (D qy) qy.  stiff != 0 marks cells whose
//              information matrix is too ill-conditioned for that form to stay
//              within 1e-7 of the reference's grouping (max|I| cell^2 > 1e7, inf
//              or NaN); the search then evaluates them from rec[] with the
//              reference's own operation order.
//   rec_vtx[r*6] = mean_x, mean_y, D, Bh = B / (2 D), S = A - B^2 / (4 D), {float c2 = D h^2,
//              float stiff}: the same quadratic in "vertex form" along a column of candidates
//              (x fixed, y = y0 + b h):  log2 L = D (qy + Bh qx)^2 + S qx^2, which the region
//              kernel evaluates per row b as c2 b'^2 + d1 b' + e0 in float with b' = b - b*
//              counted from the row nearest the vertex (no cancellation between the terms).
//              stiff additionally covers D >= 0 (the form needs D < 0).
//   thr_x[k] = smallest double x whose reference grid_x is >= k  (k=0: origin)
struct ModelView
{
  GridDesc g;
  const uint2 * occ;
  const uint32_t * occ_dilated;  // D[c] = OR of E over columns c .. c + g.dil_x, rows c, c + pitch
  const double * rec;
  const double * rec_fast;  // 6 doubles per occupied cell, see above
  const double * rec_vtx;   // 6 doubles per occupied cell, see above
  const double * thr_x;  // size_x + 2 entries (the last is +inf)
  const double * thr_y;  // size_y + 2 entries
  uint32_t n_valid_cap;
  const uint32_t * n_stiff;  // number of cells flagged stiff in rec_vtx (null: unknown, assume some)
};

#define NDT2D_REC_DOUBLES 6

// One staged query scan + candidate lattice.
struct SearchView
{
  const double2 * pts;   // n subsampled sensor-frame points
  const double2 * trig;  // n_ang (cos, sin) of pose.theta + dth
  const double * dth;    // n_ang
  const double * dlin;   // n_lin
  double pose_x, pose_y;
  double linear_res;     // search_linear_resolution (region sizing, row step of the region kernel)
  double inv_linear_res; // 1 / linear_res
  uint32_t n_pts, n_ang, n_lin;
  uint32_t * coords;        // scratch of the coordinate pre-pass (may be null)
  size_t coords_cap_bytes;
  double * chunk_sums;      // scratch of the point-chunked mode of small searches (may be null)
  size_t chunk_cap_doubles;
  uint32_t theta_stride;  // a search covers theta_begin, theta_begin + stride, ... (< theta_end)
  uint32_t tally;         // != 0: the region kernel tallies useful evaluations / items (search_stats)
};

// Per-block partial of the search (8 doubles):
//   [0] best score  [1] best global index (as double; < 2^53)
//   [2] S  [3] Sx  [4] Sy  [5] Sxx  [6] Sxy  [7] Syy     (sums over the block's
//   candidates of score * {1, dx, dy, dx^2, dx*dy, dy^2}); [8] = dth of the block
#define NDT2D_BLOCK_PARTIAL 9

struct Counters
{
  uint64_t launches, h2d_bytes, d2h_bytes;
};

// ---- build.cu -----------------------------------------------------------
#define NDT2D_RCP_TABLE 8192
struct BuildScratch
{
  // all device pointers, capacities in elements
  double * wx, * wy;        // world coordinates per map point
  double * sx, * sy;        // the same in sorted (cell) order
  uint2 * heads;            // (sorted position, length) of every cell with n >= 5
  uint32_t * n_heads;
  uint32_t * key[2];        // ping-pong sort keys (cell index, n_cells = outside)
  uint32_t * val[2];        // ping-pong values (point index)
  uint32_t * seglen;        // per sorted position: segment length (heads only)
  uint32_t * hist;          // sort scratch: digit totals, tickets, look-back status (ndt2d_sort_scratch_bytes)
  uint32_t * scan_tmp;      // scratch of the scan kernel
  double * rcp;             // rcp[0] = 1.0, rcp[k] = RN(1 / k), k <= NDT2D_RCP_TABLE (filled by the first build)
  bool rcp_ready;
  size_t cap_points, cap_hist;
};

// One small model built by a single CTA (build.cu: build_small_kernel / _batch_kernel).
struct BuildEntry
{
  GridDesc g;
  const double4 * scan_tf;
  const uint64_t * offsets;
  const double2 * pts;
  uint2 * occ;
  uint32_t * occd;
  double * rec;
  double * rec_fast;
  double * rec_vtx;
  uint32_t * n_valid;
  // parity-dump outputs (sorted keys / values / run lengths / coordinates); null in batches
  uint32_t * key_out, * val_out, * seglen;
  double * sx, * sy;
  uint32_t n_scans, n_points, rec_cap, pad_;
};
bool ndt2d_build_is_small(const GridDesc & g, size_t n_points);
size_t ndt2d_sort_scratch_bytes(size_t n_points);
// n models, one CTA each, one launch; d_entries: device array.
int ndt2d_launch_build_small_batch(const BuildEntry * d_entries, uint32_t n, cudaStream_t stream,
  Counters * ctr);

// Launches K1..K3 on `stream`: transform + key, stable radix sort by cell key,
// per-cell sequential moments, occupancy bitmap + rank prefix, packed records.
// d_scan_tf: per scan {x, y, cos, sin}; d_offsets: n_scans + 1 point offsets;
// d_pts: sensor-frame points.  Returns the index (0/1) of the sorted buffers.
int ndt2d_launch_build(
  const GridDesc & g, const double4 * d_scan_tf, const uint64_t * d_offsets, size_t n_scans,
  const double2 * d_pts, size_t n_points, BuildScratch & s, uint2 * d_occ, uint32_t * d_occ_dilated,
  double * d_rec, double * d_rec_fast, double * d_rec_vtx, uint32_t rec_cap, uint32_t * d_n_valid,
  cudaStream_t stream, Counters * ctr, int * sorted_buf);

// Debug/parity: dense dump (16 doubles per reference cell) from the sorted
// buffers of the last build.
int ndt2d_launch_dump_cells(
  const GridDesc & g, const BuildScratch & s, int sorted_buf, size_t n_points, double * d_out,
  cudaStream_t stream, Counters * ctr);

// Generic exclusive scan of n uint32 (in place), single launch sequence.
int ndt2d_launch_exclusive_scan(uint32_t * d_data, size_t n, uint32_t * d_tmp,
  cudaStream_t stream, Counters * ctr);

// Cross-GPU exchange fused into the last kernel of a search (search.cu:
// search_finish_kernel).  Every rank owns a MAILBOX in its device memory, mapped
// into the other ranks' address spaces through CUDA IPC:
//   records  [2 parities][kExchangeMaxRanks][16] doubles   slot [p][r]: rank r's partial record
//   flags    [2 parities][kExchangeMaxRanks]     u64       slot [p][r]: sequence number it belongs to
// A rank finishes its slices, stores its 16-double record into slot [seq & 1][rank] of
// EVERY mailbox (peer stores over NVLink), fences system-wide, then stores seq into the
// matching flags (release); it then waits (acquire, bounded) until its own mailbox holds
// all ranks' flags >= seq and reduces the records like ndt2d_launch_combine.  Two
// parities are enough: a rank can only publish seq + 2 after every rank published seq + 1,
// i.e. after every rank has finished reading seq.
constexpr uint32_t kExchangeMaxRanks = 64;
constexpr size_t kExchangeRecordBytes = 2 * kExchangeMaxRanks * 16 * sizeof(double);
constexpr size_t kExchangeMailboxBytes = kExchangeRecordBytes + 2 * kExchangeMaxRanks * 8;
struct ExchangeView
{
  void * const * peers;          // device array: base of every rank's mailbox (this rank's own included)
  uint32_t world, rank;
  unsigned long long seq;        // > 0, the same on every rank, +1 per search
  unsigned long long timeout_ns; // bound of the wait (a rank that never arrives must not hang the GPU)
};

// Result mailbox in MAPPED PINNED HOST memory: the last kernel of a search stores the
// 32-double result record there itself (zero-copy store over PCIe), fences system-wide and
// releases `seq` into the flag; the host polls the flag instead of paying for a device-to-host
// copy plus a stream synchronisation (the local match of every scan is latency-bound).
struct HostMailbox
{
  double * out32;                // device-visible address of the pinned record (null: unused)
  unsigned long long * flag;     // device-visible address of the pinned sequence flag
  unsigned long long seq;        // value released when the record is complete
};

// ---- search.cu ----------------------------------------------------------
// Search theta slices [theta_begin, theta_end); writes the 32-double record
// (partial + finished outputs, see ndt2d_launch_combine) to d_partial32.  d_block_partials: scratch, >= capacity returned by
// ndt2d_search_scratch_doubles().  d_scores (optional): per-candidate scores.
size_t ndt2d_search_scratch_doubles(uint32_t n_ang, uint32_t n_lin, double cell_size,
  double linear_res);

// search_region.cu: the production kernel (variant 0): one warp per
// (theta, region of candidates) job, jobs handed out through *d_counter.
size_t ndt2d_region_scratch_doubles(double cell_size, uint32_t n_ang, uint32_t n_lin,
  double linear_res);
// Extra cell columns the dilated bitmap has to cover for this lattice (1 or 2): a region is
// up to 32 candidate columns wide.
uint32_t ndt2d_region_dilate_x(double cell_size, double linear_res, uint32_t n_lin);
// Bytes of the coordinate pre-pass table for a search of this shape (0 if above cap).
size_t ndt2d_region_coords_bytes(double cell_size, uint32_t n_ang, uint32_t n_lin,
  double linear_res, uint32_t n_pts, size_t cap_bytes);
size_t ndt2d_region_chunk_doubles(double cell_size, uint32_t n_ang, uint32_t n_lin,
  double linear_res, uint32_t n_pts);
int ndt2d_launch_search_region(
  const ModelView & mv, const SearchView & sv, double linear_res, uint32_t theta_begin,
  uint32_t n_theta, double * d_job_partials, double * d_scores, uint32_t * d_counter,
  uint32_t * d_coords, size_t coords_cap_bytes, cudaStream_t stream, Counters * ctr,
  uint32_t * n_jobs);
int ndt2d_launch_search(
  const ModelView & mv, const SearchView & sv, uint32_t theta_begin, uint32_t theta_end,
  int variant, double * d_block_partials, double * d_partial32, double * d_scores,
  uint32_t * d_counter, cudaStream_t stream, Counters * ctr, cudaEvent_t ev_begin = nullptr,
  cudaEvent_t ev_end = nullptr,   // events (optional) bracket the search kernel alone
  const ExchangeView * exchange = nullptr,   // non-null: fused cross-GPU exchange + combine
  const HostMailbox * host = nullptr);       // non-null: result record also stored to the host

// Combine n 16-double partial records (device) into one 32-double record
// (device): [0..15] partial, [16..18] delta, [19] delta_written,
// [20..28] covariance, [29] best / n.
int ndt2d_launch_combine(const double * d_partials, uint32_t n, const double * d_dth,
  const double * d_dlin, uint32_t n_lin, double * d_out32, cudaStream_t stream, Counters * ctr);

// Batched pose scoring: d_pose_tf[p] = {x, y, cos, sin}; out[p] = scorePoints.
// sign = -1 / normalise = 1 reproduces scorePoints (scan_matcher_ndt.cpp:156-178);
// sign = +1 / normalise = 0 reproduces NDT::likelihood(ScanPtr) (ndt_model.cpp:189-201).
int ndt2d_launch_score_poses(
  const ModelView & mv, const double2 * d_pts, uint32_t n_pts, const double4 * d_pose_tf,
  uint32_t n_poses, double sign, int normalise, double * d_out, cudaStream_t stream,
  Counters * ctr, const HostMailbox * host = nullptr);  // host: n_poses <= 8 only (one block)

// ---- batched searches (match_scan_batch): several models / scans, one launch each for
// build, search and final reduction ------------------------------------------------------
// One search of a batch: its own model, scan and outputs.
struct BatchEntry
{
  ModelView mv;
  SearchView sv;
  double * job_partials;   // n_jobs records of NDT2D_BLOCK_PARTIAL doubles
  double * chunk_sums;     // n_jobs * P * Rw^2 doubles (P > 1)
};
struct RegionBatchPlan
{
  uint32_t RX, RY, Qx, Qy, n_jobs, P, chunk_points;
  size_t chunk_doubles;    // per search
};
// Plan shared by every search of a batch (same lattice); max_pts = most points of any scan.
int ndt2d_region_batch_plan(double cell_size, uint32_t n_ang, uint32_t n_lin, double linear_res,
  uint32_t max_pts, uint32_t n_searches, RegionBatchPlan * out);
// Search kernel (+ chunk reduction) over all entries; job records land in entry.job_partials.
int ndt2d_launch_search_region_batch(
  const BatchEntry * d_batch, uint32_t n_batch, const RegionBatchPlan & pl, uint32_t * d_counter,
  cudaStream_t stream, Counters * ctr);
// ---- search_window.cu: small windows (a few cells wide), thread per candidate with the
// per-point work shared by a CTA.  ndt2d_window_cells() = cells per axis the window can touch
// (2..4), 0 = not eligible; records as the dense kernel's, ndt2d_window_records() per search.
uint32_t ndt2d_window_cells(double cell_size, uint32_t n_lin, double linear_res);
uint32_t ndt2d_window_records(uint32_t n_theta, uint32_t n_lin);
int ndt2d_launch_search_window(
  const ModelView & mv, const SearchView & sv, uint32_t K, uint32_t theta_begin, uint32_t n_theta,
  double * d_block_partials, double * d_scores, cudaStream_t stream, Counters * ctr);
int ndt2d_launch_search_window_batch(
  const BatchEntry * d_batch, uint32_t n_batch, uint32_t K, uint32_t n_ang, uint32_t n_lin,
  cudaStream_t stream, Counters * ctr);
// Coarse lattices: the dense (warp per candidate) kernel over all entries instead;
// ndt2d_dense_batch_records() records of NDT2D_BLOCK_PARTIAL doubles land in entry.job_partials.
uint32_t ndt2d_dense_batch_records(uint32_t n_ang, uint32_t n_lin);
int ndt2d_launch_search_dense_batch(
  const BatchEntry * d_batch, uint32_t n_batch, uint32_t n_ang, uint32_t n_lin, cudaStream_t stream,
  Counters * ctr);
// One block per entry folds its job records (<= 4096) into results32 + 32 * entry and
// finishes it; clears the job counter block for the next launch.
int ndt2d_launch_finish_batch(
  const BatchEntry * d_batch, uint32_t n_batch, uint32_t n_jobs, double n_candidates,
  double * d_results32, uint32_t * d_counter, cudaStream_t stream, Counters * ctr);

// ---- filter.cu ----------------------------------------------------------
struct FilterView
{
  double * particles;  // 3 * cap
  double * weights;    // cap
  double * stats;      // [0..2] mean, [3..11] cov (row-major), persistent
};
int ndt2d_launch_pose_tf(const double * d_particles, uint32_t n, double4 * d_pose_tf,
  cudaStream_t stream, Counters * ctr);
int ndt2d_launch_filter_stats(FilterView f, uint32_t n, cudaStream_t stream, Counters * ctr);
int ndt2d_launch_filter_resample(
  FilterView f, uint32_t n, uint32_t min_particles, uint32_t max_particles, double kld_err,
  double kld_z, const double * d_uniforms, uint64_t seed, double * d_cdf, double * d_new_particles,
  double * d_new_weights, uint32_t * d_draws, uint32_t * d_first, uint32_t * d_canon,
  uint32_t * d_table, uint32_t table_size, uint32_t * d_new_n, cudaStream_t stream,
  Counters * ctr);
int ndt2d_launch_filter_init(FilterView f, uint32_t n, double x, double y, double th, double sx,
  double sy, double sth, uint64_t seed, cudaStream_t stream, Counters * ctr);
int ndt2d_launch_filter_motion(FilterView f, uint32_t n, double rot1, double trans, double rot2,
  double s_rot1, double s_trans, double s_rot2, uint64_t seed, cudaStream_t stream,
  Counters * ctr);

// ---- error plumbing -----------------------------------------------------
void ndt2d_set_error(const char * what, cudaError_t e, const char * file, int line);

#define NDT2D_CUDA_TRY(expr)                                        \
  do {                                                              \
    cudaError_t e__ = (expr);                                       \
    if (e__ != cudaSuccess) {                                       \
      ndt2d_set_error(#expr, e__, __FILE__, __LINE__);              \
      return NDT2D_ERR_CUDA;                                        \
    }                                                               \
  } while (0)

#define NDT2D_LAUNCH_CHECK(ctr)                                     \
  do {                                                              \
    cudaError_t e__ = cudaGetLastError();                           \
    if (e__ != cudaSuccess) {                                       \
      ndt2d_set_error("kernel launch", e__, __FILE__, __LINE__);    \
      return NDT2D_ERR_CUDA;                                        \
    }                                                               \
    if (ctr) {(ctr)->launches += 1;}                                \
  } while (0)

#endif  // NDT2D_INTERNAL_H_
