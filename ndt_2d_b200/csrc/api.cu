// api.cu -- the C ABI of libndt2d_b200.so (include/ndt2d_b200.h): handles,
// host-side replay of the reference's loop bounds and trigonometry, staging,
// and the launch sequences of build.cu / search.cu / filter.cu.
//
// Host work kept deliberately (it is what makes cell indices bit-exact):
//   * the (dth, dx, dy) lattices are replayed with the reference's accumulating
//     double loops (scan_matcher_ndt.cpp:103,117,119), never as -size + i*res;
//   * cos/sin of scan poses and of pose.theta + dth come from the host libm,
//     the same library the reference calls (ndt_model.cpp:135-136,
//     scan_matcher_ndt.cpp:106-107);
//   * the scan subsampling index size_t(i * step) (:95-96,110).
// Everything per point / per candidate runs on the device.
#include <cuda_runtime.h>

#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <limits>
#include <mutex>
#include <thread>
#include <new>
#include <vector>

#include "ndt2d_internal.h"

// ---------------------------------------------------------------- errors
static thread_local char g_last_error[512] = "";

void ndt2d_set_error(const char * what, cudaError_t e, const char * file, int line)
{
  snprintf(g_last_error, sizeof(g_last_error), "%s:%d: %s -> %s (%s)", file, line, what,
    cudaGetErrorName(e), cudaGetErrorString(e));
}

namespace
{

struct DeviceBuffer
{
  void * p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes)
  {
    if (bytes <= cap) {return NDT2D_OK;}
    if (p) {
      cudaFree(p);
      p = nullptr;
      cap = 0;
    }
    // grow geometrically so rolling-window rebuilds do not re-allocate
    size_t want = bytes + bytes / 4 + 256;
    NDT2D_CUDA_TRY(cudaMalloc(&p, want));
    cap = want;
    return NDT2D_OK;
  }
  void release()
  {
    if (p) {cudaFree(p);}
    p = nullptr;
    cap = 0;
  }
  template<typename T> T * as() const {return static_cast<T *>(p);}
};

struct PinnedBuffer
{
  void * p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes)
  {
    if (bytes <= cap) {return NDT2D_OK;}
    if (p) {
      cudaFreeHost(p);
      p = nullptr;
      cap = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    NDT2D_CUDA_TRY(cudaMallocHost(&p, want));
    cap = want;
    return NDT2D_OK;
  }
  void release()
  {
    if (p) {cudaFreeHost(p);}
    p = nullptr;
    cap = 0;
  }
  template<typename T> T * as() const {return static_cast<T *>(p);}
};

struct DeviceGuard
{
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev)
  {
    if (cudaGetDevice(&prev) != cudaSuccess) {prev = -1;}
    if (prev != dev) {ok = cudaSetDevice(dev) == cudaSuccess;}
  }
  ~DeviceGuard()
  {
    if (prev >= 0) {cudaSetDevice(prev);}
  }
};

// Replays `for (v = -size; v < size; v += res)`.
int replay_loop(double size, double res, size_t limit, std::vector<double> & out)
{
  out.clear();
  if (!(res > 0.0) || !std::isfinite(res) || !std::isfinite(size)) {return NDT2D_ERR_INVALID;}
  if (size > 0.0 && (2.0 * size / res) > static_cast<double>(limit)) {return NDT2D_ERR_SIZE;}
  for (double v = -size; v < size; v += res) {
    out.push_back(v);
    if (out.size() > limit) {return NDT2D_ERR_SIZE;}
  }
  return NDT2D_OK;
}

// The reference's grid coordinate on one axis (NDT::getIndex,
// ndt_model.cpp:205-215) as a monotone function of v:
//   -1 below the origin, else min(unsigned((v - origin) / cell), size).
inline int64_t ref_coord(double v, double origin, double cell, uint32_t size)
{
  if (v < origin) {return -1;}
  const double q = (v - origin) / cell;
  if (!(q < static_cast<double>(size))) {return size;}
  return static_cast<int64_t>(static_cast<unsigned int>(q));
}

// thr[k], k = 0..size: the smallest double v with ref_coord(v) >= k; thr[size + 1] = +inf.
// ref_coord is monotone non-decreasing in v (subtraction, division by a
// positive constant and truncation all are), so the thresholds partition the
// axis exactly like the reference's own arithmetic does.
void axis_thresholds(double origin, double cell, uint32_t size, std::vector<double> & thr)
{
  thr.resize(static_cast<size_t>(size) + 2);
  thr[size + 1] = std::numeric_limits<double>::infinity();
  thr[0] = origin;
  for (uint32_t k = 1; k <= size; ++k) {
    double v = origin + static_cast<double>(k) * cell;
    int steps = 0;
    // walk down while still >= k, then up until >= k
    while (ref_coord(v, origin, cell, size) >= static_cast<int64_t>(k) && steps < 4096) {
      v = std::nextafter(v, -std::numeric_limits<double>::infinity());
      ++steps;
    }
    while (ref_coord(v, origin, cell, size) < static_cast<int64_t>(k) && steps < 8192) {
      v = std::nextafter(v, std::numeric_limits<double>::infinity());
      ++steps;
    }
    thr[k] = v;
  }
}

}  // namespace

// ---------------------------------------------------------------- matcher
constexpr size_t kBatchLanes = 4;

// Host thread bound to one device of a multi-device handle: the per-device host work of a
// search (staging, the libm cos / sin of the device's theta slices, the launches) runs on all
// devices at once instead of one after the other.
struct GroupWorker
{
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> task;
  bool has_task = false, done = false, quit = false;
  // hints for the short polling phases (the mutex-protected flags above stay the truth): a worker
  // that has just finished a task polls `pending` for a while before it sleeps, so a burst of
  // searches -- the loop-closure thread matching one candidate after the other -- finds it awake,
  // and the dispatching thread polls `finished` instead of sleeping on the condition variable
  std::atomic<bool> pending{false}, finished{false};
  int rc = 0;
  char err[512] = "";
};

struct ndt2d_matcher
{
  std::mutex mu;
  ndt2d_params prm;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  Counters ctr{0, 0, 0};

  size_t max_beams = 0;             // laser_max_beams as size_t (hpp:99)
  std::vector<double> dth, dlin;    // replayed loops
  DeviceBuffer d_dth, d_dlin;

  bool has_model = false;
  GridDesc g{};
  DeviceBuffer d_occ, d_occd, d_rec, d_rec_fast, d_rec_vtx, d_nvalid;
  double * d_thr = nullptr;          // thr_x, thr_y: the head of d_build_in
  uint32_t rec_cap = 0;
  uint32_t n_valid = 0;

  BuildScratch bs{};
  DeviceBuffer d_sx, d_sy, d_heads, d_nheads, d_wx, d_wy, d_key0, d_key1, d_val0, d_val1, d_seglen, d_hist, d_scantmp;
  DeviceBuffer d_build_in;           // [thresholds | scan transforms | offsets | map points]
  DeviceBuffer d_rcp;                // reciprocals of the point counts (BuildScratch::rcp)
  size_t n_map_points = 0;
  int sorted_buf = 0;

  bool staged = false;
  DeviceBuffer d_pts, d_trig, d_blockpart, d_partial, d_pose_tf, d_out, d_counter, d_coords, d_chunk;
  double pose_x = 0, pose_y = 0;
  uint32_t n_pts = 0;
  size_t trig_offset_bytes = 0;     // (cos, sin) table inside d_pts, after the points
  // staged API: the (cos, sin) of a theta slice is computed and uploaded when a search first
  // asks for it (a rank of a sharded search only ever needs its own slices)
  bool trig_lazy = false;
  double pose_th = 0.0;
  std::vector<double> h_trig_all;   // host mirror, 2 * n_ang
  std::vector<uint8_t> trig_done;

  PinnedBuffer h_stage, h_result;
  // Pipelined mode (match_scan_batch): host staging comes from a pinned arena that is
  // only recycled after a stream synchronisation, and no call waits for the device.
  // fused cross-GPU exchange (ndt2d_matcher_exchange_*): this rank's mailbox, the peers'
  // mailboxes mapped through CUDA IPC, and the device table of all of them
  DeviceBuffer d_mailbox, d_peer_table;
  std::vector<void *> peer_ptrs;        // index = rank; own rank = d_mailbox.p
  uint32_t x_world = 0, x_rank = 0;
  bool x_connected = false;
  bool x_local = false;                 // peers are raw pointers of this process (no IPC handles)
  // single-process multi-GPU (ndt2d_params.n_devices > 1): this handle (rank 0, devices[0]) owns
  // one sub-handle per further device; group[r] = the handle of rank r (group[0] == this)
  std::vector<ndt2d_matcher *> group;
  GroupWorker * worker = nullptr;       // of a sub-handle (rank >= 1)
  bool group_p2p = false;               // mailboxes mapped: fused exchange; else host combine
  unsigned long long group_seq = 0;
  unsigned long long group_searches = 0;  // matchScans that ran on all devices
  double group_min_pairs = 1.0e10;        // smallest search spread over the devices
  bool pipelined = false;
  std::vector<ndt2d_matcher *> lanes;   // sub-handles of match_scan_batch (created on first use)
  PinnedBuffer h_arena;
  size_t arena_off = 0;
  DeviceBuffer d_batch_results, d_batch_arena;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;  // bracket the last search kernel
  bool ev_valid = false;
  // small searches / builds skip the event records (two driver calls and a bubble between the
  // kernels each) unless asked for: ndt2d_matcher_set_timing
  bool time_small = false;
  bool tally = false;                // region kernel tallies (ndt2d_matcher_set_tallies)
  unsigned long long host_seq = 0;   // HostMailbox sequence of the last search
  cudaEvent_t evb_begin = nullptr, evb_end = nullptr;  // bracket the kernels of the last build
  bool evb_valid = false;
};

namespace
{

ModelView model_view(const ndt2d_matcher * m)
{
  ModelView mv;
  mv.g = m->g;
  mv.occ = m->d_occ.as<uint2>();
  mv.occ_dilated = m->d_occd.as<uint32_t>();
  mv.rec = m->d_rec.as<double>();
  mv.rec_fast = m->d_rec_fast.as<double>();
  mv.rec_vtx = m->d_rec_vtx.as<double>();
  mv.thr_x = m->d_thr;
  mv.thr_y = m->d_thr + (m->g.size_x + 2);
  mv.n_valid_cap = m->rec_cap;
  mv.n_stiff = m->d_nvalid.as<uint32_t>() + 1;
  return mv;
}

SearchView search_view(const ndt2d_matcher * m)
{
  SearchView sv;
  sv.pts = m->d_pts.as<double2>();
  sv.trig = reinterpret_cast<const double2 *>(m->d_pts.as<char>() + m->trig_offset_bytes);
  sv.dth = m->d_dth.as<double>();
  sv.dlin = m->d_dlin.as<double>();
  sv.pose_x = m->pose_x;
  sv.pose_y = m->pose_y;
  sv.linear_res = m->prm.search_linear_resolution;
  sv.inv_linear_res = 1.0 / sv.linear_res;
  sv.n_pts = m->n_pts;
  sv.n_ang = static_cast<uint32_t>(m->dth.size());
  sv.n_lin = static_cast<uint32_t>(m->dlin.size());
  sv.theta_stride = 1;
  sv.tally = m->tally ? 1u : 0u;
  sv.coords = m->d_coords.as<uint32_t>();
  sv.coords_cap_bytes = m->d_coords.cap;
  sv.chunk_sums = m->d_chunk.as<double>();
  sv.chunk_cap_doubles = m->d_chunk.cap / sizeof(double);
  return sv;
}

// Host staging for `bytes` of H2D source data.  Normal mode: the handle's pinned buffer
// (the caller synchronises before returning, so it can be reused by the next call).
// Pipelined mode: a slice of the arena; when the arena is exhausted the stream is
// drained once and the arena starts over.
int stage_alloc(ndt2d_matcher * m, size_t bytes, char ** out)
{
  bytes = (bytes + 63) & ~size_t(63);
  if (!m->pipelined) {
    const int rc = m->h_stage.ensure(bytes);
    *out = m->h_stage.as<char>();
    return rc;
  }
  if (m->arena_off + bytes > m->h_arena.cap) {
    NDT2D_CUDA_TRY(cudaStreamSynchronize(m->stream));
    m->arena_off = 0;
    if (bytes > m->h_arena.cap) {
      const int rc = m->h_arena.ensure(bytes * 2);
      if (rc) {return rc;}
    }
  }
  *out = m->h_arena.as<char>() + m->arena_off;
  m->arena_off += bytes;
  return NDT2D_OK;
}

// scan_matcher_ndt.cpp:95-96,110 : n = min(laser_max_beams, N); j = size_t(i * (N / n))
size_t subsample_count(const ndt2d_matcher * m, size_t npts)
{
  return m->max_beams < npts ? m->max_beams : npts;
}

void subsample_points(const double * pts_xy, size_t npts, size_t n_use, double * out_xy)
{
  const double step = static_cast<double>(npts) / static_cast<double>(n_use);
  for (size_t i = 0; i < n_use; ++i) {
    const size_t j = static_cast<size_t>(i * step);
    out_xy[2 * i] = pts_xy[2 * j];
    out_xy[2 * i + 1] = pts_xy[2 * j + 1];
  }
}

// Bounding box of the scan poses +- range_max (scan_matcher_ndt.cpp:53-64; max_* start at
// DBL_MIN > 0) and the grid NDT::NDT makes of it (ndt_model.cpp:118-126).
int grid_from_poses(const ndt2d_matcher * m, size_t n_scans, const double * poses, GridDesc * out)
{
  double min_x = DBL_MAX, max_x = DBL_MIN, min_y = DBL_MAX, max_y = DBL_MIN;
  for (size_t k = 0; k < n_scans; ++k) {
    const double * pose = poses + 3 * k;
    min_x = std::min(pose[0] - m->prm.range_max, min_x);
    max_x = std::max(pose[0] + m->prm.range_max, max_x);
    min_y = std::min(pose[1] - m->prm.range_max, min_y);
    max_y = std::max(pose[1] + m->prm.range_max, max_y);
  }
  // NDT::NDT (ndt_model.cpp:118-126): size = size_t(extent / cell + 1)
  const double cell = m->prm.ndt_resolution;
  const double fx = ((max_x - min_x) / cell) + 1, fy = ((max_y - min_y) / cell) + 1;
  if (!(fx >= 0.0) || !(fy >= 0.0) || !std::isfinite(fx) || !std::isfinite(fy) ||
    !std::isfinite(min_x) || !std::isfinite(min_y))
  {
    return NDT2D_ERR_INVALID;
  }
  if (fx > 65000.0 || fy > 65000.0) {return NDT2D_ERR_SIZE;}
  GridDesc g;
  g.origin_x = min_x;
  g.origin_y = min_y;
  g.cell_size = cell;
  g.size_x = static_cast<uint32_t>(static_cast<size_t>(fx));
  g.size_y = static_cast<uint32_t>(static_cast<size_t>(fy));
  const uint64_t n_cells64 = static_cast<uint64_t>(g.size_x) * g.size_y;
  const uint64_t n_padded64 = static_cast<uint64_t>(g.size_x + 2) * (g.size_y + 2);
  if (n_padded64 >= (1ull << 28)) {return NDT2D_ERR_SIZE;}
  g.pitch = g.size_x + 2;
  g.n_cells = static_cast<uint32_t>(n_cells64);
  g.n_padded = static_cast<uint32_t>(n_padded64);
  g.n_words = (g.n_padded + 31) / 32;
  g.lin_res = m->prm.search_linear_resolution;
  g.dil_x = ndt2d_region_dilate_x(cell, g.lin_res, static_cast<uint32_t>(m->dlin.size()));
  *out = g;
  return NDT2D_OK;
}

int add_scans_impl(
  ndt2d_matcher * m, size_t n_scans, const double * poses, const uint64_t * pt_offsets,
  const double * pts_xy);

// addScans of a small model (a rolling window) does not wait for the device: its host
// staging comes from the pinned arena (recycled only after a stream synchronisation), so the
// caller goes on to scoreScan / matchScan while the build runs -- the way the node calls
// them back to back (ndt_mapper.cpp:508-515).  Large models keep the synchronous path.
int add_scans_locked(
  ndt2d_matcher * m, size_t n_scans, const double * poses, const uint64_t * pt_offsets,
  const double * pts_xy)
{
  const size_t n_points = n_scans ? static_cast<size_t>(pt_offsets[n_scans] - pt_offsets[0]) : 0;
  if (m->pipelined || n_points * sizeof(double2) > (size_t(1) << 20)) {
    return add_scans_impl(m, n_scans, poses, pt_offsets, pts_xy);
  }
  int rc = m->h_arena.ensure(size_t(8) << 20);
  if (rc) {return rc;}
  m->pipelined = true;
  rc = add_scans_impl(m, n_scans, poses, pt_offsets, pts_xy);
  m->pipelined = false;
  return rc;
}

int add_scans_impl(
  ndt2d_matcher * m, size_t n_scans, const double * poses, const uint64_t * pt_offsets,
  const double * pts_xy)
{
  m->has_model = false;  // (a staged scan stays valid across rebuilds)
  GridDesc g;
  {
    const int grc = grid_from_poses(m, n_scans, poses, &g);
    if (grc) {return grc;}
  }
  const uint64_t n_cells64 = static_cast<uint64_t>(g.size_x) * g.size_y;

  const size_t n_points = n_scans ? static_cast<size_t>(pt_offsets[n_scans] - pt_offsets[0]) : 0;
  if (n_points >= (1ull << 32) - 1) {return NDT2D_ERR_SIZE;}
  const uint64_t off0 = n_scans ? pt_offsets[0] : 0;

  const size_t tf_bytes = n_scans * sizeof(double4);
  const size_t off_bytes = (n_scans + 1) * sizeof(uint64_t);
  const size_t thr_bytes = (static_cast<size_t>(g.size_x) + 2 + g.size_y + 2) * sizeof(double);
  int rc = NDT2D_OK;

  // ---- device buffers.  Everything the build reads from the host sits in ONE buffer,
  // [thr_x thr_y | per-scan transforms | offsets | points], so that a small model (every
  // rolling window) is uploaded by a single copy from one pinned staging block.
  const size_t np1 = n_points ? n_points : 1;
  const size_t o_tf = (thr_bytes + 31) & ~size_t(31);
  const size_t o_off = o_tf + tf_bytes;
  const size_t o_pts = (o_off + off_bytes + 15) & ~size_t(15);
  const size_t in_bytes = o_pts + np1 * sizeof(double2);
  if ((rc = m->d_build_in.ensure(in_bytes))) {return rc;}
  char * const d_in = m->d_build_in.as<char>();
  m->d_thr = reinterpret_cast<double *>(d_in);
  double4 * const d_scan_tf = reinterpret_cast<double4 *>(d_in + o_tf);
  uint64_t * const d_offsets = reinterpret_cast<uint64_t *>(d_in + o_off);
  double2 * const d_mappts = reinterpret_cast<double2 *>(d_in + o_pts);
  if ((rc = m->d_wx.ensure(np1 * sizeof(double)))) {return rc;}
  if ((rc = m->d_wy.ensure(np1 * sizeof(double)))) {return rc;}
  if ((rc = m->d_sx.ensure(np1 * sizeof(double)))) {return rc;}
  if ((rc = m->d_sy.ensure(np1 * sizeof(double)))) {return rc;}
  if ((rc = m->d_nheads.ensure(sizeof(uint32_t)))) {return rc;}
  if ((rc = m->d_key0.ensure(np1 * sizeof(uint32_t)))) {return rc;}
  if ((rc = m->d_key1.ensure(np1 * sizeof(uint32_t)))) {return rc;}
  if ((rc = m->d_val0.ensure(np1 * sizeof(uint32_t)))) {return rc;}
  if ((rc = m->d_val1.ensure(np1 * sizeof(uint32_t)))) {return rc;}
  if ((rc = m->d_seglen.ensure(np1 * sizeof(uint32_t)))) {return rc;}
  const size_t sort_blocks = (np1 + 4095) / 4096;
  if ((rc = m->d_hist.ensure(ndt2d_sort_scratch_bytes(np1)))) {return rc;}
  const size_t scan_n = std::max<size_t>(sort_blocks * 256, g.n_words);
  if ((rc = m->d_scantmp.ensure(((scan_n + 8191) / 8192 + 1) * sizeof(uint32_t)))) {return rc;}
  // +4 words of slack: the search kernel's bulk copies round sizes up to 16 bytes
  if ((rc = m->d_occ.ensure((static_cast<size_t>(g.n_words) + 4) * sizeof(uint2)))) {return rc;}
  if ((rc = m->d_occd.ensure((static_cast<size_t>(g.n_words) + 4) * sizeof(uint32_t)))) {return rc;}
  if ((rc = m->d_nvalid.ensure(2 * sizeof(uint32_t)))) {return rc;}   // n_valid, n_stiff
  const uint64_t cap64 = std::min<uint64_t>(n_cells64, n_points / 5) + 1;
  m->rec_cap = static_cast<uint32_t>(cap64);
  if ((rc = m->d_rec.ensure(cap64 * NDT2D_REC_DOUBLES * sizeof(double)))) {return rc;}
  if ((rc = m->d_rec_fast.ensure(cap64 * NDT2D_REC_DOUBLES * sizeof(double)))) {return rc;}
  if ((rc = m->d_rec_vtx.ensure(cap64 * NDT2D_REC_DOUBLES * sizeof(double)))) {return rc;}
  if ((rc = m->d_heads.ensure(cap64 * sizeof(uint2)))) {return rc;}

  m->bs.wx = m->d_wx.as<double>();
  m->bs.wy = m->d_wy.as<double>();
  m->bs.sx = m->d_sx.as<double>();
  m->bs.sy = m->d_sy.as<double>();
  m->bs.heads = m->d_heads.as<uint2>();
  m->bs.n_heads = m->d_nheads.as<uint32_t>();
  m->bs.key[0] = m->d_key0.as<uint32_t>();
  m->bs.key[1] = m->d_key1.as<uint32_t>();
  m->bs.val[0] = m->d_val0.as<uint32_t>();
  m->bs.val[1] = m->d_val1.as<uint32_t>();
  m->bs.seglen = m->d_seglen.as<uint32_t>();
  m->bs.hist = m->d_hist.as<uint32_t>();
  m->bs.scan_tmp = m->d_scantmp.as<uint32_t>();
  if (!m->d_rcp.p) {
    if ((rc = m->d_rcp.ensure((NDT2D_RCP_TABLE + 1) * sizeof(double)))) {return rc;}
    m->bs.rcp_ready = false;
  }
  m->bs.rcp = m->d_rcp.as<double>();

  // ---- uploads.  Large model: the points (the bulk of the bytes) go first, straight from the
  // caller's buffer, and the host-side staging below overlaps that copy.  Small model
  // (pipelined: nobody waits for the device): points and staging share one arena block and
  // one copy.
  cudaStream_t st = m->stream;
  const double * src_pts = pts_xy + 2 * off0;
  char * hs = nullptr;
  const bool one_copy = m->pipelined;
  if (one_copy) {
    if ((rc = stage_alloc(m, in_bytes, &hs))) {return rc;}
    if (n_points) {memcpy(hs + o_pts, src_pts, n_points * sizeof(double2));}
  } else {
    if (n_points) {
      NDT2D_CUDA_TRY(cudaMemcpyAsync(d_mappts, src_pts, n_points * sizeof(double2),
        cudaMemcpyHostToDevice, st));
    }
    if ((rc = stage_alloc(m, o_pts, &hs))) {
      cudaStreamSynchronize(st);  // the points copy reads the caller's buffer
      return rc;
    }
  }

  // ---- host staging: axis thresholds, per-scan transform, rebased offsets
  std::vector<double> thr_x, thr_y;
  axis_thresholds(g.origin_x, g.cell_size, g.size_x, thr_x);
  axis_thresholds(g.origin_y, g.cell_size, g.size_y, thr_y);
  double * h_thr = reinterpret_cast<double *>(hs);
  double4 * h_tf = reinterpret_cast<double4 *>(hs + o_tf);
  uint64_t * h_off = reinterpret_cast<uint64_t *>(hs + o_off);
  for (size_t k = 0; k < n_scans; ++k) {
    const double * pose = poses + 3 * k;
    // ndt_model.cpp:135-136
    h_tf[k] = make_double4(pose[0], pose[1], cos(pose[2]), sin(pose[2]));
    h_off[k] = pt_offsets[k] - off0;
  }
  h_off[n_scans] = n_points;
  memcpy(h_thr, thr_x.data(), thr_x.size() * sizeof(double));
  memcpy(h_thr + thr_x.size(), thr_y.data(), thr_y.size() * sizeof(double));

  NDT2D_CUDA_TRY(cudaMemcpyAsync(d_in, hs, one_copy ? o_pts + n_points * sizeof(double2) : o_off + off_bytes,
    cudaMemcpyHostToDevice, st));
  m->ctr.h2d_bytes += tf_bytes + off_bytes + thr_bytes + n_points * sizeof(double2);

  const bool small_build = ndt2d_build_is_small(g, n_points);
  const bool timed_build = !small_build || m->time_small;
  m->evb_valid = false;
  if (timed_build && m->evb_begin) {cudaEventRecord(m->evb_begin, st);}
  rc = ndt2d_launch_build(g, d_scan_tf, d_offsets, n_scans,
      d_mappts, n_points, m->bs, m->d_occ.as<uint2>(), m->d_occd.as<uint32_t>(),
      m->d_rec.as<double>(), m->d_rec_fast.as<double>(), m->d_rec_vtx.as<double>(), m->rec_cap,
      m->d_nvalid.as<uint32_t>(), st,
      &m->ctr, &m->sorted_buf);
  if (rc) {return rc;}
  if (timed_build && m->evb_end) {
    cudaEventRecord(m->evb_end, st);
    m->evb_valid = true;
  }
  if (!m->pipelined) {
    NDT2D_CUDA_TRY(cudaStreamSynchronize(st));  // host staging is reused by the next call
  }
  m->g = g;
  m->n_map_points = n_points;
  m->has_model = true;
  return NDT2D_OK;
}

int stage_scan_locked(ndt2d_matcher * m, const double * pose3, const double * pts_xy, size_t npts,
  bool lazy_trig = false)
{
  const size_t n_use = subsample_count(m, npts);
  if (n_use >= (1u << 30)) {return NDT2D_ERR_SIZE;}
  const size_t n_ang = m->dth.size();
  const size_t pts_bytes = n_use * sizeof(double2);
  const size_t trig_bytes = n_ang * sizeof(double2);
  char * hs = nullptr;
  const size_t up_bytes = pts_bytes + (lazy_trig ? 0 : trig_bytes);
  // small uploads are staged from the pinned arena (recycled only after a stream
  // synchronisation), so the search is enqueued right behind the copy without waiting for it
  const bool from_arena = m->pipelined || up_bytes + 64 <= (size_t(1) << 20);
  int rc = NDT2D_OK;
  if (from_arena && !m->pipelined) {
    if ((rc = m->h_arena.ensure(size_t(8) << 20))) {return rc;}
    m->pipelined = true;
    rc = stage_alloc(m, up_bytes + 64, &hs);
    m->pipelined = false;
  } else {
    rc = stage_alloc(m, up_bytes + 64, &hs);
  }
  if (rc) {return rc;}
  // points and per-theta (cos, sin) share one device buffer: one H2D copy per scan
  if ((rc = m->d_pts.ensure(pts_bytes + trig_bytes + 16))) {return rc;}
  double * h_pts = reinterpret_cast<double *>(hs);
  double * h_trig = h_pts + 2 * n_use;
  if (n_use) {subsample_points(pts_xy, npts, n_use, h_pts);}
  m->trig_lazy = lazy_trig;
  m->pose_th = pose3[2];
  if (lazy_trig) {
    m->h_trig_all.assign(2 * n_ang, 0.0);
    m->trig_done.assign(n_ang, 0);
  } else {
    for (size_t k = 0; k < n_ang; ++k) {
      // scan_matcher_ndt.cpp:106-107
      h_trig[2 * k] = cos(pose3[2] + m->dth[k]);
      h_trig[2 * k + 1] = sin(pose3[2] + m->dth[k]);
    }
  }
  cudaStream_t st = m->stream;
  if (up_bytes) {
    NDT2D_CUDA_TRY(cudaMemcpyAsync(m->d_pts.p, h_pts, up_bytes, cudaMemcpyHostToDevice, st));
  }
  m->trig_offset_bytes = pts_bytes;
  m->ctr.h2d_bytes += up_bytes;
  m->pose_x = pose3[0];
  m->pose_y = pose3[1];
  m->n_pts = static_cast<uint32_t>(n_use);
  // search scratch
  const size_t scratch = ndt2d_search_scratch_doubles(
    static_cast<uint32_t>(n_ang), static_cast<uint32_t>(m->dlin.size()), m->prm.ndt_resolution,
    m->prm.search_linear_resolution);
  if ((rc = m->d_blockpart.ensure(scratch * sizeof(double)))) {return rc;}
  if ((rc = m->d_partial.ensure(32 * sizeof(double)))) {return rc;}
  {
    void * before = m->d_counter.p;
    if ((rc = m->d_counter.ensure(64))) {return rc;}
    if (m->d_counter.p != before) {
      // job counter + statistics start at zero; every search's finish kernel re-zeroes them
      NDT2D_CUDA_TRY(cudaMemsetAsync(m->d_counter.p, 0, m->d_counter.cap, m->stream));
    }
  }
  {
    // coordinate pre-pass table of the search kernel (skipped above 512 MiB)
    const size_t cb = ndt2d_region_coords_bytes(m->prm.ndt_resolution, static_cast<uint32_t>(n_ang),
        static_cast<uint32_t>(m->dlin.size()), m->prm.search_linear_resolution,
        static_cast<uint32_t>(n_use), size_t(512) << 20);
    if (cb && (rc = m->d_coords.ensure(cb))) {return rc;}
    const size_t cd = ndt2d_region_chunk_doubles(m->prm.ndt_resolution, static_cast<uint32_t>(n_ang),
        static_cast<uint32_t>(m->dlin.size()), m->prm.search_linear_resolution,
        static_cast<uint32_t>(n_use));
    if (cd && (rc = m->d_chunk.ensure(cd * sizeof(double)))) {return rc;}
  }
  if ((rc = m->h_result.ensure(64 * sizeof(double)))) {return rc;}
  // the single pinned staging buffer is reused by the next call: wait for the copies
  if (!from_arena) {
    NDT2D_CUDA_TRY(cudaStreamSynchronize(st));
  }
  m->staged = true;
  return NDT2D_OK;
}

void unpack_result(const double * r32, double * out_delta3, int * delta_written,
  double * out_cov9, double * out_score)
{
  const bool written = r32[19] != 0.0;
  if (delta_written) {*delta_written = written ? 1 : 0;}
  if (written && out_delta3) {
    out_delta3[0] = r32[16];
    out_delta3[1] = r32[17];
    out_delta3[2] = r32[18];
  }
  if (out_cov9) {memcpy(out_cov9, r32 + 20, 9 * sizeof(double));}
  if (out_score) {*out_score = r32[29];}
}

// Host side of the HostMailbox (ndt2d_internal.h): the record lives in h_result[0..31], the
// flag in h_result[32].  mailbox_arm() before the launch, mailbox_wait() after it.
HostMailbox mailbox_arm(ndt2d_matcher * m)
{
  double * h = m->h_result.as<double>();
  unsigned long long * flag = reinterpret_cast<unsigned long long *>(h + 32);
  *reinterpret_cast<volatile unsigned long long *>(flag) = 0ull;
  m->host_seq += 1;
  return HostMailbox{h, flag, m->host_seq};
}

inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#elif defined(__aarch64__)
  asm volatile("yield" ::: "memory");
#else
  asm volatile("" ::: "memory");
#endif
}

// `spin`: poll the flag (a local match finishes in tens of microseconds: waking up from a
// blocking synchronisation would cost as much as the search); falls back to the blocking wait
// after ~2 ms or on a stream error.  Large searches block right away.
int mailbox_wait(ndt2d_matcher * m, const HostMailbox & hm, bool spin, double * r32)
{
  const volatile unsigned long long * flag = hm.flag;
  bool seen = false;
  if (spin) {
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t it = 0;; ++it) {
      if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) == hm.seq) {
        seen = true;
        break;
      }
      cpu_relax();
      if ((it & 1023u) == 1023u) {
        if (cudaStreamQuery(m->stream) != cudaErrorNotReady) {break;}   // done or failed
        if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) {break;}
      }
    }
  }
  if (!seen) {
    NDT2D_CUDA_TRY(cudaStreamSynchronize(m->stream));
    if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) != hm.seq) {
      ndt2d_set_error("search result mailbox not written", cudaErrorUnknown, __FILE__, __LINE__);
      return NDT2D_ERR_CUDA;
    }
  }
  m->ctr.d2h_bytes += 32 * sizeof(double);
  memcpy(r32, hm.out32, 32 * sizeof(double));
  return NDT2D_OK;
}

int fetch_result_locked(ndt2d_matcher * m, double * r32)
{
  double * h = m->h_result.as<double>();
  NDT2D_CUDA_TRY(cudaMemcpyAsync(h, m->d_partial.p, 32 * sizeof(double), cudaMemcpyDeviceToHost,
    m->stream));
  NDT2D_CUDA_TRY(cudaStreamSynchronize(m->stream));
  m->ctr.d2h_bytes += 32 * sizeof(double);
  memcpy(r32, h, 32 * sizeof(double));
  return NDT2D_OK;
}

// (cos, sin) of the theta slices begin, begin + stride, ... < end of a lazily staged scan:
// the missing ones are computed on the host (libm, scan_matcher_ndt.cpp:106-107) into the
// mirror, then the covering range of the mirror is uploaded -- entries of other slices in
// that range are either already valid or never read before they are computed.
int ensure_trig_locked(ndt2d_matcher * m, size_t begin, size_t end, size_t stride)
{
  if (!m->trig_lazy || begin >= end) {return NDT2D_OK;}
  size_t lo = end, hi = begin;
  for (size_t k = begin; k < end; k += stride) {
    if (m->trig_done[k]) {continue;}
    m->h_trig_all[2 * k] = cos(m->pose_th + m->dth[k]);
    m->h_trig_all[2 * k + 1] = sin(m->pose_th + m->dth[k]);
    m->trig_done[k] = 1;
    lo = std::min(lo, k);
    hi = std::max(hi, k);
  }
  if (lo > hi) {return NDT2D_OK;}
  const size_t bytes = (hi - lo + 1) * sizeof(double2);
  // staged from the pinned arena (recycled only after a stream synchronisation): the search
  // that follows is enqueued without waiting for this copy
  char * hs = nullptr;
  int rc = m->h_arena.ensure(size_t(8) << 20);
  if (rc) {return rc;}
  const bool was_pipelined = m->pipelined;
  m->pipelined = true;
  rc = stage_alloc(m, bytes, &hs);
  m->pipelined = was_pipelined;
  if (rc) {return rc;}
  memcpy(hs, m->h_trig_all.data() + 2 * lo, bytes);
  NDT2D_CUDA_TRY(cudaMemcpyAsync(m->d_pts.as<char>() + m->trig_offset_bytes + lo * sizeof(double2), hs,
    bytes, cudaMemcpyHostToDevice, m->stream));
  m->ctr.h2d_bytes += bytes;
  return NDT2D_OK;
}

int match_scan_locked(
  ndt2d_matcher * m, const double * pose3, const double * pts_xy, size_t npts,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score)
{
  int rc = stage_scan_locked(m, pose3, pts_xy, npts);
  if (rc) {return rc;}
  // a small search (the local match of every scan): no event records, result through the
  // host mailbox with a polling wait
  const bool small = static_cast<double>(m->dth.size()) * m->dlin.size() * m->dlin.size() *
    m->n_pts < 2.0e7;
  const bool timed = !small || m->time_small;
  const HostMailbox hm = mailbox_arm(m);
  rc = ndt2d_launch_search(model_view(m), search_view(m), 0, static_cast<uint32_t>(m->dth.size()),
      m->prm.kernel_variant, m->d_blockpart.as<double>(), m->d_partial.as<double>(), nullptr,
      m->d_counter.as<uint32_t>(), m->stream, &m->ctr, timed ? m->ev_begin : nullptr,
      timed ? m->ev_end : nullptr, nullptr, &hm);
  if (rc) {return rc;}
  m->ev_valid = timed;
  double r32[32];
  if ((rc = mailbox_wait(m, hm, small, r32))) {return rc;}
  unpack_result(r32, out_delta3, delta_written, out_cov9, out_score);
  return NDT2D_OK;
}

int score_poses_locked(
  ndt2d_matcher * m, const double * pts_xy, size_t npts, const double * poses3, size_t n_poses,
  double sign, int normalise, bool subsample, double * out_scores)
{
  const size_t n_use = subsample ? subsample_count(m, npts) : npts;
  if (n_use >= (1u << 30) || n_poses >= (1u << 30)) {return NDT2D_ERR_SIZE;}
  const size_t pts_bytes = n_use * sizeof(double2);
  const size_t tf_bytes = n_poses * sizeof(double4);
  const size_t out_bytes = n_poses * sizeof(double);
  const size_t up_bytes = tf_bytes + pts_bytes;
  // a handful of poses (scoreScan / scorePoints of the node: one pose) is latency-bound: staged
  // from the pinned arena (no wait for the upload), scores through the host mailbox
  const bool small = n_poses <= 8 && up_bytes + 64 <= (size_t(1) << 20);
  int rc = NDT2D_OK;
  char * hs = nullptr;
  if (small) {
    if ((rc = m->h_arena.ensure(size_t(8) << 20))) {return rc;}
    const bool was = m->pipelined;
    m->pipelined = true;
    rc = stage_alloc(m, up_bytes + 64, &hs);
    m->pipelined = was;
  } else {
    rc = stage_alloc(m, up_bytes + 64, &hs);
  }
  if (rc) {return rc;}
  // poses and points share one device buffer: one H2D copy per call
  if ((rc = m->d_pose_tf.ensure(up_bytes + 32))) {return rc;}
  if ((rc = m->d_out.ensure(out_bytes ? out_bytes : 8))) {return rc;}
  if ((rc = m->h_result.ensure(std::max<size_t>(out_bytes, 64 * sizeof(double))))) {return rc;}
  double4 * h_tf = reinterpret_cast<double4 *>(hs);
  double * h_pts = reinterpret_cast<double *>(h_tf + n_poses);
  for (size_t p = 0; p < n_poses; ++p) {
    const double * pose = poses3 + 3 * p;
    h_tf[p] = make_double4(pose[0], pose[1], cos(pose[2]), sin(pose[2]));
  }
  if (n_use) {
    if (subsample) {
      subsample_points(pts_xy, npts, n_use, h_pts);
    } else {
      memcpy(h_pts, pts_xy, pts_bytes);
    }
  }
  cudaStream_t st = m->stream;
  if (up_bytes) {
    NDT2D_CUDA_TRY(cudaMemcpyAsync(m->d_pose_tf.p, hs, up_bytes, cudaMemcpyHostToDevice, st));
  }
  m->ctr.h2d_bytes += up_bytes;
  const double2 * d_pts = reinterpret_cast<const double2 *>(m->d_pose_tf.as<char>() + tf_bytes);
  if (small && n_poses) {
    const HostMailbox hm = mailbox_arm(m);
    rc = ndt2d_launch_score_poses(model_view(m), d_pts, static_cast<uint32_t>(n_use),
        m->d_pose_tf.as<double4>(), static_cast<uint32_t>(n_poses), sign, normalise,
        m->d_out.as<double>(), st, &m->ctr, &hm);
    if (rc) {return rc;}
    double r32[32];
    if ((rc = mailbox_wait(m, hm, true, r32))) {return rc;}
    m->ctr.d2h_bytes -= 32 * sizeof(double);
    m->ctr.d2h_bytes += out_bytes;
    memcpy(out_scores, r32, out_bytes);
    return NDT2D_OK;
  }
  rc = ndt2d_launch_score_poses(model_view(m), d_pts, static_cast<uint32_t>(n_use),
      m->d_pose_tf.as<double4>(), static_cast<uint32_t>(n_poses), sign, normalise,
      m->d_out.as<double>(), st, &m->ctr);
  if (rc) {return rc;}
  if (out_bytes) {
    NDT2D_CUDA_TRY(cudaMemcpyAsync(m->h_result.p, m->d_out.p, out_bytes, cudaMemcpyDeviceToHost, st));
  }
  NDT2D_CUDA_TRY(cudaStreamSynchronize(st));
  m->ctr.d2h_bytes += out_bytes;
  if (out_bytes) {memcpy(out_scores, m->h_result.p, out_bytes);}
  return NDT2D_OK;
}

}  // namespace

static int match_scan_batch_locked(
  ndt2d_matcher * m, size_t n_jobs,
  const uint64_t * job_scan_offsets, const double * map_poses, const uint64_t * map_pt_offsets,
  const double * map_pts_xy,
  const double * query_poses, const uint64_t * query_pt_offsets, const double * query_pts_xy,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score);
static int group_create(ndt2d_matcher * m);
static int group_run(ndt2d_matcher * m, const std::function<int(size_t, ndt2d_matcher *)> & fn);
static bool group_worthwhile(const ndt2d_matcher * m, size_t npts);
static int match_scan_group(
  ndt2d_matcher * m, const double * pose3, const double * pts_xy, size_t npts,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score);
static void exchange_close(ndt2d_matcher * m)
{
  for (uint32_t r = 0; r < m->peer_ptrs.size(); ++r) {
    if (!m->x_local && r != m->x_rank && m->peer_ptrs[r]) {cudaIpcCloseMemHandle(m->peer_ptrs[r]);}
  }
  m->x_local = false;
  m->peer_ptrs.clear();
  m->x_connected = false;
}


extern "C" {

NDT2D_API const char * ndt2d_version(void) {return "ndt2d_b200 0.1 (sm_100a)";}
NDT2D_API const char * ndt2d_last_error(void) {return g_last_error;}

NDT2D_API int ndt2d_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

NDT2D_API void ndt2d_default_params(ndt2d_params * p)
{
  if (!p) {return;}
  p->ndt_resolution = 0.25;
  p->search_angular_resolution = 0.0025;
  p->search_angular_size = 0.1;
  p->search_linear_resolution = 0.005;
  p->search_linear_size = 0.05;
  p->laser_max_beams = 100;
  p->range_max = 0.0;
  p->device = -1;
  p->stream = nullptr;
  p->kernel_variant = 0;
  p->n_devices = 0;
  for (int k = 0; k < NDT2D_MAX_DEVICES; ++k) {p->devices[k] = 0;}
}

NDT2D_API int ndt2d_matcher_create(const ndt2d_params * params, ndt2d_matcher ** out)
{
  if (!params || !out) {return NDT2D_ERR_INVALID;}
  *out = nullptr;
  if (!(params->ndt_resolution > 0.0) || !std::isfinite(params->ndt_resolution) ||
    !std::isfinite(params->range_max))
  {
    return NDT2D_ERR_INVALID;
  }
  if (ndt2d_device_count() <= 0) {
    snprintf(g_last_error, sizeof(g_last_error),
      "no CUDA device visible: libndt2d_b200 has no CPU fallback");
    return NDT2D_ERR_NO_DEVICE;
  }
  if (params->n_devices < 0 || params->n_devices > NDT2D_MAX_DEVICES) {return NDT2D_ERR_INVALID;}
  const bool multi = params->n_devices > 1;
  if (multi) {
    const int n_vis = ndt2d_device_count();
    for (int a = 0; a < params->n_devices; ++a) {
      if (params->devices[a] < 0 || params->devices[a] >= n_vis) {return NDT2D_ERR_NO_DEVICE;}
      for (int b = 0; b < a; ++b) {
        if (params->devices[a] == params->devices[b]) {return NDT2D_ERR_INVALID;}
      }
    }
  }
  int dev = multi ? params->devices[0] : (params->n_devices == 1 ? params->devices[0] : params->device);
  if (dev < 0) {
    NDT2D_CUDA_TRY(cudaGetDevice(&dev));
  }
  ndt2d_matcher * m = new (std::nothrow) ndt2d_matcher();
  if (!m) {return NDT2D_ERR_INVALID;}
  m->prm = *params;
  m->device = dev;
  // declare_parameter<int> stored into a size_t (scan_matcher_ndt.cpp:44, hpp:99)
  m->max_beams = static_cast<size_t>(params->laser_max_beams);
  int rc = replay_loop(params->search_angular_size, params->search_angular_resolution,
      1u << 24, m->dth);
  if (!rc) {
    rc = replay_loop(params->search_linear_size, params->search_linear_resolution, 65535, m->dlin);
  }
  if (rc) {
    delete m;
    return rc;
  }
  DeviceGuard guard(dev);
  if (!guard.ok) {
    delete m;
    return NDT2D_ERR_NO_DEVICE;
  }
  if (params->stream && !multi) {
    m->stream = static_cast<cudaStream_t>(params->stream);
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      ndt2d_set_error("cudaStreamCreate", e, __FILE__, __LINE__);
      delete m;
      return NDT2D_ERR_CUDA;
    }
    m->own_stream = true;
  }
  if (cudaEventCreate(&m->ev_begin) != cudaSuccess || cudaEventCreate(&m->ev_end) != cudaSuccess ||
    cudaEventCreate(&m->evb_begin) != cudaSuccess || cudaEventCreate(&m->evb_end) != cudaSuccess)
  {
    m->ev_begin = m->ev_end = m->evb_begin = m->evb_end = nullptr;  // timing is optional
    cudaGetLastError();
  }
  const size_t na = m->dth.size(), nl = m->dlin.size();
  rc = m->d_dth.ensure((na ? na : 1) * sizeof(double));
  if (!rc) {rc = m->d_dlin.ensure((nl ? nl : 1) * sizeof(double));}
  if (!rc && na) {
    if (cudaMemcpy(m->d_dth.p, m->dth.data(), na * sizeof(double), cudaMemcpyHostToDevice) !=
      cudaSuccess) {rc = NDT2D_ERR_CUDA;}
  }
  if (!rc && nl) {
    if (cudaMemcpy(m->d_dlin.p, m->dlin.data(), nl * sizeof(double), cudaMemcpyHostToDevice) !=
      cudaSuccess) {rc = NDT2D_ERR_CUDA;}
  }
  if (rc) {
    ndt2d_matcher_destroy(m);
    return rc;
  }
  m->ctr.h2d_bytes += (na + nl) * sizeof(double);
  if (multi) {
    rc = group_create(m);
    if (rc) {
      ndt2d_matcher_destroy(m);
      return rc;
    }
  }
  *out = m;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_destroy(ndt2d_matcher * m)
{
  if (!m) {return NDT2D_OK;}
  for (ndt2d_matcher * sub : m->lanes) {ndt2d_matcher_destroy(sub);}
  m->lanes.clear();
  for (size_t r = 1; r < m->group.size(); ++r) {
    ndt2d_matcher * s = m->group[r];
    if (s->worker) {
      {
        std::lock_guard<std::mutex> wl(s->worker->mu);
        s->worker->quit = true;
      }
      s->worker->cv.notify_all();
      if (s->worker->th.joinable()) {s->worker->th.join();}
      delete s->worker;
      s->worker = nullptr;
    }
    ndt2d_matcher_destroy(s);
  }
  m->group.clear();
  {
    DeviceGuard guard(m->device);
    if (m->stream) {cudaStreamSynchronize(m->stream);}
    exchange_close(m);
    m->d_mailbox.release();
    m->d_peer_table.release();
  }
  {
    DeviceGuard guard(m->device);
    if (m->stream) {cudaStreamSynchronize(m->stream);}
    DeviceBuffer * bufs[] = {&m->d_dth, &m->d_dlin, &m->d_occ, &m->d_occd, &m->d_rec, &m->d_rec_fast, &m->d_rec_vtx, &m->d_nvalid,
      &m->d_sx, &m->d_sy, &m->d_heads, &m->d_nheads, &m->d_wx, &m->d_wy, &m->d_key0, &m->d_key1, &m->d_val0, &m->d_val1, &m->d_seglen,
      &m->d_hist, &m->d_scantmp, &m->d_build_in, &m->d_rcp, &m->d_pts,
      &m->d_trig, &m->d_blockpart, &m->d_partial, &m->d_pose_tf, &m->d_out, &m->d_counter, &m->d_coords, &m->d_chunk, &m->d_batch_results, &m->d_batch_arena};
    for (DeviceBuffer * b : bufs) {b->release();}
    m->h_stage.release();
    m->h_result.release();
    m->h_arena.release();
    if (m->ev_begin) {cudaEventDestroy(m->ev_begin);}
    if (m->ev_end) {cudaEventDestroy(m->ev_end);}
    if (m->evb_begin) {cudaEventDestroy(m->evb_begin);}
    if (m->evb_end) {cudaEventDestroy(m->evb_end);}
    if (m->own_stream && m->stream) {cudaStreamDestroy(m->stream);}
  }
  delete m;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_reset(ndt2d_matcher * m)
{
  if (!m) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  m->has_model = false;  // scan_matcher_ndt.cpp:180-183
  m->n_map_points = 0;
  for (size_t r = 1; r < m->group.size(); ++r) {
    m->group[r]->has_model = false;
    m->group[r]->n_map_points = 0;
  }
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_add_scans(
  ndt2d_matcher * m, size_t n_scans, const double * poses, const uint64_t * pt_offsets,
  const double * pts_xy)
{
  if (!m || (n_scans && (!poses || !pt_offsets))) {return NDT2D_ERR_INVALID;}
  if (n_scans && pt_offsets[n_scans] > pt_offsets[0] && !pts_xy) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  // a multi-device handle replicates the model: every device builds it from the same scans,
  // all devices at once (one host thread per device)
  if (m->group.size() > 1) {
    return group_run(m, [&](size_t, ndt2d_matcher * s) {
               return add_scans_locked(s, n_scans, poses, pt_offsets, pts_xy);
             });
  }
  DeviceGuard guard(m->device);
  return add_scans_locked(m, n_scans, poses, pt_offsets, pts_xy);
}

NDT2D_API int ndt2d_matcher_match_scan(
  ndt2d_matcher * m, const double * pose3, const double * pts_xy, size_t npts,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score)
{
  if (!m || !pose3 || (npts && !pts_xy)) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  if (delta_written) {*delta_written = 0;}
  if (!m->has_model) {
    if (out_score) {*out_score = 0.0;}  // scan_matcher_ndt.cpp:80
    return NDT2D_ERR_NO_MAP;
  }
  if (group_worthwhile(m, npts)) {
    return match_scan_group(m, pose3, pts_xy, npts, out_delta3, delta_written, out_cov9, out_score);
  }
  DeviceGuard guard(m->device);
  return match_scan_locked(m, pose3, pts_xy, npts, out_delta3, delta_written, out_cov9, out_score);
}

NDT2D_API int ndt2d_matcher_score_points(
  ndt2d_matcher * m, const double * pts_xy, size_t npts, const double * pose3, double * out_score)
{
  if (!m || !pose3 || !out_score || (npts && !pts_xy)) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  if (!m->has_model) {
    *out_score = 0.0;  // scan_matcher_ndt.cpp:159
    return NDT2D_ERR_NO_MAP;
  }
  DeviceGuard guard(m->device);
  return score_poses_locked(m, pts_xy, npts, pose3, 1, -1.0, 1, true, out_score);
}

NDT2D_API int ndt2d_matcher_score_poses(
  ndt2d_matcher * m, const double * pts_xy, size_t npts, const double * poses3, size_t n_poses,
  double * out_scores)
{
  if (!m || (n_poses && (!poses3 || !out_scores)) || (npts && !pts_xy)) {
    return NDT2D_ERR_INVALID;
  }
  std::lock_guard<std::mutex> lock(m->mu);
  if (!m->has_model) {
    for (size_t i = 0; i < n_poses; ++i) {out_scores[i] = 0.0;}
    return NDT2D_ERR_NO_MAP;
  }
  DeviceGuard guard(m->device);
  return score_poses_locked(m, pts_xy, npts, poses3, n_poses, -1.0, 1, true, out_scores);
}

NDT2D_API int ndt2d_matcher_likelihood_scan(
  ndt2d_matcher * m, const double * pose3, const double * pts_xy, size_t npts,
  double * out_likelihood)
{
  if (!m || !pose3 || !out_likelihood || (npts && !pts_xy)) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  if (!m->has_model) {
    *out_likelihood = 0.0;
    return NDT2D_ERR_NO_MAP;
  }
  DeviceGuard guard(m->device);
  return score_poses_locked(m, pts_xy, npts, pose3, 1, 1.0, 0, false, out_likelihood);
}

NDT2D_API int ndt2d_matcher_match_scan_batch(
  ndt2d_matcher * m, size_t n_jobs,
  const uint64_t * job_scan_offsets, const double * map_poses, const uint64_t * map_pt_offsets,
  const double * map_pts_xy,
  const double * query_poses, const uint64_t * query_pt_offsets, const double * query_pts_xy,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score)
{
  if (!m || (n_jobs && (!job_scan_offsets || !map_poses || !map_pt_offsets || !query_poses ||
    !query_pt_offsets)))
  {
    return NDT2D_ERR_INVALID;
  }
  std::lock_guard<std::mutex> lock(m->mu);
  DeviceGuard guard(m->device);
  return match_scan_batch_locked(m, n_jobs, job_scan_offsets, map_poses, map_pt_offsets, map_pts_xy,
           query_poses, query_pt_offsets, query_pts_xy, out_delta3, delta_written, out_cov9,
           out_score);
}

}  // extern "C"

// match_scan_batch as THREE launches (all models built by one grid of single-CTA builds,
// all searches by one persistent search kernel, all final reductions by one grid of
// blocks), one H2D copy of a packed staging area and one D2H copy of the results.
// Returns kBatchNotEligible when a job does not fit the small-model build or the plan
// (the caller then pipelines the jobs over the lanes instead).
constexpr int kBatchNotEligible = -1;

static int match_scan_batch_fused(
  ndt2d_matcher * m, size_t n_jobs,
  const uint64_t * job_scan_offsets, const double * map_poses, const uint64_t * map_pt_offsets,
  const double * map_pts_xy,
  const double * query_poses, const uint64_t * query_pt_offsets, const double * query_pts_xy,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score)
{
  if (m->prm.kernel_variant != 0 || n_jobs == 0 || n_jobs > 4096) {return kBatchNotEligible;}
  const size_t n_ang = m->dth.size(), n_lin = m->dlin.size();
  if (n_ang == 0 || n_lin == 0) {return kBatchNotEligible;}
  struct Job
  {
    GridDesc g;
    size_t n_scans, n_points, n_use, npts_q, rec_cap;
    uint64_t s0, p0, q0;
    size_t o_tf, o_off, o_pts, o_thr, o_q, o_trig;                       // upload region
    size_t o_occ, o_occd, o_rec, o_recf, o_recv, o_nvalid, o_jobp, o_chunk;      // device-only region
  };
  std::vector<Job> jobs(n_jobs);
  uint32_t max_use = 0;
  for (size_t j = 0; j < n_jobs; ++j) {
    Job & J = jobs[j];
    J.s0 = job_scan_offsets[j];
    J.n_scans = static_cast<size_t>(job_scan_offsets[j + 1] - J.s0);
    if (J.n_scans == 0) {return kBatchNotEligible;}
    const int grc = grid_from_poses(m, J.n_scans, map_poses + 3 * J.s0, &J.g);
    if (grc) {return kBatchNotEligible;}
    J.p0 = map_pt_offsets[J.s0];
    J.n_points = static_cast<size_t>(map_pt_offsets[J.s0 + J.n_scans] - J.p0);
    if (!ndt2d_build_is_small(J.g, J.n_points)) {return kBatchNotEligible;}
    J.q0 = query_pt_offsets[j];
    J.npts_q = static_cast<size_t>(query_pt_offsets[j + 1] - J.q0);
    J.n_use = subsample_count(m, J.npts_q);
    if (J.n_use == 0 || J.n_use >= (1u << 20)) {return kBatchNotEligible;}
    max_use = std::max<uint32_t>(max_use, static_cast<uint32_t>(J.n_use));
    J.rec_cap = std::min<uint64_t>(static_cast<uint64_t>(J.g.size_x) * J.g.size_y, J.n_points / 5) + 1;
  }
  RegionBatchPlan pl;
  ndt2d_region_batch_plan(m->prm.ndt_resolution, static_cast<uint32_t>(n_ang),
    static_cast<uint32_t>(n_lin), m->prm.search_linear_resolution, max_use,
    static_cast<uint32_t>(std::min<size_t>(n_jobs, 1u << 16)), &pl);
  // small searches (local-match windows: a region of the region kernel holds too few
  // candidates to fill a warp) go to the window kernel, else the dense one, by the same rule
  // as a single search
  const bool small_search = static_cast<double>(n_ang) * n_lin * n_lin * max_use < 2.0e7 &&
    n_ang <= 65535 && n_jobs <= 65535;
  const uint32_t win_k = small_search ? ndt2d_window_cells(m->prm.ndt_resolution,
      static_cast<uint32_t>(n_lin), m->prm.search_linear_resolution) : 0u;
  const uint32_t dense_records = win_k ?
    ndt2d_window_records(static_cast<uint32_t>(n_ang), static_cast<uint32_t>(n_lin)) :
    ndt2d_dense_batch_records(static_cast<uint32_t>(n_ang), static_cast<uint32_t>(n_lin));
  const bool dense = small_search && dense_records <= 4096;
  if (dense) {
    pl.n_jobs = dense_records;
    pl.P = 1;
    pl.chunk_doubles = 0;
  }
  if (pl.n_jobs == 0 || pl.n_jobs > 4096) {return kBatchNotEligible;}
  // ---- layout: [upload region][device-only region], everything 256-byte aligned
  size_t off = 0;
  auto take = [&off](size_t bytes) {
      const size_t o = off;
      off += (bytes + 255) & ~size_t(255);
      return o;
    };
  const size_t o_counter = take(64);
  const size_t o_build = take(n_jobs * sizeof(BuildEntry));
  const size_t o_batch = take(n_jobs * sizeof(BatchEntry));
  for (Job & J : jobs) {
    J.o_tf = take(J.n_scans * sizeof(double4));
    J.o_off = take((J.n_scans + 1) * sizeof(uint64_t));
    J.o_pts = take(J.n_points * sizeof(double2));
    J.o_thr = take((static_cast<size_t>(J.g.size_x) + J.g.size_y + 4) * sizeof(double));
    J.o_q = take(J.n_use * sizeof(double2));
    J.o_trig = take(n_ang * sizeof(double2));
  }
  const size_t upload_bytes = off;
  for (Job & J : jobs) {
    J.o_occ = take((static_cast<size_t>(J.g.n_words) + 4) * sizeof(uint2));
    J.o_occd = take((static_cast<size_t>(J.g.n_words) + 4) * sizeof(uint32_t));
    J.o_rec = take(J.rec_cap * NDT2D_REC_DOUBLES * sizeof(double));
    J.o_recf = take(J.rec_cap * NDT2D_REC_DOUBLES * sizeof(double));
    J.o_recv = take(J.rec_cap * NDT2D_REC_DOUBLES * sizeof(double));
    J.o_nvalid = take(2 * sizeof(uint32_t));
    J.o_jobp = take(static_cast<size_t>(pl.n_jobs) * NDT2D_BLOCK_PARTIAL * sizeof(double));
    J.o_chunk = take(pl.chunk_doubles ? pl.chunk_doubles * sizeof(double) : 8);
  }
  const size_t o_results = take(n_jobs * 32 * sizeof(double));
  const size_t total_bytes = off;
  if (total_bytes > (size_t(2) << 30)) {return kBatchNotEligible;}
  // the arena may still feed an earlier asynchronous addScans: drain before reusing / growing it
  NDT2D_CUDA_TRY(cudaStreamSynchronize(m->stream));
  int rc = m->d_batch_arena.ensure(total_bytes);
  if (!rc) {rc = m->h_arena.ensure(upload_bytes);}
  if (!rc) {rc = m->h_result.ensure(std::max<size_t>(n_jobs * 32, 64) * sizeof(double));}
  if (rc) {return rc;}
  m->arena_off = 0;
  char * hb = m->h_arena.as<char>();
  char * db = m->d_batch_arena.as<char>();
  memset(hb + o_counter, 0, 64);
  BuildEntry * h_build = reinterpret_cast<BuildEntry *>(hb + o_build);
  BatchEntry * h_batch = reinterpret_cast<BatchEntry *>(hb + o_batch);
  std::vector<double> thr_x, thr_y;
  for (size_t j = 0; j < n_jobs; ++j) {
    const Job & J = jobs[j];
    // map side (as add_scans_locked)
    double4 * h_tf = reinterpret_cast<double4 *>(hb + J.o_tf);
    uint64_t * h_off = reinterpret_cast<uint64_t *>(hb + J.o_off);
    for (size_t k = 0; k < J.n_scans; ++k) {
      const double * pose = map_poses + 3 * (J.s0 + k);
      h_tf[k] = make_double4(pose[0], pose[1], cos(pose[2]), sin(pose[2]));
      h_off[k] = map_pt_offsets[J.s0 + k] - J.p0;
    }
    h_off[J.n_scans] = J.n_points;
    memcpy(hb + J.o_pts, map_pts_xy + 2 * J.p0, J.n_points * sizeof(double2));
    axis_thresholds(J.g.origin_x, J.g.cell_size, J.g.size_x, thr_x);
    axis_thresholds(J.g.origin_y, J.g.cell_size, J.g.size_y, thr_y);
    double * h_thr = reinterpret_cast<double *>(hb + J.o_thr);
    memcpy(h_thr, thr_x.data(), thr_x.size() * sizeof(double));
    memcpy(h_thr + thr_x.size(), thr_y.data(), thr_y.size() * sizeof(double));
    // query side (as stage_scan_locked)
    const double * pose3 = query_poses + 3 * j;
    subsample_points(query_pts_xy + 2 * J.q0, J.npts_q, J.n_use, reinterpret_cast<double *>(hb + J.o_q));
    double * h_trig = reinterpret_cast<double *>(hb + J.o_trig);
    for (size_t k = 0; k < n_ang; ++k) {
      h_trig[2 * k] = cos(pose3[2] + m->dth[k]);
      h_trig[2 * k + 1] = sin(pose3[2] + m->dth[k]);
    }
    BuildEntry & be = h_build[j];
    memset(&be, 0, sizeof(be));
    be.g = J.g;
    be.scan_tf = reinterpret_cast<const double4 *>(db + J.o_tf);
    be.offsets = reinterpret_cast<const uint64_t *>(db + J.o_off);
    be.pts = reinterpret_cast<const double2 *>(db + J.o_pts);
    be.occ = reinterpret_cast<uint2 *>(db + J.o_occ);
    be.occd = reinterpret_cast<uint32_t *>(db + J.o_occd);
    be.rec = reinterpret_cast<double *>(db + J.o_rec);
    be.rec_fast = reinterpret_cast<double *>(db + J.o_recf);
    be.rec_vtx = reinterpret_cast<double *>(db + J.o_recv);
    be.n_valid = reinterpret_cast<uint32_t *>(db + J.o_nvalid);
    be.n_scans = static_cast<uint32_t>(J.n_scans);
    be.n_points = static_cast<uint32_t>(J.n_points);
    be.rec_cap = static_cast<uint32_t>(J.rec_cap);
    BatchEntry & se = h_batch[j];
    memset(&se, 0, sizeof(se));
    se.mv.g = J.g;
    se.mv.occ = be.occ;
    se.mv.occ_dilated = be.occd;
    se.mv.rec = be.rec;
    se.mv.rec_fast = be.rec_fast;
    se.mv.rec_vtx = be.rec_vtx;
    se.mv.thr_x = reinterpret_cast<const double *>(db + J.o_thr);
    se.mv.thr_y = se.mv.thr_x + (J.g.size_x + 2);
    se.mv.n_valid_cap = static_cast<uint32_t>(J.rec_cap);
    se.mv.n_stiff = be.n_valid + 1;
    se.sv.pts = reinterpret_cast<const double2 *>(db + J.o_q);
    se.sv.trig = reinterpret_cast<const double2 *>(db + J.o_trig);
    se.sv.dth = m->d_dth.as<double>();
    se.sv.dlin = m->d_dlin.as<double>();
    se.sv.pose_x = pose3[0];
    se.sv.pose_y = pose3[1];
    se.sv.linear_res = m->prm.search_linear_resolution;
    se.sv.inv_linear_res = 1.0 / se.sv.linear_res;
    se.sv.n_pts = static_cast<uint32_t>(J.n_use);
    se.sv.n_ang = static_cast<uint32_t>(n_ang);
    se.sv.n_lin = static_cast<uint32_t>(n_lin);
    se.sv.theta_stride = 1;
    se.job_partials = reinterpret_cast<double *>(db + J.o_jobp);
    se.chunk_sums = reinterpret_cast<double *>(db + J.o_chunk);
  }
  cudaStream_t st = m->stream;
  NDT2D_CUDA_TRY(cudaMemcpyAsync(db, hb, upload_bytes, cudaMemcpyHostToDevice, st));
  m->ctr.h2d_bytes += upload_bytes;
  rc = ndt2d_launch_build_small_batch(reinterpret_cast<const BuildEntry *>(db + o_build),
      static_cast<uint32_t>(n_jobs), st, &m->ctr);
  if (!rc && dense && win_k) {
    rc = ndt2d_launch_search_window_batch(reinterpret_cast<const BatchEntry *>(db + o_batch),
        static_cast<uint32_t>(n_jobs), win_k, static_cast<uint32_t>(n_ang),
        static_cast<uint32_t>(n_lin), st, &m->ctr);
  } else if (!rc && dense) {
    rc = ndt2d_launch_search_dense_batch(reinterpret_cast<const BatchEntry *>(db + o_batch),
        static_cast<uint32_t>(n_jobs), static_cast<uint32_t>(n_ang), static_cast<uint32_t>(n_lin), st,
        &m->ctr);
  } else if (!rc) {
    rc = ndt2d_launch_search_region_batch(reinterpret_cast<const BatchEntry *>(db + o_batch),
        static_cast<uint32_t>(n_jobs), pl, reinterpret_cast<uint32_t *>(db + o_counter), st, &m->ctr);
  }
  if (!rc) {
    rc = ndt2d_launch_finish_batch(reinterpret_cast<const BatchEntry *>(db + o_batch),
        static_cast<uint32_t>(n_jobs), pl.n_jobs, static_cast<double>(n_ang) * n_lin * n_lin,
        reinterpret_cast<double *>(db + o_results), reinterpret_cast<uint32_t *>(db + o_counter), st,
        &m->ctr);
  }
  m->has_model = false;
  m->staged = false;
  m->ev_valid = false;
  if (rc) {
    cudaStreamSynchronize(st);
    return rc;
  }
  NDT2D_CUDA_TRY(cudaMemcpyAsync(m->h_result.p, db + o_results, n_jobs * 32 * sizeof(double),
    cudaMemcpyDeviceToHost, st));
  NDT2D_CUDA_TRY(cudaStreamSynchronize(st));
  m->ctr.d2h_bytes += n_jobs * 32 * sizeof(double);
  for (size_t j = 0; j < n_jobs; ++j) {
    if (delta_written) {delta_written[j] = 0;}
    unpack_result(m->h_result.as<double>() + 32 * j, out_delta3 ? out_delta3 + 3 * j : nullptr,
      delta_written ? delta_written + j : nullptr, out_cov9 ? out_cov9 + 9 * j : nullptr,
      out_score ? out_score + j : nullptr);
  }
  return NDT2D_OK;
}

static int match_scan_batch_locked(
  ndt2d_matcher * m, size_t n_jobs,
  const uint64_t * job_scan_offsets, const double * map_poses, const uint64_t * map_pt_offsets,
  const double * map_pts_xy,
  const double * query_poses, const uint64_t * query_pt_offsets, const double * query_pts_xy,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score)
{
  if (n_jobs == 0) {
    m->has_model = false;
    return NDT2D_OK;
  }
  {
    const int frc = match_scan_batch_fused(m, n_jobs, job_scan_offsets, map_poses, map_pt_offsets,
        map_pts_xy, query_poses, query_pt_offsets, query_pts_xy, out_delta3, delta_written, out_cov9,
        out_score);
    if (frc != kBatchNotEligible) {return frc;}
  }
  // Jobs too large for the single-launch path: pipelined over kBatchLanes internal sub-handles (own stream, own model + scratch
  // buffers): job j's build + search is enqueued on lane j % kBatchLanes without waiting
  // for the device (host staging from that lane's pinned arena), so the dependent chains
  // of small kernels of different jobs overlap; every lane collects its 32-double result
  // records on the device and they are fetched with one copy per lane at the end.
  const size_t n_lanes = std::min<size_t>(kBatchLanes, n_jobs);
  while (m->lanes.size() < n_lanes) {
    ndt2d_params p = m->prm;
    p.device = m->device;
    p.stream = nullptr;
    ndt2d_matcher * sub = nullptr;
    const int rc0 = ndt2d_matcher_create(&p, &sub);
    if (rc0) {return rc0;}
    m->lanes.push_back(sub);
  }
  int rc = NDT2D_OK;
  const size_t per_lane = (n_jobs + n_lanes - 1) / n_lanes;
  for (size_t l = 0; l < n_lanes && !rc; ++l) {
    ndt2d_matcher * s = m->lanes[l];
    rc = s->h_arena.ensure(size_t(4) << 20);
    if (!rc) {rc = s->d_batch_results.ensure(per_lane * 32 * sizeof(double));}
    if (!rc) {rc = s->h_result.ensure(std::max<size_t>(per_lane * 32, 64) * sizeof(double));}
    s->pipelined = true;
    s->arena_off = 0;
  }
  for (size_t j = 0; j < n_jobs && !rc; ++j) {
    ndt2d_matcher * s = m->lanes[j % n_lanes];
    const size_t slot = j / n_lanes;
    const uint64_t s0 = job_scan_offsets[j], s1 = job_scan_offsets[j + 1];
    rc = add_scans_locked(s, static_cast<size_t>(s1 - s0), map_poses + 3 * s0,
        map_pt_offsets + s0, map_pts_xy);
    if (rc) {break;}
    const uint64_t q0 = query_pt_offsets[j], q1 = query_pt_offsets[j + 1];
    rc = stage_scan_locked(s, query_poses + 3 * j, query_pts_xy + 2 * q0,
        static_cast<size_t>(q1 - q0));
    if (rc) {break;}
    // the search's finish kernel writes the job's 32-double record straight into its slot
    rc = ndt2d_launch_search(model_view(s), search_view(s), 0,
        static_cast<uint32_t>(s->dth.size()), m->prm.kernel_variant, s->d_blockpart.as<double>(),
        s->d_batch_results.as<double>() + 32 * slot, nullptr, s->d_counter.as<uint32_t>(),
        s->stream, &m->ctr);
    if (rc) {break;}
  }
  for (size_t l = 0; l < n_lanes; ++l) {
    ndt2d_matcher * s = m->lanes[l];
    s->pipelined = false;
    s->has_model = false;
    s->staged = false;
    m->ctr.launches += s->ctr.launches;
    m->ctr.h2d_bytes += s->ctr.h2d_bytes;
    s->ctr = Counters{0, 0, 0};
  }
  m->has_model = false;
  m->staged = false;
  m->ev_valid = false;
  if (rc) {
    for (size_t l = 0; l < n_lanes; ++l) {cudaStreamSynchronize(m->lanes[l]->stream);}
    return rc;
  }
  for (size_t l = 0; l < n_lanes; ++l) {
    ndt2d_matcher * s = m->lanes[l];
    const size_t mine = (n_jobs - l + n_lanes - 1) / n_lanes;   // jobs l, l + n_lanes, ...
    NDT2D_CUDA_TRY(cudaMemcpyAsync(s->h_result.p, s->d_batch_results.p, mine * 32 * sizeof(double),
      cudaMemcpyDeviceToHost, s->stream));
  }
  for (size_t l = 0; l < n_lanes; ++l) {
    NDT2D_CUDA_TRY(cudaStreamSynchronize(m->lanes[l]->stream));
  }
  m->ctr.d2h_bytes += n_jobs * 32 * sizeof(double);
  for (size_t j = 0; j < n_jobs; ++j) {
    const double * r32 = m->lanes[j % n_lanes]->h_result.as<double>() + 32 * (j / n_lanes);
    if (delta_written) {delta_written[j] = 0;}
    unpack_result(r32, out_delta3 ? out_delta3 + 3 * j : nullptr,
      delta_written ? delta_written + j : nullptr, out_cov9 ? out_cov9 + 9 * j : nullptr,
      out_score ? out_score + j : nullptr);
  }
  return NDT2D_OK;
}

extern "C" {

// Mapper::loopClosureThread's inner loop (ndt_mapper.cpp:619-671) with its exact sequential
// semantics, evaluated speculatively: all remaining candidates are matched in one batch with
// the query scan's CURRENT pose; results are consumed in order up to and including the first
// acceptance (finite score < typical_response, :645), which moves the scan pose (:652-655) and
// thereby invalidates the later results of that batch -- the remainder is re-issued.
NDT2D_API int ndt2d_matcher_close_loop(
  ndt2d_matcher * m, size_t n_scans, const double * scan_poses, const uint64_t * scan_pt_offsets,
  const double * scan_pts_xy, const uint64_t * candidates, size_t n_candidates, size_t rolling,
  size_t search_limit, double typical_response, double * query_pose3, const double * query_pts_xy,
  size_t query_npts, uint64_t * out_candidate, double * out_score, int * out_accepted,
  double * out_pose3, double * out_cov9, size_t * n_processed, size_t * n_batches)
{
  if (!m || !query_pose3 || !n_processed || (n_scans && (!scan_poses || !scan_pt_offsets)) ||
    (n_candidates && !candidates) || (query_npts && !query_pts_xy))
  {
    return NDT2D_ERR_INVALID;
  }
  *n_processed = 0;
  if (n_batches) {*n_batches = 0;}
  std::lock_guard<std::mutex> lock(m->mu);
  DeviceGuard guard(m->device);
  // candidates the reference would look at, in order: empty scans are skipped without
  // counting (:625), at most search_limit are processed (:671) -- the reference counts down
  // `--num_scans_to_check == 0` on a size_t, so a limit of 0 wraps and means "no limit"
  const size_t limit = search_limit == 0 ? std::numeric_limits<size_t>::max() : search_limit;
  std::vector<uint64_t> todo;
  for (size_t k = 0; k < n_candidates && todo.size() < limit; ++k) {
    const uint64_t i = candidates[k];
    if (i >= n_scans) {return NDT2D_ERR_INVALID;}
    if (scan_pt_offsets[i + 1] == scan_pt_offsets[i]) {continue;}
    todo.push_back(i);
  }
  size_t next = 0;
  std::vector<uint64_t> job_offsets, q_offsets;
  std::vector<double> q_poses, delta, cov, score;
  std::vector<int> written;
  std::vector<size_t> slot_of;   // candidate of the round -> job of the batch (npos: empty window)
  constexpr size_t npos = std::numeric_limits<size_t>::max();
  while (next < todo.size()) {
    const size_t n_round = todo.size() - next;
    // job j: window [i-1 (or i), i+1 if i < rolling else i) of the graph's scans (:628-631);
    // the windows are contiguous scan ranges, so they index the caller's arrays directly.  A
    // candidate whose window is empty (i == 0 with rolling == 0) has no map: it scores 0.0 like
    // matchScan without a model (scan_matcher_ndt.cpp:80) and is never accepted -- the batch
    // goes on (the reference itself would size an NDT from an empty bounding box there).
    job_offsets.assign(1, 0);
    slot_of.assign(n_round, npos);
    std::vector<double> w_poses;
    std::vector<uint64_t> w_offsets(1, 0);
    std::vector<double> w_points;
    size_t n_jobs = 0;
    for (size_t j = 0; j < n_round; ++j) {
      const uint64_t i = todo[next + j];
      const uint64_t begin = i > 0 ? i - 1 : i, end = i < rolling ? i + 1 : i;
      if (end <= begin) {continue;}
      for (uint64_t s = begin; s < end; ++s) {
        w_poses.insert(w_poses.end(), scan_poses + 3 * s, scan_poses + 3 * s + 3);
        w_points.insert(w_points.end(), scan_pts_xy + 2 * scan_pt_offsets[s],
          scan_pts_xy + 2 * scan_pt_offsets[s + 1]);
        w_offsets.push_back(w_points.size() / 2);
      }
      job_offsets.push_back(w_poses.size() / 3);
      slot_of[j] = n_jobs++;
    }
    q_poses.resize(3 * n_jobs);
    q_offsets.resize(n_jobs + 1);
    std::vector<double> q_points(2 * query_npts * n_jobs);
    for (size_t j = 0; j < n_jobs; ++j) {
      memcpy(&q_poses[3 * j], query_pose3, 3 * sizeof(double));
      q_offsets[j] = j * query_npts;
      if (query_npts) {
        memcpy(&q_points[2 * query_npts * j], query_pts_xy, 2 * query_npts * sizeof(double));
      }
    }
    q_offsets[n_jobs] = n_jobs * query_npts;
    delta.assign(3 * n_jobs, 0.0);
    cov.assign(9 * n_jobs, 0.0);
    score.assign(n_jobs, 0.0);
    written.assign(n_jobs, 0);
    if (n_jobs) {
      const int rc = match_scan_batch_locked(m, n_jobs, job_offsets.data(), w_poses.data(),
          w_offsets.data(), w_points.data(), q_poses.data(), q_offsets.data(), q_points.data(),
          delta.data(), written.data(), cov.data(), score.data());
      if (rc) {return rc;}
      if (n_batches) {*n_batches += 1;}
    }
    // consume in order up to the first acceptance
    size_t jr = 0;
    for (; jr < n_round; ++jr) {
      const size_t o = *n_processed;
      const size_t j = slot_of[jr];
      if (j == npos) {
        if (out_candidate) {out_candidate[o] = todo[next + jr];}
        if (out_score) {out_score[o] = 0.0;}
        if (out_accepted) {out_accepted[o] = 0;}
        if (out_pose3) {memcpy(out_pose3 + 3 * o, query_pose3, 3 * sizeof(double));}
        if (out_cov9) {memset(out_cov9 + 9 * o, 0, 9 * sizeof(double));}
        *n_processed = o + 1;
        continue;
      }
      const bool accept = std::isfinite(score[j]) && score[j] < typical_response;  // :645
      if (accept) {
        // correction += scan pose; scan->setPose(correction)  (:652-655); an unwritten
        // correction is the default Pose2d (0, 0, 0)
        query_pose3[0] = (written[j] ? delta[3 * j] : 0.0) + query_pose3[0];
        query_pose3[1] = (written[j] ? delta[3 * j + 1] : 0.0) + query_pose3[1];
        query_pose3[2] = (written[j] ? delta[3 * j + 2] : 0.0) + query_pose3[2];
      }
      if (out_candidate) {out_candidate[o] = todo[next + jr];}
      if (out_score) {out_score[o] = score[j];}
      if (out_accepted) {out_accepted[o] = accept ? 1 : 0;}
      if (out_pose3) {memcpy(out_pose3 + 3 * o, query_pose3, 3 * sizeof(double));}
      if (out_cov9) {memcpy(out_cov9 + 9 * o, &cov[9 * j], 9 * sizeof(double));}
      *n_processed = o + 1;
      if (accept) {
        ++jr;
        break;
      }
    }
    next += jr;
  }
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_search_shape(
  const ndt2d_matcher * m, uint64_t * n_angular, uint64_t * n_linear)
{
  if (!m) {return NDT2D_ERR_INVALID;}
  if (n_angular) {*n_angular = m->dth.size();}
  if (n_linear) {*n_linear = m->dlin.size();}
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_search_values(const ndt2d_matcher * m, double * dth, double * dlin)
{
  if (!m) {return NDT2D_ERR_INVALID;}
  if (dth) {memcpy(dth, m->dth.data(), m->dth.size() * sizeof(double));}
  if (dlin) {memcpy(dlin, m->dlin.data(), m->dlin.size() * sizeof(double));}
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_stage_scan(
  ndt2d_matcher * m, const double * pose3, const double * pts_xy, size_t npts)
{
  if (!m || !pose3 || (npts && !pts_xy)) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  DeviceGuard guard(m->device);
  return stage_scan_locked(m, pose3, pts_xy, npts, true);
}

NDT2D_API int ndt2d_matcher_search_staged_strided(
  ndt2d_matcher * m, uint64_t theta_begin, uint64_t theta_end, uint64_t theta_stride,
  void * d_partial)
{
  if (!m || theta_stride == 0 || theta_stride > 0xffffffffull) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  if (!m->has_model) {return NDT2D_ERR_NO_MAP;}
  if (!m->staged) {return NDT2D_ERR_STATE;}
  const uint64_t n_ang = m->dth.size();
  if (theta_begin > theta_end || theta_end > n_ang) {return NDT2D_ERR_INVALID;}
  DeviceGuard guard(m->device);
  int rc = ensure_trig_locked(m, theta_begin, theta_end, theta_stride);
  if (rc) {return rc;}
  SearchView sv = search_view(m);
  sv.theta_stride = static_cast<uint32_t>(theta_stride);
  rc = ndt2d_launch_search(model_view(m), sv, static_cast<uint32_t>(theta_begin),
      static_cast<uint32_t>(theta_end), m->prm.kernel_variant, m->d_blockpart.as<double>(),
      m->d_partial.as<double>(), nullptr, m->d_counter.as<uint32_t>(), m->stream, &m->ctr,
      m->ev_begin, m->ev_end);
  if (rc) {return rc;}
  m->ev_valid = theta_end > theta_begin;
  if (d_partial) {
    NDT2D_CUDA_TRY(cudaMemcpyAsync(d_partial, m->d_partial.p,
      NDT2D_PARTIAL_DOUBLES * sizeof(double), cudaMemcpyDeviceToDevice, m->stream));
  }
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_search_staged(
  ndt2d_matcher * m, uint64_t theta_begin, uint64_t theta_end, void * d_partial)
{
  return ndt2d_matcher_search_staged_strided(m, theta_begin, theta_end, 1, d_partial);
}

NDT2D_API int ndt2d_matcher_exchange_init(
  ndt2d_matcher * m, uint32_t world, uint32_t rank, unsigned char * handle64)
{
  if (!m || !handle64 || world == 0 || world > kExchangeMaxRanks || rank >= world) {
    return NDT2D_ERR_INVALID;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::lock_guard<std::mutex> lock(m->mu);
  DeviceGuard guard(m->device);
  exchange_close(m);
  int rc = m->d_mailbox.ensure(kExchangeMailboxBytes);
  if (!rc) {rc = m->d_peer_table.ensure(kExchangeMaxRanks * sizeof(void *));}
  if (rc) {return rc;}
  NDT2D_CUDA_TRY(cudaMemsetAsync(m->d_mailbox.p, 0, kExchangeMailboxBytes, m->stream));
  NDT2D_CUDA_TRY(cudaStreamSynchronize(m->stream));  // cleared before any peer can learn the handle
  cudaIpcMemHandle_t h;
  NDT2D_CUDA_TRY(cudaIpcGetMemHandle(&h, m->d_mailbox.p));
  memcpy(handle64, &h, 64);
  m->x_world = world;
  m->x_rank = rank;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_exchange_connect(ndt2d_matcher * m, const unsigned char * handles)
{
  if (!m || !handles) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  if (m->x_world == 0 || !m->d_mailbox.p) {return NDT2D_ERR_STATE;}
  DeviceGuard guard(m->device);
  exchange_close(m);
  m->peer_ptrs.assign(m->x_world, nullptr);
  for (uint32_t r = 0; r < m->x_world; ++r) {
    if (r == m->x_rank) {
      m->peer_ptrs[r] = m->d_mailbox.p;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * static_cast<size_t>(r), 64);
    void * p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      ndt2d_set_error("cudaIpcOpenMemHandle", e, __FILE__, __LINE__);
      exchange_close(m);
      return NDT2D_ERR_CUDA;
    }
    m->peer_ptrs[r] = p;
  }
  NDT2D_CUDA_TRY(cudaMemcpy(m->d_peer_table.p, m->peer_ptrs.data(), m->x_world * sizeof(void *),
    cudaMemcpyHostToDevice));
  m->x_connected = true;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_search_exchange(
  ndt2d_matcher * m, uint64_t theta_begin, uint64_t theta_end, uint64_t theta_stride, uint64_t seq)
{
  if (!m || theta_stride == 0 || theta_stride > 0xffffffffull || seq == 0) {
    return NDT2D_ERR_INVALID;
  }
  std::lock_guard<std::mutex> lock(m->mu);
  if (!m->has_model) {return NDT2D_ERR_NO_MAP;}
  if (!m->staged || !m->x_connected) {return NDT2D_ERR_STATE;}
  const uint64_t n_ang = m->dth.size();
  if (theta_begin > theta_end || theta_end > n_ang) {return NDT2D_ERR_INVALID;}
  DeviceGuard guard(m->device);
  {
    const int trc = ensure_trig_locked(m, theta_begin, theta_end, theta_stride);
    if (trc) {return trc;}
  }
  SearchView sv = search_view(m);
  sv.theta_stride = static_cast<uint32_t>(theta_stride);
  ExchangeView xv;
  xv.peers = m->d_peer_table.as<void *>();
  xv.world = m->x_world;
  xv.rank = m->x_rank;
  xv.seq = seq;
  xv.timeout_ns = 10ull * 1000ull * 1000ull * 1000ull;
  const int rc = ndt2d_launch_search(model_view(m), sv, static_cast<uint32_t>(theta_begin),
      static_cast<uint32_t>(theta_end), m->prm.kernel_variant, m->d_blockpart.as<double>(),
      m->d_partial.as<double>(), nullptr, m->d_counter.as<uint32_t>(), m->stream, &m->ctr,
      m->ev_begin, m->ev_end, &xv);
  if (rc) {return rc;}
  m->ev_valid = theta_end > theta_begin;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_fetch_result(
  ndt2d_matcher * m, double * out_delta3, int * delta_written, double * out_cov9,
  double * out_score)
{
  if (!m) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  if (!m->staged) {return NDT2D_ERR_STATE;}
  DeviceGuard guard(m->device);
  double r32[32];
  const int rc = fetch_result_locked(m, r32);
  if (rc) {return rc;}
  if (delta_written) {*delta_written = 0;}
  unpack_result(r32, out_delta3, delta_written, out_cov9, out_score);
  if (r32[31] != 0.0) {
    snprintf(g_last_error, sizeof(g_last_error),
      "fused exchange timed out: not every rank published its partial record");
    return NDT2D_ERR_STATE;
  }
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_fetch_partial(ndt2d_matcher * m, double * partial16)
{
  if (!m || !partial16) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  if (!m->staged) {return NDT2D_ERR_STATE;}
  DeviceGuard guard(m->device);
  double r32[32];
  int rc = fetch_result_locked(m, r32);
  if (rc) {return rc;}
  memcpy(partial16, r32, NDT2D_PARTIAL_DOUBLES * sizeof(double));
  return NDT2D_OK;
}

NDT2D_API int ndt2d_search_lattice(double size, double resolution, double * out, size_t cap,
  size_t * n)
{
  // `for (v = -size; v < size; v += resolution)` replayed (scan_matcher_ndt.cpp:103,117,119)
  if (!n) {return NDT2D_ERR_INVALID;}
  std::vector<double> v;
  const int rc = replay_loop(size, resolution, 1u << 24, v);
  if (rc) {return rc;}
  *n = v.size();
  if (out) {
    if (cap < v.size()) {return NDT2D_ERR_SIZE;}
    memcpy(out, v.data(), v.size() * sizeof(double));
  }
  return NDT2D_OK;
}

NDT2D_API int ndt2d_combine_partials_host(
  const double * dth, size_t n_ang, const double * dlin, size_t n_lin, const double * partials,
  size_t n_partials, double * out_delta3, int * delta_written, double * out_cov9,
  double * out_score)
{
  if ((n_partials && !partials) || (n_ang && !dth) || (n_lin && !dlin)) {return NDT2D_ERR_INVALID;}
  double best = 0.0, best_idx = 1.0e300, s[10] = {0}, npts = 0.0;
  for (size_t r = 0; r < n_partials; ++r) {
    const double * p = partials + r * NDT2D_PARTIAL_DOUBLES;
    if (p[0] < best || (p[0] == best && p[1] < best_idx)) {
      best = p[0];
      best_idx = p[1];
    }
    for (int k = 0; k < 10; ++k) {s[k] += p[2 + k];}
    npts = std::max(npts, p[13]);
  }
  const bool written = best < 0.0;
  if (delta_written) {*delta_written = written ? 1 : 0;}
  if (written && out_delta3) {
    const uint64_t n_cand = static_cast<uint64_t>(n_lin) * n_lin;
    const uint64_t idx = static_cast<uint64_t>(best_idx);
    const uint64_t it = idx / n_cand, rem = idx - it * n_cand;
    if (it >= n_ang) {return NDT2D_ERR_INVALID;}
    out_delta3[0] = dlin[rem / n_lin];
    out_delta3[1] = dlin[rem % n_lin];
    out_delta3[2] = dth[it];
  }
  if (out_cov9) {
    // covariance = (1/s) k + ((1/(s*s)) u) u^T      (scan_matcher_ndt.cpp:146)
    const double sum = s[9], inv_s = 1.0 / sum, inv_s2 = 1.0 / (sum * sum);
    const double k[9] = {s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5]};
    const double u[3] = {s[6], s[7], s[8]};
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) {
        out_cov9[r * 3 + c] = inv_s * k[r * 3 + c] + (inv_s2 * u[r]) * u[c];
      }
    }
  }
  if (out_score) {*out_score = (written ? best : 0.0) / npts;}
  return NDT2D_OK;
}

NDT2D_API int ndt2d_combine_partials(
  const ndt2d_matcher * m, const double * partials, size_t n_partials,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score)
{
  if (!m) {return NDT2D_ERR_INVALID;}
  return ndt2d_combine_partials_host(m->dth.data(), m->dth.size(), m->dlin.data(), m->dlin.size(),
           partials, n_partials, out_delta3, delta_written, out_cov9, out_score);
}

NDT2D_API int ndt2d_matcher_combine_device(
  ndt2d_matcher * m, const void * d_partials, size_t n_partials,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score)
{
  if (!m || !d_partials || n_partials == 0) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  DeviceGuard guard(m->device);
  int rc = m->d_partial.ensure(32 * sizeof(double));
  if (!rc) {rc = m->h_result.ensure(64 * sizeof(double));}
  if (rc) {return rc;}
  rc = ndt2d_launch_combine(static_cast<const double *>(d_partials),
      static_cast<uint32_t>(n_partials), m->d_dth.as<double>(), m->d_dlin.as<double>(),
      static_cast<uint32_t>(m->dlin.size()), m->d_partial.as<double>(), m->stream, &m->ctr);
  if (rc) {return rc;}
  double r32[32];
  if ((rc = fetch_result_locked(m, r32))) {return rc;}
  if (delta_written) {*delta_written = 0;}
  unpack_result(r32, out_delta3, delta_written, out_cov9, out_score);
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_grid_info(ndt2d_matcher * m, double * info5)
{
  if (!m || !info5) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  if (!m->has_model) {return NDT2D_ERR_NO_MAP;}
  info5[0] = m->g.size_x;
  info5[1] = m->g.size_y;
  info5[2] = m->g.origin_x;
  info5[3] = m->g.origin_y;
  info5[4] = m->g.cell_size;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_dump_cells(ndt2d_matcher * m, double * out)
{
  if (!m || !out) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  if (!m->has_model) {return NDT2D_ERR_NO_MAP;}
  DeviceGuard guard(m->device);
  const size_t bytes = static_cast<size_t>(m->g.n_cells) * 16 * sizeof(double);
  DeviceBuffer tmp;
  int rc = tmp.ensure(bytes ? bytes : 8);
  if (rc) {return rc;}
  rc = ndt2d_launch_dump_cells(m->g, m->bs, m->sorted_buf, m->n_map_points, tmp.as<double>(),
      m->stream, &m->ctr);
  if (!rc && bytes) {
    cudaError_t e = cudaMemcpyAsync(out, tmp.p, bytes, cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) {e = cudaStreamSynchronize(m->stream);}
    if (e != cudaSuccess) {
      ndt2d_set_error("dump_cells copy", e, __FILE__, __LINE__);
      rc = NDT2D_ERR_CUDA;
    }
    m->ctr.d2h_bytes += bytes;
  }
  tmp.release();
  return rc;
}

NDT2D_API int ndt2d_matcher_dump_keys(ndt2d_matcher * m, int32_t * out, size_t n_points)
{
  if (!m || !out) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  if (!m->has_model) {return NDT2D_ERR_NO_MAP;}
  if (n_points != m->n_map_points) {return NDT2D_ERR_INVALID;}
  if (n_points == 0) {return NDT2D_OK;}
  DeviceGuard guard(m->device);
  // sorted (key, point index) pairs -> keys in add_scans order
  std::vector<uint32_t> keys(n_points), vals(n_points);
  NDT2D_CUDA_TRY(cudaMemcpyAsync(keys.data(), m->bs.key[m->sorted_buf],
    n_points * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
  NDT2D_CUDA_TRY(cudaMemcpyAsync(vals.data(), m->bs.val[m->sorted_buf],
    n_points * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
  NDT2D_CUDA_TRY(cudaStreamSynchronize(m->stream));
  m->ctr.d2h_bytes += 2 * n_points * sizeof(uint32_t);
  for (size_t i = 0; i < n_points; ++i) {
    out[vals[i]] = keys[i] >= m->g.n_cells ? -1 : static_cast<int32_t>(keys[i]);
  }
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_dump_scores(
  ndt2d_matcher * m, const double * pose3, const double * pts_xy, size_t npts, double * out,
  size_t n_out)
{
  if (!m || !pose3 || !out || (npts && !pts_xy)) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  if (!m->has_model) {return NDT2D_ERR_NO_MAP;}
  const size_t n_ang = m->dth.size(), n_lin = m->dlin.size();
  if (n_out != n_ang * n_lin * n_lin) {return NDT2D_ERR_INVALID;}
  if (n_out == 0) {return NDT2D_OK;}
  DeviceGuard guard(m->device);
  int rc = stage_scan_locked(m, pose3, pts_xy, npts);
  if (rc) {return rc;}
  DeviceBuffer tmp;
  if ((rc = tmp.ensure(n_out * sizeof(double)))) {return rc;}
  rc = ndt2d_launch_search(model_view(m), search_view(m), 0, static_cast<uint32_t>(n_ang),
      m->prm.kernel_variant, m->d_blockpart.as<double>(), m->d_partial.as<double>(),
      tmp.as<double>(), m->d_counter.as<uint32_t>(), m->stream, &m->ctr);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(out, tmp.p, n_out * sizeof(double), cudaMemcpyDeviceToHost,
        m->stream);
    if (e == cudaSuccess) {e = cudaStreamSynchronize(m->stream);}
    if (e != cudaSuccess) {
      ndt2d_set_error("dump_scores copy", e, __FILE__, __LINE__);
      rc = NDT2D_ERR_CUDA;
    }
    m->ctr.d2h_bytes += n_out * sizeof(double);
  }
  tmp.release();
  return rc;
}

NDT2D_API int ndt2d_matcher_counters(ndt2d_matcher * m, uint64_t * out4)
{
  if (!m || !out4) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  out4[0] = m->ctr.launches;
  out4[1] = m->ctr.h2d_bytes;
  out4[2] = m->ctr.d2h_bytes;
  out4[3] = 0;
  if (m->has_model) {
    DeviceGuard guard(m->device);
    uint32_t nv = 0;
    cudaStreamSynchronize(m->stream);   // the build may still be running (asynchronous addScans)
    if (cudaMemcpy(&nv, m->d_nvalid.p, sizeof(nv), cudaMemcpyDeviceToHost) == cudaSuccess) {
      out4[3] = nv;
    }
  }
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_search_stats(ndt2d_matcher * m, uint64_t * out4)
{
  if (!m || !out4) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  out4[0] = out4[1] = out4[2] = out4[3] = 0;
  if (!m->d_counter.p) {return NDT2D_ERR_STATE;}
  DeviceGuard guard(m->device);
  uint64_t h[6] = {0, 0, 0, 0, 0, 0};
  NDT2D_CUDA_TRY(cudaStreamSynchronize(m->stream));
  NDT2D_CUDA_TRY(cudaMemcpy(h, m->d_counter.p, sizeof(h), cudaMemcpyDeviceToHost));
  // the finish kernel moved the launch's tallies to the "last search" slots [3..5]
  out4[0] = h[3];                  // useful evaluations
  out4[1] = h[4];                  // (point, region) items
  out4[2] = h[5] & 0xffffffffu;    // job counter at exit (>= jobs)
  if (m->ev_valid && m->ev_begin && m->ev_end) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, m->ev_begin, m->ev_end) == cudaSuccess) {
      out4[3] = static_cast<uint64_t>(static_cast<double>(ms) * 1.0e6);  // search kernel, ns
    } else {
      cudaGetLastError();
    }
  }
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_build_stats(ndt2d_matcher * m, uint64_t * out4)
{
  if (!m || !out4) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  out4[0] = out4[1] = out4[2] = out4[3] = 0;
  if (!m->has_model) {return NDT2D_ERR_NO_MAP;}
  DeviceGuard guard(m->device);
  NDT2D_CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (m->evb_valid && m->evb_begin && m->evb_end) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, m->evb_begin, m->evb_end) == cudaSuccess) {
      out4[0] = static_cast<uint64_t>(static_cast<double>(ms) * 1.0e6);
    } else {
      cudaGetLastError();
    }
  }
  out4[1] = m->n_map_points;
  uint32_t nv = 0;
  NDT2D_CUDA_TRY(cudaMemcpy(&nv, m->d_nvalid.p, sizeof(nv), cudaMemcpyDeviceToHost));
  out4[2] = nv;
  out4[3] = static_cast<uint64_t>(m->g.size_x) * m->g.size_y;
  return NDT2D_OK;
}

NDT2D_API void * ndt2d_matcher_stream(ndt2d_matcher * m)
{
  return m ? static_cast<void *>(m->stream) : nullptr;
}

}  // extern "C"

// ---------------------------------------------------------------- filter
struct ndt2d_filter
{
  std::mutex mu;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  Counters ctr{0, 0, 0};
  size_t min_particles = 0, max_particles = 0;
  uint32_t n = 0;
  int cur = 0;
  DeviceBuffer d_particles[2], d_weights[2], d_stats, d_cdf, d_draws, d_first, d_canon, d_table,
    d_new_n, d_uniforms, d_pose_tf, d_pts;
  uint32_t table_size = 0;
  PinnedBuffer h_stage;
};

namespace
{

FilterView filter_view(ndt2d_filter * f)
{
  FilterView v;
  v.particles = f->d_particles[f->cur].as<double>();
  v.weights = f->d_weights[f->cur].as<double>();
  v.stats = f->d_stats.as<double>();
  return v;
}

int filter_reserve(ndt2d_filter * f, size_t n)
{
  int rc = 0;
  const size_t cap = std::max<size_t>(std::max(n, f->max_particles), 1);
  for (int b = 0; b < 2 && !rc; ++b) {
    rc = f->d_particles[b].ensure(cap * 3 * sizeof(double));
    if (!rc) {rc = f->d_weights[b].ensure(cap * sizeof(double));}
  }
  if (!rc) {rc = f->d_cdf.ensure(cap * sizeof(double));}
  if (!rc) {rc = f->d_draws.ensure(cap * sizeof(uint32_t));}
  if (!rc) {rc = f->d_first.ensure(cap * sizeof(uint32_t));}
  if (!rc) {rc = f->d_canon.ensure(cap * sizeof(uint32_t));}
  if (!rc) {rc = f->d_pose_tf.ensure(cap * sizeof(double4));}
  if (!rc) {rc = f->d_uniforms.ensure(cap * sizeof(double));}
  uint32_t ts = 64;
  while (ts < 2 * cap) {ts <<= 1;}
  if (!rc) {rc = f->d_table.ensure(static_cast<size_t>(ts) * sizeof(uint32_t));}
  f->table_size = ts;
  if (!rc) {rc = f->h_stage.ensure(cap * 4 * sizeof(double) + 256);}
  return rc;
}

// When the current buffers had to grow the contents are lost: callers that
// need them re-upload.  Used only by set_particles (which overwrites anyway).

}  // namespace

extern "C" {

NDT2D_API int ndt2d_filter_create(
  size_t min_particles, size_t max_particles, int device, void * stream, ndt2d_filter ** out)
{
  if (!out || max_particles == 0 || max_particles >= (1u << 28) ||
    min_particles >= (1u << 28))
  {
    return NDT2D_ERR_INVALID;
  }
  *out = nullptr;
  if (ndt2d_device_count() <= 0) {
    snprintf(g_last_error, sizeof(g_last_error),
      "no CUDA device visible: libndt2d_b200 has no CPU fallback");
    return NDT2D_ERR_NO_DEVICE;
  }
  int dev = device;
  if (dev < 0) {NDT2D_CUDA_TRY(cudaGetDevice(&dev));}
  ndt2d_filter * f = new (std::nothrow) ndt2d_filter();
  if (!f) {return NDT2D_ERR_INVALID;}
  f->device = dev;
  f->min_particles = min_particles;
  f->max_particles = max_particles;
  DeviceGuard guard(dev);
  if (stream) {
    f->stream = static_cast<cudaStream_t>(stream);
  } else {
    if (cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete f;
      return NDT2D_ERR_CUDA;
    }
    f->own_stream = true;
  }
  int rc = filter_reserve(f, std::max(min_particles, max_particles));
  if (!rc) {rc = f->d_stats.ensure(12 * sizeof(double));}
  if (!rc) {rc = f->d_new_n.ensure(sizeof(uint32_t));}
  if (rc) {
    ndt2d_filter_destroy(f);
    return rc;
  }
  // particle_filter.cpp:42-50: mean/cov zero, min_particles at the origin
  cudaMemsetAsync(f->d_stats.p, 0, 12 * sizeof(double), f->stream);
  f->n = static_cast<uint32_t>(min_particles);
  if (f->n) {
    double * h = f->h_stage.as<double>();
    for (uint32_t i = 0; i < f->n; ++i) {h[i] = 1.0 / static_cast<double>(min_particles);}
    cudaMemsetAsync(f->d_particles[0].p, 0, static_cast<size_t>(f->n) * 3 * sizeof(double),
      f->stream);
    cudaMemcpyAsync(f->d_weights[0].p, h, f->n * sizeof(double), cudaMemcpyHostToDevice, f->stream);
    rc = ndt2d_launch_filter_stats(filter_view(f), f->n, f->stream, &f->ctr);
  }
  if (cudaStreamSynchronize(f->stream) != cudaSuccess) {rc = NDT2D_ERR_CUDA;}
  if (rc) {
    ndt2d_filter_destroy(f);
    return rc;
  }
  *out = f;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_filter_destroy(ndt2d_filter * f)
{
  if (!f) {return NDT2D_OK;}
  {
    DeviceGuard guard(f->device);
    if (f->stream) {cudaStreamSynchronize(f->stream);}
    DeviceBuffer * bufs[] = {&f->d_particles[0], &f->d_particles[1], &f->d_weights[0],
      &f->d_weights[1], &f->d_stats, &f->d_cdf, &f->d_draws, &f->d_first, &f->d_canon,
      &f->d_table, &f->d_new_n, &f->d_uniforms, &f->d_pose_tf, &f->d_pts};
    for (DeviceBuffer * b : bufs) {b->release();}
    f->h_stage.release();
    if (f->own_stream && f->stream) {cudaStreamDestroy(f->stream);}
  }
  delete f;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_filter_set_particles(
  ndt2d_filter * f, const double * particles3, const double * weights, size_t n)
{
  if (!f || (n && (!particles3 || !weights)) || n >= (1u << 28)) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(f->mu);
  DeviceGuard guard(f->device);
  int rc = filter_reserve(f, n);
  if (rc) {return rc;}
  f->n = static_cast<uint32_t>(n);
  if (n) {
    double * h = f->h_stage.as<double>();
    memcpy(h, particles3, n * 3 * sizeof(double));
    memcpy(h + 3 * n, weights, n * sizeof(double));
    NDT2D_CUDA_TRY(cudaMemcpyAsync(f->d_particles[f->cur].p, h, n * 3 * sizeof(double),
      cudaMemcpyHostToDevice, f->stream));
    NDT2D_CUDA_TRY(cudaMemcpyAsync(f->d_weights[f->cur].p, h + 3 * n, n * sizeof(double),
      cudaMemcpyHostToDevice, f->stream));
    NDT2D_CUDA_TRY(cudaStreamSynchronize(f->stream));
    f->ctr.h2d_bytes += n * 4 * sizeof(double);
  }
  return NDT2D_OK;
}

NDT2D_API int ndt2d_filter_size(ndt2d_filter * f, size_t * n)
{
  if (!f || !n) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(f->mu);
  *n = f->n;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_filter_get_particles(ndt2d_filter * f, double * particles3, double * weights)
{
  if (!f) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(f->mu);
  DeviceGuard guard(f->device);
  const size_t n = f->n;
  if (n == 0) {return NDT2D_OK;}
  double * h = f->h_stage.as<double>();
  NDT2D_CUDA_TRY(cudaMemcpyAsync(h, f->d_particles[f->cur].p, n * 3 * sizeof(double),
    cudaMemcpyDeviceToHost, f->stream));
  NDT2D_CUDA_TRY(cudaMemcpyAsync(h + 3 * n, f->d_weights[f->cur].p, n * sizeof(double),
    cudaMemcpyDeviceToHost, f->stream));
  NDT2D_CUDA_TRY(cudaStreamSynchronize(f->stream));
  f->ctr.d2h_bytes += n * 4 * sizeof(double);
  if (particles3) {memcpy(particles3, h, n * 3 * sizeof(double));}
  if (weights) {memcpy(weights, h + 3 * n, n * sizeof(double));}
  return NDT2D_OK;
}

NDT2D_API int ndt2d_filter_init(
  ndt2d_filter * f, double x, double y, double theta, double sigma_x, double sigma_y,
  double sigma_theta, uint64_t seed)
{
  if (!f) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(f->mu);
  DeviceGuard guard(f->device);
  int rc = ndt2d_launch_filter_init(filter_view(f), f->n, x, y, theta, sigma_x, sigma_y,
      sigma_theta, seed, f->stream, &f->ctr);
  if (!rc) {rc = ndt2d_launch_filter_stats(filter_view(f), f->n, f->stream, &f->ctr);}
  if (rc) {return rc;}
  NDT2D_CUDA_TRY(cudaStreamSynchronize(f->stream));
  return NDT2D_OK;
}

NDT2D_API int ndt2d_filter_update(
  ndt2d_filter * f, double dx, double dy, double dth, const double * alphas5, uint64_t seed)
{
  if (!f || !alphas5) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(f->mu);
  DeviceGuard guard(f->device);
  // MotionModel::sample, scalar part (motion_model.cpp:48-66); angle helpers
  // follow the ROS 2 `angles` package.
  auto norm = [](double a) {
      const double r = fmod(a + M_PI, 2.0 * M_PI);
      return (r <= 0.0) ? r + M_PI : r - M_PI;
    };
  auto diff = [&](double from, double to) {return norm(to - from);};
  const double a1 = alphas5[0], a2 = alphas5[1], a3 = alphas5[2], a4 = alphas5[3];
  const double trans = std::hypot(dx, dy);
  const double rot1 = (trans > 0.01) ? atan2(dy, dx) : 0.0;
  const double rot2 = diff(rot1, dth);
  const double rot1_ = std::min(std::fabs(diff(rot1, 0.0)), std::fabs(diff(rot1, M_PI)));
  const double rot2_ = std::min(std::fabs(diff(rot2, 0.0)), std::fabs(diff(rot2, M_PI)));
  const double s_rot1 = std::sqrt(a1 * rot1_ * rot1_ + a2 * trans * trans);
  const double s_trans = std::sqrt(a3 * trans * trans + a4 * rot1_ * rot1_ + a4 * rot2_ * rot2_);
  const double s_rot2 = std::sqrt(a1 * rot2_ * rot2_ + a2 * trans * trans);
  int rc = ndt2d_launch_filter_motion(filter_view(f), f->n, rot1, trans, rot2, s_rot1, s_trans,
      s_rot2, seed, f->stream, &f->ctr);
  if (!rc) {rc = ndt2d_launch_filter_stats(filter_view(f), f->n, f->stream, &f->ctr);}
  if (rc) {return rc;}
  NDT2D_CUDA_TRY(cudaStreamSynchronize(f->stream));
  return NDT2D_OK;
}

NDT2D_API int ndt2d_filter_measure(
  ndt2d_filter * f, ndt2d_matcher * m, const double * pts_xy, size_t npts)
{
  if (!f || !m || (npts && !pts_xy)) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(f->mu);
  std::lock_guard<std::mutex> lock_m(m->mu);
  if (f->device != m->device) {return NDT2D_ERR_INVALID;}
  DeviceGuard guard(f->device);
  const size_t n = f->n;
  cudaStream_t st = f->stream;
  if (!m->has_model) {
    // scorePoints returns 0.0 for every particle (scan_matcher_ndt.cpp:159);
    // updateStatistics then divides by a zero sum, as the reference would.
    if (n) {
      NDT2D_CUDA_TRY(cudaMemsetAsync(f->d_weights[f->cur].p, 0, n * sizeof(double), st));
    }
  } else if (n) {
    // make sure the model build (matcher stream) is complete
    NDT2D_CUDA_TRY(cudaStreamSynchronize(m->stream));
    const size_t n_use = subsample_count(m, npts);
    int rc = f->d_pts.ensure(n_use ? n_use * sizeof(double2) : 16);
    if (!rc) {rc = f->h_stage.ensure((n * 4 + n_use * 2) * sizeof(double) + 256);}
    if (rc) {return rc;}
    // particle poses -> host, cos/sin by the host libm (toEigen: conversions.hpp:64-68)
    double * h_part = f->h_stage.as<double>();
    double4 * h_tf = reinterpret_cast<double4 *>(h_part);  // written after the read below
    std::vector<double> part(n * 3);
    NDT2D_CUDA_TRY(cudaMemcpyAsync(h_part, f->d_particles[f->cur].p, n * 3 * sizeof(double),
      cudaMemcpyDeviceToHost, st));
    NDT2D_CUDA_TRY(cudaStreamSynchronize(st));
    memcpy(part.data(), h_part, n * 3 * sizeof(double));
    f->ctr.d2h_bytes += n * 3 * sizeof(double);
    for (size_t i = 0; i < n; ++i) {
      h_tf[i] = make_double4(part[3 * i], part[3 * i + 1], cos(part[3 * i + 2]),
          sin(part[3 * i + 2]));
    }
    double * h_pts = reinterpret_cast<double *>(h_tf + n);
    if (n_use) {subsample_points(pts_xy, npts, n_use, h_pts);}
    NDT2D_CUDA_TRY(cudaMemcpyAsync(f->d_pose_tf.p, h_tf, n * sizeof(double4),
      cudaMemcpyHostToDevice, st));
    if (n_use) {
      NDT2D_CUDA_TRY(cudaMemcpyAsync(f->d_pts.p, h_pts, n_use * sizeof(double2),
        cudaMemcpyHostToDevice, st));
    }
    f->ctr.h2d_bytes += n * sizeof(double4) + n_use * sizeof(double2);
    rc = ndt2d_launch_score_poses(model_view(m), f->d_pts.as<double2>(),
        static_cast<uint32_t>(n_use), f->d_pose_tf.as<double4>(), static_cast<uint32_t>(n), -1.0,
        1, f->d_weights[f->cur].as<double>(), st, &f->ctr);
    if (rc) {return rc;}
  }
  int rc = ndt2d_launch_filter_stats(filter_view(f), f->n, st, &f->ctr);
  if (rc) {return rc;}
  NDT2D_CUDA_TRY(cudaStreamSynchronize(st));
  return NDT2D_OK;
}

NDT2D_API int ndt2d_filter_resample(
  ndt2d_filter * f, double kld_err, double kld_z, const double * uniforms, size_t n_uniforms,
  uint64_t seed)
{
  if (!f) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(f->mu);
  if (uniforms && n_uniforms < f->max_particles) {return NDT2D_ERR_INVALID;}
  DeviceGuard guard(f->device);
  cudaStream_t st = f->stream;
  const double * d_u = nullptr;
  if (uniforms) {
    int rc = f->h_stage.ensure(f->max_particles * sizeof(double));
    if (rc) {return rc;}
    memcpy(f->h_stage.p, uniforms, f->max_particles * sizeof(double));
    NDT2D_CUDA_TRY(cudaMemcpyAsync(f->d_uniforms.p, f->h_stage.p,
      f->max_particles * sizeof(double), cudaMemcpyHostToDevice, st));
    f->ctr.h2d_bytes += f->max_particles * sizeof(double);
    d_u = f->d_uniforms.as<double>();
  }
  const int nxt = f->cur ^ 1;
  int rc = ndt2d_launch_filter_resample(filter_view(f), f->n,
      static_cast<uint32_t>(f->min_particles), static_cast<uint32_t>(f->max_particles), kld_err,
      kld_z, d_u, seed, f->d_cdf.as<double>(), f->d_particles[nxt].as<double>(),
      f->d_weights[nxt].as<double>(), f->d_draws.as<uint32_t>(), f->d_first.as<uint32_t>(),
      f->d_canon.as<uint32_t>(), f->d_table.as<uint32_t>(), f->table_size,
      f->d_new_n.as<uint32_t>(), st, &f->ctr);
  if (rc) {return rc;}
  uint32_t new_n = 0;
  NDT2D_CUDA_TRY(cudaMemcpyAsync(&new_n, f->d_new_n.p, sizeof(new_n), cudaMemcpyDeviceToHost, st));
  NDT2D_CUDA_TRY(cudaStreamSynchronize(st));
  f->ctr.d2h_bytes += sizeof(new_n);
  f->cur = nxt;
  f->n = new_n;
  rc = ndt2d_launch_filter_stats(filter_view(f), f->n, st, &f->ctr);  // :136
  if (rc) {return rc;}
  NDT2D_CUDA_TRY(cudaStreamSynchronize(st));
  return NDT2D_OK;
}

NDT2D_API int ndt2d_filter_stats(ndt2d_filter * f, double * mean3, double * cov9)
{
  if (!f) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(f->mu);
  DeviceGuard guard(f->device);
  double h[12];
  NDT2D_CUDA_TRY(cudaMemcpyAsync(h, f->d_stats.p, sizeof(h), cudaMemcpyDeviceToHost, f->stream));
  NDT2D_CUDA_TRY(cudaStreamSynchronize(f->stream));
  f->ctr.d2h_bytes += sizeof(h);
  if (mean3) {memcpy(mean3, h, 3 * sizeof(double));}
  if (cov9) {memcpy(cov9, h + 3, 9 * sizeof(double));}
  return NDT2D_OK;
}

NDT2D_API int ndt2d_filter_set_cov(ndt2d_filter * f, const double * cov9)
{
  if (!f || !cov9) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(f->mu);
  DeviceGuard guard(f->device);
  NDT2D_CUDA_TRY(cudaMemcpyAsync(f->d_stats.as<double>() + 3, cov9, 9 * sizeof(double),
    cudaMemcpyHostToDevice, f->stream));
  NDT2D_CUDA_TRY(cudaStreamSynchronize(f->stream));
  return NDT2D_OK;
}

NDT2D_API int ndt2d_filter_last_draws(ndt2d_filter * f, uint64_t * out)
{
  if (!f || !out) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(f->mu);
  DeviceGuard guard(f->device);
  const size_t n = f->n;
  if (n == 0) {return NDT2D_OK;}
  std::vector<uint32_t> h(n);
  NDT2D_CUDA_TRY(cudaMemcpyAsync(h.data(), f->d_draws.p, n * sizeof(uint32_t),
    cudaMemcpyDeviceToHost, f->stream));
  NDT2D_CUDA_TRY(cudaStreamSynchronize(f->stream));
  for (size_t i = 0; i < n; ++i) {out[i] = h[i];}
  return NDT2D_OK;
}

}  // extern "C"


// ---------------------------------------------------------------- one handle, several GPUs
// (ndt2d_params.n_devices > 1).  The C++ plugin is loaded into ONE process (the node creates its
// matchers at ndt_mapper.cpp:54, 299-312 and calls matchScan at :552-553, 638-643), so the
// N-GPU search has to be reachable from one host thread: rank r = devices[r] has its own
// sub-handle (stream, buffers, replicated model and scan), the host enqueues the strided
// searches on all devices back to back and fetches the combined record from rank 0.  The
// exchange is the same mailbox protocol as the one-process-per-GPU path
// (ndt2d_matcher_search_exchange), the peers' mailboxes being plain device pointers here:
// peer access is enabled between every pair of devices.

static int group_create(ndt2d_matcher * m)
{
  const int n = m->prm.n_devices;
  m->group.assign(1, m);
  for (int r = 1; r < n; ++r) {
    ndt2d_params p = m->prm;
    p.n_devices = 0;
    p.device = m->prm.devices[r];
    p.stream = nullptr;
    ndt2d_matcher * sub = nullptr;
    const int rc = ndt2d_matcher_create(&p, &sub);
    if (rc) {return rc;}
    m->group.push_back(sub);
  }
  // peer access, every ordered pair
  bool p2p = true;
  for (int a = 0; a < n && p2p; ++a) {
    for (int b = 0; b < n && p2p; ++b) {
      if (a == b) {continue;}
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, m->prm.devices[a], m->prm.devices[b]) != cudaSuccess || !can) {
        cudaGetLastError();
        p2p = false;
        break;
      }
      DeviceGuard guard(m->prm.devices[a]);
      const cudaError_t e = cudaDeviceEnablePeerAccess(m->prm.devices[b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {p2p = false;}
      cudaGetLastError();
    }
  }
  if (p2p) {
    std::vector<void *> boxes(n, nullptr);
    for (int r = 0; r < n; ++r) {
      ndt2d_matcher * s = m->group[r];
      DeviceGuard guard(s->device);
      exchange_close(s);
      int rc = s->d_mailbox.ensure(kExchangeMailboxBytes);
      if (!rc) {rc = s->d_peer_table.ensure(kExchangeMaxRanks * sizeof(void *));}
      if (rc) {return rc;}
      NDT2D_CUDA_TRY(cudaMemset(s->d_mailbox.p, 0, kExchangeMailboxBytes));
      boxes[r] = s->d_mailbox.p;
    }
    for (int r = 0; r < n; ++r) {
      ndt2d_matcher * s = m->group[r];
      DeviceGuard guard(s->device);
      s->peer_ptrs = boxes;
      s->x_world = static_cast<uint32_t>(n);
      s->x_rank = static_cast<uint32_t>(r);
      s->x_local = true;
      NDT2D_CUDA_TRY(cudaMemcpy(s->d_peer_table.p, boxes.data(), n * sizeof(void *),
        cudaMemcpyHostToDevice));
      s->x_connected = true;
    }
  }
  m->group_p2p = p2p;
  for (int r = 1; r < n; ++r) {
    ndt2d_matcher * s = m->group[r];
    GroupWorker * w = new (std::nothrow) GroupWorker();
    if (!w) {return NDT2D_ERR_INVALID;}
    s->worker = w;
    const int dev = s->device;
    w->th = std::thread([w, dev]() {
          cudaSetDevice(dev);
          std::unique_lock<std::mutex> lk(w->mu);
          bool hot = false;
          for (;;) {
            if (hot && !w->has_task && !w->quit) {
              // just finished a task: poll for the next one for 200 us before sleeping (waking a
              // sleeping thread costs tens of microseconds, which every device of a search pays)
              lk.unlock();
              const auto t0 = std::chrono::steady_clock::now();
              while (!w->pending.load(std::memory_order_acquire) &&
                std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(200))
              {
                cpu_relax();
              }
              lk.lock();
            }
            w->cv.wait(lk, [w]() {return w->has_task || w->quit;});
            if (w->quit) {return;}
            std::function<int()> task = std::move(w->task);
            w->has_task = false;
            w->pending.store(false, std::memory_order_relaxed);
            lk.unlock();
            g_last_error[0] = 0;
            const int rc = task();
            lk.lock();
            w->rc = rc;
            snprintf(w->err, sizeof(w->err), "%s", g_last_error);
            w->done = true;
            w->finished.store(true, std::memory_order_release);
            w->cv.notify_all();
            hot = true;
          }
        });
  }
  return NDT2D_OK;
}

// fn(rank, handle) on every device of the group at once: ranks >= 1 on their worker threads
// (already bound to their device), rank 0 on the calling thread.  Returns the first failure.
static int group_run(ndt2d_matcher * m, const std::function<int(size_t, ndt2d_matcher *)> & fn)
{
  const size_t world = m->group.size();
  for (size_t r = 1; r < world; ++r) {
    ndt2d_matcher * s = m->group[r];
    GroupWorker * w = s->worker;
    {
      std::lock_guard<std::mutex> wl(w->mu);
      w->task = [&fn, r, s]() {return fn(r, s);};
      w->has_task = true;
      w->done = false;
      w->finished.store(false, std::memory_order_relaxed);
      w->pending.store(true, std::memory_order_release);
    }
    w->cv.notify_all();
  }
  int rc = NDT2D_OK;
  {
    DeviceGuard guard(m->device);
    rc = fn(0, m);
  }
  for (size_t r = 1; r < world; ++r) {
    GroupWorker * w = m->group[r]->worker;
    {
      // the workers only enqueue (tens of microseconds): poll before falling back to the cv
      const auto t0 = std::chrono::steady_clock::now();
      while (!w->finished.load(std::memory_order_acquire) &&
        std::chrono::steady_clock::now() - t0 < std::chrono::milliseconds(2))
      {
        cpu_relax();
      }
    }
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [w]() {return w->done;});
    if (w->rc && !rc) {
      rc = w->rc;
      snprintf(g_last_error, sizeof(g_last_error), "%s", w->err);
    }
  }
  return rc;
}

// A search is spread over the devices when every device gets enough slices to amortise the
// replicated staging (a local match of 20,000 candidates is latency-bound on one GPU already).
static bool group_worthwhile(const ndt2d_matcher * m, size_t npts)
{
  if (m->group.size() < 2) {return false;}
  const double n_use = static_cast<double>(subsample_count(m, npts));
  const double pairs = static_cast<double>(m->dth.size()) * m->dlin.size() * m->dlin.size() * n_use;
  // (the default, 1e10 pairs, is ~0.4 ms of one B200)
  return pairs >= m->group_min_pairs && m->dth.size() >= 4 * m->group.size();
}

static int match_scan_group(
  ndt2d_matcher * m, const double * pose3, const double * pts_xy, size_t npts,
  double * out_delta3, int * delta_written, double * out_cov9, double * out_score)
{
  const size_t world = m->group.size();
  const size_t n_ang = m->dth.size();
  for (size_t r = 0; r < world; ++r) {
    if (!m->group[r]->has_model) {return NDT2D_ERR_STATE;}
  }
  const unsigned long long seq = ++m->group_seq;
  const bool p2p = m->group_p2p;
  const int variant = m->prm.kernel_variant;
  // every device at once (one host thread each): stage the scan, the libm (cos, sin) of the
  // device's own theta slices, the strided search + fused exchange.  Nobody waits for a stream
  // here: with the fused exchange rank 0's finish kernel holds the combined record only after
  // every device has published its own, and it stores that record into rank 0's host mailbox.
  HostMailbox hm{nullptr, nullptr, 0ull};
  int rc = group_run(m, [&](size_t r, ndt2d_matcher * s) {
        int rc1 = stage_scan_locked(s, pose3, pts_xy, npts, true);
        if (!rc1) {rc1 = ensure_trig_locked(s, r, n_ang, world);}
        if (rc1) {return rc1;}
        if (r == 0 && p2p) {hm = mailbox_arm(s);}
        SearchView sv = search_view(s);
        sv.theta_stride = static_cast<uint32_t>(world);
        ExchangeView xv;
        xv.peers = s->d_peer_table.as<void *>();
        xv.world = static_cast<uint32_t>(world);
        xv.rank = static_cast<uint32_t>(r);
        xv.seq = seq;
        xv.timeout_ns = 10ull * 1000ull * 1000ull * 1000ull;
        rc1 = ndt2d_launch_search(model_view(s), sv, static_cast<uint32_t>(r),
            static_cast<uint32_t>(n_ang), variant, s->d_blockpart.as<double>(),
            s->d_partial.as<double>(), nullptr, s->d_counter.as<uint32_t>(), s->stream, &s->ctr,
            s->ev_begin, s->ev_end, p2p ? &xv : nullptr, (r == 0 && p2p) ? &hm : nullptr);
        s->ev_valid = rc1 == NDT2D_OK;
        return rc1;
      });
  if (rc) {
    for (size_t r = 0; r < world; ++r) {cudaStreamSynchronize(m->group[r]->stream);}
    return rc;
  }
  m->group_searches += 1;
  if (m->group_p2p) {
    // every rank holds the combined record; rank 0's is the answer
    double r32[32];
    {
      DeviceGuard guard(m->device);
      if ((rc = mailbox_wait(m, hm, false, r32))) {return rc;}
    }
    for (size_t r = 1; r < world; ++r) {
      ndt2d_matcher * s = m->group[r];
      m->ctr.launches += s->ctr.launches;
      m->ctr.h2d_bytes += s->ctr.h2d_bytes;
      s->ctr = Counters{0, 0, 0};
    }
    if (r32[31] != 0.0) {
      snprintf(g_last_error, sizeof(g_last_error),
        "fused exchange timed out: not every device published its partial record");
      return NDT2D_ERR_STATE;
    }
    unpack_result(r32, out_delta3, delta_written, out_cov9, out_score);
    return NDT2D_OK;
  }
  // no peer access between the devices: fetch the N partial records and fold them on the host
  std::vector<double> parts(world * NDT2D_PARTIAL_DOUBLES);
  for (size_t r = 0; r < world; ++r) {
    ndt2d_matcher * s = m->group[r];
    DeviceGuard guard(s->device);
    double r32[32];
    if ((rc = fetch_result_locked(s, r32))) {return rc;}
    memcpy(parts.data() + r * NDT2D_PARTIAL_DOUBLES, r32, NDT2D_PARTIAL_DOUBLES * sizeof(double));
    if (r) {
      m->ctr.launches += s->ctr.launches;
      m->ctr.h2d_bytes += s->ctr.h2d_bytes;
      m->ctr.d2h_bytes += s->ctr.d2h_bytes;
      s->ctr = Counters{0, 0, 0};
    }
  }
  if (delta_written) {*delta_written = 0;}
  return ndt2d_combine_partials_host(m->dth.data(), m->dth.size(), m->dlin.data(), m->dlin.size(),
           parts.data(), world, out_delta3, delta_written, out_cov9, out_score);
}

extern "C" {

/* info4 = devices of the handle, whether the fused peer-to-peer exchange is in use, matchScans
 * that ran on all devices so far, sequence number of the last exchange */
NDT2D_API int ndt2d_matcher_group_info(ndt2d_matcher * m, uint64_t * info4)
{
  if (!m || !info4) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  info4[0] = m->group.empty() ? 1 : m->group.size();
  info4[1] = m->group_p2p ? 1 : 0;
  info4[2] = m->group_searches;
  info4[3] = m->group_seq;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_set_group_threshold(ndt2d_matcher * m, double min_pairs)
{
  if (!m || !(min_pairs >= 0.0)) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  m->group_min_pairs = min_pairs;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_set_tallies(ndt2d_matcher * m, int on)
{
  if (!m) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  m->tally = on != 0;
  for (size_t r = 1; r < m->group.size(); ++r) {m->group[r]->tally = m->tally;}
  for (ndt2d_matcher * s : m->lanes) {s->tally = m->tally;}
  return NDT2D_OK;
}

NDT2D_API int ndt2d_matcher_set_timing(ndt2d_matcher * m, int on)
{
  if (!m) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  m->time_small = on != 0;
  return NDT2D_OK;
}

/* Multi-device handle: duration (ms, CUDA events on each device's stream) of the search kernels of
 * the last matchScan on every device and the tallies of that search summed over the devices:
 * out_ms[n_devices]; totals3 = useful evaluations, (point, region) items, 0. */
NDT2D_API int ndt2d_matcher_group_search_stats(
  ndt2d_matcher * m, double * out_ms, size_t cap, uint64_t * totals3)
{
  if (!m || !out_ms) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(m->mu);
  const size_t n = m->group.empty() ? 1 : m->group.size();
  if (cap < n) {return NDT2D_ERR_SIZE;}
  if (totals3) {totals3[0] = totals3[1] = totals3[2] = 0;}
  for (size_t r = 0; r < n; ++r) {
    ndt2d_matcher * s = m->group.empty() ? m : m->group[r];
    out_ms[r] = 0.0;
    if (!s->d_counter.p) {continue;}
    DeviceGuard guard(s->device);
    NDT2D_CUDA_TRY(cudaStreamSynchronize(s->stream));
    uint64_t h[6] = {0, 0, 0, 0, 0, 0};
    NDT2D_CUDA_TRY(cudaMemcpy(h, s->d_counter.p, sizeof(h), cudaMemcpyDeviceToHost));
    if (totals3) {
      totals3[0] += h[3];
      totals3[1] += h[4];
    }
    if (s->ev_valid && s->ev_begin && s->ev_end) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, s->ev_begin, s->ev_end) == cudaSuccess) {
        out_ms[r] = ms;
      } else {
        cudaGetLastError();
      }
    }
  }
  return NDT2D_OK;
}

NDT2D_API int ndt2d_probe_call_latency(
  ndt2d_matcher * m, int what, size_t n_scans, const double * map_poses,
  const uint64_t * map_pt_offsets, const double * map_pts_xy, const double * pose3,
  const double * pts_xy, size_t npts, size_t calls, double * out_us)
{
  if (!m || !pose3 || !out_us || (npts && !pts_xy) || (what != 0 && what != 1)) {
    return NDT2D_ERR_INVALID;
  }
  for (size_t k = 0; k < calls; ++k) {
    double delta[3], cov[9], score = 0.0, s0 = 0.0;
    int written = 0, rc = NDT2D_OK;
    const auto t0 = std::chrono::steady_clock::now();
    if (what == 1) {
      rc = ndt2d_matcher_reset(m);
      if (!rc) {rc = ndt2d_matcher_add_scans(m, n_scans, map_poses, map_pt_offsets, map_pts_xy);}
      if (!rc) {rc = ndt2d_matcher_score_points(m, pts_xy, npts, pose3, &s0);}
    }
    if (!rc) {rc = ndt2d_matcher_match_scan(m, pose3, pts_xy, npts, delta, &written, cov, &score);}
    const auto t1 = std::chrono::steady_clock::now();
    if (rc) {return rc;}
    out_us[k] = std::chrono::duration<double, std::micro>(t1 - t0).count();
  }
  return NDT2D_OK;
}

}  // extern "C"
