// occupancy.cu -- occupancy-grid export: replaces
// ndt_2d::OccupancyGrid::getMsg / updateBounds (src/occupancy_grid.cpp:47-185).
//
//   bounds     min / max of the world coordinates of the NEW scans' points (the reference
//              keeps its bounds across calls and only looks at scans it has not seen,
//              :155-177), then floor / ceil to the resolution on the host (:180-183)
//   raytrace   one thread per scan point: "simplified Bresenham" from the scan pose's cell
//              to the point's cell (:96-133), integer atomicAdd on the empty / hit counters
//   finalize   hit / (hit + empty) > occ_thresh -> 100, else 0, untouched -> -1 (:138-151)
//
// Everything that decides a cell is the reference's own double arithmetic (no FMA, host
// libm cos / sin per scan), the counters are integers: the grid is bit-exact.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "ndt2d_internal.h"

namespace
{

// order-preserving map double -> u64 (for atomicMin / atomicMax on doubles)
__device__ __forceinline__ unsigned long long ordered_key(double v)
{
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__host__ double ordered_value(unsigned long long k)
{
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  double v;
  memcpy(&v, &b, sizeof(v));
  return v;
}

// updateBounds (:155-177): p = pose + R(theta) * point for the points of scans >= first_scan.
// bounds[0..3] = ordered keys of min_x, max_x, min_y, max_y (seeded with the previous bounds).
__global__ void __launch_bounds__(256) occupancy_bounds_kernel(
  const double4 * __restrict__ scan_tf, const uint64_t * __restrict__ offsets, uint32_t first_scan,
  uint32_t n_scans, const double2 * __restrict__ pts, unsigned long long * __restrict__ bounds)
{
  for (uint32_t s = first_scan + blockIdx.x; s < n_scans; s += gridDim.x) {
    const double4 tf = scan_tf[s];
    double mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
    for (uint64_t p = offsets[s] + threadIdx.x; p < offsets[s + 1]; p += blockDim.x) {
      const double2 pt = pts[p];
      const double X = __dadd_rn(tf.x, __dsub_rn(__dmul_rn(pt.x, tf.z), __dmul_rn(pt.y, tf.w)));
      const double Y = __dadd_rn(tf.y, __dadd_rn(__dmul_rn(pt.x, tf.w), __dmul_rn(pt.y, tf.z)));
      mnx = fmin(mnx, X);
      mxx = fmax(mxx, X);
      mny = fmin(mny, Y);
      mxy = fmax(mxy, Y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
      mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
      mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o));
      mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((threadIdx.x & 31u) == 0u) {
      if (mnx != INFINITY) {
        atomicMin(bounds + 0, ordered_key(mnx));
        atomicMax(bounds + 1, ordered_key(mxx));
        atomicMin(bounds + 2, ordered_key(mny));
        atomicMax(bounds + 3, ordered_key(mxy));
      }
    }
  }
}

struct GridInfo
{
  double origin_x, origin_y, resolution;
  uint32_t width, height;
};

__device__ __forceinline__ void touch(int * __restrict__ counter, const GridInfo & gi, int x, int y)
{
  // the reference indexes without a check (:112-113); cells outside the grid can only be
  // reached if poses moved after the bounds were taken -- those updates are dropped here
  if (x >= 0 && y >= 0 && static_cast<uint32_t>(x) < gi.width && static_cast<uint32_t>(y) < gi.height) {
    atomicAdd(counter + (static_cast<size_t>(y) * gi.width + x), 1);
  }
}

__global__ void __launch_bounds__(128) occupancy_raytrace_kernel(
  GridInfo gi, const double4 * __restrict__ scan_tf, const uint64_t * __restrict__ offsets,
  uint32_t n_scans, const double2 * __restrict__ pts, int * __restrict__ hit,
  int * __restrict__ empty)
{
  for (uint32_t s = blockIdx.x; s < n_scans; s += gridDim.x) {
    const double4 tf = scan_tf[s];   // pose x, y, cos, sin
    // :82-83   int((pose - origin) / resolution), truncation toward zero
    const int start_x = static_cast<int>(__ddiv_rn(__dsub_rn(tf.x, gi.origin_x), gi.resolution));
    const int start_y = static_cast<int>(__ddiv_rn(__dsub_rn(tf.y, gi.origin_y), gi.resolution));
    for (uint64_t p = offsets[s] + threadIdx.x; p < offsets[s + 1]; p += blockDim.x) {
      const double2 pt = pts[p];
      // :87-88   point.x * cos - point.y * sin + pose_x
      const double px = __dadd_rn(__dsub_rn(__dmul_rn(pt.x, tf.z), __dmul_rn(pt.y, tf.w)), tf.x);
      const double py = __dadd_rn(__dadd_rn(__dmul_rn(pt.x, tf.w), __dmul_rn(pt.y, tf.z)), tf.y);
      const int end_x = static_cast<int>(__ddiv_rn(__dsub_rn(px, gi.origin_x), gi.resolution));
      const int end_y = static_cast<int>(__ddiv_rn(__dsub_rn(py, gi.origin_y), gi.resolution));
      // Simplified Bresenham (:93-133)
      const int dx = abs(end_x - start_x);
      const int sx = (start_x < end_x) ? 1 : -1;
      const int dy = -abs(end_y - start_y);
      const int sy = (start_y < end_y) ? 1 : -1;
      int error = dx + dy;
      int x = start_x, y = start_y;
      while (true) {
        if (x == end_x && y == end_y) {
          touch(hit, gi, x, y);
          break;
        }
        touch(empty, gi, x, y);
        const int cx = x, cy = y;   // `index` of this iteration: both hits below use it
        if (2 * error >= dy) {
          if (x == end_x) {
            touch(hit, gi, cx, cy);
            break;
          }
          error = error + dy;
          x += sx;
        }
        if (2 * error <= dx) {
          if (y == end_y) {
            touch(hit, gi, cx, cy);
            break;
          }
          error = error + dx;
          y += sy;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) occupancy_finalize_kernel(
  const int * __restrict__ hit, const int * __restrict__ empty, size_t n, double occ_thresh,
  signed char * __restrict__ data)
{
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) {return;}
  const int h = hit[i];
  const double touches = static_cast<double>(h + empty[i]);   // int sum, then double (:140)
  signed char v = -1;
  if (touches > 0.5) {
    v = (__ddiv_rn(static_cast<double>(h), touches) > occ_thresh) ? 100 : 0;
  }
  data[i] = v;
}

struct Buf
{
  void * p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes)
  {
    if (bytes <= cap) {return NDT2D_OK;}
    if (p) {cudaFree(p);}
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
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
};

}  // namespace

struct ndt2d_occupancy
{
  std::mutex mu;
  int device = 0;
  cudaStream_t stream = nullptr;
  double resolution = 0.05, occ_thresh = 0.25;
  // OccupancyGrid's persistent state (occupancy_grid.cpp:35-44)
  double min_x = 0, max_x = 0, min_y = 0, max_y = 0;
  size_t num_scans = 0;
  GridInfo gi{};
  bool rendered = false;
  Buf d_tf, d_off, d_pts, d_hit, d_empty, d_data, d_bounds;
  uint64_t launches = 0;
};

extern "C" {

NDT2D_API int ndt2d_occupancy_create(
  double resolution, double occ_thresh, int device, ndt2d_occupancy ** out)
{
  if (!out || !(resolution > 0.0) || !std::isfinite(resolution)) {return NDT2D_ERR_INVALID;}
  *out = nullptr;
  if (ndt2d_device_count() <= 0) {return NDT2D_ERR_NO_DEVICE;}
  int dev = device;
  if (dev < 0) {NDT2D_CUDA_TRY(cudaGetDevice(&dev));}
  ndt2d_occupancy * g = new (std::nothrow) ndt2d_occupancy();
  if (!g) {return NDT2D_ERR_INVALID;}
  g->device = dev;
  g->resolution = resolution;
  g->occ_thresh = occ_thresh;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(dev);
  const cudaError_t e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
  if (prev >= 0) {cudaSetDevice(prev);}
  if (e != cudaSuccess) {
    ndt2d_set_error("cudaStreamCreate", e, __FILE__, __LINE__);
    delete g;
    return NDT2D_ERR_CUDA;
  }
  *out = g;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_occupancy_destroy(ndt2d_occupancy * g)
{
  if (!g) {return NDT2D_OK;}
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(g->device);
  if (g->stream) {
    cudaStreamSynchronize(g->stream);
    cudaStreamDestroy(g->stream);
  }
  for (Buf * b : {&g->d_tf, &g->d_off, &g->d_pts, &g->d_hit, &g->d_empty, &g->d_data, &g->d_bounds}) {
    b->release();
  }
  if (prev >= 0) {cudaSetDevice(prev);}
  delete g;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_occupancy_render(
  ndt2d_occupancy * g, size_t n_scans, const double * poses, const uint64_t * pt_offsets,
  const double * pts_xy, double * info5)
{
  if (!g || !info5 || (n_scans && (!poses || !pt_offsets))) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(g->mu);
  int prev = -1;
  cudaGetDevice(&prev);
  NDT2D_CUDA_TRY(cudaSetDevice(g->device));
  struct Restore {int d; ~Restore() {if (d >= 0) {cudaSetDevice(d);}}} restore{prev};
  cudaStream_t st = g->stream;
  const size_t n_points = n_scans ? static_cast<size_t>(pt_offsets[n_scans] - pt_offsets[0]) : 0;
  if (n_points && !pts_xy) {return NDT2D_ERR_INVALID;}
  const uint64_t off0 = n_scans ? pt_offsets[0] : 0;
  // ---- uploads: per-scan (x, y, cos, sin) with host libm (:77-80), rebased offsets, points
  std::vector<double4> tf(n_scans ? n_scans : 1);
  std::vector<uint64_t> off(n_scans + 1);
  for (size_t k = 0; k < n_scans; ++k) {
    tf[k] = make_double4(poses[3 * k], poses[3 * k + 1], cos(poses[3 * k + 2]), sin(poses[3 * k + 2]));
    off[k] = pt_offsets[k] - off0;
  }
  off[n_scans] = n_points;
  int rc = g->d_tf.ensure(tf.size() * sizeof(double4));
  if (!rc) {rc = g->d_off.ensure(off.size() * sizeof(uint64_t));}
  if (!rc) {rc = g->d_pts.ensure((n_points ? n_points : 1) * sizeof(double2));}
  if (!rc) {rc = g->d_bounds.ensure(4 * sizeof(unsigned long long));}
  if (rc) {return rc;}
  NDT2D_CUDA_TRY(cudaMemcpyAsync(g->d_tf.p, tf.data(), tf.size() * sizeof(double4),
    cudaMemcpyHostToDevice, st));
  NDT2D_CUDA_TRY(cudaMemcpyAsync(g->d_off.p, off.data(), off.size() * sizeof(uint64_t),
    cudaMemcpyHostToDevice, st));
  if (n_points) {
    NDT2D_CUDA_TRY(cudaMemcpyAsync(g->d_pts.p, pts_xy + 2 * off0, n_points * sizeof(double2),
      cudaMemcpyHostToDevice, st));
  }
  // ---- updateBounds (:47-52, :155-184): only when the number of scans changed, only new scans
  if (n_scans != g->num_scans) {
    const size_t start_idx = g->num_scans;
    g->num_scans = n_scans;
    if (start_idx < n_scans) {
      // seed the reduction with "nothing seen"; the previous bounds are folded in on the host
      const double seed[4] = {INFINITY, -INFINITY, INFINITY, -INFINITY};
      unsigned long long keys[4];
      for (int k = 0; k < 4; ++k) {
        unsigned long long b;
        memcpy(&b, &seed[k], 8);
        keys[k] = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
      }
      NDT2D_CUDA_TRY(cudaMemcpyAsync(g->d_bounds.p, keys, sizeof(keys), cudaMemcpyHostToDevice, st));
      const uint32_t nb = static_cast<uint32_t>(std::min<size_t>(n_scans - start_idx, 4096));
      occupancy_bounds_kernel<<<nb, 256, 0, st>>>(
        static_cast<const double4 *>(g->d_tf.p), static_cast<const uint64_t *>(g->d_off.p),
        static_cast<uint32_t>(start_idx), static_cast<uint32_t>(n_scans),
        static_cast<const double2 *>(g->d_pts.p), static_cast<unsigned long long *>(g->d_bounds.p));
      NDT2D_CUDA_TRY(cudaGetLastError());
      g->launches += 1;
      NDT2D_CUDA_TRY(cudaMemcpyAsync(keys, g->d_bounds.p, sizeof(keys), cudaMemcpyDeviceToHost, st));
      NDT2D_CUDA_TRY(cudaStreamSynchronize(st));
      const double nmin_x = ordered_value(keys[0]), nmax_x = ordered_value(keys[1]);
      const double nmin_y = ordered_value(keys[2]), nmax_y = ordered_value(keys[3]);
      if (nmin_x != INFINITY) {
        g->min_x = std::min(nmin_x, g->min_x);
        g->max_x = std::max(nmax_x, g->max_x);
        g->min_y = std::min(nmin_y, g->min_y);
        g->max_y = std::max(nmax_y, g->max_y);
      }
    }
    // Make min/max values a multiple of resolution (:180-183)
    g->min_x = std::floor(g->min_x / g->resolution) * g->resolution;
    g->max_x = std::ceil(g->max_x / g->resolution) * g->resolution;
    g->min_y = std::floor(g->min_y / g->resolution) * g->resolution;
    g->max_y = std::ceil(g->max_y / g->resolution) * g->resolution;
  }
  // ---- meta data (:54-64)
  const double pad = 5 * g->resolution;
  GridInfo gi;
  gi.resolution = g->resolution;
  gi.width = static_cast<uint32_t>((g->max_x - g->min_x + 2 * pad) / g->resolution);
  gi.height = static_cast<uint32_t>((g->max_y - g->min_y + 2 * pad) / g->resolution);
  gi.origin_x = g->min_x - pad;
  gi.origin_y = g->min_y - pad;
  const size_t n_cells = static_cast<size_t>(gi.width) * gi.height;
  if (n_cells >= (size_t(1) << 31)) {return NDT2D_ERR_SIZE;}
  rc = g->d_hit.ensure((n_cells ? n_cells : 1) * sizeof(int));
  if (!rc) {rc = g->d_empty.ensure((n_cells ? n_cells : 1) * sizeof(int));}
  if (!rc) {rc = g->d_data.ensure(n_cells ? n_cells : 1);}
  if (rc) {return rc;}
  if (n_cells) {
    NDT2D_CUDA_TRY(cudaMemsetAsync(g->d_hit.p, 0, n_cells * sizeof(int), st));
    NDT2D_CUDA_TRY(cudaMemsetAsync(g->d_empty.p, 0, n_cells * sizeof(int), st));
    if (n_scans) {
      const uint32_t nb = static_cast<uint32_t>(std::min<size_t>(n_scans, 65535u * 4u));
      occupancy_raytrace_kernel<<<nb, 128, 0, st>>>(
        gi, static_cast<const double4 *>(g->d_tf.p), static_cast<const uint64_t *>(g->d_off.p),
        static_cast<uint32_t>(n_scans), static_cast<const double2 *>(g->d_pts.p),
        static_cast<int *>(g->d_hit.p), static_cast<int *>(g->d_empty.p));
      NDT2D_CUDA_TRY(cudaGetLastError());
      g->launches += 1;
    }
    occupancy_finalize_kernel<<<static_cast<uint32_t>((n_cells + 255) / 256), 256, 0, st>>>(
      static_cast<const int *>(g->d_hit.p), static_cast<const int *>(g->d_empty.p), n_cells,
      g->occ_thresh, static_cast<signed char *>(g->d_data.p));
    NDT2D_CUDA_TRY(cudaGetLastError());
    g->launches += 1;
  }
  NDT2D_CUDA_TRY(cudaStreamSynchronize(st));
  g->gi = gi;
  g->rendered = true;
  info5[0] = gi.width;
  info5[1] = gi.height;
  info5[2] = gi.origin_x;
  info5[3] = gi.origin_y;
  info5[4] = gi.resolution;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_occupancy_fetch(ndt2d_occupancy * g, int8_t * data, size_t capacity)
{
  if (!g || (!data && capacity)) {return NDT2D_ERR_INVALID;}
  std::lock_guard<std::mutex> lock(g->mu);
  if (!g->rendered) {return NDT2D_ERR_STATE;}
  const size_t n_cells = static_cast<size_t>(g->gi.width) * g->gi.height;
  if (capacity < n_cells) {return NDT2D_ERR_SIZE;}
  if (n_cells == 0) {return NDT2D_OK;}
  int prev = -1;
  cudaGetDevice(&prev);
  NDT2D_CUDA_TRY(cudaSetDevice(g->device));
  const cudaError_t e = cudaMemcpy(data, g->d_data.p, n_cells, cudaMemcpyDeviceToHost);
  if (prev >= 0) {cudaSetDevice(prev);}
  if (e != cudaSuccess) {
    ndt2d_set_error("ndt2d_occupancy_fetch", e, __FILE__, __LINE__);
    return NDT2D_ERR_CUDA;
  }
  return NDT2D_OK;
}

}  // extern "C"
