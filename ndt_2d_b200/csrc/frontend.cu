// frontend.cu -- the step before the path: sensor_msgs/LaserScan -> ndt_2d::Scan points
// (Mapper::laserCallback, ndt_mapper.cpp:385-453): polar -> Cartesian in the laser frame,
// laser -> robot transform, per-beam motion de-skew, NaN / max-range filter, order kept.
//
// One CTA; a thread per beam, ordered stream compaction by ballots + a running offset.
// The beam angle is formed in float exactly like the reference's expression
// `msg->angle_min + i * msg->angle_increment` (size_t -> float, float multiply, float add);
// everything after it is double.  cos / sin are the device's double-precision functions
// (<= 2 ulp), so points agree with the reference to ~1e-15 relative, not bit for bit;
// which beams are kept, and their order, is exact.
//
// Also here: Graph::findNearest (graph.cpp:167-189), the candidate selection in front of the
// loop-closure batch -- a brute-force radius search over the scan positions.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "ndt2d_internal.h"

namespace
{

struct LaserParams
{
  float angle_min, angle_increment;
  double range_max;
  double lt_x, lt_y, cos_lt, sin_lt;      // laser_transform_ (:403-404)
  double tr_x, tr_y, tr_th;               // translation during the scan (:386-389)
  double pm_x, pm_y, pm_th;               // trans_per_meas (:392-395)
  int inverted;
  uint32_t n;
};

__global__ void __launch_bounds__(1024) laser_to_points_kernel(
  LaserParams P, const float * __restrict__ ranges, double2 * __restrict__ out,
  uint32_t * __restrict__ n_out)
{
  __shared__ uint32_t warp_counts[32];
  __shared__ uint32_t base_shared;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {base_shared = 0;}
  __syncthreads();
  // the inverted branch walks i = n-1 .. 1 (index 0 is never visited, :411)
  const uint32_t count = P.inverted ? (P.n > 0 ? P.n - 1u : 0u) : P.n;
  for (uint32_t k0 = 0; k0 < count; k0 += blockDim.x) {
    const uint32_t k = k0 + threadIdx.x;
    bool keep = false;
    double2 pt = make_double2(0.0, 0.0);
    if (k < count) {
      const uint32_t i = P.inverted ? (P.n - 1u - k) : k;
      const float r = ranges[i];
      // Filter out NANs and scans beyond max range (:414, :436)
      keep = !(isnan(r) || static_cast<double>(r) > P.range_max);
      if (keep) {
        const float af = __fadd_rn(P.angle_min, __fmul_rn(static_cast<float>(i), P.angle_increment));
        const double angle = P.inverted ? -static_cast<double>(af) : static_cast<double>(af);
        const double di = static_cast<double>(i);
        double sa, ca;
        sincos(angle, &sa, &ca);
        const double lx = __dmul_rn(ca, static_cast<double>(r)), ly = __dmul_rn(sa, static_cast<double>(r));
        // robot frame (:420-421, :442-443)
        const double px = __dadd_rn(__dsub_rn(__dmul_rn(P.cos_lt, lx), __dmul_rn(P.sin_lt, ly)), P.lt_x);
        const double py = __dadd_rn(__dadd_rn(__dmul_rn(P.sin_lt, lx), __dmul_rn(P.cos_lt, ly)), P.lt_y);
        // laser movement (:423-427, :445-448)
        const double th = P.inverted ? __dsub_rn(P.tr_th, __dmul_rn(P.pm_th, di)) : __dmul_rn(P.pm_th, di);
        const double ox = P.inverted ? __dsub_rn(P.tr_x, __dmul_rn(P.pm_x, di)) : __dmul_rn(P.pm_x, di);
        const double oy = P.inverted ? __dsub_rn(P.tr_y, __dmul_rn(P.pm_y, di)) : __dmul_rn(P.pm_y, di);
        double st, ct;
        sincos(th, &st, &ct);
        pt.x = __dadd_rn(__dsub_rn(__dmul_rn(ct, px), __dmul_rn(st, py)), ox);
        pt.y = __dadd_rn(__dadd_rn(__dmul_rn(st, px), __dmul_rn(ct, py)), oy);
      }
    }
    // ordered compaction of this batch of blockDim.x beams
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) {warp_counts[warp] = __popc(bal);}
    __syncthreads();
    uint32_t before = base_shared;
    for (uint32_t w = 0; w < warp; ++w) {before += warp_counts[w];}
    if (keep) {out[before + __popc(bal & ((1u << lane) - 1u))] = pt;}
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
      for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) {t += warp_counts[w];}
      base_shared += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {*n_out = base_shared;}
}

// Graph::findNearest (graph.cpp:167-189): every scan position closer than the (squared)
// radius to the query.  A thread per scan; the squared distance is accumulated as
// nanoflann's L2_Simple_Adaptor::evalMetric does (0 + dx*dx, then + dy*dy, no FMA), matches
// are appended through one warp-aggregated atomic.  The host orders them by (distance, index).
__global__ void __launch_bounds__(256) find_nearest_kernel(
  const double2 * __restrict__ scan_xy, uint32_t n, double qx, double qy, double radius_sq,
  double * __restrict__ out_dist, uint32_t * __restrict__ out_index, uint32_t * __restrict__ n_out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  bool match = false;
  double d = 0.0;
  if (i < n) {
    const double2 p = scan_xy[i];
    const double dx = __dsub_rn(qx, p.x), dy = __dsub_rn(qy, p.y);
    d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    match = d < radius_sq;   // RadiusResultSet::addPoint keeps dist < radius (strict)
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, match);
  if (bal == 0u) {return;}
  uint32_t base = 0;
  if (lane == 0) {base = atomicAdd(n_out, static_cast<uint32_t>(__popc(bal)));}
  base = __shfl_sync(0xffffffffu, base, 0);
  if (match) {
    const uint32_t slot = base + __popc(bal & ((1u << lane) - 1u));
    out_dist[slot] = d;
    out_index[slot] = i;
  }
}

}  // namespace

extern "C" {

NDT2D_API int ndt2d_laser_to_points(
  int device, const float * ranges, size_t n, float angle_min, float angle_increment,
  double range_max, const double * laser_tf3, const double * translation3, int laser_inverted,
  double * out_pts_xy, size_t * n_out)
{
  if (!laser_tf3 || !translation3 || !n_out || (n && (!ranges || !out_pts_xy)) || n >= (1u << 30)) {
    return NDT2D_ERR_INVALID;
  }
  *n_out = 0;
  if (ndt2d_device_count() <= 0) {return NDT2D_ERR_NO_DEVICE;}
  if (n == 0) {return NDT2D_OK;}
  int prev = -1;
  cudaGetDevice(&prev);
  if (device >= 0 && device != prev) {NDT2D_CUDA_TRY(cudaSetDevice(device));}
  LaserParams P;
  P.angle_min = angle_min;
  P.angle_increment = angle_increment;
  P.range_max = range_max;
  P.lt_x = laser_tf3[0];
  P.lt_y = laser_tf3[1];
  P.cos_lt = cos(laser_tf3[2]);   // host libm, as the reference's "minor optimization" (:403-404)
  P.sin_lt = sin(laser_tf3[2]);
  P.tr_x = translation3[0];
  P.tr_y = translation3[1];
  P.tr_th = translation3[2];
  P.pm_x = translation3[0] / static_cast<double>(n);
  P.pm_y = translation3[1] / static_cast<double>(n);
  P.pm_th = translation3[2] / static_cast<double>(n);
  P.inverted = laser_inverted ? 1 : 0;
  P.n = static_cast<uint32_t>(n);
  float * d_ranges = nullptr;
  double2 * d_out = nullptr;
  uint32_t * d_n = nullptr;
  int rc = NDT2D_OK;
  uint32_t h_n = 0;
  do {
    if (cudaMalloc(&d_ranges, n * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d_out, n * sizeof(double2)) != cudaSuccess ||
      cudaMalloc(&d_n, sizeof(uint32_t)) != cudaSuccess)
    {
      rc = NDT2D_ERR_CUDA;
      break;
    }
    if (cudaMemcpy(d_ranges, ranges, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
      rc = NDT2D_ERR_CUDA;
      break;
    }
    laser_to_points_kernel<<<1, 1024>>>(P, d_ranges, d_out, d_n);
    if (cudaGetLastError() != cudaSuccess ||
      cudaMemcpy(&h_n, d_n, sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess)
    {
      rc = NDT2D_ERR_CUDA;
      break;
    }
    if (h_n && cudaMemcpy(out_pts_xy, d_out, static_cast<size_t>(h_n) * sizeof(double2),
      cudaMemcpyDeviceToHost) != cudaSuccess)
    {
      rc = NDT2D_ERR_CUDA;
      break;
    }
  } while (false);
  if (rc == NDT2D_ERR_CUDA) {ndt2d_set_error("ndt2d_laser_to_points", cudaGetLastError(), __FILE__, __LINE__);}
  cudaFree(d_ranges);
  cudaFree(d_out);
  cudaFree(d_n);
  if (prev >= 0 && device >= 0 && device != prev) {cudaSetDevice(prev);}
  if (rc == NDT2D_OK) {*n_out = h_n;}
  return rc;
}

NDT2D_API int ndt2d_find_nearest(
  int device, const double * scan_xy, size_t n_scans, int64_t limit_scan_index,
  const double * query_xy, double radius_sq, uint64_t * out_indices, double * out_dist_sq,
  size_t capacity, size_t * n_found)
{
  if (!query_xy || !n_found || (n_scans && !scan_xy) || (capacity && !out_indices) ||
    n_scans >= (1u << 30))
  {
    return NDT2D_ERR_INVALID;
  }
  *n_found = 0;
  if (ndt2d_device_count() <= 0) {return NDT2D_ERR_NO_DEVICE;}
  // graph.cpp:171: limit_scan_index > 0 restricts the search to scans [0, limit)
  size_t n = n_scans;
  if (limit_scan_index > 0 && static_cast<uint64_t>(limit_scan_index) < n_scans) {
    n = static_cast<size_t>(limit_scan_index);
  }
  if (n == 0) {return NDT2D_OK;}
  int prev = -1;
  cudaGetDevice(&prev);
  if (device >= 0 && device != prev) {NDT2D_CUDA_TRY(cudaSetDevice(device));}
  double2 * d_xy = nullptr;
  double * d_dist = nullptr;
  uint32_t * d_idx = nullptr;
  uint32_t * d_n = nullptr;
  int rc = NDT2D_OK;
  uint32_t h_n = 0;
  std::vector<double> dist;
  std::vector<uint32_t> idx;
  do {
    if (cudaMalloc(&d_xy, n * sizeof(double2)) != cudaSuccess ||
      cudaMalloc(&d_dist, n * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&d_idx, n * sizeof(uint32_t)) != cudaSuccess ||
      cudaMalloc(&d_n, sizeof(uint32_t)) != cudaSuccess)
    {
      rc = NDT2D_ERR_CUDA;
      break;
    }
    if (cudaMemcpy(d_xy, scan_xy, n * sizeof(double2), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemset(d_n, 0, sizeof(uint32_t)) != cudaSuccess)
    {
      rc = NDT2D_ERR_CUDA;
      break;
    }
    find_nearest_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(
      d_xy, static_cast<uint32_t>(n), query_xy[0], query_xy[1], radius_sq, d_dist, d_idx, d_n);
    if (cudaGetLastError() != cudaSuccess ||
      cudaMemcpy(&h_n, d_n, sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess)
    {
      rc = NDT2D_ERR_CUDA;
      break;
    }
    dist.resize(h_n);
    idx.resize(h_n);
    if (h_n && (cudaMemcpy(dist.data(), d_dist, h_n * sizeof(double), cudaMemcpyDeviceToHost) !=
      cudaSuccess ||
      cudaMemcpy(idx.data(), d_idx, h_n * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess))
    {
      rc = NDT2D_ERR_CUDA;
      break;
    }
  } while (false);
  if (rc == NDT2D_ERR_CUDA) {ndt2d_set_error("ndt2d_find_nearest", cudaGetLastError(), __FILE__, __LINE__);}
  cudaFree(d_xy);
  cudaFree(d_dist);
  cudaFree(d_idx);
  cudaFree(d_n);
  if (prev >= 0 && device >= 0 && device != prev) {cudaSetDevice(prev);}
  if (rc != NDT2D_OK) {return rc;}
  // nanoflann returns the matches sorted by ascending distance (SearchParams::sorted); the
  // append order above is arbitrary, (distance, index) makes the result deterministic
  std::vector<uint32_t> order(h_n);
  for (uint32_t k = 0; k < h_n; ++k) {order[k] = k;}
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
      return dist[a] < dist[b] || (dist[a] == dist[b] && idx[a] < idx[b]);
    });
  for (size_t k = 0; k < h_n && k < capacity; ++k) {
    out_indices[k] = idx[order[k]];
    if (out_dist_sq) {out_dist_sq[k] = dist[order[k]];}
  }
  *n_found = h_n;
  return NDT2D_OK;
}

}  // extern "C"
