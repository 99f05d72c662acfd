// search_common.cuh -- device helpers shared by the search kernels
// (search.cu: plain per-candidate kernel, pose scoring, final reduce;
//  search_tiled.cu: the tiled production kernel).
#ifndef NDT2D_SEARCH_COMMON_CUH_
#define NDT2D_SEARCH_COMMON_CUH_

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ndt2d_internal.h"

namespace ndt2d_dev
{

constexpr double kLog2e = 1.44269504088896340736;

// ------------------------------------------------------------------ lookup
// Padded cell coordinate of x on one axis, by the reference's own arithmetic
// (NDT::getIndex, ndt_model.cpp:205-215): 0 = below the origin,
// size + 1 = at or beyond size.
__device__ __forceinline__ uint32_t padded_coord_exact(
  double v, double origin, double cell_size, uint32_t size)
{
  if (v < origin) {return 0u;}
  const uint32_t gi = __double2uint_rz(__ddiv_rn(__dsub_rn(v, origin), cell_size));
  return (gi >= size ? size : gi) + 1u;
}

// The same padded coordinate from the host-tabulated exact thresholds (api.cu:
// axis_thresholds; thr[k] = smallest double whose reference coordinate is >= k,
// thr[size + 1] = +inf): pc(v) = #{k in [0, size] : thr[k] <= v}.  The product with 1 / cell
// is only a starting guess, the thresholds fix the answer -- no division.
__device__ __forceinline__ uint32_t padded_coord_thr_g(
  double v, const double * __restrict__ thr, uint32_t size, double origin, double inv_cell)
{
  if (!(v >= __ldg(thr))) {return 0u;}
  const double q = (v - origin) * inv_cell;
  uint32_t pc = (q >= static_cast<double>(size)) ? size : static_cast<uint32_t>(q);
  pc += 1u;
  while (pc <= size && v >= __ldg(thr + pc)) {++pc;}
  while (pc > 1u && v < __ldg(thr + pc - 1u)) {--pc;}
  return pc;
}

// Likelihood of one map-frame point given its padded cell index: 0 for an
// unoccupied cell, else exp(-0.5 q^T I q) (Cell::score, ndt_model.cpp:105-116).
__device__ __forceinline__ double cell_likelihood(
  const uint2 * __restrict__ occ, const double * __restrict__ rec, uint32_t pidx, double x,
  double y)
{
  const uint2 w = occ[pidx >> 5];
  const uint32_t bit = pidx & 31u;
  if (((w.x >> bit) & 1u) == 0u) {return 0.0;}
  const uint32_t rank = w.y + __popc(w.x & ((1u << bit) - 1u));
  const double * r = rec + static_cast<size_t>(rank) * NDT2D_REC_DOUBLES;
  // exponent = ((-0.5 q^T) I) q with the reference's own grouping and no FMA
  // (Eigen evaluates it left to right, ndt_model.cpp:113-114): near-singular
  // information matrices cancel exactly where the reference's do.
  const double qx = __dsub_rn(x, r[0]), qy = __dsub_rn(y, r[1]);
  const double r0 = __dadd_rn(__dmul_rn(qx, r[2]), __dmul_rn(qy, r[3]));
  const double r1 = __dadd_rn(__dmul_rn(qx, r[4]), __dmul_rn(qy, r[5]));
  const double e = __dadd_rn(__dmul_rn(r0, qx), __dmul_rn(r1, qy));
  return static_cast<double>(exp2f(static_cast<float>(e * kLog2e)));
}

__device__ __forceinline__ double point_likelihood_exact(const ModelView & mv, double x, double y)
{
  const uint32_t ex = padded_coord_exact(x, mv.g.origin_x, mv.g.cell_size, mv.g.size_x);
  const uint32_t ey = padded_coord_exact(y, mv.g.origin_y, mv.g.cell_size, mv.g.size_y);
  return cell_likelihood(mv.occ, mv.rec, ey * mv.g.pitch + ex, x, y);
}

// ------------------------------------------------------------------ reduce
struct Best
{
  double score;
  double index;
};

// strict '<' with lowest index on ties == the reference's first-wins rule
// (scan_matcher_ndt.cpp:128); NaN never wins.
__device__ __forceinline__ void best_merge(Best & a, double score, double index)
{
  if (score < a.score || (score == a.score && index < a.index)) {
    a.score = score;
    a.index = index;
  }
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;
}

__device__ __forceinline__ void warp_best(Best & b)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double s = __shfl_xor_sync(0xffffffffu, b.score, o);
    const double i = __shfl_xor_sync(0xffffffffu, b.index, o);
    best_merge(b, s, i);
  }
}

// Block-level reduction of (best, 6 sums) into out[0..7]; all threads call.
template<int THREADS>
__device__ __forceinline__ void block_reduce_partial(Best b, double (&sum)[6], double * out)
{
  constexpr int W = THREADS / 32;
  __shared__ double red[W][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  warp_best(b);
#pragma unroll
  for (int k = 0; k < 6; ++k) {sum[k] = warp_sum(sum[k]);}
  if (lane == 0) {
    red[warp][0] = b.score;
    red[warp][1] = b.index;
#pragma unroll
    for (int k = 0; k < 6; ++k) {red[warp][2 + k] = sum[k];}
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Best t{red[0][0], red[0][1]};
    double s[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {s[k] = red[0][2 + k];}
    for (int w = 1; w < W; ++w) {
      best_merge(t, red[w][0], red[w][1]);
#pragma unroll
      for (int k = 0; k < 6; ++k) {s[k] += red[w][2 + k];}
    }
    out[0] = t.score;
    out[1] = t.index;
#pragma unroll
    for (int k = 0; k < 6; ++k) {out[2 + k] = s[k];}
  }
  __syncthreads();
}

constexpr double kNoIndex = 1.0e300;


// Same reduction for a block whose size is only known at run time (<= 1024).
__device__ __forceinline__ void block_reduce_partial_dyn(Best b, double (&sum)[6], double * out)
{
  __shared__ double red[32][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = (blockDim.x + 31) >> 5;
  warp_best(b);
#pragma unroll
  for (int k = 0; k < 6; ++k) {sum[k] = warp_sum(sum[k]);}
  if (lane == 0) {
    red[warp][0] = b.score;
    red[warp][1] = b.index;
#pragma unroll
    for (int k = 0; k < 6; ++k) {red[warp][2 + k] = sum[k];}
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Best t{red[0][0], red[0][1]};
    double s[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {s[k] = red[0][2 + k];}
    for (int w = 1; w < n_warps; ++w) {
      best_merge(t, red[w][0], red[w][1]);
#pragma unroll
      for (int k = 0; k < 6; ++k) {s[k] += red[w][2 + k];}
    }
    out[0] = t.score;
    out[1] = t.index;
#pragma unroll
    for (int k = 0; k < 6; ++k) {out[2 + k] = s[k];}
  }
  __syncthreads();
}

}  // namespace ndt2d_dev

#endif  // NDT2D_SEARCH_COMMON_CUH_
