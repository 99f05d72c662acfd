// search_tiled.cu -- the round's FIRST production correlative-search kernel, kept as
// kernel_variant 2 for A/B runs (260 ms at config 4 against 29 ms for search_region.cu).
//
// Replaces the three nested loops of ScanMatcherNDT::matchScan
// (scan_matcher_ndt.cpp:103-143).  One CTA scores a TILE of (dx, dy)
// candidates for one theta slice; one thread owns an R x R PATCH of adjacent
// candidates and keeps their R*R score sums in registers while it walks the
// scan points in order (so every candidate's sum is accumulated in the
// reference's point order).
//
// What makes it fast is what it does NOT do per (candidate, point):
//
//  * no divisions.  For a fixed point, x = outer.x + dx depends only on the
//    candidate's column and y only on its row, and the reference's cell
//    coordinate unsigned((x - origin) / cell) is a monotone step function of
//    x.  The host tabulates the exact step positions (thr_x / thr_y: the
//    smallest double whose reference coordinate is >= k), so a point's padded
//    cell coordinate is found by comparing against thresholds -- bit-exact,
//    never by re-doing the division.  Per chunk of points each CTA walks the
//    patch starts of its tile once per axis (GX / GY tables in shared memory).
//  * no per-candidate work for empty space.  (R-1) * step < cell, so the
//    candidates of a patch can put a point in at most the 2 x 2 cells starting
//    at the patch's first cell.  The model carries a DILATED occupancy bitmap
//    D[c] = E[c] | E[c+1] | E[c+pitch] | E[c+pitch+1]; one bit test of D
//    rejects the point for all R*R candidates.  Most of a large search is
//    empty space.
//  * only on a D hit: exact per-column / per-row cell assignment from the
//    precomputed crossing counts (NX / NY), occupancy test in E, record fetch,
//    and the Gaussian with the reference's own operation order
//    (cell_gaussian in search_common.cuh).
//
// Staging: D and E (+ rank prefix) are bulk-copied into shared memory with
// cp.async.bulk (TMA, SASS UBLKCP) signalled through an mbarrier when the grid
// is small enough; otherwise they are read through L1/L2.  Cell records stay
// in global memory (48 B, L1-resident for the cells a tile touches).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ndt2d_internal.h"
#include "search_common.cuh"

namespace
{

using namespace ndt2d_dev;

constexpr int kMaxR = 5;
constexpr uint32_t kMaxPatchesPerAxis = 20;     // 20 x 20 patches = 400 threads
constexpr uint32_t kChunkPoints = 256;          // scan points staged per pass
constexpr size_t kSmemOccBudget = 96 * 1024;    // D + E in shared memory up to this size

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void * p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
    "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void * dst, const void * src, uint32_t bytes,
  uint64_t * bar)
{
  asm volatile(
    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
    ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t phase)
{
  uint32_t done = 0;
  while (!done) {
    asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  }
}

// ---------------------------------------------------------------- cell coordinate
// Padded coordinate pc(v) = #{k in [0, size] : thr[k] <= v}  (0 = below the
// origin, size + 1 = beyond the grid).  The product with 1/cell only provides a
// starting guess; the answer is fixed by the thresholds.
__device__ __forceinline__ uint32_t padded_coord_thr(
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

struct TileShared
{
  double * dl_x;     // P*R candidate offsets of this tile, x
  double * dl_y;     // P*R, y
  double2 * outer;   // PC rotated + translated points
  uint32_t * gx;     // PC*P  padded x coordinate of the patch's first column
  uint32_t * gy;     // PC*P  padded y coordinate of the patch's first row, times pitch
  uint8_t * nx;      // PC*P  columns of the patch still in the first cell (R = no crossing)
  uint8_t * ny;      // PC*P
  uint32_t * occd;   // dilated occupancy (shared copy or global)
  uint2 * occ;       // exact occupancy + rank prefix
};

// One axis of the GX/GY tables for one point: walk the P patch starts.
template<int R>
__device__ __forceinline__ void walk_axis(
  double o, const double * __restrict__ dl, uint32_t P, const double * __restrict__ thr,
  uint32_t size, double origin, double inv_cell, uint32_t scale, uint32_t * __restrict__ g_out,
  uint8_t * __restrict__ n_out)
{
  uint32_t pc = padded_coord_thr(__dadd_rn(o, dl[0]), thr, size, origin, inv_cell);
  for (uint32_t k = 0; k < P; ++k) {
    // coordinate of the LAST column of patch k (monotone: >= pc)
    uint32_t pcl = pc;
    if (R > 1) {
      const double vl = __dadd_rn(o, dl[k * R + (R - 1)]);
      while (pcl <= size && vl >= __ldg(thr + pcl)) {++pcl;}
    }
    uint32_t n_in = R;
    if (R > 1 && pcl != pc) {
      // the patch crosses into the next cell: count the columns still in the first
      const double t = __ldg(thr + pc);
      n_in = 1;
#pragma unroll
      for (int a = 1; a < R - 1; ++a) {
        n_in += (__dadd_rn(o, dl[k * R + a]) < t) ? 1u : 0u;
      }
    }
    g_out[k] = pc * scale;
    n_out[k] = static_cast<uint8_t>(n_in);
    if (k + 1 < P) {
      const double vn = __dadd_rn(o, dl[(k + 1) * R]);
      pc = pcl;
      while (pc <= size && vn >= __ldg(thr + pc)) {++pc;}
    }
  }
}

template<int R, bool SMEM_OCC>
__global__ void __launch_bounds__(kMaxPatchesPerAxis * kMaxPatchesPerAxis + 16)
search_tiled_kernel(
  ModelView mv, SearchView sv, uint32_t theta_begin, uint32_t P, uint32_t tiles_per_axis,
  double * __restrict__ block_partials, double * __restrict__ scores)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // ---- carve shared memory
  uint64_t * mbar = reinterpret_cast<uint64_t *>(smem_raw);
  unsigned char * sp = smem_raw + 16;
  TileShared ts;
  ts.dl_x = reinterpret_cast<double *>(sp);
  sp += sizeof(double) * P * R;
  ts.dl_y = reinterpret_cast<double *>(sp);
  sp += sizeof(double) * P * R;
  sp = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(sp) + 15) & ~uintptr_t(15));
  ts.outer = reinterpret_cast<double2 *>(sp);
  sp += sizeof(double2) * kChunkPoints;
  ts.gx = reinterpret_cast<uint32_t *>(sp);
  sp += sizeof(uint32_t) * kChunkPoints * P;
  ts.gy = reinterpret_cast<uint32_t *>(sp);
  sp += sizeof(uint32_t) * kChunkPoints * P;
  ts.nx = sp;
  sp += kChunkPoints * P;
  ts.ny = sp;
  sp += kChunkPoints * P;
  sp = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(sp) + 15) & ~uintptr_t(15));
  const uint32_t occd_bytes = ((mv.g.n_words * 4u) + 15u) & ~15u;
  const uint32_t occ_bytes = ((mv.g.n_words * 8u) + 15u) & ~15u;
  if (SMEM_OCC) {
    ts.occd = reinterpret_cast<uint32_t *>(sp);
    sp += occd_bytes;
    ts.occ = reinterpret_cast<uint2 *>(sp);
  } else {
    ts.occd = const_cast<uint32_t *>(mv.occ_dilated);
    ts.occ = const_cast<uint2 *>(mv.occ);
  }

  const uint32_t tid = threadIdx.x;
  const uint32_t itheta = theta_begin + blockIdx.y * sv.theta_stride;
  const uint32_t tile_x = blockIdx.x / tiles_per_axis, tile_y = blockIdx.x - tile_x * tiles_per_axis;
  const uint32_t n_lin = sv.n_lin;
  const uint32_t jx0 = tile_x * P * R, jy0 = tile_y * P * R;
  const bool active = tid < P * P;
  const uint32_t ty = active ? tid / P : 0u, tx = active ? tid - ty * P : 0u;

  // ---- stage the occupancy bitmaps with TMA bulk copies (one elected thread)
  if (SMEM_OCC && tid == 0) {
    mbar_init(mbar, 1);
    mbar_expect_tx(mbar, occd_bytes + occ_bytes);
    bulk_copy_g2s(ts.occd, mv.occ_dilated, occd_bytes, mbar);
    bulk_copy_g2s(ts.occ, mv.occ, occ_bytes, mbar);
  }
  for (uint32_t k = tid; k < P * R; k += blockDim.x) {
    ts.dl_x[k] = sv.dlin[min(jx0 + k, n_lin - 1u)];
    ts.dl_y[k] = sv.dlin[min(jy0 + k, n_lin - 1u)];
  }
  const double2 cs = sv.trig[itheta];
  const double inv_cell = 1.0 / mv.g.cell_size;
  const uint32_t pitch = mv.g.pitch;

  double acc[R * R];
#pragma unroll
  for (int k = 0; k < R * R; ++k) {acc[k] = 0.0;}

  bool occ_ready = !SMEM_OCC;
  for (uint32_t p0 = 0; p0 < sv.n_pts; p0 += kChunkPoints) {
    const uint32_t np = min(kChunkPoints, sv.n_pts - p0);
    const uint32_t np_pad = (np + 3u) & ~3u;
    __syncthreads();  // previous chunk fully consumed (also publishes dl_x / dl_y)
    // outer = (p.x*c - p.y*s) + pose.x , (p.x*s + p.y*c) + pose.y   (scan_matcher_ndt.cpp:111-114)
    for (uint32_t i = tid; i < np; i += blockDim.x) {
      const double2 p = sv.pts[p0 + i];
      double2 o;
      o.x = __dadd_rn(__dsub_rn(__dmul_rn(p.x, cs.x), __dmul_rn(p.y, cs.y)), sv.pose_x);
      o.y = __dadd_rn(__dadd_rn(__dmul_rn(p.x, cs.y), __dmul_rn(p.y, cs.x)), sv.pose_y);
      ts.outer[i] = o;
    }
    __syncthreads();
    // GX / GY tables: one task per (point, axis)
    for (uint32_t t = tid; t < 2u * np_pad; t += blockDim.x) {
      const uint32_t axis = t >= np_pad ? 1u : 0u;
      const uint32_t i = t - axis * np_pad;
      uint32_t * g_out = (axis ? ts.gy : ts.gx) + i * P;
      uint8_t * n_out = (axis ? ts.ny : ts.nx) + i * P;
      if (i >= np) {
        // padding rows: padded cell 0 with "no crossing" can never reach an occupied cell
        for (uint32_t k = 0; k < P; ++k) {
          g_out[k] = 0u;
          n_out[k] = R;
        }
      } else if (axis == 0) {
        walk_axis<R>(ts.outer[i].x, ts.dl_x, P, mv.thr_x, mv.g.size_x, mv.g.origin_x, inv_cell, 1u,
          g_out, n_out);
      } else {
        walk_axis<R>(ts.outer[i].y, ts.dl_y, P, mv.thr_y, mv.g.size_y, mv.g.origin_y, inv_cell,
          pitch, g_out, n_out);
      }
    }
    __syncthreads();
    if (!occ_ready) {
      mbar_wait(mbar, 0);
      occ_ready = true;
    }
    if (active) {
#pragma unroll 1
      for (uint32_t i0 = 0; i0 < np_pad; i0 += 4) {
        uint32_t hit = 0;
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) {
          const uint32_t idx = ts.gx[(i0 + u) * P + tx] + ts.gy[(i0 + u) * P + ty];
          const uint32_t w = ts.occd[idx >> 5];
          hit |= ((w >> (idx & 31u)) & 1u) << u;
        }
        while (hit) {
          const uint32_t i = i0 + __ffs(hit) - 1u;
          hit &= hit - 1u;
          // ---- some cell of the 2 x 2 neighbourhood is occupied: exact evaluation
          const uint32_t base = ts.gx[i * P + tx] + ts.gy[i * P + ty];
          const uint32_t nx = ts.nx[i * P + tx], ny = ts.ny[i * P + ty];
          const double2 o = ts.outer[i];
          double ys[R];
#pragma unroll
          for (int b = 0; b < R; ++b) {ys[b] = __dadd_rn(o.y, ts.dl_y[ty * R + b]);}
#pragma unroll
          for (uint32_t vy = 0; vy < 2; ++vy) {
            if (vy == 1 && ny == R) {continue;}
#pragma unroll
            for (uint32_t vx = 0; vx < 2; ++vx) {
              if (vx == 1 && nx == R) {continue;}
              const uint32_t cidx = base + vx + vy * pitch;
              const uint2 w = ts.occ[cidx >> 5];
              const uint32_t bit = cidx & 31u;
              if (((w.x >> bit) & 1u) == 0u) {continue;}
              const uint32_t rank = w.y + __popc(w.x & ((1u << bit) - 1u));
              const double2 * r2 = reinterpret_cast<const double2 *>(
                mv.rec + static_cast<size_t>(rank) * NDT2D_REC_DOUBLES);
              const double2 mean = __ldg(r2), i0010 = __ldg(r2 + 1), i0111 = __ldg(r2 + 2);
              // row terms: qy, qy*I(1,0), qy*I(1,1)   (already scaled by -0.5)
              double qy[R], t10[R], t11[R];
#pragma unroll
              for (int b = 0; b < R; ++b) {
                qy[b] = __dsub_rn(ys[b], mean.y);
                t10[b] = __dmul_rn(qy[b], i0010.y);
                t11[b] = __dmul_rn(qy[b], i0111.y);
              }
#pragma unroll
              for (int a = 0; a < R; ++a) {
                const bool in_x = vx ? (static_cast<uint32_t>(a) >= nx) : (static_cast<uint32_t>(a) < nx);
                if (!in_x) {continue;}
                const double qx = __dsub_rn(__dadd_rn(o.x, ts.dl_x[tx * R + a]), mean.x);
                const double t00 = __dmul_rn(qx, i0010.x), t01 = __dmul_rn(qx, i0111.x);
#pragma unroll
                for (int b = 0; b < R; ++b) {
                  const bool in_y = vy ? (static_cast<uint32_t>(b) >= ny) : (static_cast<uint32_t>(b) < ny);
                  if (in_y) {
                    // ((-0.5 q)^T I) q, reference grouping (ndt_model.cpp:113-114)
                    const double r0 = __dadd_rn(t00, t10[b]);
                    const double r1 = __dadd_rn(t01, t11[b]);
                    const double e = __dadd_rn(__dmul_rn(r0, qx), __dmul_rn(r1, qy[b]));
                    acc[a * R + b] += static_cast<double>(exp2f(static_cast<float>(e * kLog2e)));
                  }
                }
              }
            }
          }
        }
      }
    }
  }

  // ---- epilogue: per-candidate score, block partial
  Best best{0.0, kNoIndex};
  double sum[6] = {0, 0, 0, 0, 0, 0};
  if (active) {
    const uint64_t n_cand = static_cast<uint64_t>(n_lin) * n_lin;
#pragma unroll
    for (int a = 0; a < R; ++a) {
      const uint32_t jx = jx0 + tx * R + a;
#pragma unroll
      for (int b = 0; b < R; ++b) {
        const uint32_t jy = jy0 + ty * R + b;
        if (jx < n_lin && jy < n_lin) {
          const double score = -acc[a * R + b];
          const double dx = ts.dl_x[tx * R + a], dy = ts.dl_y[ty * R + b];
          const uint64_t gi = static_cast<uint64_t>(itheta) * n_cand + static_cast<uint64_t>(jx) * n_lin + jy;
          if (scores) {scores[gi] = score;}
          best_merge(best, score, static_cast<double>(gi));
          sum[0] += score;
          sum[1] += dx * score;
          sum[2] += dy * score;
          sum[3] += (dx * dx) * score;
          sum[4] += (dx * dy) * score;
          sum[5] += (dy * dy) * score;
        }
      }
    }
  }
  double * out = block_partials +
    (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * NDT2D_BLOCK_PARTIAL;
  block_reduce_partial_dyn(best, sum, out);
  if (tid == 0) {out[8] = sv.dth[itheta];}
}

struct TiledPlan
{
  int R;
  uint32_t P, tiles_per_axis, threads;
  bool smem_occ;
  size_t smem_bytes;
};

TiledPlan make_plan(const GridDesc & g, uint32_t n_lin, double linear_res)
{
  TiledPlan pl;
  // R * step <= cell keeps a patch inside a 2 x 2 cell neighbourhood
  int R = 1;
  if (linear_res > 0.0) {
    const double ratio = g.cell_size / linear_res * (1.0 - 1e-9);
    R = ratio >= kMaxR ? kMaxR : (ratio < 1.0 ? 1 : static_cast<int>(ratio));
  }
  // small searches: prefer more threads over bigger patches
  while (R > 1 && (n_lin / R) * (n_lin / R) < 64) {--R;}
  pl.R = R;
  const uint32_t np = (n_lin + R - 1) / R;                       // patches per axis
  const uint32_t nt = (np + kMaxPatchesPerAxis - 1) / kMaxPatchesPerAxis;
  pl.tiles_per_axis = nt;
  pl.P = (np + nt - 1) / nt;
  pl.threads = ((pl.P * pl.P + 31u) / 32u) * 32u;
  const size_t occ_bytes = (((size_t)g.n_words * 4 + 15) & ~size_t(15)) +
    (((size_t)g.n_words * 8 + 15) & ~size_t(15));
  pl.smem_occ = occ_bytes <= kSmemOccBudget;
  size_t b = 16 + 2 * sizeof(double) * pl.P * R + 16 + sizeof(double2) * kChunkPoints +
    2 * sizeof(uint32_t) * kChunkPoints * pl.P + 2 * (size_t)kChunkPoints * pl.P + 16;
  if (pl.smem_occ) {b += occ_bytes;}
  pl.smem_bytes = b;
  return pl;
}

template<int R, bool S>
int launch_one(const TiledPlan & pl, const ModelView & mv, const SearchView & sv,
  uint32_t theta_begin, uint32_t n_theta, double * d_block_partials, double * d_scores,
  cudaStream_t stream, Counters * ctr)
{
  auto kernel = search_tiled_kernel<R, S>;
  NDT2D_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
    static_cast<int>(pl.smem_bytes)));
  const uint32_t tiles = pl.tiles_per_axis * pl.tiles_per_axis;
  uint32_t done = 0;
  while (done < n_theta) {
    const uint32_t ny = min(n_theta - done, 65535u);
    dim3 grid(tiles, ny);
    kernel<<<grid, pl.threads, pl.smem_bytes, stream>>>(
      mv, sv, theta_begin + done * sv.theta_stride, pl.P, pl.tiles_per_axis,
      d_block_partials + static_cast<size_t>(done) * tiles * NDT2D_BLOCK_PARTIAL, d_scores);
    NDT2D_LAUNCH_CHECK(ctr);
    done += ny;
  }
  return NDT2D_OK;
}

}  // namespace

size_t ndt2d_tiled_scratch_doubles(const GridDesc & g, uint32_t n_ang, uint32_t n_lin,
  double linear_res)
{
  const TiledPlan pl = make_plan(g, n_lin ? n_lin : 1, linear_res);
  return static_cast<size_t>(n_ang ? n_ang : 1) * pl.tiles_per_axis * pl.tiles_per_axis *
         NDT2D_BLOCK_PARTIAL;
}

int ndt2d_launch_search_tiled(
  const ModelView & mv, const SearchView & sv, double linear_res, uint32_t theta_begin,
  uint32_t n_theta, double * d_block_partials, double * d_scores, cudaStream_t stream,
  Counters * ctr, uint32_t * n_blocks)
{
  const TiledPlan pl = make_plan(mv.g, sv.n_lin, linear_res);
  *n_blocks = n_theta * pl.tiles_per_axis * pl.tiles_per_axis;
#define NDT2D_TILED_CASE(RR)                                                                   \
  case RR:                                                                                     \
    return pl.smem_occ                                                                         \
           ? launch_one<RR, true>(pl, mv, sv, theta_begin, n_theta, d_block_partials, d_scores, \
             stream, ctr)                                                                      \
           : launch_one<RR, false>(pl, mv, sv, theta_begin, n_theta, d_block_partials,         \
             d_scores, stream, ctr);
  switch (pl.R) {
    NDT2D_TILED_CASE(1)
    NDT2D_TILED_CASE(2)
    NDT2D_TILED_CASE(3)
    NDT2D_TILED_CASE(4)
    NDT2D_TILED_CASE(5)
  }
#undef NDT2D_TILED_CASE
  return NDT2D_ERR_INVALID;
}
