// search_window.cu -- the correlative search for SMALL WINDOWS (K4, local matching).
//
// Same contract as search_region.cu / search.cu (ScanMatcherNDT::matchScan's nested loops,
// scan_matcher_ndt.cpp:103-143, with NDT::likelihood / getIndex / Cell::score inside,
// ndt_model.cpp:105-116, 162-187, 203-218), for searches whose linear window is a few
// cells wide: the rolling-window match of the node (ndt_mapper.cpp:508-515), the plugin's
// default window (21 x 21 steps of 5 mm) and every job of the loop-closure batch.
//
// There the candidates of one theta slice put a scan point into at most K x K cells,
// K = floor(window / cell) + 2 <= 4, so everything that depends on the point alone is done
// ONCE per (theta slice, point) and shared by all candidates:
//   phase A (thread per point): rotate + translate with the reference's operation order,
//            exact padded cell of the window's first column / row (threshold tables);
//            then, for every candidate column ix, which of the K columns of cells the
//            exact coordinate outer.x + dlin[ix] (the reference's own addition) falls in
//            -- one byte per ix, likewise per row iy --, plus occupancy + record rank of
//            the K x K cells -> shared memory;
//   phase B (thread per candidate): cell = the point's byte for its ix + the byte for its
//            iy (exact, no floating point), record rank from the point's table, Gaussian
//            exactly as the dense kernel evaluates it (double
//            differences and quadratic form, 2^t on the SFU), summed in scan-point order
//            like the reference's own loop (float blocks of 8 points into a double).
// A CTA = a tile of <= 128 candidates of one theta slice x G groups of threads that split
// the scan points between them (their partial sums are added in group = point order at the
// end: a local match is latency-bound, the shorter serial loop matters more than anything
// else); it leaves one 9-double record for the final reduction (same records as the dense
// kernel).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <type_traits>

#include "ndt2d_internal.h"
#include "search_common.cuh"

namespace
{

using namespace ndt2d_dev;

constexpr uint32_t kWinTile = 128;               // candidates per CTA (at most)
constexpr uint32_t kWinMaxThreads = 512;         // tile x point groups, rounded up to a warp
constexpr uint32_t kWinBatchThreads = 256;       // the same for the CTAs of a batch launch
constexpr uint32_t kWinMaxK = 4;
constexpr uint32_t kWinMaxLin = 64;              // steps per axis (byte tables per point)
constexpr uint32_t kWinSmemBytes = 32 * 1024;    // point tables of one pass (+ 11.5 KB static < 48 KB)
constexpr uint32_t kWinBlockPts = 8;             // points per float accumulation block

// One shared-memory record per scan point (16-byte aligned, all tables of a point side by
// side so that the candidate loop advances ONE pointer per point):
//   +0            outer (double2)
//   +16           K * K record ranks (int32, -1 = unoccupied)
//   +16 + 4 K K   padded index of the first cell (uint32)
//   +20 + 4 K K   byte per candidate column: cell column 0 .. K-1   (LT bytes, LT = n_lin up to a multiple of 4)
//   ... + LT      byte per candidate row: cell row * K              (LT bytes)
__host__ __device__ inline uint32_t win_point_bytes(uint32_t K, uint32_t n_lin)
{
  return (20u + 4u * K * K + 2u * ((n_lin + 3u) & ~3u) + 15u) & ~15u;
}
__host__ __device__ inline uint32_t win_pass_points(uint32_t K, uint32_t n_lin)
{
  return (kWinSmemBytes / win_point_bytes(K, n_lin)) & ~31u;
}

template<uint32_t K>
__device__ __forceinline__ void window_block(
  const ModelView & mv, const SearchView & sv, uint32_t theta_begin, uint32_t tile,
  uint32_t n_groups, double * __restrict__ block_partials, double * __restrict__ scores)
{
  constexpr uint32_t KK = K * K;
  const uint32_t n_lin = sv.n_lin, n_cand = n_lin * n_lin;
  const uint32_t LT = (n_lin + 3u) & ~3u;                     // bytes per axis table
  const uint32_t PP = win_pass_points(K, n_lin);
  extern __shared__ __align__(16) unsigned char win_smem[];
  const uint32_t S = win_point_bytes(K, n_lin);                // bytes per point record
  constexpr uint32_t kOffRank = 16u, kOffBase = 16u + 4u * KK, kOffKx = 20u + 4u * KK;
  const uint32_t off_ky = kOffKx + LT;
  __shared__ double group_sums[kWinMaxThreads];
  __shared__ double dlin_s[kWinMaxLin];

  const uint32_t itheta = theta_begin + blockIdx.y * sv.theta_stride;
  const double2 cs = sv.trig[itheta];
  const uint32_t size_x = mv.g.size_x, size_y = mv.g.size_y, pitch = mv.g.pitch;
  const double inv_cell = 1.0 / mv.g.cell_size;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  // (in registers: for a batch launch mv / sv live in shared memory)
  const double * const rec_fast = mv.rec_fast;
  const uint2 * const occ = mv.occ;
  const bool any_stiff = mv.n_stiff == nullptr || __ldg(mv.n_stiff) != 0u;

  for (uint32_t k = threadIdx.x; k < n_lin; k += blockDim.x) {dlin_s[k] = sv.dlin[k];}

  const bool worker = threadIdx.x < tile * n_groups;
  const uint32_t group = worker ? threadIdx.x / tile : 0u;
  const uint32_t lc = threadIdx.x - group * tile;
  const uint32_t c = blockIdx.x * tile + lc;
  const bool active = worker && c < n_cand;
  const uint32_t ix = active ? c / n_lin : 0u, iy = active ? c - ix * n_lin : 0u;
  const double dx = sv.dlin[ix], dy = sv.dlin[iy];
  double acc = 0.0;

  for (uint32_t p0 = 0; p0 < sv.n_pts; p0 += PP) {
    const uint32_t np = min(PP, sv.n_pts - p0);
    __syncthreads();
    // ---- phase A: everything that depends on (theta, point) only
    for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
      const double2 p = sv.pts[p0 + i];
      double2 o;
      // outer = (p.x*c - p.y*s) + pose.x , (p.x*s + p.y*c) + pose.y  (scan_matcher_ndt.cpp:111-114)
      o.x = __dadd_rn(__dsub_rn(__dmul_rn(p.x, cs.x), __dmul_rn(p.y, cs.y)), sv.pose_x);
      o.y = __dadd_rn(__dadd_rn(__dmul_rn(p.x, cs.y), __dmul_rn(p.y, cs.x)), sv.pose_y);
      unsigned char * const rec_i = win_smem + static_cast<size_t>(i) * S;
      *reinterpret_cast<double2 *>(rec_i) = o;
      // padded cell of the window's first column / row
      const uint32_t bx = padded_coord_thr_g(__dadd_rn(o.x, dlin_s[0]), mv.thr_x, size_x, mv.g.origin_x, inv_cell);
      const uint32_t by = padded_coord_thr_g(__dadd_rn(o.y, dlin_s[0]), mv.thr_y, size_y, mv.g.origin_y, inv_cell);
      // upper bounds of the cells bx .. bx + K - 2 (thr[size + 1] = +inf: nothing lies beyond)
      double tx[K - 1u], ty[K - 1u];
#pragma unroll
      for (uint32_t j = 0; j < K - 1u; ++j) {
        tx[j] = (bx + j <= size_x + 1u) ? __ldg(mv.thr_x + bx + j) : inf;
        ty[j] = (by + j <= size_y + 1u) ? __ldg(mv.thr_y + by + j) : inf;
      }
      // cell column / row of every candidate column / row: the candidate's coordinate is
      // outer + d, one rounding (scan_matcher_ndt.cpp:123-124), counted against the thresholds
      for (uint32_t k = 0; k < n_lin; ++k) {
        const double d = dlin_s[k];
        const double xa = __dadd_rn(o.x, d), ya = __dadd_rn(o.y, d);
        uint32_t ccx = 0, ccy = 0;
#pragma unroll
        for (uint32_t j = 0; j < K - 1u; ++j) {
          ccx += (xa >= tx[j]) ? 1u : 0u;
          ccy += (ya >= ty[j]) ? 1u : 0u;
        }
        rec_i[kOffKx + k] = static_cast<uint8_t>(ccx);
        rec_i[off_ky + k] = static_cast<uint8_t>(ccy * K);
      }
      *reinterpret_cast<uint32_t *>(rec_i + kOffBase) = by * pitch + bx;
#pragma unroll
      for (uint32_t cy = 0; cy < K; ++cy) {
#pragma unroll
        for (uint32_t cx = 0; cx < K; ++cx) {
          int32_t r = -1;
          if (bx + cx <= size_x + 1u && by + cy <= size_y + 1u) {
            const uint32_t pidx = (by + cy) * pitch + bx + cx;
            const uint2 w = __ldg(occ + (pidx >> 5));
            const uint32_t bit = pidx & 31u;
            if ((w.x >> bit) & 1u) {
              r = static_cast<int32_t>(w.y + __popc(w.x & ((1u << bit) - 1u)));
            }
          }
          reinterpret_cast<int32_t *>(rec_i + kOffRank)[cy * K + cx] = r;
        }
      }
    }
    __syncthreads();
    // ---- phase B: this thread's candidate against its group's share of the pass, in point order.
    // Two instantiations of the loop: models without stiff cells (the usual case; the build counts
    // them, ModelView::n_stiff) run it without the stiff-flag bookkeeping and branch.
    if (active) {
      const uint32_t per_group = (np + n_groups - 1u) / n_groups;
      const uint32_t i_end = min(np, (group + 1u) * per_group);
      const uint32_t i_begin = min(np, group * per_group);
      const uint32_t my_kx = kOffKx + ix, my_ky = off_ky + iy;
      auto phase_b = [&](auto stiff_tag) {
        constexpr bool STIFF = decltype(stiff_tag)::value;
        int32_t r_cached = -1;
        bool stiff_cached = false;
        double2 mean = make_double2(0.0, 0.0), AB = mean, Ds = mean;
        // one pointer walks the point records; this candidate's two bytes sit at fixed offsets
        const unsigned char * prec = win_smem + static_cast<size_t>(i_begin) * S;
        for (uint32_t i0 = i_begin; i0 < i_end; i0 += kWinBlockPts) {
          const uint32_t n_blk = min(i_end - i0, kWinBlockPts);
          float blk = 0.0f;
          for (uint32_t j = 0; j < n_blk; ++j, prec += S) {
            // one (candidate, point) evaluation
            const uint32_t k = static_cast<uint32_t>(prec[my_kx]) + static_cast<uint32_t>(prec[my_ky]);
            const int32_t r = reinterpret_cast<const int32_t *>(prec + kOffRank)[k];
            // consecutive beams mostly stay in one cell: the record is fetched only when this
            // candidate's cell changes; an unoccupied cell evaluates the record at hand and drops
            // the result -- cheaper than diverging around the arithmetic
            if (r >= 0 && r != r_cached) {
              const double2 * f2 = reinterpret_cast<const double2 *>(
                rec_fast + static_cast<size_t>(r) * NDT2D_REC_DOUBLES);
              mean = __ldg(f2);
              AB = __ldg(f2 + 1);
              Ds = __ldg(f2 + 2);
              r_cached = r;
              if (STIFF) {
                stiff_cached = ((__double2hiint(Ds.y) & 0x7fffffff) | __double2loint(Ds.y)) != 0;
              }
            }
            const double2 o = *reinterpret_cast<const double2 *>(prec);
            const double x = __dadd_rn(o.x, dx), y = __dadd_rn(o.y, dy);   // scan_matcher_ndt.cpp:123-124
            const double qx = x - mean.x, qy = y - mean.y;
            const double e = qx * (AB.x * qx + AB.y * qy) + (Ds.x * qy) * qy;   // log2 of the likelihood
            float f;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(static_cast<float>(e)));
            if (STIFF && r >= 0 && stiff_cached) {
              // stiff cell: the reference's own grouping (see search_common.cuh)
              f = 0.0f;
              const uint32_t b0 = *reinterpret_cast<const uint32_t *>(prec + kOffBase);
              acc += cell_likelihood(occ, mv.rec, b0 + (k / K) * pitch + (k % K), x, y);
            }
            blk += (r >= 0) ? f : 0.0f;
          }
          acc += static_cast<double>(blk);
        }
      };
      if (any_stiff) {
        phase_b(std::true_type{});
      } else {
        phase_b(std::false_type{});
      }
    }
  }

  // the groups' partial sums of a candidate, added in group (= point) order
  group_sums[threadIdx.x] = acc;
  __syncthreads();
  Best best{0.0, kNoIndex};
  double sum[6] = {0, 0, 0, 0, 0, 0};
  if (active && group == 0u) {
    for (uint32_t g = 1; g < n_groups; ++g) {acc += group_sums[g * tile + lc];}
    const double score = -acc;
    const uint64_t gi = static_cast<uint64_t>(itheta) * n_cand + c;
    if (scores) {scores[gi] = score;}
    best_merge(best, score, static_cast<double>(gi));
    sum[0] = score;
    sum[1] = dx * score;
    sum[2] = dy * score;
    sum[3] = (dx * dx) * score;
    sum[4] = (dx * dy) * score;
    sum[5] = (dy * dy) * score;
  }
  double * out = block_partials +
    (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * NDT2D_BLOCK_PARTIAL;
  block_reduce_partial_dyn(best, sum, out);
  if (threadIdx.x == 0) {out[8] = sv.dth[itheta];}
}

template<uint32_t K>
__global__ void __launch_bounds__(kWinMaxThreads) search_window_kernel(
  ModelView mv, SearchView sv, uint32_t theta_begin, uint32_t tile, uint32_t n_groups,
  double * __restrict__ block_partials, double * __restrict__ scores)
{
  // the finish kernel of this search may be scheduled now (it waits for this grid's completion
  // with griddepcontrol.wait): its launch latency overlaps the search
  asm volatile("griddepcontrol.launch_dependents;");
  window_block<K>(mv, sv, theta_begin, tile, n_groups, block_partials, scores);
}

// Several searches in one launch (match_scan_batch): blockIdx.z picks the search, its
// descriptor is copied to shared memory first.
template<uint32_t K>
__global__ void __launch_bounds__(kWinMaxThreads) search_window_batch_kernel(
  const BatchEntry * __restrict__ batch, uint32_t tile, uint32_t n_groups)
{
  __shared__ __align__(16) uint32_t entry_words[(sizeof(BatchEntry) + 3) / 4];
  const uint32_t * src = reinterpret_cast<const uint32_t *>(batch + blockIdx.z);
  for (uint32_t k = threadIdx.x; k < sizeof(BatchEntry) / 4; k += blockDim.x) {entry_words[k] = src[k];}
  __syncthreads();
  const BatchEntry & e = *reinterpret_cast<const BatchEntry *>(entry_words);
  window_block<K>(e.mv, e.sv, 0u, tile, n_groups, e.job_partials, nullptr);
}

// Tile (candidates per CTA), point groups and block size for a lattice of n_lin x n_lin.
struct WinShape
{
  uint32_t tile, tiles, groups, threads;
};
WinShape window_shape(uint32_t n_lin, uint32_t max_threads)
{
  const uint64_t n_cand = static_cast<uint64_t>(n_lin) * n_lin;
  WinShape w;
  w.tile = n_cand < kWinTile ? static_cast<uint32_t>(n_cand ? n_cand : 1u) : kWinTile;
  w.tiles = static_cast<uint32_t>((n_cand + w.tile - 1) / w.tile);
  w.groups = max_threads / w.tile;
  w.threads = (w.tile * w.groups + 31u) & ~31u;
  return w;
}

size_t window_smem(uint32_t K, uint32_t n_lin)
{
  return static_cast<size_t>(win_pass_points(K, n_lin)) * win_point_bytes(K, n_lin);
}

}  // namespace

// Cells per axis a window of n_lin steps can touch (0 = too wide for this kernel).  The
// thresholds sit within a few ulps of origin + k * cell, the 1e-9 margin covers that and the
// rounding of the accumulated loop values.
uint32_t ndt2d_window_cells(double cell_size, uint32_t n_lin, double linear_res)
{
  if (!(cell_size > 0.0) || !(linear_res > 0.0) || n_lin == 0 || n_lin > kWinMaxLin) {return 0u;}
  const double span = static_cast<double>(n_lin - 1u) * linear_res / cell_size * (1.0 + 1e-9);
  if (!(span < static_cast<double>(kWinMaxK))) {return 0u;}
  const uint32_t K = static_cast<uint32_t>(span) + 2u;
  return K <= kWinMaxK ? K : 0u;
}

uint32_t ndt2d_window_records(uint32_t n_theta, uint32_t n_lin)
{
  const uint64_t r = static_cast<uint64_t>(n_theta) * window_shape(n_lin, kWinMaxThreads).tiles;
  return r > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(r);
}

int ndt2d_launch_search_window(
  const ModelView & mv, const SearchView & sv, uint32_t K, uint32_t theta_begin, uint32_t n_theta,
  double * d_block_partials, double * d_scores, cudaStream_t stream, Counters * ctr)
{
  if (K < 2 || K > kWinMaxK) {return NDT2D_ERR_INVALID;}
  const WinShape ws = window_shape(sv.n_lin, kWinMaxThreads);   // one search: shortest serial loop
  const uint32_t bx = ws.tiles;
  const uint32_t stride = sv.theta_stride ? sv.theta_stride : 1u;
  uint32_t done = 0;
  while (done < n_theta) {
    const uint32_t ny = min(n_theta - done, 65535u);
    dim3 grid(bx, ny);
    double * out = d_block_partials + static_cast<size_t>(done) * bx * NDT2D_BLOCK_PARTIAL;
    const uint32_t tb = theta_begin + done * stride;
    if (K == 2) {
      search_window_kernel<2><<<grid, ws.threads, window_smem(2, sv.n_lin), stream>>>(mv, sv, tb, ws.tile, ws.groups, out,
        d_scores);
    } else if (K == 3) {
      search_window_kernel<3><<<grid, ws.threads, window_smem(3, sv.n_lin), stream>>>(mv, sv, tb, ws.tile, ws.groups, out,
        d_scores);
    } else {
      search_window_kernel<4><<<grid, ws.threads, window_smem(4, sv.n_lin), stream>>>(mv, sv, tb, ws.tile, ws.groups, out,
        d_scores);
    }
    NDT2D_LAUNCH_CHECK(ctr);
    done += ny;
  }
  return NDT2D_OK;
}

int ndt2d_launch_search_window_batch(
  const BatchEntry * d_batch, uint32_t n_batch, uint32_t K, uint32_t n_ang, uint32_t n_lin,
  cudaStream_t stream, Counters * ctr)
{
  if (n_batch == 0 || n_ang == 0 || n_lin == 0) {return NDT2D_OK;}
  if (K < 2 || K > kWinMaxK) {return NDT2D_ERR_INVALID;}
  if (n_ang > 65535u || n_batch > 65535u) {return NDT2D_ERR_SIZE;}
  // a batch has CTAs to spare: smaller ones overlap the two phases of different CTAs better
  const WinShape ws = window_shape(n_lin, kWinBatchThreads);
  dim3 grid(ws.tiles, n_ang, n_batch);
  if (K == 2) {
    search_window_batch_kernel<2><<<grid, ws.threads, window_smem(2, n_lin), stream>>>(d_batch, ws.tile,
      ws.groups);
  } else if (K == 3) {
    search_window_batch_kernel<3><<<grid, ws.threads, window_smem(3, n_lin), stream>>>(d_batch, ws.tile,
      ws.groups);
  } else {
    search_window_batch_kernel<4><<<grid, ws.threads, window_smem(4, n_lin), stream>>>(d_batch, ws.tile,
      ws.groups);
  }
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}
