// search.cu -- candidate scoring on the device (kernels K4, K5).
//
// Replaces ScanMatcherNDT::matchScan's three nested loops
// (scan_matcher_ndt.cpp:103-143) together with NDT::likelihood(vector<Point>)
// / NDT::likelihood(Vector2d) / NDT::getIndex / Cell::score
// (ndt_model.cpp:105-116, 162-187, 203-218), and ScanMatcherNDT::scorePoints
// (scan_matcher_ndt.cpp:156-178) for batches of poses.
//
// Exactness contract:
//  * A point's position is formed with the reference's own operation sequence
//    in IEEE double, using the never-contracted *_rn intrinsics:
//       outer = (px*c - py*s) + pose     (scan_matcher_ndt.cpp:111-114)
//       inner = outer + d                (:123-124)
//    and cos/sin come from the HOST libm (staged per theta), so the cell a
//    point falls in is the reference's cell, bit for bit.
//  * The Gaussian exponent is evaluated in double from double means; only the
//    final 2^t uses the FP32 special-function unit (relative error ~2e-7,
//    inside the 1e-5 contract).  Per-candidate sums are double.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ndt2d_internal.h"
#include "search_common.cuh"

namespace
{

constexpr int kPlainThreads = 256;
constexpr int kPlainChunk = 2048;     // outer points staged per pass (32 KB)
constexpr int kPlainMaxBlocksX = 32;  // candidate chunks per theta slice

using namespace ndt2d_dev;

// ------------------------------------------------------------------ K4, plain
// One thread per candidate, one theta slice per blockIdx.y, reference
// arithmetic for every cell lookup (two double divides per evaluation).  Kept
// as the simple on-device cross-check of the production kernels.
__global__ void __launch_bounds__(kPlainThreads) search_plain_kernel(
  ModelView mv, SearchView sv, uint32_t theta_begin, double * __restrict__ block_partials,
  double * __restrict__ scores)
{
  __shared__ double2 outer[kPlainChunk];
  const uint32_t itheta = theta_begin + blockIdx.y * sv.theta_stride;
  const double2 cs = sv.trig[itheta];
  const double dth = sv.dth[itheta];
  const uint32_t n_lin = sv.n_lin;
  const uint32_t n_cand = n_lin * n_lin;

  Best best{0.0, kNoIndex};
  double sum[6] = {0, 0, 0, 0, 0, 0};

  for (uint32_t base = blockIdx.x * kPlainThreads; base < n_cand;
    base += gridDim.x * kPlainThreads)
  {
    const uint32_t c = base + threadIdx.x;
    const bool active = c < n_cand;
    const uint32_t ix = active ? c / n_lin : 0u;
    const uint32_t iy = active ? c - ix * n_lin : 0u;
    const double dx = sv.dlin[ix], dy = sv.dlin[iy];
    double acc = 0.0;
    for (uint32_t p0 = 0; p0 < sv.n_pts; p0 += kPlainChunk) {
      const uint32_t np = min(static_cast<uint32_t>(kPlainChunk), sv.n_pts - p0);
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < np; i += kPlainThreads) {
        const double2 p = sv.pts[p0 + i];
        double2 o;
        o.x = __dadd_rn(__dsub_rn(__dmul_rn(p.x, cs.x), __dmul_rn(p.y, cs.y)), sv.pose_x);
        o.y = __dadd_rn(__dadd_rn(__dmul_rn(p.x, cs.y), __dmul_rn(p.y, cs.x)), sv.pose_y);
        outer[i] = o;
      }
      __syncthreads();
      if (active) {
        for (uint32_t i = 0; i < np; ++i) {
          const double2 o = outer[i];
          acc += point_likelihood_exact(mv, __dadd_rn(o.x, dx), __dadd_rn(o.y, dy));
        }
      }
    }
    if (active) {
      const double score = -acc;
      const double gidx =
        static_cast<double>(static_cast<uint64_t>(itheta) * n_cand + c);
      if (scores) {scores[static_cast<uint64_t>(itheta) * n_cand + c] = score;}
      best_merge(best, score, gidx);
      sum[0] += score;
      sum[1] += dx * score;
      sum[2] += dy * score;
      sum[3] += (dx * dx) * score;
      sum[4] += (dx * dy) * score;
      sum[5] += (dy * dy) * score;
    }
  }
  double * out = block_partials +
    (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * NDT2D_BLOCK_PARTIAL;
  block_reduce_partial<kPlainThreads>(best, sum, out);
  if (threadIdx.x == 0) {out[8] = dth;}
}

// ------------------------------------------------------------------ K4, dense (small searches)
// Local matching (config 1: 20,000 candidates x 360 points, most points landing in occupied
// cells) is too small and too dense for the region kernel's machinery to pay: here ONE WARP
// scores one candidate, lanes stride over the scan points (rotated once per CTA into shared
// memory), cell coordinates come from the exact thresholds, the likelihood from the packed
// records (short form for well-conditioned cells, the reference's grouping for stiff ones),
// and a shuffle tree adds the 32 partial sums.  A CTA walks kDenseCandPerBlock candidates
// of one theta slice and leaves one 9-double record for the final reduction.
constexpr uint32_t kDenseWarps = 8;
constexpr uint32_t kDenseCandPerBlock = 32;
constexpr uint32_t kDenseMaxPts = 2048;         // outer points staged per CTA (32 KB)

// Likelihood of one map-frame point, thresholds + packed records (see ModelView).
__device__ __forceinline__ double point_likelihood_fast(
  const ModelView & mv, double inv_cell, double x, double y)
{
  const uint32_t ex = padded_coord_thr_g(x, mv.thr_x, mv.g.size_x, mv.g.origin_x, inv_cell);
  const uint32_t ey = padded_coord_thr_g(y, mv.thr_y, mv.g.size_y, mv.g.origin_y, inv_cell);
  const uint32_t pidx = ey * mv.g.pitch + ex;
  const uint2 w = __ldg(mv.occ + (pidx >> 5));
  const uint32_t bit = pidx & 31u;
  if (((w.x >> bit) & 1u) == 0u) {return 0.0;}
  const uint32_t rank = w.y + __popc(w.x & ((1u << bit) - 1u));
  const double2 * f2 = reinterpret_cast<const double2 *>(
    mv.rec_fast + static_cast<size_t>(rank) * NDT2D_REC_DOUBLES);
  const double2 mean = __ldg(f2), AB = __ldg(f2 + 1), Ds = __ldg(f2 + 2);
  const double qx = x - mean.x, qy = y - mean.y;
  if (Ds.y == 0.0) {
    const double e = qx * (AB.x * qx + AB.y * qy) + (Ds.x * qy) * qy;   // log2 of the likelihood
    float f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(static_cast<float>(e)));
    return static_cast<double>(f);
  }
  return cell_likelihood(mv.occ, mv.rec, pidx, x, y);   // stiff cell: reference grouping
}

__device__ __forceinline__ void dense_block(
  const ModelView & mv, const SearchView & sv, uint32_t theta_begin,
  double * __restrict__ block_partials, double * __restrict__ scores)
{
  __shared__ double2 outer[kDenseMaxPts];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t itheta = theta_begin + blockIdx.y * sv.theta_stride;
  const double2 cs = sv.trig[itheta];
  const uint32_t n_lin = sv.n_lin, n_cand = n_lin * n_lin;
  const double inv_cell = 1.0 / mv.g.cell_size;
  __shared__ double cand_acc[kDenseCandPerBlock];
  const uint32_t c_lo = blockIdx.x * kDenseCandPerBlock;
  const uint32_t c_hi = min(n_cand, c_lo + kDenseCandPerBlock);
  if (threadIdx.x < kDenseCandPerBlock) {cand_acc[threadIdx.x] = 0.0;}
  // scans longer than kDenseMaxPts are walked in several passes
  for (uint32_t p0 = 0; p0 < sv.n_pts; p0 += kDenseMaxPts) {
    const uint32_t np = min(kDenseMaxPts, sv.n_pts - p0);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) {
      const double2 p = sv.pts[p0 + i];
      double2 o;
      // outer = (p.x*c - p.y*s) + pose.x , (p.x*s + p.y*c) + pose.y  (scan_matcher_ndt.cpp:111-114)
      o.x = __dadd_rn(__dsub_rn(__dmul_rn(p.x, cs.x), __dmul_rn(p.y, cs.y)), sv.pose_x);
      o.y = __dadd_rn(__dadd_rn(__dmul_rn(p.x, cs.y), __dmul_rn(p.y, cs.x)), sv.pose_y);
      outer[i] = o;
    }
    __syncthreads();
    for (uint32_t c = c_lo + warp; c < c_hi; c += kDenseWarps) {
      const uint32_t ix = c / n_lin, iy = c - ix * n_lin;
      const double dx = sv.dlin[ix], dy = sv.dlin[iy];
      double acc = 0.0;
      for (uint32_t i = lane; i < np; i += 32) {
        const double2 o = outer[i];
        acc += point_likelihood_fast(mv, inv_cell, __dadd_rn(o.x, dx), __dadd_rn(o.y, dy));
      }
      acc = warp_sum(acc);
      if (lane == 0) {cand_acc[c - c_lo] += acc;}
    }
  }
  __syncthreads();
  Best best{0.0, kNoIndex};
  double sum[6] = {0, 0, 0, 0, 0, 0};
  if (threadIdx.x < c_hi - c_lo) {
    const uint32_t c = c_lo + threadIdx.x;
    const uint32_t ix = c / n_lin, iy = c - ix * n_lin;
    const double dx = sv.dlin[ix], dy = sv.dlin[iy];
    const double score = -cand_acc[threadIdx.x];
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
  block_reduce_partial<kDenseWarps * 32>(best, sum, out);
  if (threadIdx.x == 0) {out[8] = sv.dth[itheta];}
}

__global__ void __launch_bounds__(kDenseWarps * 32) search_dense_kernel(
  ModelView mv, SearchView sv, uint32_t theta_begin, double * __restrict__ block_partials,
  double * __restrict__ scores)
{
  dense_block(mv, sv, theta_begin, block_partials, scores);
}

// The same CTA body over SEVERAL searches in one launch (match_scan_batch with coarse
// lattices, where a region of the region kernel holds too few candidates to fill a warp):
// blockIdx.z picks the search, its descriptor is copied to shared memory first.
__global__ void __launch_bounds__(kDenseWarps * 32) search_dense_batch_kernel(
  const BatchEntry * __restrict__ batch)
{
  __shared__ __align__(16) uint32_t entry_words[(sizeof(BatchEntry) + 3) / 4];
  const uint32_t * src = reinterpret_cast<const uint32_t *>(batch + blockIdx.z);
  for (uint32_t k = threadIdx.x; k < sizeof(BatchEntry) / 4; k += blockDim.x) {entry_words[k] = src[k];}
  __syncthreads();
  const BatchEntry & e = *reinterpret_cast<const BatchEntry *>(entry_words);
  dense_block(e.mv, e.sv, 0u, e.job_partials, nullptr);
}

uint32_t dense_blocks_x(uint32_t n_lin)
{
  const uint64_t n_cand = static_cast<uint64_t>(n_lin) * n_lin;
  return static_cast<uint32_t>((n_cand + kDenseCandPerBlock - 1) / kDenseCandPerBlock);
}

// ------------------------------------------------------------------ finish
// rec[0..15] partial record (see ndt2d_b200.h); rec[16..31] finished outputs:
//   [16..18] delta (dx, dy, dth)  [19] delta_written  [20..28] covariance
//   [29] best / n
__device__ void finish_record(double * rec, const double * dth, const double * dlin,
  uint32_t n_lin)
{
  const double best = rec[0];
  const double n = rec[13];
  const bool written = best < 0.0;
  rec[19] = written ? 1.0 : 0.0;
  rec[16] = rec[17] = rec[18] = 0.0;
  if (written && dth && dlin) {
    const uint64_t idx = static_cast<uint64_t>(rec[1]);
    const uint64_t n_cand = static_cast<uint64_t>(n_lin) * n_lin;
    const uint64_t it = idx / n_cand, rem = idx - it * n_cand;
    const uint64_t ix = rem / n_lin, iy = rem - ix * n_lin;
    rec[16] = dlin[ix];
    rec[17] = dlin[iy];
    rec[18] = dth[it];
  }
  // covariance = (1/s) k + ((1/(s*s)) u) u^T      (scan_matcher_ndt.cpp:146)
  const double s = rec[11];
  const double inv_s = 1.0 / s, inv_s2 = 1.0 / (s * s);
  const double k[9] = {rec[2], rec[3], rec[4], rec[3], rec[5], rec[6], rec[4], rec[6], rec[7]};
  const double u[3] = {rec[8], rec[9], rec[10]};
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {
      rec[20 + r * 3 + c] = inv_s * k[r * 3 + c] + (inv_s2 * u[r]) * u[c];
    }
  }
  // :148 (n == 0 -> NaN, as in the reference); `best` stays the +0.0 it was
  // initialised with when no candidate scored below zero (:83,128)
  rec[29] = written ? best / n : 0.0 / n;
  rec[30] = rec[31] = 0.0;
}

// Reduction of the per-job partials of one launch, two deterministic stages:
//   stage 1  search_reduce_kernel  block b folds the jobs of its contiguous chunk into
//            one 12-double record {best, index, kxx kxy kxt kyy kyt ktt ux uy ut s}
//            (the job's dtheta is folded into the theta terms here)
//   stage 2  search_finish_kernel  one block folds the <= kReduceBlocks records and
//            writes the 32-double result record.
constexpr uint32_t kReduceBlocks = 592;
constexpr int kStage1Doubles = 12;

__device__ __forceinline__ void block_fold_12(
  Best best, double (&acc)[10], double * __restrict__ out12)
{
  __shared__ double red[8][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  warp_best(best);
#pragma unroll
  for (int k = 0; k < 10; ++k) {acc[k] = warp_sum(acc[k]);}
  if (lane == 0) {
    red[warp][0] = best.score;
    red[warp][1] = best.index;
#pragma unroll
    for (int k = 0; k < 10; ++k) {red[warp][2 + k] = acc[k];}
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Best t{red[0][0], red[0][1]};
    double s[10];
    for (int k = 0; k < 10; ++k) {s[k] = red[0][2 + k];}
    for (int w = 1; w < 8; ++w) {
      best_merge(t, red[w][0], red[w][1]);
      for (int k = 0; k < 10; ++k) {s[k] += red[w][2 + k];}
    }
    out12[0] = t.score;
    out12[1] = t.index;
    for (int k = 0; k < 10; ++k) {out12[2 + k] = s[k];}
  }
}

__global__ void __launch_bounds__(256) search_reduce_kernel(
  const double * __restrict__ block_partials, uint32_t n_blocks, double * __restrict__ stage1)
{
  const uint32_t chunk = (n_blocks + gridDim.x - 1) / gridDim.x;
  const uint32_t lo = blockIdx.x * chunk, hi = min(n_blocks, lo + chunk);
  Best best{0.0, kNoIndex};
  double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // kxx kxy kxt kyy kyt ktt ux uy ut s
  for (uint32_t b = lo + threadIdx.x; b < hi; b += blockDim.x) {
    const double * p = block_partials + static_cast<size_t>(b) * NDT2D_BLOCK_PARTIAL;
    best_merge(best, p[0], p[1]);
    const double S = p[2], Sx = p[3], Sy = p[4], Sxx = p[5], Sxy = p[6], Syy = p[7], t = p[8];
    acc[0] += Sxx;
    acc[1] += Sxy;
    acc[2] += t * Sx;
    acc[3] += Syy;
    acc[4] += t * Sy;
    acc[5] += (t * t) * S;
    acc[6] += Sx;
    acc[7] += Sy;
    acc[8] += t * S;
    acc[9] += S;
  }
  block_fold_12(best, acc, stage1 + static_cast<size_t>(blockIdx.x) * kStage1Doubles);
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long * p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long * p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Warp 0 of the finish kernel: publish this rank's record to every mailbox, wait for
// everybody's, reduce (see ExchangeView in ndt2d_internal.h).  out32[0..15] holds this
// rank's record on entry and the combined record on exit; out32[31] = 1 flags a timeout.
__device__ void exchange_and_combine(ExchangeView xv, double * out32, const double * dth,
  const double * dlin, uint32_t n_lin)
{
  const uint32_t lane = threadIdx.x;
  const uint32_t parity = static_cast<uint32_t>(xv.seq & 1ull);
  const size_t slot_mine = static_cast<size_t>(parity) * kExchangeMaxRanks + xv.rank;
  if (lane < xv.world) {
    char * base = static_cast<char *>(xv.peers[lane]);
    double * dst = reinterpret_cast<double *>(base) + slot_mine * 16;
#pragma unroll
    for (int k = 0; k < 16; ++k) {dst[k] = out32[k];}
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long *>(base + kExchangeRecordBytes) + slot_mine,
      xv.seq);
  }
  char * mine = static_cast<char *>(xv.peers[xv.rank]);
  const unsigned long long * flags =
    reinterpret_cast<const unsigned long long *>(mine + kExchangeRecordBytes) +
    static_cast<size_t>(parity) * kExchangeMaxRanks;
  bool ok = true;
  if (lane < xv.world) {
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flags + lane) < xv.seq) {
      if (global_timer_ns() - t0 > xv.timeout_ns) {
        ok = false;
        break;
      }
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    const double * recs = reinterpret_cast<const double *>(mine) +
      static_cast<size_t>(parity) * kExchangeMaxRanks * 16;
    Best best{0.0, kNoIndex};
    double s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    double cands = 0.0, npts = 0.0;
    for (uint32_t r = 0; r < xv.world; ++r) {
      // the reducing thread acquires every flag itself, so the record reads below are
      // ordered after the peer's release
      if (ok) {(void)ld_acquire_sys(flags + r);}
      const volatile double * p = recs + static_cast<size_t>(r) * 16;
      best_merge(best, p[0], p[1]);
      for (int k = 0; k < 10; ++k) {s[k] += p[2 + k];}
      cands += p[12];
      npts = fmax(npts, p[13]);
    }
    out32[0] = best.score;
    out32[1] = best.index;
    for (int k = 0; k < 10; ++k) {out32[2 + k] = s[k];}
    out32[12] = cands;
    out32[13] = npts;
    out32[14] = out32[15] = 0.0;
    finish_record(out32, dth, dlin, n_lin);
    if (!ok) {
      out32[29] = nan("");   // a rank never arrived: no valid result
      out32[31] = 1.0;
    }
  }
}

// `direct`: the records are the 9-double job records themselves (few jobs: stage 1 skipped).
// `counters` (may be null): the job counter / statistics block of the production kernel,
// cleared here for the next search after the statistics were moved to their "last" slots.
__global__ void __launch_bounds__(256) search_finish_kernel(
  const double * __restrict__ stage1, uint32_t n_records, int direct, SearchView sv,
  double n_candidates, double * __restrict__ out32, ExchangeView xv,
  unsigned long long * __restrict__ counters, HostMailbox hm)
{
  // launched with programmatic stream serialisation: the block may already be resident while
  // the search kernel still runs; everything below needs that kernel's results
  asm volatile("griddepcontrol.wait;" ::: "memory");
  Best best{0.0, kNoIndex};
  double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (threadIdx.x == 0 && counters) {
    counters[3] = counters[1];
    counters[4] = counters[2];
    counters[5] = counters[0];
    counters[0] = counters[1] = counters[2] = 0ull;
  }
  for (uint32_t b = threadIdx.x; b < n_records; b += blockDim.x) {
    if (direct) {
      const double * p = stage1 + static_cast<size_t>(b) * NDT2D_BLOCK_PARTIAL;
      best_merge(best, p[0], p[1]);
      const double S = p[2], Sx = p[3], Sy = p[4], Sxx = p[5], Sxy = p[6], Syy = p[7], t = p[8];
      acc[0] += Sxx;
      acc[1] += Sxy;
      acc[2] += t * Sx;
      acc[3] += Syy;
      acc[4] += t * Sy;
      acc[5] += (t * t) * S;
      acc[6] += Sx;
      acc[7] += Sy;
      acc[8] += t * S;
      acc[9] += S;
    } else {
      const double * p = stage1 + static_cast<size_t>(b) * kStage1Doubles;
      best_merge(best, p[0], p[1]);
#pragma unroll
      for (int k = 0; k < 10; ++k) {acc[k] += p[2 + k];}
    }
  }
  __shared__ double folded[12];
  block_fold_12(best, acc, folded);
  if (threadIdx.x == 0) {
    for (int k = 0; k < 12; ++k) {out32[k] = folded[k];}
    out32[12] = n_candidates;
    out32[13] = static_cast<double>(sv.n_pts);
    out32[14] = out32[15] = 0.0;
    finish_record(out32, sv.dth, sv.dlin, sv.n_lin);
  }
  if (xv.world > 1) {
    __threadfence();
    __syncthreads();   // this rank's record is complete and visible to warp 0
    if (threadIdx.x < 32) {exchange_and_combine(xv, out32, sv.dth, sv.dlin, sv.n_lin);}
  }
  if (hm.out32) {
    // the result record goes straight into the caller's mapped pinned memory (HostMailbox);
    // warp 0 wrote / combined out32 above
    if (threadIdx.x < 32) {
      __syncwarp();
      hm.out32[threadIdx.x] = out32[threadIdx.x];
      __threadfence_system();
      __syncwarp();
      if (threadIdx.x == 0) {st_release_sys(hm.flag, hm.seq);}
    }
  }
}

// Batched finish: block b folds entry b's job records directly and writes results32 + 32 b.
__global__ void __launch_bounds__(256) search_finish_batch_kernel(
  const BatchEntry * __restrict__ batch, uint32_t n_jobs, double n_candidates,
  double * __restrict__ results32, unsigned long long * __restrict__ counters)
{
  const BatchEntry & e = batch[blockIdx.x];
  Best best{0.0, kNoIndex};
  double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (blockIdx.x == 0 && threadIdx.x == 0 && counters) {
    counters[3] = counters[1];
    counters[4] = counters[2];
    counters[5] = counters[0];
    counters[0] = counters[1] = counters[2] = 0ull;
  }
  for (uint32_t b = threadIdx.x; b < n_jobs; b += blockDim.x) {
    const double * p = e.job_partials + static_cast<size_t>(b) * NDT2D_BLOCK_PARTIAL;
    best_merge(best, p[0], p[1]);
    const double S = p[2], Sx = p[3], Sy = p[4], Sxx = p[5], Sxy = p[6], Syy = p[7], t = p[8];
    acc[0] += Sxx;
    acc[1] += Sxy;
    acc[2] += t * Sx;
    acc[3] += Syy;
    acc[4] += t * Sy;
    acc[5] += (t * t) * S;
    acc[6] += Sx;
    acc[7] += Sy;
    acc[8] += t * S;
    acc[9] += S;
  }
  __shared__ double folded[12];
  block_fold_12(best, acc, folded);
  if (threadIdx.x == 0) {
    double * out32 = results32 + 32 * static_cast<size_t>(blockIdx.x);
    for (int k = 0; k < 12; ++k) {out32[k] = folded[k];}
    out32[12] = n_candidates;
    out32[13] = static_cast<double>(e.sv.n_pts);
    out32[14] = out32[15] = 0.0;
    finish_record(out32, e.sv.dth, e.sv.dlin, e.sv.n_lin);
  }
}

// block_partials: n_blocks records of NDT2D_BLOCK_PARTIAL doubles, followed by room for
// kReduceBlocks stage-1 records (ndt2d_search_scratch_doubles accounts for it).
constexpr uint32_t kDirectFinishMax = 4096;   // job records one block folds without stage 1

int launch_final(const double * d_block_partials, uint32_t n_blocks, double * d_stage1,
  const SearchView & sv, double n_candidates, double * d_partial32, cudaStream_t stream,
  Counters * ctr, const ExchangeView * exchange, uint32_t * d_counter, const HostMailbox * host)
{
  ExchangeView xv{};
  if (exchange) {xv = *exchange;}
  HostMailbox hm{nullptr, nullptr, 0ull};
  if (host) {hm = *host;}
  unsigned long long * counters = reinterpret_cast<unsigned long long *>(d_counter);
  if (n_blocks <= kDirectFinishMax) {
    // programmatic dependent launch: a search kernel that executes griddepcontrol.launch_dependents
    // (the window kernel does, at its start) lets this block become resident while it runs, so the
    // launch latency of the finish is off the critical path of a local match; the kernel waits
    // for the search's completion + memory flush itself (griddepcontrol.wait)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NDT2D_CUDA_TRY(cudaLaunchKernelEx(&cfg, search_finish_kernel, d_block_partials, n_blocks, 1, sv,
      n_candidates, d_partial32, xv, counters, hm));
    NDT2D_LAUNCH_CHECK(ctr);
    return NDT2D_OK;
  }
  const uint32_t nb = min(kReduceBlocks, max(1u, (n_blocks + 255u) / 256u));
  search_reduce_kernel<<<nb, 256, 0, stream>>>(d_block_partials, n_blocks, d_stage1);
  NDT2D_LAUNCH_CHECK(ctr);
  search_finish_kernel<<<1, 256, 0, stream>>>(d_stage1, nb, 0, sv, n_candidates, d_partial32, xv,
    counters, hm);
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}

// Lexicographic min + sums over n partial records (one per GPU / theta range).
__global__ void combine_kernel(
  const double * __restrict__ partials, uint32_t n, const double * dth, const double * dlin,
  uint32_t n_lin, double * __restrict__ out32)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) {return;}
  Best best{0.0, kNoIndex};
  double s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  double cands = 0.0, npts = 0.0;
  for (uint32_t r = 0; r < n; ++r) {
    const double * p = partials + static_cast<size_t>(r) * NDT2D_PARTIAL_DOUBLES;
    best_merge(best, p[0], p[1]);
    for (int k = 0; k < 10; ++k) {s[k] += p[2 + k];}
    cands += p[12];
    npts = fmax(npts, p[13]);
  }
  out32[0] = best.score;
  out32[1] = best.index;
  for (int k = 0; k < 10; ++k) {out32[2 + k] = s[k];}
  out32[12] = cands;
  out32[13] = npts;
  out32[14] = out32[15] = 0.0;
  finish_record(out32, dth, dlin, n_lin);
}

// An empty theta range still has to produce a neutral record.
__global__ void empty_partial_kernel(SearchView sv, double * __restrict__ out32, HostMailbox hm)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) {return;}
  for (int k = 0; k < 32; ++k) {out32[k] = 0.0;}
  out32[1] = kNoIndex;
  out32[13] = static_cast<double>(sv.n_pts);
  finish_record(out32, sv.dth, sv.dlin, sv.n_lin);
  if (hm.out32) {
    for (int k = 0; k < 32; ++k) {hm.out32[k] = out32[k];}
    __threadfence_system();
    st_release_sys(hm.flag, hm.seq);
  }
}

// ------------------------------------------------------------------ K5
// One warp per pose; lanes stride over the points.
//   x' = (c*px + (-s)*py) + tx ,  y' = (s*px + c*py) + ty
// which is toEigen(pose) * (px, py, 1) (conversions.hpp:64-68; the z column
// of the rotation is exactly zero).
__global__ void __launch_bounds__(256) score_poses_kernel(
  ModelView mv, const double2 * __restrict__ pts, uint32_t n_pts,
  const double4 * __restrict__ pose_tf, uint32_t n_poses, double sign, int normalise,
  double * __restrict__ out, HostMailbox hm)
{
  const uint32_t pose = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pose < n_poses) {
    const double4 tf = pose_tf[pose];  // x, y, cos, sin
    double acc = 0.0;
    for (uint32_t i = lane; i < n_pts; i += 32) {
      const double2 p = pts[i];
      const double x = __dadd_rn(__dadd_rn(__dmul_rn(tf.z, p.x), __dmul_rn(-tf.w, p.y)), tf.x);
      const double y = __dadd_rn(__dadd_rn(__dmul_rn(tf.w, p.x), __dmul_rn(tf.z, p.y)), tf.y);
      acc += point_likelihood_exact(mv, x, y);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      double v = sign * acc;
      if (normalise) {v = v / static_cast<double>(n_pts);}
      out[pose] = v;
      if (hm.out32) {
        hm.out32[pose] = v;   // single-block launch: the scores go straight to the host mailbox
        __threadfence_system();
      }
    }
  }
  if (hm.out32) {
    __syncthreads();
    if (threadIdx.x == 0) {st_release_sys(hm.flag, hm.seq);}
  }
}

// A handful of poses (scoreScan / scorePoints of the node: ONE pose per scan): one 512-thread
// block, a thread per scan point, poses one after the other -- a warp per pose walks 360 beams in
// 12 dependent rounds (24 us), a block does it in one (the call is pure latency).  Results
// go to `out` and, when given, straight to the host mailbox.
constexpr uint32_t kFewPosesThreads = 512;
__global__ void __launch_bounds__(kFewPosesThreads) score_few_poses_kernel(
  ModelView mv, const double2 * __restrict__ pts, uint32_t n_pts,
  const double4 * __restrict__ pose_tf, uint32_t n_poses, double sign, int normalise,
  double * __restrict__ out, HostMailbox hm)
{
  __shared__ double warp_acc[kFewPosesThreads / 32];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t pose = 0; pose < n_poses; ++pose) {
    const double4 tf = pose_tf[pose];  // x, y, cos, sin
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < n_pts; i += kFewPosesThreads) {
      const double2 p = pts[i];
      const double x = __dadd_rn(__dadd_rn(__dmul_rn(tf.z, p.x), __dmul_rn(-tf.w, p.y)), tf.x);
      const double y = __dadd_rn(__dadd_rn(__dmul_rn(tf.w, p.x), __dmul_rn(tf.z, p.y)), tf.y);
      acc += point_likelihood_exact(mv, x, y);
    }
    acc = warp_sum(acc);
    if (lane == 0) {warp_acc[warp] = acc;}
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (uint32_t w = 0; w < kFewPosesThreads / 32; ++w) {t += warp_acc[w];}
      double v = sign * t;
      if (normalise) {v = v / static_cast<double>(n_pts);}
      out[pose] = v;
      if (hm.out32) {hm.out32[pose] = v;}
    }
    __syncthreads();
  }
  if (hm.out32 && threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(hm.flag, hm.seq);
  }
}

uint32_t plain_blocks_x(uint32_t n_lin)
{
  const uint64_t n_cand = static_cast<uint64_t>(n_lin) * n_lin;
  const uint64_t chunks = (n_cand + kPlainThreads - 1) / kPlainThreads;
  return static_cast<uint32_t>(chunks < kPlainMaxBlocksX ? (chunks ? chunks : 1) : kPlainMaxBlocksX);
}

}  // namespace

size_t ndt2d_search_scratch_doubles(uint32_t n_ang, uint32_t n_lin, double cell_size,
  double linear_res)
{
  const size_t plain =
    static_cast<size_t>(n_ang ? n_ang : 1) * plain_blocks_x(n_lin) * NDT2D_BLOCK_PARTIAL;
  const size_t region = ndt2d_region_scratch_doubles(cell_size, n_ang, n_lin, linear_res);
  const size_t dense =
    static_cast<size_t>(n_ang ? n_ang : 1) * dense_blocks_x(n_lin) * NDT2D_BLOCK_PARTIAL;
  size_t a = plain;
  a = a > region ? a : region;
  a = a > dense ? a : dense;
  return a + static_cast<size_t>(kReduceBlocks) * kStage1Doubles;
}

int ndt2d_launch_search(
  const ModelView & mv, const SearchView & sv, uint32_t theta_begin, uint32_t theta_end,
  int variant, double * d_block_partials, double * d_partial32, double * d_scores,
  uint32_t * d_counter, cudaStream_t stream, Counters * ctr, cudaEvent_t ev_begin,
  cudaEvent_t ev_end, const ExchangeView * exchange, const HostMailbox * host)
{
  if (theta_end <= theta_begin || sv.n_lin == 0) {
    if (exchange) {
      // a rank without slices still takes part in the exchange: neutral record through
      // the same finish kernel (zero stage-1 records)
      ExchangeView xv = *exchange;
      search_finish_kernel<<<1, 256, 0, stream>>>(d_block_partials, 0, 1, sv, 0.0, d_partial32, xv,
        nullptr, host ? *host : HostMailbox{nullptr, nullptr, 0ull});
      NDT2D_LAUNCH_CHECK(ctr);
      return NDT2D_OK;
    }
    empty_partial_kernel<<<1, 32, 0, stream>>>(sv, d_partial32,
      host ? *host : HostMailbox{nullptr, nullptr, 0ull});
    NDT2D_LAUNCH_CHECK(ctr);
    return NDT2D_OK;
  }
  // stage-1 records of the final reduction live at the end of the scratch buffer
  const size_t stage1_offset = ndt2d_search_scratch_doubles(sv.n_ang, sv.n_lin, mv.g.cell_size,
      sv.linear_res) - static_cast<size_t>(kReduceBlocks) * kStage1Doubles;
  const uint32_t stride = sv.theta_stride ? sv.theta_stride : 1u;
  const uint32_t n_theta = (theta_end - theta_begin + stride - 1u) / stride;
  const double n_candidates = static_cast<double>(n_theta) * sv.n_lin * sv.n_lin;
  // variant 0 = auto: the dense warp-per-candidate kernel when there is little work altogether
  // (< 2e7 (candidate, point) pairs: local matches, plugin-default windows -- measured 10-20 %
  // faster there, both kernels being latency-bound), else the region kernel (7x faster
  // already at 1/100 of config 4)
  bool dense = variant == 3 || variant == 5;
  if (variant == 0) {
    dense = n_candidates * static_cast<double>(sv.n_pts) < 2.0e7;
  }
  // small windows (a few cells wide -- the usual local match): thread per candidate, the
  // per-point work shared by the CTA (search_window.cu); variant 5 forces it where eligible
  const uint32_t win_k = (variant == 0 || variant == 5) ?
    ndt2d_window_cells(mv.g.cell_size, sv.n_lin, sv.linear_res) : 0u;
  if (dense && win_k != 0u) {
    if (ev_begin) {NDT2D_CUDA_TRY(cudaEventRecord(ev_begin, stream));}
    const int rc = ndt2d_launch_search_window(mv, sv, win_k, theta_begin, n_theta, d_block_partials,
        d_scores, stream, ctr);
    if (rc != NDT2D_OK) {return rc;}
    if (ev_end) {NDT2D_CUDA_TRY(cudaEventRecord(ev_end, stream));}
    return launch_final(d_block_partials, ndt2d_window_records(n_theta, sv.n_lin),
             d_block_partials + stage1_offset, sv, n_candidates, d_partial32, stream, ctr, exchange,
             nullptr, host);
  }
  if (dense) {
    const uint32_t bx = dense_blocks_x(sv.n_lin);
    uint32_t done = 0;
    if (ev_begin) {NDT2D_CUDA_TRY(cudaEventRecord(ev_begin, stream));}
    while (done < n_theta) {
      const uint32_t ny = min(n_theta - done, 65535u);
      dim3 grid(bx, ny);
      search_dense_kernel<<<grid, kDenseWarps * 32, 0, stream>>>(
        mv, sv, theta_begin + done * stride,
        d_block_partials + static_cast<size_t>(done) * bx * NDT2D_BLOCK_PARTIAL, d_scores);
      NDT2D_LAUNCH_CHECK(ctr);
      done += ny;
    }
    if (ev_end) {NDT2D_CUDA_TRY(cudaEventRecord(ev_end, stream));}
    return launch_final(d_block_partials, n_theta * bx, d_block_partials + stage1_offset, sv,
             n_candidates, d_partial32, stream, ctr, exchange, nullptr, host);
  }
  if (variant != 1) {
    uint32_t n_jobs = 0;
    if (ev_begin) {NDT2D_CUDA_TRY(cudaEventRecord(ev_begin, stream));}
    const int rc = ndt2d_launch_search_region(mv, sv, sv.linear_res, theta_begin, n_theta,
        d_block_partials, d_scores, d_counter, sv.coords, sv.coords_cap_bytes, stream, ctr,
        &n_jobs);
    if (rc != NDT2D_OK) {return rc;}
    if (ev_end) {NDT2D_CUDA_TRY(cudaEventRecord(ev_end, stream));}
    return launch_final(d_block_partials, n_jobs, d_block_partials + stage1_offset, sv,
             n_candidates, d_partial32, stream, ctr, exchange, d_counter, host);
  }
  const uint32_t bx = plain_blocks_x(sv.n_lin);
  // gridDim.y is limited to 65535: slice the theta range if needed
  uint32_t done = 0;
  if (ev_begin) {NDT2D_CUDA_TRY(cudaEventRecord(ev_begin, stream));}
  while (done < n_theta) {
    const uint32_t ny = min(n_theta - done, 65535u);
    dim3 grid(bx, ny);
    search_plain_kernel<<<grid, kPlainThreads, 0, stream>>>(
      mv, sv, theta_begin + done * stride,
      d_block_partials + static_cast<size_t>(done) * bx * NDT2D_BLOCK_PARTIAL, d_scores);
    NDT2D_LAUNCH_CHECK(ctr);
    done += ny;
  }
  if (ev_end) {NDT2D_CUDA_TRY(cudaEventRecord(ev_end, stream));}
  return launch_final(d_block_partials, n_theta * bx, d_block_partials + stage1_offset, sv,
           n_candidates, d_partial32, stream, ctr, exchange, nullptr, host);
}

uint32_t ndt2d_dense_batch_records(uint32_t n_ang, uint32_t n_lin)
{
  const uint64_t r = static_cast<uint64_t>(n_ang) * dense_blocks_x(n_lin);
  return r > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(r);
}

int ndt2d_launch_search_dense_batch(
  const BatchEntry * d_batch, uint32_t n_batch, uint32_t n_ang, uint32_t n_lin, cudaStream_t stream,
  Counters * ctr)
{
  if (n_batch == 0 || n_ang == 0 || n_lin == 0) {return NDT2D_OK;}
  if (n_ang > 65535u || n_batch > 65535u) {return NDT2D_ERR_SIZE;}
  dim3 grid(dense_blocks_x(n_lin), n_ang, n_batch);
  search_dense_batch_kernel<<<grid, kDenseWarps * 32, 0, stream>>>(d_batch);
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}

int ndt2d_launch_finish_batch(
  const BatchEntry * d_batch, uint32_t n_batch, uint32_t n_jobs, double n_candidates,
  double * d_results32, uint32_t * d_counter, cudaStream_t stream, Counters * ctr)
{
  if (n_batch == 0) {return NDT2D_OK;}
  search_finish_batch_kernel<<<n_batch, 256, 0, stream>>>(d_batch, n_jobs, n_candidates,
    d_results32, reinterpret_cast<unsigned long long *>(d_counter));
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}

int ndt2d_launch_combine(const double * d_partials, uint32_t n, const double * d_dth,
  const double * d_dlin, uint32_t n_lin, double * d_out32, cudaStream_t stream, Counters * ctr)
{
  combine_kernel<<<1, 32, 0, stream>>>(d_partials, n, d_dth, d_dlin, n_lin, d_out32);
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}

int ndt2d_launch_score_poses(
  const ModelView & mv, const double2 * d_pts, uint32_t n_pts, const double4 * d_pose_tf,
  uint32_t n_poses, double sign, int normalise, double * d_out, cudaStream_t stream,
  Counters * ctr, const HostMailbox * host)
{
  if (n_poses == 0) {return NDT2D_OK;}
  const uint32_t warps_per_block = 256 / 32;
  const uint32_t grid = (n_poses + warps_per_block - 1) / warps_per_block;
  HostMailbox hm{nullptr, nullptr, 0ull};
  if (host && grid == 1) {hm = *host;}   // (the flag protocol needs a single block)
  if (n_poses <= 8) {
    score_few_poses_kernel<<<1, kFewPosesThreads, 0, stream>>>(
      mv, d_pts, n_pts, d_pose_tf, n_poses, sign, normalise, d_out, hm);
    NDT2D_LAUNCH_CHECK(ctr);
    return NDT2D_OK;
  }
  score_poses_kernel<<<grid, 256, 0, stream>>>(
    mv, d_pts, n_pts, d_pose_tf, n_poses, sign, normalise, d_out, hm);
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}
