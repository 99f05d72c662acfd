// search_region.cu -- the production correlative-search kernel (K4).
//
// Replaces the three nested loops of ScanMatcherNDT::matchScan
// (scan_matcher_ndt.cpp:103-143) and, inside them, NDT::likelihood /
// NDT::getIndex / Cell::score (ndt_model.cpp:105-116, 162-187, 203-218).
//
// Work decomposition
//   job    = (theta slice, REGION of RX x RY adjacent (dx, dy) candidates): RX <= 32 columns
//            (one per lane) with (RX - 1) * search_linear_resolution < 2 cells, RY <= 25 rows
//            with (RY - 1) * search_linear_resolution < 1 cell: for a fixed scan point the
//            candidates of a region can only put it into the (<= 3) x 2 cells starting at
//            the cell of the region's first candidate.
//   warp   = one job at a time, taken from a global atomic job counter
//            (persistent CTAs, dynamic balance: job cost varies with how much
//            of the scan overlaps the map).  Warps never synchronise with each
//            other inside the job loop.
//   scan   = 32 scan points per step, one per lane: rotate + translate with the
//            reference's own operation order (scan_matcher_ndt.cpp:111-114),
//            padded cell coordinate of the region's first candidate from the
//            host-tabulated thresholds (bit-exact, no division), one bit test
//            in the DILATED occupancy bitmap (OR of E over that (<= 3) x 2 window).  A
//            clear bit rejects the point for all RX * RY candidates at once -- most of a
//            large search is empty space.
//   item   = a (point, region) pair whose D bit is set, processed by the whole
//            warp, LANE = CANDIDATE COLUMN: the lane forms its exact column coordinate (one
//            __dadd_rn) and compares it with the next two thresholds -> its cell column; a
//            ballot over the exact row coordinates gives the row where the region crosses
//            into the next cell row.  Each lane therefore has (at most) two cells, one per
//            row range, and sets up, in double, the cell's quadratic along its column in
//            VERTEX FORM (rows counted from the row nearest the ridge of the Gaussian):
//            log2 L(b) = c2 b'^2 + d1 b' + e0 with all three terms <= 0 -- no cancellation,
//            so the row loop runs in float: FADD, 2 FFMA, MUFU.EX2, FADD per row, the
//            accumulators being REGISTERS (rows are unrolled, entered by a jump table).
//   sums   = per-candidate float block sums in registers (25 per lane), flushed every 4
//            steps with hits into double totals in shared memory (lane-private columns: no
//            synchronisation anywhere in the job loop).
//
// Staging: D and the threshold tables are bulk-copied into shared memory once
// per CTA with cp.async.bulk (TMA, SASS UBLKCP) completing on an mbarrier when
// they fit; E (+ rank prefix) and the 48-byte cell records are read with
// loads through L1 (at most three distinct addresses per warp).  DRAM traffic is the
// coordinate pre-pass table, written and read once.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <stdint.h>

#include "ndt2d_internal.h"
#include "search_common.cuh"

namespace
{

using namespace ndt2d_dev;

constexpr uint32_t kMaxRX = 32;                 // region columns (candidates along dx) = lanes
constexpr uint32_t kMaxRY = 25;                 // region rows (candidates along dy) = registers
constexpr uint32_t kPairs = (kMaxRY + 1) / 2;   // rows are held and evaluated two at a time
#ifndef NDT2D_REGION_WARPS
#define NDT2D_REGION_WARPS 28
#endif
constexpr uint32_t kWarps = NDT2D_REGION_WARPS;  // warps per CTA; one persistent CTA per SM
// per warp: double totals, [row][lane], then one 32-byte record per lane for the scan points of
// the current step (exact rotated position + packed splits / occupancy / record ranks): an item
// reads its point's record with two broadcast loads instead of seven shuffles, and the
// values do not occupy registers across the item loop
constexpr uint32_t kWarpTotBytes = kMaxRY * 32u * static_cast<uint32_t>(sizeof(double));
constexpr uint32_t kWarpSmemBytes = kWarpTotBytes + 32u * 32u;
// D + thresholds in shared memory up to this: what the warps' totals leave of the 227 KB a CTA
// can opt in to, at most 64 KB
constexpr size_t kSmemLeft = 227 * 1024 - 1024 - static_cast<size_t>(kWarpSmemBytes) * kWarps;
constexpr size_t kSmemTabBudget = kSmemLeft < 64 * 1024 ? kSmemLeft : 64 * 1024;
constexpr uint32_t kBatchSlotBytes = 320;        // per-warp BatchEntry slot of the batch kernel (>= sizeof)
constexpr uint32_t kFlushSteps = 4;             // 32-point steps per float accumulation block
constexpr uint32_t kTargetJobs = 148 * 16;      // shrink regions of small searches
constexpr uint32_t kChunkTargetWork = 148 * kWarps * 3;  // (job, point chunk) pairs wanted in flight
constexpr double kRoundMagic = 6755399441055744.0;      // 1.5 * 2^52: (x + M) - M == rint(x)
constexpr uint32_t kCoordsSlices = 1;           // theta slices per thread of the coordinate pre-pass

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
// 2^x on the special-function unit; results below 2^-126 flush to zero (the
// parity tests carry the matching absolute floor).
__device__ __forceinline__ float ex2_ftz(float x)
{
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct RegionPlan
{
  uint32_t RX, RY;    // region columns / rows
  uint32_t Qx, Qy;    // regions per axis
  uint32_t n_jobs;    // n_theta * Qx * Qy
  uint32_t thr_doubles;  // size_x + 1 + size_y + 1
  uint32_t tab_bytes; // D + thresholds, 16-byte rounded (0 = keep in global memory)
  bool smem_tab;
  uint32_t grid;
  size_t smem_bytes;
  uint32_t P;             // point chunks per job (1 = whole scan in one job)
  uint32_t chunk_points;  // scan points per chunk (multiple of 32)
};

// Padded coordinate pc(v) = #{k in [0, size] : thr[k] <= v}  (0 = below the
// origin, size + 1 = beyond the grid).  The product with 1/cell only provides a
// starting guess; the answer is fixed by the thresholds, which the host derived
// from the reference's own arithmetic (api.cu: axis_thresholds).
template<bool SMEM>
__device__ __forceinline__ uint32_t padded_coord(
  double v, const double * __restrict__ thr, uint32_t size, double origin, double inv_cell)
{
  if (!(v >= thr[0])) {return 0u;}
  const double q = (v - origin) * inv_cell;
  uint32_t pc = (q >= static_cast<double>(size)) ? size : static_cast<uint32_t>(q);
  pc += 1u;
  while (pc <= size && v >= thr[pc]) {++pc;}
  while (pc > 1u && v < thr[pc - 1u]) {--pc;}
  return pc;
}

// Two rows (2J, 2J + 1) of the region for this lane's column, packed FP32 (sm_100 FFMA2 / FADD2):
// b' = row - bs, log2 L = c2 b'^2 + d1 b' + e0.  bp = (b'(2J), b'(2J + 1)) is carried from pair
// to pair (step -2 going down, +2 going up), EE carries e0 per row (-inf masks a row).  The
// chain is software-pipelined by one pair: a block first adds the PREVIOUS pair's two
// likelihoods to their accumulators (JPREV) and then starts its own pair, so the SFU latency
// of a pair overlaps the arithmetic of the next one (every block is a jump target, the
// scheduler cannot move code across them by itself).
#ifdef NDT2D_COUNT_ZERO_PAIRS   // experiment: how many executed pairs are zero on every lane
#define NDT2D_ZERO_TALLY(e) \
  { \
    dbg_pairs += 1u; \
    const bool z = __all_sync(0xffffffffu, e.x < -126.0f && e.y < -126.0f); \
    dbg_zero += z ? 1u : 0u; \
    dbg_phase_nz |= z ? 0u : 1u; \
  }
#else
#define NDT2D_ZERO_TALLY(e)
#endif
#define NDT2D_PAIR_FIRST(EE) \
  { \
    const float2 e = __ffma2_rn(__ffma2_rn(c2p, bp, d1p), bp, EE); \
    NDT2D_ZERO_TALLY(e) \
    prev = make_float2(ex2_ftz(e.x), ex2_ftz(e.y)); \
    bp = __fadd2_rn(bp, stepp); \
  }
#define NDT2D_PAIR_NEXT(JPREV, EE) \
  { \
    acc2[JPREV] = __fadd2_rn(acc2[JPREV], prev); \
    const float2 e = __ffma2_rn(__ffma2_rn(c2p, bp, d1p), bp, EE); \
    NDT2D_ZERO_TALLY(e) \
    prev = make_float2(ex2_ftz(e.x), ex2_ftz(e.y)); \
    bp = __fadd2_rn(bp, stepp); \
  }
#define NDT2D_PAIR_LAST(JPREV) {acc2[JPREV] = __fadd2_rn(acc2[JPREV], prev);}

// Number of k in [0, n) with o + dl[k] < thr, for the replayed lattice dl[k] ~ dl[0] + k h.
// g = (thr - (o + dl[0])) / h counts the steps below the threshold.  The accumulated lattice and
// the roundings of o + dl[k] deviate from the nominal positions dl[0] + k h by at most
// n ulp(|o| + |dl|) / h steps -- below 4e-8 for n <= 16384 positions whose coordinates stay under
// 1e7 steps -- so whenever g is farther than 1e-6 from an integer the count is floor(g) + 1 with
// no further work.  Everything else (near-integer g, huge lattices or coordinates, NaN / inf)
// is decided by the reference's own additions, exact either way.
__device__ __forceinline__ uint32_t count_below(
  double o, const double * __restrict__ dl, uint32_t n, double thr, double inv_h)
{
  const double dl0 = __ldg(dl);
  const double g = (thr - __dadd_rn(o, dl0)) * inv_h;
  const double gf = floor(g);
  const double fr = g - gf;
  if (fr > 1.0e-6 && fr < 1.0 - 1.0e-6 && n <= 16384u && (fabs(o) + fabs(dl0)) * inv_h < 1.0e7) {
    if (g <= 0.0) {return 0u;}
    return g >= static_cast<double>(n) ? n : static_cast<uint32_t>(gf) + 1u;
  }
  uint32_t k = g >= static_cast<double>(n) ? n : (g > 0.0 ? static_cast<uint32_t>(g) + 1u : 0u);
  k = min(k, n);
  while (k > 0u && !(__dadd_rn(o, __ldg(dl + k - 1u)) < thr)) {--k;}
  while (k < n && __dadd_rn(o, __ldg(dl + k)) < thr) {++k;}
  return k;
}

// Occupancy bits of padded cells p, p + 1, p + 2 and the record rank of the first occupied
// one among them (ranks follow the bit order, also across a word boundary).
__device__ __forceinline__ void occ3(
  const uint2 * __restrict__ occ, uint32_t p, uint32_t & bits3, uint32_t & rank_base)
{
  const uint32_t w = p >> 5, sh = p & 31u;
  const uint2 a = __ldg(occ + w);
  const uint32_t hi = sh > 29u ? __ldg(occ + w + 1u).x : 0u;   // (the buffer has 4 words of slack)
  bits3 = __funnelshift_r(a.x, hi, sh) & 7u;
  rank_base = a.y + __popc(a.x & ((1u << sh) - 1u));
}

// Float block sums -> double totals (lane-private column of shared memory).
__device__ __forceinline__ void flush_pairs(
  float2 (&acc2)[kPairs], double * __restrict__ tot, uint32_t RY, uint32_t lane)
{
#pragma unroll
  for (uint32_t j = 0; j < kPairs; ++j) {
    if (2u * j < RY) {
      tot[(2u * j) * 32u + lane] += static_cast<double>(acc2[j].x);
      if (2u * j + 1u < RY) {tot[(2u * j + 1u) * 32u + lane] += static_cast<double>(acc2[j].y);}
      acc2[j] = make_float2(0.0f, 0.0f);
    }
  }
}

// A lane's cell along its candidate column, vertex form, float coefficients.
struct VtxLine
{
  float c2, d1, e0, nbs;
  bool stiff;
};

// Per-lane setup in double.  Record (ndt2d_internal.h): mean, D, Bh, S, {float c2, float stiff}.
//   w0 = (y0 - mean.y) + Bh (x - mean.x)         the quadratic's argument at row 0
//   bs = rint(-w0 / h), dl = w0 + bs h           the row nearest the ridge, |dl| <= h / 2
//   log2 L(b) = c2 b'^2 + (2 h D dl) b' + (D dl^2 + S qx^2),  b' = b - bs
// A lane without an occupied cell gets e0 = -inf (2^-inf == +0).
__device__ __forceinline__ VtxLine vtx_setup(
  bool occ, const double * __restrict__ rec_vtx, uint32_t rank, double xa, double y0, double h,
  double inv_h)
{
  VtxLine ln{0.0f, 0.0f, -INFINITY, 0.0f, false};
  if (occ) {
    const double2 * r = reinterpret_cast<const double2 *>(
      rec_vtx + static_cast<size_t>(rank) * NDT2D_REC_DOUBLES);
    const double2 mean = __ldg(r), DB = __ldg(r + 1), SF = __ldg(r + 2);
    ln.stiff = __double2hiint(SF.y) != 0;
    const double qu = xa - mean.x;
    const double w0 = fma(DB.y, qu, y0 - mean.y);
    if (fabs(w0) < 2097152.0 * h) {   // else the ridge is > 2e6 rows away: L == 0 on this column
      const double rm = fma(-w0, inv_h, kRoundMagic);
      const double bs = __dadd_rn(rm, -kRoundMagic);      // a row within 1/2 (+ 1 ulp) of -w0 / h
      const double dl = fma(bs, h, w0);
      const double Dd = DB.x * dl;
      ln.e0 = static_cast<float>(fma(Dd, dl, (SF.x * qu) * qu));
      ln.d1 = static_cast<float>((h + h) * Dd);
      ln.c2 = __int_as_float(__double2loint(SF.y));
      ln.nbs = __int_as_float(0x4B400000 - __double2loint(rm)) - 12582912.0f;   // float(-bs)
    }
  }
  return ln;
}

// Rows [lo, hi) of a phase in which some lane's cell is stiff (a cluster of near-identical
// points: |I| ~ 1e17, inf / NaN, or D >= 0): every lane takes the reference's own grouping
// ((q^T I) q, ndt_model.cpp:113-114) without FMA, so that it cancels exactly where the
// reference cancels.
__device__ __forceinline__ void stiff_rows(
  float2 (&acc2)[kPairs], bool occ, const double * __restrict__ rec, uint32_t rank, double xa,
  double poy, const double * __restrict__ dlin_rows, uint32_t lo, uint32_t hi)
{
  double2 mean = make_double2(0.0, 0.0), i0010 = mean, i0111 = mean;
  if (occ) {
    const double2 * r2 = reinterpret_cast<const double2 *>(
      rec + static_cast<size_t>(rank) * NDT2D_REC_DOUBLES);
    mean = __ldg(r2);
    i0010 = __ldg(r2 + 1);
    i0111 = __ldg(r2 + 2);
  }
  const double qx = __dsub_rn(xa, mean.x);
#pragma unroll
  for (uint32_t b = 0; b < kMaxRY; ++b) {
    if (b >= lo && b < hi) {
      const double yb = __dadd_rn(poy, dlin_rows[b]);
      const double qy = __dsub_rn(yb, mean.y);
      const double r0 = __dadd_rn(__dmul_rn(qx, i0010.x), __dmul_rn(qy, i0010.y));
      const double r1 = __dadd_rn(__dmul_rn(qx, i0111.x), __dmul_rn(qy, i0111.y));
      const double e = __dadd_rn(__dmul_rn(r0, qx), __dmul_rn(r1, qy));
      const float f = exp2f(static_cast<float>(e * kLog2e));
      if (b & 1u) {acc2[b >> 1].y += occ ? f : 0.0f;} else {acc2[b >> 1].x += occ ? f : 0.0f;}
    }
  }
}

// Pre-pass of the search: everything an item needs that depends on (theta slice, scan point,
// region column) or (theta slice, scan point, region row) only -- the Qx * Qy regions of a slice
// share it, so computing it here instead of inside every job removes a factor ~Q of work:
//   x entry (region column q):  pcx | kx1 << 16 | kx2 << 22
//       pcx  padded cell coordinate of the region's FIRST column,
//       kx1 / kx2  columns of the region left of the next / second-next x threshold
//   y entry (region row q):     pcy | ky << 16      (same for rows; a region spans <= 2 cell rows)
// all by the reference's own additions against the tabulated thresholds (exact).
// Layout: coords[(it * (Qx + Qy) + q) * n_pts_pad + i], u32; q < Qx: x, q >= Qx: y of row q - Qx.
__global__ void __launch_bounds__(128) region_coords_kernel(
  ModelView mv, SearchView sv, uint32_t theta_begin, uint32_t n_theta, uint32_t RX, uint32_t RY,
  uint32_t Qx, uint32_t Qy, uint32_t n_pts_pad, uint32_t * __restrict__ coords)
{
  // one thread per (scan point, axis, theta slice): it walks the axis' Q regions in order.  All
  // entries of the row follow from K(c) = #{k < n_lin : o + dlin[k] < thr[c]} for the ~n_lin h /
  // cell thresholds c the lattice crosses (each found exactly, with the reference's additions):
  // lattice position j lies in padded cell #{c : K(c) <= j}, and a region starting at j0 in
  // cell c has min(n, K(c) - j0) positions left of the next threshold -- about half the
  // searches of doing every region on its own.
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sv.n_pts) {return;}
  const bool is_x = blockIdx.y == 0u;
  const double2 p = sv.pts[i];
  const double inv_cell = 1.0 / mv.g.cell_size, inv_h = sv.inv_linear_res;
  const uint32_t n_lin = sv.n_lin;
  const uint32_t R = is_x ? RX : RY, Q = is_x ? Qx : Qy, q_base = is_x ? 0u : Qx;
  const double * __restrict__ thr = is_x ? mv.thr_x : mv.thr_y;
  const uint32_t size = is_x ? mv.g.size_x : mv.g.size_y;
  const double origin = is_x ? mv.g.origin_x : mv.g.origin_y;
  const double d0 = sv.dlin[0];
  const uint32_t it_end = min(n_theta, (blockIdx.z + 1u) * kCoordsSlices);
  for (uint32_t it = blockIdx.z * kCoordsSlices; it < it_end; ++it) {
    const double2 cs = sv.trig[theta_begin + it * sv.theta_stride];
    // outer = (p.x*c - p.y*s) + pose.x , (p.x*s + p.y*c) + pose.y   (scan_matcher_ndt.cpp:111-114)
    const double o = is_x ?
      __dadd_rn(__dsub_rn(__dmul_rn(p.x, cs.x), __dmul_rn(p.y, cs.y)), sv.pose_x) :
      __dadd_rn(__dadd_rn(__dmul_rn(p.x, cs.y), __dmul_rn(p.y, cs.x)), sv.pose_y);
    uint32_t c = padded_coord<false>(__dadd_rn(o, d0), thr, size, origin, inv_cell);   // cell of position 0
    uint32_t Kc = count_below(o, sv.dlin, n_lin, thr[c], inv_h);                        // > 0
    uint32_t Kn = Kc >= n_lin ? n_lin : count_below(o, sv.dlin, n_lin, thr[min(c + 1u, size + 1u)], inv_h);
    uint32_t * out = coords + (static_cast<size_t>(it) * (Qx + Qy) + q_base) * n_pts_pad + i;
    for (uint32_t q = 0; q < Q; ++q) {
      const uint32_t j0 = q * R, n = min(R, n_lin - j0);
      while (Kc <= j0) {   // the region starts beyond threshold c: next cell (thr[size + 1] = +inf ends it)
        ++c;
        Kc = Kn;
        Kn = Kc >= n_lin ? n_lin : count_below(o, sv.dlin, n_lin, thr[min(c + 1u, size + 1u)], inv_h);
      }
      uint32_t e = c | (min(n, Kc - j0) << 16);
      if (is_x) {e |= min(n, Kn - j0) << 22;}
      out[static_cast<size_t>(q) * n_pts_pad] = e;
    }
  }
}

// Per-candidate score of one job from its sums, warp argmin + the six covariance
// sums -> the job's 9-double record.  sums(b) = this lane's column (dx index jx0 + lane), row b.
template<typename SUMS>
__device__ __forceinline__ void job_epilogue(
  SUMS sums, const SearchView & sv, uint32_t job, uint32_t itheta,
  uint32_t jx0, uint32_t jy0, uint32_t nxc, uint32_t nyc, uint32_t lane,
  double * __restrict__ job_partials, double * __restrict__ scores)
{
  const uint32_t n_lin = sv.n_lin;
  const uint64_t n_cand = static_cast<uint64_t>(n_lin) * n_lin;
  Best best{0.0, kNoIndex};
  double sum[6] = {0, 0, 0, 0, 0, 0};
  if (lane < nxc) {
    // this lane's column: sums over its rows of score * {1, dy, dy^2}, the dx factors once
    const double dx = sv.dlin[jx0 + lane];
    const uint64_t g0 = static_cast<uint64_t>(itheta) * n_cand +
      static_cast<uint64_t>(jx0 + lane) * n_lin + jy0;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (uint32_t b = 0; b < nyc; ++b) {
      const double score = -sums(b);
      const double dy = sv.dlin[jy0 + b];
      const uint64_t gi = g0 + b;
      if (scores) {scores[gi] = score;}
      best_merge(best, score, static_cast<double>(gi));
      s0 += score;
      s1 += dy * score;
      s2 += (dy * dy) * score;
    }
    sum[0] = s0;
    sum[1] = dx * s0;
    sum[2] = s1;
    sum[3] = (dx * dx) * s0;
    sum[4] = dx * s1;
    sum[5] = s2;
  }
  warp_best(best);
#pragma unroll
  for (int k = 0; k < 6; ++k) {sum[k] = warp_sum(sum[k]);}
  if (lane == 0) {
    double * out = job_partials + static_cast<size_t>(job) * NDT2D_BLOCK_PARTIAL;
    out[0] = best.score;
    out[1] = best.index;
#pragma unroll
    for (int k = 0; k < 6; ++k) {out[2 + k] = sum[k];}
    out[8] = sv.dth[itheta];
  }
}

// Small searches split every job's scan points into P chunks (more warps in flight,
// shorter critical path); this kernel adds the chunk sums of a job in chunk order
// (deterministic) and finishes it.  One warp per job; chunk sums are [row][lane].
__global__ void __launch_bounds__(256) region_chunk_reduce_kernel(
  SearchView sv, uint32_t theta_begin, uint32_t RX, uint32_t RY, uint32_t Qx, uint32_t Qy,
  uint32_t n_jobs, uint32_t P, double * __restrict__ chunk_sums,
  double * __restrict__ job_partials, double * __restrict__ scores)
{
  const uint32_t job = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (job >= n_jobs) {return;}
  const uint32_t QQ = Qx * Qy, RR = 32u * RY;
  const uint32_t it = job / QQ, rr = job - it * QQ;
  const uint32_t rx = rr / Qy, ry = rr - rx * Qy;
  const uint32_t jx0 = rx * RX, jy0 = ry * RY;
  const uint32_t nxc = min(RX, sv.n_lin - jx0), nyc = min(RY, sv.n_lin - jy0);
  double * first = chunk_sums + static_cast<size_t>(job) * P * RR;
  for (uint32_t k = lane; k < RR; k += 32) {
    double t = first[k];
    for (uint32_t c = 1; c < P; ++c) {t += first[static_cast<size_t>(c) * RR + k];}
    first[k] = t;
  }
  __syncwarp();
  job_epilogue([first, lane](uint32_t b) {return first[b * 32u + lane];}, sv, job,
    theta_begin + it * sv.theta_stride, jx0, jy0, nxc, nyc, lane, job_partials, scores);
}

// STATS: tally the useful evaluations and items of the launch (ndt2d_matcher_set_tallies; the
// bookkeeping costs ~2 % of the kernel, so it is off unless asked for).
template<bool SMEM_TAB, bool PRE, bool STATS>
__global__ void __launch_bounds__(kWarps * 32, 1)
search_region_kernel(
  ModelView mv, SearchView sv, uint32_t theta_begin, uint32_t RX, uint32_t RY, uint32_t Qx,
  uint32_t Qy, uint32_t n_jobs, uint32_t tab_d_bytes, uint32_t tab_thr_bytes,
  double * __restrict__ job_partials, double * __restrict__ scores,
  uint32_t * __restrict__ job_counter, unsigned long long * __restrict__ stats,
  const uint32_t * __restrict__ coords, uint32_t n_pts_pad, uint32_t P, uint32_t chunk_points,
  double * __restrict__ chunk_sums)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t * mbar = reinterpret_cast<uint64_t *>(smem_raw);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;

  // ---- shared memory: [mbar 16][D][thr_x thr_y][per-warp double totals]
  const uint32_t * occd_tab;
  const double * thr_x_tab;
  const double * thr_y_tab;
  unsigned char * sp = smem_raw + 16;
  if (SMEM_TAB) {
    occd_tab = reinterpret_cast<const uint32_t *>(sp);
    sp += tab_d_bytes;
    thr_x_tab = reinterpret_cast<const double *>(sp);
    thr_y_tab = thr_x_tab + (mv.g.size_x + 2);
    sp += tab_thr_bytes;
    if (threadIdx.x == 0) {
      mbar_init(mbar, 1);
      mbar_expect_tx(mbar, tab_d_bytes + tab_thr_bytes);
      bulk_copy_g2s(const_cast<uint32_t *>(occd_tab), mv.occ_dilated, tab_d_bytes, mbar);
      bulk_copy_g2s(const_cast<double *>(thr_x_tab), mv.thr_x, tab_thr_bytes, mbar);
    }
  } else {
    occd_tab = mv.occ_dilated;
    thr_x_tab = mv.thr_x;
    thr_y_tab = mv.thr_y;
  }
  double * tot = reinterpret_cast<double *>(sp + static_cast<size_t>(warp) * kWarpSmemBytes);

  if (SMEM_TAB) {
    __syncthreads();       // mbarrier initialised before anyone polls it
    mbar_wait(mbar, 0);
  }

#define MV mv
#define SV sv
#define NDT2D_BODY_BATCH 0
#define NDT2D_BODY_TOTAL_WORK (n_jobs * P)
#define NDT2D_BODY_OCCD occd_tab
#define NDT2D_BODY_THRX thr_x_tab
#define NDT2D_BODY_THRY thr_y_tab
#define NDT2D_BODY_JOBP job_partials
#define NDT2D_BODY_CHUNKS chunk_sums
#include "search_region_body.inc"
#undef MV
#undef SV
#undef NDT2D_BODY_BATCH
#undef NDT2D_BODY_TOTAL_WORK
#undef NDT2D_BODY_OCCD
#undef NDT2D_BODY_THRX
#undef NDT2D_BODY_THRY
#undef NDT2D_BODY_JOBP
#undef NDT2D_BODY_CHUNKS
}

static_assert(sizeof(BatchEntry) <= kBatchSlotBytes && sizeof(BatchEntry) % 4 == 0, "batch slot size");

// The same job loop over SEVERAL searches in one launch: the job counter enumerates
// (search, job, chunk) triples, a warp copies the descriptor of the search it is working for
// into a shared-memory slot.  Tables are read through L1 (they differ per search), there is
// no coordinate pre-pass.  All searches share the lattice (theta_begin = 0, stride 1).
__global__ void __launch_bounds__(kWarps * 32, 1)
search_region_batch_kernel(
  const BatchEntry * __restrict__ batch, uint32_t n_batch, uint32_t RX, uint32_t RY, uint32_t Qx,
  uint32_t Qy, uint32_t n_jobs, uint32_t P, uint32_t chunk_points,
  uint32_t * __restrict__ job_counter, unsigned long long * __restrict__ stats)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  unsigned char * sp = smem_raw + static_cast<size_t>(warp) * (kWarpSmemBytes + kBatchSlotBytes);
  double * tot = reinterpret_cast<double *>(sp);
  BatchEntry * slot = reinterpret_cast<BatchEntry *>(sp + kWarpSmemBytes);
  uint32_t cur_search = 0xffffffffu;
  constexpr bool PRE = false, SMEM_TAB = false, STATS = false;
  const uint32_t theta_begin = 0;
  double * const scores = nullptr;
  const uint32_t * const coords = nullptr;
  const uint32_t n_pts_pad = 0;
  (void)coords;
  (void)n_pts_pad;

#define MV (slot->mv)
#define SV (slot->sv)
#define NDT2D_BODY_BATCH 1
#define NDT2D_BODY_TOTAL_WORK (n_batch * n_jobs * P)
#define NDT2D_BODY_OCCD (slot->mv.occ_dilated)
#define NDT2D_BODY_THRX (slot->mv.thr_x)
#define NDT2D_BODY_THRY (slot->mv.thr_y)
#define NDT2D_BODY_JOBP (slot->job_partials)
#define NDT2D_BODY_CHUNKS (slot->chunk_sums)
#include "search_region_body.inc"
#undef MV
#undef SV
#undef NDT2D_BODY_BATCH
#undef NDT2D_BODY_TOTAL_WORK
#undef NDT2D_BODY_OCCD
#undef NDT2D_BODY_THRX
#undef NDT2D_BODY_THRY
#undef NDT2D_BODY_JOBP
#undef NDT2D_BODY_CHUNKS
}

// Largest region of this lattice: RY rows inside one cell ((RY - 1) * step < cell), RX
// columns inside two ((RX - 1) * step < 2 cells).
void full_region(double cell_size, double linear_res, uint32_t n_lin, uint32_t * RX, uint32_t * RY)
{
  uint32_t ry = 1, rx = 1;
  if (linear_res > 0.0 && cell_size > 0.0) {
    const double ratio = cell_size / linear_res * (1.0 - 1e-9);
    ry = ratio >= static_cast<double>(kMaxRY) ? kMaxRY : static_cast<uint32_t>(ratio) + 1u;
    rx = 2.0 * ratio >= static_cast<double>(kMaxRX) ? kMaxRX : static_cast<uint32_t>(2.0 * ratio) + 1u;
  }
  if (ry > n_lin) {ry = n_lin ? n_lin : 1u;}
  if (rx > n_lin) {rx = n_lin ? n_lin : 1u;}
  *RX = rx;
  *RY = ry;
}

// n_searches > 1: a batch of searches of this shape shares the launch, so regions shrink (and
// points get chunked) only as far as the whole batch needs to fill the machine.
RegionPlan make_plan(const GridDesc & g, uint32_t n_theta, uint32_t n_lin, double linear_res,
  uint32_t n_searches = 1)
{
  RegionPlan pl{};
  uint32_t RX = 1, RY = 1;
  full_region(g.cell_size, linear_res, n_lin, &RX, &RY);
  // small searches: more, smaller regions so that every SM gets work (rows first: the
  // columns are the lanes)
  for (;;) {
    const uint32_t qx = (n_lin + RX - 1) / RX, qy = (n_lin + RY - 1) / RY;
    if (static_cast<uint64_t>(n_theta) * qx * qy * n_searches >= kTargetJobs) {break;}
    if (RY > 6) {
      RY = (RY + 1) / 2;
    } else if (RX > 8) {
      RX = (RX + 1) / 2;
    } else {
      break;
    }
  }
  pl.RX = RX;
  pl.RY = RY;
  pl.Qx = (n_lin + RX - 1) / RX;
  pl.Qy = (n_lin + RY - 1) / RY;
  pl.n_jobs = n_theta * pl.Qx * pl.Qy;
  pl.thr_doubles = g.size_x + 2 + g.size_y + 2;
  const size_t d_bytes = (static_cast<size_t>(g.n_words) * 4 + 15) & ~size_t(15);
  const size_t t_bytes = (static_cast<size_t>(pl.thr_doubles) * 8 + 15) & ~size_t(15);
  pl.smem_tab = d_bytes + t_bytes <= kSmemTabBudget;
  pl.tab_bytes = pl.smem_tab ? static_cast<uint32_t>(d_bytes + t_bytes) : 0u;
  pl.smem_bytes = 16 + pl.tab_bytes + static_cast<size_t>(kWarpSmemBytes) * kWarps;
  pl.P = 1;
  pl.chunk_points = 0;
  return pl;
}

// Small searches: split the scan points of every job into chunks until there are
// about kChunkTargetWork (job, chunk) pairs, bounded by the chunk-sum scratch.
void plan_chunks(RegionPlan & pl, uint32_t n_pts, size_t chunk_cap_doubles, uint32_t n_searches = 1)
{
  pl.P = 1;
  pl.chunk_points = (n_pts + 31u) & ~31u;
  const uint32_t steps = (n_pts + 31u) / 32u;
  const uint64_t jobs_in_flight = static_cast<uint64_t>(pl.n_jobs) * n_searches;
  if (pl.n_jobs == 0 || steps < 2 || jobs_in_flight >= kChunkTargetWork) {return;}
  uint32_t P = static_cast<uint32_t>((kChunkTargetWork + jobs_in_flight - 1) / jobs_in_flight);
  if (P > steps) {P = steps;}
  const size_t per_chunk = static_cast<size_t>(pl.n_jobs) * 32u * pl.RY;
  if (per_chunk * P > chunk_cap_doubles) {P = static_cast<uint32_t>(chunk_cap_doubles / per_chunk);}
  if (P < 2) {return;}
  const uint32_t spc = (steps + P - 1) / P;
  pl.P = (steps + spc - 1) / spc;
  pl.chunk_points = spc * 32u;
}

size_t coords_bytes(const RegionPlan & pl, uint32_t n_theta, uint32_t n_pts)
{
  const size_t n_pts_pad = (static_cast<size_t>(n_pts) + 31u) & ~size_t(31);
  return static_cast<size_t>(n_theta) * (pl.Qx + pl.Qy) * n_pts_pad * sizeof(uint32_t);
}

template<bool S, bool PRE, bool STATS>
int launch_one(RegionPlan & pl, const ModelView & mv, const SearchView & sv,
  uint32_t theta_begin, uint32_t n_theta, double * d_job_partials, double * d_scores,
  uint32_t * d_counter, uint32_t * d_coords, double * d_chunk_sums, cudaStream_t stream,
  Counters * ctr)
{
  auto kernel = search_region_kernel<S, PRE, STATS>;
  // per (instantiation, device): opt in to the large dynamic shared memory once, and
  // remember the SM count (both calls cost microseconds that a 100 us search notices)
  static int configured_sms[64] = {0};
  int dev = 0;
  NDT2D_CUDA_TRY(cudaGetDevice(&dev));
  const int slot = dev & 63;
  if (configured_sms[slot] == 0) {
    int n = 148;
    NDT2D_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
      static_cast<int>(16 + kSmemTabBudget + static_cast<size_t>(kWarpSmemBytes) * kWarps)));
    NDT2D_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    configured_sms[slot] = n;
  }
  const int sms = configured_sms[slot];
  // one persistent CTA per SM; small searches still spread over all SMs (the
  // warps of every CTA draw jobs from the same counter)
  pl.grid = min(pl.n_jobs * pl.P, static_cast<uint32_t>(sms));
  // d_counter: [0] job counter (u32, + pad), [1..2] u64 statistics of this launch; zero on
  // entry -- cleared at allocation and again by the finish kernel of every search, which
  // first moves the statistics to [3..4] for ndt2d_matcher_search_stats
  const uint32_t n_pts_pad = (sv.n_pts + 31u) & ~31u;
  if (PRE) {
    dim3 grid((sv.n_pts + 127u) / 128u, 2u, (n_theta + kCoordsSlices - 1u) / kCoordsSlices);
    region_coords_kernel<<<grid, 128, 0, stream>>>(mv, sv, theta_begin, n_theta, pl.RX, pl.RY,
      pl.Qx, pl.Qy, n_pts_pad, d_coords);
    NDT2D_LAUNCH_CHECK(ctr);
  }
  const uint32_t d_bytes = pl.smem_tab ? ((mv.g.n_words * 4u + 15u) & ~15u) : 0u;
  const uint32_t t_bytes = pl.smem_tab ? pl.tab_bytes - d_bytes : 0u;
  kernel<<<pl.grid, kWarps * 32, pl.smem_bytes, stream>>>(
    mv, sv, theta_begin, pl.RX, pl.RY, pl.Qx, pl.Qy, pl.n_jobs, d_bytes, t_bytes, d_job_partials,
    d_scores, d_counter, reinterpret_cast<unsigned long long *>(d_counter) + 1, d_coords,
    n_pts_pad, pl.P, pl.chunk_points, d_chunk_sums);
  NDT2D_LAUNCH_CHECK(ctr);
  if (pl.P > 1) {
    region_chunk_reduce_kernel<<<(pl.n_jobs + 7u) / 8u, 256, 0, stream>>>(
      sv, theta_begin, pl.RX, pl.RY, pl.Qx, pl.Qy, pl.n_jobs, pl.P, d_chunk_sums, d_job_partials,
      d_scores);
    NDT2D_LAUNCH_CHECK(ctr);
  }
  return NDT2D_OK;
}

// region_chunk_reduce_kernel for a batch: one warp per (search, job).
__global__ void __launch_bounds__(256) region_chunk_reduce_batch_kernel(
  const BatchEntry * __restrict__ batch, uint32_t n_batch, uint32_t RX, uint32_t RY, uint32_t Qx,
  uint32_t Qy, uint32_t n_jobs, uint32_t P)
{
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (gw >= n_batch * n_jobs) {return;}
  const uint32_t bj = gw / n_jobs, job = gw - bj * n_jobs;
  const BatchEntry & e = batch[bj];
  const uint32_t QQ = Qx * Qy, RR = 32u * RY;
  const uint32_t it = job / QQ, rr = job - it * QQ;
  const uint32_t rx = rr / Qy, ry = rr - rx * Qy;
  const uint32_t jx0 = rx * RX, jy0 = ry * RY;
  const uint32_t n_lin = e.sv.n_lin;
  const uint32_t nxc = min(RX, n_lin - jx0), nyc = min(RY, n_lin - jy0);
  double * first = e.chunk_sums + static_cast<size_t>(job) * P * RR;
  for (uint32_t k = lane; k < RR; k += 32) {
    double t = first[k];
    for (uint32_t c = 1; c < P; ++c) {t += first[static_cast<size_t>(c) * RR + k];}
    first[k] = t;
  }
  __syncwarp();
  job_epilogue([first, lane](uint32_t b) {return first[b * 32u + lane];}, e.sv, job, it, jx0, jy0,
    nxc, nyc, lane, e.job_partials, nullptr);
}

}  // namespace

uint32_t ndt2d_region_dilate_x(double cell_size, double linear_res, uint32_t n_lin)
{
  uint32_t RX = 1, RY = 1;
  full_region(cell_size, linear_res, n_lin ? n_lin : 1u, &RX, &RY);
  // columns of one region reach into a third cell column once (RX - 1) * step >= cell
  return (linear_res > 0.0 && static_cast<double>(RX - 1u) * linear_res >= cell_size * (1.0 - 1e-9)) ?
         2u : 1u;
}

int ndt2d_region_batch_plan(double cell_size, uint32_t n_ang, uint32_t n_lin, double linear_res,
  uint32_t max_pts, uint32_t n_searches, RegionBatchPlan * out)
{
  GridDesc g{};
  g.cell_size = cell_size;
  RegionPlan pl = make_plan(g, n_ang, n_lin, linear_res, n_searches ? n_searches : 1u);
  // chunk the points only while the batch as a whole is short of work, capped so that a
  // search's chunk sums stay small
  plan_chunks(pl, max_pts, size_t(1) << 22, n_searches ? n_searches : 1u);
  out->RX = pl.RX;
  out->RY = pl.RY;
  out->Qx = pl.Qx;
  out->Qy = pl.Qy;
  out->n_jobs = pl.n_jobs;
  out->P = pl.P;
  out->chunk_points = pl.chunk_points;
  out->chunk_doubles = pl.P > 1 ? static_cast<size_t>(pl.n_jobs) * pl.P * 32u * pl.RY : 0;
  return NDT2D_OK;
}

int ndt2d_launch_search_region_batch(
  const BatchEntry * d_batch, uint32_t n_batch, const RegionBatchPlan & pl, uint32_t * d_counter,
  cudaStream_t stream, Counters * ctr)
{
  if (n_batch == 0 || pl.n_jobs == 0) {return NDT2D_OK;}
  static int configured_sms[64] = {0};
  int dev = 0;
  NDT2D_CUDA_TRY(cudaGetDevice(&dev));
  const int slot = dev & 63;
  const size_t smem = static_cast<size_t>(kWarps) * (kWarpSmemBytes + kBatchSlotBytes);
  if (configured_sms[slot] == 0) {
    int n = 148;
    NDT2D_CUDA_TRY(cudaFuncSetAttribute(search_region_batch_kernel,
      cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    NDT2D_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    configured_sms[slot] = n;
  }
  const uint64_t total = static_cast<uint64_t>(n_batch) * pl.n_jobs * pl.P;
  if (total >= (1ull << 32)) {return NDT2D_ERR_SIZE;}
  const uint32_t grid = static_cast<uint32_t>(std::min<uint64_t>(total, configured_sms[slot]));
  search_region_batch_kernel<<<grid, kWarps * 32, smem, stream>>>(
    d_batch, n_batch, pl.RX, pl.RY, pl.Qx, pl.Qy, pl.n_jobs, pl.P, pl.chunk_points, d_counter,
    reinterpret_cast<unsigned long long *>(d_counter) + 1);
  NDT2D_LAUNCH_CHECK(ctr);
  if (pl.P > 1) {
    const uint32_t warps = n_batch * pl.n_jobs;
    region_chunk_reduce_batch_kernel<<<(warps + 7u) / 8u, 256, 0, stream>>>(
      d_batch, n_batch, pl.RX, pl.RY, pl.Qx, pl.Qy, pl.n_jobs, pl.P);
    NDT2D_LAUNCH_CHECK(ctr);
  }
  return NDT2D_OK;
}

size_t ndt2d_region_scratch_doubles(double cell_size, uint32_t n_ang, uint32_t n_lin,
  double linear_res)
{
  GridDesc g{};
  g.cell_size = cell_size;
  // the plan's region size shrinks with the number of theta slices searched (more,
  // smaller regions for small searches); bound the job count of any sub-range
  // [nt, 2 nt) by the plan of nt slices applied to 2 nt slices
  const uint32_t na = n_ang ? n_ang : 1;
  size_t worst = 0;
  for (uint64_t nt = 1;; nt *= 2) {
    const uint32_t t = static_cast<uint32_t>(nt < na ? nt : na);
    const RegionPlan pl = make_plan(g, t, n_lin ? n_lin : 1, linear_res);
    const uint64_t upto = (2 * nt < na) ? 2 * nt : na;
    const size_t jobs = static_cast<size_t>(upto) * pl.Qx * pl.Qy;
    worst = jobs > worst ? jobs : worst;
    if (nt >= na) {break;}
  }
  return worst * NDT2D_BLOCK_PARTIAL + 8;
}

size_t ndt2d_region_coords_bytes(double cell_size, uint32_t n_ang, uint32_t n_lin,
  double linear_res, uint32_t n_pts, size_t cap_bytes)
{
  GridDesc g{};
  g.cell_size = cell_size;
  // sized for the full theta range; a launch over a sub-range (other region plan)
  // uses the table only if its own needs fit (ndt2d_launch_search_region)
  const RegionPlan pl = make_plan(g, n_ang ? n_ang : 1, n_lin ? n_lin : 1, linear_res);
  const size_t worst = pl.Qx >= 3 ? coords_bytes(pl, n_ang ? n_ang : 1, n_pts) : 0;
  return worst <= cap_bytes ? worst : 0;  // 0: the search computes coordinates per job
}

size_t ndt2d_region_chunk_doubles(double cell_size, uint32_t n_ang, uint32_t n_lin,
  double linear_res, uint32_t n_pts)
{
  // any theta sub-range: at most ~2 * kChunkTargetWork (job, chunk) pairs of 32 x RY sums
  // (regions of small searches are shrunk to <= 13 rows before chunking matters)
  (void)cell_size;
  (void)n_ang;
  (void)n_lin;
  (void)linear_res;
  if (n_pts < 64) {return 0;}
  return static_cast<size_t>(2) * kChunkTargetWork * 32 * 13;
}

int ndt2d_launch_search_region(
  const ModelView & mv, const SearchView & sv, double linear_res, uint32_t theta_begin,
  uint32_t n_theta, double * d_job_partials, double * d_scores, uint32_t * d_counter,
  uint32_t * d_coords, size_t coords_cap_bytes, cudaStream_t stream, Counters * ctr,
  uint32_t * n_jobs)
{
  RegionPlan pl = make_plan(mv.g, n_theta, sv.n_lin, linear_res);
  *n_jobs = pl.n_jobs;
  plan_chunks(pl, sv.n_pts, sv.chunk_sums ? sv.chunk_cap_doubles : 0);
  // the pre-pass pays off when a slice has several regions per axis to share it
  const bool pre = d_coords && pl.Qx >= 3 && sv.n_pts > 0 &&
    coords_bytes(pl, n_theta, sv.n_pts) <= coords_cap_bytes;
#define NDT2D_REGION_LAUNCH(S, P) \
  (sv.tally ? \
   launch_one<S, P, true>(pl, mv, sv, theta_begin, n_theta, d_job_partials, d_scores, d_counter, \
     d_coords, sv.chunk_sums, stream, ctr) : \
   launch_one<S, P, false>(pl, mv, sv, theta_begin, n_theta, d_job_partials, d_scores, d_counter, \
     d_coords, sv.chunk_sums, stream, ctr))
  if (pl.smem_tab) {
    return pre ? NDT2D_REGION_LAUNCH(true, true) : NDT2D_REGION_LAUNCH(true, false);
  }
  return pre ? NDT2D_REGION_LAUNCH(false, true) : NDT2D_REGION_LAUNCH(false, false);
#undef NDT2D_REGION_LAUNCH
}
