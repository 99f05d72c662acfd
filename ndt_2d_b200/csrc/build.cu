// build.cu -- NDT model construction on the device (kernels K1..K3).
//
// Replaces NDT::addScan / Cell::addPoint / NDT::compute / Cell::compute
// (ndt_model.cpp:50-103, 132-160) with a sort-by-cell-key segmented reduction:
//
//   K1  transform_key   every map point -> world coordinates + cell key
//   K2  radix sort      STABLE LSD sort of (key, point index), 8-bit digits
//   K3a segment_count   per occupied cell: point count, occupancy bit if n >= 5
//       scan            rank prefix over the occupancy words
//   K3b segment_moments per occupied cell with n >= 5: the reference's running
//                       mean / second-moment recurrence replayed in order,
//                       covariance, eigenvalue clamp, information -> packed record
//
// Because the sort is stable, the points of one cell stay in (scan, point)
// order, so the sequential recurrence of Cell::addPoint is reproduced term by
// term: cell statistics are bit-identical to a sequential CPU build.  No float
// atomics anywhere (only an integer atomicOr for the occupancy bit).
//
// All arithmetic that decides a cell index or enters the recurrence uses the
// round-to-nearest intrinsics (__dadd_rn, __dmul_rn, __ddiv_rn), which nvcc
// never contracts into FMAs -- the reference's x86-64 build has none either.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ndt2d_internal.h"
#include "build_common.cuh"

namespace
{

using ndt2d_dev::div_by_count;

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per block
constexpr int kSortWarps = kSortThreads / 32;

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 8192 values per block

// ------------------------------------------------------------------ K1
// NDT::getIndex (ndt_model.cpp:203-218); returns n_cells for "outside".
__device__ __forceinline__ uint32_t cell_key(const GridDesc & g, double x, double y)
{
  if (x < g.origin_x || y < g.origin_y) {
    return g.n_cells;
  }
  const uint32_t gx = __double2uint_rz(__ddiv_rn(__dsub_rn(x, g.origin_x), g.cell_size));
  const uint32_t gy = __double2uint_rz(__ddiv_rn(__dsub_rn(y, g.origin_y), g.cell_size));
  if (gx >= g.size_x || gy >= g.size_y) {
    return g.n_cells;
  }
  return gy * g.size_x + gx;
}

// Digits of the radix sort: `passes` digits of `width` <= 8 bits cover bits_needed(n_cells).
struct DigitPlan
{
  int passes, width;
  uint32_t mask;
};
constexpr int kMaxPasses = 4;

// NDT::addScan (ndt_model.cpp:132-152): p = pose; p += R(theta) * point.  The block also
// histograms every digit of its keys (shared-memory counters, flushed to the global
// histograms at the end): the one-sweep sort below needs the digit totals of all passes up
// front and the keys are in registers here anyway.
__global__ void __launch_bounds__(128) transform_key_kernel(
  GridDesc g, const double4 * __restrict__ scan_tf, const uint64_t * __restrict__ offsets,
  uint32_t n_scans, const double2 * __restrict__ pts, double * __restrict__ wx,
  double * __restrict__ wy, uint32_t * __restrict__ key, uint32_t * __restrict__ val,
  DigitPlan dp, uint32_t * __restrict__ ghist)
{
  __shared__ uint32_t h[kMaxPasses][256];
  for (uint32_t k = threadIdx.x; k < kMaxPasses * 256; k += blockDim.x) {(&h[0][0])[k] = 0u;}
  __syncthreads();
  for (uint32_t s = blockIdx.x; s < n_scans; s += gridDim.x) {
    const double4 tf = scan_tf[s];  // x, y, cos, sin
    const uint64_t lo = offsets[s], hi = offsets[s + 1];
    for (uint64_t p = lo + threadIdx.x; p < hi; p += blockDim.x) {
      const double2 pt = pts[p];
      const double X =
        __dadd_rn(tf.x, __dsub_rn(__dmul_rn(pt.x, tf.z), __dmul_rn(pt.y, tf.w)));
      const double Y =
        __dadd_rn(tf.y, __dadd_rn(__dmul_rn(pt.x, tf.w), __dmul_rn(pt.y, tf.z)));
      wx[p] = X;
      wy[p] = Y;
      const uint32_t k = cell_key(g, X, Y);
      key[p] = k;
      val[p] = static_cast<uint32_t>(p);
#pragma unroll
      for (int q = 0; q < kMaxPasses; ++q) {
        if (q < dp.passes) {atomicAdd(&h[q][(k >> (q * dp.width)) & dp.mask], 1u);}
      }
    }
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < static_cast<uint32_t>(dp.passes) * 256u; k += blockDim.x) {
    const uint32_t c = (&h[0][0])[k];
    if (c) {atomicAdd(ghist + k, c);}
  }
}

// ------------------------------------------------------------------ scan
// Exclusive scan of uint32 values, three phases for arrays above one tile.
// MODE 0: plain array in place.  MODE 1: uint2 occupancy words, in = popc(.x),
// out -> .y.
template<int MODE>
__device__ __forceinline__ uint32_t scan_load(const void * data, size_t i)
{
  if (MODE == 0) {
    return static_cast<const uint32_t *>(data)[i];
  }
  return __popc(static_cast<const uint2 *>(data)[i].x);
}
template<int MODE>
__device__ __forceinline__ void scan_store(void * data, size_t i, uint32_t v)
{
  if (MODE == 0) {
    static_cast<uint32_t *>(data)[i] = v;
  } else {
    static_cast<uint2 *>(data)[i].y = v;
  }
}

__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t * total)
{
  __shared__ uint32_t warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) {incl += t;}
  }
  if (lane == 31) {warp_sums[warp] = incl;}
  __syncthreads();
  if (warp == 0) {
    uint32_t ws = warp_sums[lane];
    uint32_t wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) {wi += t;}
    }
    warp_sums[lane] = wi - ws;  // exclusive warp prefix
    if (lane == 31) {*total = wi;}
  }
  __syncthreads();
  const uint32_t r = warp_sums[warp] + incl - v;
  __syncthreads();
  return r;
}

template<int MODE>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(
  const void * data, size_t n, uint32_t * block_sums)
{
  __shared__ uint32_t total;
  const size_t base = static_cast<size_t>(blockIdx.x) * kScanTile + threadIdx.x * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) {s += scan_load<MODE>(data, base + k);}
  }
  block_exclusive_scan_1024(s, &total);
  if (threadIdx.x == 0) {block_sums[blockIdx.x] = total;}
}

// single block: exclusive scan of up to kScanTile block sums, in place
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(uint32_t * sums, uint32_t n)
{
  __shared__ uint32_t total;
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) {carry_s = 0;}
  __syncthreads();
  for (uint32_t start = 0; start < n; start += kScanTile) {
    uint32_t v[kScanItems];
    uint32_t s = 0;
    const uint32_t base = start + threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      v[k] = (base + k < n) ? sums[base + k] : 0u;
      s += v[k];
    }
    uint32_t ex = block_exclusive_scan_1024(s, &total) + carry_s;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      if (base + k < n) {sums[base + k] = ex;}
      ex += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {carry_s += total;}
    __syncthreads();
  }
}

template<int MODE>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(
  void * data, size_t n, const uint32_t * block_sums)
{
  __shared__ uint32_t total;
  const size_t base = static_cast<size_t>(blockIdx.x) * kScanTile + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? scan_load<MODE>(data, base + k) : 0u;
    s += v[k];
  }
  uint32_t ex = block_exclusive_scan_1024(s, &total) + (block_sums ? block_sums[blockIdx.x] : 0u);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) {scan_store<MODE>(data, base + k, ex);}
    ex += v[k];
  }
}

template<int MODE>
int launch_scan(void * d_data, size_t n, uint32_t * d_tmp, cudaStream_t stream, Counters * ctr)
{
  if (n == 0) {return NDT2D_OK;}
  const uint32_t nblk = static_cast<uint32_t>((n + kScanTile - 1) / kScanTile);
  if (nblk == 1) {
    scan_apply_kernel<MODE><<<1, kScanThreads, 0, stream>>>(d_data, n, nullptr);
    NDT2D_LAUNCH_CHECK(ctr);
    return NDT2D_OK;
  }
  scan_reduce_kernel<MODE><<<nblk, kScanThreads, 0, stream>>>(d_data, n, d_tmp);
  NDT2D_LAUNCH_CHECK(ctr);
  scan_sums_kernel<<<1, kScanThreads, 0, stream>>>(d_tmp, nblk);
  NDT2D_LAUNCH_CHECK(ctr);
  scan_apply_kernel<MODE><<<nblk, kScanThreads, 0, stream>>>(d_data, n, d_tmp);
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}

// ------------------------------------------------------------------ K2
// One-sweep LSD radix sort (Adinets & Merrill's "onesweep" scheme), stable: per digit ONE
// kernel reads the (key, value) pairs once and writes them once.
//   * the digit totals of every pass come from transform_key_kernel (ghist), so a block knows
//     where each digit's output range starts after a 256-entry scan of its own;
//   * a block takes its tile from a ticket counter (tiles are processed in ticket order, so a
//     block only ever waits for blocks that started before it), ranks its 4096 pairs stably
//     (warp w owns 512 consecutive pairs and walks them in order: __match_any_sync + running
//     per-warp digit counters), publishes its per-digit counts in status[tile][digit] and adds
//     up its predecessors' by decoupled look-back (thread d follows digit d: a word is
//     `flag << 30 | count`, flag 1 = this tile's count alone, 2 = inclusive prefix up to this
//     tile) -- no histogram / scan launches between the passes;
//   * the tile is first permuted into digit order in shared memory, then written out: the runs
//     of one digit are contiguous in the output, so consecutive threads store to consecutive
//     addresses.
constexpr uint32_t kOsFlagAgg = 1u << 30, kOsFlagInc = 2u << 30, kOsCount = (1u << 30) - 1u;

struct OnesweepSmem
{
  uint32_t cnt[kSortWarps][256];   // per-warp digit counters -> per-warp exclusive offsets
  uint32_t toff[256];              // start of digit d's run in the sorted tile
  uint32_t gbase[256];             // global position of tile-sorted index j of digit d: gbase[d] + j
  uint32_t keys[kSortTile];
  uint32_t vals[kSortTile];
  uint32_t tile;
  uint32_t warp_sums[kSortWarps];
};

// LAST: the final pass also writes the world coordinates in sorted (cell) order -- the moment
// kernel then reads a cell's points contiguously -- instead of a separate gather launch.
template<bool LAST>
__global__ void __launch_bounds__(kSortThreads, 3) radix_onesweep_kernel(
  const uint32_t * __restrict__ keys_in, const uint32_t * __restrict__ vals_in, uint32_t n,
  int shift, uint32_t mask, const uint32_t * __restrict__ ghist, uint32_t * __restrict__ status,
  uint32_t * __restrict__ ticket, uint32_t * __restrict__ keys_out, uint32_t * __restrict__ vals_out,
  const double * __restrict__ wx, const double * __restrict__ wy, double * __restrict__ sx,
  double * __restrict__ sy)
{
  extern __shared__ __align__(16) unsigned char os_raw[];
  OnesweepSmem & sm = *reinterpret_cast<OnesweepSmem *>(os_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) {sm.tile = atomicAdd(ticket, 1u);}
  for (uint32_t i = tid; i < kSortWarps * 256; i += kSortThreads) {(&sm.cnt[0][0])[i] = 0u;}
  __syncthreads();
  const uint32_t tile = sm.tile;

  // ---- stable ranks inside the tile
  const uint32_t sub = tile * kSortTile + warp * 32u * kSortItems;
  uint32_t k[kSortItems], v[kSortItems], local[kSortItems];
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int s = 0; s < kSortItems; ++s) {
    const uint32_t i = sub + static_cast<uint32_t>(s) * 32u + lane;
    const bool active = i < n;
    k[s] = active ? keys_in[i] : 0u;
    v[s] = active ? vals_in[i] : 0u;
    // inactive lanes get a private pseudo digit so they match nobody
    const uint32_t d = active ? ((k[s] >> shift) & mask) : (256u + lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t r = __popc(peers & lt_mask);
    uint32_t base = 0;
    if (active) {base = sm.cnt[warp][d];}
    __syncwarp();
    if (active && r == 0) {sm.cnt[warp][d] = base + __popc(peers);}
    __syncwarp();
    local[s] = base + r;
  }
  __syncthreads();

  // ---- thread d: digit d's count in this tile, per-warp offsets, the look-back
  {
    const uint32_t d = tid;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t t = sm.cnt[w][d];
      sm.cnt[w][d] = run;
      run += t;
    }
    const uint32_t count = run;
    volatile uint32_t * st = status + static_cast<size_t>(tile) * 256u + d;
    *st = (tile == 0u ? kOsFlagInc : kOsFlagAgg) | count;
    // exclusive scans over the digits: of the tile's counts (run starts in the sorted tile) and
    // of the global totals (where digit d's output range starts)
    const uint32_t total = ghist[d];
    uint32_t in_t = count, in_g = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t a = __shfl_up_sync(0xffffffffu, in_t, o);
      const uint32_t b = __shfl_up_sync(0xffffffffu, in_g, o);
      if (lane >= static_cast<uint32_t>(o)) {
        in_t += a;
        in_g += b;
      }
    }
    __shared__ uint32_t ws_t[kSortWarps], ws_g[kSortWarps];
    if (lane == 31u) {
      ws_t[warp] = in_t;
      ws_g[warp] = in_g;
    }
    __syncthreads();
    uint32_t off_t = 0, off_g = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      if (static_cast<uint32_t>(w) < warp) {
        off_t += ws_t[w];
        off_g += ws_g[w];
      }
    }
    const uint32_t toff = off_t + in_t - count, gdigit = off_g + in_g - total;
    // look back over the tiles before this one
    uint32_t excl = 0;
    if (tile > 0u) {
      uint32_t t = tile - 1u;
      for (;;) {
        volatile const uint32_t * ps = status + static_cast<size_t>(t) * 256u + d;
        uint32_t w;
        do {w = *ps;} while ((w >> 30) == 0u);
        excl += w & kOsCount;
        if ((w >> 30) == 2u || t == 0u) {break;}
        --t;
      }
      *st = kOsFlagInc | (excl + count);
    }
    sm.toff[d] = toff;
    sm.gbase[d] = gdigit + excl - toff;
  }
  __syncthreads();

  // ---- permute the tile into digit order in shared memory
#pragma unroll
  for (int s = 0; s < kSortItems; ++s) {
    const uint32_t i = sub + static_cast<uint32_t>(s) * 32u + lane;
    if (i < n) {
      const uint32_t d = (k[s] >> shift) & mask;
      const uint32_t pos = sm.toff[d] + sm.cnt[warp][d] + local[s];
      sm.keys[pos] = k[s];
      sm.vals[pos] = v[s];
    }
  }
  __syncthreads();
  // ---- write out: tile-sorted index j of digit d goes to gbase[d] + j
  const uint32_t n_tile = min(static_cast<uint32_t>(kSortTile), n - tile * kSortTile);
  for (uint32_t j = tid; j < n_tile; j += kSortThreads) {
    const uint32_t kk = sm.keys[j];
    const uint32_t pos = sm.gbase[(kk >> shift) & mask] + j;
    const uint32_t p = sm.vals[j];
    keys_out[pos] = kk;
    vals_out[pos] = p;
    if (LAST) {
      sx[pos] = wx[p];
      sy[pos] = wy[p];
    }
  }
}

// ------------------------------------------------------------------ K3
__device__ __forceinline__ uint32_t padded_index(const GridDesc & g, uint32_t key)
{
  const uint32_t gy = key / g.size_x, gx = key - gy * g.size_x;
  return (gy + 1) * g.pitch + gx + 1;
}

// K3a: heads measure their segment (galloping + binary search for the end of the run of
// equal keys: no serial walk); cells with n >= 5 get their occupancy bit and are appended
// to the list of cells that need a record.
__global__ void __launch_bounds__(256) segment_count_kernel(
  GridDesc g, const uint32_t * __restrict__ key, size_t n, uint32_t * __restrict__ seglen,
  uint2 * __restrict__ occ, uint2 * __restrict__ heads, uint32_t * __restrict__ n_heads)
{
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) {return;}
  const uint32_t k = key[i];
  if (k >= g.n_cells) {return;}
  if (i > 0 && key[i - 1] == k) {return;}
  // first position j > i with key[j] != k
  size_t lo = i, step = 1;            // key[lo] == k
  size_t hi = n;                      // key[hi] != k (or hi == n)
  while (lo + step < n) {
    if (key[lo + step] == k) {
      lo += step;
      step <<= 1;
    } else {
      hi = lo + step;
      break;
    }
  }
  while (hi - lo > 1) {
    const size_t mid = lo + ((hi - lo) >> 1);
    if (key[mid] == k) {lo = mid;} else {hi = mid;}
  }
  const uint32_t len = static_cast<uint32_t>(hi - i);
  seglen[i] = len;
  if (len >= 5) {
    const uint32_t p = padded_index(g, k);
    atomicOr(&occ[p >> 5].x, 1u << (p & 31u));
    heads[atomicAdd(n_heads, 1u)] = make_uint2(static_cast<uint32_t>(i), len);
  }
}

struct CellStats
{
  double n, mean[2], corr[3] /*xx, xy, yy*/, cov[3], info[4];
};

// Cell::addPoint recurrence (ndt_model.cpp:50-63), one point.
__device__ __forceinline__ void stats_add(CellStats & c, double x, double y)
{
  const double n = c.n, n1 = __dadd_rn(n, 1.0);
  c.mean[0] = __ddiv_rn(__dadd_rn(__dmul_rn(c.mean[0], n), x), n1);
  c.mean[1] = __ddiv_rn(__dadd_rn(__dmul_rn(c.mean[1], n), y), n1);
  c.corr[0] = __ddiv_rn(__dadd_rn(__dmul_rn(c.corr[0], n), __dmul_rn(x, x)), n1);
  c.corr[1] = __ddiv_rn(__dadd_rn(__dmul_rn(c.corr[1], n), __dmul_rn(x, y)), n1);
  c.corr[2] = __ddiv_rn(__dadd_rn(__dmul_rn(c.corr[2], n), __dmul_rn(y, y)), n1);
  c.n = n1;
}

// Cell::compute (ndt_model.cpp:65-103) for n >= 3.
__device__ __forceinline__ void stats_finalize(CellStats & c)
{
  const double scale = __ddiv_rn(c.n, __dsub_rn(c.n, 1.0));
  c.cov[0] = __dmul_rn(__dsub_rn(c.corr[0], __dmul_rn(c.mean[0], c.mean[0])), scale);
  c.cov[1] = __dmul_rn(__dsub_rn(c.corr[1], __dmul_rn(c.mean[0], c.mean[1])), scale);
  c.cov[2] = __dmul_rn(__dsub_rn(c.corr[2], __dmul_rn(c.mean[1], c.mean[1])), scale);
  // eigenvalues of the symmetric 2x2 (closed form; the reference asks Eigen's
  // EigenSolver, ndt_model.cpp:84-87 -- only the clamp decision depends on it)
  const double a = c.cov[0], b = c.cov[1], d = c.cov[2];
  const double mid = __dmul_rn(0.5, __dadd_rn(a, d));
  const double half = __dmul_rn(0.5, __dsub_rn(a, d));
  const double rad = __dsqrt_rn(__dadd_rn(__dmul_rn(half, half), __dmul_rn(b, b)));
  double small = __dsub_rn(mid, rad), large = __dadd_rn(mid, rad);
  if (small > large) {
    const double t = small;
    small = large;
    large = t;
  }
  if (small < __dmul_rn(0.001, large)) {
    const double det = __dmul_rn(__dmul_rn(0.001, large), large);
    c.info[0] = __ddiv_rn(d, det);
    c.info[1] = __ddiv_rn(-b, det);
    c.info[2] = __ddiv_rn(-b, det);
    c.info[3] = __ddiv_rn(a, det);
  } else {
    const double invdet = __ddiv_rn(1.0, __dsub_rn(__dmul_rn(a, d), __dmul_rn(b, b)));
    c.info[0] = __dmul_rn(d, invdet);
    c.info[1] = __dmul_rn(-b, invdet);
    c.info[2] = __dmul_rn(-b, invdet);
    c.info[3] = __dmul_rn(a, invdet);
  }
}

__device__ __forceinline__ void stats_walk(
  CellStats & c, const double * __restrict__ sx, const double * __restrict__ sy, size_t i,
  uint32_t len)
{
  c.n = 0.0;
  c.mean[0] = c.mean[1] = 0.0;
  c.corr[0] = c.corr[1] = c.corr[2] = 0.0;
  for (uint32_t j = 0; j < len; ++j) {
    stats_add(c, sx[i + j], sy[i + j]);
  }
}

// The packed records of one finished cell (layout: ndt2d_internal.h, ModelView).
__device__ __forceinline__ void write_records(
  const CellStats & c, const GridDesc & g, double * __restrict__ rr, double * __restrict__ f,
  double * __restrict__ vt, uint32_t * __restrict__ n_stiff)
{
  // -0.5 * information: scaling by a power of two is exact, so the device
  // exponent (-0.5 q)^T I q keeps the reference's rounding term by term.
  rr[0] = c.mean[0];
  rr[1] = c.mean[1];
  rr[2] = -0.5 * c.info[0];  // (0,0)
  rr[3] = -0.5 * c.info[2];  // (1,0)
  rr[4] = -0.5 * c.info[1];  // (0,1)
  rr[5] = -0.5 * c.info[3];  // (1,1)
  // short form of the window / dense search kernels
  constexpr double kLog2e = 1.44269504088896340736;
  const double mag = fmax(fmax(fabs(c.info[0]), fabs(c.info[3])), fmax(fabs(c.info[1]), fabs(c.info[2])));
  const double A = rr[2] * kLog2e, B = (rr[3] + rr[4]) * kLog2e, D = rr[5] * kLog2e;
  const bool stiff = !(mag * (g.cell_size * g.cell_size) <= 1.0e7);  // NaN -> stiff
  f[0] = c.mean[0];
  f[1] = c.mean[1];
  f[2] = A;
  f[3] = B;
  f[4] = D;
  f[5] = stiff ? 1.0 : 0.0;
  // vertex form of the region kernel: log2 L = D (qy + Bh qx)^2 + S qx^2 ; needs D < 0
  const bool vstiff = stiff || !(D < 0.0);
  const double Bh = B / (2.0 * D), S = A - (B * B) / (4.0 * D);
  vt[0] = c.mean[0];
  vt[1] = c.mean[1];
  vt[2] = D;
  vt[3] = Bh;
  vt[4] = S;
  const float2 tail = make_float2(static_cast<float>(D * (g.lin_res * g.lin_res)), vstiff ? 1.0f : 0.0f);
  *reinterpret_cast<float2 *>(vt + 5) = tail;
  // the region kernel skips its per-item stiff-cell vote when the model has none
  if (vstiff) {atomicAdd(n_stiff, 1u);}
}

// K3b: the cells of the head list replay Cell::addPoint's recurrence and emit their
// records.  The five running quantities (mean x, mean y, second moments xx, xy, yy) are
// independent recurrences of the same form v <- (v * n + term) / (n + 1), so a cell is
// given to a group of 5 lanes, each carrying one of them through the cell's points in order
// (bit-identical to the sequential CPU build, one divide per point on the critical path
// instead of five); 6 cells per warp (lanes 30, 31 idle), warps stride over the head list
// (its length is only known on the device: the grid is sized for the machine, not for the
// worst case).  The kernel is issue-bound -- config 5: 51,725 cells of 5..469 points --, so
// what counts is instructions per step and lanes busy per instruction
// (profiles/r01_build_kernels_config5.md): a lane's term is ONE product a[i] * b[i * stride]
// -- (x, 1), (y, 1), (x, x), (x, y), (y, y) picked by two lane-invariant pointers, x * 1.0 being
// exact -- and the divide by the point count is div_by_count with the reciprocal read from a
// table (rcp[0] = 1.0 doubles as the unit factor; counts beyond the table compute it in line).
constexpr uint32_t kMomentCellsPerWarp = 6;

__global__ void rcp_table_kernel(double * __restrict__ rcp)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= NDT2D_RCP_TABLE) {rcp[k] = k == 0u ? 1.0 : __drcp_rn(static_cast<double>(k));}
}

__global__ void __launch_bounds__(256) segment_moments_kernel(
  GridDesc g, const uint32_t * __restrict__ key, const uint2 * __restrict__ heads,
  const uint32_t * __restrict__ n_heads, const double * __restrict__ sx,
  const double * __restrict__ sy, const double * __restrict__ rcp,
  const uint2 * __restrict__ occ, double * __restrict__ rec,
  double * __restrict__ rec_fast, double * __restrict__ rec_vtx, uint32_t rec_cap,
  uint32_t * __restrict__ n_valid)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) {
    // total number of occupied cells = prefix + popc of the last word
    const uint2 last = occ[g.n_words - 1];
    *n_valid = last.y + __popc(last.x);
  }
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = t >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t sub = lane / 5u, r = lane - sub * 5u;      // lanes 30, 31: sub == 6, no cell
  const uint32_t src = (sub < kMomentCellsPerWarp ? sub : kMomentCellsPerWarp - 1u) * 5u;
  const double * const pa = (r == 1u || r == 4u) ? sy : sx;
  const double * const pb = r < 2u ? rcp : (r == 2u ? sx : sy);
  const size_t sb = r < 2u ? 0u : 1u;
  const uint32_t total = *n_heads;
  for (uint32_t first = warp * kMomentCellsPerWarp; first < total;
    first += n_warps * kMomentCellsPerWarp)
  {
    const uint32_t cell = first + sub;
    const bool have = sub < kMomentCellsPerWarp && cell < total;
    const uint2 h = have ? heads[cell] : make_uint2(0u, 0u);
    const size_t i = h.x;
    const uint32_t len = h.y;
    double v = 0.0, n = 0.0;
    if (have) {
      // the loads of four points are issued before the (dependent) recurrence steps
      // that consume them
      const double * a = pa + i;
      const double * b = pb + i * sb;
      const uint32_t tab = min(len, static_cast<uint32_t>(NDT2D_RCP_TABLE));
      uint32_t j = 0;
      for (; j + 4 <= tab; j += 4) {
        double tm[4], q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          tm[u] = __dmul_rn(a[j + u], b[(j + u) * sb]);
          q[u] = rcp[j + u + 1];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double n1 = __dadd_rn(n, 1.0);
          v = div_by_count(__dadd_rn(__dmul_rn(v, n), tm[u]), n1, q[u]);
          n = n1;
        }
      }
      for (; j < len; ++j) {
        const double term = __dmul_rn(a[j], b[j * sb]);
        const double n1 = __dadd_rn(n, 1.0);
        const double q = j < NDT2D_RCP_TABLE ? rcp[j + 1] : __drcp_rn(n1);
        v = n1 <= 1048576.0 ? div_by_count(__dadd_rn(__dmul_rn(v, n), term), n1, q) :
          __ddiv_rn(__dadd_rn(__dmul_rn(v, n), term), n1);
        n = n1;
      }
    }
    // collect the five quantities in the first lane of the group (all 32 lanes shuffle)
    CellStats c;
    c.n = static_cast<double>(len);
    c.mean[0] = __shfl_sync(0xffffffffu, v, src + 0u);
    c.mean[1] = __shfl_sync(0xffffffffu, v, src + 1u);
    c.corr[0] = __shfl_sync(0xffffffffu, v, src + 2u);
    c.corr[1] = __shfl_sync(0xffffffffu, v, src + 3u);
    c.corr[2] = __shfl_sync(0xffffffffu, v, src + 4u);
    if (have && r == 0u) {
      stats_finalize(c);
      const uint32_t k = key[i];
      const uint32_t p = padded_index(g, k);
      const uint2 w = occ[p >> 5];
      const uint32_t rank = w.y + __popc(w.x & ((1u << (p & 31u)) - 1u));
      if (rank < rec_cap) {
        write_records(c, g, rec + static_cast<size_t>(rank) * NDT2D_REC_DOUBLES,
          rec_fast + static_cast<size_t>(rank) * NDT2D_REC_DOUBLES,
          rec_vtx + static_cast<size_t>(rank) * NDT2D_REC_DOUBLES, n_valid + 1);
      }
    }
  }
}

// Parity dump: every occupied cell, dense, in the layout of ndt_2d::Cell.
__global__ void __launch_bounds__(128) segment_dump_kernel(
  GridDesc g, const uint32_t * __restrict__ key, size_t n,
  const uint32_t * __restrict__ seglen, const double * __restrict__ sx,
  const double * __restrict__ sy, double * __restrict__ out)
{
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) {return;}
  const uint32_t k = key[i];
  if (k >= g.n_cells) {return;}
  if (i > 0 && key[i - 1] == k) {return;}
  const uint32_t len = seglen[i];
  CellStats c;
  stats_walk(c, sx, sy, i, len);
  double * o = out + static_cast<size_t>(k) * 16;
  o[1] = c.n;
  o[2] = c.mean[0];
  o[3] = c.mean[1];
  o[8] = c.corr[0];
  o[9] = c.corr[1];
  o[10] = 0.0;  // correlation(1,0) is never written (ndt_model.cpp:55)
  o[11] = c.corr[2];
  if (len >= 3) {
    stats_finalize(c);
    o[0] = 1.0;
    o[4] = c.cov[0];
    o[5] = c.cov[1];
    o[6] = c.cov[1];
    o[7] = c.cov[2];
    o[12] = c.info[0];
    o[13] = c.info[1];
    o[14] = c.info[2];
    o[15] = c.info[3];
  }
}

// 32 occupancy bits starting at padded cell `pos` (may straddle two words).
__device__ __forceinline__ uint32_t occ_bits_at(
  const uint2 * __restrict__ occ, uint32_t n_words, uint64_t pos)
{
  const uint64_t w = pos >> 5;
  const uint32_t sh = static_cast<uint32_t>(pos & 31u);
  const uint32_t lo = w < n_words ? occ[w].x : 0u;
  const uint32_t hi = (w + 1) < n_words ? occ[w + 1].x : 0u;
  return __funnelshift_r(lo, hi, sh);
}

// K3c: dilated occupancy D[c] = OR of E over columns c .. c + dil_x and rows c, c + pitch.
// A region of search candidates whose first candidate puts a point in padded cell c can
// only touch those (dil_x + 1) x 2 cells (search_region.cu: a region is up to 32 columns
// wide -- at most 3 cell columns -- and at most one cell high), so one D bit rejects the
// whole region.  Past the last column the window wraps into the next row's border cell and
// first real cells: a conservative superset, the search probes the real cells afterwards.
__global__ void __launch_bounds__(256) dilate_kernel(
  GridDesc g, const uint2 * __restrict__ occ, uint32_t * __restrict__ occ_dilated)
{
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= g.n_words) {return;}
  const uint64_t pos = static_cast<uint64_t>(w) << 5;
  uint32_t d = 0;
  for (uint32_t i = 0; i <= g.dil_x; ++i) {
    d |= occ_bits_at(occ, g.n_words, pos + i) | occ_bits_at(occ, g.n_words, pos + g.pitch + i);
  }
  occ_dilated[w] = d;
}


// ------------------------------------------------------------------ small models
// A rolling-window model (config 1: 10 scans, 3,600 points; a loop-closure window:
// 1-2 scans) is far too small for the multi-launch pipeline above: ~15 dependent
// launches of a few microseconds each.  build_small_kernel does the whole build in ONE
// CTA: transform + key, a stable in-shared-memory radix sort of (key << 32 | point
// index), run lengths, occupancy bitmap + rank prefix + dilated bitmap, and the same
// 8-lanes-per-cell moment recurrences -- identical results (the sort is stable because
// the point index is part of the sort item; the recurrence walks the points in order).
constexpr uint32_t kSmallMaxPoints = 4096;
constexpr uint32_t kSmallMaxWords = 2048;   // padded cells <= 65,536
constexpr uint32_t kSmallThreads = 1024;
constexpr uint32_t kSmallMaxHeads = kSmallMaxPoints / 5 + 1;

struct SmallSmem
{
  unsigned long long items[kSmallMaxPoints];
  unsigned long long items_alt[kSmallMaxPoints];   // ping-pong buffer of the radix sort
  uint32_t cnt[32][256];                           // per-warp digit counters / offsets
  double wx[kSmallMaxPoints];
  double wy[kSmallMaxPoints];
  double rcp[kSmallMaxPoints + 1];                 // rcp[n] = RN(1 / n): div_by_count's reciprocals
  double one;                                      // 1.0: the second factor of the mean lanes
  uint32_t occw[kSmallMaxWords];
  uint32_t pref[kSmallMaxWords];
  uint2 heads[kSmallMaxHeads];
  uint32_t n_heads;
  uint32_t total;
};

__device__ __forceinline__ uint32_t small_bits_at(const uint32_t * occw, uint32_t n_words, uint32_t pos)
{
  const uint32_t w = pos >> 5, sh = pos & 31u;
  const uint32_t lo = w < n_words ? occw[w] : 0u;
  const uint32_t hi = (w + 1) < n_words ? occw[w + 1] : 0u;
  return __funnelshift_r(lo, hi, sh);
}

__device__ __forceinline__ void build_small_body(const BuildEntry & e, SmallSmem & sm)
{
  const GridDesc g = e.g;
  const double4 * __restrict__ scan_tf = e.scan_tf;
  const uint64_t * __restrict__ offsets = e.offsets;
  const uint32_t n_scans = e.n_scans, n_points = e.n_points, rec_cap = e.rec_cap;
  const double2 * __restrict__ pts = e.pts;
  uint32_t * __restrict__ key_out = e.key_out;    // the five dump outputs may be null
  uint32_t * __restrict__ val_out = e.val_out;
  uint32_t * __restrict__ seglen = e.seglen;
  double * __restrict__ sx = e.sx;
  double * __restrict__ sy = e.sy;
  uint2 * __restrict__ occ = e.occ;
  uint32_t * __restrict__ occ_dilated = e.occd;
  double * __restrict__ rec = e.rec;
  double * __restrict__ rec_fast = e.rec_fast;
  double * __restrict__ rec_vtx = e.rec_vtx;
  uint32_t * __restrict__ n_valid = e.n_valid;
  const uint32_t tid = threadIdx.x;
  uint32_t n2 = 128;
  while (n2 < n_points) {n2 <<= 1;}

  for (uint32_t w = tid; w < g.n_words; w += kSmallThreads) {sm.occw[w] = 0u;}
  for (uint32_t k = tid + 1u; k <= n_points; k += kSmallThreads) {sm.rcp[k] = __drcp_rn(static_cast<double>(k));}
  if (tid == 0) {
    sm.n_heads = 0u;
    sm.one = 1.0;
    n_valid[1] = 0u;   // stiff-cell count (write_records)
  }
  // ---- K1: transform + key (NDT::addScan, ndt_model.cpp:132-152)
  for (uint32_t p = tid; p < n2; p += kSmallThreads) {
    if (p < n_points) {
      uint32_t lo = 0, hi = n_scans;          // scan s with offsets[s] <= p < offsets[s + 1]
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= p) {lo = mid;} else {hi = mid;}
      }
      const double4 tf = scan_tf[lo];
      const double2 pt = pts[p];
      const double X = __dadd_rn(tf.x, __dsub_rn(__dmul_rn(pt.x, tf.z), __dmul_rn(pt.y, tf.w)));
      const double Y = __dadd_rn(tf.y, __dadd_rn(__dmul_rn(pt.x, tf.w), __dmul_rn(pt.y, tf.z)));
      sm.wx[p] = X;
      sm.wy[p] = Y;
      sm.items[p] = (static_cast<unsigned long long>(cell_key(g, X, Y)) << 32) | p;
    } else {
      sm.items[p] = ~0ull;
    }
  }
  __syncthreads();
  // ---- K2: stable LSD radix sort on the cell key, 8-bit digits, entirely in shared memory
  // (the same warp-ordered ranking as radix_scatter_kernel: warp w owns the 128 consecutive
  // items [128 w, 128 w + 128) and walks them in order, so equal keys keep their order).
  {
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool warp_active = warp * 128u < n2;
    const int bits = 32 - __clz(g.n_cells | 1u);           // keys are <= n_cells (= "outside")
    unsigned long long * src = sm.items;
    unsigned long long * dst = sm.items_alt;
    for (int shift = 0; shift < bits; shift += 8) {
      for (uint32_t k = lane; k < 256u; k += 32u) {sm.cnt[warp][k] = 0u;}
      __syncwarp();
      unsigned long long item[4];
      uint32_t local[4], dig[4];
      if (warp_active) {
#pragma unroll
        for (int st = 0; st < 4; ++st) {
          item[st] = src[warp * 128u + st * 32u + lane];
          // padding items (~0) get digit 255 in every pass and stay at the end
          const uint32_t d = static_cast<uint32_t>(item[st] >> (32 + shift)) & 255u;
          const uint32_t peers = __match_any_sync(0xffffffffu, d);
          const uint32_t r = __popc(peers & lt_mask);
          const uint32_t base = sm.cnt[warp][d];
          __syncwarp();
          if (r == 0u) {sm.cnt[warp][d] = base + __popc(peers);}
          __syncwarp();
          local[st] = base + r;
          dig[st] = d;
        }
      }
      __syncthreads();
      {
        // exclusive scan over (digit, warp) in digit-major order: entry e = d * 32 + w
        uint32_t v[8], tsum = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t e = tid * 8u + q;
          v[q] = sm.cnt[e & 31u][e >> 5];
          tsum += v[q];
        }
        uint32_t run = block_exclusive_scan_1024(tsum, &sm.total);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t e = tid * 8u + q;
          sm.cnt[e & 31u][e >> 5] = run;
          run += v[q];
        }
      }
      __syncthreads();
      if (warp_active) {
#pragma unroll
        for (int st = 0; st < 4; ++st) {dst[sm.cnt[warp][dig[st]] + local[st]] = item[st];}
      }
      __syncthreads();
      unsigned long long * t = src;
      src = dst;
      dst = t;
    }
    if (src != sm.items) {
      for (uint32_t i = tid; i < n2; i += kSmallThreads) {sm.items[i] = src[i];}
      __syncthreads();
    }
  }
  // ---- K3a: run lengths, occupancy bits, head list; sorted copies for the parity dumps
  for (uint32_t i = tid; i < n_points; i += kSmallThreads) {
    const unsigned long long it = sm.items[i];
    const uint32_t k = static_cast<uint32_t>(it >> 32), p = static_cast<uint32_t>(it);
    if (key_out) {
      key_out[i] = k;
      val_out[i] = p;
      sx[i] = sm.wx[p];
      sy[i] = sm.wy[p];
    }
    if (k >= g.n_cells) {continue;}
    if (i > 0 && static_cast<uint32_t>(sm.items[i - 1] >> 32) == k) {continue;}
    uint32_t lo = i, step = 1, hi = n_points;
    while (lo + step < n_points) {
      if (static_cast<uint32_t>(sm.items[lo + step] >> 32) == k) {
        lo += step;
        step <<= 1;
      } else {
        hi = lo + step;
        break;
      }
    }
    while (hi - lo > 1) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (static_cast<uint32_t>(sm.items[mid] >> 32) == k) {lo = mid;} else {hi = mid;}
    }
    const uint32_t len = hi - i;
    if (seglen) {seglen[i] = len;}
    if (len >= 5) {
      const uint32_t pi = padded_index(g, k);
      atomicOr(&sm.occw[pi >> 5], 1u << (pi & 31u));
      sm.heads[atomicAdd(&sm.n_heads, 1u)] = make_uint2(i, len);
    }
  }
  __syncthreads();
  // ---- rank prefix over the occupancy words (<= 2 words per thread), dilated bitmap
  {
    const uint32_t w0 = 2u * tid, w1 = w0 + 1u;
    const uint32_t c0 = w0 < g.n_words ? __popc(sm.occw[w0]) : 0u;
    const uint32_t c1 = w1 < g.n_words ? __popc(sm.occw[w1]) : 0u;
    // (the scan helper has one thread store the block total; it syncs before returning)
    const uint32_t ex = block_exclusive_scan_1024(c0 + c1, &sm.total);
    if (w0 < g.n_words) {sm.pref[w0] = ex;}
    if (w1 < g.n_words) {sm.pref[w1] = ex + c0;}
  }
  __syncthreads();
  if (tid == 0) {*n_valid = sm.total;}
  for (uint32_t w = tid; w < g.n_words; w += kSmallThreads) {
    occ[w] = make_uint2(sm.occw[w], sm.pref[w]);
    const uint32_t pos = w << 5;
    uint32_t d = 0;
    for (uint32_t i = 0; i <= g.dil_x; ++i) {
      d |= small_bits_at(sm.occw, g.n_words, pos + i) | small_bits_at(sm.occw, g.n_words, pos + g.pitch + i);
    }
    occ_dilated[w] = d;
  }
  // ---- K3b: moment recurrences, 5 lanes per listed cell, 6 cells per warp (as
  // segment_moments_kernel): 192 cells per round of the CTA, so a rolling window (config 1: 141
  // cells) is one round.  One SM runs all of it, so the phase is bound by the instructions per
  // recurrence step: every lane forms its term as ONE product a[p] * b[p * stride] -- (x, 1),
  // (y, 1), (x, x), (x, y), (y, y) chosen by two lane-invariant pointers, x * 1.0 being exact --
  // instead of selecting among five candidates, the reciprocal of the point count comes from
  // a table (div_by_count), and the loads of four points are issued ahead of the steps that
  // consume them.
  {
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const uint32_t sub = lane / 5u, r = lane - sub * 5u;      // lanes 30, 31: sub == 6, no cell
    const uint32_t src = (sub < kMomentCellsPerWarp ? sub : kMomentCellsPerWarp - 1u) * 5u;
    const double * const pa = (r == 1u || r == 4u) ? sm.wy : sm.wx;
    const double * const pb = r < 2u ? &sm.one : (r == 2u ? sm.wx : sm.wy);
    const uint32_t sb = r < 2u ? 0u : 1u;
    const uint32_t n_heads = sm.n_heads;
    constexpr uint32_t kPerRound = (kSmallThreads / 32u) * kMomentCellsPerWarp;
    for (uint32_t base = 0; base < n_heads; base += kPerRound) {
      const uint32_t cell = base + warp * kMomentCellsPerWarp + sub;
      const bool have = sub < kMomentCellsPerWarp && cell < n_heads;
      const uint2 h = have ? sm.heads[cell] : make_uint2(0u, 0u);
      double v = 0.0, n = 0.0;
      if (have) {
        const unsigned long long * it = sm.items + h.x;
        const double * rc = sm.rcp + 1;
        uint32_t j = 0;
        for (; j + 4 <= h.y; j += 4) {
          double t[4], q[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t p = static_cast<uint32_t>(it[j + u]);
            t[u] = __dmul_rn(pa[p], pb[p * sb]);
            q[u] = rc[j + u];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double n1 = __dadd_rn(n, 1.0);
            v = div_by_count(__dadd_rn(__dmul_rn(v, n), t[u]), n1, q[u]);
            n = n1;
          }
        }
        for (; j < h.y; ++j) {
          const uint32_t p = static_cast<uint32_t>(it[j]);
          const double term = __dmul_rn(pa[p], pb[p * sb]);
          const double n1 = __dadd_rn(n, 1.0);
          v = div_by_count(__dadd_rn(__dmul_rn(v, n), term), n1, rc[j]);
          n = n1;
        }
      }
      CellStats c;
      c.n = static_cast<double>(h.y);
      c.mean[0] = __shfl_sync(0xffffffffu, v, src + 0u);
      c.mean[1] = __shfl_sync(0xffffffffu, v, src + 1u);
      c.corr[0] = __shfl_sync(0xffffffffu, v, src + 2u);
      c.corr[1] = __shfl_sync(0xffffffffu, v, src + 3u);
      c.corr[2] = __shfl_sync(0xffffffffu, v, src + 4u);
      if (have && r == 0u) {
        stats_finalize(c);
        const uint32_t k = static_cast<uint32_t>(sm.items[h.x] >> 32);
        const uint32_t pi = padded_index(g, k);
        const uint32_t bits = sm.occw[pi >> 5];
        const uint32_t rank = sm.pref[pi >> 5] + __popc(bits & ((1u << (pi & 31u)) - 1u));
        if (rank < rec_cap) {
          write_records(c, g, rec + static_cast<size_t>(rank) * NDT2D_REC_DOUBLES,
            rec_fast + static_cast<size_t>(rank) * NDT2D_REC_DOUBLES,
            rec_vtx + static_cast<size_t>(rank) * NDT2D_REC_DOUBLES, n_valid + 1);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kSmallThreads, 1) build_small_kernel(BuildEntry e)
{
  extern __shared__ __align__(16) unsigned char small_raw[];
  build_small_body(e, *reinterpret_cast<SmallSmem *>(small_raw));
}

// One CTA per model of a batch (the loop-closure windows of match_scan_batch).
__global__ void __launch_bounds__(kSmallThreads, 1) build_small_batch_kernel(
  const BuildEntry * __restrict__ entries)
{
  extern __shared__ __align__(16) unsigned char small_raw[];
  __shared__ BuildEntry e;
  {
    const uint32_t * src = reinterpret_cast<const uint32_t *>(entries + blockIdx.x);
    uint32_t * dst = reinterpret_cast<uint32_t *>(&e);
    for (uint32_t k = threadIdx.x; k < sizeof(BuildEntry) / 4; k += blockDim.x) {dst[k] = src[k];}
  }
  __syncthreads();
  build_small_body(e, *reinterpret_cast<SmallSmem *>(small_raw));
}

// opt in to the large dynamic shared memory of the single-CTA build, once per device
int configure_small_kernels()
{
  static bool configured[64] = {false};
  int dev = 0;
  NDT2D_CUDA_TRY(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    NDT2D_CUDA_TRY(cudaFuncSetAttribute(build_small_kernel,
      cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SmallSmem))));
    NDT2D_CUDA_TRY(cudaFuncSetAttribute(build_small_batch_kernel,
      cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SmallSmem))));
    configured[dev & 63] = true;
  }
  return NDT2D_OK;
}

int bits_needed(uint32_t max_value)
{
  int b = 1;
  while (b < 32 && (max_value >> b) != 0u) {++b;}
  return b;
}

}  // namespace

size_t ndt2d_sort_scratch_bytes(size_t n_points)
{
  const size_t nblk = (n_points + kSortTile - 1) / kSortTile;
  return (static_cast<size_t>(kMaxPasses) * 256 + 64 + static_cast<size_t>(kMaxPasses) * nblk * 256) *
         sizeof(uint32_t);
}

bool ndt2d_build_is_small(const GridDesc & g, size_t n_points)
{
  return n_points > 0 && n_points <= kSmallMaxPoints && g.n_words <= kSmallMaxWords;
}

int ndt2d_launch_build_small_batch(const BuildEntry * d_entries, uint32_t n, cudaStream_t stream,
  Counters * ctr)
{
  if (n == 0) {return NDT2D_OK;}
  const int rc0 = configure_small_kernels();
  if (rc0 != NDT2D_OK) {return rc0;}
  build_small_batch_kernel<<<n, kSmallThreads, sizeof(SmallSmem), stream>>>(d_entries);
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}

int ndt2d_launch_exclusive_scan(uint32_t * d_data, size_t n, uint32_t * d_tmp,
  cudaStream_t stream, Counters * ctr)
{
  return launch_scan<0>(d_data, n, d_tmp, stream, ctr);
}

int ndt2d_launch_build(
  const GridDesc & g, const double4 * d_scan_tf, const uint64_t * d_offsets, size_t n_scans,
  const double2 * d_pts, size_t n_points, BuildScratch & s, uint2 * d_occ, uint32_t * d_occ_dilated,
  double * d_rec, double * d_rec_fast, double * d_rec_vtx, uint32_t rec_cap, uint32_t * d_n_valid,
  cudaStream_t stream, Counters * ctr, int * sorted_buf)
{
  if (n_points > 0 && n_points <= kSmallMaxPoints && g.n_words <= kSmallMaxWords) {
    // small model: the whole build in one CTA / one launch
    const int rc0 = configure_small_kernels();
    if (rc0 != NDT2D_OK) {return rc0;}
    BuildEntry e{};
    e.g = g;
    e.scan_tf = d_scan_tf;
    e.offsets = d_offsets;
    e.n_scans = static_cast<uint32_t>(n_scans);
    e.pts = d_pts;
    e.n_points = static_cast<uint32_t>(n_points);
    e.occ = d_occ;
    e.occd = d_occ_dilated;
    e.rec = d_rec;
    e.rec_fast = d_rec_fast;
    e.rec_vtx = d_rec_vtx;
    e.rec_cap = rec_cap;
    e.n_valid = d_n_valid;
    e.key_out = s.key[0];
    e.val_out = s.val[0];
    e.seglen = s.seglen;
    e.sx = s.sx;
    e.sy = s.sy;
    build_small_kernel<<<1, kSmallThreads, sizeof(SmallSmem), stream>>>(e);
    NDT2D_LAUNCH_CHECK(ctr);
    *sorted_buf = 0;
    return NDT2D_OK;
  }
  NDT2D_CUDA_TRY(cudaMemsetAsync(d_occ, 0, (static_cast<size_t>(g.n_words) + 4) * sizeof(uint2), stream));
  NDT2D_CUDA_TRY(cudaMemsetAsync(d_occ_dilated, 0, (static_cast<size_t>(g.n_words) + 4) * sizeof(uint32_t), stream));
  NDT2D_CUDA_TRY(cudaMemsetAsync(d_n_valid, 0, 2 * sizeof(uint32_t), stream));   // n_valid, n_stiff
  NDT2D_CUDA_TRY(cudaMemsetAsync(s.n_heads, 0, sizeof(uint32_t), stream));
  int cur = 0;
  if (n_points > 0) {
    if (n_points >= (size_t(1) << 30)) {return NDT2D_ERR_SIZE;}   // 30-bit counts in the look-back words
    const uint32_t nblk = static_cast<uint32_t>((n_points + kSortTile - 1) / kSortTile);
    DigitPlan dp;
    const int bits = bits_needed(g.n_cells);
    dp.passes = (bits + 7) / 8;
    dp.width = (bits + dp.passes - 1) / dp.passes;
    dp.mask = (1u << dp.width) - 1u;
    // sort scratch (s.hist): [kMaxPasses][256] digit totals | tickets | [passes][nblk][256] status
    uint32_t * ghist = s.hist;
    uint32_t * tickets = s.hist + kMaxPasses * 256;
    uint32_t * status = tickets + 64;
    NDT2D_CUDA_TRY(cudaMemsetAsync(s.hist, 0, ndt2d_sort_scratch_bytes(n_points), stream));
    // a few blocks per SM, each looping over scans (one flush of its digit counters per block)
    const uint32_t grid = static_cast<uint32_t>(n_scans < 148u * 16u ? n_scans : 148u * 16u);
    transform_key_kernel<<<grid, 128, 0, stream>>>(
      g, d_scan_tf, d_offsets, static_cast<uint32_t>(n_scans), d_pts, s.wx, s.wy, s.key[0],
      s.val[0], dp, ghist);
    NDT2D_LAUNCH_CHECK(ctr);
    static bool configured[64] = {false};
    int dev = 0;
    NDT2D_CUDA_TRY(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
      NDT2D_CUDA_TRY(cudaFuncSetAttribute(radix_onesweep_kernel<false>,
        cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(OnesweepSmem))));
      NDT2D_CUDA_TRY(cudaFuncSetAttribute(radix_onesweep_kernel<true>,
        cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(OnesweepSmem))));
      configured[dev & 63] = true;
    }
    for (int q = 0; q < dp.passes; ++q) {
      auto kernel = q + 1 == dp.passes ? radix_onesweep_kernel<true> : radix_onesweep_kernel<false>;
      kernel<<<nblk, kSortThreads, sizeof(OnesweepSmem), stream>>>(
        s.key[cur], s.val[cur], static_cast<uint32_t>(n_points), q * dp.width, dp.mask,
        ghist + q * 256, status + static_cast<size_t>(q) * nblk * 256u, tickets + q, s.key[cur ^ 1],
        s.val[cur ^ 1], s.wx, s.wy, s.sx, s.sy);
      NDT2D_LAUNCH_CHECK(ctr);
      cur ^= 1;
    }
    const uint32_t nb = static_cast<uint32_t>((n_points + 255) / 256);
    segment_count_kernel<<<nb, 256, 0, stream>>>(g, s.key[cur], n_points, s.seglen, d_occ,
      s.heads, s.n_heads);
    NDT2D_LAUNCH_CHECK(ctr);
  }
  {
    const int rc = launch_scan<1>(d_occ, g.n_words, s.scan_tmp, stream, ctr);
    if (rc != NDT2D_OK) {return rc;}
    dilate_kernel<<<(g.n_words + 255) / 256, 256, 0, stream>>>(g, d_occ, d_occ_dilated);
    NDT2D_LAUNCH_CHECK(ctr);
  }
  if (n_points > 0) {
    // one 5-lane group per listed cell, warps stride over the list (at most rec_cap long)
    const uint32_t want = (rec_cap + 8u * kMomentCellsPerWarp - 1u) / (8u * kMomentCellsPerWarp);
    const uint32_t nb = want < 148u * 8u ? (want ? want : 1u) : 148u * 8u;
    if (!s.rcp_ready) {
      rcp_table_kernel<<<(NDT2D_RCP_TABLE + 256) / 256, 256, 0, stream>>>(s.rcp);
      NDT2D_LAUNCH_CHECK(ctr);
      s.rcp_ready = true;
    }
    segment_moments_kernel<<<nb, 256, 0, stream>>>(
      g, s.key[cur], s.heads, s.n_heads, s.sx, s.sy, s.rcp, d_occ, d_rec, d_rec_fast, d_rec_vtx, rec_cap,
      d_n_valid);
    NDT2D_LAUNCH_CHECK(ctr);
  }
  *sorted_buf = cur;
  return NDT2D_OK;
}

int ndt2d_launch_dump_cells(
  const GridDesc & g, const BuildScratch & s, int sorted_buf, size_t n_points, double * d_out,
  cudaStream_t stream, Counters * ctr)
{
  NDT2D_CUDA_TRY(
    cudaMemsetAsync(d_out, 0, static_cast<size_t>(g.n_cells) * 16 * sizeof(double), stream));
  if (n_points > 0) {
    const uint32_t nb = static_cast<uint32_t>((n_points + 127) / 128);
    segment_dump_kernel<<<nb, 128, 0, stream>>>(
      g, s.key[sorted_buf], n_points, s.seglen, s.sx, s.sy, d_out);
    NDT2D_LAUNCH_CHECK(ctr);
  }
  return NDT2D_OK;
}
