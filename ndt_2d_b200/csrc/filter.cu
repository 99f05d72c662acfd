// filter.cu -- particle-filter measurement update pieces on the device
// (kernels K6, K7): weight normalisation + statistics, multinomial draw with
// the KLD stopping rule, Gaussian initialisation and the odometry motion model.
//
// Replaces ParticleFilter::updateStatistics / resample / init / update
// (particle_filter.cpp:53-76, 91-137, 163-218), KDTree's leaf counting
// (kd_tree.hpp:97-189) and MotionModel::sample (motion_model.cpp:45-83).
// The particle set is small (hundreds to thousands), so each step is ONE
// single-block launch working out of L1/L2: latency-bound by design.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ndt2d_internal.h"

namespace
{

constexpr int kThreads = 1024;
constexpr uint32_t kEmpty = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;
}

// Sum of K values per thread over the whole block; result valid in all threads.
template<int K>
__device__ __forceinline__ void block_sum(double (&v)[K])
{
  __shared__ double red[32][K];
  __shared__ double tot[K];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {v[k] = warp_sum(v[k]);}
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {red[warp][k] = v[k];}
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double t = (lane < (blockDim.x >> 5)) ? red[lane][k] : 0.0;
      t = warp_sum(t);
      if (lane == 0) {tot[k] = t;}
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {v[k] = tot[k];}
  __syncthreads();
}

// angles::normalize_angle (ROS 2 `angles`): fmod(a + pi, 2pi), then shift.
__device__ __forceinline__ double normalize_angle(double a)
{
  const double pi = 3.14159265358979323846;
  const double r = fmod(a + pi, 2.0 * pi);
  return (r <= 0.0) ? r + pi : r - pi;
}

// Same counter-based uniform stream as ndt2d_synth_uniform (synth.cpp).
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double uniform_at(uint64_t seed, uint64_t i)
{
  const uint64_t z = splitmix64(seed * 0xD1342543DE82EF95ull + i * 0x9E3779B97F4A7C15ull);
  return static_cast<double>(z >> 11) * (1.0 / 9007199254740992.0);
}
// N(0,1) as float, the type std::normal_distribution<float> produces.
__device__ __forceinline__ float normal_at(uint64_t seed, uint64_t i)
{
  const double u1 = uniform_at(seed, 2 * i), u2 = uniform_at(seed, 2 * i + 1);
  return static_cast<float>(sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586476925 * u2));
}

// ------------------------------------------------------------------ pose tf
__global__ void pose_tf_kernel(const double * __restrict__ particles, uint32_t n,
  double4 * __restrict__ tf)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {return;}
  const double th = particles[3 * i + 2];
  double s, c;
  sincos(th, &s, &c);
  tf[i] = make_double4(particles[3 * i], particles[3 * i + 1], c, s);
}

// ------------------------------------------------------------------ K6
// ParticleFilter::updateStatistics (particle_filter.cpp:163-218).
// stats[0..2] = mean_, stats[3..11] = cov_ (row-major, persistent: (2,2) is
// accumulated with '+=' and (0,2),(1,2),(2,0),(2,1) are never written, :216).
__global__ void __launch_bounds__(kThreads) filter_stats_kernel(FilterView f, uint32_t n)
{
  double sw[1] = {0.0};
  for (uint32_t i = threadIdx.x; i < n; i += kThreads) {sw[0] += f.weights[i];}
  block_sum<1>(sw);
  const double sum_weight = sw[0];

  double a[7] = {0, 0, 0, 0, 0, 0, 0};  // wx wy wcos wsin wxx wxy wyy
  for (uint32_t i = threadIdx.x; i < n; i += kThreads) {
    const double w = __ddiv_rn(f.weights[i], sum_weight);
    f.weights[i] = w;
    const double x = f.particles[3 * i], y = f.particles[3 * i + 1], th = f.particles[3 * i + 2];
    double s, c;
    sincos(th, &s, &c);
    a[0] += w * x;
    a[1] += w * y;
    a[2] += w * c;
    a[3] += w * s;
    a[4] += w * x * x;
    a[5] += w * x * y;
    a[6] += w * y * y;
  }
  block_sum<7>(a);
  const double mean_t = atan2(a[3], a[2]);

  double d2[1] = {0.0};
  for (uint32_t i = threadIdx.x; i < n; i += kThreads) {
    // shortest_angular_distance(theta_i, mean_theta) = normalize(mean - theta_i)
    const double d = normalize_angle(mean_t - f.particles[3 * i + 2]);
    d2[0] += f.weights[i] * d * d;
  }
  block_sum<1>(d2);

  if (threadIdx.x == 0) {
    double * mean = f.stats, * cov = f.stats + 3;
    mean[0] = a[0];
    mean[1] = a[1];
    mean[2] = mean_t;
    cov[0] = a[4] - a[0] * a[0];
    cov[1] = a[5] - a[0] * a[1];
    cov[3] = cov[1];
    cov[4] = a[6] - a[1] * a[1];
    cov[8] += d2[0];
  }
}

// ------------------------------------------------------------------ K7
// KD bin key: static_cast<int>(coord / size), truncation toward zero
// (kd_tree.hpp:99-102), bin sizes (0.5, 0.5, 0.2671) (particle_filter.cpp:44).
struct Key3 {int k[3];};
__device__ __forceinline__ Key3 kd_key(const double * __restrict__ particles, uint32_t p)
{
  Key3 r;
  r.k[0] = __double2int_rz(__ddiv_rn(particles[3 * p + 0], 0.5));
  r.k[1] = __double2int_rz(__ddiv_rn(particles[3 * p + 1], 0.5));
  r.k[2] = __double2int_rz(__ddiv_rn(particles[3 * p + 2], 0.2671));
  return r;
}
__device__ __forceinline__ bool key_eq(const Key3 & a, const Key3 & b)
{
  return a.k[0] == b.k[0] && a.k[1] == b.k[1] && a.k[2] == b.k[2];
}
__device__ __forceinline__ uint32_t key_hash(const Key3 & k, uint32_t mask)
{
  uint64_t h = static_cast<uint32_t>(k.k[0]) * 0x9E3779B97F4A7C15ull;
  h ^= static_cast<uint32_t>(k.k[1]) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
  h ^= static_cast<uint32_t>(k.k[2]) * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
  return static_cast<uint32_t>(h ^ (h >> 29)) & mask;
}

// Inclusive block scan (double) of one value per thread with a running carry.
__device__ __forceinline__ double block_inclusive_scan_d(double v, double * carry_s)
{
  __shared__ double wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) {incl += t;}
  }
  if (lane == 31) {wsum[warp] = incl;}
  __syncthreads();
  if (warp == 0) {
    double w = wsum[lane];
    double wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) {wi += t;}
    }
    wsum[lane] = wi - w;
  }
  __syncthreads();
  const double r = *carry_s + wsum[warp] + incl;
  __syncthreads();
  if (threadIdx.x == kThreads - 1) {*carry_s = r;}
  __syncthreads();
  return r;
}
__device__ __forceinline__ uint32_t block_inclusive_scan_u(uint32_t v, uint32_t * carry_s)
{
  __shared__ uint32_t wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) {incl += t;}
  }
  if (lane == 31) {wsum[warp] = incl;}
  __syncthreads();
  if (warp == 0) {
    uint32_t w = wsum[lane];
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) {wi += t;}
    }
    wsum[lane] = wi - w;
  }
  __syncthreads();
  const uint32_t r = *carry_s + wsum[warp] + incl;
  __syncthreads();
  if (threadIdx.x == kThreads - 1) {*carry_s = r;}
  __syncthreads();
  return r;
}

// ParticleFilter::resample (particle_filter.cpp:91-137) without the trailing
// updateStatistics.  Every draw is independent given its uniform variate, so
// all max_particles draws are made at once and the sequential stopping rule
// is evaluated as a prefix computation:
//   cdf      : std::discrete_distribution's normalised partial sums, last = 1
//   draw i   : lower_bound(cdf, u_i)
//   canon[p] : representative particle of p's KD bin (lock-free hash claim)
//   first[c] : earliest draw that landed in bin c  -> "new leaf" flag per draw
//   k(i)     : inclusive scan of the flags == KDTree::getLeafCount() after draw i
//   stop     : first m with m >= max(min_particles, Mx(k(m-1))) or m >= max_particles
__global__ void __launch_bounds__(kThreads) filter_resample_kernel(
  FilterView f, uint32_t n, uint32_t min_particles, uint32_t max_particles, double kld_err,
  double kld_z, const double * __restrict__ uniforms, uint64_t seed, double * __restrict__ cdf,
  double * __restrict__ new_particles, double * __restrict__ new_weights,
  uint32_t * __restrict__ draws, uint32_t * __restrict__ first, uint32_t * __restrict__ canon,
  uint32_t * __restrict__ table, uint32_t table_size, uint32_t * __restrict__ new_n)
{
  __shared__ double carry_d;
  __shared__ uint32_t carry_u;
  __shared__ uint32_t stop_at;
  const uint32_t tid = threadIdx.x;

  // --- cdf
  double sw[1] = {0.0};
  for (uint32_t i = tid; i < n; i += kThreads) {sw[0] += f.weights[i];}
  block_sum<1>(sw);
  const double sum = sw[0];
  if (tid == 0) {
    carry_d = 0.0;
    carry_u = 0;
    stop_at = max_particles;
  }
  __syncthreads();
  for (uint32_t base = 0; base < n; base += kThreads) {
    const uint32_t i = base + tid;
    const double p = (i < n) ? __ddiv_rn(f.weights[i], sum) : 0.0;
    const double c = block_inclusive_scan_d(p, &carry_d);
    if (i < n) {cdf[i] = (i == n - 1) ? 1.0 : c;}
  }
  // --- bins of the particles
  for (uint32_t i = tid; i < table_size; i += kThreads) {table[i] = kEmpty;}
  for (uint32_t i = tid; i < n; i += kThreads) {first[i] = kEmpty;}
  __syncthreads();
  for (uint32_t p = tid; p < n; p += kThreads) {
    const Key3 key = kd_key(f.particles, p);
    uint32_t h = key_hash(key, table_size - 1);
    for (;; ) {
      const uint32_t prev = atomicCAS(&table[h], kEmpty, p);
      if (prev == kEmpty) {
        canon[p] = p;
        break;
      }
      if (key_eq(kd_key(f.particles, prev), key)) {
        canon[p] = prev;
        break;
      }
      h = (h + 1) & (table_size - 1);
    }
  }
  __syncthreads();
  // --- draws
  for (uint32_t i = tid; i < max_particles; i += kThreads) {
    uint32_t idx = 0;
    if (n >= 2) {
      const double u = uniforms ? uniforms[i] : uniform_at(seed, i);
      uint32_t lo = 0, hi = n;
      while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (cdf[mid] < u) {lo = mid + 1;} else {hi = mid;}
      }
      idx = lo < n ? lo : n - 1;
    }
    draws[i] = idx;
    atomicMin(&first[canon[idx]], i);
  }
  __syncthreads();
  // --- leaf counts, stopping rule
  for (uint32_t base = 0; base < max_particles; base += kThreads) {
    const uint32_t i = base + tid;
    const uint32_t flag = (i < max_particles && first[canon[draws[i]]] == i) ? 1u : 0u;
    const uint32_t k = block_inclusive_scan_u(flag, &carry_u);
    if (i < max_particles) {
      unsigned long long Mx = max_particles;
      if (k > 1) {
        const double km1 = static_cast<double>(k - 1);
        const double a = __ddiv_rn(km1, __dmul_rn(2.0, kld_err));
        const double b = __ddiv_rn(2.0, __dmul_rn(9.0, km1));
        const double c = __dadd_rn(__dsub_rn(1.0, b), __dmul_rn(__dsqrt_rn(b), kld_z));
        Mx = __double2ull_rz(__dmul_rn(__dmul_rn(__dmul_rn(a, c), c), c));
      }
      const unsigned long long need = Mx > min_particles ? Mx : min_particles;
      const uint32_t m = i + 1;
      if (m >= need || m >= max_particles) {atomicMin(&stop_at, m);}
    }
  }
  __syncthreads();
  const uint32_t N = stop_at;
  for (uint32_t i = tid; i < N; i += kThreads) {
    const uint32_t p = draws[i];
    new_particles[3 * i + 0] = f.particles[3 * p + 0];
    new_particles[3 * i + 1] = f.particles[3 * p + 1];
    new_particles[3 * i + 2] = f.particles[3 * p + 2];
    new_weights[i] = f.weights[p];  // the OLD weight (particle_filter.cpp:114)
  }
  if (tid == 0) {*new_n = N;}
}

// ParticleFilter::init (particle_filter.cpp:53-69); float normal variates.
__global__ void filter_init_kernel(FilterView f, uint32_t n, double x, double y, double th,
  double sx, double sy, double sth, uint64_t seed)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {return;}
  const float vx = static_cast<float>(x) + static_cast<float>(sx) * normal_at(seed, 3ull * i);
  const float vy = static_cast<float>(y) + static_cast<float>(sy) * normal_at(seed, 3ull * i + 1);
  const float vt = static_cast<float>(th) + static_cast<float>(sth) * normal_at(seed, 3ull * i + 2);
  f.particles[3 * i + 0] = vx;
  f.particles[3 * i + 1] = vy;
  f.particles[3 * i + 2] = normalize_angle(static_cast<double>(vt));
  f.weights[i] = 1.0 / static_cast<double>(n);
}

// MotionModel::sample's per-pose loop (motion_model.cpp:73-82); the scalar
// decomposition (:48-66) is done on the host.
__global__ void filter_motion_kernel(FilterView f, uint32_t n, double rot1, double trans,
  double rot2, double s_rot1, double s_trans, double s_rot2, uint64_t seed)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) {return;}
  const float r1 = static_cast<float>(rot1) + static_cast<float>(s_rot1) * normal_at(seed, 3ull * i);
  const float t = static_cast<float>(trans) + static_cast<float>(s_trans) * normal_at(seed, 3ull * i + 1);
  const float r2 = static_cast<float>(rot2) + static_cast<float>(s_rot2) * normal_at(seed, 3ull * i + 2);
  const double th = f.particles[3 * i + 2];
  double s, c;
  sincos(th + r1, &s, &c);
  f.particles[3 * i + 0] += t * c;
  f.particles[3 * i + 1] += t * s;
  f.particles[3 * i + 2] = normalize_angle(th + r1 + r2);
}

}  // namespace

int ndt2d_launch_pose_tf(const double * d_particles, uint32_t n, double4 * d_pose_tf,
  cudaStream_t stream, Counters * ctr)
{
  if (n == 0) {return NDT2D_OK;}
  pose_tf_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_particles, n, d_pose_tf);
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}

int ndt2d_launch_filter_stats(FilterView f, uint32_t n, cudaStream_t stream, Counters * ctr)
{
  filter_stats_kernel<<<1, kThreads, 0, stream>>>(f, n);
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}

int ndt2d_launch_filter_resample(
  FilterView f, uint32_t n, uint32_t min_particles, uint32_t max_particles, double kld_err,
  double kld_z, const double * d_uniforms, uint64_t seed, double * d_cdf, double * d_new_particles,
  double * d_new_weights, uint32_t * d_draws, uint32_t * d_first, uint32_t * d_canon,
  uint32_t * d_table, uint32_t table_size, uint32_t * d_new_n, cudaStream_t stream,
  Counters * ctr)
{
  filter_resample_kernel<<<1, kThreads, 0, stream>>>(
    f, n, min_particles, max_particles, kld_err, kld_z, d_uniforms, seed, d_cdf, d_new_particles,
    d_new_weights, d_draws, d_first, d_canon, d_table, table_size, d_new_n);
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}

int ndt2d_launch_filter_init(FilterView f, uint32_t n, double x, double y, double th, double sx,
  double sy, double sth, uint64_t seed, cudaStream_t stream, Counters * ctr)
{
  if (n == 0) {return NDT2D_OK;}
  filter_init_kernel<<<(n + 255) / 256, 256, 0, stream>>>(f, n, x, y, th, sx, sy, sth, seed);
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}

int ndt2d_launch_filter_motion(FilterView f, uint32_t n, double rot1, double trans, double rot2,
  double s_rot1, double s_trans, double s_rot2, uint64_t seed, cudaStream_t stream,
  Counters * ctr)
{
  if (n == 0) {return NDT2D_OK;}
  filter_motion_kernel<<<(n + 255) / 256, 256, 0, stream>>>(
    f, n, rot1, trans, rot2, s_rot1, s_trans, s_rot2, seed);
  NDT2D_LAUNCH_CHECK(ctr);
  return NDT2D_OK;
}
