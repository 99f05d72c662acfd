// synth.cpp -- deterministic synthetic laser world (host code; test / bench infrastructure,
// its own library libndt2d_synth.so -- not part of the product).
//
// square arena with outer walls and axis-aligned
// rectangular obstacles (denser and smaller than the survey's first sketch so
// that most beams return within range_max = 10 m); beams ray-cast in double; Gaussian range noise;
// returns beyond range_max dropped like the node does (ndt_mapper.cpp:433-451
// keeps only finite ranges <= range_max); points stored in the sensor frame
// (identity laser transform), the layout Scan::setPoints receives.
// Shared by tests and bench so the oracle and the device see the same input.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "ndt2d_synth.h"

namespace
{

inline uint64_t splitmix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// i-th uniform of stream `seed`, in [0, 1)
inline double uniform_at(uint64_t seed, uint64_t i)
{
  const uint64_t z = splitmix64(seed * 0xD1342543DE82EF95ull + i * 0x9E3779B97F4A7C15ull);
  return static_cast<double>(z >> 11) * (1.0 / 9007199254740992.0);
}

inline double normal_at(uint64_t seed, uint64_t i)
{
  const double u1 = uniform_at(seed, 2 * i);
  const double u2 = uniform_at(seed, 2 * i + 1);
  return std::sqrt(-2.0 * std::log(1.0 - u1)) * std::cos(6.283185307179586476925 * u2);
}

// Distance along the ray to the boundary of box [x0,x1]x[y0,y1]; <0 = no hit.
inline double ray_box(double px, double py, double dx, double dy,
  double x0, double y0, double x1, double y1)
{
  const double inf = 1e300;
  double tx_lo = -inf, tx_hi = inf, ty_lo = -inf, ty_hi = inf;
  if (dx != 0.0) {
    const double a = (x0 - px) / dx, b = (x1 - px) / dx;
    tx_lo = std::min(a, b);
    tx_hi = std::max(a, b);
  } else if (px < x0 || px > x1) {
    return -1.0;
  }
  if (dy != 0.0) {
    const double a = (y0 - py) / dy, b = (y1 - py) / dy;
    ty_lo = std::min(a, b);
    ty_hi = std::max(a, b);
  } else if (py < y0 || py > y1) {
    return -1.0;
  }
  const double t_in = std::max(tx_lo, ty_lo);
  const double t_out = std::min(tx_hi, ty_hi);
  if (t_out < 0.0 || t_in > t_out) {
    return -1.0;
  }
  return t_in > 0.0 ? t_in : t_out;  // from outside: entry; from inside: exit
}

}  // namespace

extern "C" {

NDT2D_SYNTH_API void ndt2d_synth_uniform(uint64_t seed, size_t n, double * out)
{
  for (size_t i = 0; i < n; ++i) {out[i] = uniform_at(seed, i);}
}

NDT2D_SYNTH_API void ndt2d_synth_normal(uint64_t seed, size_t n, double * out)
{
  for (size_t i = 0; i < n; ++i) {out[i] = normal_at(seed, i);}
}

NDT2D_SYNTH_API int ndt2d_synth_world(
  uint64_t seed, double arena, int n_obstacles, double side_min, double side_max, double * rects4)
{
  if (!rects4 || n_obstacles < 0 || !(arena > 10.0) || !(side_min > 0.0) || !(side_max >= side_min)) {
    return 1;  /* invalid argument */
  }
  for (int k = 0; k < n_obstacles; ++k) {
    const double cx = 5.0 + (arena - 10.0) * uniform_at(seed, 4 * k + 0);
    const double cy = 5.0 + (arena - 10.0) * uniform_at(seed, 4 * k + 1);
    const double w = side_min + (side_max - side_min) * uniform_at(seed, 4 * k + 2);
    const double h = side_min + (side_max - side_min) * uniform_at(seed, 4 * k + 3);
    rects4[4 * k + 0] = cx - 0.5 * w;
    rects4[4 * k + 1] = cy - 0.5 * h;
    rects4[4 * k + 2] = cx + 0.5 * w;
    rects4[4 * k + 3] = cy + 0.5 * h;
  }
  return 0;
}

NDT2D_SYNTH_API int ndt2d_synth_scans(
  const double * rects4, int n_rects, double arena, const double * poses3, size_t n_scans,
  int beams, double range_max, double noise_sigma, uint64_t seed, uint64_t * pt_offsets,
  double * pts_xy)
{
  if (!poses3 || !pt_offsets || !pts_xy || beams <= 0 || n_rects < 0 || (n_rects && !rects4)) {
    return 1;  /* invalid argument */
  }
  // Pass 1 (parallel): every scan writes into its own fixed-size slot and
  // records its count; pass 2 compacts in scan order.
  std::vector<double> slots(static_cast<size_t>(2) * beams * n_scans);
  std::vector<uint32_t> counts(n_scans);

  auto work = [&](size_t lo, size_t hi) {
      for (size_t s = lo; s < hi; ++s) {
        const double px = poses3[3 * s], py = poses3[3 * s + 1], th = poses3[3 * s + 2];
        double * out = slots.data() + static_cast<size_t>(2) * beams * s;
        uint32_t n = 0;
        for (int b = 0; b < beams; ++b) {
          const double alpha = -M_PI + b * (2.0 * M_PI / beams);
          const double dx = std::cos(th + alpha), dy = std::sin(th + alpha);
          double t = ray_box(px, py, dx, dy, 0.0, 0.0, arena, arena);
          if (t < 0.0) {t = 1e300;}
          for (int k = 0; k < n_rects; ++k) {
            const double * r = rects4 + 4 * k;
            const double tk = ray_box(px, py, dx, dy, r[0], r[1], r[2], r[3]);
            if (tk >= 0.0 && tk < t) {t = tk;}
          }
          const double range = t + noise_sigma * normal_at(seed + s, static_cast<uint64_t>(b));
          if (!std::isfinite(range) || range > range_max || range <= 0.0) {
            continue;
          }
          out[2 * n] = range * std::cos(alpha);
          out[2 * n + 1] = range * std::sin(alpha);
          ++n;
        }
        counts[s] = n;
      }
    };

  unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
  if (n_scans < 64) {nthreads = 1;}
  std::vector<std::thread> pool;
  const size_t chunk = (n_scans + nthreads - 1) / nthreads;
  for (unsigned t = 0; t < nthreads; ++t) {
    const size_t lo = t * chunk, hi = std::min(n_scans, lo + chunk);
    if (lo >= hi) {break;}
    if (nthreads == 1) {work(lo, hi);} else {pool.emplace_back(work, lo, hi);}
  }
  for (auto & th : pool) {th.join();}

  uint64_t off = 0;
  for (size_t s = 0; s < n_scans; ++s) {
    pt_offsets[s] = off;
    const double * src = slots.data() + static_cast<size_t>(2) * beams * s;
    std::copy(src, src + 2 * counts[s], pts_xy + 2 * off);
    off += counts[s];
  }
  pt_offsets[n_scans] = off;
  return 0;
}

}  // extern "C"
