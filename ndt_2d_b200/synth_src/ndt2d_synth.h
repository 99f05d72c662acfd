/* ndt2d_synth.h -- synthetic laser world (TEST / BENCH INFRASTRUCTURE, host code only).
 *
 * Built into its own library, ndt_2d_b200/lib/libndt2d_synth.so (synth_src/synth.cpp, g++): it is
 * not part of the product ABI (include/ndt2d_b200.h) and links nothing of it, so the CPU
 * reference arm of bench.py maps no product code.  Shared by tests, bench.py, the golden
 * generators and the oracle so that every implementation sees bit-identical inputs. */
#ifndef NDT2D_SYNTH_H_
#define NDT2D_SYNTH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDT2D_SYNTH_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------
 * Synthetic laser world (host code; shared by tests and bench so that the
 * oracle and the device path see identical inputs).  Not part of the
 * reference; SURVEY.md section 8(d) defines it.
 * ---------------------------------------------------------------------- */

/* n_obstacles axis-aligned rectangles (xmin, ymin, xmax, ymax) inside a square
 * arena of side `arena`: sides U[side_min, side_max] m, centres
 * U[5, arena-5]^2, SplitMix64(seed). */
NDT2D_SYNTH_API int ndt2d_synth_world(
  uint64_t seed, double arena, int n_obstacles, double side_min, double side_max,
  double * rects4);

/* Ray-casts `beams` beams (angle -pi + i*2pi/beams in the sensor frame) from
 * each pose against the arena walls and the rectangles, adds N(0, sigma)
 * range noise (stream seed + scan index), drops returns beyond range_max and
 * writes sensor-frame points.  pt_offsets gets n_scans+1 entries; pts_xy must
 * hold 2*beams*n_scans doubles.  Multi-threaded on the host. */
NDT2D_SYNTH_API int ndt2d_synth_scans(
  const double * rects4, int n_rects, double arena, const double * poses3, size_t n_scans,
  int beams, double range_max, double noise_sigma, uint64_t seed, uint64_t * pt_offsets,
  double * pts_xy);

/* SplitMix64-based uniform doubles in [0,1): out[i] for stream `seed`. */
NDT2D_SYNTH_API void ndt2d_synth_uniform(uint64_t seed, size_t n, double * out);
/* Standard normal variates (Box-Muller on the uniform stream). */
NDT2D_SYNTH_API void ndt2d_synth_normal(uint64_t seed, size_t n, double * out);

#ifdef __cplusplus
}
#endif

#endif  /* NDT2D_SYNTH_H_ */
