/*
 * ndt2d_oracle.c -- CPU restatement of the ndt_2d scan-matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * are the only callers.  The product (ndt_2d_b200/) never links, loads or
 * calls anything in oracle/.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_*.py)
 *  (1) against every known-answer value in the reference's own gtests
 *      (test/ndt_model_tests.cpp:32-230, test/particle_tests.cpp:47-72), and
 *  (2) against the reference's own sources compiled in place into
 *      oracle/_ref/libndt2d_ref.so (see oracle/Makefile, oracle/ref_driver.cpp).
 *
 * Everything is IEEE double, scalar, sequential -- the same operation order
 * as the reference (all file:line citations are relative to /root/reference).
 * Compile with -ffp-contract=off (no FMA contraction), like the reference's
 * default x86-64 Release build (CMakeLists.txt:4-11).
 *
 * The data layout follows the reference too (AoS cell of 128 bytes, a fresh
 * `points_inner` vector rewritten per candidate, per-call point-vector
 * copies), so that timing this file is a fair stand-in for the reference's
 * CPU path when oracle/_ref is not available.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* Cell  (include/ndt_2d/ndt_model.hpp:43-65, src/ndt_model.cpp:40-116) */
/* ------------------------------------------------------------------ */

/* Same member order and 16-byte alignment as the reference's Eigen members:
 * valid(+pad) | n | mean[2] | covariance[4] | correlation[4] | information[4]
 * = 128 bytes.  Matrices are stored (row, col) -> m[row * 2 + col]. */
typedef struct
{
  int valid;
  int pad_;
  double n;
  double mean[2];
  double covariance[4];
  double correlation[4];
  double information[4];
} orc_cell;

static void cell_init(orc_cell * c)
{
  memset(c, 0, sizeof(*c)); /* ndt_model.cpp:40-48 */
}

/* ndt_model.cpp:50-63 -- running mean and running second moment, upper
 * triangle only (correlation(1,0) is never written). */
static void cell_add_point(orc_cell * c, double px, double py)
{
  const double p[2] = {px, py};
  const double n = c->n;
  /* mean = (mean * n + point) / (n + 1)                       :52 */
  c->mean[0] = (c->mean[0] * n + p[0]) / (n + 1);
  c->mean[1] = (c->mean[1] * n + p[1]) / (n + 1);
  for (size_t i = 0; i < 2; ++i) {
    for (size_t j = i; j < 2; ++j) {
      /* :57 */
      c->correlation[i * 2 + j] = (c->correlation[i * 2 + j] * n + p[i] * p[j]) / (n + 1);
    }
  }
  c->n += 1; /* :61 */
  c->valid = 0;
}

/* Eigenvalues of the symmetric 2x2 covariance.  The reference calls
 * Eigen::EigenSolver<Matrix2d>(cov).eigenvalues().real() (ndt_model.cpp:84-85),
 * a Hessenberg + real-Schur iteration in the un-vendored Eigen3 dependency
 * (package.xml:12, version unpinned).  For a symmetric 2x2 input its result
 * is the pair  (a+d)/2 -/+ sqrt(((a-d)/2)^2 + b^2)  up to a few ulps; the
 * values are only used to pick a branch and the clamp determinant, so ulps
 * only matter on the measure-zero boundary small == 0.001*large. */
static void sym2_eigenvalues(double a, double b, double d, double * small, double * large)
{
  const double mid = 0.5 * (a + d);
  const double half = 0.5 * (a - d);
  const double rad = sqrt(half * half + b * b);
  *small = mid - rad;
  *large = mid + rad;
}

/* ndt_model.cpp:65-103 */
static void cell_compute(orc_cell * c)
{
  if (c->valid || c->n < 3) {
    return; /* :68-71 */
  }
  const double scale = c->n / (c->n - 1); /* :73 */
  for (size_t i = 0; i < 2; ++i) {
    for (size_t j = i; j < 2; ++j) {
      /* :78-79 */
      c->covariance[i * 2 + j] =
        (c->correlation[i * 2 + j] - (c->mean[i] * c->mean[j])) * scale;
      c->covariance[j * 2 + i] = c->covariance[i * 2 + j];
    }
  }
  double small, large;
  sym2_eigenvalues(c->covariance[0], c->covariance[1], c->covariance[3], &small, &large);
  if (small > large) { /* :87 */
    double t = small;
    small = large;
    large = t;
  }
  if (small < 0.001 * large) {
    /* :88-96  clamp: determinant of the clamped matrix, element-wise divide */
    const double determinant = (0.001 * large) * large;
    c->information[0] = c->covariance[3] / determinant;
    c->information[1] = -c->covariance[2] / determinant;
    c->information[2] = -c->covariance[1] / determinant;
    c->information[3] = c->covariance[0] / determinant;
  } else {
    /* :99  Eigen 2x2 inverse: invdet = 1/det, then multiply */
    const double a = c->covariance[0], b = c->covariance[1];
    const double cc = c->covariance[2], d = c->covariance[3];
    const double invdet = 1.0 / (a * d - b * cc);
    c->information[0] = d * invdet;
    c->information[2] = -cc * invdet;
    c->information[1] = -b * invdet;
    c->information[3] = a * invdet;
  }
  c->valid = 1; /* :102 */
}

/* ndt_model.cpp:105-116.  Eigen evaluates  -0.5 * q^T * I * q  left to
 * right: ((-0.5 * q^T) * I) * q. */
static double cell_score(const orc_cell * c, double px, double py)
{
  if (c->n < 5) {
    return 0.0; /* :107-111 */
  }
  const double qx = px - c->mean[0];
  const double qy = py - c->mean[1];
  const double hx = -0.5 * qx, hy = -0.5 * qy;
  const double r0 = hx * c->information[0] + hy * c->information[2];
  const double r1 = hx * c->information[1] + hy * c->information[3];
  const double exponent = r0 * qx + r1 * qy;
  return exp(exponent); /* :115 */
}

/* Cell-level entry points so the reference's Cell gtests can be replayed. */
ORC_API void * orc_cell_new(void)
{
  orc_cell * c = (orc_cell *)malloc(sizeof(orc_cell));
  cell_init(c);
  return c;
}
ORC_API void orc_cell_free(void * c) {free(c);}
ORC_API void orc_cell_add_point(void * c, double x, double y) {cell_add_point((orc_cell *)c, x, y);}
ORC_API void orc_cell_compute(void * c) {cell_compute((orc_cell *)c);}
ORC_API double orc_cell_score(void * c, double x, double y)
{
  return cell_score((const orc_cell *)c, x, y);
}
/* out[16]: valid, n, mean[2], covariance[4], correlation[4], information[4] */
ORC_API void orc_cell_get(const void * cv, double * out)
{
  const orc_cell * c = (const orc_cell *)cv;
  out[0] = c->valid;
  out[1] = c->n;
  memcpy(out + 2, c->mean, 2 * sizeof(double));
  memcpy(out + 4, c->covariance, 4 * sizeof(double));
  memcpy(out + 8, c->correlation, 4 * sizeof(double));
  memcpy(out + 12, c->information, 4 * sizeof(double));
}

/* ------------------------------------------------------------------ */
/* NDT  (include/ndt_2d/ndt_model.hpp:67-134, src/ndt_model.cpp:118-218) */
/* ------------------------------------------------------------------ */

typedef struct
{
  double cell_size;
  size_t size_x, size_y;
  double origin_x, origin_y;
  orc_cell * cells;
} orc_ndt;

/* ndt_model.cpp:118-126 : size = size_t(size / cell + 1) (truncation) */
ORC_API void * orc_ndt_create(
  double cell_size, double size_x, double size_y, double origin_x, double origin_y)
{
  orc_ndt * m = (orc_ndt *)malloc(sizeof(orc_ndt));
  m->cell_size = cell_size;
  m->size_x = (size_t)((size_x / cell_size) + 1);
  m->size_y = (size_t)((size_y / cell_size) + 1);
  m->origin_x = origin_x;
  m->origin_y = origin_y;
  const size_t n = m->size_x * m->size_y;
  m->cells = (orc_cell *)calloc(n ? n : 1, sizeof(orc_cell));
  return m;
}

ORC_API void orc_ndt_destroy(void * mv)
{
  orc_ndt * m = (orc_ndt *)mv;
  if (!m) {return;}
  free(m->cells);
  free(m);
}

/* ndt_model.cpp:203-218 */
static int ndt_get_index(const orc_ndt * m, double x, double y)
{
  if (x < m->origin_x || y < m->origin_y) {
    return -1;
  }
  unsigned int grid_x = (unsigned int)((x - m->origin_x) / m->cell_size);
  unsigned int grid_y = (unsigned int)((y - m->origin_y) / m->cell_size);
  if (grid_x >= m->size_x || grid_y >= m->size_y) {
    return -1;
  }
  return (int)((grid_y * m->size_x) + grid_x);
}

ORC_API int orc_ndt_get_index(const void * m, double x, double y)
{
  return ndt_get_index((const orc_ndt *)m, x, y);
}

/* ndt_model.cpp:132-152.  p = pose; p += rotated point (add to pose LAST). */
ORC_API void orc_ndt_add_scan(void * mv, const double * pose, const double * pts_xy, size_t npts)
{
  orc_ndt * m = (orc_ndt *)mv;
  const double cos_th = cos(pose[2]);
  const double sin_th = sin(pose[2]);
  /* scan->getPoints() returns the vector by value (scan.cpp:67-70) */
  double * pts = (double *)malloc((npts ? npts : 1) * 2 * sizeof(double));
  memcpy(pts, pts_xy, npts * 2 * sizeof(double));
  for (size_t i = 0; i < npts; ++i) {
    double px = pose[0], py = pose[1];
    px += pts[2 * i] * cos_th - pts[2 * i + 1] * sin_th;
    py += pts[2 * i] * sin_th + pts[2 * i + 1] * cos_th;
    const int index = ndt_get_index(m, px, py);
    if (index >= 0) {
      cell_add_point(&m->cells[index], px, py);
    }
  }
  free(pts);
}

/* ndt_model.cpp:154-160 : every cell, occupied or not */
ORC_API void orc_ndt_compute(void * mv)
{
  orc_ndt * m = (orc_ndt *)mv;
  const size_t n = m->size_x * m->size_y;
  for (size_t i = 0; i < n; ++i) {
    cell_compute(&m->cells[i]);
  }
}

/* ndt_model.cpp:162-170 */
static double ndt_likelihood_xy(const orc_ndt * m, double x, double y)
{
  const int index = ndt_get_index(m, x, y);
  if (index >= 0) {
    return cell_score(&m->cells[index], x, y);
  }
  return 0.0;
}

ORC_API double orc_ndt_likelihood_point(const void * m, double x, double y)
{
  return ndt_likelihood_xy((const orc_ndt *)m, x, y);
}

/* ndt_model.cpp:178-187 : sequential f64 sum in point order */
ORC_API double orc_ndt_likelihood_points(const void * mv, const double * pts_xy, size_t npts)
{
  const orc_ndt * m = (const orc_ndt *)mv;
  double score = 0.0;
  for (size_t i = 0; i < npts; ++i) {
    score += ndt_likelihood_xy(m, pts_xy[2 * i], pts_xy[2 * i + 1]);
  }
  return score;
}

/* conversions.hpp:64-68 applied to (px, py, 1):  Translation3d(x,y,0) *
 * AngleAxisd(theta, Z).  Eigen's AngleAxis::toRotationMatrix with axis (0,0,1)
 * gives R = [[c, -s, 0], [s, c, 0], [0, 0, (1-c)+c]]; the isometry product is
 * translation + (R(i,0)*px + R(i,1)*py + R(i,2)*pz). */
static void pose_apply(double c, double s, double tx, double ty, double px, double py,
  double * ox, double * oy)
{
  const double pz = 1.0;
  *ox = ((c * px + (-s) * py) + 0.0 * pz) + tx;
  *oy = ((s * px + c * py) + 0.0 * pz) + ty;
}

/* ndt_model.cpp:189-201 : all points, positive sign, not normalised */
ORC_API double orc_ndt_likelihood_scan(
  const void * mv, const double * pose, const double * pts_xy, size_t npts)
{
  const orc_ndt * m = (const orc_ndt *)mv;
  const double c = cos(pose[2]), s = sin(pose[2]);
  double score = 0.0;
  for (size_t i = 0; i < npts; ++i) {
    double x, y;
    pose_apply(c, s, pose[0], pose[1], pts_xy[2 * i], pts_xy[2 * i + 1], &x, &y);
    score += ndt_likelihood_xy(m, x, y);
  }
  return score;
}

/* info[5]: size_x, size_y, origin_x, origin_y, cell_size */
ORC_API void orc_ndt_grid(const void * mv, double * info)
{
  const orc_ndt * m = (const orc_ndt *)mv;
  info[0] = (double)m->size_x;
  info[1] = (double)m->size_y;
  info[2] = m->origin_x;
  info[3] = m->origin_y;
  info[4] = m->cell_size;
}

/* Dump per-cell state, 16 doubles per cell (layout of orc_cell_get). */
ORC_API void orc_ndt_dump_cells(const void * mv, double * out)
{
  const orc_ndt * m = (const orc_ndt *)mv;
  const size_t n = m->size_x * m->size_y;
  for (size_t i = 0; i < n; ++i) {
    orc_cell_get(&m->cells[i], out + 16 * i);
  }
}

/* ------------------------------------------------------------------ */
/* ScanMatcherNDT  (src/scan_matcher_ndt.cpp:35-183)                    */
/* ------------------------------------------------------------------ */

typedef struct
{
  double resolution;                 /* ndt_resolution            :37 */
  double angular_res, angular_size;  /* search_angular_*          :39-40 */
  double linear_res, linear_size;    /* search_linear_*           :41-42 */
  size_t laser_max_beams;            /* laser_max_beams           :44 */
  double range_max;                  /* :46 */
  orc_ndt * ndt;
} orc_matcher;

ORC_API void * orc_matcher_create(
  double ndt_resolution, double search_angular_resolution, double search_angular_size,
  double search_linear_resolution, double search_linear_size, int laser_max_beams,
  double range_max)
{
  orc_matcher * m = (orc_matcher *)calloc(1, sizeof(orc_matcher));
  m->resolution = ndt_resolution;
  m->angular_res = search_angular_resolution;
  m->angular_size = search_angular_size;
  m->linear_res = search_linear_resolution;
  m->linear_size = search_linear_size;
  /* declare_parameter<int> assigned to a size_t member (:44, hpp:99) */
  m->laser_max_beams = (size_t)laser_max_beams;
  m->range_max = range_max;
  m->ndt = NULL;
  return m;
}

/* scan_matcher_ndt.cpp:180-183 */
ORC_API void orc_matcher_reset(void * mv)
{
  orc_matcher * m = (orc_matcher *)mv;
  orc_ndt_destroy(m->ndt);
  m->ndt = NULL;
}

ORC_API void orc_matcher_destroy(void * mv)
{
  if (!mv) {return;}
  orc_matcher_reset(mv);
  free(mv);
}

/* scan_matcher_ndt.cpp:49-74.  poses = 3 doubles per scan, pt_offsets has
 * n_scans+1 entries (in points), pts_xy interleaved x,y in the sensor frame.
 * Note max_* start at numeric_limits<double>::min() = DBL_MIN > 0 (:54,56). */
ORC_API void orc_matcher_add_scans(
  void * mv, size_t n_scans, const double * poses, const uint64_t * pt_offsets,
  const double * pts_xy)
{
  orc_matcher * m = (orc_matcher *)mv;
  double min_x = DBL_MAX, max_x = DBL_MIN, min_y = DBL_MAX, max_y = DBL_MIN;
  for (size_t k = 0; k < n_scans; ++k) {
    const double * pose = poses + 3 * k;
    min_x = fmin(pose[0] - m->range_max, min_x);
    max_x = fmax(pose[0] + m->range_max, max_x);
    min_y = fmin(pose[1] - m->range_max, min_y);
    max_y = fmax(pose[1] + m->range_max, max_y);
  }
  orc_ndt_destroy(m->ndt); /* make_unique replaces the old model (:66) */
  m->ndt = (orc_ndt *)orc_ndt_create(m->resolution, (max_x - min_x), (max_y - min_y), min_x, min_y);
  for (size_t k = 0; k < n_scans; ++k) {
    orc_ndt_add_scan(m->ndt, poses + 3 * k, pts_xy + 2 * pt_offsets[k],
      (size_t)(pt_offsets[k + 1] - pt_offsets[k]));
  }
  orc_ndt_compute(m->ndt);
}

ORC_API void * orc_matcher_ndt(void * mv) {return ((orc_matcher *)mv)->ndt;}

/* Replay of the accumulated-double loop `for (v = -size; v < size; v += res)`
 * (scan_matcher_ndt.cpp:103,117,119).  Returns the count; fills out[] when
 * non-NULL (up to cap entries). */
ORC_API size_t orc_loop_values(double size, double res, double * out, size_t cap)
{
  size_t n = 0;
  for (double v = -size; v < size; v += res) {
    if (out && n < cap) {out[n] = v;}
    ++n;
  }
  return n;
}

/* scan_matcher_ndt.cpp:76-149.
 *  pose3      : scan pose (x, y, theta)
 *  out_delta  : written only when a candidate scores < 0   (:128-134)
 *  delta_written : 1 if out_delta was written
 *  out_cov    : 3x3 row-major; always written when an NDT exists (:146)
 *  all_scores : optional, one entry per candidate in (dth, dx, dy) loop order
 *  theta_lo/hi: candidate theta-index window [lo, hi) actually evaluated; pass
 *               0 / SIZE_MAX for the full search.  The window exists only so
 *               the CPU baseline can time a bounded sample of a huge search;
 *               the loop variable is still replayed from -angular_size.
 * returns best_score / n  (0.0 and nothing written when there is no NDT). */
static double match_scan_window_ex(
  void * mv, const double * pose3, const double * pts_xy, size_t npts,
  double * out_delta, int * delta_written, double * out_cov, double * all_scores,
  size_t theta_lo, size_t theta_hi, uint64_t * n_candidates, double * partial16);

ORC_API double orc_matcher_match_scan_window(
  void * mv, const double * pose3, const double * pts_xy, size_t npts,
  double * out_delta, int * delta_written, double * out_cov, double * all_scores,
  size_t theta_lo, size_t theta_hi, uint64_t * n_candidates)
{
  return match_scan_window_ex(mv, pose3, pts_xy, npts, out_delta, delta_written, out_cov,
           all_scores, theta_lo, theta_hi, n_candidates, NULL);
}

/* The sequential search restricted to theta indices [theta_lo, theta_hi), reported as
 * the 16-double partial record of include/ndt2d_b200.h (ndt2d_matcher_search_staged):
 * what one rank of a theta-sliced search contributes.  Test infrastructure for the
 * multi-rank host logic (tests/test_multi_rank.py). */
ORC_API void orc_matcher_partial(
  void * mv, const double * pose3, const double * pts_xy, size_t npts,
  size_t theta_lo, size_t theta_hi, double * partial16)
{
  match_scan_window_ex(mv, pose3, pts_xy, npts, NULL, NULL, NULL, NULL, theta_lo, theta_hi, NULL,
    partial16);
}

static double match_scan_window_ex(
  void * mv, const double * pose3, const double * pts_xy, size_t npts,
  double * out_delta, int * delta_written, double * out_cov, double * all_scores,
  size_t theta_lo, size_t theta_hi, uint64_t * n_candidates, double * partial16)
{
  orc_matcher * m = (orc_matcher *)mv;
  if (partial16) {
    memset(partial16, 0, 16 * sizeof(double));
    partial16[1] = 1.0e300;
  }
  if (delta_written) {*delta_written = 0;}
  if (n_candidates) {*n_candidates = 0;}
  if (!m->ndt) {
    return 0.0; /* :80 */
  }
  double best_score = 0;
  double k[9] = {0}, u[3] = {0}, s = 0.0; /* :86-88 */

  /* local copies :91-92 */
  const double scan_pose[3] = {pose3[0], pose3[1], pose3[2]};
  double * points = (double *)malloc((npts ? npts : 1) * 2 * sizeof(double));
  memcpy(points, pts_xy, npts * 2 * sizeof(double));

  /* :95-96 */
  const size_t scan_points_to_use = m->laser_max_beams < npts ? m->laser_max_beams : npts;
  const double scan_step = (double)npts / (double)scan_points_to_use;

  double * outer = (double *)calloc((scan_points_to_use ? scan_points_to_use : 1) * 2, sizeof(double));
  double * inner = (double *)calloc((scan_points_to_use ? scan_points_to_use : 1) * 2, sizeof(double));

  uint64_t cand = 0;
  size_t ith = 0;
  size_t n_lin = 0;
  for (double d = -m->linear_size; d < m->linear_size; d += m->linear_res) {++n_lin;}
  double best_index = 1.0e300;
  for (double dth = -m->angular_size; dth < m->angular_size; dth += m->angular_res, ++ith) {
    if (ith < theta_lo || ith >= theta_hi) {
      continue;
    }
    uint64_t in_slice = 0;
    const double costh = cos(scan_pose[2] + dth); /* :106-107 */
    const double sinth = sin(scan_pose[2] + dth);
    for (size_t i = 0; i < scan_points_to_use; ++i) {
      const size_t scan_idx = (size_t)(i * scan_step); /* :110 */
      outer[2 * i] = points[2 * scan_idx] * costh - points[2 * scan_idx + 1] * sinth + scan_pose[0];
      outer[2 * i + 1] = points[2 * scan_idx] * sinth + points[2 * scan_idx + 1] * costh + scan_pose[1];
    }
    for (double dx = -m->linear_size; dx < m->linear_size; dx += m->linear_res) {
      for (double dy = -m->linear_size; dy < m->linear_size; dy += m->linear_res) {
        for (size_t i = 0; i < scan_points_to_use; ++i) { /* :121-125 */
          inner[2 * i] = outer[2 * i] + dx;
          inner[2 * i + 1] = outer[2 * i + 1] + dy;
        }
        const double score = -orc_ndt_likelihood_points(m->ndt, inner, scan_points_to_use);
        if (score < best_score) { /* :128 strict <, first wins */
          best_score = score;
          best_index = (double)((uint64_t)ith * n_lin * n_lin + in_slice);
          if (out_delta) {
            out_delta[0] = dx;
            out_delta[1] = dy;
            out_delta[2] = dth;
          }
          if (delta_written) {*delta_written = 1;}
        }
        /* :137-140   k += (x x^T) * score ; u += x * score ; s += score */
        const double x[3] = {dx, dy, dth};
        for (int r = 0; r < 3; ++r) {
          for (int c = 0; c < 3; ++c) {
            k[r * 3 + c] += (x[r] * x[c]) * score;
          }
          u[r] += x[r] * score;
        }
        s += score;
        if (all_scores) {all_scores[cand] = score;}
        ++cand;
        ++in_slice;
      }
    }
  }
  if (partial16) {
    partial16[0] = best_score;
    partial16[1] = best_index;
    partial16[2] = k[0]; partial16[3] = k[1]; partial16[4] = k[2];
    partial16[5] = k[4]; partial16[6] = k[5]; partial16[7] = k[8];
    partial16[8] = u[0]; partial16[9] = u[1]; partial16[10] = u[2];
    partial16[11] = s;
    partial16[12] = (double)cand;
    partial16[13] = (double)scan_points_to_use;
  }
  /* :146  covariance = (1/s) * k + ((1/(s*s)) * u) * u^T   (note the '+') */
  if (out_cov) {
    const double inv_s = 1 / s;
    const double inv_s2 = 1 / (s * s);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) {
        out_cov[r * 3 + c] = inv_s * k[r * 3 + c] + (inv_s2 * u[r]) * u[c];
      }
    }
  }
  if (n_candidates) {*n_candidates = cand;}
  free(points);
  free(outer);
  free(inner);
  return best_score / scan_points_to_use; /* :148 */
}

ORC_API double orc_matcher_match_scan(
  void * mv, const double * pose3, const double * pts_xy, size_t npts,
  double * out_delta, int * delta_written, double * out_cov, double * all_scores)
{
  return orc_matcher_match_scan_window(mv, pose3, pts_xy, npts, out_delta, delta_written,
           out_cov, all_scores, 0, SIZE_MAX, NULL);
}

/* scan_matcher_ndt.cpp:156-178 */
ORC_API double orc_matcher_score_points(
  void * mv, const double * pts_xy, size_t npts, const double * pose3)
{
  orc_matcher * m = (orc_matcher *)mv;
  if (!m->ndt) {
    return 0.0; /* :159 */
  }
  const double c = cos(pose3[2]), s = sin(pose3[2]);
  const size_t scan_points_to_use = m->laser_max_beams < npts ? m->laser_max_beams : npts;
  const double scan_step = (double)npts / (double)scan_points_to_use;
  double score = 0.0;
  for (size_t i = 0; i < scan_points_to_use; ++i) {
    const size_t scan_idx = (size_t)(i * scan_step);
    double x, y;
    pose_apply(c, s, pose3[0], pose3[1], pts_xy[2 * scan_idx], pts_xy[2 * scan_idx + 1], &x, &y);
    score += -ndt_likelihood_xy(m->ndt, x, y);
  }
  return score / scan_points_to_use;
}

/* ------------------------------------------------------------------ */
/* ParticleFilter measure / updateStatistics / resample                 */
/* (src/particle_filter.cpp:78-137, 163-218; kd_tree.hpp:97-189)        */
/* ------------------------------------------------------------------ */

/* angles::normalize_angle / shortest_angular_distance (ROS 2 `angles` package,
 * un-vendored dependency, package.xml:16, version unpinned; published
 * algorithm of the ROS 2 releases:
 *   normalize_angle(a) = r = fmod(a + pi, 2pi); r <= 0 ? r + pi : r - pi
 *   shortest_angular_distance(from, to) = normalize_angle(to - from) ) */
static double ang_normalize(double a)
{
  const double r = fmod(a + M_PI, 2.0 * M_PI);
  if (r <= 0.0) {return r + M_PI;}
  return r - M_PI;
}
static double ang_shortest(double from, double to)
{
  return ang_normalize(to - from);
}
ORC_API double orc_normalize_angle(double a) {return ang_normalize(a);}
ORC_API double orc_shortest_angular_distance(double from, double to) {return ang_shortest(from, to);}

/* particle_filter.cpp:78-89 (without the trailing updateStatistics) */
ORC_API void orc_pf_measure(
  void * matcher, const double * particles /*3*P*/, size_t P, const double * pts_xy, size_t npts,
  double * weights)
{
  for (size_t i = 0; i < P; ++i) {
    weights[i] = orc_matcher_score_points(matcher, pts_xy, npts, particles + 3 * i);
  }
}

/* particle_filter.cpp:163-218.  weights are normalised in place; mean[3] is
 * overwritten; cov[9] (row-major) keeps its previous (0,2),(1,2),(2,0),(2,1)
 * entries and ACCUMULATES into (2,2) (`+=` at :216, never reset). */
ORC_API void orc_pf_update_statistics(
  const double * particles, double * weights, size_t P, double * mean_out, double * cov)
{
  double sum_weight = 0.0;
  for (size_t i = 0; i < P; ++i) {sum_weight += weights[i];}
  for (size_t i = 0; i < P; ++i) {weights[i] /= sum_weight;}

  double mean[3] = {0, 0, 0};
  double corr[9] = {0};
  double sum_cos_th = 0.0, sum_sin_th = 0.0;
  for (size_t i = 0; i < P; ++i) {
    const double * p = particles + 3 * i;
    for (int j = 0; j < 3; ++j) {mean[j] += weights[i] * p[j];}
    sum_cos_th += weights[i] * cos(p[2]);
    sum_sin_th += weights[i] * sin(p[2]);
    for (size_t j = 0; j < 2; ++j) {
      for (size_t k = j; k < 2; ++k) {
        corr[j * 3 + k] += weights[i] * p[j] * p[k];
      }
    }
  }
  mean_out[0] = mean[0];
  mean_out[1] = mean[1];
  mean_out[2] = atan2(sum_sin_th, sum_cos_th);
  for (size_t j = 0; j < 2; ++j) {
    for (size_t k = j; k < 2; ++k) {
      cov[j * 3 + k] = corr[j * 3 + k] - mean[j] * mean[k];
      cov[k * 3 + j] = cov[j * 3 + k];
    }
  }
  for (size_t i = 0; i < P; ++i) {
    const double d = ang_shortest(particles[3 * i + 2], mean_out[2]);
    cov[8] += weights[i] * d * d;
  }
}

/* KD-tree bin key: static_cast<int>(coord / size), truncation toward zero
 * (kd_tree.hpp:99-102).  Only getLeafCount() is consumed by the filter
 * (particle_filter.cpp:117), and the leaf count equals the number of distinct
 * keys inserted so far (a new leaf is created exactly when the key is not
 * already present: kd_tree.hpp:104-110, 146-153, 168-177). */
ORC_API void orc_kd_key(const double * pose3, const double * sizes3, int * key3)
{
  for (int i = 0; i < 3; ++i) {
    key3[i] = (int)(pose3[i] / sizes3[i]);
  }
}

typedef struct {int k[3]; int used;} kd_slot;

static size_t kd_hash(const int * k, size_t mask)
{
  uint64_t h = (uint32_t)k[0] * 0x9E3779B97F4A7C15ull;
  h ^= (uint32_t)k[1] * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
  h ^= (uint32_t)k[2] * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
  return (size_t)(h ^ (h >> 29)) & mask;
}

/* leaf_counts[i] = KDTree::getLeafCount() after inserting poses[0..i]. */
ORC_API void orc_kd_leaf_counts(
  const double * poses /*3*N*/, size_t N, const double * sizes3, uint64_t * leaf_counts)
{
  size_t cap = 16;
  while (cap < 4 * N) {cap <<= 1;}
  kd_slot * tab = (kd_slot *)calloc(cap, sizeof(kd_slot));
  uint64_t leaves = 0;
  for (size_t i = 0; i < N; ++i) {
    int key[3];
    orc_kd_key(poses + 3 * i, sizes3, key);
    size_t h = kd_hash(key, cap - 1);
    for (;; ) {
      if (!tab[h].used) {
        tab[h].used = 1;
        memcpy(tab[h].k, key, sizeof(key));
        ++leaves;
        break;
      }
      if (tab[h].k[0] == key[0] && tab[h].k[1] == key[1] && tab[h].k[2] == key[2]) {
        break;
      }
      h = (h + 1) & (cap - 1);
    }
    leaf_counts[i] = leaves;
  }
  free(tab);
}

/* particle_filter.cpp:91-137 with the random source made injectable:
 * uniforms[i] is the value std::generate_canonical<double,53>(gen_) would
 * return for draw i; std::discrete_distribution (libstdc++ 13
 * bits/random.tcc:2655-2714) normalises the weights by their sequential sum,
 * takes sequential partial sums, forces the last to 1.0 and returns
 * lower_bound(cp, u).  KD bin sizes are (0.5, 0.5, 0.2671)
 * (particle_filter.cpp:44).  Does NOT run the trailing updateStatistics.
 * out_particles (3*max), out_weights (max), out_index (max) receive the
 * resampled set; returns its size.  n_uniforms_used = draws consumed. */
ORC_API size_t orc_pf_resample(
  const double * particles, const double * weights, size_t P,
  size_t min_particles, size_t max_particles, double kld_err, double kld_z,
  const double * uniforms, double * out_particles, double * out_weights, uint64_t * out_index)
{
  const double sizes[3] = {0.5, 0.5, 0.2671};
  double * cp = (double *)malloc((P ? P : 1) * sizeof(double));
  int degenerate = P < 2; /* _M_prob.size() < 2 -> always returns 0 */
  if (!degenerate) {
    double sum = 0.0;
    for (size_t i = 0; i < P; ++i) {sum += weights[i];}
    double acc = 0.0;
    for (size_t i = 0; i < P; ++i) {
      acc += weights[i] / sum;
      cp[i] = acc;
    }
    cp[P - 1] = 1.0;
  }

  size_t cap = 16;
  while (cap < 4 * max_particles) {cap <<= 1;}
  kd_slot * tab = (kd_slot *)calloc(cap, sizeof(kd_slot));
  size_t leaves = 0;

  size_t Mx = max_particles; /* :106 */
  size_t size = 0;
  while (size < (min_particles > Mx ? min_particles : Mx)) { /* :108 */
    size_t p = 0;
    if (!degenerate) {
      const double u = uniforms[size];
      size_t lo = 0, hi = P; /* lower_bound: first cp[k] >= u */
      while (lo < hi) {
        const size_t mid = lo + (hi - lo) / 2;
        if (cp[mid] < u) {lo = mid + 1;} else {hi = mid;}
      }
      p = lo;
    }
    /* kd_tree_.insert(particles_[p], weights_[p])  :112 */
    int key[3];
    orc_kd_key(particles + 3 * p, sizes, key);
    size_t h = kd_hash(key, cap - 1);
    for (;; ) {
      if (!tab[h].used) {
        tab[h].used = 1;
        memcpy(tab[h].k, key, sizeof(key));
        ++leaves;
        break;
      }
      if (tab[h].k[0] == key[0] && tab[h].k[1] == key[1] && tab[h].k[2] == key[2]) {break;}
      h = (h + 1) & (cap - 1);
    }
    memcpy(out_particles + 3 * size, particles + 3 * p, 3 * sizeof(double)); /* :113 */
    out_weights[size] = weights[p]; /* :114 old weight */
    if (out_index) {out_index[size] = p;}
    ++size;

    const size_t k = leaves; /* :117-124 */
    if (k > 1) {
      const double a = (k - 1) / (2.0 * kld_err);
      const double b = 2.0 / (9.0 * (k - 1));
      const double c = 1.0 - b + sqrt(b) * kld_z;
      Mx = (size_t)(a * c * c * c);
    }
    if (size >= max_particles) { /* :127-130 */
      break;
    }
  }
  free(tab);
  free(cp);
  return size;
}

/* ------------------------------------------------------------------ */
/* Candidate selection of the loop closure:        */
/* Scan::update's barycenter (src/scan.cpp:72-91) and Graph::findNearest            */
/* (src/graph.cpp:167-189).  findNearest runs nanoflann (un-vendored dependency,    */
/* package.xml, version unpinned): KDTreeSingleIndexAdaptor<L2_Simple_Adaptor<double,*/
/* GraphAdapter>, GraphAdapter, 2>::radiusSearch with default SearchParams.  Its     */
/* published behaviour, restated: evalMetric = sum over the dimensions of            */
/* (query[d] - point[d])^2 accumulated from 0 in dimension order; RadiusResultSet    */
/* keeps a point when dist < radius (strict; radius is a SQUARED distance); the      */
/* result is sorted by ascending distance (order among equal distances unspecified;  */
/* here: ascending index).  graph.cpp cannot be compiled here (nanoflann, rosbag2)   */
/* and no reference test calls it: findNearest is PARITY UNPINNED.                   */
/* ------------------------------------------------------------------ */
ORC_API void orc_scan_barycenter(const double * pose3, const double * pts_xy, size_t n, double * out3)
{
  const double cos_theta = cos(pose3[2]), sin_theta = sin(pose3[2]); /* scan.cpp:74-75 */
  out3[0] = pose3[0]; /* :78 */
  out3[1] = pose3[1];
  out3[2] = pose3[2];
  if (n) {
    double cx = 0.0, cy = 0.0;
    for (size_t i = 0; i < n; ++i) { /* :82-86 */
      cx += cos_theta * pts_xy[2 * i] - sin_theta * pts_xy[2 * i + 1];
      cy += sin_theta * pts_xy[2 * i] + cos_theta * pts_xy[2 * i + 1];
    }
    out3[0] += cx / n; /* :87-88 */
    out3[1] += cy / n;
  }
}

ORC_API size_t orc_find_nearest(
  const double * scan_xy, size_t n_scans, long long limit_scan_index, const double * query_xy,
  double radius_sq, unsigned long long * out_indices, double * out_dist_sq)
{
  /* graph.cpp:171 */
  size_t limit = (limit_scan_index > 0) ? (size_t)limit_scan_index : n_scans;
  if (limit > n_scans) {limit = n_scans;}
  size_t m = 0;
  for (size_t i = 0; i < limit; ++i) {
    double result = 0.0;
    for (size_t d = 0; d < 2; ++d) {
      const double diff = query_xy[d] - scan_xy[2 * i + d];
      result += diff * diff;
    }
    if (result < radius_sq) {
      /* insertion keeps (distance, index) ascending */
      size_t k = m;
      while (k > 0 && out_dist_sq[k - 1] > result) {
        out_dist_sq[k] = out_dist_sq[k - 1];
        out_indices[k] = out_indices[k - 1];
        --k;
      }
      out_dist_sq[k] = result;
      out_indices[k] = i;
      ++m;
    }
  }
  return m;
}

/* ------------------------------------------------------------------ */
/* LaserScan -> Scan points (Mapper::laserCallback, src/ndt_mapper.cpp:385-453)   */
/* SURVEY.md section 8(f) rank 3.  The reference code sits inside the ROS node and */
/* cannot be compiled here, and no reference test covers it: this restatement is   */
/* PARITY UNPINNED (checked only against hand-computed cases).                     */
/* ------------------------------------------------------------------ */
ORC_API size_t orc_laser_to_points(
  const float * ranges, size_t n, float angle_min, float angle_increment, double range_max,
  const double * laser_tf3, const double * translation3, int laser_inverted, double * out_xy)
{
  /* :392-395 */
  const double pm_x = translation3[0] / n, pm_y = translation3[1] / n, pm_th = translation3[2] / n;
  /* :403-404 */
  const double cos_lt = cos(laser_tf3[2]), sin_lt = sin(laser_tf3[2]);
  size_t m = 0;
  if (laser_inverted) {
    for (size_t i = n - 1; i > 0 && n > 0; --i) { /* :411 -- index 0 is never visited */
      if (isnan(ranges[i]) || ranges[i] > range_max) {continue;} /* :414 */
      const double angle = -(angle_min + i * angle_increment);   /* float arithmetic, :416 */
      const double lx = cos(angle) * ranges[i], ly = sin(angle) * ranges[i];
      const double px = cos_lt * lx - sin_lt * ly + laser_tf3[0];
      const double py = sin_lt * lx + cos_lt * ly + laser_tf3[1];
      const double cos_tt = cos(translation3[2] - (pm_th * i));
      const double sin_tt = sin(translation3[2] - (pm_th * i));
      out_xy[2 * m] = cos_tt * px - sin_tt * py + (translation3[0] - (pm_x * i));
      out_xy[2 * m + 1] = sin_tt * px + cos_tt * py + (translation3[1] - (pm_y * i));
      ++m;
    }
  } else {
    for (size_t i = 0; i < n; ++i) {
      if (isnan(ranges[i]) || ranges[i] > range_max) {continue;} /* :436 */
      const double angle = (angle_min + i * angle_increment);    /* float arithmetic, :438 */
      const double lx = cos(angle) * ranges[i], ly = sin(angle) * ranges[i];
      const double px = cos_lt * lx - sin_lt * ly + laser_tf3[0];
      const double py = sin_lt * lx + cos_lt * ly + laser_tf3[1];
      const double cos_tt = cos(pm_th * i);
      const double sin_tt = sin(pm_th * i);
      out_xy[2 * m] = cos_tt * px - sin_tt * py + (pm_x * i);
      out_xy[2 * m + 1] = sin_tt * px + cos_tt * py + (pm_y * i);
      ++m;
    }
  }
  return m;
}

/* ------------------------------------------------------------------ */
/* OccupancyGrid (src/occupancy_grid.cpp:35-185), SURVEY.md 8(f) rank 4 */
/* ------------------------------------------------------------------ */
typedef struct
{
  double resolution, occ_thresh;
  double min_x, max_x, min_y, max_y;
  size_t num_scans;
} orc_occ;

ORC_API void * orc_occ_create(double resolution, double occ_thresh)
{
  orc_occ * g = (orc_occ *)calloc(1, sizeof(orc_occ)); /* bounds and num_scans start at 0 (:38-43) */
  g->resolution = resolution;
  g->occ_thresh = occ_thresh;
  return g;
}
ORC_API void orc_occ_destroy(void * g) {free(g);}

/* updateBounds (:155-185) */
static void occ_update_bounds(
  orc_occ * g, size_t n_scans, const double * poses, const uint64_t * offsets, const double * pts)
{
  const size_t start_idx = g->num_scans;
  g->num_scans = n_scans;
  for (size_t i = start_idx; i < g->num_scans; ++i) {
    const double x = poses[3 * i], y = poses[3 * i + 1];
    const double cos_th = cos(poses[3 * i + 2]), sin_th = sin(poses[3 * i + 2]);
    for (uint64_t p = offsets[i]; p < offsets[i + 1]; ++p) {
      double px = x, py = y;
      px += pts[2 * p] * cos_th - pts[2 * p + 1] * sin_th;
      py += pts[2 * p] * sin_th + pts[2 * p + 1] * cos_th;
      g->min_x = px < g->min_x ? px : g->min_x; /* std::min(p.x, min_x_) */
      g->max_x = g->max_x < px ? px : g->max_x; /* std::max(p.x, max_x_) */
      g->min_y = py < g->min_y ? py : g->min_y;
      g->max_y = g->max_y < py ? py : g->max_y;
    }
  }
  g->min_x = floor(g->min_x / g->resolution) * g->resolution;
  g->max_x = ceil(g->max_x / g->resolution) * g->resolution;
  g->min_y = floor(g->min_y / g->resolution) * g->resolution;
  g->max_y = ceil(g->max_y / g->resolution) * g->resolution;
}

/* getMsg (:47-152); info5 = {width, height, origin_x, origin_y, (float)resolution};
 * data (capacity cells) filled if large enough; returns the number of cells. */
ORC_API size_t orc_occ_render(
  void * gv, size_t n_scans, const double * poses, const uint64_t * offsets, const double * pts,
  double * info5, int8_t * data, size_t capacity)
{
  orc_occ * g = (orc_occ *)gv;
  if (n_scans != g->num_scans) {occ_update_bounds(g, n_scans, poses, offsets, pts);} /* :51-54 */
  const double pad = 5 * g->resolution;
  const uint32_t width = (uint32_t)((g->max_x - g->min_x + 2 * pad) / g->resolution);
  const uint32_t height = (uint32_t)((g->max_y - g->min_y + 2 * pad) / g->resolution);
  const double origin_x = g->min_x - pad, origin_y = g->min_y - pad;
  info5[0] = width;
  info5[1] = height;
  info5[2] = origin_x;
  info5[3] = origin_y;
  info5[4] = (float)g->resolution;
  const size_t n_cells = (size_t)width * height;
  if (!data || capacity < n_cells) {return n_cells;}
  int * hit = (int *)calloc(n_cells ? n_cells : 1, sizeof(int));
  int * empty = (int *)calloc(n_cells ? n_cells : 1, sizeof(int));
  for (size_t s = 0; s < n_scans; ++s) {
    const double pose_x = poses[3 * s], pose_y = poses[3 * s + 1];
    const double cos_th = cos(poses[3 * s + 2]), sin_th = sin(poses[3 * s + 2]);
    const int start_x = (int)((pose_x - origin_x) / g->resolution);
    const int start_y = (int)((pose_y - origin_y) / g->resolution);
    for (uint64_t p = offsets[s]; p < offsets[s + 1]; ++p) {
      const double point_x = pts[2 * p] * cos_th - pts[2 * p + 1] * sin_th + pose_x;
      const double point_y = pts[2 * p] * sin_th + pts[2 * p + 1] * cos_th + pose_y;
      const int end_x = (int)((point_x - origin_x) / g->resolution);
      const int end_y = (int)((point_y - origin_y) / g->resolution);
      int dx = abs(end_x - start_x);
      int sx = (start_x < end_x) ? 1 : -1;
      int dy = -abs(end_y - start_y);
      int sy = (start_y < end_y) ? 1 : -1;
      int error = dx + dy;
      int x = start_x, y = start_y;
      while (1) {
        const int index = x + y * (int)width;
        /* the reference does not bounds-check; stale bounds could leave the grid */
        const int inside = x >= 0 && y >= 0 && (uint32_t)x < width && (uint32_t)y < height;
        if (x == end_x && y == end_y) {
          if (inside) {++hit[index];}
          break;
        }
        if (inside) {++empty[index];}
        if (2 * error >= dy) {
          if (x == end_x) {
            if (inside) {++hit[index];}
            break;
          }
          error = error + dy;
          x += sx;
        }
        if (2 * error <= dx) {
          if (y == end_y) {
            if (inside) {++hit[index];}
            break;
          }
          error = error + dx;
          y += sy;
        }
      }
    }
  }
  for (size_t i = 0; i < n_cells; ++i) {
    const double touches = hit[i] + empty[i];
    data[i] = -1;
    if (touches > 0.5) {
      data[i] = ((double)hit[i] / touches > g->occ_thresh) ? 100 : 0;
    }
  }
  free(hit);
  free(empty);
  return n_cells;
}
