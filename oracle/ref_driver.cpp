// ref_driver.cpp -- C entry points around the REFERENCE's own classes.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile together with the
// reference sources compiled unmodified, in place, from /root/reference/src
// (ndt_model.cpp, scan.cpp, scan_matcher_ndt.cpp, particle_filter.cpp,
// motion_model.cpp) against the stand-in headers in oracle/ref_shim/.  The
// output oracle/_ref/libndt2d_ref.so is git-ignored, travels to the GPU box,
// and is used to (a) pin oracle/ndt2d_oracle.c, (b) generate tests/golden/,
// (c) serve as bench.py's cpu_baseline / --impl reference ("kind":"reference").
// No reference source is copied into this repository.
//
// The ref_* functions mirror the orc_* functions of ndt2d_oracle.c one to one.

#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

// The reference keeps the filter state private and seeds its generator from
// std::random_device (particle_filter.cpp:38).  To replay it deterministically
// this translation unit (only) looks at those members; access specifiers do
// not change the class layout.  Standard headers are included above, before
// the redefinition.
#define private public
#define protected public
#include <angles/angles.h>
#include <ndt_2d/kd_tree.hpp>
#include <ndt_2d/motion_model.hpp>
#include <ndt_2d/ndt_model.hpp>
#include <ndt_2d/occupancy_grid.hpp>
#include <ndt_2d/particle_filter.hpp>
#include <ndt_2d/scan.hpp>
#include <ndt_2d/scan_matcher_ndt.hpp>
#undef private
#undef protected

#define REF_API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------- Cell
REF_API void * ref_cell_new() {return new ndt_2d::Cell();}
REF_API void ref_cell_free(void * c) {delete static_cast<ndt_2d::Cell *>(c);}
REF_API void ref_cell_add_point(void * c, double x, double y)
{
  Eigen::Vector2d p(x, y);
  static_cast<ndt_2d::Cell *>(c)->addPoint(p);
}
REF_API void ref_cell_compute(void * c) {static_cast<ndt_2d::Cell *>(c)->compute();}
REF_API double ref_cell_score(void * c, double x, double y)
{
  Eigen::Vector2d p(x, y);
  return static_cast<ndt_2d::Cell *>(c)->score(p);
}
static void cell_get(const ndt_2d::Cell & c, double * out)
{
  out[0] = c.valid ? 1.0 : 0.0;
  out[1] = c.n;
  out[2] = c.mean(0);
  out[3] = c.mean(1);
  for (int i = 0; i < 2; ++i) {
    for (int j = 0; j < 2; ++j) {
      out[4 + i * 2 + j] = c.covariance(i, j);
      out[8 + i * 2 + j] = c.correlation(i, j);
      out[12 + i * 2 + j] = c.information(i, j);
    }
  }
}
REF_API void ref_cell_get(const void * c, double * out)
{
  cell_get(*static_cast<const ndt_2d::Cell *>(c), out);
}

// ---------------------------------------------------------------- NDT
static ndt_2d::ScanPtr make_scan(const double * pose, const double * pts_xy, size_t npts)
{
  ndt_2d::ScanPtr scan(new ndt_2d::Scan(0));
  scan->setPose(ndt_2d::Pose2d(pose[0], pose[1], pose[2]));
  std::vector<ndt_2d::Point> points(npts);
  for (size_t i = 0; i < npts; ++i) {
    points[i].x = pts_xy[2 * i];
    points[i].y = pts_xy[2 * i + 1];
  }
  scan->setPoints(points);
  return scan;
}

// Scan::getBarycenterPose (scan.cpp:55-59, 72-91) of the reference's own Scan class
REF_API void ref_scan_barycenter(const double * pose3, const double * pts_xy, size_t n, double * out3)
{
  const ndt_2d::Pose2d b = make_scan(pose3, pts_xy, n)->getBarycenterPose();
  out3[0] = b.x;
  out3[1] = b.y;
  out3[2] = b.theta;
}

REF_API void * ref_ndt_create(double cell, double sx, double sy, double ox, double oy)
{
  return new ndt_2d::NDT(cell, sx, sy, ox, oy);
}
REF_API void ref_ndt_destroy(void * m) {delete static_cast<ndt_2d::NDT *>(m);}
REF_API int ref_ndt_get_index(void * m, double x, double y)
{
  return static_cast<ndt_2d::NDT *>(m)->getIndex(x, y);
}
REF_API void ref_ndt_add_scan(void * m, const double * pose, const double * pts_xy, size_t npts)
{
  static_cast<ndt_2d::NDT *>(m)->addScan(make_scan(pose, pts_xy, npts));
}
REF_API void ref_ndt_compute(void * m) {static_cast<ndt_2d::NDT *>(m)->compute();}
REF_API double ref_ndt_likelihood_point(void * m, double x, double y)
{
  Eigen::Vector2d p(x, y);
  return static_cast<ndt_2d::NDT *>(m)->likelihood(p);
}
REF_API double ref_ndt_likelihood_points(void * m, const double * pts_xy, size_t npts)
{
  std::vector<ndt_2d::Point> points(npts);
  for (size_t i = 0; i < npts; ++i) {
    points[i].x = pts_xy[2 * i];
    points[i].y = pts_xy[2 * i + 1];
  }
  return static_cast<ndt_2d::NDT *>(m)->likelihood(points);
}
REF_API double ref_ndt_likelihood_scan(
  void * m, const double * pose, const double * pts_xy, size_t npts)
{
  return static_cast<ndt_2d::NDT *>(m)->likelihood(make_scan(pose, pts_xy, npts));
}
REF_API void ref_ndt_grid(void * mv, double * info)
{
  auto * m = static_cast<ndt_2d::NDT *>(mv);
  info[0] = static_cast<double>(m->size_x_);
  info[1] = static_cast<double>(m->size_y_);
  info[2] = m->origin_x_;
  info[3] = m->origin_y_;
  info[4] = m->cell_size_;
}
REF_API void ref_ndt_dump_cells(void * mv, double * out)
{
  auto * m = static_cast<ndt_2d::NDT *>(mv);
  for (size_t i = 0; i < m->cells_.size(); ++i) {
    cell_get(m->cells_[i], out + 16 * i);
  }
}

// ---------------------------------------------------------------- matcher
struct RefMatcher
{
  rclcpp::Node node;
  ndt_2d::ScanMatcherNDT matcher;
};

REF_API void * ref_matcher_create(
  double ndt_resolution, double search_angular_resolution, double search_angular_size,
  double search_linear_resolution, double search_linear_size, int laser_max_beams,
  double range_max)
{
  auto * m = new RefMatcher();
  const std::string ns = "m";
  m->node.overrides[ns + ".ndt_resolution"] = ndt_resolution;
  m->node.overrides[ns + ".search_angular_resolution"] = search_angular_resolution;
  m->node.overrides[ns + ".search_angular_size"] = search_angular_size;
  m->node.overrides[ns + ".search_linear_resolution"] = search_linear_resolution;
  m->node.overrides[ns + ".search_linear_size"] = search_linear_size;
  m->node.overrides[ns + ".laser_max_beams"] = laser_max_beams;
  m->matcher.initialize(ns, &m->node, range_max);
  return m;
}
// A matcher initialised with NO overrides: exposes the plugin defaults.
REF_API void ref_matcher_defaults(double * out6)
{
  RefMatcher m;
  m.matcher.initialize("d", &m.node, 1.0);
  out6[0] = m.matcher.resolution_;
  out6[1] = m.matcher.angular_res_;
  out6[2] = m.matcher.angular_size_;
  out6[3] = m.matcher.linear_res_;
  out6[4] = m.matcher.linear_size_;
  out6[5] = static_cast<double>(m.matcher.laser_max_beams_);
}
REF_API void ref_matcher_destroy(void * m) {delete static_cast<RefMatcher *>(m);}
REF_API void ref_matcher_reset(void * m) {static_cast<RefMatcher *>(m)->matcher.reset();}
REF_API void * ref_matcher_ndt(void * m) {return static_cast<RefMatcher *>(m)->matcher.ndt_.get();}

REF_API void ref_matcher_add_scans(
  void * mv, size_t n_scans, const double * poses, const uint64_t * pt_offsets,
  const double * pts_xy)
{
  std::vector<ndt_2d::ScanPtr> scans;
  for (size_t k = 0; k < n_scans; ++k) {
    scans.push_back(make_scan(poses + 3 * k, pts_xy + 2 * pt_offsets[k],
      static_cast<size_t>(pt_offsets[k + 1] - pt_offsets[k])));
  }
  std::vector<ndt_2d::ScanPtr>::const_iterator b = scans.begin(), e = scans.end();
  static_cast<RefMatcher *>(mv)->matcher.addScans(b, e);
}

// out_delta is pre-filled by the caller; delta_written reports whether the
// reference changed it (scan_matcher_ndt.cpp:128-134).  all_scores is not
// available from the unmodified reference (must be NULL).
REF_API double ref_matcher_match_scan(
  void * mv, const double * pose3, const double * pts_xy, size_t npts,
  double * out_delta, int * delta_written, double * out_cov, double * all_scores)
{
  (void)all_scores;
  auto * m = static_cast<RefMatcher *>(mv);
  ndt_2d::ScanPtr scan = make_scan(pose3, pts_xy, npts);
  // sentinel pattern to detect "not written"
  const double sentinel = -12345.678;
  ndt_2d::Pose2d pose(sentinel, sentinel, sentinel);
  Eigen::Matrix3d cov;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {cov(i, j) = out_cov ? out_cov[i * 3 + j] : 0.0;}
  }
  const double score = m->matcher.matchScan(scan, pose, cov);
  const bool written = !(pose.x == sentinel && pose.y == sentinel && pose.theta == sentinel);
  if (delta_written) {*delta_written = written ? 1 : 0;}
  if (written && out_delta) {
    out_delta[0] = pose.x;
    out_delta[1] = pose.y;
    out_delta[2] = pose.theta;
  }
  if (out_cov) {
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {out_cov[i * 3 + j] = cov(i, j);}
    }
  }
  return score;
}

REF_API double ref_matcher_score_points(
  void * mv, const double * pts_xy, size_t npts, const double * pose3)
{
  std::vector<ndt_2d::Point> points(npts);
  for (size_t i = 0; i < npts; ++i) {
    points[i].x = pts_xy[2 * i];
    points[i].y = pts_xy[2 * i + 1];
  }
  return static_cast<RefMatcher *>(mv)->matcher.scorePoints(
    points, ndt_2d::Pose2d(pose3[0], pose3[1], pose3[2]));
}

REF_API double ref_matcher_score_scan(
  void * mv, const double * pose3, const double * pts_xy, size_t npts)
{
  return static_cast<RefMatcher *>(mv)->matcher.scoreScan(make_scan(pose3, pts_xy, npts));
}

// ---------------------------------------------------------------- angles
REF_API double ref_normalize_angle(double a) {return angles::normalize_angle(a);}
REF_API double ref_shortest_angular_distance(double f, double t)
{
  return angles::shortest_angular_distance(f, t);
}

// ---------------------------------------------------------------- KD tree
REF_API void ref_kd_leaf_counts(
  const double * poses, size_t N, const double * sizes3, uint64_t * leaf_counts)
{
  ndt_2d::KDTree<double> tree(sizes3[0], sizes3[1], sizes3[2], 1);
  for (size_t i = 0; i < N; ++i) {
    Eigen::Vector3d p(poses[3 * i], poses[3 * i + 1], poses[3 * i + 2]);
    double w = 1.0;
    tree.insert(p, w);
    leaf_counts[i] = tree.getLeafCount();
  }
}

// ---------------------------------------------------------------- filter
struct RefFilter
{
  ndt_2d::MotionModelPtr model;
  std::unique_ptr<ndt_2d::ParticleFilter> filter;
};

REF_API void * ref_pf_create(size_t min_particles, size_t max_particles, const double * alphas5)
{
  auto * f = new RefFilter();
  f->model = std::make_shared<ndt_2d::MotionModel>(
    alphas5[0], alphas5[1], alphas5[2], alphas5[3], alphas5[4]);
  f->filter.reset(new ndt_2d::ParticleFilter(min_particles, max_particles, f->model));
  return f;
}
REF_API void ref_pf_destroy(void * f) {delete static_cast<RefFilter *>(f);}

// Reseed both generators (they are seeded from std::random_device otherwise).
REF_API void ref_pf_seed(void * fv, uint32_t filter_seed, uint32_t motion_seed)
{
  auto * f = static_cast<RefFilter *>(fv);
  f->filter->gen_.seed(filter_seed);
  f->model->gen_.seed(motion_seed);
}
REF_API void ref_pf_set(void * fv, const double * particles, const double * weights, size_t P)
{
  auto * f = static_cast<RefFilter *>(fv);
  f->filter->particles_.resize(P);
  f->filter->weights_.resize(P);
  for (size_t i = 0; i < P; ++i) {
    f->filter->particles_[i] =
      Eigen::Vector3d(particles[3 * i], particles[3 * i + 1], particles[3 * i + 2]);
    f->filter->weights_[i] = weights[i];
  }
}
REF_API size_t ref_pf_size(void * fv) {return static_cast<RefFilter *>(fv)->filter->particles_.size();}
REF_API void ref_pf_get(void * fv, double * particles, double * weights)
{
  auto * f = static_cast<RefFilter *>(fv);
  for (size_t i = 0; i < f->filter->particles_.size(); ++i) {
    for (int j = 0; j < 3; ++j) {particles[3 * i + j] = f->filter->particles_[i](j);}
    weights[i] = f->filter->weights_[i];
  }
}
REF_API void ref_pf_stats(void * fv, double * mean3, double * cov9)
{
  auto * f = static_cast<RefFilter *>(fv);
  const Eigen::Vector3d m = f->filter->getMean();
  const Eigen::Matrix3d c = f->filter->getCovariance();
  for (int i = 0; i < 3; ++i) {
    mean3[i] = m(i);
    for (int j = 0; j < 3; ++j) {cov9[i * 3 + j] = c(i, j);}
  }
}
REF_API void ref_pf_set_cov(void * fv, const double * cov9)
{
  auto * f = static_cast<RefFilter *>(fv);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {f->filter->cov_(i, j) = cov9[i * 3 + j];}
  }
}
REF_API void ref_pf_update_statistics(void * fv)
{
  static_cast<RefFilter *>(fv)->filter->updateStatistics();
}
REF_API void ref_pf_init(void * fv, double x, double y, double th, double sx, double sy, double sth)
{
  static_cast<RefFilter *>(fv)->filter->init(x, y, th, sx, sy, sth);
}
REF_API void ref_pf_update(void * fv, double dx, double dy, double dth)
{
  static_cast<RefFilter *>(fv)->filter->update(dx, dy, dth);
}
// ParticleFilter::measure through the ScanMatcher interface
// (particle_filter.cpp:78-89), matcher = a ref_matcher_create() handle.
REF_API void ref_pf_measure(void * fv, void * mv, const double * pts_xy, size_t npts)
{
  auto * f = static_cast<RefFilter *>(fv);
  auto * m = static_cast<RefMatcher *>(mv);
  // non-owning shared_ptr: the matcher lives in RefMatcher
  ndt_2d::ScanMatcherPtr matcher(&m->matcher, [](ndt_2d::ScanMatcher *) {});
  const double pose[3] = {0, 0, 0};
  f->filter->measure(matcher, make_scan(pose, pts_xy, npts));
}
REF_API void ref_pf_resample(void * fv, double kld_err, double kld_z)
{
  static_cast<RefFilter *>(fv)->filter->resample(kld_err, kld_z);
}

// The uniform variates std::discrete_distribution<size_t> consumes from a
// std::mt19937 seeded with `seed` (libstdc++: one generate_canonical<double,53>
// per draw).  Lets the oracle / device resampler replay the reference's draws.
REF_API void ref_canonical_uniforms(uint32_t seed, size_t n, double * out)
{
  std::mt19937 gen(seed);
  for (size_t i = 0; i < n; ++i) {
    out[i] = std::generate_canonical<double, std::numeric_limits<double>::digits>(gen);
  }
}

// ---------------------------------------------------------------- timing aid
// Bounded sample of a large search for the CPU baseline: the UNMODIFIED
// reference matchScan run with a narrower search_angular_size (a legal
// parameter value) -- same per-candidate work as the full search.
REF_API uint64_t ref_matcher_candidate_count(void * mv)
{
  auto & m = static_cast<RefMatcher *>(mv)->matcher;
  uint64_t na = 0, nl = 0;
  for (double d = -m.angular_size_; d < m.angular_size_; d += m.angular_res_) {++na;}
  for (double d = -m.linear_size_; d < m.linear_size_; d += m.linear_res_) {++nl;}
  return na * nl * nl;
}

// ---------------------------------------------------------------- occupancy grid
REF_API void * ref_occ_create(double resolution, double occ_thresh)
{
  return new ndt_2d::OccupancyGrid(resolution, occ_thresh);
}
REF_API void ref_occ_destroy(void * g) {delete static_cast<ndt_2d::OccupancyGrid *>(g);}
// getMsg: info5 = {width, height, origin_x, origin_y, resolution (float)}; data copied to
// `data` if it has room (capacity cells); returns the number of cells.
REF_API size_t ref_occ_render(
  void * g, size_t n_scans, const double * poses, const uint64_t * offsets, const double * pts_xy,
  double * info5, int8_t * data, size_t capacity)
{
  std::vector<ndt_2d::ScanPtr> scans;
  for (size_t k = 0; k < n_scans; ++k) {
    scans.push_back(make_scan(poses + 3 * k, pts_xy + 2 * offsets[k], offsets[k + 1] - offsets[k]));
  }
  nav_msgs::msg::OccupancyGrid grid;
  static_cast<ndt_2d::OccupancyGrid *>(g)->getMsg(scans, grid);
  info5[0] = grid.info.width;
  info5[1] = grid.info.height;
  info5[2] = grid.info.origin.position.x;
  info5[3] = grid.info.origin.position.y;
  info5[4] = grid.info.resolution;
  if (data && capacity >= grid.data.size()) {
    memcpy(data, grid.data.data(), grid.data.size());
  }
  return grid.data.size();
}
