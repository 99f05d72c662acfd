// ndt_2d_b200::ScanMatcherNDT -- see include/ndt_2d_b200/scan_matcher_ndt.hpp.
#include <ndt_2d_b200/scan_matcher_ndt.hpp>

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace ndt_2d_b200
{

// std::vector<ndt_2d::Point> is passed to the C ABI as interleaved doubles
// (point.hpp:35-51: two doubles, no padding).
static_assert(sizeof(ndt_2d::Point) == 2 * sizeof(double), "ndt_2d::Point must be {double x, y}");
static_assert(std::is_standard_layout<ndt_2d::Point>::value, "ndt_2d::Point layout");

namespace
{

const double * xy(const std::vector<ndt_2d::Point> & points)
{
  return points.empty() ? nullptr : &points[0].x;
}

[[noreturn]] void fail(const char * where, int status)
{
  throw std::runtime_error(
          std::string("ndt_2d_b200::ScanMatcherNDT::") + where + ": libndt2d_b200 status " +
          std::to_string(status) + " (" + ndt2d_last_error() + "); there is no CPU fallback");
}

// Flattens scans into the add_scans layout: poses (3 per scan), offsets, points.
struct FlatScans
{
  std::vector<double> poses;
  std::vector<uint64_t> offsets;
  std::vector<double> points;
  void append(const ndt_2d::ScanPtr & scan)
  {
    const ndt_2d::Pose2d pose = scan->getPose();
    const std::vector<ndt_2d::Point> pts = scan->getPoints();
    poses.push_back(pose.x);
    poses.push_back(pose.y);
    poses.push_back(pose.theta);
    if (offsets.empty()) {offsets.push_back(0);}
    for (const auto & p : pts) {
      points.push_back(p.x);
      points.push_back(p.y);
    }
    offsets.push_back(points.size() / 2);
  }
  size_t n_scans() const {return poses.size() / 3;}
  const uint64_t * offsets_ptr()
  {
    if (offsets.empty()) {offsets.push_back(0);}
    return offsets.data();
  }
};

}  // namespace

ScanMatcherNDT::~ScanMatcherNDT()
{
  if (handle_) {ndt2d_matcher_destroy(handle_);}
}

void ScanMatcherNDT::setDevice(int device, void * cuda_stream)
{
  device_ = device;
  stream_ = cuda_stream;
}

void ScanMatcherNDT::setDevices(const std::vector<int> & devices)
{
  if (devices.size() > NDT2D_MAX_DEVICES) {
    throw std::invalid_argument("ndt_2d_b200::ScanMatcherNDT::setDevices: too many devices");
  }
  devices_ = devices;
}

void ScanMatcherNDT::require_handle(const char * where) const
{
  if (!handle_) {
    throw std::logic_error(
            std::string("ndt_2d_b200::ScanMatcherNDT::") + where + " called before initialize()");
  }
}

void ScanMatcherNDT::initialize(const std::string & name, rclcpp::Node * node, double range_max)
{
  // The same six parameters, names and defaults as the reference
  // (scan_matcher_ndt.cpp:37-44).
  ndt2d_default_params(&params_);
  params_.ndt_resolution = node->declare_parameter<double>(name + ".ndt_resolution", 0.25);
  params_.search_angular_resolution =
    node->declare_parameter<double>(name + ".search_angular_resolution", 0.0025);
  params_.search_angular_size = node->declare_parameter<double>(name + ".search_angular_size", 0.1);
  params_.search_linear_resolution =
    node->declare_parameter<double>(name + ".search_linear_resolution", 0.005);
  params_.search_linear_size = node->declare_parameter<double>(name + ".search_linear_size", 0.05);
  params_.laser_max_beams = node->declare_parameter<int>(name + ".laser_max_beams", 100);
  params_.range_max = range_max;
  params_.device = device_;
  params_.stream = stream_;
  // extension: "<name>.n_gpus" GPUs (0 .. n_gpus - 1) behind this matcher; setDevices() wins
  const int n_gpus = node->declare_parameter<int>(name + ".n_gpus", 1);
  std::vector<int> devices = devices_;
  if (devices.empty() && n_gpus > 1) {
    for (int d = 0; d < n_gpus && d < NDT2D_MAX_DEVICES; ++d) {devices.push_back(d);}
  }
  params_.n_devices = static_cast<int>(devices.size());
  for (size_t d = 0; d < devices.size(); ++d) {params_.devices[d] = devices[d];}
  if (handle_) {
    ndt2d_matcher_destroy(handle_);
    handle_ = nullptr;
  }
  const int rc = ndt2d_matcher_create(&params_, &handle_);
  if (rc != NDT2D_OK) {fail("initialize", rc);}
}

void ScanMatcherNDT::addScans(
  const std::vector<ndt_2d::ScanPtr>::const_iterator & begin,
  const std::vector<ndt_2d::ScanPtr>::const_iterator & end)
{
  require_handle("addScans");
  FlatScans flat;
  for (auto scan = begin; scan != end; ++scan) {flat.append(*scan);}
  const int rc = ndt2d_matcher_add_scans(
    handle_, flat.n_scans(), flat.poses.data(), flat.offsets_ptr(), flat.points.data());
  if (rc != NDT2D_OK) {fail("addScans", rc);}
}

double ScanMatcherNDT::matchScan(
  const ndt_2d::ScanPtr & scan, ndt_2d::Pose2d & pose, Eigen::Matrix3d & covariance) const
{
  require_handle("matchScan");
  const ndt_2d::Pose2d scan_pose = scan->getPose();
  const std::vector<ndt_2d::Point> points = scan->getPoints();
  const double pose3[3] = {scan_pose.x, scan_pose.y, scan_pose.theta};
  double delta[3] = {0.0, 0.0, 0.0}, cov[9], score = 0.0;
  int written = 0;
  const int rc = ndt2d_matcher_match_scan(
    handle_, pose3, xy(points), points.size(), delta, &written, cov, &score);
  if (rc == NDT2D_ERR_NO_MAP) {
    return 0.0;  // scan_matcher_ndt.cpp:80 -- pose and covariance untouched
  }
  if (rc != NDT2D_OK) {fail("matchScan", rc);}
  if (written) {
    // :130-133 -- the correction (a delta), only if some candidate scored below zero
    pose.x = delta[0];
    pose.y = delta[1];
    pose.theta = delta[2];
  }
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {covariance(r, c) = cov[3 * r + c];}
  }
  return score;
}

double ScanMatcherNDT::scoreScan(const ndt_2d::ScanPtr & scan) const
{
  return scorePoints(scan->getPoints(), scan->getPose());
}

double ScanMatcherNDT::scorePoints(
  const std::vector<ndt_2d::Point> & points, const ndt_2d::Pose2d & pose) const
{
  require_handle("scorePoints");
  const double pose3[3] = {pose.x, pose.y, pose.theta};
  double score = 0.0;
  const int rc = ndt2d_matcher_score_points(handle_, xy(points), points.size(), pose3, &score);
  if (rc == NDT2D_ERR_NO_MAP) {return 0.0;}  // scan_matcher_ndt.cpp:159
  if (rc != NDT2D_OK) {fail("scorePoints", rc);}
  return score;
}

void ScanMatcherNDT::scorePoses(
  const std::vector<ndt_2d::Point> & points, const double * poses3, size_t n_poses,
  double * out_scores) const
{
  require_handle("scorePoses");
  const int rc = ndt2d_matcher_score_poses(
    handle_, xy(points), points.size(), poses3, n_poses, out_scores);
  if (rc == NDT2D_ERR_NO_MAP) {return;}  // every score 0.0, as n calls of scorePoints would give
  if (rc != NDT2D_OK) {fail("scorePoses", rc);}
}

void ScanMatcherNDT::reset()
{
  require_handle("reset");
  const int rc = ndt2d_matcher_reset(handle_);
  if (rc != NDT2D_OK) {fail("reset", rc);}
}

std::vector<ScanMatcherNDT::BatchResult> ScanMatcherNDT::matchScanBatch(
  const std::vector<std::vector<ndt_2d::ScanPtr>> & maps,
  const std::vector<ndt_2d::ScanPtr> & scans)
{
  require_handle("matchScanBatch");
  if (maps.size() != scans.size()) {
    throw std::invalid_argument("matchScanBatch: one query scan per map window");
  }
  const size_t n_jobs = maps.size();
  FlatScans map_flat, query_flat;
  std::vector<uint64_t> job_offsets(1, 0);
  for (size_t j = 0; j < n_jobs; ++j) {
    for (const auto & s : maps[j]) {map_flat.append(s);}
    job_offsets.push_back(map_flat.n_scans());
    query_flat.append(scans[j]);
  }
  std::vector<double> delta(3 * n_jobs, 0.0), cov(9 * n_jobs, 0.0), score(n_jobs, 0.0);
  std::vector<int> written(n_jobs, 0);
  const int rc = ndt2d_matcher_match_scan_batch(
    handle_, n_jobs, job_offsets.data(), map_flat.poses.data(), map_flat.offsets_ptr(),
    map_flat.points.data(), query_flat.poses.data(), query_flat.offsets_ptr(),
    query_flat.points.data(), delta.data(), written.data(), cov.data(), score.data());
  if (rc != NDT2D_OK) {fail("matchScanBatch", rc);}
  std::vector<BatchResult> out(n_jobs);
  for (size_t j = 0; j < n_jobs; ++j) {
    out[j].score = score[j];
    out[j].pose_written = written[j] != 0;
    out[j].pose = ndt_2d::Pose2d(delta[3 * j], delta[3 * j + 1], delta[3 * j + 2]);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) {out[j].covariance(r, c) = cov[9 * j + 3 * r + c];}
    }
  }
  return out;
}

std::vector<ScanMatcherNDT::LoopClosure> ScanMatcherNDT::closeLoop(
  const std::vector<ndt_2d::ScanPtr> & graph_scans, const ndt_2d::ScanPtr & scan,
  const std::vector<size_t> & candidates, size_t rolling, size_t search_limit,
  double typical_response, size_t * n_batches)
{
  require_handle("closeLoop");
  FlatScans graph;
  for (const auto & s : graph_scans) {graph.append(s);}
  std::vector<uint64_t> cand(candidates.begin(), candidates.end());
  const ndt_2d::Pose2d pose0 = scan->getPose();
  double query_pose[3] = {pose0.x, pose0.y, pose0.theta};
  const std::vector<ndt_2d::Point> points = scan->getPoints();
  // (a limit of 0 means no limit: the reference's size_t countdown wraps, ndt_mapper.cpp:619,671)
  const size_t cap = std::max<size_t>(1, search_limit ? std::min(search_limit, cand.size()) : cand.size());
  std::vector<uint64_t> out_cand(cap);
  std::vector<double> out_score(cap), out_pose(3 * cap), out_cov(9 * cap);
  std::vector<int> out_acc(cap);
  size_t n = 0, nb = 0;
  const int rc = ndt2d_matcher_close_loop(
    handle_, graph.n_scans(), graph.poses.data(), graph.offsets_ptr(), graph.points.data(),
    cand.data(), cand.size(), rolling, search_limit, typical_response, query_pose, xy(points),
    points.size(), out_cand.data(), out_score.data(), out_acc.data(), out_pose.data(),
    out_cov.data(), &n, &nb);
  if (rc != NDT2D_OK) {fail("closeLoop", rc);}
  if (n_batches) {*n_batches = nb;}
  std::vector<LoopClosure> out(n);
  bool any = false;
  for (size_t k = 0; k < n; ++k) {
    out[k].candidate = static_cast<size_t>(out_cand[k]);
    out[k].score = out_score[k];
    out[k].accepted = out_acc[k] != 0;
    out[k].scan_pose = ndt_2d::Pose2d(out_pose[3 * k], out_pose[3 * k + 1], out_pose[3 * k + 2]);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) {out[k].covariance(r, c) = out_cov[9 * k + 3 * r + c];}
    }
    any = any || out[k].accepted;
  }
  if (any) {
    scan->setPose(ndt_2d::Pose2d(query_pose[0], query_pose[1], query_pose[2]));  // :655
  }
  return out;
}

std::vector<size_t> ScanMatcherNDT::findNearest(
  const std::vector<ndt_2d::ScanPtr> & graph_scans, const ndt_2d::ScanPtr & scan, double dist,
  int limit_scan_index, bool use_barycenter) const
{
  std::vector<double> xy(2 * graph_scans.size());
  for (size_t i = 0; i < graph_scans.size(); ++i) {
    // GraphAdapter::kdtree_get_pt (graph.hpp:97-108)
    const ndt_2d::Pose2d p =
      use_barycenter ? graph_scans[i]->getBarycenterPose() : graph_scans[i]->getPose();
    xy[2 * i] = p.x;
    xy[2 * i + 1] = p.y;
  }
  const ndt_2d::Pose2d q = use_barycenter ? scan->getBarycenterPose() : scan->getPose();
  const double query[2] = {q.x, q.y};
  std::vector<uint64_t> idx(std::max<size_t>(1, graph_scans.size()));
  size_t n = 0;
  const int rc = ndt2d_find_nearest(
    device_, xy.data(), graph_scans.size(), limit_scan_index, query, dist, idx.data(), nullptr,
    idx.size(), &n);
  if (rc != NDT2D_OK) {fail("findNearest", rc);}
  return std::vector<size_t>(idx.begin(), idx.begin() + std::min(n, idx.size()));
}

}  // namespace ndt_2d_b200

#include <pluginlib/class_list_macros.hpp>
PLUGINLIB_EXPORT_CLASS(ndt_2d_b200::ScanMatcherNDT, ndt_2d::ScanMatcher)
