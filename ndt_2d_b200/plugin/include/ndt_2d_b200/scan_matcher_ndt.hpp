// ndt_2d_b200::ScanMatcherNDT -- the drop-in pluginlib class.
//
// Implements the reference's abstract plugin interface ndt_2d::ScanMatcher
// (include/ndt_2d/scan_matcher.hpp:42-91) on top of the C ABI of
// libndt2d_b200.so (include/ndt2d_b200.h).  It replaces the reference class
// ndt_2d::ScanMatcherNDT (include/ndt_2d/scan_matcher_ndt.hpp,
// src/scan_matcher_ndt.cpp): same six ROS parameters with the same defaults,
// same call semantics, same return values.  plugins.xml registers it under the
// reference's lookup name "ndt_2d::ScanMatcherNDT", so a node configured with
// scan_matcher_type = ndt_2d::ScanMatcherNDT loads this class unchanged.
//
// There is no CPU path behind it: every method that needs the device throws
// std::runtime_error if the C ABI reports anything but success / "no map".
#ifndef NDT_2D_B200__SCAN_MATCHER_NDT_HPP_
#define NDT_2D_B200__SCAN_MATCHER_NDT_HPP_

#include <Eigen/Core>

#include <cstddef>
#include <string>
#include <vector>

#include <ndt_2d/scan_matcher.hpp>

#include "ndt2d_b200.h"

namespace ndt_2d_b200
{

class ScanMatcherNDT : public ndt_2d::ScanMatcher
{
public:
  ScanMatcherNDT() = default;
  virtual ~ScanMatcherNDT();
  ScanMatcherNDT(const ScanMatcherNDT &) = delete;
  ScanMatcherNDT & operator=(const ScanMatcherNDT &) = delete;

  // scan_matcher_ndt.cpp:35-47
  void initialize(const std::string & name, rclcpp::Node * node, double range_max) override;

  // scan_matcher_ndt.cpp:49-74
  void addScans(
    const std::vector<ndt_2d::ScanPtr>::const_iterator & begin,
    const std::vector<ndt_2d::ScanPtr>::const_iterator & end) override;

  // scan_matcher_ndt.cpp:76-149
  double matchScan(
    const ndt_2d::ScanPtr & scan, ndt_2d::Pose2d & pose,
    Eigen::Matrix3d & covariance) const override;

  // scan_matcher_ndt.cpp:151-154
  double scoreScan(const ndt_2d::ScanPtr & scan) const override;

  // scan_matcher_ndt.cpp:156-178
  double scorePoints(
    const std::vector<ndt_2d::Point> & points, const ndt_2d::Pose2d & pose) const override;

  // scan_matcher_ndt.cpp:180-183
  void reset() override;

  // ---- extensions (not part of ndt_2d::ScanMatcher) ----------------------

  // One scorePoints per pose in a single launch: the body of
  // ParticleFilter::measure's loop (particle_filter.cpp:81-87).
  // poses3: n * {x, y, theta}; out_scores: n doubles.
  void scorePoses(
    const std::vector<ndt_2d::Point> & points, const double * poses3, size_t n_poses,
    double * out_scores) const;

  // The loop-closure inner loop (ndt_mapper.cpp:619-671) for several candidate
  // windows in one submission: job j = reset + addScans(maps[j]) + matchScan(scan).
  struct BatchResult
  {
    double score;
    bool pose_written;
    ndt_2d::Pose2d pose;
    Eigen::Matrix3d covariance;
  };
  std::vector<BatchResult> matchScanBatch(
    const std::vector<std::vector<ndt_2d::ScanPtr>> & maps,
    const std::vector<ndt_2d::ScanPtr> & scans);

  // The inner loop of Mapper::loopClosureThread (ndt_mapper.cpp:619-671) for one new scan,
  // sequential semantics kept (an accepted match moves `scan`'s pose before the next
  // candidate is matched), evaluated as speculative batches.  `graph_scans` is
  // graph_->scans, `candidates` the result of Graph::findNearest, in its order.  The
  // caller still owns the graph: it adds a constraint per accepted entry
  // (makeConstraint(graph_scans[candidate], scan, covariance), :658-661).
  struct LoopClosure
  {
    size_t candidate;
    double score;
    bool accepted;
    ndt_2d::Pose2d scan_pose;       // pose of `scan` after this candidate
    Eigen::Matrix3d covariance;
  };
  std::vector<LoopClosure> closeLoop(
    const std::vector<ndt_2d::ScanPtr> & graph_scans, const ndt_2d::ScanPtr & scan,
    const std::vector<size_t> & candidates, size_t rolling, size_t search_limit,
    double typical_response, size_t * n_batches = nullptr);

  // Graph::findNearest(scan, dist, limit_scan_index) (graph.cpp:167-189) over
  // graph_->scans: the candidate list closeLoop takes, nearest first.  `dist` is a squared
  // radius, as nanoflann's radiusSearch takes it; use_barycenter = Graph::use_barycenter_.
  std::vector<size_t> findNearest(
    const std::vector<ndt_2d::ScanPtr> & graph_scans, const ndt_2d::ScanPtr & scan, double dist,
    int limit_scan_index = -1, bool use_barycenter = false) const;

  // CUDA device / stream selection, before initialize() (defaults: current device,
  // a stream owned by the handle).
  void setDevice(int device, void * cuda_stream = nullptr);

  // Several GPUs of this process behind this one matcher (ndt2d_params.n_devices), before
  // initialize(): the model and the query scan are replicated, a large matchScan (a global
  // search) scores theta slices r, r + N, ... on device r and the partial records are
  // exchanged GPU to GPU by the searches' last kernels; local matches stay on devices[0].
  // Also settable without code: the extra ROS parameter "<name>.n_gpus" (default 1) uses
  // devices 0 .. n_gpus - 1.
  void setDevices(const std::vector<int> & devices);

  ndt2d_matcher * handle() const {return handle_;}
  const ndt2d_params & params() const {return params_;}

private:
  void require_handle(const char * where) const;

  ndt2d_params params_{};
  int device_ = -1;
  void * stream_ = nullptr;
  std::vector<int> devices_;
  ndt2d_matcher * handle_ = nullptr;  // device state lives behind the handle, so the
                                      // const methods of the interface can use it
};

}  // namespace ndt_2d_b200

#endif  // NDT_2D_B200__SCAN_MATCHER_NDT_HPP_
