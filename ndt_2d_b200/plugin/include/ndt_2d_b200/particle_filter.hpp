// ndt_2d_b200::ParticleFilter / MotionModel -- device-resident replacement of
// ndt_2d::ParticleFilter (include/ndt_2d/particle_filter.hpp:45-115,
// src/particle_filter.cpp) with the same public API.  The particle set stays on
// the GPU across update -> measure -> resample; only mean / covariance (and the
// PoseArray for visualisation) come back to the host.
//
// ndt_2d::MotionModel keeps its alphas private and owns a host RNG
// (motion_model.hpp:48-67), so the device filter takes its own MotionModel class
// with the same constructor; swapping the two `using` lines in the node is the
// whole integration.
#ifndef NDT_2D_B200__PARTICLE_FILTER_HPP_
#define NDT_2D_B200__PARTICLE_FILTER_HPP_

#include <Eigen/Core>

#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#include <geometry_msgs/msg/pose_array.hpp>
#include <ndt_2d/scan_matcher.hpp>

#include "ndt2d_b200.h"

namespace ndt_2d_b200
{

// Same constructor as ndt_2d::MotionModel (motion_model.cpp:39-43): odom_alpha1..5.
class MotionModel
{
public:
  MotionModel(double a1, double a2, double a3, double a4, double a5)
  : alphas_{a1, a2, a3, a4, a5} {}
  const double * alphas() const {return alphas_;}

private:
  double alphas_[5];
};
using MotionModelPtr = std::shared_ptr<MotionModel>;

class ParticleFilter
{
public:
  // particle_filter.cpp:36-51.  `seed` starts the counter-based device generator
  // (the reference seeds an mt19937 from std::random_device; only the
  // distributions can agree).
  ParticleFilter(
    size_t min_particles, size_t max_particles, MotionModelPtr & motion_model,
    uint64_t seed = 0x9E3779B97F4A7C15ull, int device = -1, void * cuda_stream = nullptr);
  ~ParticleFilter();
  ParticleFilter(const ParticleFilter &) = delete;
  ParticleFilter & operator=(const ParticleFilter &) = delete;

  // particle_filter.cpp:53-69
  void init(
    const double x, const double y, const double theta,
    const double sigma_x, const double sigma_y, const double sigma_theta);
  // particle_filter.cpp:71-76
  void update(const double dx, const double dy, const double dth);
  // particle_filter.cpp:78-89.  `matcher` must be an ndt_2d_b200::ScanMatcherNDT
  // (std::invalid_argument otherwise: there is no host scoring path).
  void measure(const ndt_2d::ScanMatcherPtr & matcher, const ndt_2d::ScanPtr & scan);
  // particle_filter.cpp:91-137
  void resample(const double kld_err, const double kld_z);
  // particle_filter.cpp:139-147
  Eigen::Vector3d getMean();
  Eigen::Matrix3d getCovariance();
  // particle_filter.cpp:149-161
  void getMsg(geometry_msgs::msg::PoseArray & msg);

  // ---- extensions ---------------------------------------------------------
  size_t size() const;
  // resample with caller-provided uniform variates (one per draw): the exact
  // replay hook used by the parity tests
  void resample(const double kld_err, const double kld_z, const std::vector<double> & uniforms);
  void setParticles(const std::vector<double> & particles3, const std::vector<double> & weights);
  void getParticles(std::vector<double> & particles3, std::vector<double> & weights) const;
  ndt2d_filter * handle() const {return handle_;}

private:
  uint64_t next_seed() {return seed_ += 0x9E3779B97F4A7C15ull;}

  MotionModelPtr motion_model_;
  uint64_t seed_;
  ndt2d_filter * handle_ = nullptr;
};

}  // namespace ndt_2d_b200

#endif  // NDT_2D_B200__PARTICLE_FILTER_HPP_
