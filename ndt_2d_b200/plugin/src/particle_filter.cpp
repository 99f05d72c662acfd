// ndt_2d_b200::ParticleFilter -- see include/ndt_2d_b200/particle_filter.hpp.
#include <ndt_2d_b200/particle_filter.hpp>

#include <cmath>
#include <stdexcept>
#include <string>

#include <ndt_2d_b200/scan_matcher_ndt.hpp>

namespace ndt_2d_b200
{

namespace
{
void check(const char * where, int status)
{
  if (status != NDT2D_OK) {
    throw std::runtime_error(
            std::string("ndt_2d_b200::ParticleFilter::") + where + ": libndt2d_b200 status " +
            std::to_string(status) + " (" + ndt2d_last_error() + "); there is no CPU fallback");
  }
}
}  // namespace

ParticleFilter::ParticleFilter(
  size_t min_particles, size_t max_particles, MotionModelPtr & motion_model, uint64_t seed,
  int device, void * cuda_stream)
: motion_model_(motion_model), seed_(seed)
{
  check("ParticleFilter", ndt2d_filter_create(min_particles, max_particles, device, cuda_stream,
    &handle_));
}

ParticleFilter::~ParticleFilter()
{
  if (handle_) {ndt2d_filter_destroy(handle_);}
}

void ParticleFilter::init(
  const double x, const double y, const double theta,
  const double sigma_x, const double sigma_y, const double sigma_theta)
{
  check("init", ndt2d_filter_init(handle_, x, y, theta, sigma_x, sigma_y, sigma_theta, next_seed()));
}

void ParticleFilter::update(const double dx, const double dy, const double dth)
{
  check("update", ndt2d_filter_update(handle_, dx, dy, dth, motion_model_->alphas(), next_seed()));
}

void ParticleFilter::measure(const ndt_2d::ScanMatcherPtr & matcher, const ndt_2d::ScanPtr & scan)
{
  auto * ours = dynamic_cast<ScanMatcherNDT *>(matcher.get());
  if (!ours || !ours->handle()) {
    throw std::invalid_argument(
            "ndt_2d_b200::ParticleFilter::measure needs an initialised ndt_2d_b200::ScanMatcherNDT: "
            "particles are scored on the device, there is no host scoring path");
  }
  const std::vector<ndt_2d::Point> points = scan->getPoints();
  const int rc = ndt2d_filter_measure(
    handle_, ours->handle(), points.empty() ? nullptr : &points[0].x, points.size());
  check("measure", rc);
}

void ParticleFilter::resample(const double kld_err, const double kld_z)
{
  check("resample", ndt2d_filter_resample(handle_, kld_err, kld_z, nullptr, 0, next_seed()));
}

void ParticleFilter::resample(
  const double kld_err, const double kld_z, const std::vector<double> & uniforms)
{
  check("resample", ndt2d_filter_resample(handle_, kld_err, kld_z, uniforms.data(), uniforms.size(),
    0));
}

Eigen::Vector3d ParticleFilter::getMean()
{
  double mean[3];
  check("getMean", ndt2d_filter_stats(handle_, mean, nullptr));
  return Eigen::Vector3d(mean[0], mean[1], mean[2]);
}

Eigen::Matrix3d ParticleFilter::getCovariance()
{
  double cov[9];
  check("getCovariance", ndt2d_filter_stats(handle_, nullptr, cov));
  Eigen::Matrix3d out;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {out(r, c) = cov[3 * r + c];}
  }
  return out;
}

size_t ParticleFilter::size() const
{
  size_t n = 0;
  check("size", ndt2d_filter_size(handle_, &n));
  return n;
}

void ParticleFilter::setParticles(
  const std::vector<double> & particles3, const std::vector<double> & weights)
{
  if (particles3.size() != 3 * weights.size()) {
    throw std::invalid_argument("setParticles: 3 doubles per particle, one weight each");
  }
  check("setParticles", ndt2d_filter_set_particles(handle_, particles3.data(), weights.data(),
    weights.size()));
}

void ParticleFilter::getParticles(std::vector<double> & particles3, std::vector<double> & weights) const
{
  const size_t n = size();
  particles3.assign(3 * n, 0.0);
  weights.assign(n, 0.0);
  check("getParticles", ndt2d_filter_get_particles(handle_, particles3.data(), weights.data()));
}

void ParticleFilter::getMsg(geometry_msgs::msg::PoseArray & msg)
{
  std::vector<double> p, w;
  getParticles(p, w);
  msg.poses.reserve(w.size());
  for (size_t i = 0; i < w.size(); ++i) {
    geometry_msgs::msg::Pose pose;
    pose.position.x = p[3 * i];
    pose.position.y = p[3 * i + 1];
    pose.orientation.z = sin(p[3 * i + 2] / 2.0);
    pose.orientation.w = cos(p[3 * i + 2] / 2.0);
    msg.poses.push_back(pose);
  }
}

}  // namespace ndt_2d_b200
