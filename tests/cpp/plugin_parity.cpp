// plugin_parity.cpp -- drives the REFERENCE plugin and the B200 plugin through the
// same abstract interface (ndt_2d::ScanMatcher, scan_matcher.hpp:42-91) on identical
// synthetic inputs, the way Mapper::laserCallback does (ndt_mapper.cpp:508-515), and
// ndt_2d::ParticleFilter next to ndt_2d_b200::ParticleFilter.
//
// TEST INFRASTRUCTURE.  Built by oracle/Makefile (target plugin_parity) only where
// /root/reference exists: the reference's sources are compiled in place, unmodified,
// against oracle/ref_shim; the binary lands in oracle/_ref/ and travels to the GPU box.
//
//   plugin_parity            full comparison (needs a CUDA device), exit 0 = parity
//   plugin_parity --no-gpu   checks that the B200 plugin fails loudly without a device
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#define private public
#include <ndt_2d/particle_filter.hpp>
#undef private
#include <ndt_2d/motion_model.hpp>
#include <ndt_2d/scan_matcher_ndt.hpp>

#include <ndt_2d/occupancy_grid.hpp>
#include <ndt_2d_b200/occupancy_grid.hpp>
#include <ndt_2d_b200/particle_filter.hpp>
#include <ndt_2d_b200/scan_matcher_ndt.hpp>
#include <ndt2d_synth.h>

namespace
{

int g_failures = 0;

void expect(bool ok, const char * what)
{
  if (!ok) {
    ++g_failures;
    std::printf("FAIL: %s\n", what);
  }
}

bool close_rel(double a, double b, double rtol, double atol = 0.0)
{
  if (std::isnan(a) || std::isnan(b)) {return std::isnan(a) && std::isnan(b);}
  return std::fabs(a - b) <= atol + rtol * std::fabs(b);
}

const double kPi = 3.14159265358979323846;

struct World
{
  std::vector<double> rects;
  int n_rects = 60;
  double arena = 100.0;
};

std::vector<ndt_2d::ScanPtr> make_scans(
  const World & w, const std::vector<double> & poses3, int beams, double range_max, uint64_t seed)
{
  const size_t n = poses3.size() / 3;
  std::vector<uint64_t> offsets(n + 1);
  std::vector<double> pts(2 * static_cast<size_t>(beams) * n);
  ndt2d_synth_scans(w.rects.data(), w.n_rects, w.arena, poses3.data(), n, beams, range_max, 0.01,
    seed, offsets.data(), pts.data());
  std::vector<ndt_2d::ScanPtr> out;
  for (size_t k = 0; k < n; ++k) {
    ndt_2d::ScanPtr scan(new ndt_2d::Scan(k));
    scan->setPose(ndt_2d::Pose2d(poses3[3 * k], poses3[3 * k + 1], poses3[3 * k + 2]));
    std::vector<ndt_2d::Point> points;
    for (uint64_t i = offsets[k]; i < offsets[k + 1]; ++i) {
      points.emplace_back(pts[2 * i], pts[2 * i + 1]);
    }
    scan->setPoints(points);
    out.push_back(scan);
  }
  return out;
}

void compare_match(
  const ndt_2d::ScanMatcherPtr & ref, const ndt_2d::ScanMatcherPtr & dev, const ndt_2d::ScanPtr & scan,
  const char * label)
{
  // both start from a default pose and an "uninitialised" covariance, as the node's callers do
  ndt_2d::Pose2d pr, pd;
  Eigen::Matrix3d cr, cd;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {cr(r, c) = cd(r, c) = 7.0;}
  }
  const double sr = ref->matchScan(scan, pr, cr);
  const double sd = dev->matchScan(scan, pd, cd);
  std::printf("%s: matchScan ref %.12g dev %.12g  pose ref (%.6f %.6f %.6f) dev (%.6f %.6f %.6f)\n",
    label, sr, sd, pr.x, pr.y, pr.theta, pd.x, pd.y, pd.theta);
  expect(close_rel(sd, sr, 1e-5, 1e-30), "matchScan score within 1e-5");
  expect(pr.x == pd.x && pr.y == pd.y && pr.theta == pd.theta, "matchScan pose identical");
  double cmax = 0.0;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {
      if (std::isfinite(cr(r, c))) {cmax = std::fmax(cmax, std::fabs(cr(r, c)));}
    }
  }
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {
      expect(close_rel(cd(r, c), cr(r, c), 1e-5, 1e-5 * cmax), "matchScan covariance within 1e-5");
    }
  }
  const double qr = ref->scoreScan(scan), qd = dev->scoreScan(scan);
  expect(close_rel(qd, qr, 1e-5, 1e-30), "scoreScan within 1e-5");
}

int run_no_gpu()
{
  rclcpp::Node node;
  ndt_2d_b200::ScanMatcherNDT m;
  bool threw = false;
  try {
    m.initialize("local_scan_matcher", &node, 10.0);
  } catch (const std::runtime_error & e) {
    threw = true;
    std::printf("initialize without a device: %s\n", e.what());
  }
  if (ndt2d_device_count() > 0) {
    std::printf("a CUDA device is visible; --no-gpu check not applicable\n");
    return 0;
  }
  return threw ? 0 : 1;
}

}  // namespace

int main(int argc, char ** argv)
{
  if (argc > 1 && !std::strcmp(argv[1], "--no-gpu")) {return run_no_gpu();}

  World w;
  w.rects.resize(4 * w.n_rects);
  ndt2d_synth_world(42, w.arena, w.n_rects, 1.0, 8.0, w.rects.data());

  // ---- mapping: rolling window of 10 scans + the 11th matched against it -----------
  std::vector<double> poses;
  for (int k = 0; k < 11; ++k) {
    poses.push_back(50.0 + 0.2 * k);
    poses.push_back(50.0);
    poses.push_back(0.01 * ((k * 7) % 11 - 5));
  }
  std::vector<ndt_2d::ScanPtr> scans = make_scans(w, poses, 360, 10.0, 43);
  ndt_2d::ScanPtr query = scans.back();
  scans.pop_back();
  const ndt_2d::Pose2d truth = query->getPose();
  query->setPose(ndt_2d::Pose2d(truth.x - 0.12, truth.y + 0.07, truth.theta - 0.06));

  rclcpp::Node node;  // the params file: config 1 of BASELINE.json
  node.overrides["local_scan_matcher.ndt_resolution"] = 0.25;
  node.overrides["local_scan_matcher.search_linear_size"] = 0.25;
  node.overrides["local_scan_matcher.search_linear_resolution"] = 0.05;
  node.overrides["local_scan_matcher.search_angular_size"] = 0.25;
  node.overrides["local_scan_matcher.search_angular_resolution"] = 0.0025;
  node.overrides["local_scan_matcher.laser_max_beams"] = 360;

  ndt_2d::ScanMatcherPtr ref(new ndt_2d::ScanMatcherNDT());
  ndt_2d::ScanMatcherPtr dev(new ndt_2d_b200::ScanMatcherNDT());
  for (auto & m : {ref, dev}) {
    m->initialize("local_scan_matcher", &node, 10.0);
  }

  // no map yet: 0.0 and outputs untouched (scan_matcher_ndt.cpp:80,159)
  {
    ndt_2d::Pose2d p(1.0, 2.0, 3.0);
    Eigen::Matrix3d c;
    for (int r = 0; r < 3; ++r) {for (int q = 0; q < 3; ++q) {c(r, q) = 7.0;}}
    const double s = dev->matchScan(query, p, c);
    expect(s == 0.0 && p.x == 1.0 && p.y == 2.0 && p.theta == 3.0 && c(1, 1) == 7.0,
      "no map: matchScan returns 0.0 and leaves pose / covariance untouched");
    expect(dev->scoreScan(query) == 0.0 && ref->scoreScan(query) == 0.0, "no map: scoreScan 0.0");
  }

  for (auto & m : {ref, dev}) {
    m->reset();
    m->addScans(scans.begin(), scans.end());   // ndt_mapper.cpp:508-509
  }
  compare_match(ref, dev, query, "config1");

  // a scan nowhere near the map: every candidate scores 0, pose untouched, NaN covariance
  {
    ndt_2d::ScanPtr far(new ndt_2d::Scan(99));
    far->setPose(ndt_2d::Pose2d(500.0, 500.0, 0.0));
    far->setPoints(query->getPoints());
    compare_match(ref, dev, far, "far scan");
  }

  // plugin defaults (no overrides): 21 x 21 x 80 candidates, 100 beams
  {
    rclcpp::Node defaults;
    ndt_2d::ScanMatcherPtr r2(new ndt_2d::ScanMatcherNDT()), d2(new ndt_2d_b200::ScanMatcherNDT());
    for (auto & m : {r2, d2}) {
      m->initialize("global_scan_matcher", &defaults, 10.0);
      m->addScans(scans.begin(), scans.end());
    }
    ndt_2d::ScanPtr q2(new ndt_2d::Scan(100));
    q2->setPose(ndt_2d::Pose2d(truth.x - 0.02, truth.y + 0.03, truth.theta - 0.04));
    q2->setPoints(query->getPoints());
    compare_match(r2, d2, q2, "plugin defaults");
  }

  // ---- loop closure: the reference's sequential loop (ndt_mapper.cpp:619-671, written out
  // here against the reference plugin) next to ScanMatcherNDT::closeLoop
  {
    const std::vector<size_t> candidates = {3, 8, 0, 9, 6, 1};
    const size_t rolling = 8, limit = 5;
    const double typical = -0.05;
    ndt_2d::ScanPtr sr(new ndt_2d::Scan(200)), sd(new ndt_2d::Scan(201));
    for (auto & s : {sr, sd}) {
      s->setPose(query->getPose());
      s->setPoints(query->getPoints());
    }
    struct Seq {size_t i; double score; bool accepted; ndt_2d::Pose2d pose;};
    std::vector<Seq> seq;
    size_t left = limit;
    for (auto i : candidates) {
      if (scans[i]->getPoints().empty()) {continue;}
      const size_t begin_idx = (i > 0) ? i - 1 : i, end_idx = (i < rolling) ? i + 1 : i;
      ref->reset();
      ref->addScans(scans.begin() + begin_idx, scans.begin() + end_idx);
      ndt_2d::Pose2d correction;
      Eigen::Matrix3d covariance;
      const double score = ref->matchScan(sr, correction, covariance);
      const bool accept = std::isfinite(score) && (score < typical);
      if (accept) {
        correction.x += sr->getPose().x;
        correction.y += sr->getPose().y;
        correction.theta += sr->getPose().theta;
        sr->setPose(correction);
      }
      seq.push_back({i, score, accept, sr->getPose()});
      if (--left == 0) {break;}
    }
    size_t n_batches = 0;
    auto * ours = dynamic_cast<ndt_2d_b200::ScanMatcherNDT *>(dev.get());
    const auto got = ours->closeLoop(scans, sd, candidates, rolling, limit, typical, &n_batches);
    expect(got.size() == seq.size(), "closeLoop processes the same candidates");
    size_t accepted = 0;
    for (size_t k = 0; k < got.size() && k < seq.size(); ++k) {
      expect(got[k].candidate == seq[k].i && got[k].accepted == seq[k].accepted, "closeLoop decisions");
      expect(close_rel(got[k].score, seq[k].score, 1e-5, 1e-30), "closeLoop scores within 1e-5");
      expect(got[k].scan_pose.x == seq[k].pose.x && got[k].scan_pose.y == seq[k].pose.y &&
        got[k].scan_pose.theta == seq[k].pose.theta, "closeLoop poses identical");
      accepted += seq[k].accepted ? 1 : 0;
    }
    expect(sd->getPose().x == sr->getPose().x && sd->getPose().theta == sr->getPose().theta,
      "closeLoop leaves the scan at the reference's pose");
    std::printf("closeLoop: %zu candidates, %zu accepted, %zu batch submissions\n", got.size(),
      accepted, n_batches);
    // restore the rolling-window model for the checks below
    for (auto & m : {ref, dev}) {
      m->reset();
      m->addScans(scans.begin(), scans.end());
    }
  }

  // ---- candidate selection: Graph::findNearest (graph.cpp:167-189) -- nanoflann is not
  // installed here, so the expectation is its radius search written out (squared L2 over the
  // reference's own Scan poses / barycenters, dist < radius, nearest first)
  {
    auto * ours = dynamic_cast<ndt_2d_b200::ScanMatcherNDT *>(dev.get());
    for (const bool bary : {false, true}) {
      for (const int limit : {-1, 7}) {
        const double radius_sq = 0.7;
        const ndt_2d::Pose2d q = bary ? query->getBarycenterPose() : query->getPose();
        std::vector<std::pair<double, size_t>> want;
        const size_t lim = limit > 0 ? static_cast<size_t>(limit) : scans.size();
        for (size_t i = 0; i < lim; ++i) {
          const ndt_2d::Pose2d p = bary ? scans[i]->getBarycenterPose() : scans[i]->getPose();
          double result = 0.0;
          const double a[2] = {q.x, q.y}, b[2] = {p.x, p.y};
          for (int d = 0; d < 2; ++d) {
            const double diff = a[d] - b[d];
            result += diff * diff;
          }
          if (result < radius_sq) {want.emplace_back(result, i);}
        }
        std::sort(want.begin(), want.end());
        const std::vector<size_t> got = ours->findNearest(scans, query, radius_sq, limit, bary);
        bool same = got.size() == want.size();
        for (size_t k = 0; same && k < got.size(); ++k) {same = got[k] == want[k].second;}
        expect(same && !want.empty() && want.size() < lim, "findNearest: same scans, same order");
      }
    }
  }

  // ---- scorePoints at assorted poses ------------------------------------------------
  const std::vector<ndt_2d::Point> qpts = query->getPoints();
  for (int k = 0; k < 8; ++k) {
    ndt_2d::Pose2d p(truth.x + 0.03 * k - 0.1, truth.y - 0.02 * k, truth.theta + 0.01 * k);
    expect(close_rel(dev->scorePoints(qpts, p), ref->scorePoints(qpts, p), 1e-5, 1e-30),
      "scorePoints within 1e-5");
  }

  // ---- particle filter: measure + statistics ------------------------------------------
  {
    const size_t P = 600;
    std::vector<double> u(3 * P), particles(3 * P), weights(P, 1.0 / P);
    ndt2d_synth_normal(7, 3 * P, u.data());
    for (size_t i = 0; i < P; ++i) {
      particles[3 * i] = truth.x + 0.3 * u[3 * i];
      particles[3 * i + 1] = truth.y + 0.3 * u[3 * i + 1];
      particles[3 * i + 2] = truth.theta + 0.1 * u[3 * i + 2];
    }
    ndt_2d::MotionModelPtr rmm(new ndt_2d::MotionModel(0.2, 0.2, 0.2, 0.2, 0.2));
    ndt_2d::ParticleFilter rf(100, P, rmm);
    rf.particles_.clear();
    for (size_t i = 0; i < P; ++i) {
      rf.particles_.emplace_back(particles[3 * i], particles[3 * i + 1], particles[3 * i + 2]);
    }
    rf.weights_.assign(P, 1.0 / P);
    rf.cov_ = Eigen::Matrix3d::Zero();
    rf.measure(ref, query);

    ndt_2d_b200::MotionModelPtr dmm(new ndt_2d_b200::MotionModel(0.2, 0.2, 0.2, 0.2, 0.2));
    ndt_2d_b200::ParticleFilter df(100, P, dmm);
    expect(df.size() == 100, "filter starts with min_particles particles");
    df.setParticles(particles, weights);
    df.measure(dev, query);
    std::vector<double> dp, dw;
    df.getParticles(dp, dw);
    bool w_ok = dw.size() == P;
    for (size_t i = 0; w_ok && i < P; ++i) {w_ok = close_rel(dw[i], rf.weights_[i], 1e-5, 1e-30);}
    expect(w_ok, "measure: normalised weights within 1e-5");
    const Eigen::Vector3d mr = rf.getMean(), md = df.getMean();
    const Eigen::Matrix3d cr = rf.getCovariance(), cd = df.getCovariance();
    std::printf("filter mean ref (%.9f %.9f %.9f) dev (%.9f %.9f %.9f)\n", mr(0), mr(1), mr(2),
      md(0), md(1), md(2));
    for (int k = 0; k < 3; ++k) {expect(close_rel(md(k), mr(k), 1e-5, 1e-9), "filter mean");}
    double cmax = 0.0;
    for (int r = 0; r < 3; ++r) {for (int c = 0; c < 3; ++c) {cmax = std::fmax(cmax, std::fabs(cr(r, c)));}}
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) {expect(close_rel(cd(r, c), cr(r, c), 1e-5, 1e-5 * cmax), "filter covariance");}
    }
    geometry_msgs::msg::PoseArray msg;
    df.getMsg(msg);
    expect(msg.poses.size() == P && msg.poses[3].position.x == particles[9], "getMsg");
    // a foreign matcher cannot be scored on the device: loud failure, no host path
    bool threw = false;
    try {
      df.measure(ref, query);
    } catch (const std::invalid_argument &) {
      threw = true;
    }
    expect(threw, "measure with a non-B200 matcher throws");
    df.resample(0.01, 2.3);
    expect(df.size() >= 100 && df.size() <= P, "resample keeps min <= n <= max");
    (void)kPi;
  }

  // ---- occupancy grid: two calls on one instance (the bounds persist), bit-exact grid
  {
    ndt_2d::OccupancyGrid rg(0.05, 0.25);
    ndt_2d_b200::OccupancyGrid dg(0.05, 0.25);
    for (size_t n : {size_t(4), scans.size()}) {
      std::vector<ndt_2d::ScanPtr> some(scans.begin(), scans.begin() + n);
      nav_msgs::msg::OccupancyGrid a, b;
      rg.getMsg(some, a);
      dg.getMsg(some, b);
      expect(a.info.width == b.info.width && a.info.height == b.info.height &&
        a.info.resolution == b.info.resolution &&
        a.info.origin.position.x == b.info.origin.position.x &&
        a.info.origin.position.y == b.info.origin.position.y, "occupancy grid meta data identical");
      expect(a.data == b.data, "occupancy grid data bit-identical");
      size_t occ = 0;
      for (auto v : b.data) {occ += v == 100 ? 1 : 0;}
      std::printf("occupancy grid: %zu scans -> %u x %u, %zu occupied cells\n", n, b.info.width,
        b.info.height, occ);
    }
  }

  // ---- several GPUs behind ONE plugin instance (one process, one host thread): a global
  // search through ndt_2d::ScanMatcher::matchScan with "<name>.n_gpus" = all visible devices
  // against the same class on one device -- same pose, same score, same covariance
  if (ndt2d_device_count() >= 2) {
    const int n_gpus = std::min(ndt2d_device_count(), 8);
    rclcpp::Node big;
    big.overrides["global_scan_matcher.ndt_resolution"] = 0.25;
    big.overrides["global_scan_matcher.search_linear_size"] = 1.0;
    big.overrides["global_scan_matcher.search_linear_resolution"] = 0.01;
    big.overrides["global_scan_matcher.search_angular_size"] = 0.8;
    big.overrides["global_scan_matcher.search_angular_resolution"] = 0.004;
    big.overrides["global_scan_matcher.laser_max_beams"] = 360;
    rclcpp::Node big_n = big;
    big_n.overrides["global_scan_matcher.n_gpus"] = n_gpus;
    ndt_2d::ScanMatcherPtr one(new ndt_2d_b200::ScanMatcherNDT()), many(new ndt_2d_b200::ScanMatcherNDT());
    one->initialize("global_scan_matcher", &big, 10.0);
    many->initialize("global_scan_matcher", &big_n, 10.0);
    for (auto & m : {one, many}) {m->addScans(scans.begin(), scans.end());}
    // (this test's search is smaller than the default threshold for spreading a search)
    ndt2d_matcher_set_group_threshold(
      dynamic_cast<ndt_2d_b200::ScanMatcherNDT *>(many.get())->handle(), 1.0e8);
    ndt_2d::ScanPtr q3(new ndt_2d::Scan(300));
    q3->setPose(ndt_2d::Pose2d(truth.x - 0.4, truth.y + 0.3, truth.theta - 0.3));
    q3->setPoints(query->getPoints());
    compare_match(one, many, q3, "global search, 1 GPU vs all GPUs of the process");
    compare_match(one, many, q3, "global search again (sequence numbers advance)");
    uint64_t info[4] = {0, 0, 0, 0};
    auto * mm = dynamic_cast<ndt_2d_b200::ScanMatcherNDT *>(many.get());
    ndt2d_matcher_group_info(mm->handle(), info);
    std::printf("multi-GPU plugin: %llu devices, p2p exchange %llu, %llu searches spread over them\n",
      static_cast<unsigned long long>(info[0]), static_cast<unsigned long long>(info[1]),
      static_cast<unsigned long long>(info[2]));
    expect(info[0] == static_cast<uint64_t>(n_gpus) && info[2] == 2, "matchScan ran on every GPU of the handle");
    // a local match through the same instance stays on one device and still agrees
    many->reset();
    one->reset();
  } else {
    std::printf("multi-GPU plugin case skipped: %d device(s) visible\n", ndt2d_device_count());
  }

  std::printf("{\"plugin_parity\": \"%s\", \"failures\": %d}\n", g_failures ? "FAILED" : "ok",
    g_failures);
  return g_failures ? 1 : 0;
}
