// ndt_2d_b200::OccupancyGrid -- see include/ndt_2d_b200/occupancy_grid.hpp.
#include <ndt_2d_b200/occupancy_grid.hpp>

#include <cstdint>
#include <stdexcept>
#include <string>

namespace ndt_2d_b200
{

namespace
{
void check(const char * where, int status)
{
  if (status != NDT2D_OK) {
    throw std::runtime_error(
            std::string("ndt_2d_b200::OccupancyGrid::") + where + ": libndt2d_b200 status " +
            std::to_string(status) + " (" + ndt2d_last_error() + "); there is no CPU fallback");
  }
}
}  // namespace

OccupancyGrid::OccupancyGrid(const double resolution, const double occ_thresh, int device)
{
  check("OccupancyGrid", ndt2d_occupancy_create(resolution, occ_thresh, device, &handle_));
}

OccupancyGrid::~OccupancyGrid()
{
  if (handle_) {ndt2d_occupancy_destroy(handle_);}
}

void OccupancyGrid::getMsg(std::vector<ndt_2d::ScanPtr> & scans, nav_msgs::msg::OccupancyGrid & grid)
{
  std::vector<double> poses, points;
  std::vector<uint64_t> offsets(1, 0);
  for (auto & scan : scans) {
    const ndt_2d::Pose2d pose = scan->getPose();
    poses.push_back(pose.x);
    poses.push_back(pose.y);
    poses.push_back(pose.theta);
    for (const auto & p : scan->getPoints()) {
      points.push_back(p.x);
      points.push_back(p.y);
    }
    offsets.push_back(points.size() / 2);
  }
  double info[5];
  check("getMsg", ndt2d_occupancy_render(handle_, scans.size(), poses.data(), offsets.data(),
    points.data(), info));
  // meta data (occupancy_grid.cpp:57-63)
  grid.info.resolution = info[4];
  grid.info.width = static_cast<uint32_t>(info[0]);
  grid.info.height = static_cast<uint32_t>(info[1]);
  grid.info.origin.position.x = info[2];
  grid.info.origin.position.y = info[3];
  grid.info.origin.orientation.w = 1.0;
  grid.data.assign(static_cast<size_t>(grid.info.width) * grid.info.height, -1);
  check("getMsg", ndt2d_occupancy_fetch(handle_, grid.data.data(), grid.data.size()));
}

}  // namespace ndt_2d_b200
