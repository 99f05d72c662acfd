// ndt_2d_b200::OccupancyGrid -- device replacement of ndt_2d::OccupancyGrid
// (include/ndt_2d/occupancy_grid.hpp:40-68, src/occupancy_grid.cpp) with the same
// constructor and getMsg(); the int8 grid, its size and origin are bit-identical.
#ifndef NDT_2D_B200__OCCUPANCY_GRID_HPP_
#define NDT_2D_B200__OCCUPANCY_GRID_HPP_

#include <memory>
#include <vector>

#include <nav_msgs/msg/occupancy_grid.hpp>
#include <ndt_2d/scan.hpp>

#include "ndt2d_b200.h"

namespace ndt_2d_b200
{

class OccupancyGrid
{
public:
  explicit OccupancyGrid(const double resolution, const double occ_thresh, int device = -1);
  ~OccupancyGrid();
  OccupancyGrid(const OccupancyGrid &) = delete;
  OccupancyGrid & operator=(const OccupancyGrid &) = delete;

  // occupancy_grid.cpp:47-152
  void getMsg(std::vector<ndt_2d::ScanPtr> & scans, nav_msgs::msg::OccupancyGrid & grid);

private:
  ndt2d_occupancy * handle_ = nullptr;
};

typedef std::shared_ptr<OccupancyGrid> OccupancyGridPtr;

}  // namespace ndt_2d_b200

#endif  // NDT_2D_B200__OCCUPANCY_GRID_HPP_
