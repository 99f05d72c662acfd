// Minimal nav_msgs/msg/OccupancyGrid stand-in (test infrastructure): the fields
// src/occupancy_grid.cpp writes (nav_msgs/MapMetaData: float32 resolution, uint32 width /
// height, geometry_msgs/Pose origin; int8[] data).
#ifndef NDT2D_ORACLE_NAV_MSGS_SHIM_HPP_
#define NDT2D_ORACLE_NAV_MSGS_SHIM_HPP_
#include <cstdint>
#include <vector>
#include "../../geometry_msgs/msg/pose.hpp"
namespace nav_msgs
{
namespace msg
{
struct MapMetaData
{
  float resolution = 0.0f;
  uint32_t width = 0, height = 0;
  geometry_msgs::msg::Pose origin;
};
struct OccupancyGrid
{
  MapMetaData info;
  std::vector<int8_t> data;
};
}  // namespace msg
}  // namespace nav_msgs
#endif
