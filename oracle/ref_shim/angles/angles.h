// Stand-in for the ROS 2 `angles` package (un-vendored dependency of the
// reference, package.xml:16): the two published helpers the hot path uses.
#ifndef NDT2D_ORACLE_ANGLES_SHIM_H_
#define NDT2D_ORACLE_ANGLES_SHIM_H_
#include <cmath>
namespace angles
{
static inline double normalize_angle_positive(double angle)
{
  const double result = std::fmod(angle, 2.0 * M_PI);
  if (result < 0) {return result + 2.0 * M_PI;}
  return result;
}
static inline double normalize_angle(double angle)
{
  const double result = std::fmod(angle + M_PI, 2.0 * M_PI);
  if (result <= 0.0) {return result + M_PI;}
  return result - M_PI;
}
static inline double shortest_angular_distance(double from, double to)
{
  return normalize_angle(to - from);
}
}  // namespace angles
#endif
