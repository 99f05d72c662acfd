// tf2::getYaw stand-in; only referenced by conversions.hpp's fromMsg helpers,
// which are outside the hot path.
#ifndef NDT2D_ORACLE_TF2_SHIM_H_
#define NDT2D_ORACLE_TF2_SHIM_H_
#include <cmath>
namespace tf2
{
template<typename Q>
double getYaw(const Q & q)
{
  return std::atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
}
}  // namespace tf2
#endif
