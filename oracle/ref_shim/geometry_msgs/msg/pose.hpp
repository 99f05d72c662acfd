#ifndef NDT2D_ORACLE_GEOMETRY_MSGS_SHIM_HPP_
#define NDT2D_ORACLE_GEOMETRY_MSGS_SHIM_HPP_
#include <vector>
namespace geometry_msgs
{
namespace msg
{
struct Point {double x = 0, y = 0, z = 0;};
struct Vector3 {double x = 0, y = 0, z = 0;};
struct Quaternion {double x = 0, y = 0, z = 0, w = 1;};
struct Pose {Point position; Quaternion orientation;};
struct PoseStamped {Pose pose;};
struct Transform {Vector3 translation; Quaternion rotation;};
struct TransformStamped {Transform transform;};
struct PoseArray {std::vector<Pose> poses;};
}  // namespace msg
}  // namespace geometry_msgs
#endif
