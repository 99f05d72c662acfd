// Minimal rclcpp::Node stand-in (test infrastructure): only declare_parameter,
// the single rclcpp facility the reference's matcher plugin uses
// (src/scan_matcher_ndt.cpp:35-47).  Overrides play the role of a params file.
#ifndef NDT2D_ORACLE_RCLCPP_SHIM_HPP_
#define NDT2D_ORACLE_RCLCPP_SHIM_HPP_
#include <map>
#include <string>
namespace rclcpp
{
class Node
{
public:
  template<typename T>
  T declare_parameter(const std::string & name, const T & default_value)
  {
    auto it = overrides.find(name);
    if (it == overrides.end()) {return default_value;}
    return static_cast<T>(it->second);
  }
  std::map<std::string, double> overrides;
};
}  // namespace rclcpp
#endif
