// TEST INFRASTRUCTURE ONLY -- stand-in for <ros/ros.h>: a NodeHandle whose param<T>()
// always yields the caller's default, which is all param_config.h needs to compile.
#ifndef RSM_STANDIN_ROS_ROS_H
#define RSM_STANDIN_ROS_ROS_H
#include <string>
namespace ros {
class NodeHandle {
 public:
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  template <typename T, typename D>
  bool param(const std::string&, T& value, const D& dflt) const { value = T(dflt); return false; }
};
}  // namespace ros
#endif
