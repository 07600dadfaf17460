// ORACLE / TEST INFRASTRUCTURE: stand-in for <nav_msgs/Path.h>.
#pragma once
#include <geometry_msgs/PoseStamped.h>
namespace nav_msgs {
struct Path {
  std_msgs::Header header;
  std::vector<geometry_msgs::PoseStamped> poses;
};
} // namespace nav_msgs
