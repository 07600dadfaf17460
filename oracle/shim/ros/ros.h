// ORACLE / TEST INFRASTRUCTURE: stand-in for <ros/ros.h>.  src/meas/*.h include it but use nothing of
// ROS on the hot path; they do rely on the standard headers it pulls in (std::set, std::map, ...).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>
