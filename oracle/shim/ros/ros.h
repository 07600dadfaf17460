// ORACLE / TEST INFRASTRUCTURE: stand-in for <ros/ros.h>.
// src/meas/*.h include it but use nothing of ROS on the hot path; they do rely on the standard headers it pulls in
// (std::set, std::map, ...).  src/solver/ViconGraphSolver.cpp uses a NodeHandle as its parameter source, publishers
// for RViz and the ROS_* log macros: parameters come from a process-wide table the wrapper fills (ref_wrap.cpp),
// publishers and logging do nothing.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_INFO_THROTTLE(...) ((void)0)

namespace ros {

struct ParamTable {
  std::map<std::string, double> scalars;               // double / int / bool parameters
  std::map<std::string, std::vector<double>> vectors;  // std::vector<double> parameters
  static ParamTable &instance() {
    static ParamTable t;
    return t;
  }
};

inline bool ok() { return true; }

struct Time {
  double sec = 0;
  Time() {}
  explicit Time(double s) : sec(s) {}
  static Time now() { return Time(0.0); }
};

class Publisher {
  std::string topic_;

public:
  Publisher() {}
  explicit Publisher(const std::string &t) : topic_(t) {}
  template <class M> void publish(const M &) const {}
  std::string getTopic() const { return topic_; }
};

class NodeHandle {
public:
  NodeHandle() {}
  explicit NodeHandle(const std::string &) {}
  template <class T> bool param(const std::string &name, T &out, const T &def) const {
    auto &t = ParamTable::instance().scalars;
    auto it = t.find(name);
    if (it == t.end()) {
      out = def;
      return false;
    }
    out = static_cast<T>(it->second);
    return true;
  }
  template <class M> Publisher advertise(const std::string &topic, int) { return Publisher(topic); }
};
template <>
inline bool NodeHandle::param<std::vector<double>>(const std::string &name, std::vector<double> &out, const std::vector<double> &def) const {
  auto &t = ParamTable::instance().vectors;
  auto it = t.find(name);
  if (it == t.end()) {
    out = def;
    return false;
  }
  out = it->second;
  return true;
}

} // namespace ros
