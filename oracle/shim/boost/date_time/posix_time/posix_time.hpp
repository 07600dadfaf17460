// ORACLE / TEST INFRASTRUCTURE: stand-in for <boost/date_time/posix_time/posix_time.hpp> (ptime differences in
// microseconds, used by src/solver/ViconGraphSolver.cpp only for its timing printout).
#pragma once
#include <chrono>
namespace boost {
namespace posix_time {
struct time_duration {
  long long us = 0;
  long long total_microseconds() const { return us; }
};
struct ptime {
  std::chrono::steady_clock::time_point t;
};
inline time_duration operator-(const ptime &a, const ptime &b) {
  time_duration d;
  d.us = std::chrono::duration_cast<std::chrono::microseconds>(a.t - b.t).count();
  return d;
}
struct microsec_clock {
  static ptime local_time() {
    ptime p;
    p.t = std::chrono::steady_clock::now();
    return p;
  }
};
} // namespace posix_time
} // namespace boost
