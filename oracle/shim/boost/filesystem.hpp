// ORACLE / TEST INFRASTRUCTURE: stand-in for <boost/filesystem.hpp> (exists / remove / create_directories / path as
// used by src/solver/ViconGraphSolver.cpp:198-211), forwarding to std::filesystem.
#pragma once
#include <filesystem>
#include <string>
namespace boost {
namespace filesystem {
class path {
  std::filesystem::path p_;

public:
  path() {}
  path(const std::string &s) : p_(s) {}
  path(const std::filesystem::path &p) : p_(p) {}
  path parent_path() const { return path(p_.parent_path()); }
  const std::filesystem::path &std_path() const { return p_; }
  std::string string() const { return p_.string(); }
};
inline bool exists(const path &p) { return std::filesystem::exists(p.std_path()); }
inline bool remove(const path &p) { return std::filesystem::remove(p.std_path()); }
inline bool create_directories(const path &p) {
  if (p.std_path().empty())
    return false;
  return std::filesystem::create_directories(p.std_path());
}
} // namespace filesystem
} // namespace boost
