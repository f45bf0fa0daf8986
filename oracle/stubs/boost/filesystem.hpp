// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <boost/filesystem.hpp>: create_directories (src/utils.cpp, src/odomEstimationNode.cpp:67).
#pragma once
#include <string>
#include <sys/stat.h>
#include <sys/types.h>
namespace boost {
namespace filesystem {
inline bool create_directories(const std::string& dir) {
  bool made = false;
  for (size_t i = 1; i <= dir.size(); ++i)
    if (i == dir.size() || dir[i] == '/') { if (::mkdir(dir.substr(0, i).c_str(), 0777) == 0) made = true; }
  return made;
}
}  // namespace filesystem
}  // namespace boost
