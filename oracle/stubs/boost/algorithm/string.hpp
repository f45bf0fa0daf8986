// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <boost/algorithm/string.hpp>: to_lower_copy (src/odomEstimationClass.cpp:23).
#pragma once
#include <cctype>
#include <string>
namespace boost {
namespace algorithm {
inline std::string to_lower_copy(const std::string& in) {
  std::string out(in);
  for (size_t i = 0; i < out.size(); ++i) out[i] = static_cast<char>(std::tolower(static_cast<unsigned char>(out[i])));
  return out;
}
}  // namespace algorithm
using algorithm::to_lower_copy;
}  // namespace boost
