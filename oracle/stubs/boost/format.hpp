// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <boost/format.hpp> (include/utils.h:10,17-18; src/utils.cpp:64,94).
// Like Boost.Format, a printf-style directive only sets STREAM formatting state (fill, width, precision, fixed / dec) and the argument
// is then written with operator<< — so "%lf" applied to an integer prints the integer, "%06d" zero-pads to six columns.
#pragma once
#include <cctype>
#include <cstdio>
#include <iomanip>
#include <ostream>
#include <sstream>
#include <string>
namespace boost {
class format {
 public:
  explicit format(const std::string& f) : out_(f), cursor_(0) {}
  template <class T> format& operator%(const T& v) {
    size_t a = out_.find('%', cursor_);
    if (a == std::string::npos) return *this;
    size_t b = a + 1;
    std::ostringstream os;
    if (b < out_.size() && out_[b] == '0') { os.fill('0'); ++b; }
    int width = 0;
    while (b < out_.size() && std::isdigit(static_cast<unsigned char>(out_[b]))) width = width * 10 + (out_[b++] - '0');
    if (b < out_.size() && out_[b] == '.') {
      ++b; int prec = 0;
      while (b < out_.size() && std::isdigit(static_cast<unsigned char>(out_[b]))) prec = prec * 10 + (out_[b++] - '0');
      os.precision(prec);
    }
    while (b < out_.size() && (out_[b] == 'l' || out_[b] == 'h')) ++b;   // length modifiers are ignored
    const char conv = b < out_.size() ? out_[b] : 's';
    if (conv == 'f') os.setf(std::ios_base::fixed, std::ios_base::floatfield);
    if (conv == 'e') os.setf(std::ios_base::scientific, std::ios_base::floatfield);
    if (width) os.width(width);
    os << v;
    const std::string piece = os.str();
    out_.replace(a, b - a + 1, piece);
    cursor_ = a + piece.size();
    return *this;
  }
  std::string str() const { return out_; }
 private:
  std::string out_;
  size_t cursor_;
};
inline std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }
inline std::string str(const format& f) { return f.str(); }
namespace io { template <class T> const T& group(const T& t) { return t; } }
}  // namespace boost
