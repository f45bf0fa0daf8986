// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <boost/format.hpp> (include/utils.h:10,17-18; only "%0Nd"-style int fields are used).
#pragma once
#include <cstdio>
#include <ostream>
#include <string>
namespace boost {
class format {
 public:
  explicit format(const std::string& f) : fmt_(f) {}
  template <class T> format& operator%(const T& v) {
    // substitute the first remaining %...d / %...s / %...f directive
    size_t p = out_.empty() && !started_ ? 0 : 0;
    (void)p;
    if (!started_) { out_ = fmt_; started_ = true; }
    size_t a = out_.find('%', cursor_);
    if (a == std::string::npos) return *this;
    size_t b = a + 1;
    while (b < out_.size() && !std::isalpha(static_cast<unsigned char>(out_[b]))) ++b;
    std::string spec = out_.substr(a, b - a + 1);
    char buf[128];
    render(buf, sizeof(buf), spec, v);
    out_.replace(a, b - a + 1, buf);
    cursor_ = a + std::string(buf).size();
    return *this;
  }
  std::string str() const { return started_ ? out_ : fmt_; }
 private:
  static void render(char* buf, size_t n, const std::string& spec, int v) { std::snprintf(buf, n, spec.c_str(), v); }
  static void render(char* buf, size_t n, const std::string& spec, long v) { std::string s = spec; s.insert(s.size() - 1, "l"); std::snprintf(buf, n, s.c_str(), v); }
  static void render(char* buf, size_t n, const std::string& spec, unsigned long v) { std::string s = spec; s[s.size() - 1] = 'u'; s.insert(s.size() - 1, "l"); std::snprintf(buf, n, s.c_str(), v); }
  static void render(char* buf, size_t n, const std::string& spec, double v) { std::snprintf(buf, n, spec.c_str(), v); }
  static void render(char* buf, size_t n, const std::string&, const std::string& v) { std::snprintf(buf, n, "%s", v.c_str()); }
  std::string fmt_, out_;
  bool started_ = false;
  size_t cursor_ = 0;
};
inline std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }
inline std::string str(const format& f) { return f.str(); }
namespace io { template <class T> const T& group(const T& t) { return t; } }
}  // namespace boost
