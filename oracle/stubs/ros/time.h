// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <ros/time.h> (rostime, ROS Melodic): Time / Duration arithmetic as the reference uses it.
#pragma once
#include <math.h>  // rostime's duration.h includes the C header
#include <chrono>
#include <cmath>
#include <cstdint>
#include <ostream>
namespace ros {
struct Duration {
  std::int32_t sec = 0, nsec = 0;
  Duration() {}
  explicit Duration(double t) { fromSec(t); }
  Duration& fromSec(double t) { sec = (std::int32_t)floor(t); nsec = (std::int32_t)std::llround((t - sec) * 1e9); return *this; }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
  bool operator>(const Duration& o) const { return sec > o.sec || (sec == o.sec && nsec > o.nsec); }
  bool operator<(const Duration& o) const { return o > *this; }
};
struct Time {
  std::uint32_t sec = 0, nsec = 0;
  Time() {}
  Time(std::uint32_t s, std::uint32_t ns) : sec(s), nsec(ns) {}
  explicit Time(double t) { fromSec(t); }
  // TimeBase::fromSec (rostime 0.6.x): sec = floor(t); nsec = round((t - sec) * 1e9); carry
  Time& fromSec(double t) {
    std::int64_t sec64 = static_cast<std::int64_t>(floor(t));
    sec = static_cast<std::uint32_t>(sec64);
    nsec = static_cast<std::uint32_t>(std::llround((t - sec) * 1e9));
    sec += (nsec / 1000000000ul);
    nsec %= 1000000000ul;
    return *this;
  }
  Time& fromNSec(std::uint64_t t) { sec = (std::uint32_t)(t / 1000000000ull); nsec = (std::uint32_t)(t % 1000000000ull); return *this; }
  std::uint64_t toNSec() const { return (std::uint64_t)sec * 1000000000ull + (std::uint64_t)nsec; }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
  static Time now() {
    const auto ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
    Time t; t.fromNSec((std::uint64_t)ns); return t;
  }
  Duration operator-(const Time& o) const { Duration d; d.fromSec(toSec() - o.toSec()); return d; }
  bool operator<(const Time& o) const { return sec < o.sec || (sec == o.sec && nsec < o.nsec); }
  bool operator==(const Time& o) const { return sec == o.sec && nsec == o.nsec; }
};
inline std::ostream& operator<<(std::ostream& os, const Time& t) { return os << t.sec << "." << t.nsec; }
inline std::ostream& operator<<(std::ostream& os, const Duration& t) { return os << t.toSec(); }
}  // namespace ros
