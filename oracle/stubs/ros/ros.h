// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <ros/ros.h>: just enough of roscpp for the reference's sources to compile
// (class sources: ros::Time::now() in dead timing code; node sources: syntax check only, nothing here talks to a ROS master).
#pragma once
#include <ros/time.h>
#include <boost/shared_ptr.hpp>
#include <cstdio>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <string>
#define ROS_INFO(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_WARN(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_ERROR(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_WARN_ONCE(...) do { static bool once_ = false; if (!once_) { once_ = true; std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)
#define ROS_INFO_STREAM(x) do { std::cout << x << std::endl; } while (0)
#define ROS_INFO_STREAM_THROTTLE(t, x) do { std::cout << x << std::endl; } while (0)
namespace ros {
class Publisher {
 public:
  template <class M> void publish(const M&) const {}
};
class Subscriber {};
class NodeHandle {
 public:
  NodeHandle(const std::string& ns = std::string()) { (void)ns; }
  template <class T> bool getParam(const std::string&, T&) const { return false; }
  template <class M> Subscriber subscribe(const std::string&, std::uint32_t, void (*)(const boost::shared_ptr<M const>&)) { return Subscriber(); }
  template <class M> Publisher advertise(const std::string&, std::uint32_t, bool latch = false) { (void)latch; return Publisher(); }
};
class Rate {
 public:
  explicit Rate(double) {}
  bool sleep() { return true; }
};
inline void init(int&, char**, const std::string&) {}
inline void spin() {}
inline void spinOnce() {}
inline bool ok() { return false; }
}  // namespace ros
