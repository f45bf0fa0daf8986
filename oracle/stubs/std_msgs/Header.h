// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <std_msgs/Header.h>.
#pragma once
#include <ros/time.h>
#include <string>
namespace std_msgs {
struct Header {
  std::uint32_t seq = 0;
  ros::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs
