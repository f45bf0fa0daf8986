// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <sensor_msgs/PointCloud2.h> / PointField.h.
#pragma once
#include <boost/shared_ptr.hpp>
#include <std_msgs/Header.h>
#include <cstdint>
#include <string>
#include <vector>
namespace sensor_msgs {
struct PointField {
  enum { INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8 };
  std::string name;
  std::uint32_t offset = 0;
  std::uint8_t datatype = 0;
  std::uint32_t count = 0;
};
struct PointCloud2 {
  std_msgs::Header header;
  std::uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  std::uint32_t point_step = 0, row_step = 0;
  std::vector<std::uint8_t> data;
  bool is_dense = false;
  typedef boost::shared_ptr<PointCloud2> Ptr;
  typedef boost::shared_ptr<PointCloud2 const> ConstPtr;
};
typedef boost::shared_ptr<PointCloud2> PointCloud2Ptr;
typedef boost::shared_ptr<PointCloud2 const> PointCloud2ConstPtr;
}  // namespace sensor_msgs
