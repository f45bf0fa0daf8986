// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl_ros/impl/transforms.hpp> (include/laserMappingClass.h:14): brings pcl/common/transforms.h in.
#pragma once
#include <pcl/common/transforms.h>
