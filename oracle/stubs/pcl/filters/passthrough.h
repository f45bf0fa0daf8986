// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <pcl/filters/passthrough.h> (included by the reference, nothing in it is used).
#pragma once
#include <pcl/filters/filter.h>
