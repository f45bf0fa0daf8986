// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <tf/transform_broadcaster.h> (node sources only: syntax check).
#pragma once
#include <tf/transform_datatypes.h>
namespace tf {
class TransformBroadcaster {
 public:
  void sendTransform(const StampedTransform&) {}
};
}  // namespace tf
