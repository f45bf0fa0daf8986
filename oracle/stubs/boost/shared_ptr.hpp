// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <boost/shared_ptr.hpp> (Boost is absent from this image).
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::make_shared;
using std::const_pointer_cast;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
}  // namespace boost
