// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for <ceres/rotation.h> (included by the reference, nothing in it is used).
#pragma once
