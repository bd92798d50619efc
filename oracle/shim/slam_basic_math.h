// ORACLE SHIM (test infrastructure).  The hot-path sources include this header but use nothing from it.
#ifndef _ORACLE_SHIM_SLAM_BASIC_MATH_H_
#define _ORACLE_SHIM_SLAM_BASIC_MATH_H_
#include "basic_type.h"
#endif
