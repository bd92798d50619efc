// ORACLE SHIM (test infrastructure).  Control-flow macros of Slam_Utility.
#ifndef _ORACLE_SHIM_SLAM_OPERATIONS_H_
#define _ORACLE_SHIM_SLAM_OPERATIONS_H_
#define RETURN_FALSE_IF(c) \
    if (c) {               \
        return false;      \
    }
#define RETURN_FALSE_IF_FALSE(c) \
    if (!(c)) {                  \
        return false;            \
    }
#define CONTINUE_IF(c) \
    if (c) {           \
        continue;      \
    }
#define BREAK_IF(c) \
    if (c) {        \
        break;      \
    }
#endif
