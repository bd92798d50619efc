// ORACLE SHIM (test infrastructure).  The hot-path sources include the log reporter but never log.
#ifndef _ORACLE_SHIM_SLAM_LOG_REPORTER_H_
#define _ORACLE_SHIM_SLAM_LOG_REPORTER_H_
#include <iostream>
#define ReportInfo(...) \
    do {                \
    } while (0)
#define ReportError(...) \
    do {                 \
    } while (0)
#endif
