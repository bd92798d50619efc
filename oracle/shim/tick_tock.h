// ORACLE SHIM (test infrastructure).  Slam_Utility's timer header; nn_feature_matcher.cpp includes it but never uses it.
#ifndef _ORACLE_SHIM_TICK_TOCK_H_
#define _ORACLE_SHIM_TICK_TOCK_H_
#endif
