// ORACLE SHIM (test infrastructure).  Stand-in for Slam_Utility's `slam_basic_math.h`.  The KLT sources include it without
// using it; dense_optical_flow.cpp uses slam_utility::Utility::Interpolate(Mat, row, col) (:75-76, :319-324), a bilinear
// lookup in a float matrix.  Upstream is absent, so its boundary behaviour is FROZEN here: the position is clamped to
// [0, rows-1] x [0, cols-1], the "+1" neighbours are clamped to the last row / column, and the four weighted terms are added
// left to right: (1-dr)(1-dc) m00 + (1-dr) dc m01 + dr (1-dc) m10 + dr dc m11.
#ifndef _ORACLE_SHIM_SLAM_BASIC_MATH_H_
#define _ORACLE_SHIM_SLAM_BASIC_MATH_H_
#include <algorithm>
#include <cmath>

#include "basic_type.h"

namespace slam_utility {

class Utility {
public:
    static float Interpolate(const Mat &m, float row, float col) {
        const float max_r = static_cast<float>(m.rows() - 1), max_c = static_cast<float>(m.cols() - 1);
        const float r = row < 0.0f ? 0.0f : (row > max_r ? max_r : row);
        const float c = col < 0.0f ? 0.0f : (col > max_c ? max_c : col);
        const float fr = std::floor(r), fc = std::floor(c);
        const int r0 = static_cast<int>(fr), c0 = static_cast<int>(fc);
        const int r1 = r0 + 1 < m.rows() ? r0 + 1 : m.rows() - 1, c1 = c0 + 1 < m.cols() ? c0 + 1 : m.cols() - 1;
        const float dr = r - fr, dc = c - fc;
        const float ir = 1.0f - dr, ic = 1.0f - dc;
        return ir * ic * m(r0, c0) + ir * dc * m(r0, c1) + dr * ic * m(r1, c0) + dr * dc * m(r1, c1);
    }
};

}  // namespace slam_utility
#endif
