// ORACLE SHIM (test infrastructure, not product code).
// Minimal stand-in for Slam_Utility's `basic_type.h` (Eigen float typedefs), which is absent from
// /root/reference.  Only the operations the hot-path sources actually use are provided (SURVEY.md
// Appendix A item 1).  All arithmetic is plain sequential IEEE fp32, evaluated left to right; products and
// sums of length n are `a0*b0 + a1*b1 + ...` starting from the first product.  These semantics ARE the parity
// specification for the external (un-vendored) Eigen dependency; see DESIGN.md "Oracle".
#ifndef _ORACLE_SHIM_BASIC_TYPE_H_
#define _ORACLE_SHIM_BASIC_TYPE_H_

#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

constexpr int32_t kMaxInt32 = 2147483647;

namespace shim {

template <int N> struct Ldlt;
template <int R, int C> struct Mat;

template <int R, int C>
struct CommaInit {
    Mat<R, C> &m;
    int idx;
    template <typename T> CommaInit &operator,(T v) {
        m.d[idx / C][idx % C] = static_cast<float>(v);
        ++idx;
        return *this;
    }
};

// Mutable view on a sub-block (used for col(i) +=, block<>() = / setIdentity()).
template <int R, int C, int BR, int BC>
struct BlockRef {
    Mat<R, C> &m;
    int r0, c0;
    void setIdentity() {
        for (int i = 0; i < BR; ++i)
            for (int j = 0; j < BC; ++j) m.d[r0 + i][c0 + j] = (i == j) ? 1.0f : 0.0f;
    }
    BlockRef &operator=(const Mat<BR, BC> &o) {
        for (int i = 0; i < BR; ++i)
            for (int j = 0; j < BC; ++j) m.d[r0 + i][c0 + j] = o.d[i][j];
        return *this;
    }
    BlockRef &operator+=(const Mat<BR, BC> &o) {
        for (int i = 0; i < BR; ++i)
            for (int j = 0; j < BC; ++j) m.d[r0 + i][c0 + j] = m.d[r0 + i][c0 + j] + o.d[i][j];
        return *this;
    }
    float norm() const {
        float s = 0.0f;
        bool first = true;
        for (int i = 0; i < BR; ++i)
            for (int j = 0; j < BC; ++j) {
                const float p = m.d[r0 + i][c0 + j] * m.d[r0 + i][c0 + j];
                s = first ? p : s + p;
                first = false;
            }
        return std::sqrt(s);
    }
};

struct IsNanResult {
    bool v;
    bool any() const { return v; }
};

template <int R, int C>
struct ArrayView {
    const Mat<R, C> &m;
};

template <int R, int C>
struct Mat {
    float d[R][C];

    Mat() {}  // uninitialised, like Eigen
    // Two-coefficient constructors (Vec2(int,int), Vec2(float,float), Mat1x2(a,b)).
    template <typename A, typename B>
    Mat(A a, B b) {
        static_assert(R * C == 2, "two-coefficient constructor needs a 2-vector");
        (&d[0][0])[0] = static_cast<float>(a);
        (&d[0][0])[1] = static_cast<float>(b);
    }
    explicit Mat(float a) {
        static_assert(R * C == 1, "one-coefficient constructor needs a 1x1");
        d[0][0] = a;
    }

    static Mat Zero() {
        Mat m;
        m.setZero();
        return m;
    }
    static Mat Identity() {
        Mat m;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) m.d[i][j] = (i == j) ? 1.0f : 0.0f;
        return m;
    }
    void setZero() {
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) d[i][j] = 0.0f;
    }

    static constexpr int rows() { return R; }  // nn_feature_matcher.cpp:98-101 sizes its input tensors with them
    static constexpr int cols() { return C; }
    float &operator()(int i, int j) { return d[i][j]; }
    const float &operator()(int i, int j) const { return d[i][j]; }
    float &operator()(int i) { return (&d[0][0])[i]; }
    const float &operator()(int i) const { return (&d[0][0])[i]; }
    float &x() { return (&d[0][0])[0]; }
    const float &x() const { return (&d[0][0])[0]; }
    float &y() { return (&d[0][0])[1]; }
    const float &y() const { return (&d[0][0])[1]; }
    float &z() { return (&d[0][0])[2]; }
    const float &z() const { return (&d[0][0])[2]; }

    Mat operator+(const Mat &o) const {
        Mat r;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) r.d[i][j] = d[i][j] + o.d[i][j];
        return r;
    }
    Mat operator-(const Mat &o) const {
        Mat r;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) r.d[i][j] = d[i][j] - o.d[i][j];
        return r;
    }
    Mat &operator+=(const Mat &o) {
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) d[i][j] = d[i][j] + o.d[i][j];
        return *this;
    }
    Mat &operator-=(const Mat &o) {
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) d[i][j] = d[i][j] - o.d[i][j];
        return *this;
    }
    Mat operator*(float s) const {
        Mat r;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) r.d[i][j] = d[i][j] * s;
        return r;
    }
    Mat operator/(float s) const {
        Mat r;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) r.d[i][j] = d[i][j] / s;
        return r;
    }
    Mat &operator*=(float s) {
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) d[i][j] = d[i][j] * s;
        return *this;
    }
    Mat &operator/=(float s) {
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) d[i][j] = d[i][j] / s;
        return *this;
    }
    // Matrix product: r(i,j) = a(i,0)*b(0,j) + a(i,1)*b(1,j) + ... (left to right).
    template <int K>
    Mat<R, K> operator*(const Mat<C, K> &o) const {
        Mat<R, K> r;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < K; ++j) {
                float s = d[i][0] * o.d[0][j];
                for (int k = 1; k < C; ++k) s = s + d[i][k] * o.d[k][j];
                r.d[i][j] = s;
            }
        return r;
    }
    // In-place right multiplication (R_cr *= delta_R); evaluated into a temporary first, like Eigen.
    Mat &operator*=(const Mat<C, C> &o) {
        const Mat t = (*this) * o;
        *this = t;
        return *this;
    }
    Mat<C, R> transpose() const {
        Mat<C, R> r;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) r.d[j][i] = d[i][j];
        return r;
    }
    float squaredNorm() const {
        const float *p = &d[0][0];
        float s = p[0] * p[0];
        for (int i = 1; i < R * C; ++i) s = s + p[i] * p[i];
        return s;
    }
    float norm() const { return std::sqrt(squaredNorm()); }
    float dot(const Mat &o) const {
        const float *p = &d[0][0];
        const float *q = &o.d[0][0];
        float s = p[0] * q[0];
        for (int i = 1; i < R * C; ++i) s = s + p[i] * q[i];
        return s;
    }

    template <int N> Mat<N, 1> head() const { return segment<N>(0); }
    template <int N> Mat<N, 1> tail() const { return segment<N>(R * C - N); }
    template <int N> Mat<N, 1> segment(int start) const {
        Mat<N, 1> r;
        for (int i = 0; i < N; ++i) r.d[i][0] = (&d[0][0])[start + i];
        return r;
    }
    BlockRef<R, C, R, 1> col(int j) { return BlockRef<R, C, R, 1>{*this, 0, j}; }
    template <int BR, int BC> BlockRef<R, C, BR, BC> block(int r0, int c0) { return BlockRef<R, C, BR, BC>{*this, r0, c0}; }

    template <typename T> CommaInit<R, C> operator<<(T v) {
        d[0][0] = static_cast<float>(v);
        return CommaInit<R, C>{*this, 1};
    }

    ArrayView<R, C> array() const { return ArrayView<R, C>{*this}; }
    Ldlt<R> ldlt() const;

    // dense_optical_flow.cpp:216-218.  trace(): d00 + d11 + ...; inverse() of a 2x2 restates Eigen's
    // compute_inverse_size2_helper: invdet = 1 / (a*d - b*c); [d, -b; -c, a] * invdet.
    float trace() const {
        static_assert(R == C, "trace needs a square matrix");
        float s = d[0][0];
        for (int i = 1; i < R; ++i) s = s + d[i][i];
        return s;
    }
    Mat inverse() const {
        static_assert(R == 2 && C == 2, "only the 2x2 inverse is used");
        const float invdet = 1.0f / (d[0][0] * d[1][1] - d[1][0] * d[0][1]);
        Mat r;
        r.d[0][0] = d[1][1] * invdet;
        r.d[1][0] = -d[1][0] * invdet;
        r.d[0][1] = -d[0][1] * invdet;
        r.d[1][1] = d[0][0] * invdet;
        return r;
    }
};

// Restatement of Eigen 3.3/3.4 LDLT<MatrixType, Lower>: in-place unblocked
// factorisation with diagonal pivoting on the largest |diagonal| (first maximum wins), and the
// pseudo-inverse-of-D solve with tolerance numeric_limits<float>::min().  Inner sums are accumulated
// sequentially (ascending index, starting from the first product) and then subtracted.
template <int N>
struct Ldlt {
    float a[N][N];
    int tr[N];

    explicit Ldlt(const Mat<N, N> &m) {
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) a[i][j] = m.d[i][j];
        factor();
    }
    static void swapf(float &p, float &q) {
        const float t = p;
        p = q;
        q = t;
    }
    void factor() {
        if (N <= 1) {
            tr[0] = 0;
            return;
        }
        float temp[N];
        for (int k = 0; k < N; ++k) {
            int p = k;
            float best = std::fabs(a[k][k]);
            for (int i = k + 1; i < N; ++i) {
                const float v = std::fabs(a[i][i]);
                if (v > best) {
                    best = v;
                    p = i;
                }
            }
            tr[k] = p;
            if (p != k) {
                for (int j = 0; j < k; ++j) swapf(a[k][j], a[p][j]);
                for (int i = p + 1; i < N; ++i) swapf(a[i][k], a[i][p]);
                swapf(a[k][k], a[p][p]);
                for (int i = k + 1; i < p; ++i) swapf(a[i][k], a[p][i]);
            }
            if (k > 0) {
                for (int j = 0; j < k; ++j) temp[j] = a[j][j] * a[k][j];
                {
                    float s = a[k][0] * temp[0];
                    for (int j = 1; j < k; ++j) s = s + a[k][j] * temp[j];
                    a[k][k] = a[k][k] - s;
                }
                for (int i = k + 1; i < N; ++i) {
                    float s = a[i][0] * temp[0];
                    for (int j = 1; j < k; ++j) s = s + a[i][j] * temp[j];
                    a[i][k] = a[i][k] - s;
                }
            }
            const float akk = a[k][k];
            const bool pivot_ok = std::fabs(akk) > 0.0f;
            if (k == 0 && !pivot_ok) {
                for (int j = 0; j < N; ++j) tr[j] = j;
                return;
            }
            if (pivot_ok) {
                for (int i = k + 1; i < N; ++i) a[i][k] = a[i][k] / akk;
            }
        }
    }
    Mat<N, 1> solve(const Mat<N, 1> &b) const {
        float x[N];
        for (int i = 0; i < N; ++i) x[i] = b.d[i][0];
        for (int k = 0; k < N; ++k) {
            const float t = x[k];
            x[k] = x[tr[k]];
            x[tr[k]] = t;
        }
        for (int i = 1; i < N; ++i) {
            float s = a[i][0] * x[0];
            for (int j = 1; j < i; ++j) s = s + a[i][j] * x[j];
            x[i] = x[i] - s;
        }
        const float tol = std::numeric_limits<float>::min();
        for (int i = 0; i < N; ++i) x[i] = (std::fabs(a[i][i]) > tol) ? x[i] / a[i][i] : 0.0f;
        for (int i = N - 2; i >= 0; --i) {
            float s = a[i + 1][i] * x[i + 1];
            for (int j = i + 2; j < N; ++j) s = s + a[j][i] * x[j];
            x[i] = x[i] - s;
        }
        for (int k = N - 1; k >= 0; --k) {
            const float t = x[k];
            x[k] = x[tr[k]];
            x[tr[k]] = t;
        }
        Mat<N, 1> r;
        for (int i = 0; i < N; ++i) r.d[i][0] = x[i];
        return r;
    }
};

template <int R, int C>
Ldlt<R> Mat<R, C>::ldlt() const {
    static_assert(R == C, "ldlt needs a square matrix");
    return Ldlt<R>(*this);
}

// scalar * matrix (direct_method_tracker.cpp:176 `residual * jacobian`): each coefficient s * m(i,j).
template <int R, int C>
inline Mat<R, C> operator*(float s, const Mat<R, C> &m) {
    Mat<R, C> r;
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) r.d[i][j] = s * m.d[i][j];
    return r;
}

// Stand-in for Eigen::Quaternionf (direct_method_tracker.cpp only).  Restates Eigen 3.3/3.4's scalar code paths:
//   product            quat_product<..., false>::run
//   quat * vec3        QuaternionBase::_transformVector: uv = vec x v; uv += uv; v + w * uv + vec x uv
//   inverse            conjugate / squaredNorm (a zero quaternion stays zero)
//   normalize(d)       coeffs / norm, norm = sqrt(x*x + y*y + z*z + w*w) summed in coefficient order (x, y, z, w)
// Sums run left to right as everywhere in this shim.
struct Quat {
    float qw, qx, qy, qz;
    Quat() {}
    Quat(float w, float x, float y, float z) : qw(w), qx(x), qy(y), qz(z) {}
    static Quat Identity() { return Quat(1.0f, 0.0f, 0.0f, 0.0f); }
    float w() const { return qw; }
    float x() const { return qx; }
    float y() const { return qy; }
    float z() const { return qz; }
    float squaredNorm() const { return qx * qx + qy * qy + qz * qz + qw * qw; }
    float norm() const { return std::sqrt(squaredNorm()); }
    Quat inverse() const {
        const float n2 = squaredNorm();
        if (n2 > 0.0f) return Quat(qw / n2, -qx / n2, -qy / n2, -qz / n2);
        return Quat(0.0f, 0.0f, 0.0f, 0.0f);
    }
    Quat normalized() const {
        const float n = norm();
        return Quat(qw / n, qx / n, qy / n, qz / n);
    }
    void normalize() { *this = normalized(); }
    Quat operator*(const Quat &b) const {
        return Quat(qw * b.qw - qx * b.qx - qy * b.qy - qz * b.qz, qw * b.qx + qx * b.qw + qy * b.qz - qz * b.qy,
                    qw * b.qy + qy * b.qw + qz * b.qx - qx * b.qz, qw * b.qz + qz * b.qw + qx * b.qy - qy * b.qx);
    }
    Mat<3, 1> operator*(const Mat<3, 1> &v) const {
        const float vx = v.d[0][0], vy = v.d[1][0], vz = v.d[2][0];
        float ux = qy * vz - qz * vy, uy = qz * vx - qx * vz, uz = qx * vy - qy * vx;  // vec x v
        ux = ux + ux, uy = uy + uy, uz = uz + uz;
        Mat<3, 1> r;
        r.d[0][0] = vx + qw * ux + (qy * uz - qz * uy);
        r.d[1][0] = vy + qw * uy + (qz * ux - qx * uz);
        r.d[2][0] = vz + qw * uz + (qx * uy - qy * ux);
        return r;
    }
};

// Dynamic float matrix (Slam_Utility's `Mat` = Eigen::MatrixXf; dense_optical_flow.cpp only): setZero / resize / element
// access / rows / cols / scalar division.  Storage order does not matter to any result.
struct MatDyn {
    std::vector<float> v;
    int r = 0, c = 0;
    void setZero(int rows, int cols) {
        r = rows;
        c = cols;
        v.assign(static_cast<size_t>(rows) * cols, 0.0f);
    }
    void resize(int rows, int cols) {
        r = rows;
        c = cols;
        v.resize(static_cast<size_t>(rows) * cols);
    }
    int rows() const { return r; }
    int cols() const { return c; }
    float &operator()(int i, int j) { return v[static_cast<size_t>(i) * c + j]; }
    const float &operator()(int i, int j) const { return v[static_cast<size_t>(i) * c + j]; }
    MatDyn &operator/=(float s) {
        for (float &x : v) x = x / s;
        return *this;
    }
};

// Dynamic int matrix (only setConstant + element access are used, lssd_klt.cpp:136-137).
struct MatIntDyn {
    std::vector<int32_t> v;
    int r = 0, c = 0;
    void setConstant(int rows, int cols, int32_t value) {
        r = rows;
        c = cols;
        v.assign(static_cast<size_t>(rows) * cols, value);
    }
    int32_t &operator()(int i, int j) { return v[static_cast<size_t>(i) * c + j]; }
    const int32_t &operator()(int i, int j) const { return v[static_cast<size_t>(i) * c + j]; }
};

}  // namespace shim

namespace Eigen {
template <int R, int C>
inline shim::IsNanResult isnan(const shim::ArrayView<R, C> &a) {
    const float *p = &a.m.d[0][0];
    bool any = false;
    for (int i = 0; i < R * C; ++i) any = any || std::isnan(p[i]);
    return shim::IsNanResult{any};
}
}  // namespace Eigen

using Vec1 = shim::Mat<1, 1>;
using Vec2 = shim::Mat<2, 1>;
using Vec3 = shim::Mat<3, 1>;
using Vec6 = shim::Mat<6, 1>;
using Mat2 = shim::Mat<2, 2>;
using Mat3 = shim::Mat<3, 3>;
using Mat6 = shim::Mat<6, 6>;
using Mat2x3 = shim::Mat<2, 3>;
using Mat1x2 = shim::Mat<1, 2>;
using Mat1x3 = shim::Mat<1, 3>;
using Mat2x6 = shim::Mat<2, 6>;
using Quat = shim::Quat;
// Slam_Utility's "treat as zero" threshold (direct_method_tracker.cpp:129,140 depth tests).  Frozen here, like the rest of the shim.
constexpr float kZeroFloat = 1e-6f;
using MatInt = shim::MatIntDyn;
using Mat = shim::MatDyn;

#endif
