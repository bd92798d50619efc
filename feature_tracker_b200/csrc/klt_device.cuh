// Device-side building blocks shared by the KLT kernels.
//
// Arithmetic contract (DESIGN.md "Bit-exact parity"): every floating-point operation of the reference is executed
// as the same IEEE-754 binary32 round-to-nearest operation, without FMA contraction (explicit __fmul_rn / __fadd_rn
// / __fsub_rn / __fdiv_rn / __fsqrt_rn; the file is also compiled with -fmad=false), and every reduction over patch
// pixels is accumulated in the reference's row-major pixel order.  The parallelism comes from (a) evaluating the
// per-pixel terms of a chunk of pixels on different lanes and (b) giving every accumulator ("chain") of the normal
// equations its own lane, which then folds the chunk's terms sequentially.
#ifndef FTK_KLT_DEVICE_CUH_
#define FTK_KLT_DEVICE_CUH_

#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "ftk_internal.h"

namespace ftk {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// Correctly rounded a / b for many numerators and ONE divisor (the patch means of the LSSD tracker divide every pixel of a patch):
// y = RN(1 / b) once (one correctly rounded division), then per quotient q0 = RN(a * y), r = a - b * q0 (exact in one FMA),
// q = RN(q0 + r * y) -- Markstein's theorem: with a correctly rounded reciprocal and a faithful q0 the corrected quotient IS
// RN(a / b), i.e. bit-identical to the reference's division (3 instructions instead of the ~10 of a general IEEE division; checked
// against a / b on 1.5e9 random and adversarial operand pairs on the CPU, 0 mismatches).  The theorem needs results away from the
// subnormal range: a divisor outside [2^-20, 2^20] (never seen for pixel means in (0, 255]) takes the general division.
struct SharedDivisor {
    float b, y;
    bool fast;
};
__device__ __forceinline__ SharedDivisor MakeSharedDivisor(float divisor) {
    SharedDivisor d;
    d.b = divisor;
    d.y = __fdiv_rn(1.0f, divisor);
    d.fast = divisor >= 9.5367431640625e-07f && divisor <= 1048576.0f;
    return d;
}
// |a| is 0 or in [2^-100, 2^100] for every caller (bilinear samples of 8-bit pixels and their differences).  FAST is hoisted out of
// the pixel loops by the callers (both divisors in range), so the loop bodies carry no branch and no general-division slow path.
template <bool FAST>
__device__ __forceinline__ float DivideBy(const SharedDivisor &d, float a) {
    if (!FAST) return __fdiv_rn(a, d.b);
    const float q0 = __fmul_rn(a, d.y);
    const float r = __fmaf_rn(-d.b, q0, a);
    return __fmaf_rn(r, d.y, q0);
}

// ---- images ---------------------------------------------------------------------------------------------------
struct Img {
    const uint8_t *p;
    int rows, cols, pitch;
};

__device__ __forceinline__ Img LevelImage(const PyramidView &v, int image, int level) {
    Img im;
    im.p = v.base[level] + image * v.image_stride[level];
    // Opaque to the optimiser: otherwise ptxas re-derives this 64-bit base from the kernel parameters inside every pixel loop
    // (~20 integer instructions per chunk) instead of keeping two registers live.
    asm("" : "+l"(im.p));
    im.rows = v.rows[level];
    im.cols = v.cols[level];
    im.pitch = v.pitch[level];
    return im;
}

// uint8 pixel -> float.  cvt.rn.f32.s32 (I2FP, ALU pipe, 64 lanes/clk/SM) instead of the I2F.U16 (XU pipe, 16 lanes/clk/SM)
// the compiler picks for a byte source: profiles/r1_microbench_pipe_rates.txt.
__device__ __forceinline__ float PxToFloat(const uint8_t *p) {
    const int v = __ldg(p);
    float f;
    asm("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(v));
    return f;
}

__device__ __forceinline__ float PxI(const Img &im, int row, int col) { return PxToFloat(im.p + row * im.pitch + col); }

// GrayImage::GetPixelValueNoCheck(float,float) (oracle/shim/datatype_image.h): base pixel by truncation, fractions by
// floor, ((ic*ir)*p00 + (sc*ir)*p01) + (ic*sr)*p10) + (sc*sr)*p11.
__device__ __forceinline__ float PxF(const Img &im, float row, float col) {
    // Callers only sample positions inside the image (0 <= row, col < 2^24): there truncation == floor, so one F2I.FLOOR per
    // axis yields both the base pixel and (converted back on the ALU pipe) the value floor() would return.
    const int r = __float2int_rd(row), c = __float2int_rd(col);
    const uint8_t *v = im.p + r * im.pitch + c;
    float fr, fc;
    asm("cvt.rn.f32.s32 %0, %1;" : "=f"(fr) : "r"(r));
    asm("cvt.rn.f32.s32 %0, %1;" : "=f"(fc) : "r"(c));
    const float sr = fsub(row, fr);
    const float sc = fsub(col, fc);
    const float ir = fsub(1.0f, sr);
    const float ic = fsub(1.0f, sc);
    const float p00 = PxToFloat(v);
    const float p01 = PxToFloat(v + 1);
    const float p10 = PxToFloat(v + im.pitch);
    const float p11 = PxToFloat(v + im.pitch + 1);
    return fadd(fadd(fadd(fmul(fmul(ic, ir), p00), fmul(fmul(sc, ir), p01)), fmul(fmul(ic, sr), p10)), fmul(fmul(sc, sr), p11));
}

// GrayImage::GetPixelValueNoCheck(float, float) at a position that nobody bounds-checked (lssd_klt_fast.cpp:182-193: the
// "patch totally inside" test looks at the un-rotated bounding box only, so a diverged R_cr sends samples anywhere).  The
// reference then reads data[(int)row * cols + (int)col + {0, 1, cols, cols + 1}] from its tightly packed image: inside the
// image that wraps into neighbouring rows, outside it is undefined behaviour (garbage or a crash).  Here: positions whose 2x2
// footprint is inside take the normal path; the others reproduce the linear addressing for indices inside the image and read
// 0 elsewhere -- never an address outside the pyramid.
__device__ __forceinline__ float PxFUnchecked(const Img &im, float row, float col) {
    const int r = __float2int_rz(row), c = __float2int_rz(col);  // static_cast<int32_t>: truncation
    if (row >= 0.0f && col >= 0.0f && r < im.rows - 1 && c < im.cols - 1) return PxF(im, row, col);
    const float sr = fsub(row, floorf(row)), sc = fsub(col, floorf(col));
    const float ir = fsub(1.0f, sr), ic = fsub(1.0f, sc);
    const long long n = static_cast<long long>(im.rows) * im.cols, base = static_cast<long long>(r) * im.cols + c;
    float p[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const long long l = base + (q & 1) + (q >> 1) * im.cols;
        p[q] = 0.0f;
        if (l >= 0 && l < n) {
            const int rr = static_cast<int>(l / im.cols), cc = static_cast<int>(l - static_cast<long long>(rr) * im.cols);
            p[q] = PxToFloat(im.p + rr * im.pitch + cc);
        }
    }
    return fadd(fadd(fadd(fmul(fmul(ic, ir), p[0]), fmul(fmul(sc, ir), p[1])), fmul(fmul(ic, sr), p[2])), fmul(fmul(sc, sr), p[3]));
}

// GrayImage::GetPixelValue(float,float,float*): fails iff outside [0, cols-1] x [0, rows-1].
__device__ __forceinline__ bool PxInside(const Img &im, float row, float col) {
    return !(col < 0.0f || row < 0.0f || col > static_cast<float>(im.cols - 1) || row > static_cast<float>(im.rows - 1));
}
__device__ __forceinline__ bool PxChecked(const Img &im, float row, float col, float *out) {
    if (!PxInside(im, row, col)) return false;
    *out = PxF(im, row, col);
    return true;
}

// The gradient stencil of the direct / inverse methods: five GetPixelValue calls at (row, col -/+ 1), (row -/+ 1, col) and
// (row, col) (basic_klt.cpp:127-134, affine_klt.cpp:140-147, lssd_klt.cpp:202-207).  False when any of them is outside.
// col - 1 and row - 1 are exact for coordinates >= 1 (guaranteed once their bounds tests pass); when col + 1 and row + 1 are
// exact too (always, except where the coordinate crosses a binade) the five positions have the same fractions, hence the
// same four weight products, and overlap in 12 distinct pixels.  Every sample is still the reference's own sequence of
// rounded operations, ((ic*ir)*p00 + (sc*ir)*p01) + (ic*sr)*p10) + (sc*sr)*p11; only the common sub-expressions are shared.
__device__ __forceinline__ bool PxStencil5(const Img &im, float row, float col, float *left, float *right, float *up, float *down, float *centre) {
    const float cm = fsub(col, 1.0f), cp = fadd(col, 1.0f), rm = fsub(row, 1.0f), rp = fadd(row, 1.0f);
    // The five GetPixelValue bounds tests (each !(col < 0 || row < 0 || col > cols - 1 || row > rows - 1)) in four comparisons:
    // cm <= col <= cp and rm <= row <= rp for every float (rounding is monotonic), so the extreme coordinates decide -- cm < 0 covers
    // col < 0 and cp < 0, cp > cols - 1 covers col > and cm >, likewise for the rows; with a NaN coordinate every comparison of either
    // form is false and both forms say "inside", as the reference does.
    if (cm < 0.0f || cp > static_cast<float>(im.cols - 1) || rm < 0.0f || rp > static_cast<float>(im.rows - 1)) return false;
    const int r = __float2int_rd(row), c = __float2int_rd(col);
    float fr, fc;
    asm("cvt.rn.f32.s32 %0, %1;" : "=f"(fr) : "r"(r));
    asm("cvt.rn.f32.s32 %0, %1;" : "=f"(fc) : "r"(c));
    const float sr = fsub(row, fr), sc = fsub(col, fc);
    const float ir = fsub(1.0f, sr), ic = fsub(1.0f, sc);
    const float w00 = fmul(ic, ir), w01 = fmul(sc, ir), w10 = fmul(ic, sr), w11 = fmul(sc, sr);
    const uint8_t *v = im.p + r * im.pitch + c;
    const uint8_t *vm = v - im.pitch, *vp = v + im.pitch, *vq = vp + im.pitch;
    const float a0 = PxToFloat(vm), a1 = PxToFloat(vm + 1);
    const float b0 = PxToFloat(v - 1), b1 = PxToFloat(v), b2 = PxToFloat(v + 1), b3 = PxToFloat(v + 2);
    const float c0 = PxToFloat(vp - 1), c1 = PxToFloat(vp), c2 = PxToFloat(vp + 1), c3 = PxToFloat(vp + 2);
    const float d0 = PxToFloat(vq), d1 = PxToFloat(vq + 1);
    *left = fadd(fadd(fadd(fmul(w00, b0), fmul(w01, b1)), fmul(w10, c0)), fmul(w11, c1));
    *up = fadd(fadd(fadd(fmul(w00, a0), fmul(w01, a1)), fmul(w10, b1)), fmul(w11, b2));
    *centre = fadd(fadd(fadd(fmul(w00, b1), fmul(w01, b2)), fmul(w10, c1)), fmul(w11, c2));
    float rv = fadd(fadd(fadd(fmul(w00, b2), fmul(w01, b3)), fmul(w10, c2)), fmul(w11, c3));
    float dv = fadd(fadd(fadd(fmul(w00, c1), fmul(w01, c2)), fmul(w10, d0)), fmul(w11, d1));
    if (fsub(cp, 1.0f) != col) rv = PxF(im, row, cp);  // col + 1 was rounded: its own fraction (and possibly base pixel)
    if (fsub(rp, 1.0f) != row) dv = PxF(im, rp, col);
    *right = rv;
    *down = dv;
    return true;
}

__device__ __forceinline__ bool IsOutside(const Img &im, float x, float y) {
    return x < 0.0f || x > static_cast<float>(im.cols - 1) || y < 0.0f || y > static_cast<float>(im.rows - 1);
}

// ---- lane groups: G lanes of a warp cooperate on one feature ---------------------------------------------------
template <int G>
struct Group {
    int lane;        // 0..G-1
    int base;        // first lane of the group inside the warp
    unsigned mask;   // mask of the warp-level operations: the group's lanes, or the whole warp (see `whole_warp`)
    unsigned lanes;  // the group's lanes
    // whole_warp: every group of the warp runs the same instruction stream (the kernel guarantees it: finished groups keep executing
    // with their results discarded), so barriers / votes / shuffles may name the full warp.  ptxas then emits the plain instruction; a
    // per-lane sub-warp mask costs a REDUX.OR + R2UR + BRA.DIV uniformity check around every one of them (3 % of the affine kDirect
    // kernel's instructions).
    __device__ __forceinline__ explicit Group(bool whole_warp = false) {
        const int l = threadIdx.x & 31;
        lane = l % G;
        base = l - lane;
        lanes = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << base);
        mask = whole_warp ? 0xFFFFFFFFu : lanes;
    }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ int count(bool pred) const { return __popc(__ballot_sync(mask, pred) & lanes); }
    __device__ __forceinline__ float get(float v, int src) const { return __shfl_sync(mask, v, base + src); }
    // true when `pred` holds on every group that takes part in the group's warp-level operations
    __device__ __forceinline__ bool all(bool pred) const { return __all_sync(mask, pred); }
};

// ---- chains: K accumulators, one per lane, folded sequentially over the G terms of a chunk -----------------------
// Shared layout: term[k * (G + 4) + lane]; the +4 float padding keeps the float4 reads of lanes 0..K-1 on distinct
// banks.
template <int G>
struct Chain {
    static constexpr int kStride = G + 4;
    float *term;
    float acc;     // chain `lane`
    float acc_hi;  // chain `G + lane` when a pass has more chains than the group has lanes (fold_wide)
    __device__ __forceinline__ void reset() { acc = 0.0f, acc_hi = 0.0f; }
    __device__ __forceinline__ void put(int lane, int k, float v) const { term[k * kStride + lane] = v; }
    // All lanes call; lanes >= K idle during the fold.
    template <int K>
    __device__ __forceinline__ void fold(const Group<G> &g) {
        g.sync();
        if (g.lane < K) {
            const float4 *t4 = reinterpret_cast<const float4 *>(term + g.lane * kStride);
#pragma unroll
            for (int q = 0; q < G / 4; ++q) {
                const float4 v = t4[q];
                acc = fadd(acc, v.x);
                acc = fadd(acc, v.y);
                acc = fadd(acc, v.z);
                acc = fadd(acc, v.w);
            }
        }
        g.sync();
    }
    // fold<K> for chains FIRST .. FIRST + K - 1 whose terms were stored in slots 0 .. K - 1: lane FIRST + k folds slot k (a pass that sends its
    // chains through a buffer of K chains in several parts)
    template <int K, int FIRST>
    __device__ __forceinline__ void fold_at(const Group<G> &g) {
        static_assert(FIRST + K <= G, "one chain per lane");
        g.sync();
        if (g.lane >= FIRST && g.lane < FIRST + K) {
            const float4 *t4 = reinterpret_cast<const float4 *>(term + (g.lane - FIRST) * kStride);
#pragma unroll
            for (int q = 0; q < G / 4; ++q) {
                const float4 v = t4[q];
                acc = fadd(acc, v.x);
                acc = fadd(acc, v.y);
                acc = fadd(acc, v.z);
                acc = fadd(acc, v.w);
            }
        }
        g.sync();
    }
    // fold<K> into the second accumulator: chains G .. G + K - 1 of a pass whose terms were stored in slots 0 .. K - 1 (a pass with more
    // chains than lanes that reuses a G-chain buffer instead of holding all of its terms at once)
    template <int K>
    __device__ __forceinline__ void fold_hi(const Group<G> &g) {
        g.sync();
        if (g.lane < K) {
            const float4 *t4 = reinterpret_cast<const float4 *>(term + g.lane * kStride);
#pragma unroll
            for (int q = 0; q < G / 4; ++q) {
                const float4 v = t4[q];
                acc_hi = fadd(acc_hi, v.x);
                acc_hi = fadd(acc_hi, v.y);
                acc_hi = fadd(acc_hi, v.z);
                acc_hi = fadd(acc_hi, v.w);
            }
        }
        g.sync();
    }
    // ---- paired layout: chain k and chain G + k travel together (affine kDirect / kInverse on 16 lanes: 27 chains) ----------------
    // term2[k * kPairStride + 2 * pixel + {0, 1}] = term of chain k / chain G + k.  A pixel lane writes one 64-bit store per chain pair
    // (16 instead of 27 stores), a chain lane reads its row as float4s = two pixels of both chains, and advances both accumulators with
    // ONE packed add per pixel (Blackwell FADD2, add.rn.f32x2: two independent IEEE round-to-nearest adds, so each chain still sees
    // exactly its own sequence of sums).  Row stride 2 G + 4 floats: the float4 reads of lanes 0..7 fall on distinct banks.
    static constexpr int kPairStride = 2 * G + 4;
    __device__ __forceinline__ void put_pair(int lane, int k, float lo, float hi) const {
        *reinterpret_cast<float2 *>(term + k * kPairStride + 2 * lane) = make_float2(lo, hi);
    }
    // All lanes call; chain pairs 0..G-1 (chains G + k beyond the real count accumulate the zeros the pixel lanes stored for them).
    __device__ __forceinline__ void fold_pairs(const Group<G> &g) {
        g.sync();
        {
            const float4 *t4 = reinterpret_cast<const float4 *>(term + g.lane * kPairStride);
            unsigned long long a2;
            asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "f"(acc), "f"(acc_hi));
#pragma unroll
            for (int q = 0; q < G / 2; ++q) {
                const float4 v = t4[q];  // pixel 2q: (chain lane, chain G + lane), pixel 2q + 1: the same two chains
                unsigned long long p0, p1;
                asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(v.x), "f"(v.y));
                asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(v.z), "f"(v.w));
                asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a2) : "l"(p0));
                asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a2) : "l"(p1));
            }
            asm("mov.b64 {%0, %1}, %2;" : "=f"(acc), "=f"(acc_hi) : "l"(a2));
        }
        g.sync();
    }
    // K chains with G < K <= 2 G: lanes fold chains 0..G-1, then lanes 0..K-G-1 fold chains G..K-1.
    template <int K>
    __device__ __forceinline__ void fold_wide(const Group<G> &g) {
        if constexpr (K <= G) {
            fold<K>(g);
        } else {
            static_assert(K <= 2 * G, "at most two chains per lane");
            g.sync();
            {
                const float4 *t4 = reinterpret_cast<const float4 *>(term + g.lane * kStride);
#pragma unroll
                for (int q = 0; q < G / 4; ++q) {
                    const float4 v = t4[q];
                    acc = fadd(acc, v.x);
                    acc = fadd(acc, v.y);
                    acc = fadd(acc, v.z);
                    acc = fadd(acc, v.w);
                }
            }
            if (g.lane < K - G) {
                const float4 *t4 = reinterpret_cast<const float4 *>(term + (G + g.lane) * kStride);
#pragma unroll
                for (int q = 0; q < G / 4; ++q) {
                    const float4 v = t4[q];
                    acc_hi = fadd(acc_hi, v.x);
                    acc_hi = fadd(acc_hi, v.y);
                    acc_hi = fadd(acc_hi, v.z);
                    acc_hi = fadd(acc_hi, v.w);
                }
            }
            g.sync();
        }
    }
};

// ---- LDLT (Eigen LDLT<Lower> restated; oracle/shim/basic_type.h) ----------------------------------
// Executed redundantly by every lane of a group on identical inputs (uniform control flow, no broadcast needed).
// Factorisation and solve are separate so that a Hessian that stays constant over the iterations of a level (the fast
// methods) is factorised once; the reference calls hessian.ldlt() every iteration, which yields the same factors each time.
template <int N>
struct LdltFactors {
    float a[N][N];
    int tr[N];
};

template <int N>
__device__ __forceinline__ void LdltFactor(const float (&A)[N][N], LdltFactors<N> &f) {
    float(&a)[N][N] = f.a;
    int(&tr)[N] = f.tr;
    float temp[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) a[i][j] = A[i][j];

    bool zero_diag = false;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (zero_diag) break;
        int p = k;
        float best = fabsf(a[k][k]);
        for (int i = k + 1; i < N; ++i) {
            const float v = fabsf(a[i][i]);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        tr[k] = p;
        if (p != k) {
            for (int j = 0; j < k; ++j) {
                const float t = a[k][j];
                a[k][j] = a[p][j];
                a[p][j] = t;
            }
            for (int i = p + 1; i < N; ++i) {
                const float t = a[i][k];
                a[i][k] = a[i][p];
                a[i][p] = t;
            }
            {
                const float t = a[k][k];
                a[k][k] = a[p][p];
                a[p][p] = t;
            }
            for (int i = k + 1; i < p; ++i) {
                const float t = a[i][k];
                a[i][k] = a[p][i];
                a[p][i] = t;
            }
        }
        if (k > 0) {
            for (int j = 0; j < k; ++j) temp[j] = fmul(a[j][j], a[k][j]);
            float s = fmul(a[k][0], temp[0]);
            for (int j = 1; j < k; ++j) s = fadd(s, fmul(a[k][j], temp[j]));
            a[k][k] = fsub(a[k][k], s);
            for (int i = k + 1; i < N; ++i) {
                float t = fmul(a[i][0], temp[0]);
                for (int j = 1; j < k; ++j) t = fadd(t, fmul(a[i][j], temp[j]));
                a[i][k] = fsub(a[i][k], t);
            }
        }
        const float akk = a[k][k];
        const bool pivot_ok = fabsf(akk) > 0.0f;
        if (k == 0 && !pivot_ok) {
            for (int j = 0; j < N; ++j) tr[j] = j;
            zero_diag = true;
        } else if (pivot_ok) {
            for (int i = k + 1; i < N; ++i) a[i][k] = fdiv(a[i][k], akk);
        }
    }
}

template <int N>
__device__ __forceinline__ void LdltSolveFactored(const LdltFactors<N> &f, const float (&b)[N], float (&x)[N]) {
    const float(&a)[N][N] = f.a;
    const int(&tr)[N] = f.tr;
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = b[i];
    for (int k = 0; k < N; ++k) {
        const float t = x[k];
        x[k] = x[tr[k]];
        x[tr[k]] = t;
    }
    for (int i = 1; i < N; ++i) {
        float s = fmul(a[i][0], x[0]);
        for (int j = 1; j < i; ++j) s = fadd(s, fmul(a[i][j], x[j]));
        x[i] = fsub(x[i], s);
    }
    for (int i = 0; i < N; ++i) x[i] = (fabsf(a[i][i]) > FLT_MIN) ? fdiv(x[i], a[i][i]) : 0.0f;
    for (int i = N - 2; i >= 0; --i) {
        float s = fmul(a[i + 1][i], x[i + 1]);
        for (int j = i + 2; j < N; ++j) s = fadd(s, fmul(a[j][i], x[j]));
        x[i] = fsub(x[i], s);
    }
    for (int k = N - 1; k >= 0; --k) {
        const float t = x[k];
        x[k] = x[tr[k]];
        x[tr[k]] = t;
    }
}

// ---- cooperative LDLT<N> in shared memory (affine: N = 6, LSSD: N = 3) ------------------------------------------------
// The register version above indexes its arrays with the data-dependent pivot, which puts them in local memory (the affine
// kernels carried ~780 LDL/STL) and makes every lane of a group repeat ~1300 instructions per 6 x 6 solve.  Here the lower
// triangle lives in the group's shared scratch and lane i owns row i: pivot search, symmetric swap, column update and scaling of
// one elimination step run on N lanes at once.  Every matrix element goes through exactly the operations of LdltFactor, in the
// same order, so factors, transpositions and solutions are bit-identical to it (tests: the trackers against the oracle).
template <int N>
struct LdltShared {
    float a[N * N];  // row-major, lower triangle + diagonal used
    int perm[8];     // x_permuted[i] = b[perm[i]] (the forward transpositions applied to the identity)
    float b[8];      // right-hand side, written by the lanes that own the bias chains
    float z[8];      // solution after the backward transpositions
};
using Ldlt6Shared = LdltShared<6>;
constexpr int kLdlt6Floats = 36 + 8 + 8 + 8;
constexpr int kLdlt3Floats = 9 + 8 + 8 + 8;

// Factorises s.a in place; fills s.perm.  All lanes of the group call (uniform control flow inside the group).
template <int N, int G>
__device__ __forceinline__ void LdltFactorShared(const Group<G> &g, LdltShared<N> &s) {
    float *a = s.a;
    const int lane = g.lane;
    int tr[N];
    bool zero_diag = false;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        // After a zero first pivot (zero_diag) the reference stops: the remaining steps then change nothing, but still execute their
        // barriers -- the groups of a warp may share them (Group::mask), so every group must pass the same number.
        int p = k;
        float best = fabsf(a[k * N + k]);
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const float v = fabsf(a[i * N + i]);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        if (zero_diag) p = k;
        tr[k] = p;
        g.sync();  // every lane has read the diagonal
        {
            // symmetric exchange of rows / columns k and p of the lower triangle: one element pair per lane (p == k: no-ops)
            int ia = -1, ib = -1;
            if (lane < k) ia = k * N + lane, ib = p * N + lane;
            else if (lane == k) ia = k * N + k, ib = p * N + p;
            else if (lane < p) ia = lane * N + k, ib = p * N + lane;
            else if (lane > p && lane < N) ia = lane * N + k, ib = lane * N + p;
            if (ia >= 0) {
                const float t = a[ia];
                a[ia] = a[ib];
                a[ib] = t;
            }
        }
        g.sync();
        if (k > 0) {
            if (!zero_diag && lane >= k && lane < N) {
                // lane == k: a[k][k] -= sum_j a[k][j] * temp[j]; lane > k: a[i][k] -= sum_j a[i][j] * temp[j]; temp[j] = a[j][j] * a[k][j]
                float sum = fmul(a[lane * N + 0], fmul(a[0 * N + 0], a[k * N + 0]));
#pragma unroll
                for (int j = 1; j < k; ++j) sum = fadd(sum, fmul(a[lane * N + j], fmul(a[j * N + j], a[k * N + j])));
                a[lane * N + k] = fsub(a[lane * N + k], sum);
            }
            g.sync();
        }
        const float akk = a[k * N + k];
        const bool pivot_ok = fabsf(akk) > 0.0f;
        if (k == 0 && !pivot_ok) {
            tr[0] = 0;
            zero_diag = true;
        } else if (pivot_ok && !zero_diag) {
            if (lane > k && lane < N) a[lane * N + k] = fdiv(a[lane * N + k], akk);
        }
        g.sync();
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) s.perm[i] = i;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int t = s.perm[k];
            s.perm[k] = s.perm[tr[k]];
            s.perm[tr[k]] = t;
        }
    }
    g.sync();
}

// Solves with the factors of LdltFactorShared and the right-hand side s.b; every lane gets the solution.
template <int N, int G>
__device__ __forceinline__ void LdltSolveShared(const Group<G> &g, LdltShared<N> &s, float (&z)[N]) {
    const float *a = s.a;
    float x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = s.b[s.perm[i]];
#pragma unroll
    for (int i = 1; i < N; ++i) {
        float sum = fmul(a[i * N + 0], x[0]);
#pragma unroll
        for (int j = 1; j < i; ++j) sum = fadd(sum, fmul(a[i * N + j], x[j]));
        x[i] = fsub(x[i], sum);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = (fabsf(a[i * N + i]) > FLT_MIN) ? fdiv(x[i], a[i * N + i]) : 0.0f;
#pragma unroll
    for (int i = N - 2; i >= 0; --i) {
        float sum = fmul(a[(i + 1) * N + i], x[i + 1]);
#pragma unroll
        for (int j = i + 2; j < N; ++j) sum = fadd(sum, fmul(a[j * N + i], x[j]));
        x[i] = fsub(x[i], sum);
    }
    // backward transpositions = the inverse permutation (every lane writes the same values)
#pragma unroll
    for (int i = 0; i < N; ++i) s.z[s.perm[i]] = x[i];
    g.sync();
#pragma unroll
    for (int i = 0; i < N; ++i) z[i] = s.z[i];
    g.sync();
}

template <int G>
__device__ __forceinline__ void Ldlt6FactorShared(const Group<G> &g, Ldlt6Shared &s) { LdltFactorShared<6, G>(g, s); }
template <int G>
__device__ __forceinline__ void Ldlt6SolveShared(const Group<G> &g, Ldlt6Shared &s, float (&z)[6]) { LdltSolveShared<6, G>(g, s, z); }

template <int N>
__device__ __forceinline__ void LdltSolve(const float (&A)[N][N], const float (&b)[N], float (&x)[N]) {
    LdltFactors<N> f;
    LdltFactor<N>(A, f);
    LdltSolveFactored<N>(f, b, x);
}

}  // namespace ftk

#endif
