// Helpers shared by the 16-lanes-per-feature basic-KLT kernels (klt_basic_fastpath.cu).
#ifndef FTK_KLT_FAST_COMMON_CUH_
#define FTK_KLT_FAST_COMMON_CUH_

#include "klt_device.cuh"

namespace ftk {
namespace fastk {

constexpr int kG = 16;
constexpr int kThreads = 128;
constexpr int kGroupsPerBlock = kThreads / kG;

// Separable part of a bilinear sample along one axis.
struct __align__(16) Entry {
    int base;  // floor(x) as an integer (limited to [-64, n + 64]); for in-bounds x it is the reference's static_cast<int32_t>(x)
    float s;   // fraction  (x - floor(x))
    float i;   // 1 - fraction
    int off;   // row tables: byte offset (clamped, always addressable) of the image row the rolling strip loads next
};

// ok = !(x < 0 || x > n - 1): the GrayImage::GetPixelValue bounds test.
__device__ __forceinline__ Entry MakeEntry(float x, int n, bool *ok) {
    Entry e;
    const float f = floorf(x);
    e.s = fsub(x, f);
    e.i = fsub(1.0f, e.s);
    *ok = !(x < 0.0f || x > static_cast<float>(n - 1));
    e.base = min(max(static_cast<int>(f), -64), n + 64);
    e.off = 0;
    return e;
}

__device__ __forceinline__ int Clamp(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ((ic*ir)*p00 + (sc*ir)*p01) + (ic*sr)*p10) + (sc*sr)*p11 -- GrayImage::GetPixelValueNoCheck(float, float).
__device__ __forceinline__ float Bilerp(const Entry &R, const Entry &C, float p00, float p01, float p10, float p11) {
    return fadd(fadd(fadd(fmul(fmul(C.i, R.i), p00), fmul(fmul(C.s, R.i), p01)), fmul(fmul(C.i, R.s), p10)), fmul(fmul(C.s, R.s), p11));
}

__device__ __forceinline__ float LoadPx(const uint8_t *p) {
    // cvt.rn.f32.s32 keeps the conversion on the ALU pipe (I2FP) instead of the quarter-rate XU pipe (I2F.U16).
    const int v = __ldg(p);
    float f;
    asm("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(v));
    return f;
}

// A sample addressed through (row entry, column entry): loads its four bytes directly.  Out-of-image entries are
// clamped to an addressable pixel; their value is never used (the pixel's validity bit is clear).
__device__ __forceinline__ float SampleDirect(const Img &im, const Entry &R, const Entry &C) {
    const uint8_t *p = im.p + Clamp(R.base, 0, im.rows - 1) * im.pitch + Clamp(C.base, 0, im.cols - 1);
    return Bilerp(R, C, LoadPx(p), LoadPx(p + 1), LoadPx(p + im.pitch), LoadPx(p + im.pitch + 1));
}

// The two 16-lane groups of a warp always execute the same instruction stream (a finished or idle group keeps
// computing on clamped, addressable data and simply discards its results), so every vote / shuffle / barrier below is
// a cheap full-warp one.
constexpr unsigned kFull = 0xFFFFFFFFu;

struct Lanes {
    int lane;            // 0..15 inside the group
    int base;            // 0 or 16
    unsigned mask;       // the group's 16 lanes
    // bit r of the result = pred of group lane r
    __device__ __forceinline__ unsigned bits(bool pred) const { return (__ballot_sync(kFull, pred) >> base) & 0xFFFFu; }
    __device__ __forceinline__ bool any(bool pred) const { return (__ballot_sync(kFull, pred) & mask) != 0u; }
    __device__ __forceinline__ float get(float v, int src) const { return __shfl_sync(kFull, v, base + src); }
    __device__ __forceinline__ int sum(int v) const {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        return v;
    }
};

}  // namespace fastk
}  // namespace ftk

#endif
