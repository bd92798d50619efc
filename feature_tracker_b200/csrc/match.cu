// K5-K7 (first generation): descriptor matching on sm_100a.
//
// Replaces DescriptorMatcher<T>::ForceMatch / NearbyMatch (src/descriptor_matcher/descriptor_matcher.h:55-79,
// 90-124) together with the user-supplied ComputeDistance bodies the reference's demos use:
//   BRIEF   : number of differing elements      (test/test_descriptor_matcher_brief.cpp:33-45)
//   float   : 0.5 - dot / |a| / |b| * 0.5       (test/test_descriptor_matcher_superpoint.cpp:32-34, ..._disk.cpp:32-34)
//
// Result contract (bit-exact): idx[i] = the LOWEST j minimising d(i, j) among the admissible j, provided that
// minimum is < max_dist; otherwise idx[i] keeps its previous content.  "Admissible" is every j for ForceMatch and
// the window-gated j for NearbyMatch.  The reference's `break` on d == 0 (descriptor_matcher.h:119) is honoured
// exactly: candidates after the first admissible j with d == 0 are ignored (this only matters when a float
// distance rounds below zero).
//
// Kernels:
//   HammingForceKernel : XOR + POPC, ref descriptor in registers, cur descriptors streamed through shared memory as
//                        warp-wide broadcasts; packed (distance << 20 | j) keys make "lowest j wins ties" a plain
//                        integer min; the cur dimension is split across CTAs and merged with atomicMin.
//   Nearby*            : cur features are bucketed into a uniform grid (count -> scan -> fill); one warp per ref
//                        descriptor walks the cells overlapping its window, applies the reference's exact fp32 gate
//                        and reduces (distance, j) lexicographically.
//   CosineForceKernel  : exact fp32 distances with the reference's sequential k = 0..dim-1 summation order
//                        (one chain per (ref, cur) pair, 4 independent pairs per thread for ILP).
#include <cfloat>
#include <cstring>

#include "ftk_internal.h"

namespace ftk {

namespace {

constexpr unsigned kNoKey32 = 0xFFFFFFFFu;
constexpr unsigned long long kNoKey64 = 0xFFFFFFFFFFFFFFFFull;
constexpr int kJBits = 20;  // packed 32-bit keys: j < 2^20, distance < 2^12

__global__ void FillU32Kernel(unsigned *p, int n, unsigned v) {
    GridDepLaunchDependents();
    GridDepWait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void FillU64Kernel(unsigned long long *p, int n, unsigned long long v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// Hamming force matching
// ---------------------------------------------------------------------------------------------------------------
constexpr int kHamThreads = 256;
constexpr int kHamTile = 256;  // cur descriptors per shared-memory tile

// Population count of several words with fewer POPC instructions.  POPC runs on the XU pipe (16 lanes/clk/SM), LOP3 on the ALU pipe
// (64 lanes/clk/SM, profiles/r2_microbench_pipe_rates.txt), and the force kernel was XU bound (93.5 % busy): a carry-save adder
// (sum = a ^ b ^ c, carry = majority(a, b, c): one LOP3 each) turns three words of weight 1 into one of weight 1 and one of weight 2,
// so popc(x0..x7) = popc(s3) + popc(x7) + 2 * (popc(c1) + popc(c2) + popc(c3)) needs 5 POPC + 6 LOP3 instead of 8 POPC -- the point
// where both pipes take about the same time.  Integer arithmetic: the distance is the same number.
__device__ __forceinline__ unsigned Maj3(unsigned a, unsigned b, unsigned c) { return (a & b) | (c & (a ^ b)); }
__device__ __forceinline__ unsigned Popc8(unsigned x0, unsigned x1, unsigned x2, unsigned x3, unsigned x4, unsigned x5, unsigned x6, unsigned x7) {
    const unsigned s1 = x0 ^ x1 ^ x2, c1 = Maj3(x0, x1, x2);
    const unsigned s2 = x3 ^ x4 ^ x5, c2 = Maj3(x3, x4, x5);
    const unsigned s3 = s1 ^ s2 ^ x6, c3 = Maj3(s1, s2, x6);
    return __popc(s3) + __popc(x7) + 2u * (__popc(c1) + __popc(c2) + __popc(c3));
}
__device__ __forceinline__ unsigned Popc4(unsigned x0, unsigned x1, unsigned x2, unsigned x3) {
    return __popc(x0 ^ x1 ^ x2) + __popc(x3) + 2u * __popc(Maj3(x0, x1, x2));
}

template <int W>
__global__ void __launch_bounds__(kHamThreads) HammingForceKernel(const uint32_t *__restrict__ ref, int n_ref, const uint32_t *__restrict__ cur, int n_cur,
                                                                 int cur_per_split, unsigned *__restrict__ best) {
    __shared__ __align__(16) uint32_t tile[kHamTile * W];
    GridDepLaunchDependents();
    GridDepWait();
    const int i = blockIdx.x * kHamThreads + threadIdx.x;
    const int j_begin = blockIdx.y * cur_per_split;
    const int j_end = min(n_cur, j_begin + cur_per_split);

    uint32_t r[W];
    if (i < n_ref) {
#pragma unroll
        for (int w = 0; w < W; ++w) r[w] = __ldg(ref + static_cast<size_t>(i) * W + w);
    } else {
#pragma unroll
        for (int w = 0; w < W; ++w) r[w] = 0u;
    }

    unsigned key = kNoKey32;
    for (int j0 = j_begin; j0 < j_end; j0 += kHamTile) {
        const int n_tile = min(kHamTile, j_end - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < n_tile * W; t += kHamThreads) tile[t] = __ldg(cur + static_cast<size_t>(j0) * W + t);
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < n_tile; ++t) {
            unsigned d = 0;
            if constexpr (W % 8 == 0) {
#pragma unroll
                for (int w = 0; w < W; w += 8) {
                    const uint4 c = *reinterpret_cast<const uint4 *>(&tile[t * W + w]), e = *reinterpret_cast<const uint4 *>(&tile[t * W + w + 4]);
                    d += Popc8(r[w] ^ c.x, r[w + 1] ^ c.y, r[w + 2] ^ c.z, r[w + 3] ^ c.w, r[w + 4] ^ e.x, r[w + 5] ^ e.y, r[w + 6] ^ e.z, r[w + 7] ^ e.w);
                }
            } else if constexpr (W % 4 == 0) {
#pragma unroll
                for (int w = 0; w < W; w += 4) {
                    const uint4 c = *reinterpret_cast<const uint4 *>(&tile[t * W + w]);
                    d += Popc4(r[w] ^ c.x, r[w + 1] ^ c.y, r[w + 2] ^ c.z, r[w + 3] ^ c.w);
                }
            } else {
#pragma unroll
                for (int w = 0; w < W; ++w) d += __popc(r[w] ^ tile[t * W + w]);
            }
            key = min(key, (d << kJBits) | static_cast<unsigned>(j0 + t));
        }
    }
    if (i < n_ref && key != kNoKey32) atomicMin(&best[i], key);
}

// Any descriptor length: ref words re-read from global (L1), 64-bit keys.
__global__ void __launch_bounds__(kHamThreads) HammingForceGenericKernel(const uint32_t *__restrict__ ref, int n_ref, const uint32_t *__restrict__ cur,
                                                                        int n_cur, int words, int cur_per_split, unsigned long long *__restrict__ best) {
    const int i = blockIdx.x * kHamThreads + threadIdx.x;
    if (i >= n_ref) return;
    const int j_begin = blockIdx.y * cur_per_split;
    const int j_end = min(n_cur, j_begin + cur_per_split);
    unsigned long long key = kNoKey64;
    for (int j = j_begin; j < j_end; ++j) {
        unsigned d = 0;
        for (int w = 0; w < words; ++w) d += __popc(__ldg(ref + static_cast<size_t>(i) * words + w) ^ __ldg(cur + static_cast<size_t>(j) * words + w));
        const unsigned long long k = (static_cast<unsigned long long>(d) << 32) | static_cast<unsigned>(j);
        key = k < key ? k : key;
    }
    if (key != kNoKey64) atomicMin(&best[i], key);
}

// fill_unmatched: rows without a match get -1 (no index input); otherwise the caller's entry stays (descriptor_matcher.h:75-77).
__global__ void HammingFinalize32Kernel(const unsigned *best, int n_ref, float max_dist, int fill_unmatched, int *idx) {
    GridDepWait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ref) return;
    const unsigned key = best[i];
    if (key != kNoKey32 && static_cast<float>(key >> kJBits) < max_dist)
        idx[i] = static_cast<int>(key & ((1u << kJBits) - 1u));
    else if (fill_unmatched)
        idx[i] = -1;
}

__global__ void HammingFinalize64Kernel(const unsigned long long *best, int n_ref, float max_dist, int *idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ref) return;
    const unsigned long long key = best[i];
    if (key == kNoKey64) return;
    const float d = static_cast<float>(static_cast<unsigned>(key >> 32));
    if (d < max_dist) idx[i] = static_cast<int>(key & 0xFFFFFFFFull);
}

// ---------------------------------------------------------------------------------------------------------------
// Uniform grid over the cur feature positions (nearby matching)
// ---------------------------------------------------------------------------------------------------------------
struct GridDesc {
    float x0, y0, inv_cw, inv_ch;
    int gw, gh;
};

// Everything the grid build keeps on the device.  Between calls: bounds = kBoundsInit, ticket = 0, every count 0 -- restored by the
// kernels that consume them (BoundsGridKernel, ScanKernel), so a call queues neither a memset nor a host copy.
constexpr int kMaxCells = 256 * 256;
struct NearbyState {
    int bounds[4];  // min x, min y, max x, max y over finite positions, as order-preserving integers
    int ticket;     // blocks of BoundsGridKernel that have finished
    int n_special;  // non-finite positions ("always a candidate" list)
    GridDesc grid;
    int counts[kMaxCells];
    int starts[kMaxCells + 1];
    int cursor[kMaxCells];
};

__device__ __forceinline__ int FloatToOrdered(float f) {
    const int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float OrderedToFloat(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7FFFFFFF); }
__device__ __forceinline__ bool Finite2(float2 p) { return isfinite(p.x) && isfinite(p.y); }

__global__ void NearbyStateInitKernel(NearbyState *s) {
    s->bounds[0] = s->bounds[1] = 0x7FFFFFFF;
    s->bounds[2] = s->bounds[3] = static_cast<int>(0x80000000);
    s->ticket = 0;
    s->n_special = 0;
}

// Bounding box of the finite positions (one atomic set per warp), then -- in the last block to finish -- the grid geometry: cells about
// one search window wide, at most 256 x 256.  That block also puts bounds / ticket back to their initial values for the next call.
__global__ void __launch_bounds__(256) BoundsGridKernel(const float2 *pos, int n, int max_dcol, int max_drow, NearbyState *s) {
    GridDepLaunchDependents();
    GridDepWait();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    int lo_x = 0x7FFFFFFF, lo_y = 0x7FFFFFFF, hi_x = static_cast<int>(0x80000000), hi_y = static_cast<int>(0x80000000);
    if (j < n) {
        const float2 p = pos[j];
        if (Finite2(p)) lo_x = hi_x = FloatToOrdered(p.x), lo_y = hi_y = FloatToOrdered(p.y);
    }
    lo_x = __reduce_min_sync(0xFFFFFFFFu, lo_x), lo_y = __reduce_min_sync(0xFFFFFFFFu, lo_y);
    hi_x = __reduce_max_sync(0xFFFFFFFFu, hi_x), hi_y = __reduce_max_sync(0xFFFFFFFFu, hi_y);
    if ((threadIdx.x & 31) == 0 && lo_x != 0x7FFFFFFF) {
        atomicMin(&s->bounds[0], lo_x);
        atomicMin(&s->bounds[1], lo_y);
        atomicMax(&s->bounds[2], hi_x);
        atomicMax(&s->bounds[3], hi_y);
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    __threadfence();
    if (atomicAdd(&s->ticket, 1) != static_cast<int>(gridDim.x) - 1) return;
    __threadfence();
    volatile int *bounds = s->bounds;
    const int b0 = bounds[0], b1 = bounds[1], b2 = bounds[2], b3 = bounds[3];
    GridDesc g;
    g.gw = g.gh = 1;
    g.x0 = g.y0 = 0.0f;
    g.inv_cw = g.inv_ch = 0.0f;
    if (b0 != 0x7FFFFFFF) {
        const float x0 = OrderedToFloat(b0), y0 = OrderedToFloat(b1);
        const float x1 = OrderedToFloat(b2), y1 = OrderedToFloat(b3);
        float cw = static_cast<float>(max_dcol > 1 ? max_dcol : 1), ch = static_cast<float>(max_drow > 1 ? max_drow : 1);
        const float span_x = x1 - x0, span_y = y1 - y0;
        if (span_x / cw > 255.0f) cw = span_x / 255.0f;
        if (span_y / ch > 255.0f) ch = span_y / 255.0f;
        g.x0 = x0;
        g.y0 = y0;
        g.inv_cw = 1.0f / cw;
        g.inv_ch = 1.0f / ch;
        g.gw = min(static_cast<int>(span_x / cw) + 1, 256);
        g.gh = min(static_cast<int>(span_y / ch) + 1, 256);
    }
    s->grid = g;
    bounds[0] = bounds[1] = 0x7FFFFFFF;
    bounds[2] = bounds[3] = static_cast<int>(0x80000000);
    s->ticket = 0;
    s->n_special = 0;
}

__device__ __forceinline__ int CellCoord(float v, float v0, float inv, int n) {
    const float c = floorf((v - v0) * inv);
    return c < 0.0f ? 0 : (c > static_cast<float>(n - 1) ? n - 1 : static_cast<int>(c));
}

// counts[cell] for finite positions; non-finite positions go to the "always a candidate" list.
__global__ void CellCountKernel(const float2 *pos, int n, const GridDesc *gp, int *counts, int *special, int *n_special) {
    GridDepLaunchDependents();
    GridDepWait();
    const GridDesc g = *gp;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float2 p = pos[j];
    if (!Finite2(p)) {
        special[atomicAdd(n_special, 1)] = j;
        return;
    }
    atomicAdd(&counts[CellCoord(p.y, g.y0, g.inv_ch, g.gh) * g.gw + CellCoord(p.x, g.x0, g.inv_cw, g.gw)], 1);
}

// Exclusive scan of counts[0..n) into starts[0..n]; single block: per-thread range sums, then a shuffle scan over the 1024 partials.
// The counts are zeroed again on the way out (the state's "between calls" invariant).
__global__ void __launch_bounds__(1024) ScanKernel(int *counts, const GridDesc *gp, int *starts, int *cursor) {
    GridDepLaunchDependents();
    GridDepWait();
    const int n = gp->gw * gp->gh;
    __shared__ int warp_total[32];
    const int per = (n + 1023) / 1024;
    const int begin = threadIdx.x * per, end = min(n, begin + per);
    int sum = 0;
    for (int i = begin; i < end; ++i) sum += counts[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_total[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warp_total[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xFFFFFFFFu, w, o);
            if (lane >= o) w += up;
        }
        warp_total[lane] = w;  // inclusive totals of the warps
    }
    __syncthreads();
    int acc = incl - sum + (warp > 0 ? warp_total[warp - 1] : 0);
    if (threadIdx.x == 1023) starts[n] = acc + sum;
    for (int i = begin; i < end; ++i) {
        starts[i] = acc;
        cursor[i] = acc;
        acc += counts[i];
        counts[i] = 0;
    }
}

__global__ void CellFillKernel(const float2 *pos, int n, const GridDesc *gp, int *cursor, int *sorted) {
    GridDepLaunchDependents();
    GridDepWait();
    const GridDesc g = *gp;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float2 p = pos[j];
    if (!Finite2(p)) return;
    const int cell = CellCoord(p.y, g.y0, g.inv_ch, g.gh) * g.gw + CellCoord(p.x, g.x0, g.inv_cw, g.gw);
    sorted[atomicAdd(&cursor[cell], 1)] = j;
}

// ---------------------------------------------------------------------------------------------------------------
// Distances for the nearby kernels
// ---------------------------------------------------------------------------------------------------------------
// Order-preserving key for a non-NaN float distance (with -0 canonicalised to +0).
__device__ __forceinline__ unsigned FloatKey(float d) {
    d = d + 0.0f;
    const unsigned b = __float_as_uint(d);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float KeyFloat(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }

// A distance functor gives the scan a `Row` (what a warp keeps of its reference descriptor while it walks the candidates) and the distance
// of that row to current descriptor j.
struct HammingDist {
    const uint32_t *ref, *cur;
    int words;
    struct Row {
        int i;
    };
    __device__ __forceinline__ Row load(int i) const { return Row{i}; }
    // returns false when the distance is not comparable (never for Hamming)
    __device__ __forceinline__ bool operator()(const Row &r, int j, float *d) const {
        unsigned s = 0;
        for (int w = 0; w < words; ++w) s += __popc(__ldg(ref + static_cast<size_t>(r.i) * words + w) ^ __ldg(cur + static_cast<size_t>(j) * words + w));
        *d = words > 0 ? static_cast<float>(s) : 2147483647.0f;  // brief.cpp:34-36: empty descriptors -> kMaxInt32
        return true;
    }
};

// 256-bit descriptors in 16-byte aligned arrays (BRIEF-256, the usual case): the reference row lives in registers, a candidate costs two
// 128-bit loads and the carry-save popcount instead of 16 scalar loads and 8 POPC.
struct HammingDist256 {
    const uint4 *ref, *cur;
    struct Row {
        uint4 a, b;
    };
    __device__ __forceinline__ Row load(int i) const { return Row{__ldg(ref + 2 * static_cast<size_t>(i)), __ldg(ref + 2 * static_cast<size_t>(i) + 1)}; }
    __device__ __forceinline__ bool operator()(const Row &r, int j, float *d) const {
        const uint4 c = __ldg(cur + 2 * static_cast<size_t>(j)), e = __ldg(cur + 2 * static_cast<size_t>(j) + 1);
        *d = static_cast<float>(Popc8(r.a.x ^ c.x, r.a.y ^ c.y, r.a.z ^ c.z, r.a.w ^ c.w, r.b.x ^ e.x, r.b.y ^ e.y, r.b.z ^ e.z, r.b.w ^ e.w));
        return true;
    }
};

// Sequential dot product, k ascending, no FMA: bit-identical to the reference's scalar evaluation order.
__device__ __forceinline__ float SeqDot(const float *a, const float *b, int n) {
    float s = __fmul_rn(__ldg(a), __ldg(b));
    for (int k = 1; k < n; ++k) s = __fadd_rn(s, __fmul_rn(__ldg(a + k), __ldg(b + k)));
    return s;
}

struct CosineDist {
    const float *ref, *cur;
    const float *ref_norm, *cur_norm;
    int dim;
    struct Row {
        int i;
    };
    __device__ __forceinline__ Row load(int i) const { return Row{i}; }
    __device__ __forceinline__ bool operator()(const Row &r, int j, float *d) const {
        const int i = r.i;
        const float dot = SeqDot(ref + static_cast<size_t>(i) * dim, cur + static_cast<size_t>(j) * dim, dim);
        const float v = __fsub_rn(0.5f, __fmul_rn(__fdiv_rn(__fdiv_rn(dot, ref_norm[i]), cur_norm[j]), 0.5f));
        *d = v;
        return v == v;  // NaN never compares below anything in the reference
    }
};

__global__ void NormKernel(const float *desc, int n, int dim, float *norm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *a = desc + static_cast<size_t>(i) * dim;
    norm[i] = __fsqrt_rn(SeqDot(a, a, dim));
}

__device__ __forceinline__ unsigned long long WarpMin64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, v, o);
        v = other < v ? other : v;
    }
    return v;
}
__device__ __forceinline__ unsigned WarpMin32(unsigned v) { return __reduce_min_sync(0xFFFFFFFFu, v); }

// One warp per ref descriptor.  `limit_j`: candidates with j > limit_j are ignored (used for the d == 0 break).
template <typename Dist>
__device__ void NearbyScan(const Dist &dist, const typename Dist::Row &row, float2 pred, const float2 *pos, int n_cur, const GridDesc &g, const int *starts, const int *sorted,
                           const int *special, int n_special, float max_dcol, float max_drow, unsigned limit_j, unsigned long long *best_out,
                           unsigned *first_zero_out) {
    const int lane = threadIdx.x & 31;
    unsigned long long best = kNoKey64;
    unsigned first_zero = 0xFFFFFFFFu;

    auto consider = [&](int j) {
        if (static_cast<unsigned>(j) > limit_j) return;
        const float2 q = pos[j];
        // descriptor_matcher.h:108-111 (exact fp32 gate)
        if (fabsf(__fsub_rn(pred.x, q.x)) > max_dcol || fabsf(__fsub_rn(pred.y, q.y)) > max_drow) return;
        float d;
        if (!dist(row, j, &d)) return;
        const unsigned long long key = (static_cast<unsigned long long>(FloatKey(d)) << 32) | static_cast<unsigned>(j);
        best = key < best ? key : best;
        if (d == 0.0f) first_zero = min(first_zero, static_cast<unsigned>(j));
    };

    if (Finite2(pred)) {
        // Conservative cell range: one pixel (plus relative slack) wider than the window, cell mapping is monotonic.
        const float mx = max_dcol + 1.0f + fabsf(pred.x) * 1e-6f, my = max_drow + 1.0f + fabsf(pred.y) * 1e-6f;
        const int cx0 = CellCoord(pred.x - mx, g.x0, g.inv_cw, g.gw), cx1 = CellCoord(pred.x + mx, g.x0, g.inv_cw, g.gw);
        const int cy0 = CellCoord(pred.y - my, g.y0, g.inv_ch, g.gh), cy1 = CellCoord(pred.y + my, g.y0, g.inv_ch, g.gh);
        for (int cy = cy0; cy <= cy1; ++cy) {
            const int begin = starts[cy * g.gw + cx0], end = starts[cy * g.gw + cx1 + 1];
            for (int t = begin + lane; t < end; t += 32) consider(sorted[t]);
        }
        for (int t = lane; t < n_special; t += 32) consider(special[t]);
    } else {
        // A non-finite prediction makes every comparison of the gate false: all cur features are candidates.
        for (int j = lane; j < n_cur; j += 32) consider(j);
    }
    *best_out = WarpMin64(best);
    *first_zero_out = WarpMin32(first_zero);
}

template <typename Dist>
__global__ void __launch_bounds__(128) NearbyKernel(Dist dist, int n_ref, const float2 *pred, const float2 *pos, int n_cur, const GridDesc *gp, const int *starts,
                                                    const int *sorted, const int *special, const int *n_special_ptr, float max_dcol, float max_drow,
                                                    float max_dist, int fill_unmatched, int *idx) {
    GridDepWait();
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_ref) return;
    const int n_special = *n_special_ptr;
    const GridDesc g = *gp;
    unsigned long long best;
    unsigned first_zero;
    const typename Dist::Row row = dist.load(i);
    NearbyScan(dist, row, pred[i], pos, n_cur, g, starts, sorted, special, n_special, max_dcol, max_drow, 0xFFFFFFFFu, &best, &first_zero);
    if (first_zero != 0xFFFFFFFFu && static_cast<unsigned>(best & 0xFFFFFFFFull) > first_zero) {
        // A candidate after the reference's d == 0 break won: redo the scan over j <= first_zero only.
        unsigned dummy;
        NearbyScan(dist, row, pred[i], pos, n_cur, g, starts, sorted, special, n_special, max_dcol, max_drow, first_zero, &best, &dummy);
    }
    if ((threadIdx.x & 31) == 0) {
        if (best != kNoKey64 && KeyFloat(static_cast<unsigned>(best >> 32)) < max_dist)
            idx[i] = static_cast<int>(best & 0xFFFFFFFFull);
        else if (fill_unmatched)
            idx[i] = -1;  // otherwise the caller's entry stays (descriptor_matcher.h: only matched rows are assigned)
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Many small matching problems in one launch (one frame pair each: a few hundred descriptors per side).  One warp per ref
// descriptor; its lanes stride over the cur descriptors of the same pair, apply the reference's window gate when positions are
// given (descriptor_matcher.h:108-111) and reduce (distance, j) lexicographically.  The d == 0 break (:119) cannot change an
// integer-distance result: a later candidate never beats distance 0 at a lower index.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) HammingPairsKernel(const uint32_t *__restrict__ ref, const uint32_t *__restrict__ cur, int words, int n_ref_total,
                                                          const int *__restrict__ ref_pair, const int *__restrict__ ref_off, const int *__restrict__ cur_off,
                                                          const float2 *__restrict__ pred, const float2 *__restrict__ pos, float max_dcol, float max_drow,
                                                          float max_dist, int *__restrict__ idx) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n_ref_total) return;
    const int p = ref_pair[i];
    const int c0 = cur_off[p], c1 = cur_off[p + 1];
    (void)ref_off;
    const uint32_t *r = ref + static_cast<size_t>(i) * words;
    float2 pr = make_float2(0.0f, 0.0f);
    if (pred) pr = pred[i];
    unsigned long long best = kNoKey64;
    if (words == 8 && ((reinterpret_cast<uintptr_t>(ref) | reinterpret_cast<uintptr_t>(cur)) & 15) == 0) {
        // 256-bit descriptors, aligned: reference row in registers, two 128-bit loads and the carry-save popcount per candidate
        const uint4 ra = __ldg(reinterpret_cast<const uint4 *>(r)), rb = __ldg(reinterpret_cast<const uint4 *>(r) + 1);
        for (int j = c0 + lane; j < c1; j += 32) {
            if (pred) {
                const float2 q = pos[j];
                if (fabsf(__fsub_rn(pr.x, q.x)) > max_dcol || fabsf(__fsub_rn(pr.y, q.y)) > max_drow) continue;
            }
            const uint4 *c = reinterpret_cast<const uint4 *>(cur) + 2 * static_cast<size_t>(j);
            const uint4 ca = __ldg(c), cb = __ldg(c + 1);
            const unsigned d = Popc8(ra.x ^ ca.x, ra.y ^ ca.y, ra.z ^ ca.z, ra.w ^ ca.w, rb.x ^ cb.x, rb.y ^ cb.y, rb.z ^ cb.z, rb.w ^ cb.w);
            const unsigned long long key = (static_cast<unsigned long long>(d) << 32) | static_cast<unsigned>(j - c0);
            best = key < best ? key : best;
        }
    } else {
        for (int j = c0 + lane; j < c1; j += 32) {
            if (pred) {
                const float2 q = pos[j];
                if (fabsf(__fsub_rn(pr.x, q.x)) > max_dcol || fabsf(__fsub_rn(pr.y, q.y)) > max_drow) continue;
            }
            const uint32_t *c = cur + static_cast<size_t>(j) * words;
            unsigned d = 0;
            for (int w = 0; w < words; ++w) d += __popc(__ldg(r + w) ^ __ldg(c + w));
            const unsigned long long key = (static_cast<unsigned long long>(d) << 32) | static_cast<unsigned>(j - c0);
            best = key < best ? key : best;
        }
    }
    best = WarpMin64(best);
    if (lane == 0 && best != kNoKey64 && static_cast<float>(static_cast<unsigned>(best >> 32)) < max_dist) idx[i] = static_cast<int>(best & 0xFFFFFFFFull);
}

// The same for float descriptors (0.5 - 0.5 * cos): the warp's reference row is staged in shared memory, each lane evaluates one current
// descriptor of the pair at a time with the reference's own arithmetic (sequential fp32 dot, k ascending, no FMA, two divisions).
// Float distances can be negative, so the d == 0 break of NearbyMatch (descriptor_matcher.h:119) matters: when a candidate behind
// the first exact zero wins, the row is rescanned up to that zero only -- as NearbyKernel does.  ForceMatch has no break (:67-76).
constexpr int kCosPairsWarps = 8;
__global__ void __launch_bounds__(32 * kCosPairsWarps) CosinePairsKernel(const float *__restrict__ ref, const float *__restrict__ cur, int dim, int n_ref_total,
                                                                         const int *__restrict__ ref_pair, const int *__restrict__ cur_off,
                                                                         const float *__restrict__ ref_norm, const float *__restrict__ cur_norm,
                                                                         const float2 *__restrict__ pred, const float2 *__restrict__ pos, float max_dcol,
                                                                         float max_drow, float max_dist, int *__restrict__ idx) {
    extern __shared__ float s_rows[];  // [kCosPairsWarps][dim]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * kCosPairsWarps + warp;
    if (i >= n_ref_total) return;
    float *row = s_rows + warp * dim;
    for (int k = lane; k < dim; k += 32) row[k] = __ldg(ref + static_cast<size_t>(i) * dim + k);
    __syncwarp();
    const int p = ref_pair[i];
    const int c0 = cur_off[p], c1 = cur_off[p + 1];
    const float rn = ref_norm[i];
    float2 pr = make_float2(0.0f, 0.0f);
    if (pred) pr = pred[i];
    unsigned limit = 0xFFFFFFFFu;
    unsigned long long best = kNoKey64;
    for (int pass = 0; pass < 2; ++pass) {
        best = kNoKey64;
        unsigned first_zero = 0xFFFFFFFFu;
        for (int j = c0 + lane; j < c1; j += 32) {
            if (static_cast<unsigned>(j - c0) > limit) break;
            if (pred) {
                const float2 q = pos[j];
                if (fabsf(__fsub_rn(pr.x, q.x)) > max_dcol || fabsf(__fsub_rn(pr.y, q.y)) > max_drow) continue;
            }
            const float *c = cur + static_cast<size_t>(j) * dim;
            float dot = __fmul_rn(row[0], __ldg(c));
            for (int k = 1; k < dim; ++k) dot = __fadd_rn(dot, __fmul_rn(row[k], __ldg(c + k)));
            const float d = __fsub_rn(0.5f, __fmul_rn(__fdiv_rn(__fdiv_rn(dot, rn), cur_norm[j]), 0.5f));
            if (d != d) continue;  // NaN never compares below anything in the reference
            const unsigned long long key = (static_cast<unsigned long long>(FloatKey(d)) << 32) | static_cast<unsigned>(j - c0);
            best = key < best ? key : best;
            if (d == 0.0f) first_zero = min(first_zero, static_cast<unsigned>(j - c0));
        }
        best = WarpMin64(best);
        first_zero = WarpMin32(first_zero);
        // only NearbyMatch breaks at the first zero distance
        if (!pred || first_zero == 0xFFFFFFFFu || static_cast<unsigned>(best & 0xFFFFFFFFull) <= first_zero) break;
        limit = first_zero;
    }
    if (lane == 0 && best != kNoKey64 && KeyFloat(static_cast<unsigned>(best >> 32)) < max_dist) idx[i] = static_cast<int>(best & 0xFFFFFFFFull);
}

// ---------------------------------------------------------------------------------------------------------------
// Exact fp32 cosine force matching
// ---------------------------------------------------------------------------------------------------------------
constexpr int kCosThreads = 128;  // ref rows per CTA (one per thread)
constexpr int kCosCurTile = 32;   // cur descriptors per shared tile
constexpr int kCosK = 32;         // k-slice staged per step

// best[i] packs (FloatKey(distance) << 32 | j); first_zero[i] = lowest j with distance == 0.
__global__ void __launch_bounds__(kCosThreads) CosineForceKernel(const float *__restrict__ ref, int n_ref, const float *__restrict__ cur, int n_cur, int dim,
                                                                const float *__restrict__ ref_norm, const float *__restrict__ cur_norm,
                                                                int cur_per_split, unsigned long long *__restrict__ best) {
    // ref slice transposed [k][thread] (conflict-free), cur slice [k][j] (broadcast float4 reads)
    __shared__ float s_ref[kCosK][kCosThreads + 1];
    __shared__ __align__(16) float s_cur[kCosK][kCosCurTile];
    const int i0 = blockIdx.x * kCosThreads;
    const int i = i0 + threadIdx.x;
    const int j_begin = blockIdx.y * cur_per_split;
    const int j_end = min(n_cur, j_begin + cur_per_split);
    const float rn = i < n_ref ? ref_norm[i] : 1.0f;
    unsigned long long key = kNoKey64;

    for (int j0 = j_begin; j0 < j_end; j0 += kCosCurTile) {
        float acc[kCosCurTile];
#pragma unroll
        for (int t = 0; t < kCosCurTile; ++t) acc[t] = 0.0f;
        for (int k0 = 0; k0 < dim; k0 += kCosK) {
            __syncthreads();
            // stage ref[i0 .. i0+127][k0 .. k0+31]: consecutive threads read consecutive k of one row
            for (int t = threadIdx.x; t < kCosThreads * kCosK; t += kCosThreads) {
                const int r = t / kCosK, k = t % kCosK;
                s_ref[k][r] = (i0 + r < n_ref && k0 + k < dim) ? __ldg(ref + static_cast<size_t>(i0 + r) * dim + k0 + k) : 0.0f;
            }
            for (int t = threadIdx.x; t < kCosCurTile * kCosK; t += kCosThreads) {
                const int c = t / kCosK, k = t % kCosK;
                s_cur[k][c] = (j0 + c < j_end && k0 + k < dim) ? __ldg(cur + static_cast<size_t>(j0 + c) * dim + k0 + k) : 0.0f;
            }
            __syncthreads();
            const int kn = min(kCosK, dim - k0);
            for (int k = 0; k < kn; ++k) {
                const float a = s_ref[k][threadIdx.x];
#pragma unroll
                for (int t = 0; t < kCosCurTile; t += 4) {
                    const float4 b = *reinterpret_cast<const float4 *>(&s_cur[k][t]);
                    if (k0 + k == 0) {  // the reference's sum starts from the first product, not from 0 + product
                        acc[t] = __fmul_rn(a, b.x), acc[t + 1] = __fmul_rn(a, b.y), acc[t + 2] = __fmul_rn(a, b.z), acc[t + 3] = __fmul_rn(a, b.w);
                    } else {
                        acc[t] = __fadd_rn(acc[t], __fmul_rn(a, b.x));
                        acc[t + 1] = __fadd_rn(acc[t + 1], __fmul_rn(a, b.y));
                        acc[t + 2] = __fadd_rn(acc[t + 2], __fmul_rn(a, b.z));
                        acc[t + 3] = __fadd_rn(acc[t + 3], __fmul_rn(a, b.w));
                    }
                }
            }
        }
        const int n_tile = min(kCosCurTile, j_end - j0);
#pragma unroll
        for (int t = 0; t < kCosCurTile; ++t) {
            if (t < n_tile) {
                const float d = __fsub_rn(0.5f, __fmul_rn(__fdiv_rn(__fdiv_rn(acc[t], rn), cur_norm[j0 + t]), 0.5f));
                if (d == d) {
                    const unsigned long long k64 = (static_cast<unsigned long long>(FloatKey(d)) << 32) | static_cast<unsigned>(j0 + t);
                    key = k64 < key ? k64 : key;
                }
            }
        }
    }
    if (i < n_ref && key != kNoKey64) atomicMin(&best[i], key);
}

__global__ void CosineFinalizeKernel(const unsigned long long *best, int n_ref, float max_dist, int *idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ref) return;
    const unsigned long long key = best[i];
    if (key == kNoKey64) return;
    const float d = KeyFloat(static_cast<unsigned>(key >> 32));
    if (d < max_dist) idx[i] = static_cast<int>(key & 0xFFFFFFFFull);
}

int Blocks(int n, int threads) { return (n + threads - 1) / threads; }

// Number of cur splits so that the grid fills the GPU a few times over without making the tiles tiny.
int CurSplits(const ftk_context *ctx, int ref_blocks, int n_cur, int tile) {
    const int target = ctx->sm_count * 8;
    int splits = (target + ref_blocks - 1) / ref_blocks;
    const int max_splits = (n_cur + tile - 1) / tile;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    return splits;
}

template <typename Dist>
int RunNearby(ftk_context *ctx, const Dist &dist, int n_ref, int n_cur, const float2 *d_pred, const float2 *d_pos, int max_drow, int max_dcol,
              float max_dist, int *d_idx, bool fill_unmatched) {
    cudaStream_t st = ctx->stream;
    if (int rc = EnsureDevice(ctx, ctx->d_nearby_state, sizeof(NearbyState))) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_work2, sizeof(int) * (2 * static_cast<size_t>(n_cur) + 8))) return rc;
    NearbyState *state = static_cast<NearbyState *>(ctx->d_nearby_state.ptr);
    int *d_sorted = static_cast<int *>(ctx->d_work2.ptr);
    int *d_special = d_sorted + n_cur;
    if (!ctx->nearby_state_clean) {  // first use, or a call that failed half-way
        FTK_CUDA_CHECK(ctx, cudaMemsetAsync(state, 0, sizeof(NearbyState), st));
        NearbyStateInitKernel<<<1, 1, 0, st>>>(state);
    }
    ctx->nearby_state_clean = false;  // until this call's launches are all queued
    // 1. bounding box of the finite cur positions -> grid geometry, all on the device (no host round trip)
    // (an ordinary launch: the first kernel of a call never starts early -- see LaunchHammingForce; it costs < 0.5 us)
    BoundsGridKernel<<<Blocks(n_cur, 256), 256, 0, st>>>(d_pos, n_cur, max_dcol, max_drow, state);
    // 2. count -> scan -> fill
    FTK_CUDA_CHECK(ctx, LaunchDependent(CellCountKernel, Blocks(n_cur, 256), 256, 0, st, d_pos, n_cur, &state->grid, state->counts, d_special, &state->n_special));
    FTK_CUDA_CHECK(ctx, LaunchDependent(ScanKernel, 1, 1024, 0, st, state->counts, &state->grid, state->starts, state->cursor));
    FTK_CUDA_CHECK(ctx, LaunchDependent(CellFillKernel, Blocks(n_cur, 256), 256, 0, st, d_pos, n_cur, &state->grid, state->cursor, d_sorted));
    // 3. one warp per ref descriptor
    FTK_CUDA_CHECK(ctx, LaunchDependent(NearbyKernel<Dist>, Blocks(n_ref * 32, 128), 128, 0, st, dist, n_ref, d_pred, d_pos, n_cur, &state->grid, state->starts, d_sorted,
                                        d_special, &state->n_special, static_cast<float>(max_dcol), static_cast<float>(max_drow), max_dist, fill_unmatched ? 1 : 0,
                                        d_idx));
    ctx->launches += 5;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    ctx->nearby_state_clean = true;
    return FTK_OK;
}

int ComputeNorms(ftk_context *ctx, const float *d_ref, int n_ref, const float *d_cur, int n_cur, int dim, float **ref_norm, float **cur_norm) {
    if (int rc = EnsureDevice(ctx, ctx->d_work3, sizeof(float) * (static_cast<size_t>(n_ref) + n_cur + 8))) return rc;
    *ref_norm = static_cast<float *>(ctx->d_work3.ptr);
    *cur_norm = *ref_norm + n_ref;
    NormKernel<<<Blocks(n_ref, 128), 128, 0, ctx->stream>>>(d_ref, n_ref, dim, *ref_norm);
    NormKernel<<<Blocks(n_cur, 128), 128, 0, ctx->stream>>>(d_cur, n_cur, dim, *cur_norm);
    ctx->launches += 2;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

}  // namespace

int LaunchHammingForce(ftk_context *ctx, const uint32_t *d_ref, int n_ref, const uint32_t *d_cur, int n_cur, int words, float max_dist, int *d_idx,
                       bool fill_unmatched) {
    if (n_ref == 0) return FTK_OK;
    cudaStream_t st = ctx->stream;
    if (words == 0) {  // empty descriptors: distance kMaxInt32 never beats a sane max_dist
        if (fill_unmatched) FTK_CUDA_CHECK(ctx, cudaMemsetAsync(d_idx, 0xFF, sizeof(int) * n_ref, st));
        return FTK_OK;
    }
    const int ref_blocks = Blocks(n_ref, kHamThreads);
    const bool packed = (words == 4 || words == 8 || words == 16) && n_cur <= (1 << kJBits);
    if (packed) {
        if (int rc = EnsureDevice(ctx, ctx->d_work0, sizeof(unsigned) * static_cast<size_t>(n_ref))) return rc;
        unsigned *d_best = static_cast<unsigned *>(ctx->d_work0.ptr);
        // The first kernel of the call is an ordinary launch: were it allowed to start early too, the chain of early launches would let the
        // force kernel's CTAs of call n + 1 take whatever slots call n frees first, and its 800 CTAs would end up unevenly spread over the
        // SMs (measured: 190 us instead of 148 us per call in a back-to-back loop).
        FillU32Kernel<<<Blocks(n_ref, 256), 256, 0, st>>>(d_best, n_ref, kNoKey32);
        const int splits = CurSplits(ctx, ref_blocks, n_cur, kHamTile);
        const int per = ((n_cur + splits - 1) / splits + kHamTile - 1) / kHamTile * kHamTile;
        const dim3 grid(ref_blocks, (n_cur + per - 1) / per);
        ProfBegin(ctx);
        auto kernel = words == 8 ? HammingForceKernel<8> : words == 4 ? HammingForceKernel<4> : HammingForceKernel<16>;
        FTK_CUDA_CHECK(ctx, LaunchDependent(kernel, grid, kHamThreads, 0, st, d_ref, n_ref, d_cur, n_cur, per, d_best));
        ProfEnd(ctx);
        FTK_CUDA_CHECK(ctx, LaunchDependent(HammingFinalize32Kernel, Blocks(n_ref, 256), 256, 0, st, d_best, n_ref, max_dist, fill_unmatched ? 1 : 0, d_idx));
    } else {
        if (int rc = EnsureDevice(ctx, ctx->d_work0, sizeof(unsigned long long) * static_cast<size_t>(n_ref))) return rc;
        unsigned long long *d_best = static_cast<unsigned long long *>(ctx->d_work0.ptr);
        if (fill_unmatched) FTK_CUDA_CHECK(ctx, cudaMemsetAsync(d_idx, 0xFF, sizeof(int) * n_ref, st));  // -1
        FillU64Kernel<<<Blocks(n_ref, 256), 256, 0, st>>>(d_best, n_ref, kNoKey64);
        const int splits = CurSplits(ctx, ref_blocks, n_cur, 64);
        const int per = (n_cur + splits - 1) / splits;
        const dim3 grid(ref_blocks, (n_cur + per - 1) / per);
        HammingForceGenericKernel<<<grid, kHamThreads, 0, st>>>(d_ref, n_ref, d_cur, n_cur, words, per, d_best);
        HammingFinalize64Kernel<<<Blocks(n_ref, 256), 256, 0, st>>>(d_best, n_ref, max_dist, d_idx);
    }
    ctx->launches += 3;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

// d_pred / d_pos null: ForceMatch per pair; otherwise NearbyMatch per pair.  d_ref_pair [n_ref_total] = pair of every ref descriptor.
int LaunchHammingPairs(ftk_context *ctx, const uint32_t *d_ref, const uint32_t *d_cur, int words, int n_ref_total, const int *d_ref_pair,
                       const int *d_ref_off, const int *d_cur_off, const float2 *d_pred, const float2 *d_pos, int max_drow, int max_dcol, float max_dist,
                       int *d_idx) {
    if (n_ref_total == 0 || words == 0) return FTK_OK;
    HammingPairsKernel<<<Blocks(n_ref_total, 8), 256, 0, ctx->stream>>>(d_ref, d_cur, words, n_ref_total, d_ref_pair, d_ref_off, d_cur_off, d_pred, d_pos,
                                                                       static_cast<float>(max_dcol), static_cast<float>(max_drow), max_dist, d_idx);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

int LaunchCosinePairs(ftk_context *ctx, const float *d_ref, const float *d_cur, int dim, int n_ref_total, int n_cur_total, const int *d_ref_pair,
                      const int *d_cur_off, const float2 *d_pred, const float2 *d_pos, int max_drow, int max_dcol, float max_dist, int *d_idx) {
    if (n_ref_total == 0) return FTK_OK;
    float *ref_norm, *cur_norm;
    if (int rc = ComputeNorms(ctx, d_ref, n_ref_total, d_cur, n_cur_total, dim, &ref_norm, &cur_norm)) return rc;
    const size_t smem = sizeof(float) * static_cast<size_t>(kCosPairsWarps) * dim;
    CosinePairsKernel<<<Blocks(n_ref_total, kCosPairsWarps), 32 * kCosPairsWarps, smem, ctx->stream>>>(d_ref, d_cur, dim, n_ref_total, d_ref_pair, d_cur_off, ref_norm,
                                                                                                       cur_norm, d_pred, d_pos, static_cast<float>(max_dcol),
                                                                                                       static_cast<float>(max_drow), max_dist, d_idx);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

int LaunchHammingNearby(ftk_context *ctx, const uint32_t *d_ref, int n_ref, const uint32_t *d_cur, int n_cur, int words, const float2 *d_pred,
                        const float2 *d_pos, int max_drow, int max_dcol, float max_dist, int *d_idx, bool fill_unmatched) {
    if (n_ref == 0) return FTK_OK;
    if (words == 8 && reinterpret_cast<uintptr_t>(d_ref) % 16 == 0 && reinterpret_cast<uintptr_t>(d_cur) % 16 == 0) {
        const HammingDist256 dist{reinterpret_cast<const uint4 *>(d_ref), reinterpret_cast<const uint4 *>(d_cur)};
        return RunNearby(ctx, dist, n_ref, n_cur, d_pred, d_pos, max_drow, max_dcol, max_dist, d_idx, fill_unmatched);
    }
    const HammingDist dist{d_ref, d_cur, words};
    return RunNearby(ctx, dist, n_ref, n_cur, d_pred, d_pos, max_drow, max_dcol, max_dist, d_idx, fill_unmatched);
}

int LaunchCosineForce(ftk_context *ctx, const float *d_ref, int n_ref, const float *d_cur, int n_cur, int dim, float max_dist, int *d_idx, bool fill_unmatched) {
    if (n_ref == 0) return FTK_OK;
    ctx->d_last_scan_items = nullptr;
    if (ctx->use_fast_paths) {
        const int rc = LaunchCosineForceTensor(ctx, d_ref, n_ref, d_cur, n_cur, dim, max_dist, d_idx, fill_unmatched);
        if (rc != FTK_ERR_UNSUPPORTED) return rc;
    }
    cudaStream_t st = ctx->stream;
    if (fill_unmatched) FTK_CUDA_CHECK(ctx, cudaMemsetAsync(d_idx, 0xFF, sizeof(int) * n_ref, st));  // -1
    float *ref_norm, *cur_norm;
    if (int rc = ComputeNorms(ctx, d_ref, n_ref, d_cur, n_cur, dim, &ref_norm, &cur_norm)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_work0, sizeof(unsigned long long) * static_cast<size_t>(n_ref))) return rc;
    unsigned long long *d_best = static_cast<unsigned long long *>(ctx->d_work0.ptr);
    FillU64Kernel<<<Blocks(n_ref, 256), 256, 0, st>>>(d_best, n_ref, kNoKey64);
    const int ref_blocks = Blocks(n_ref, kCosThreads);
    const int splits = CurSplits(ctx, ref_blocks, n_cur, kCosCurTile);
    const int per = ((n_cur + splits - 1) / splits + kCosCurTile - 1) / kCosCurTile * kCosCurTile;
    const dim3 grid(ref_blocks, (n_cur + per - 1) / per);
    CosineForceKernel<<<grid, kCosThreads, 0, st>>>(d_ref, n_ref, d_cur, n_cur, dim, ref_norm, cur_norm, per, d_best);
    CosineFinalizeKernel<<<Blocks(n_ref, 256), 256, 0, st>>>(d_best, n_ref, max_dist, d_idx);
    ctx->launches += 3;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

int LaunchCosineNearby(ftk_context *ctx, const float *d_ref, int n_ref, const float *d_cur, int n_cur, int dim, const float2 *d_pred,
                       const float2 *d_pos, int max_drow, int max_dcol, float max_dist, int *d_idx, bool fill_unmatched) {
    if (n_ref == 0) return FTK_OK;
    float *ref_norm, *cur_norm;
    if (int rc = ComputeNorms(ctx, d_ref, n_ref, d_cur, n_cur, dim, &ref_norm, &cur_norm)) return rc;
    const CosineDist dist{d_ref, d_cur, ref_norm, cur_norm, dim};
    return RunNearby(ctx, dist, n_ref, n_cur, d_pred, d_pos, max_drow, max_dcol, max_dist, d_idx, fill_unmatched);
}

}  // namespace ftk
