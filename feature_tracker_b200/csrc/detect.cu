// K9 / K10: corner detection and BRIEF description on the device pyramid.
//
// Stands in for feature_detector::FeaturePointHarrisDetector::DetectGoodFeatures and feature_detector::BriefDescriptor::Compute
// (called at test/test_descriptor_matcher_brief.cpp:59-76 and test/test_optical_flow.cpp:60-66).  Those classes belong to the
// sibling repository Feature_Detector, which is not part of the reference tree: PARITY IS UNPINNED.  The arithmetic is the
// published algorithm as frozen in the detector section of oracle/ftk_oracle.c, and
// these kernels are bit-exact against that restatement.
//
// K9a ResponseKernel   one pass over the image (HBM bound: 1 B read + 8 B written per pixel): integer structure-tensor sums from a
//                      shared-memory tile, fp32 response with explicit rounding, 32-bit ordered state per candidate pixel.
// K9b SelectRoundTileKernel   the sequential "visit by falling response, take unless a taken feature is near" loop as a parallel
//                      fixed point over the state plane: per round, the maximum of every pixel's window (separable: rows, then
//                      columns, in shared memory); a candidate whose window holds a taken pixel is dropped, one that holds its
//                      window's maximum (ties: the lower row-major index) is taken, the others wait.  Every decision is final and
//                      equals the sequential loop's; rounds repeat until no candidate is undecided.  Windows too wide for
//                      shared memory use RowMaxKernel + DecideKernel (two kernels per round, 64-bit keys through HBM).
// K9c CollectKernel + SortEmitKernel   the taken set ordered by key (bitonic sort in shared memory; CUB segmented radix sort +
//                      EmitKernel for lists beyond 4096 keys); the first `needed` are the sequential loop's output.
// K10 BriefKernel      one warp per feature, one pair per lane, a ballot per 32-bit descriptor word.
#include <cub/device/device_segmented_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "ftk_internal.h"

namespace ftk {

namespace {

constexpr int kDetTile = 32;     // output tile edge
constexpr int kDetMaxMargin = 4; // half_patch (<= 3) + 1
constexpr int kRoundsPerBatch = 8;
constexpr int kMaxRounds = 1 << 16;

__device__ __forceinline__ unsigned DetOrderMap(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Pixel states of the selection, one 32-bit word per pixel in HBM: 0 = not a candidate / dropped, kTaken32 = taken, anything else =
// the ordered response bits of an undecided candidate.  The two-kernel rounds compare 64-bit keys (state << 32 | ~pixel index),
// built when a state is loaded, so that equal responses resolve towards the lower index; 0 and kTaken keep their meaning as keys.
using Key = unsigned long long;
using State = unsigned;
constexpr Key kTaken = ~0ull;
constexpr State kTaken32 = ~0u;
__device__ __forceinline__ Key StateKey(State v, unsigned pixel) {
    return v == 0u ? 0ull : (v == kTaken32 ? kTaken : ((static_cast<Key>(v) << 32) | (0xFFFFFFFFu - pixel)));
}
// -0 + 0 = +0: one state for both zeros, like the float comparison of the sequential loop; finite responses never map to 0 or kTaken32.
__device__ __forceinline__ State MakeState(float response) { return DetOrderMap(__fadd_rn(response, 0.0f)); }
constexpr int kRowMaxThreads = 256;

// blockIdx.z = image of the batch in every selection kernel: image z owns pixels [z * rows * cols, (z + 1) * rows * cols) of the
// response / state planes and the source plane img + z * image_stride.
__global__ void __launch_bounds__(256) ResponseKernel(const uint8_t *__restrict__ img, long long image_stride, int rows, int cols, int pitch,
                                                      ftk_detector_params p, float *__restrict__ response, State *__restrict__ state) {
    __shared__ uint8_t tile[kDetTile + 2 * kDetMaxMargin][kDetTile + 2 * kDetMaxMargin + 8];
    img += blockIdx.z * image_stride;
    const size_t plane = static_cast<size_t>(blockIdx.z) * rows * cols;
    response += plane;
    if (state) state += plane;
    const int h = p.half_patch, m = h + 1, edge = kDetTile + 2 * m;
    const int r0 = blockIdx.y * kDetTile, c0 = blockIdx.x * kDetTile;
    for (int tr = threadIdx.y; tr < edge; tr += 8) {
        const int r = r0 - m + tr;
        for (int tc = threadIdx.x; tc < edge; tc += 32) {
            const int c = c0 - m + tc;
            tile[tr][tc] = (r >= 0 && r < rows && c >= 0 && c < cols) ? img[static_cast<size_t>(r) * pitch + c] : 0;
        }
    }
    __syncthreads();
    const float inv = __fdiv_rn(1.0f, static_cast<float>(4 * (2 * h + 1) * (2 * h + 1)));
#pragma unroll
    for (int q = 0; q < kDetTile / 8; ++q) {
        const int lr = threadIdx.y + 8 * q, lc = threadIdx.x;
        const int r = r0 + lr, c = c0 + lc;
        if (r >= rows || c >= cols) continue;
        float resp = -INFINITY;
        if (r >= m && r < rows - m && c >= m && c < cols - m) {
            int sxx = 0, syy = 0, sxy = 0;
            for (int dr = -h; dr <= h; ++dr) {
                for (int dc = -h; dc <= h; ++dc) {
                    const int tr = lr + m + dr, tc = lc + m + dc;
                    const int gx = static_cast<int>(tile[tr][tc + 1]) - static_cast<int>(tile[tr][tc - 1]);
                    const int gy = static_cast<int>(tile[tr + 1][tc]) - static_cast<int>(tile[tr - 1][tc]);
                    sxx += gx * gx, syy += gy * gy, sxy += gx * gy;
                }
            }
            const float a = __fmul_rn(static_cast<float>(sxx), inv), b = __fmul_rn(static_cast<float>(sxy), inv), cc = __fmul_rn(static_cast<float>(syy), inv);
            if (p.kind == FTK_DETECTOR_HARRIS) {
                const float det = __fsub_rn(__fmul_rn(a, cc), __fmul_rn(b, b)), tr = __fadd_rn(a, cc);
                resp = __fsub_rn(det, __fmul_rn(p.harris_k, __fmul_rn(tr, tr)));
            } else {
                const float d = __fsub_rn(a, cc);
                const float disc = __fadd_rn(__fmul_rn(d, d), __fmul_rn(4.0f, __fmul_rn(b, b)));
                resp = __fmul_rn(0.5f, __fsub_rn(__fadd_rn(a, cc), __fsqrt_rn(disc)));
            }
        }
        const unsigned i = static_cast<unsigned>(r) * cols + c;
        response[i] = resp;
        if (state) state[i] = resp >= p.min_response ? MakeState(resp) : 0u;
    }
}

// Features the caller already holds block their window (they are not candidates themselves).
__global__ void ExistingMaskKernel(const float2 *__restrict__ existing, int rows, int cols, int dist, State *__restrict__ state) {
    const float2 f = existing[blockIdx.x];
    if (!(f.x >= 0.0f && f.y >= 0.0f && f.x < static_cast<float>(cols) && f.y < static_cast<float>(rows))) return;
    const int r = static_cast<int>(f.y), c = static_cast<int>(f.x);
    const int ra = max(0, r - (dist - 1)), rb = min(rows - 1, r + (dist - 1)), ca = max(0, c - (dist - 1)), cb = min(cols - 1, c + (dist - 1));
    const unsigned w = cb - ca + 1, count = w * (rb - ra + 1);  // <= rows * cols < 2^32
    for (unsigned k = threadIdx.x; k < count; k += blockDim.x) state[static_cast<size_t>(ra + k / w) * cols + ca + k % w] = 0u;
}

// rowmax[r][c] = max of state[r][c - (dist-1) .. c + (dist-1)].  One block per 256-column segment of a row, staged through shared memory.
__global__ void __launch_bounds__(kRowMaxThreads) RowMaxKernel(const State *__restrict__ state, int rows, int cols, int dist, Key *__restrict__ rowmax,
                                                               const unsigned *__restrict__ counters, int round) {
    if (round > 0 && counters[round - 1] == 0) return;  // converged earlier in this batch
    extern __shared__ Key seg[];
    state += static_cast<size_t>(blockIdx.z) * rows * cols;
    rowmax += static_cast<size_t>(blockIdx.z) * rows * cols;
    const int c0 = blockIdx.x * kRowMaxThreads, halo = dist - 1, width = kRowMaxThreads + 2 * halo;
    const State *row = state + static_cast<size_t>(blockIdx.y) * cols;
    for (int k = threadIdx.x; k < width; k += kRowMaxThreads) {
        const int c = c0 - halo + k;
        seg[k] = (c >= 0 && c < cols) ? StateKey(row[c], blockIdx.y * cols + c) : 0ull;
    }
    __syncthreads();
    const int c = c0 + threadIdx.x;
    if (c >= cols) return;
    Key best = 0ull;
    for (int k = 0; k <= 2 * halo; ++k) best = max(best, seg[threadIdx.x + k]);
    rowmax[static_cast<size_t>(blockIdx.y) * cols + c] = best;
}

// Window maximum = column maximum of the row maxima; then the decision for every undecided candidate.  counters[round] = candidates
// still undecided after this round.  Reads only the snapshot in `rowmax` and the pixel's own state: no ordering between threads.
__global__ void __launch_bounds__(256) DecideKernel(const Key *__restrict__ rowmax, State *__restrict__ state, int rows, int cols, int dist,
                                                    unsigned *__restrict__ counters, int round) {
    if (round > 0 && counters[round - 1] == 0) return;
    rowmax += static_cast<size_t>(blockIdx.z) * rows * cols;
    state += static_cast<size_t>(blockIdx.z) * rows * cols;
    const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
    bool waits = false;
    if (c < cols && r < rows) {
        const size_t i = static_cast<size_t>(r) * cols + c;
        const Key s = StateKey(state[i], static_cast<unsigned>(i));
        if (s != 0ull && s != kTaken) {
            const int ra = max(0, r - (dist - 1)), rb = min(rows - 1, r + (dist - 1));
            Key best = 0ull;
            for (int rr = ra; rr <= rb; ++rr) best = max(best, __ldg(rowmax + static_cast<size_t>(rr) * cols + c));
            if (best == kTaken) state[i] = 0u;
            else if (best == s) state[i] = kTaken32;
            else waits = true;
        }
    }
    const int n_waiting = __syncthreads_count(waits);
    if (n_waiting && threadIdx.x == 0 && threadIdx.y == 0) atomicAdd(counters + round, static_cast<unsigned>(n_waiting));
}

// out[k] = max of element(k) .. element(k + W - 1), k = 0 .. G-1: the G windows share elements G-1 .. W-1, window k adds the suffix
// k .. G-2 on the left and the prefix W .. W+k-1 on the right -- W + G - 1 reads instead of G * W.
template <int G, typename Element>
__device__ __forceinline__ void WindowMax(Element element, int W, State (&out)[G]) {
    if (W < G) {
#pragma unroll
        for (int k = 0; k < G; ++k) {
            State best = 0u;
            for (int j = 0; j < W; ++j) best = max(best, element(k + j));
            out[k] = best;
        }
        return;
    }
    State common = 0u;
    for (int j = G - 1; j < W; ++j) common = max(common, element(j));
    out[G - 1] = common;
#pragma unroll
    for (int k = G - 2; k >= 0; --k) out[k] = max(out[k + 1], element(k));  // common + left suffix
    State prefix = 0u;
#pragma unroll
    for (int k = 1; k < G; ++k) {
        prefix = max(prefix, element(W + k - 1));
        out[k] = max(out[k], prefix);
    }
}

// One round in one kernel for windows that fit in shared memory: a CTA stages the 32-bit states of its 32 x 32 tile plus the halo,
// takes row maxima, then column maxima (one integer max per element: 0 < any response < kTaken32), and decides its own pixels IN
// PLACE: a taken state in the window drops the pixel, a larger state makes it wait, and a pixel that holds the window's maximum is
// taken unless an equal state with a lower index is in the window (checked explicitly, in index order, for those few pixels only).
// Neighbouring CTAs may read a pixel before or after its decision; both readings lead to decisions the sequential loop also makes
// (a state only ever moves from undecided to its final value, "taken" needs every higher key of the window finally dropped,
// "dropped" needs a finally taken key in the window), so the fixed point is the same -- only the number of rounds can differ.
constexpr int kGroup = 8;
constexpr int kStageCols = 4;  // column slots per lane while staging: edge <= 32 * kStageCols
__global__ void __launch_bounds__(256, 6) SelectRoundTileKernel(State *state, int rows, int cols, int dist, unsigned *__restrict__ counters, int round) {
    if (round > 0 && counters[round - 1] == 0) return;  // converged earlier in this batch
    extern __shared__ State sm32[];
    state += static_cast<size_t>(blockIdx.z) * rows * cols;
    const int h = dist - 1, edge = kDetTile + 2 * h, stride = edge | 1, W = 2 * h + 1;
    State *B = sm32;                    // B[row][32]: row maxima for the tile's columns
    State *A = sm32 + edge * kDetTile;  // A: states of tile + halo, `stride` words per row
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int r_tile = blockIdx.y * kDetTile, c_tile = blockIdx.x * kDetTile;
    bool undecided = false;
#pragma unroll
    for (int q = 0; q < kDetTile / 8; ++q) {
        const int r = r_tile + threadIdx.y + 8 * q, c = c_tile + threadIdx.x;
        if (r < rows && c < cols) {
            const State v = __ldcg(state + static_cast<size_t>(r) * cols + c);
            undecided |= v != 0u && v != kTaken32;
        }
    }
    if (!__syncthreads_or(undecided)) return;
    {  // staging: warp = row of the region, lane = column slot; everything column-dependent is computed once
        int offset[kStageCols];
        bool inside[kStageCols];
#pragma unroll
        for (int q = 0; q < kStageCols; ++q) {
            const int tc = threadIdx.x + 32 * q, c = c_tile - h + tc;
            inside[q] = tc < edge && c >= 0 && c < cols;
            offset[q] = inside[q] ? c : 0;
        }
        for (int tr = threadIdx.y; tr < edge; tr += 8) {
            const int r = r_tile - h + tr;
            const bool row_inside = r >= 0 && r < rows;
            const State *row = state + static_cast<size_t>(row_inside ? r : 0) * cols;
            State v[kStageCols];
#pragma unroll
            for (int q = 0; q < kStageCols; ++q) v[q] = (row_inside && inside[q]) ? __ldcg(row + offset[q]) : 0u;
#pragma unroll
            for (int q = 0; q < kStageCols; ++q)
                if (threadIdx.x + 32 * q < edge) A[tr * stride + threadIdx.x + 32 * q] = v[q];
        }
    }
    __syncthreads();
    for (int item = tid; item < edge * (kDetTile / kGroup); item += 256) {
        const int row = item / (kDetTile / kGroup), g = item % (kDetTile / kGroup);
        const State *in = A + row * stride + g * kGroup;
        State out[kGroup];
        WindowMax<kGroup>([&](int j) { return in[j]; }, W, out);
#pragma unroll
        for (int k = 0; k < kGroup; ++k) B[row * kDetTile + g * kGroup + k] = out[k];
    }
    __syncthreads();
    if (tid < 32 * (kDetTile / kGroup)) {
        const int col = tid & 31, g = tid >> 5;
        const State *in = B + g * kGroup * kDetTile + col;
        State out[kGroup];
        WindowMax<kGroup>([&](int j) { return in[j * kDetTile]; }, W, out);
        unsigned waiting = 0;
#pragma unroll
        for (int k = 0; k < kGroup; ++k) {
            const int lr = g * kGroup + k;  // the same tile row for the whole warp
            const State *centre = A + (lr + h) * stride + h + col;
            const State s = *centre;
            const bool live = s != 0u && s != kTaken32;  // false for every pixel outside the image
            State *mine = state + static_cast<size_t>(r_tile + lr) * cols + c_tile + col;
            if (live && out[k] == kTaken32) __stcg(mine, 0u);
            if (live && out[k] != kTaken32 && out[k] > s) ++waiting;
            // The window's maximum response is mine; an equal one at a lower index (rows above, or left in my row) goes first.  Few
            // pixels get here, so the warp checks them one after the other, all lanes scanning one pixel's window part.
            bool holds = live && out[k] == s;
            if (holds && h > 0 && (centre[-1] == s || centre[-stride] == s)) holds = false, ++waiting;  // plateaus: the neighbour goes first
            unsigned holders = __ballot_sync(0xFFFFFFFFu, holds);
            while (holders) {
                const int src = __ffs(holders) - 1;
                holders &= holders - 1;
                const State want = __shfl_sync(0xFFFFFFFFu, s, src);
                const State *c0 = A + (lr + h) * stride + h + src;
                bool lower_tie = static_cast<int>(threadIdx.x) < h && c0[static_cast<int>(threadIdx.x) - h] == want;  // h <= 48: two rounds at most
                if (static_cast<int>(threadIdx.x) + 32 < h) lower_tie |= c0[static_cast<int>(threadIdx.x) + 32 - h] == want;
                for (int dr = -h; dr < 0; ++dr)
                    for (int dc = -h + static_cast<int>(threadIdx.x); dc <= h; dc += 32) lower_tie |= c0[dr * stride + dc] == want;
                lower_tie = __any_sync(0xFFFFFFFFu, lower_tie);
                if (col == src) {
                    if (lower_tie) ++waiting;
                    else __stcg(mine, kTaken32);
                }
            }
        }
        waiting = __reduce_add_sync(0xFFFFFFFFu, waiting);
        if (col == 0 && waiting) atomicAdd(counters + round, waiting);
    }
}

// taken[z] counts image z's taken pixels; its keys go to keys + z * max_taken.
__global__ void CollectKernel(const State *__restrict__ state, unsigned n_pixels, const float *__restrict__ response, Key *__restrict__ keys,
                              unsigned max_taken, unsigned *__restrict__ taken) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t plane = static_cast<size_t>(blockIdx.z) * n_pixels;
    if (i >= n_pixels || state[plane + i] != kTaken32) return;
    const unsigned slot = atomicAdd(taken + blockIdx.z, 1u);
    if (slot < max_taken) keys[static_cast<size_t>(blockIdx.z) * max_taken + slot] = StateKey(MakeState(response[plane + i]), i);
}

// Per-image key lists of at most kSortMax keys: one CTA sorts image z's list in shared memory (bitonic, descending) and writes the
// first `needed` features straight away -- no sorted copy in HBM, no separate emit pass.
constexpr int kSortMax = 4096;
__global__ void __launch_bounds__(256) SortEmitKernel(const Key *__restrict__ keys, const unsigned *__restrict__ taken, unsigned max_taken, int needed, int rows,
                                                      int cols, const float *__restrict__ response, float2 *__restrict__ uv, float *__restrict__ out_response,
                                                      int *__restrict__ n_out) {
    extern __shared__ Key list[];
    const int z = blockIdx.x;
    const unsigned count = taken[z];
    if (count > max_taken) {  // more taken features than can exist (never happens; the host reports it)
        if (threadIdx.x == 0) n_out[z] = -1;
        return;
    }
    int padded = 1;
    while (padded < static_cast<int>(count)) padded <<= 1;
    for (int i = threadIdx.x; i < padded; i += blockDim.x) list[i] = i < static_cast<int>(count) ? keys[static_cast<size_t>(z) * max_taken + i] : 0ull;
    __syncthreads();
    for (int size = 2; size <= padded; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < padded / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;  // lo has bit `stride` clear
                const bool descending = (lo & size) == 0;
                const Key a = list[lo], b = list[hi];
                if ((a < b) == descending) list[lo] = b, list[hi] = a;
            }
            __syncthreads();
        }
    }
    const int n = min(static_cast<int>(count), needed);
    if (threadIdx.x == 0) n_out[z] = n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned p = 0xFFFFFFFFu - static_cast<unsigned>(list[i] & 0xFFFFFFFFull);
        uv[static_cast<size_t>(z) * needed + i] = make_float2(static_cast<float>(p % cols), static_cast<float>(p / cols));
        if (out_response) out_response[static_cast<size_t>(z) * needed + i] = response[static_cast<size_t>(z) * rows * cols + p];
    }
}

// Segment bounds of the per-image key lists for the segmented sort.
__global__ void SegmentsKernel(const unsigned *__restrict__ taken, int n_images, unsigned max_taken, int *__restrict__ seg_begin, int *__restrict__ seg_end) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= n_images) return;
    seg_begin[z] = z * static_cast<int>(max_taken);
    seg_end[z] = seg_begin[z] + static_cast<int>(min(taken[z], max_taken));
}

__global__ void EmitKernel(const Key *__restrict__ sorted, const unsigned *__restrict__ taken, unsigned max_taken, int needed, int rows, int cols,
                           const float *__restrict__ response, float2 *__restrict__ uv, float *__restrict__ out_response, int *__restrict__ n_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.z;
    const int n = min(static_cast<int>(min(taken[z], max_taken)), needed);
    if (i == 0) n_out[z] = taken[z] > max_taken ? -1 : n;  // -1: more taken features than can exist (never happens; checked by the host)
    if (i >= n) return;
    const unsigned p = 0xFFFFFFFFu - static_cast<unsigned>(sorted[static_cast<size_t>(z) * max_taken + i] & 0xFFFFFFFFull);
    uv[static_cast<size_t>(z) * needed + i] = make_float2(static_cast<float>(p % cols), static_cast<float>(p / cols));
    if (out_response) out_response[static_cast<size_t>(z) * needed + i] = response[static_cast<size_t>(z) * rows * cols + p];
}

// feat_image (may be null: every feature belongs to the plane `img`) = image of the batch, relative to `img`, of every feature.
__global__ void __launch_bounds__(256) BriefKernel(const uint8_t *__restrict__ img, long long image_stride, const int *__restrict__ feat_image, int rows,
                                                   int cols, int pitch, const float2 *__restrict__ uv, int n, const char4 *__restrict__ pattern, int words,
                                                   int half, uint32_t *__restrict__ desc, uint8_t *__restrict__ valid) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (f >= n) return;
    if (feat_image) img += feat_image[f] * image_stride;
    const float2 p = uv[f];
    const bool ok = p.x >= static_cast<float>(half) && p.y >= static_cast<float>(half) && p.x < static_cast<float>(cols - half) &&
                    p.y < static_cast<float>(rows - half);
    if (lane == 0 && valid) valid[f] = ok;
    if (!ok) {
        for (int w = lane; w < words; w += 32) desc[static_cast<size_t>(f) * words + w] = 0u;
        return;
    }
    const uint8_t *centre = img + static_cast<size_t>(static_cast<int>(p.y)) * pitch + static_cast<int>(p.x);
    for (int w = 0; w < words; ++w) {
        const char4 q = __ldg(pattern + w * 32 + lane);
        const uint8_t va = __ldg(centre + q.x * pitch + q.y), vb = __ldg(centre + q.z * pitch + q.w);
        const unsigned word = __ballot_sync(0xFFFFFFFFu, va < vb);
        if (lane == 0) desc[static_cast<size_t>(f) * words + w] = word;
    }
}

int CheckDetector(ftk_context *ctx, const ftk_detector_params &p, const PyramidView &pyr, int image) {
    if (p.kind != FTK_DETECTOR_HARRIS && p.kind != FTK_DETECTOR_SHI_TOMASI) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "detector kind %d", p.kind);
    if (p.half_patch < 1 || p.half_patch + 1 > kDetMaxMargin) return SetError(ctx, FTK_ERR_UNSUPPORTED, "detector half_patch %d outside 1..3", p.half_patch);
    if (!(p.harris_k >= 0.0f && p.harris_k <= 1.0f) || !std::isfinite(p.min_response))
        return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "detector harris_k must lie in [0, 1] and min_response be finite");
    if (image < 0 || image >= pyr.n_images) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "image %d outside the pyramid batch", image);
    return FTK_OK;
}

}  // namespace

int LaunchDetectResponse(ftk_context *ctx, const ftk_detector_params &p, const PyramidView &pyr, int image, float *d_response) {
    if (int rc = CheckDetector(ctx, p, pyr, image)) return rc;
    const int rows = pyr.rows[0], cols = pyr.cols[0];
    const uint8_t *img = pyr.base[0] + image * pyr.image_stride[0];
    const dim3 grid((cols + kDetTile - 1) / kDetTile, (rows + kDetTile - 1) / kDetTile);
    ResponseKernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(img, 0, rows, cols, pyr.pitch[0], p, d_response, nullptr);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

// Images first .. first + count - 1 of the batch, `needed` features each: d_out_uv [count][needed], d_out_response [count][needed] (or
// null), d_n_out [count] (device).  The images are processed in chunks sized to the scratch budget; every kernel of a chunk covers all
// its images (blockIdx.z), so the launch count per chunk does not depend on the number of images.  Pre-existing features (count == 1
// only) block their windows.  The stream is synchronised once per batch of rounds (convergence test), not per image.
int LaunchDetectFeatures(ftk_context *ctx, const ftk_detector_params &p, const PyramidView &pyr, int first, int count, const float2 *d_existing,
                         int n_existing, int needed, float2 *d_out_uv, float *d_out_response, int *d_n_out) {
    if (int rc = CheckDetector(ctx, p, pyr, first)) return rc;
    if (count < 1 || first + count > pyr.n_images) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "images %d..%d outside the pyramid batch", first, first + count - 1);
    if (n_existing > 0 && count != 1) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "pre-existing features need a single-image call");
    const int rows = pyr.rows[0], cols = pyr.cols[0];
    const size_t n = static_cast<size_t>(rows) * cols;
    if (n >= 0x7FFFFFFFull) return SetError(ctx, FTK_ERR_UNSUPPORTED, "image too large for 32-bit pixel indices");
    cudaStream_t st = ctx->stream;
    if (needed <= 0) {
        FTK_CUDA_CHECK(ctx, cudaMemsetAsync(d_n_out, 0, sizeof(int) * count, st));
        return FTK_OK;
    }
    // a feature blocks |drow| < dist and |dcol| < dist; dist <= 1 blocks nothing but its own pixel, and no window is wider than the image
    const int dist = std::min(std::max(p.min_distance, 1), std::max(rows, cols));
    // taken features are >= dist apart: at most ceil(rows / dist) * ceil(cols / dist) of them
    const size_t max_taken = static_cast<size_t>((rows + dist - 1) / dist) * ((cols + dist - 1) / dist);
    // one fused kernel per round while tile + halo fit in shared memory with two CTAs per SM; two kernels per round beyond that
    const int edge = kDetTile + 2 * (dist - 1);
    const size_t tile_smem = sizeof(State) * (static_cast<size_t>(edge) * kDetTile + static_cast<size_t>(edge) * (edge | 1));
    const bool fused = edge <= 32 * kStageCols && !getenv("FTK_DETECT_TWO_PASS");  // <= 82 KB: at least two CTAs per SM
    const size_t row_smem = sizeof(Key) * (kRowMaxThreads + 2 * (dist - 1));
    if (fused && tile_smem > 48 * 1024)
        FTK_CUDA_CHECK(ctx, cudaFuncSetAttribute(SelectRoundTileKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(tile_smem)));
    if (!fused && row_smem > 48 * 1024)
        FTK_CUDA_CHECK(ctx, cudaFuncSetAttribute(RowMaxKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(row_smem)));
    // scratch per image: response (4 B / pixel), state (4), row maxima (8, two-kernel rounds only), two key lists
    const size_t per_image = n * (fused ? 8 : 16) + 2 * sizeof(Key) * max_taken;
    size_t budget = 1ull << 30;
    if (const char *e = getenv("FTK_DETECT_SCRATCH_BYTES")) budget = static_cast<size_t>(atoll(e));
    int chunk = static_cast<int>(std::min<size_t>(std::max<size_t>(budget / per_image, 1), static_cast<size_t>(count)));
    chunk = std::min({chunk, 65535, static_cast<int>(0x7FFFFFFFull / std::max<size_t>(max_taken, 1))});
    chunk = std::max(chunk, 1);
    if (int rc = EnsureDevice(ctx, ctx->d_det_response, sizeof(float) * n * chunk)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_det_state, (sizeof(State) + (fused ? 0 : sizeof(Key))) * n * chunk + 16)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_det_keys, sizeof(Key) * 2 * max_taken * chunk)) return rc;
    // counters: undecided per round of the current batch of rounds | taken per image | segment begin | segment end
    if (int rc = EnsureDevice(ctx, ctx->d_det_cand, sizeof(unsigned) * (kRoundsPerBatch + 3 * static_cast<size_t>(chunk)))) return rc;
    float *response = static_cast<float *>(ctx->d_det_response.ptr);
    Key *rowmax = static_cast<Key *>(ctx->d_det_state.ptr);  // two-kernel rounds only
    State *state = reinterpret_cast<State *>(fused ? rowmax : rowmax + n * chunk);
    Key *keys = static_cast<Key *>(ctx->d_det_keys.ptr), *sorted = keys + max_taken * chunk;
    unsigned *undecided_dev = static_cast<unsigned *>(ctx->d_det_cand.ptr), *taken = undecided_dev + kRoundsPerBatch;
    int *seg_begin = reinterpret_cast<int *>(taken + chunk), *seg_end = seg_begin + chunk;
    size_t tmp_bytes = 0;
    FTK_CUDA_CHECK(ctx, cub::DeviceSegmentedRadixSort::SortKeysDescending(nullptr, tmp_bytes, keys, sorted, static_cast<int>(max_taken * chunk), chunk,
                                                                          seg_begin, seg_end, 0, 64, st));
    if (int rc = EnsureDevice(ctx, ctx->d_det_tmp, tmp_bytes ? tmp_bytes : 1)) return rc;

    for (int c0 = 0; c0 < count; c0 += chunk) {
        const int m = std::min(chunk, count - c0);
        const uint8_t *img = pyr.base[0] + static_cast<long long>(first + c0) * pyr.image_stride[0];
        FTK_CUDA_CHECK(ctx, cudaMemsetAsync(undecided_dev, 0, sizeof(unsigned) * (kRoundsPerBatch + static_cast<size_t>(chunk)), st));
        const dim3 tiles((cols + kDetTile - 1) / kDetTile, (rows + kDetTile - 1) / kDetTile, m);
        ResponseKernel<<<tiles, dim3(32, 8), 0, st>>>(img, pyr.image_stride[0], rows, cols, pyr.pitch[0], p, response, state);
        ++ctx->launches;
        FTK_CUDA_CHECK(ctx, cudaGetLastError());
        if (n_existing > 0 && p.min_distance > 0) {
            ExistingMaskKernel<<<n_existing, 128, 0, st>>>(d_existing, rows, cols, dist, state);
            ++ctx->launches;
            FTK_CUDA_CHECK(ctx, cudaGetLastError());
        }
        const dim3 row_grid((cols + kRowMaxThreads - 1) / kRowMaxThreads, rows, m), decide_grid((cols + 31) / 32, (rows + 7) / 8, m);
        for (int done = 0;; done += kRoundsPerBatch) {
            if (done >= kMaxRounds) return SetError(ctx, FTK_ERR_CUDA, "feature selection did not converge in %d rounds", kMaxRounds);
            if (done) FTK_CUDA_CHECK(ctx, cudaMemsetAsync(undecided_dev, 0, sizeof(unsigned) * kRoundsPerBatch, st));
            for (int k = 0; k < kRoundsPerBatch; ++k) {
                if (fused) {
                    SelectRoundTileKernel<<<tiles, dim3(32, 8), tile_smem, st>>>(state, rows, cols, dist, undecided_dev, k);
                    ++ctx->launches;
                } else {
                    RowMaxKernel<<<row_grid, kRowMaxThreads, row_smem, st>>>(state, rows, cols, dist, rowmax, undecided_dev, k);
                    DecideKernel<<<decide_grid, dim3(32, 8), 0, st>>>(rowmax, state, rows, cols, dist, undecided_dev, k);
                    ctx->launches += 2;
                }
            }
            FTK_CUDA_CHECK(ctx, cudaGetLastError());
            unsigned undecided[kRoundsPerBatch];
            FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(undecided, undecided_dev, sizeof(undecided), cudaMemcpyDeviceToHost, st));
            FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
            bool converged = false;
            for (int k = 0; k < kRoundsPerBatch; ++k) converged |= undecided[k] == 0;
            if (converged) break;
        }
        CollectKernel<<<dim3(static_cast<unsigned>((n + 255) / 256), 1, m), 256, 0, st>>>(state, static_cast<unsigned>(n), response, keys,
                                                                                         static_cast<unsigned>(max_taken), taken);
        ++ctx->launches;
        FTK_CUDA_CHECK(ctx, cudaGetLastError());
        float2 *uv_out = d_out_uv + static_cast<size_t>(c0) * needed;
        float *resp_out = d_out_response ? d_out_response + static_cast<size_t>(c0) * needed : nullptr;
        if (max_taken <= kSortMax) {  // the usual case: a few hundred features per image
            int padded = 1;
            while (padded < static_cast<int>(max_taken)) padded <<= 1;
            SortEmitKernel<<<m, 256, sizeof(Key) * padded, st>>>(keys, taken, static_cast<unsigned>(max_taken), needed, rows, cols, response, uv_out, resp_out,
                                                                d_n_out + c0);
            ++ctx->launches;
            FTK_CUDA_CHECK(ctx, cudaGetLastError());
            continue;
        }
        SegmentsKernel<<<(m + 127) / 128, 128, 0, st>>>(taken, m, static_cast<unsigned>(max_taken), seg_begin, seg_end);
        ++ctx->launches;
        FTK_CUDA_CHECK(ctx, cudaGetLastError());
        FTK_CUDA_CHECK(ctx, cub::DeviceSegmentedRadixSort::SortKeysDescending(ctx->d_det_tmp.ptr, tmp_bytes, keys, sorted, static_cast<int>(max_taken * m), m,
                                                                              seg_begin, seg_end, 0, 64, st));
        ++ctx->launches;
        const int n_emit_max = static_cast<int>(std::min<size_t>(max_taken, static_cast<size_t>(needed)));
        EmitKernel<<<dim3((n_emit_max + 255) / 256, 1, m), 256, 0, st>>>(sorted, taken, static_cast<unsigned>(max_taken), needed, rows, cols, response,
                                                                        uv_out, resp_out, d_n_out + c0);
        ++ctx->launches;
        FTK_CUDA_CHECK(ctx, cudaGetLastError());
    }
    return FTK_OK;
}

// d_feat_image null: all n features lie on image `first`; otherwise feature f lies on image first + d_feat_image[f].
int LaunchDescribeBrief(ftk_context *ctx, const PyramidView &pyr, int first, int count, const int *d_feat_image, const float2 *d_uv, int n,
                        const char4 *d_pattern, int n_bits, int half_patch, uint32_t *d_desc, uint8_t *d_valid) {
    if (first < 0 || count < 1 || first + count > pyr.n_images)
        return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "images %d..%d outside the pyramid batch", first, first + count - 1);
    if (n == 0) return FTK_OK;
    const uint8_t *img = pyr.base[0] + first * pyr.image_stride[0];
    BriefKernel<<<(n + 7) / 8, 256, 0, ctx->stream>>>(img, pyr.image_stride[0], d_feat_image, pyr.rows[0], pyr.cols[0], pyr.pitch[0], d_uv, n, d_pattern,
                                                     n_bits / 32, half_patch, d_desc, d_valid);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

}  // namespace ftk
