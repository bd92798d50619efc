// K8: mutual arg-max post-processing of a matching score matrix.
//
// Replaces the score-matrix branch of NNFeatureMatcher::Match (src/nn_feature_matcher/nn_feature_matcher.cpp:180-216): the
// LightGlue network hands back an n_ref x n_cur matrix of log-assignment scores; the reference then
//   1. finds, for every column j, the row of its maximum         (:187-198, first maximum wins: `scores(i) > max_score`)
//   2. finds, for every row i, the column of its maximum         (:200-209, same rule)
//   3. keeps (i, j) when max_score >= kMinValidMatchScore (:210) and the column's maximum row is i (:211).
// Result contract (bit-exact indices): idx[i] = j for kept rows, -1 otherwise.  NaN follows the reference's comparisons:
// a NaN never replaces the running maximum, and a NaN in the first element of a row / column is never replaced.
//
// One pass over the matrix (HBM bound: 4 * n_ref * n_cur bytes): a CTA owns a band of rows x 2048 columns; every warp streams
// 2 x 512 B of each row (float4 per lane), keeps eight running column maxima per lane in registers and reduces the row
// maximum with shuffles; partial results are merged across CTAs with 64-bit atomicMax on (ordered score << 32 | ~index) keys,
// which makes "highest score, then lowest index" a plain integer maximum.
#include <cfloat>
#include <cmath>

#include "ftk_internal.h"

namespace ftk {

namespace {

constexpr int kMutThreads = 256;
constexpr int kMutWarps = kMutThreads / 32;
constexpr int kMutColsPerWarp = 256;  // two float4 per lane
constexpr int kMutColsPerCta = kMutWarps * kMutColsPerWarp;
constexpr int kMutMaxRowsPerCta = 64;

// Monotonic map float -> unsigned (no NaN here; -0 was folded into +0).
__device__ __forceinline__ unsigned OrderMap(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float OrderUnmap(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }
__device__ __forceinline__ unsigned long long MakeKey(float v, int index) {
    return (static_cast<unsigned long long>(OrderMap(v)) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(index));
}

template <bool VEC>
__global__ void __launch_bounds__(kMutThreads) MutualMaxKernel(const float *__restrict__ scores, int n_ref, int n_cur, int rows_per_cta,
                                                              unsigned long long *__restrict__ rowkey, unsigned long long *__restrict__ colkey) {
    __shared__ unsigned long long rowpart[kMutMaxRowsPerCta][kMutWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(n_ref, r0 + rows_per_cta);
    const int c0 = blockIdx.x * kMutColsPerCta + warp * kMutColsPerWarp + lane * 4;  // my columns: c0 .. c0+3 and c0+128 .. c0+131

    float cv[8];
    int cr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cv[k] = -INFINITY, cr[k] = r0;

    // software pipeline: the loads of row r + 1 are in flight while row r is reduced
    auto load_row = [&](int r, float (&v)[8]) {
        const float *row = scores + static_cast<size_t>(r) * n_cur;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = c0 + h * 128;
            if (VEC) {
                float4 q = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                if (c < n_cur) q = __ldg(reinterpret_cast<const float4 *>(row + c));  // n_cur % 4 == 0: all four or none
                v[4 * h + 0] = q.x, v[4 * h + 1] = q.y, v[4 * h + 2] = q.z, v[4 * h + 3] = q.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[4 * h + k] = (c + k < n_cur) ? __ldg(row + c + k) : -INFINITY;
            }
        }
    };
    float nxt[8];
    if (r0 < r1) load_row(r0, nxt);
    for (int r = r0; r < r1; ++r) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = nxt[k];
        if (r + 1 < r1) load_row(r + 1, nxt);
        // A NaN never passes `x > max`, and -0 == +0 leaves the earlier index in place: the raw values can be compared as they are.
        float lane_max = -INFINITY;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (v[k] > cv[k]) cv[k] = v[k], cr[k] = r;  // rows ascend: strict > keeps the first maximum
            lane_max = fmaxf(lane_max, v[k]);            // fmaxf drops NaN operands
        }
        float bv = lane_max;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bv = fmaxf(bv, __shfl_xor_sync(0xFFFFFFFFu, bv, o));
        // first lane holding the row maximum, then its first such column (my columns ascend with k; lanes ascend with columns
        // inside each 128-column half, so the halves are resolved separately)
        int k0 = 8, k1 = 8;
#pragma unroll
        for (int k = 3; k >= 0; --k) {
            if (v[k] == bv) k0 = k;
            if (v[4 + k] == bv) k1 = k;
        }
        const unsigned has0 = __ballot_sync(0xFFFFFFFFu, k0 < 8), has1 = __ballot_sync(0xFFFFFFFFu, k1 < 8);
        int bi;  // a row of NaN / -inf only: the band's first column (-inf ties resolve to the lowest index overall)
        if (has0) {
            const int src = __ffs(has0) - 1;
            bi = (c0 - lane * 4) + src * 4 + __shfl_sync(0xFFFFFFFFu, k0, src);
        } else if (has1) {
            const int src = __ffs(has1) - 1;
            bi = (c0 - lane * 4) + 128 + src * 4 + __shfl_sync(0xFFFFFFFFu, k1, src);
        } else {
            bi = c0 - lane * 4;
        }
        if (lane == 0) rowpart[r - r0][warp] = MakeKey(__fadd_rn(bv, 0.0f), bi);  // -0 + 0 = +0: one key for both zeros
    }
    __syncthreads();
    if (threadIdx.x < r1 - r0) {
        unsigned long long key = rowpart[threadIdx.x][0];
#pragma unroll
        for (int w = 1; w < kMutWarps; ++w) key = max(key, rowpart[threadIdx.x][w]);
        atomicMax(rowkey + r0 + threadIdx.x, key);
    }
    if (r1 > r0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = c0 + (k >> 2) * 128 + (k & 3);
            if (c < n_cur) atomicMax(colkey + c, MakeKey(__fadd_rn(cv[k], 0.0f), cr[k]));
        }
    }
}

__global__ void MutualFinalizeKernel(const float *__restrict__ scores, int n_ref, int n_cur, float min_score, const unsigned long long *__restrict__ rowkey,
                                     const unsigned long long *__restrict__ colkey, int *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ref) return;
    const unsigned long long rk = rowkey[i];
    int j = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(rk & 0xFFFFFFFFull));
    float max_score = OrderUnmap(static_cast<unsigned>(rk >> 32));
    const float first = scores[static_cast<size_t>(i) * n_cur];
    if (first != first) j = 0, max_score = first;  // a NaN in column 0 is the row's running maximum for good (:203-208)
    int result = -1;
    if (!(max_score < min_score)) {  // CONTINUE_IF(max_score < kMinValidMatchScore)
        int col_row = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(colkey[j] & 0xFFFFFFFFull));
        const float top = scores[j];
        if (top != top) col_row = 0;  // same for a NaN in row 0 of column j (:190-196)
        if (col_row == i) result = j;  // CONTINUE_IF(max_scores_in_cols_index[max_score_index] != idx_ref)
    }
    idx[i] = result;
}

// idx_fwd[i] survives only if the opposite direction agrees (two ForceMatch calls + this filter = cross-check matching).
__global__ void CrossCheckKernel(int *__restrict__ idx_fwd, int n_ref, const int *__restrict__ idx_bwd, int n_cur) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ref) return;
    const int j = idx_fwd[i];
    if (j < 0 || j >= n_cur || idx_bwd[j] != i) idx_fwd[i] = -1;
}

}  // namespace

int LaunchMutualScores(ftk_context *ctx, const float *d_scores, int n_ref, int n_cur, float min_score, int *d_idx) {
    if (n_ref == 0) return FTK_OK;
    // row keys then column keys in one buffer: one memset (key 0 is below every real key)
    if (int rc = EnsureDevice(ctx, ctx->d_work0, sizeof(unsigned long long) * (static_cast<size_t>(n_ref) + n_cur))) return rc;
    unsigned long long *rowkey = static_cast<unsigned long long *>(ctx->d_work0.ptr), *colkey = rowkey + n_ref;
    FTK_CUDA_CHECK(ctx, cudaMemsetAsync(rowkey, 0, sizeof(unsigned long long) * (static_cast<size_t>(n_ref) + n_cur), ctx->stream));
    // row band per CTA: as tall as possible (fewer column atomics) while the grid still covers the GPU about twice
    const int col_blocks = (n_cur + kMutColsPerCta - 1) / kMutColsPerCta;
    int rows_per_cta = kMutMaxRowsPerCta;
    while (rows_per_cta > 16 && static_cast<long long>(col_blocks) * ((n_ref + rows_per_cta - 1) / rows_per_cta) < 2LL * ctx->sm_count) rows_per_cta >>= 1;
    const dim3 grid(col_blocks, (n_ref + rows_per_cta - 1) / rows_per_cta);
    const bool vec = n_cur % 4 == 0 && reinterpret_cast<uintptr_t>(d_scores) % 16 == 0;
    if (vec) MutualMaxKernel<true><<<grid, kMutThreads, 0, ctx->stream>>>(d_scores, n_ref, n_cur, rows_per_cta, rowkey, colkey);
    else MutualMaxKernel<false><<<grid, kMutThreads, 0, ctx->stream>>>(d_scores, n_ref, n_cur, rows_per_cta, rowkey, colkey);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    MutualFinalizeKernel<<<(n_ref + 255) / 256, 256, 0, ctx->stream>>>(d_scores, n_ref, n_cur, min_score, rowkey, colkey, d_idx);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

int LaunchCrossCheck(ftk_context *ctx, int *d_idx_fwd, int n_ref, const int *d_idx_bwd, int n_cur) {
    if (n_ref == 0) return FTK_OK;
    CrossCheckKernel<<<(n_ref + 255) / 256, 256, 0, ctx->stream>>>(d_idx_fwd, n_ref, d_idx_bwd, n_cur);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

}  // namespace ftk
