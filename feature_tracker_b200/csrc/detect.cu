// K9 / K10: corner detection and BRIEF description on the device pyramid.
//
// Stands in for feature_detector::FeaturePointHarrisDetector::DetectGoodFeatures and feature_detector::BriefDescriptor::Compute
// (called at test/test_descriptor_matcher_brief.cpp:59-76 and test/test_optical_flow.cpp:60-66).  Those classes belong to the
// sibling repository Feature_Detector, which is not part of the reference tree: PARITY IS UNPINNED.  The arithmetic is the
// published algorithm as frozen in the detector section of oracle/ftk_oracle.c, and
// these kernels are bit-exact against that restatement.
//
// K9a ResponseKernel   one pass over the image (HBM bound: 1 B read + 8 B written per pixel): integer structure-tensor sums from a
//                      shared-memory tile, fp32 response with explicit rounding, candidate list by warp-aggregated append.
// K9b SelectRoundKernel  the sequential "visit by falling response, take unless a taken feature is near" loop as a parallel fixed
//                      point: a candidate is TAKEN once no undecided or taken candidate of higher priority is left in its
//                      window, DROPPED once a taken one is; every decision is final and equals the sequential loop's, rounds
//                      repeat until no candidate is undecided.  Priority = (response, then lower row-major index).
// K9c CollectKernel + sort + EmitKernel   the taken set ordered by priority; the first `needed` are the sequential loop's output.
// K10 BriefKernel      one warp per feature, one pair per lane, a ballot per 32-bit descriptor word.
#include <cub/device/device_radix_sort.cuh>

#include <cmath>

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

__global__ void __launch_bounds__(256) ResponseKernel(const uint8_t *__restrict__ img, int rows, int cols, int pitch, ftk_detector_params p,
                                                      float *__restrict__ response, float *__restrict__ state, unsigned *__restrict__ cand,
                                                      unsigned *__restrict__ n_cand) {
    __shared__ uint8_t tile[kDetTile + 2 * kDetMaxMargin][kDetTile + 2 * kDetMaxMargin + 8];
    const int h = p.half_patch, m = h + 1, edge = kDetTile + 2 * m;
    const int r0 = blockIdx.y * kDetTile, c0 = blockIdx.x * kDetTile;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int k = tid; k < edge * edge; k += 256) {
        const int tr = k / edge, tc = k - tr * edge;
        const int r = r0 - m + tr, c = c0 - m + tc;
        tile[tr][tc] = (r >= 0 && r < rows && c >= 0 && c < cols) ? img[static_cast<size_t>(r) * pitch + c] : 0;
    }
    __syncthreads();
    const float inv = __fdiv_rn(1.0f, static_cast<float>(4 * (2 * h + 1) * (2 * h + 1)));
#pragma unroll
    for (int q = 0; q < kDetTile / 8; ++q) {
        const int lr = threadIdx.y + 8 * q, lc = threadIdx.x;
        const int r = r0 + lr, c = c0 + lc;
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
        const bool inside = r < rows && c < cols;
        const bool is_cand = inside && resp >= p.min_response;
        if (inside) {
            const size_t i = static_cast<size_t>(r) * cols + c;
            response[i] = resp;
            if (state) state[i] = is_cand ? resp : -INFINITY;
        }
        if (cand) {  // warp-aggregated append; the list's order does not matter
            const unsigned vote = __ballot_sync(0xFFFFFFFFu, is_cand);
            if (vote) {
                unsigned base = 0;
                const int leader = __ffs(vote) - 1;
                if (threadIdx.x == leader) base = atomicAdd(n_cand, __popc(vote));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if (is_cand) cand[base + __popc(vote & ((1u << threadIdx.x) - 1u))] = static_cast<unsigned>(r) * cols + c;
            }
        }
    }
}

// Features the caller already holds block their window (they are not candidates themselves).
__global__ void ExistingMaskKernel(const float2 *__restrict__ existing, int rows, int cols, int dist, float *__restrict__ state) {
    const float2 f = existing[blockIdx.x];
    if (!(f.x >= 0.0f && f.y >= 0.0f && f.x < static_cast<float>(cols) && f.y < static_cast<float>(rows))) return;
    const int r = static_cast<int>(f.y), c = static_cast<int>(f.x), w = 2 * dist - 1;
    for (int k = threadIdx.x; k < w * w; k += blockDim.x) {
        const int rr = r - (dist - 1) + k / w, qc = c - (dist - 1) + k % w;
        if (rr >= 0 && rr < rows && qc >= 0 && qc < cols) state[static_cast<size_t>(rr) * cols + qc] = -INFINITY;
    }
}

// state[pixel]: the response while the candidate is undecided, +inf once taken, -inf when dropped / never a candidate.
// One warp per candidate.  counters[round] = candidates still undecided after this round.
__global__ void __launch_bounds__(256) SelectRoundKernel(const unsigned *__restrict__ cand, const unsigned *__restrict__ n_cand, float *state, int rows,
                                                         int cols, int dist, unsigned *__restrict__ counters, int round) {
    if (round > 0 && counters[round - 1] == 0) return;  // converged earlier in this batch
    const unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= *n_cand) return;
    const unsigned p = cand[w];
    const float s = __ldcg(state + p);
    if (s == INFINITY || s == -INFINITY) return;
    const int r = p / cols, c = p - r * cols;
    const int ra = max(0, r - (dist - 1)), rb = min(rows - 1, r + (dist - 1)), ca = max(0, c - (dist - 1)), cb = min(cols - 1, c + (dist - 1));
    bool taken_near = false, higher_near = false;
    for (int rr = ra; rr <= rb; ++rr) {
        const float *row = state + static_cast<size_t>(rr) * cols;
        for (int qc = ca + lane; qc <= cb; qc += 32) {
            const float v = __ldcg(row + qc);
            const unsigned q = static_cast<unsigned>(rr) * cols + qc;
            taken_near |= v == INFINITY;
            higher_near |= v > s || (v == s && q < p);
        }
    }
    taken_near = __any_sync(0xFFFFFFFFu, taken_near);
    higher_near = __any_sync(0xFFFFFFFFu, higher_near);
    if (lane == 0) {
        if (taken_near) __stcg(state + p, -INFINITY);
        else if (!higher_near) __stcg(state + p, INFINITY);
        else atomicAdd(counters + round, 1u);
    }
}

__global__ void CollectKernel(const unsigned *__restrict__ cand, const unsigned *__restrict__ n_cand, const float *__restrict__ state,
                              const float *__restrict__ response, unsigned long long *__restrict__ keys, unsigned *__restrict__ n_keys) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_cand) return;
    const unsigned p = cand[i];
    if (state[p] != INFINITY) return;
    // -0 + 0 = +0: one key for both zeros, like the float comparison of the sequential loop
    keys[atomicAdd(n_keys, 1u)] = (static_cast<unsigned long long>(DetOrderMap(__fadd_rn(response[p], 0.0f))) << 32) | (0xFFFFFFFFu - p);
}

__global__ void EmitKernel(const unsigned long long *__restrict__ keys, int n, int cols, const float *__restrict__ response, float2 *__restrict__ uv,
                           float *__restrict__ out_response) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned p = 0xFFFFFFFFu - static_cast<unsigned>(keys[i] & 0xFFFFFFFFull);
    uv[i] = make_float2(static_cast<float>(p % cols), static_cast<float>(p / cols));
    if (out_response) out_response[i] = response[p];
}

__global__ void __launch_bounds__(256) BriefKernel(const uint8_t *__restrict__ img, int rows, int cols, int pitch, const float2 *__restrict__ uv, int n,
                                                   const char4 *__restrict__ pattern, int words, int half, uint32_t *__restrict__ desc,
                                                   uint8_t *__restrict__ valid) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (f >= n) return;
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
    ResponseKernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(img, rows, cols, pyr.pitch[0], p, d_response, nullptr, nullptr, nullptr);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

int LaunchDetectFeatures(ftk_context *ctx, const ftk_detector_params &p, const PyramidView &pyr, int image, const float2 *d_existing, int n_existing,
                         int needed, float2 *d_out_uv, float *d_out_response, int *n_out) {
    *n_out = 0;
    if (int rc = CheckDetector(ctx, p, pyr, image)) return rc;
    const int rows = pyr.rows[0], cols = pyr.cols[0];
    const size_t n = static_cast<size_t>(rows) * cols;
    if (n >= 0xFFFFFFFFull) return SetError(ctx, FTK_ERR_UNSUPPORTED, "image too large for 32-bit pixel indices");
    if (needed <= 0) return FTK_OK;
    const uint8_t *img = pyr.base[0] + image * pyr.image_stride[0];
    cudaStream_t st = ctx->stream;
    // counters: [0] candidates, [1] taken, [2 ..] undecided per round of the current batch
    const size_t n_counters = 2 + kRoundsPerBatch;
    if (int rc = EnsureDevice(ctx, ctx->d_det_response, sizeof(float) * n)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_det_state, sizeof(float) * n)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_det_cand, sizeof(unsigned) * (n + n_counters))) return rc;
    float *response = static_cast<float *>(ctx->d_det_response.ptr), *state = static_cast<float *>(ctx->d_det_state.ptr);
    unsigned *cand = static_cast<unsigned *>(ctx->d_det_cand.ptr), *counters = cand + n;
    FTK_CUDA_CHECK(ctx, cudaMemsetAsync(counters, 0, sizeof(unsigned) * n_counters, st));
    const dim3 grid((cols + kDetTile - 1) / kDetTile, (rows + kDetTile - 1) / kDetTile);
    ResponseKernel<<<grid, dim3(32, 8), 0, st>>>(img, rows, cols, pyr.pitch[0], p, response, state, cand, counters);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    const int dist = p.min_distance > 0 ? p.min_distance : 1;  // d <= 1: a feature blocks nothing but its own pixel
    if (n_existing > 0 && p.min_distance > 0) {
        ExistingMaskKernel<<<n_existing, 128, 0, st>>>(d_existing, rows, cols, dist, state);
        ++ctx->launches;
        FTK_CUDA_CHECK(ctx, cudaGetLastError());
    }
    unsigned n_cand = 0;
    FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(&n_cand, counters, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    if (n_cand == 0) return FTK_OK;
    const unsigned round_blocks = (n_cand + 7) / 8;
    for (int done = 0;; done += kRoundsPerBatch) {
        if (done >= kMaxRounds) return SetError(ctx, FTK_ERR_CUDA, "feature selection did not converge in %d rounds", kMaxRounds);
        if (done) FTK_CUDA_CHECK(ctx, cudaMemsetAsync(counters + 2, 0, sizeof(unsigned) * kRoundsPerBatch, st));
        for (int k = 0; k < kRoundsPerBatch; ++k) {
            SelectRoundKernel<<<round_blocks, 256, 0, st>>>(cand, counters, state, rows, cols, dist, counters + 2, k);
            ++ctx->launches;
        }
        FTK_CUDA_CHECK(ctx, cudaGetLastError());
        unsigned undecided[kRoundsPerBatch];
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(undecided, counters + 2, sizeof(undecided), cudaMemcpyDeviceToHost, st));
        FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
        bool converged = false;
        for (int k = 0; k < kRoundsPerBatch; ++k) converged |= undecided[k] == 0;
        if (converged) break;
    }
    if (int rc = EnsureDevice(ctx, ctx->d_det_keys, sizeof(unsigned long long) * 2 * n_cand)) return rc;
    unsigned long long *keys = static_cast<unsigned long long *>(ctx->d_det_keys.ptr), *sorted = keys + n_cand;
    CollectKernel<<<(n_cand + 255) / 256, 256, 0, st>>>(cand, counters, state, response, keys, counters + 1);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    unsigned n_taken = 0;
    FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(&n_taken, counters + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    if (n_taken == 0) return FTK_OK;
    size_t tmp_bytes = 0;
    FTK_CUDA_CHECK(ctx, cub::DeviceRadixSort::SortKeysDescending(nullptr, tmp_bytes, keys, sorted, static_cast<int>(n_taken), 0, 64, st));
    if (int rc = EnsureDevice(ctx, ctx->d_det_tmp, tmp_bytes ? tmp_bytes : 1)) return rc;
    FTK_CUDA_CHECK(ctx, cub::DeviceRadixSort::SortKeysDescending(ctx->d_det_tmp.ptr, tmp_bytes, keys, sorted, static_cast<int>(n_taken), 0, 64, st));
    ++ctx->launches;
    const int n_emit = static_cast<int>(n_taken) < needed ? static_cast<int>(n_taken) : needed;
    EmitKernel<<<(n_emit + 255) / 256, 256, 0, st>>>(sorted, n_emit, cols, response, d_out_uv, d_out_response);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    *n_out = n_emit;
    return FTK_OK;
}

int LaunchDescribeBrief(ftk_context *ctx, const PyramidView &pyr, int image, const float2 *d_uv, int n, const char4 *d_pattern, int n_bits,
                        int half_patch, uint32_t *d_desc, uint8_t *d_valid) {
    if (image < 0 || image >= pyr.n_images) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "image %d outside the pyramid batch", image);
    if (n == 0) return FTK_OK;
    const uint8_t *img = pyr.base[0] + image * pyr.image_stride[0];
    BriefKernel<<<(n + 7) / 8, 256, 0, ctx->stream>>>(img, pyr.rows[0], pyr.cols[0], pyr.pitch[0], d_uv, n, d_pattern, n_bits / 32, half_patch, d_desc, d_valid);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

}  // namespace ftk
