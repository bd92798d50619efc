// Internal declarations shared by the translation units of libftk_b200.so (not part of the public ABI).
#ifndef FTK_INTERNAL_H_
#define FTK_INTERNAL_H_

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "ftk_c.h"

namespace ftk {

constexpr int kMaxLevels = 10;  // oracle/shim/datatype_image_pyramid.h: kPyramidMaxLevel

// Device-side view of one pyramid batch.  Level l of image i starts at base[l] + i * image_stride[l]; rows are
// `pitch[l]` bytes apart (a multiple of 16) and every plane is followed by at least pitch[l] + 16 readable bytes,
// so the bilinear sampler's weight-0 "+1" neighbour on the last row / column never leaves the allocation.
struct PyramidView {
    const uint8_t *base[kMaxLevels];
    long long image_stride[kMaxLevels];
    int rows[kMaxLevels];
    int cols[kMaxLevels];
    int pitch[kMaxLevels];
    int levels;
    int n_images;
};

}  // namespace ftk

struct ftk_pyramid {
    ftk::PyramidView view;
    uint8_t *storage = nullptr;
    size_t storage_bytes = 0;
    int device = 0;
};

// Growable device / pinned-host scratch buffer owned by a context.
struct FtkBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct ftk_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    // the pipelined entry points (ftk_track_image_pairs*, ftk_track_image_sequence): one copy stream for every host-to-device byte,
    // kStageBuffers rotating staging pyramids with one compute stream each
    static constexpr int kStageBuffers = 3;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t chunk_stream[kStageBuffers] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_copied[kStageBuffers] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_computed[kStageBuffers] = {nullptr, nullptr, nullptr};
    ftk_pyramid *stage_pyr[kStageBuffers] = {nullptr, nullptr, nullptr};
    int stage_rows = 0, stage_cols = 0, stage_levels = 0, stage_pairs = 0;
    std::string error;
    uint64_t launches = 0;
    bool profiling = false;                        // ftk_set_profiling: events around the dominant kernel of a call
    cudaEvent_t ev_prof[2] = {nullptr, nullptr};
    bool prof_recorded = false;
    int sm_count = 0;
    const int *d_last_scan_items = nullptr;  // device counter of the last tensor-core cosine match (nullptr: path not used)
    // tensor-core cosine match: two counter sets used alternately -- call n counts in set n & 1 while its first kernel zeroes the
    // other set for call n + 1 (no memset launch per call); `clean` is false until a call has left both sets in that state
    FtkBuffer d_cos_counters;
    unsigned cos_calls = 0;
    bool cos_counters_clean = false;
    // NearbyMatch grid: bounds / ticket / grid descriptor / per-cell counts live in one buffer whose "between calls" state (bounds at
    // their initial values, every count zero) is restored by the kernels themselves, so a call needs neither a memset nor a copy
    FtkBuffer d_nearby_state;
    bool nearby_state_clean = false;
    bool use_fast_paths = true;  // FTK_DISABLE_FASTPATH=1 forces the generic kernels (A/B testing)
    // device scratch
    FtkBuffer d_ref_uv, d_cur_uv, d_status, d_offsets, d_ref_img, d_cur_img, d_feat_pair;
    FtkBuffer d_chunk_offsets, d_chunk_curmap;
    FtkBuffer d_back_uv, d_back_status;  // forward-backward pass scratch
    FtkBuffer d_small;                   // small host-pointer calls: every input / output array in one device block ...
    void *h_small = nullptr;             // ... mirrored by one pinned host block (one H2D + one D2H per call)
    size_t h_small_bytes = 0;
    FtkBuffer d_dm_K, d_dm_points, d_dm_q, d_dm_p;  // direct-method staging
    FtkBuffer d_flow;  // dense-flow staging (2 planes)
    FtkBuffer d_det_response, d_det_state, d_det_cand, d_det_keys, d_det_tmp, d_det_out, d_det_pattern;  // detector / BRIEF scratch
    FtkBuffer d_desc_ref, d_desc_cur, d_idx, d_pred_uv, d_pos_cur, d_work0, d_work1, d_work2, d_work3;
};

namespace ftk {

int SetError(ftk_context *ctx, int code, const char *fmt, ...);
// ftk_set_profiling: bracket the dominant kernel of a call (no-ops unless profiling is enabled)
inline void ProfBegin(ftk_context *ctx) {
    if (ctx->profiling && ctx->ev_prof[0]) cudaEventRecord(ctx->ev_prof[0], ctx->stream);
}
inline void ProfEnd(ftk_context *ctx) {
    if (ctx->profiling && ctx->ev_prof[1]) {
        cudaEventRecord(ctx->ev_prof[1], ctx->stream);
        ctx->prof_recorded = true;
    }
}
int EnsureDevice(ftk_context *ctx, FtkBuffer &buf, size_t bytes);

#ifdef __CUDACC__
// Programmatic dependent launch: the kernels of one call are launched with the stream-serialization attribute, so that the launch
// latency and the prologue of kernel n + 1 overlap the tail of kernel n.  Every such kernel announces its dependents at once and waits
// for its predecessor (complete and flushed) right before its first access to anything an earlier kernel of the stream may have written.
__device__ __forceinline__ void GridDepLaunchDependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void GridDepWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... Params, typename... Args>
cudaError_t LaunchDependent(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}
#endif

#define FTK_CUDA_CHECK(ctx, expr)                                                                              \
    do {                                                                                                       \
        cudaError_t err__ = (expr);                                                                            \
        if (err__ != cudaSuccess) {                                                                            \
            return ftk::SetError((ctx), FTK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                                 __FILE__, __LINE__);                                                          \
        }                                                                                                      \
    } while (0)

// pyramid.cu
int LaunchPyramidBuild(ftk_context *ctx, ftk_pyramid *pyr, int first, int count);

// klt.cu
struct KltLaunch {
    ftk_klt_params p;
    PyramidView ref, cur;
    int n_pairs;
    int n_features;
    const int *ref_image;   // device, may be null
    const int *cur_image;   // device, may be null
    const int *feat_offsets;  // device [n_pairs + 1]
    const int *feat_pair;     // device [n_features]
    const float2 *ref_uv;
    float2 *cur_uv;
    uint8_t *status;
    int has_prediction;
    int has_status;
    int single_level;
};
int LaunchFeaturePairs(ftk_context *ctx, const int *d_offsets, int n_pairs, int n_features, int *d_feat_pair);
int LaunchKltTrack(ftk_context *ctx, const KltLaunch &launch);
// Forward pass, then (p.forward_backward_max_error > 0) the backward pass and the consistency test; scratch is indexed like the
// launch's feature arrays starting at scratch_offset.
int LaunchKltTrackChecked(ftk_context *ctx, const KltLaunch &launch, size_t scratch_offset = 0);
// klt_basic_fastpath.cu: FTK_ERR_UNSUPPORTED when no specialisation covers the configuration
int LaunchKltBasicFastPath(ftk_context *ctx, const KltLaunch &launch);

// direct_method.cu: one CTA per frame pair
int LaunchDirectMethod(ftk_context *ctx, const ftk_direct_params &p, const PyramidView &ref, const PyramidView &cur, int n_pairs, const int *d_ref_image,
                       const int *d_cur_image, const int *d_offsets, const float *d_K, const float *d_points, const float2 *d_ref_uv, float2 *d_cur_uv,
                       float *d_q_rc, float *d_p_rc, uint8_t *d_status, int n_features, bool has_prediction, bool has_status);

// dense_flow.cu
int LaunchDenseFlow(ftk_context *ctx, const ftk_dense_flow_params &p, const PyramidView &ref, const PyramidView &cur, int ref_image, int cur_image,
                    bool single_level, bool use_initial_flow, float *d_flow_r, float *d_flow_c);

// detect.cu: corner response, greedy min-distance selection, BRIEF bits (level 0 of one pyramid image)
int LaunchDetectResponse(ftk_context *ctx, const ftk_detector_params &p, const PyramidView &pyr, int image, float *d_response);
int LaunchDetectFeatures(ftk_context *ctx, const ftk_detector_params &p, const PyramidView &pyr, int first, int count, const float2 *d_existing,
                         int n_existing, int needed, float2 *d_out_uv, float *d_out_response, int *d_n_out);
int LaunchDescribeBrief(ftk_context *ctx, const PyramidView &pyr, int first, int count, const int *d_feat_image, const float2 *d_uv, int n,
                        const char4 *d_pattern, int n_bits, int half_patch, uint32_t *d_desc, uint8_t *d_valid);

// match.cu
int LaunchHammingForce(ftk_context *ctx, const uint32_t *d_ref, int n_ref, const uint32_t *d_cur, int n_cur, int words, float max_dist, int *d_idx,
                       bool fill_unmatched);
// fill_unmatched (here and below): d_idx holds no input -- rows without a match get -1, written with the results instead of a memset
int LaunchHammingNearby(ftk_context *ctx, const uint32_t *d_ref, int n_ref, const uint32_t *d_cur, int n_cur, int words, const float2 *d_pred,
                        const float2 *d_pos, int max_drow, int max_dcol, float max_dist, int *d_idx, bool fill_unmatched);
int LaunchHammingPairs(ftk_context *ctx, const uint32_t *d_ref, const uint32_t *d_cur, int words, int n_ref_total, const int *d_ref_pair,
                       const int *d_ref_off, const int *d_cur_off, const float2 *d_pred, const float2 *d_pos, int max_drow, int max_dcol, float max_dist,
                       int *d_idx);
int LaunchCosinePairs(ftk_context *ctx, const float *d_ref, const float *d_cur, int dim, int n_ref_total, int n_cur_total, const int *d_ref_pair,
                      const int *d_cur_off, const float2 *d_pred, const float2 *d_pos, int max_drow, int max_dcol, float max_dist, int *d_idx);
int LaunchCosineForce(ftk_context *ctx, const float *d_ref, int n_ref, const float *d_cur, int n_cur, int dim, float max_dist, int *d_idx, bool fill_unmatched);
// match_mutual.cu: mutual arg-max of a score matrix; cross-check filter
int LaunchMutualScores(ftk_context *ctx, const float *d_scores, int n_ref, int n_cur, float min_score, int *d_idx);
int LaunchCrossCheck(ftk_context *ctx, int *d_idx_fwd, int n_ref, const int *d_idx_bwd, int n_cur);
// match_cosine_tc.cu: tcgen05 GEMM + exact re-rank; FTK_ERR_UNSUPPORTED for dim > 256
int LaunchCosineForceTensor(ftk_context *ctx, const float *d_ref, int n_ref, const float *d_cur, int n_cur, int dim, float max_dist, int *d_idx, bool fill_unmatched);
int LaunchCosineNearby(ftk_context *ctx, const float *d_ref, int n_ref, const float *d_cur, int n_cur, int dim, const float2 *d_pred,
                       const float2 *d_pos, int max_drow, int max_dcol, float max_dist, int *d_idx, bool fill_unmatched);

}  // namespace ftk

#endif
