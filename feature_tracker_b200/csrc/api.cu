// C ABI of libftk_b200.so (declared in include/ftk_c.h): context, device-resident pyramid batches, and the host-side
// marshalling (H2D / launch / D2H) around the kernels in pyramid.cu, klt.cu, klt_basic_fastpath.cu and match.cu.
// There is deliberately no CPU fallback anywhere in this library.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ftk_internal.h"

namespace ftk {

int SetError(ftk_context *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    return code;
}

int EnsureDevice(ftk_context *ctx, FtkBuffer &buf, size_t bytes) {
    if (bytes <= buf.bytes && buf.ptr) return FTK_OK;
    if (buf.ptr) {
        FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaFree(buf.ptr));
        buf.ptr = nullptr;
        buf.bytes = 0;
    }
    size_t want = bytes < 256 ? 256 : bytes;
    want += want / 4;  // growth slack
    FTK_CUDA_CHECK(ctx, cudaMalloc(&buf.ptr, want));
    buf.bytes = want;
    return FTK_OK;
}

}  // namespace ftk

using ftk::EnsureDevice;
using ftk::SetError;

namespace {

inline int RoundUp(int v, int m) { return (v + m - 1) / m * m; }

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device) {
        cudaGetDevice(&prev);
        if (prev != device) cudaSetDevice(device);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

void FreeBuffer(FtkBuffer &b) {
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.bytes = 0;
}

// Host or device source -> device buffer owned by the context (or the caller's device pointer as is).
template <typename T>
int Stage(ftk_context *ctx, FtkBuffer &buf, const T *src, size_t count, bool on_device, const T **out) {
    if (on_device) {
        *out = src;
        return FTK_OK;
    }
    if (int rc = EnsureDevice(ctx, buf, sizeof(T) * (count ? count : 1))) return rc;
    if (count) FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(buf.ptr, src, sizeof(T) * count, cudaMemcpyHostToDevice, ctx->stream));
    *out = static_cast<const T *>(buf.ptr);
    return FTK_OK;
}

}  // namespace

extern "C" {

int ftk_abi_version(void) { return FTK_ABI_VERSION; }

void ftk_klt_params_default(ftk_klt_params *p) {
    if (!p) return;
    p->variant = FTK_VARIANT_BASIC;
    p->method = FTK_METHOD_FAST;
    p->max_track_points = 500;
    p->max_iteration = 15;
    p->max_tolerance_large_step = 3;
    p->patch_row_half = 6;
    p->patch_col_half = 6;
    p->max_converge_step = 4e-2f;
    p->predict[0] = 1.0f, p->predict[1] = 0.0f, p->predict[2] = 0.0f, p->predict[3] = 1.0f;
    p->consider_patch_luminance = 0;
    p->forward_backward_max_error = 0.0f;
}

void ftk_dense_flow_params_default(ftk_dense_flow_params *p) {
    if (!p) return;
    p->max_iteration = 10;  // dense_optical_flow.h:15-20
    p->half_patch_size = 2;
    p->max_converge_step = 1e-6f;
    p->max_delta_flow_step = 1.0f;
}

void ftk_detector_params_default(ftk_detector_params *p) {
    if (!p) return;
    p->kind = FTK_DETECTOR_HARRIS;
    p->half_patch = 1;
    p->harris_k = 0.04f;
    p->min_response = 40.0f;  // the values test/test_descriptor_matcher_brief.cpp:60-61 sets
    p->min_distance = 20;
}

// xorshift32 (13, 17, 5); coordinate = draw % (2 half + 1) - half in the order (drow_a, dcol_a, drow_b, dcol_b); a pair with
// a == b is redrawn.  oracle/ftk_oracle.c holds the checker's own copy of this definition.
void ftk_brief_pattern_default(int32_t n_bits, int32_t half_patch, uint32_t seed, int8_t *pattern) {
    if (!pattern || half_patch < 0 || half_patch > 127) return;
    uint32_t x = seed ? seed : 0x9E3779B9u;
    const uint32_t span = static_cast<uint32_t>(2 * half_patch + 1);
    for (int32_t k = 0; k < n_bits; ++k) {
        int8_t v[4];
        do {
            for (int j = 0; j < 4; ++j) {
                x ^= x << 13, x ^= x >> 17, x ^= x << 5;
                v[j] = static_cast<int8_t>(static_cast<int32_t>(x % span) - half_patch);
            }
        } while (half_patch > 0 && v[0] == v[2] && v[1] == v[3]);
        for (int j = 0; j < 4; ++j) pattern[4 * k + j] = v[j];
    }
}

void ftk_direct_params_default(ftk_direct_params *p) {
    if (!p) return;
    p->max_track_points = 500;  // direct_method_tracker.h:20-28
    p->max_iteration = 15;
    p->patch_row_half = 6;
    p->patch_col_half = 6;
    p->max_converge_step = 1e-6f;
    p->max_converge_residual = 2.0f;
    p->method = FTK_DIRECT_METHOD_DIRECT;
}

int ftk_create(int device, ftk_context **out) {
    if (!out) return FTK_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return FTK_ERR_CUDA;  // no GPU: fail loudly, never fall back
    if (device < 0 || device >= count) return FTK_ERR_INVALID_ARGUMENT;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FTK_ERR_CUDA;
    if (prop.major != 10) return FTK_ERR_UNSUPPORTED;  // the library only carries sm_100a code
    ftk_context *ctx = new ftk_context();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    {
        const char *e = getenv("FTK_DISABLE_FASTPATH");
        ctx->use_fast_paths = !(e && e[0] == '1');
    }
    DeviceGuard guard(device);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return FTK_ERR_CUDA;
    }
    *out = ctx;
    return FTK_OK;
}

void ftk_destroy(ftk_context *ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    FtkBuffer *all[] = {&ctx->d_ref_uv, &ctx->d_cur_uv, &ctx->d_status, &ctx->d_offsets, &ctx->d_ref_img, &ctx->d_cur_img, &ctx->d_feat_pair,
                        &ctx->d_chunk_offsets, &ctx->d_chunk_curmap, &ctx->d_back_uv, &ctx->d_back_status, &ctx->d_small, &ctx->d_dm_K, &ctx->d_dm_points, &ctx->d_dm_q, &ctx->d_dm_p, &ctx->d_flow, &ctx->d_det_response, &ctx->d_det_state, &ctx->d_det_cand, &ctx->d_det_keys, &ctx->d_det_tmp, &ctx->d_det_out, &ctx->d_det_pattern, &ctx->d_desc_ref, &ctx->d_desc_cur, &ctx->d_idx, &ctx->d_pred_uv, &ctx->d_pos_cur, &ctx->d_work0, &ctx->d_work1,
                        &ctx->d_work2, &ctx->d_work3, &ctx->d_cos_counters, &ctx->d_nearby_state};
    for (FtkBuffer *b : all) FreeBuffer(*b);
    for (int b = 0; b < ftk_context::kStageBuffers; ++b) {
        if (ctx->stage_pyr[b]) {
            if (ctx->stage_pyr[b]->storage) cudaFree(ctx->stage_pyr[b]->storage);
            delete ctx->stage_pyr[b];
        }
        if (ctx->ev_copied[b]) cudaEventDestroy(ctx->ev_copied[b]);
        if (ctx->ev_computed[b]) cudaEventDestroy(ctx->ev_computed[b]);
        if (ctx->chunk_stream[b]) cudaStreamDestroy(ctx->chunk_stream[b]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int q = 0; q < 2; ++q)
        if (ctx->ev_prof[q]) cudaEventDestroy(ctx->ev_prof[q]);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *ftk_last_error(const ftk_context *ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int ftk_synchronize(ftk_context *ctx) {
    if (!ctx) return FTK_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return FTK_OK;
}

void *ftk_stream(ftk_context *ctx) { return ctx ? static_cast<void *>(ctx->stream) : nullptr; }

int ftk_alloc_pinned(size_t bytes, void **out) {
    if (!out) return FTK_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? FTK_OK : FTK_ERR_CUDA;
}

void ftk_free_pinned(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

int ftk_set_profiling(ftk_context *ctx, int enabled) {
    if (!ctx) return FTK_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    if (enabled && !ctx->ev_prof[0]) {
        FTK_CUDA_CHECK(ctx, cudaEventCreate(&ctx->ev_prof[0]));
        FTK_CUDA_CHECK(ctx, cudaEventCreate(&ctx->ev_prof[1]));
    }
    ctx->profiling = enabled != 0;
    ctx->prof_recorded = false;
    return FTK_OK;
}

float ftk_last_kernel_ms(ftk_context *ctx) {
    if (!ctx || !ctx->prof_recorded) return -1.0f;
    DeviceGuard guard(ctx->device);
    float ms = -1.0f;
    if (cudaEventSynchronize(ctx->ev_prof[1]) != cudaSuccess) return -1.0f;
    if (cudaEventElapsedTime(&ms, ctx->ev_prof[0], ctx->ev_prof[1]) != cudaSuccess) return -1.0f;
    return ms;
}

uint64_t ftk_kernel_launches(const ftk_context *ctx) { return ctx ? ctx->launches : 0; }

// ---- pyramids -------------------------------------------------------------------------------------------------

int ftk_pyramid_create(ftk_context *ctx, int32_t rows, int32_t cols, int32_t levels, int32_t n_images, ftk_pyramid **out) {
    if (!ctx || !out) return FTK_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (rows <= 0 || cols <= 0 || n_images <= 0) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "pyramid needs positive rows/cols/n_images");
    if (levels < 1 || levels > ftk::kMaxLevels) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "levels must be in [1, %d]", ftk::kMaxLevels);
    DeviceGuard guard(ctx->device);
    ftk_pyramid *pyr = new ftk_pyramid();
    pyr->device = ctx->device;
    ftk::PyramidView &v = pyr->view;
    memset(&v, 0, sizeof(v));
    v.levels = levels;
    v.n_images = n_images;
    size_t total = 0;
    size_t offset[ftk::kMaxLevels];
    for (int l = 0; l < levels; ++l) {
        v.rows[l] = rows >> l;
        v.cols[l] = cols >> l;
        const int r = v.rows[l] > 0 ? v.rows[l] : 1, c = v.cols[l] > 0 ? v.cols[l] : 1;
        v.pitch[l] = RoundUp(c, 16);
        // one extra row + 16 bytes of slack per plane: the sampler's weight-0 "+1" neighbours stay inside it
        v.image_stride[l] = static_cast<long long>(v.pitch[l]) * (r + 1) + 16;
        offset[l] = total;
        total += static_cast<size_t>(v.image_stride[l]) * n_images;
        total = (total + 255) / 256 * 256;
    }
    total += 256;
    if (cudaMalloc(&pyr->storage, total) != cudaSuccess) {
        delete pyr;
        return SetError(ctx, FTK_ERR_CUDA, "cudaMalloc of %zu bytes for the pyramid batch failed: %s", total, cudaGetErrorString(cudaGetLastError()));
    }
    pyr->storage_bytes = total;
    cudaMemsetAsync(pyr->storage, 0, total, ctx->stream);
    for (int l = 0; l < levels; ++l) v.base[l] = pyr->storage + offset[l];
    *out = pyr;
    return FTK_OK;
}

void ftk_pyramid_destroy(ftk_context *ctx, ftk_pyramid *pyr) {
    if (!pyr) return;
    DeviceGuard guard(pyr->device);
    if (ctx) cudaStreamSynchronize(ctx->stream);
    if (pyr->storage) cudaFree(pyr->storage);
    delete pyr;
}

int32_t ftk_pyramid_levels(const ftk_pyramid *pyr) { return pyr ? pyr->view.levels : 0; }
int32_t ftk_pyramid_images(const ftk_pyramid *pyr) { return pyr ? pyr->view.n_images : 0; }

int ftk_pyramid_set_images(ftk_context *ctx, ftk_pyramid *pyr, int32_t first, int32_t count, const uint8_t *images, uint32_t flags) {
    if (!ctx || !pyr || !images) return FTK_ERR_INVALID_ARGUMENT;
    const ftk::PyramidView &v = pyr->view;
    if (first < 0 || count <= 0 || first + count > v.n_images) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "image range out of bounds");
    DeviceGuard guard(ctx->device);
    const cudaMemcpyKind kind = (flags & FTK_FLAG_DEVICE_POINTERS) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    uint8_t *dst = const_cast<uint8_t *>(v.base[0]) + first * v.image_stride[0];
    const size_t plane = static_cast<size_t>(v.rows[0]) * v.cols[0];
    if (v.pitch[0] == v.cols[0]) {
        // Rows are already pitch-sized: one 2-D copy whose "rows" are whole images (src stride = plane, dst stride = image_stride).
        FTK_CUDA_CHECK(ctx, cudaMemcpy2DAsync(dst, v.image_stride[0], images, plane, plane, count, kind, ctx->stream));
    } else {
        // One strided 2-D copy per image (rows of `cols` bytes into rows of `pitch` bytes).
        for (int i = 0; i < count; ++i) {
            FTK_CUDA_CHECK(ctx, cudaMemcpy2DAsync(dst + i * v.image_stride[0], v.pitch[0], images + i * plane, v.cols[0], v.cols[0], v.rows[0], kind,
                                                  ctx->stream));
        }
    }
    return FTK_OK;
}

int ftk_pyramid_build(ftk_context *ctx, ftk_pyramid *pyr, int32_t first, int32_t count) {
    if (!ctx || !pyr) return FTK_ERR_INVALID_ARGUMENT;
    if (first < 0 || count <= 0 || first + count > pyr->view.n_images) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "image range out of bounds");
    DeviceGuard guard(ctx->device);
    for (int done = 0; done < count;) {
        const int n = (count - done) < 32768 ? (count - done) : 32768;  // gridDim.z limit
        if (int rc = ftk::LaunchPyramidBuild(ctx, pyr, first + done, n)) return rc;
        done += n;
    }
    return FTK_OK;
}

int ftk_pyramid_set_level(ftk_context *ctx, ftk_pyramid *pyr, int32_t image, int32_t level, const uint8_t *data) {
    if (!ctx || !pyr || !data) return FTK_ERR_INVALID_ARGUMENT;
    const ftk::PyramidView &v = pyr->view;
    if (image < 0 || image >= v.n_images || level < 0 || level >= v.levels) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "image/level out of range");
    if (v.rows[level] == 0 || v.cols[level] == 0) return FTK_OK;
    DeviceGuard guard(ctx->device);
    uint8_t *dst = const_cast<uint8_t *>(v.base[level]) + image * v.image_stride[level];
    FTK_CUDA_CHECK(ctx, cudaMemcpy2DAsync(dst, v.pitch[level], data, v.cols[level], v.cols[level], v.rows[level], cudaMemcpyHostToDevice, ctx->stream));
    FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return FTK_OK;
}

int ftk_pyramid_get_level(ftk_context *ctx, const ftk_pyramid *pyr, int32_t image, int32_t level, uint8_t *data) {
    if (!ctx || !pyr || !data) return FTK_ERR_INVALID_ARGUMENT;
    const ftk::PyramidView &v = pyr->view;
    if (image < 0 || image >= v.n_images || level < 0 || level >= v.levels) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "image/level out of range");
    if (v.rows[level] == 0 || v.cols[level] == 0) return FTK_OK;
    DeviceGuard guard(ctx->device);
    const uint8_t *src = v.base[level] + image * v.image_stride[level];
    FTK_CUDA_CHECK(ctx, cudaMemcpy2DAsync(data, v.cols[level], src, v.pitch[level], v.cols[level], v.rows[level], cudaMemcpyDeviceToHost, ctx->stream));
    FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return FTK_OK;
}

// ---- KLT ------------------------------------------------------------------------------------------------------

int ftk_klt_track(ftk_context *ctx, const ftk_klt_params *params, const ftk_pyramid *ref, const ftk_pyramid *cur, int32_t n_pairs,
                  const int32_t *ref_image, const int32_t *cur_image, const int32_t *feat_offsets, const float *ref_uv, float *cur_uv, uint8_t *status,
                  uint32_t flags) {
    if (!ctx || !params || !ref || !cur || !feat_offsets || !ref_uv || !cur_uv || !status) return FTK_ERR_INVALID_ARGUMENT;
    if (n_pairs <= 0) return SetError(ctx, FTK_ERR_EMPTY_INPUT, "no frame pairs");
    // optical_flow.cpp:9: RETURN_FALSE_IF(cur_pyramid.level() != ref_pyramid.level())
    if (ref->view.levels != cur->view.levels) return SetError(ctx, FTK_ERR_LEVEL_MISMATCH, "ref has %d levels, cur has %d", ref->view.levels, cur->view.levels);
    if (ref->view.rows[0] != cur->view.rows[0] || ref->view.cols[0] != cur->view.cols[0])
        return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "ref and cur pyramids differ in image size");
    if (params->variant < 0 || params->variant > 2) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "unknown tracker variant %d", params->variant);
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;

    // Feature offsets are always needed on the host too (sizes, validation).
    std::vector<int32_t> h_offsets(n_pairs + 1);
    if (on_device) {
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(h_offsets.data(), feat_offsets, sizeof(int32_t) * (n_pairs + 1), cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        memcpy(h_offsets.data(), feat_offsets, sizeof(int32_t) * (n_pairs + 1));
    }
    if (h_offsets[0] != 0) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "feat_offsets[0] must be 0");
    for (int p = 0; p < n_pairs; ++p)
        if (h_offsets[p + 1] < h_offsets[p]) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "feat_offsets must be non-decreasing");
    const int n_features = h_offsets[n_pairs];
    // optical_flow.cpp:8: RETURN_FALSE_IF(ref_pixel_uv.empty())
    if (n_features == 0) return SetError(ctx, FTK_ERR_EMPTY_INPUT, "no features");
    {
        // pair -> image maps index device memory inside the kernels: validated on the host for device callers too (two small copies
        // next to the feat_offsets one above)
        std::vector<int32_t> h_ri, h_ci;
        if (on_device) {
            if (ref_image) {
                h_ri.resize(n_pairs);
                FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(h_ri.data(), ref_image, sizeof(int32_t) * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
            }
            if (cur_image) {
                h_ci.resize(n_pairs);
                FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(h_ci.data(), cur_image, sizeof(int32_t) * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
            }
            if (ref_image || cur_image) FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        }
        const int32_t *ri_map = on_device ? (ref_image ? h_ri.data() : nullptr) : ref_image;
        const int32_t *ci_map = on_device ? (cur_image ? h_ci.data() : nullptr) : cur_image;
        for (int p = 0; p < n_pairs; ++p) {
            const int ri = ri_map ? ri_map[p] : p, ci = ci_map ? ci_map[p] : p;
            if (ri < 0 || ri >= ref->view.n_images || ci < 0 || ci >= cur->view.n_images)
                return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "pair %d references image %d/%d outside the pyramid batches", p, ri, ci);
        }
    }

    ftk::KltLaunch a{};
    a.p = *params;
    a.ref = ref->view;
    a.cur = cur->view;
    a.n_pairs = n_pairs;
    a.n_features = n_features;
    a.has_prediction = (flags & FTK_FLAG_NO_PREDICTION) ? 0 : 1;
    a.has_status = (flags & FTK_FLAG_NO_STATUS) ? 0 : 1;
    a.single_level = (flags & FTK_FLAG_SINGLE_LEVEL) ? 1 : 0;

    // Small host-pointer calls (the reference's own use: one frame pair, a few hundred features -- BASELINE configs[0]) are bound by
    // the number of copies and launches, not by bytes: every array travels in ONE pinned block (one H2D, one D2H).
    // Block layout: [cur_uv | status (padded to 16) | ref_uv | feat_offsets | ref_image | cur_image]; results = the first two.
    const size_t sz_uv = sizeof(float2) * static_cast<size_t>(n_features), sz_st = (static_cast<size_t>(n_features) + 15) / 16 * 16;
    const size_t sz_off = (sizeof(int32_t) * (static_cast<size_t>(n_pairs) + 1) + 15) / 16 * 16, sz_map = (sizeof(int32_t) * static_cast<size_t>(n_pairs) + 15) / 16 * 16;
    const size_t small_bytes = 2 * sz_uv + sz_st + sz_off + 2 * sz_map;
    if (!on_device && small_bytes <= 256 * 1024) {
        if (int rc = EnsureDevice(ctx, ctx->d_small, small_bytes)) return rc;
        if (ctx->h_small_bytes < small_bytes) {
            if (ctx->h_small) cudaFreeHost(ctx->h_small);
            ctx->h_small = nullptr, ctx->h_small_bytes = 0;
            FTK_CUDA_CHECK(ctx, cudaHostAlloc(&ctx->h_small, 256 * 1024, cudaHostAllocDefault));
            ctx->h_small_bytes = 256 * 1024;
        }
        uint8_t *h = static_cast<uint8_t *>(ctx->h_small), *d = static_cast<uint8_t *>(ctx->d_small.ptr);
        const size_t o_st = sz_uv, o_ref = o_st + sz_st, o_off = o_ref + sz_uv, o_ri = o_off + sz_off, o_ci = o_ri + sz_map;
        if (a.has_prediction) memcpy(h, cur_uv, sz_uv);
        if (a.has_status) memcpy(h + o_st, status, n_features);
        memcpy(h + o_ref, ref_uv, sz_uv);
        memcpy(h + o_off, feat_offsets, sizeof(int32_t) * (n_pairs + 1));
        if (ref_image) memcpy(h + o_ri, ref_image, sizeof(int32_t) * n_pairs);
        if (cur_image) memcpy(h + o_ci, cur_image, sizeof(int32_t) * n_pairs);
        // inputs only: when neither a prediction nor a status comes in, the upload starts at ref_uv
        const size_t up_from = (a.has_prediction || a.has_status) ? 0 : o_ref;
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(d + up_from, h + up_from, small_bytes - up_from, cudaMemcpyHostToDevice, ctx->stream));
        a.cur_uv = reinterpret_cast<float2 *>(d);
        a.status = d + o_st;
        a.ref_uv = reinterpret_cast<const float2 *>(d + o_ref);
        a.feat_offsets = reinterpret_cast<const int *>(d + o_off);
        a.ref_image = ref_image ? reinterpret_cast<const int *>(d + o_ri) : nullptr;
        a.cur_image = cur_image ? reinterpret_cast<const int *>(d + o_ci) : nullptr;
        if (n_pairs == 1) {
            a.feat_pair = nullptr;  // every feature belongs to pair 0
        } else {
            if (int rc = EnsureDevice(ctx, ctx->d_feat_pair, sizeof(int) * n_features)) return rc;
            int *d_feat_pair = static_cast<int *>(ctx->d_feat_pair.ptr);
            if (int rc = ftk::LaunchFeaturePairs(ctx, a.feat_offsets, n_pairs, n_features, d_feat_pair)) return rc;
            a.feat_pair = d_feat_pair;
        }
        if (int rc = ftk::LaunchKltTrackChecked(ctx, a)) return rc;
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(h, d, o_st + n_features, cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(cur_uv, h, sz_uv);
        memcpy(status, h + o_st, n_features);
        return FTK_OK;
    }

    const float2 *d_ref_uv = nullptr;
    if (int rc = Stage(ctx, ctx->d_ref_uv, reinterpret_cast<const float2 *>(ref_uv), n_features, on_device, &d_ref_uv)) return rc;
    a.ref_uv = d_ref_uv;
    if (on_device) {
        a.cur_uv = reinterpret_cast<float2 *>(cur_uv);
        a.status = status;
        a.feat_offsets = feat_offsets;
        a.ref_image = ref_image;
        a.cur_image = cur_image;
    } else {
        if (int rc = EnsureDevice(ctx, ctx->d_cur_uv, sizeof(float2) * n_features)) return rc;
        if (int rc = EnsureDevice(ctx, ctx->d_status, n_features)) return rc;
        a.cur_uv = static_cast<float2 *>(ctx->d_cur_uv.ptr);
        a.status = static_cast<uint8_t *>(ctx->d_status.ptr);
        if (a.has_prediction) FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(a.cur_uv, cur_uv, sizeof(float2) * n_features, cudaMemcpyHostToDevice, ctx->stream));
        if (a.has_status) FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(a.status, status, n_features, cudaMemcpyHostToDevice, ctx->stream));
        const int32_t *d = nullptr;
        if (int rc = Stage(ctx, ctx->d_offsets, feat_offsets, n_pairs + 1, false, &d)) return rc;
        a.feat_offsets = d;
        if (ref_image) {
            if (int rc = Stage(ctx, ctx->d_ref_img, ref_image, n_pairs, false, &d)) return rc;
            a.ref_image = d;
        }
        if (cur_image) {
            if (int rc = Stage(ctx, ctx->d_cur_img, cur_image, n_pairs, false, &d)) return rc;
            a.cur_image = d;
        }
    }
    if (int rc = EnsureDevice(ctx, ctx->d_feat_pair, sizeof(int) * n_features)) return rc;
    int *d_feat_pair = static_cast<int *>(ctx->d_feat_pair.ptr);
    if (int rc = ftk::LaunchFeaturePairs(ctx, a.feat_offsets, n_pairs, n_features, d_feat_pair)) return rc;
    a.feat_pair = d_feat_pair;

    if (int rc = ftk::LaunchKltTrackChecked(ctx, a)) return rc;

    if (!on_device) {
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(cur_uv, a.cur_uv, sizeof(float2) * n_features, cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(status, a.status, n_features, cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return FTK_OK;
}

// The chunked pipeline behind ftk_track_image_pairs (cur_images != nullptr: pair p = ref_images[p] -> cur_images[p]) and
// ftk_track_image_sequence (cur_images == nullptr: ref_images holds n_pairs + 1 frames, pair p = frame p -> frame p + 1).
// n_trackers parameter sets run on every chunk while its images are resident: tracker k reads / writes cur_uv + k * 2 * n_features and
// status + k * n_features, exactly as n_trackers separate calls would, but the images cross PCIe once.
//
// Streams: one copy stream carries every host-to-device byte (a chunk's features, then its images); kStageBuffers staging pyramids
// rotate, each with its own compute stream, so the tail of chunk c's kernels overlaps the head of chunk c + 1's (different streams)
// and chunk c's results travel back (device-to-host engine) while later chunks are uploaded and computed.  Nothing is copied up
// front and nothing is left for the end but the last chunk's results.
static int TrackImagesPipelinedBody(ftk_context *ctx, const ftk_klt_params *params, int32_t n_trackers, int32_t rows, int32_t cols, int32_t levels,
                                    int32_t n_pairs, const uint8_t *ref_images, const uint8_t *cur_images, const int32_t *feat_offsets, const float *ref_uv,
                                    float *cur_uv, uint8_t *status, uint32_t flags, bool any_fb) {
    const bool sequence = cur_images == nullptr;
    const int n_features = feat_offsets[n_pairs];
    const size_t n_all = static_cast<size_t>(n_trackers) * n_features;
    constexpr int kB = ftk_context::kStageBuffers;

    // chunking: ~24 chunks per call (small enough that filling / draining the pipeline costs a few percent, large enough that a
    // chunk's kernels still fill the GPU for several waves)
    int chunk_pairs = (n_pairs + 23) / 24;  // (16 / 24 / 32 / 48 chunks measured within 0.6 % of each other on BASELINE configs[1])
    if (chunk_pairs < 8) chunk_pairs = n_pairs < 8 ? n_pairs : 8;
    const int n_chunks = (n_pairs + chunk_pairs - 1) / chunk_pairs;
    if (!ctx->copy_stream) {
        FTK_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < kB; ++b) {
            FTK_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->chunk_stream[b], cudaStreamNonBlocking));
            FTK_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming));
            FTK_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ev_computed[b], cudaEventDisableTiming));
        }
    }
    FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // earlier asynchronous calls on the context may still use the scratch buffers
    if (ctx->stage_rows != rows || ctx->stage_cols != cols || ctx->stage_levels != levels || ctx->stage_pairs < chunk_pairs) {
        for (int b = 0; b < kB; ++b) {
            if (ctx->stage_pyr[b]) ftk_pyramid_destroy(ctx, ctx->stage_pyr[b]);
            ctx->stage_pyr[b] = nullptr;
        }
        ctx->stage_rows = ctx->stage_cols = ctx->stage_levels = ctx->stage_pairs = 0;
        for (int b = 0; b < kB; ++b) {
            if (int rc = ftk_pyramid_create(ctx, rows, cols, levels, 2 * chunk_pairs, &ctx->stage_pyr[b])) {
                for (int q = 0; q < kB; ++q) {  // all or nothing: a later call must not find half of the staging set
                    if (ctx->stage_pyr[q]) ftk_pyramid_destroy(ctx, ctx->stage_pyr[q]);
                    ctx->stage_pyr[q] = nullptr;
                }
                return rc;
            }
        }
        ctx->stage_rows = rows, ctx->stage_cols = cols, ctx->stage_levels = levels, ctx->stage_pairs = chunk_pairs;
        FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // the pyramids' memset
    }
    const int stage_pairs = ctx->stage_pairs;

    if (int rc = EnsureDevice(ctx, ctx->d_ref_uv, sizeof(float2) * n_features)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_cur_uv, sizeof(float2) * n_all)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_status, n_all)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_feat_pair, sizeof(int) * n_features)) return rc;
    if (any_fb) {  // backward-pass scratch for the whole batch, so no chunk reallocates it
        if (int rc = EnsureDevice(ctx, ctx->d_back_uv, sizeof(float2) * n_all)) return rc;
        if (int rc = EnsureDevice(ctx, ctx->d_back_status, n_all)) return rc;
    }
    // chunk-local offset tables and the "cur image = stage_pairs + p" map (small; first thing on the copy stream)
    std::vector<int32_t> h_offsets;
    h_offsets.reserve(n_pairs + n_chunks);
    std::vector<int> chunk_table_start(n_chunks);
    for (int c = 0; c < n_chunks; ++c) {
        const int p_lo = c * chunk_pairs, p_hi = p_lo + chunk_pairs < n_pairs ? p_lo + chunk_pairs : n_pairs;
        chunk_table_start[c] = static_cast<int>(h_offsets.size());
        for (int p = p_lo; p <= p_hi; ++p) h_offsets.push_back(feat_offsets[p] - feat_offsets[p_lo]);
    }
    std::vector<int32_t> h_curmap(stage_pairs);
    for (int p = 0; p < stage_pairs; ++p) h_curmap[p] = sequence ? p + 1 : stage_pairs + p;
    if (int rc = EnsureDevice(ctx, ctx->d_chunk_offsets, sizeof(int32_t) * h_offsets.size())) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_chunk_curmap, sizeof(int32_t) * h_curmap.size())) return rc;
    const bool has_prediction = !(flags & FTK_FLAG_NO_PREDICTION), has_status = !(flags & FTK_FLAG_NO_STATUS);
    cudaStream_t cs = ctx->copy_stream;
    FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_chunk_offsets.ptr, h_offsets.data(), sizeof(int32_t) * h_offsets.size(), cudaMemcpyHostToDevice, cs));
    FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_chunk_curmap.ptr, h_curmap.data(), sizeof(int32_t) * h_curmap.size(), cudaMemcpyHostToDevice, cs));

    const size_t plane = static_cast<size_t>(rows) * cols;
    float2 *d_ref_uv = static_cast<float2 *>(ctx->d_ref_uv.ptr), *d_cur_uv = static_cast<float2 *>(ctx->d_cur_uv.ptr);
    uint8_t *d_status = static_cast<uint8_t *>(ctx->d_status.ptr);
    for (int c = 0; c < n_chunks; ++c) {
        const int b = c % kB;
        const int p_lo = c * chunk_pairs, p_hi = p_lo + chunk_pairs < n_pairs ? p_lo + chunk_pairs : n_pairs;
        const int np = p_hi - p_lo;
        const int f_lo = feat_offsets[p_lo], f_hi = feat_offsets[p_hi], nf = f_hi - f_lo;
        ftk_pyramid *pyr = ctx->stage_pyr[b];
        const ftk::PyramidView &v = pyr->view;
        // ---- copy stream: this chunk's features and images into staging buffer b (after its previous user finished computing)
        if (c >= kB) FTK_CUDA_CHECK(ctx, cudaStreamWaitEvent(cs, ctx->ev_computed[b], 0));
        if (nf > 0) {
            FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(d_ref_uv + f_lo, ref_uv + 2 * static_cast<size_t>(f_lo), sizeof(float2) * nf, cudaMemcpyHostToDevice, cs));
            for (int k = 0; k < n_trackers; ++k) {
                const size_t base = static_cast<size_t>(k) * n_features + f_lo;
                if (has_prediction) FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(d_cur_uv + base, cur_uv + 2 * base, sizeof(float2) * nf, cudaMemcpyHostToDevice, cs));
                if (has_status) FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(d_status + base, status + base, nf, cudaMemcpyHostToDevice, cs));
            }
        }
        uint8_t *dst_ref = const_cast<uint8_t *>(v.base[0]);
        uint8_t *dst_cur = dst_ref + static_cast<size_t>(stage_pairs) * v.image_stride[0];
        const int n_first = sequence ? np + 1 : np;  // images copied into staging slots [0, n_first)
        if (v.pitch[0] == cols) {
            FTK_CUDA_CHECK(ctx, cudaMemcpy2DAsync(dst_ref, v.image_stride[0], ref_images + p_lo * plane, plane, plane, n_first, cudaMemcpyHostToDevice, cs));
            if (!sequence)
                FTK_CUDA_CHECK(ctx, cudaMemcpy2DAsync(dst_cur, v.image_stride[0], cur_images + p_lo * plane, plane, plane, np, cudaMemcpyHostToDevice, cs));
        } else {
            for (int i = 0; i < n_first; ++i)
                FTK_CUDA_CHECK(ctx, cudaMemcpy2DAsync(dst_ref + i * v.image_stride[0], v.pitch[0], ref_images + (p_lo + i) * plane, cols, cols, rows,
                                                      cudaMemcpyHostToDevice, cs));
            for (int i = 0; i < np && !sequence; ++i)
                FTK_CUDA_CHECK(ctx, cudaMemcpy2DAsync(dst_cur + i * v.image_stride[0], v.pitch[0], cur_images + (p_lo + i) * plane, cols, cols, rows,
                                                      cudaMemcpyHostToDevice, cs));
        }
        FTK_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_copied[b], cs));
        // ---- compute stream of buffer b: pyramids of the chunk's images, its features, its results back to the host
        cudaStream_t ks = ctx->chunk_stream[b];
        ctx->stream = ks;  // every Launch* helper issues on ctx->stream (restored by the caller, also on error)
        FTK_CUDA_CHECK(ctx, cudaStreamWaitEvent(ks, ctx->ev_copied[b], 0));
        if (int rc = ftk::LaunchPyramidBuild(ctx, pyr, 0, n_first)) return rc;
        if (!sequence)
            if (int rc = ftk::LaunchPyramidBuild(ctx, pyr, stage_pairs, np)) return rc;
        if (nf > 0) {
            int *d_feat_pair = static_cast<int *>(ctx->d_feat_pair.ptr) + f_lo;
            const int *chunk_offsets = static_cast<const int *>(ctx->d_chunk_offsets.ptr) + chunk_table_start[c];
            if (int rc = ftk::LaunchFeaturePairs(ctx, chunk_offsets, np, nf, d_feat_pair)) return rc;
            for (int k = 0; k < n_trackers; ++k) {
                const size_t base = static_cast<size_t>(k) * n_features + f_lo;
                ftk::KltLaunch a{};
                a.p = params[k];
                a.ref = v;
                a.cur = v;
                a.n_pairs = np;
                a.n_features = nf;
                a.has_prediction = has_prediction ? 1 : 0;
                a.has_status = has_status ? 1 : 0;
                a.single_level = 0;
                a.ref_uv = d_ref_uv + f_lo;
                a.cur_uv = d_cur_uv + base;
                a.status = d_status + base;
                a.feat_offsets = chunk_offsets;
                a.ref_image = nullptr;  // image p of the staging batch
                a.cur_image = static_cast<const int *>(ctx->d_chunk_curmap.ptr);
                a.feat_pair = d_feat_pair;
                if (int rc = ftk::LaunchKltTrackChecked(ctx, a, base)) return rc;
                FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(cur_uv + 2 * base, d_cur_uv + base, sizeof(float2) * nf, cudaMemcpyDeviceToHost, ks));
                FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(status + base, d_status + base, nf, cudaMemcpyDeviceToHost, ks));
            }
        }
        FTK_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_computed[b], ks));
    }
    return FTK_OK;
}

static int TrackImagesPipelined(ftk_context *ctx, const ftk_klt_params *params, int32_t n_trackers, int32_t rows, int32_t cols, int32_t levels,
                                int32_t n_pairs, const uint8_t *ref_images, const uint8_t *cur_images, const int32_t *feat_offsets, const float *ref_uv,
                                float *cur_uv, uint8_t *status, uint32_t flags) {
    if (!ctx || !params || n_trackers < 1 || !ref_images || !feat_offsets || !ref_uv || !cur_uv || !status) return FTK_ERR_INVALID_ARGUMENT;
    if (n_pairs <= 0) return SetError(ctx, FTK_ERR_EMPTY_INPUT, "no frame pairs");
    if (flags & (FTK_FLAG_DEVICE_POINTERS | FTK_FLAG_SINGLE_LEVEL)) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "host images, multi-level only");
    bool any_fb = false;
    for (int k = 0; k < n_trackers; ++k) {
        if (params[k].variant < 0 || params[k].variant > 2) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "unknown tracker variant %d", params[k].variant);
        any_fb = any_fb || params[k].forward_backward_max_error > 0.0f;
    }
    if (feat_offsets[0] != 0) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "feat_offsets[0] must be 0");
    for (int p = 0; p < n_pairs; ++p)
        if (feat_offsets[p + 1] < feat_offsets[p]) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "feat_offsets must be non-decreasing");
    if (feat_offsets[n_pairs] == 0) return SetError(ctx, FTK_ERR_EMPTY_INPUT, "no features");  // optical_flow.cpp:8
    DeviceGuard guard(ctx->device);
    cudaStream_t user_stream = ctx->stream;
    const int rc = TrackImagesPipelinedBody(ctx, params, n_trackers, rows, cols, levels, n_pairs, ref_images, cur_images, feat_offsets, ref_uv, cur_uv, status,
                                            flags, any_fb);
    ctx->stream = user_stream;
    // Success or not, nothing of this call may stay in flight: the copies read and write the caller's host arrays.
    cudaError_t err = cudaSuccess;
    if (ctx->copy_stream) {
        cudaError_t e = cudaStreamSynchronize(ctx->copy_stream);
        if (e != cudaSuccess) err = e;
        for (int b = 0; b < ftk_context::kStageBuffers; ++b) {
            if (!ctx->chunk_stream[b]) continue;
            e = cudaStreamSynchronize(ctx->chunk_stream[b]);
            if (e != cudaSuccess) err = e;
        }
    }
    if (rc != FTK_OK) return rc;
    if (err != cudaSuccess) return SetError(ctx, FTK_ERR_CUDA, "pipelined tracking failed: %s", cudaGetErrorString(err));
    return FTK_OK;
}

int ftk_track_image_pairs(ftk_context *ctx, const ftk_klt_params *params, int32_t rows, int32_t cols, int32_t levels, int32_t n_pairs,
                          const uint8_t *ref_images, const uint8_t *cur_images, const int32_t *feat_offsets, const float *ref_uv, float *cur_uv,
                          uint8_t *status, uint32_t flags) {
    if (!cur_images) return FTK_ERR_INVALID_ARGUMENT;
    return TrackImagesPipelined(ctx, params, 1, rows, cols, levels, n_pairs, ref_images, cur_images, feat_offsets, ref_uv, cur_uv, status, flags);
}

int ftk_track_image_pairs_multi(ftk_context *ctx, const ftk_klt_params *params, int32_t n_trackers, int32_t rows, int32_t cols, int32_t levels,
                                int32_t n_pairs, const uint8_t *ref_images, const uint8_t *cur_images, const int32_t *feat_offsets, const float *ref_uv,
                                float *cur_uv, uint8_t *status, uint32_t flags) {
    if (!cur_images) return FTK_ERR_INVALID_ARGUMENT;
    return TrackImagesPipelined(ctx, params, n_trackers, rows, cols, levels, n_pairs, ref_images, cur_images, feat_offsets, ref_uv, cur_uv, status, flags);
}

int ftk_track_image_sequence(ftk_context *ctx, const ftk_klt_params *params, int32_t rows, int32_t cols, int32_t levels, int32_t n_frames,
                             const uint8_t *frames, const int32_t *feat_offsets, const float *ref_uv, float *cur_uv, uint8_t *status, uint32_t flags) {
    if (ctx && n_frames < 2) return SetError(ctx, FTK_ERR_EMPTY_INPUT, "a sequence needs at least two frames");
    return TrackImagesPipelined(ctx, params, 1, rows, cols, levels, n_frames - 1, frames, nullptr, feat_offsets, ref_uv, cur_uv, status, flags);
}

// ---- matching -------------------------------------------------------------------------------------------------

namespace {

// Common index-vector handling (descriptor_matcher.h:60-62, 98-100): a missing index vector starts at -1.
// defer_fill: with FTK_FLAG_NO_INDEX_INPUT the launcher writes the -1 of unmatched rows itself (no memset here)
int PrepareIndex(ftk_context *ctx, int32_t *idx, int n_ref, uint32_t flags, int **d_idx, bool defer_fill = false) {
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    if (on_device) {
        *d_idx = idx;
    } else {
        if (int rc = EnsureDevice(ctx, ctx->d_idx, sizeof(int) * (n_ref ? n_ref : 1))) return rc;
        *d_idx = static_cast<int *>(ctx->d_idx.ptr);
    }
    if (n_ref == 0) return FTK_OK;
    if (flags & FTK_FLAG_NO_INDEX_INPUT) {
        if (!defer_fill) FTK_CUDA_CHECK(ctx, cudaMemsetAsync(*d_idx, 0xFF, sizeof(int) * n_ref, ctx->stream));  // -1
    } else if (!on_device) {
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(*d_idx, idx, sizeof(int) * n_ref, cudaMemcpyHostToDevice, ctx->stream));
    }
    return FTK_OK;
}

int FinishIndex(ftk_context *ctx, int32_t *idx, int n_ref, uint32_t flags, const int *d_idx) {
    if (flags & FTK_FLAG_DEVICE_POINTERS) return FTK_OK;
    if (n_ref) FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(idx, d_idx, sizeof(int) * n_ref, cudaMemcpyDeviceToHost, ctx->stream));
    FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return FTK_OK;
}

}  // namespace

int ftk_match_hamming_force(ftk_context *ctx, const uint32_t *ref, int32_t n_ref, const uint32_t *cur, int32_t n_cur, int32_t words, float max_dist,
                            int32_t *idx, uint32_t flags) {
    if (!ctx || !idx || n_ref < 0 || words < 0 || words > 64) return FTK_ERR_INVALID_ARGUMENT;
    if (n_cur <= 0) return SetError(ctx, FTK_ERR_EMPTY_INPUT, "descriptors_cur is empty");  // descriptor_matcher.h:58
    if ((n_ref > 0 && !ref) || !cur) return FTK_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    const uint32_t *d_ref = nullptr, *d_cur = nullptr;
    if (int rc = Stage(ctx, ctx->d_desc_ref, ref, static_cast<size_t>(n_ref) * words, on_device, &d_ref)) return rc;
    if (int rc = Stage(ctx, ctx->d_desc_cur, cur, static_cast<size_t>(n_cur) * words, on_device, &d_cur)) return rc;
    int *d_idx = nullptr;
    if (int rc = PrepareIndex(ctx, idx, n_ref, flags, &d_idx, true)) return rc;
    if (int rc = ftk::LaunchHammingForce(ctx, d_ref, n_ref, d_cur, n_cur, words, max_dist, d_idx, (flags & FTK_FLAG_NO_INDEX_INPUT) != 0)) return rc;
    return FinishIndex(ctx, idx, n_ref, flags, d_idx);
}

int ftk_match_hamming_nearby(ftk_context *ctx, const uint32_t *ref, int32_t n_ref, const uint32_t *cur, int32_t n_cur, int32_t words,
                             const float *pred_uv, const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist, int32_t *idx,
                             uint32_t flags) {
    if (!ctx || !idx || n_ref < 0 || words < 0 || words > 64) return FTK_ERR_INVALID_ARGUMENT;
    if (n_cur <= 0) return SetError(ctx, FTK_ERR_EMPTY_INPUT, "descriptors_cur is empty");  // descriptor_matcher.h:94
    if ((n_ref > 0 && (!ref || !pred_uv)) || !cur || !cur_uv) return FTK_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    const uint32_t *d_ref = nullptr, *d_cur = nullptr;
    const float2 *d_pred = nullptr, *d_pos = nullptr;
    if (int rc = Stage(ctx, ctx->d_desc_ref, ref, static_cast<size_t>(n_ref) * words, on_device, &d_ref)) return rc;
    if (int rc = Stage(ctx, ctx->d_desc_cur, cur, static_cast<size_t>(n_cur) * words, on_device, &d_cur)) return rc;
    if (int rc = Stage(ctx, ctx->d_pred_uv, reinterpret_cast<const float2 *>(pred_uv), n_ref, on_device, &d_pred)) return rc;
    if (int rc = Stage(ctx, ctx->d_pos_cur, reinterpret_cast<const float2 *>(cur_uv), n_cur, on_device, &d_pos)) return rc;
    int *d_idx = nullptr;
    if (int rc = PrepareIndex(ctx, idx, n_ref, flags, &d_idx, true)) return rc;
    if (int rc = ftk::LaunchHammingNearby(ctx, d_ref, n_ref, d_cur, n_cur, words, d_pred, d_pos, max_drow, max_dcol, max_dist, d_idx,
                                          (flags & FTK_FLAG_NO_INDEX_INPUT) != 0))
        return rc;
    return FinishIndex(ctx, idx, n_ref, flags, d_idx);
}

int ftk_match_hamming_pairs(ftk_context *ctx, const uint32_t *ref, const uint32_t *cur, int32_t words, int32_t n_pairs, const int32_t *ref_offsets,
                            const int32_t *cur_offsets, const float *pred_uv, const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist,
                            int32_t *idx, uint32_t flags) {
    if (!ctx || !idx || n_pairs < 0 || words < 0 || words > 64 || !ref_offsets || !cur_offsets || (pred_uv && !cur_uv)) return FTK_ERR_INVALID_ARGUMENT;
    if (n_pairs == 0) return FTK_OK;
    for (int32_t p = 0; p < n_pairs; ++p)
        if (ref_offsets[p + 1] < ref_offsets[p] || cur_offsets[p + 1] < cur_offsets[p] || ref_offsets[0] != 0 || cur_offsets[0] != 0)
            return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "descriptor offsets must start at 0 and not decrease (pair %d)", p);
    const int32_t n_ref = ref_offsets[n_pairs], n_cur = cur_offsets[n_pairs];
    if ((n_ref > 0 && !ref) || (n_cur > 0 && !cur)) return FTK_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    const uint32_t *d_ref = nullptr, *d_cur = nullptr;
    const float2 *d_pred = nullptr, *d_pos = nullptr;
    const int32_t *d_ref_off = nullptr, *d_cur_off = nullptr;
    if (int rc = Stage(ctx, ctx->d_desc_ref, ref, static_cast<size_t>(n_ref) * words, on_device, &d_ref)) return rc;
    if (int rc = Stage(ctx, ctx->d_desc_cur, cur, static_cast<size_t>(n_cur) * words, on_device, &d_cur)) return rc;
    if (pred_uv) {
        if (int rc = Stage(ctx, ctx->d_pred_uv, reinterpret_cast<const float2 *>(pred_uv), n_ref, on_device, &d_pred)) return rc;
        if (int rc = Stage(ctx, ctx->d_pos_cur, reinterpret_cast<const float2 *>(cur_uv), n_cur, on_device, &d_pos)) return rc;
    }
    if (int rc = Stage(ctx, ctx->d_offsets, ref_offsets, static_cast<size_t>(n_pairs) + 1, false, &d_ref_off)) return rc;
    if (int rc = Stage(ctx, ctx->d_chunk_offsets, cur_offsets, static_cast<size_t>(n_pairs) + 1, false, &d_cur_off)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_feat_pair, sizeof(int) * static_cast<size_t>(n_ref ? n_ref : 1))) return rc;
    int *d_ref_pair = static_cast<int *>(ctx->d_feat_pair.ptr);
    if (n_ref > 0)
        if (int rc = ftk::LaunchFeaturePairs(ctx, d_ref_off, n_pairs, n_ref, d_ref_pair)) return rc;
    int *d_idx = nullptr;
    if (int rc = PrepareIndex(ctx, idx, n_ref, flags, &d_idx)) return rc;
    if (int rc = ftk::LaunchHammingPairs(ctx, d_ref, d_cur, words, n_ref, d_ref_pair, d_ref_off, d_cur_off, d_pred, d_pos, max_drow, max_dcol, max_dist, d_idx))
        return rc;
    return FinishIndex(ctx, idx, n_ref, flags, d_idx);
}

int ftk_match_cosine_pairs(ftk_context *ctx, const float *ref, const float *cur, int32_t dim, int32_t n_pairs, const int32_t *ref_offsets,
                           const int32_t *cur_offsets, const float *pred_uv, const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist,
                           int32_t *idx, uint32_t flags) {
    if (!ctx || !idx || n_pairs < 0 || dim <= 0 || dim > 1024 || !ref_offsets || !cur_offsets || (pred_uv && !cur_uv)) return FTK_ERR_INVALID_ARGUMENT;
    if (n_pairs == 0) return FTK_OK;
    for (int32_t p = 0; p < n_pairs; ++p)
        if (ref_offsets[p + 1] < ref_offsets[p] || cur_offsets[p + 1] < cur_offsets[p] || ref_offsets[0] != 0 || cur_offsets[0] != 0)
            return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "descriptor offsets must start at 0 and not decrease (pair %d)", p);
    const int32_t n_ref = ref_offsets[n_pairs], n_cur = cur_offsets[n_pairs];
    if ((n_ref > 0 && !ref) || (n_cur > 0 && !cur)) return FTK_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    const float *d_ref = nullptr, *d_cur = nullptr;
    const float2 *d_pred = nullptr, *d_pos = nullptr;
    const int32_t *d_ref_off = nullptr, *d_cur_off = nullptr;
    if (int rc = Stage(ctx, ctx->d_desc_ref, ref, static_cast<size_t>(n_ref) * dim, on_device, &d_ref)) return rc;
    if (int rc = Stage(ctx, ctx->d_desc_cur, cur, static_cast<size_t>(n_cur) * dim, on_device, &d_cur)) return rc;
    if (pred_uv) {
        if (int rc = Stage(ctx, ctx->d_pred_uv, reinterpret_cast<const float2 *>(pred_uv), n_ref, on_device, &d_pred)) return rc;
        if (int rc = Stage(ctx, ctx->d_pos_cur, reinterpret_cast<const float2 *>(cur_uv), n_cur, on_device, &d_pos)) return rc;
    }
    if (int rc = Stage(ctx, ctx->d_offsets, ref_offsets, static_cast<size_t>(n_pairs) + 1, false, &d_ref_off)) return rc;
    if (int rc = Stage(ctx, ctx->d_chunk_offsets, cur_offsets, static_cast<size_t>(n_pairs) + 1, false, &d_cur_off)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_feat_pair, sizeof(int) * static_cast<size_t>(n_ref ? n_ref : 1))) return rc;
    int *d_ref_pair = static_cast<int *>(ctx->d_feat_pair.ptr);
    if (n_ref > 0)
        if (int rc = ftk::LaunchFeaturePairs(ctx, d_ref_off, n_pairs, n_ref, d_ref_pair)) return rc;
    int *d_idx = nullptr;
    if (int rc = PrepareIndex(ctx, idx, n_ref, flags, &d_idx)) return rc;
    if (n_cur > 0)
        if (int rc = ftk::LaunchCosinePairs(ctx, d_ref, d_cur, dim, n_ref, n_cur, d_ref_pair, d_cur_off, d_pred, d_pos, max_drow, max_dcol, max_dist, d_idx)) return rc;
    return FinishIndex(ctx, idx, n_ref, flags, d_idx);
}

int ftk_direct_method_track(ftk_context *ctx, const ftk_direct_params *params, const ftk_pyramid *ref, const ftk_pyramid *cur, int32_t n_pairs,
                            const int32_t *ref_image, const int32_t *cur_image, const int32_t *feat_offsets, const float *K, const float *p_c_in_ref,
                            const float *ref_uv, float *cur_uv, float *q_rc, float *p_rc, uint8_t *status, uint32_t flags) {
    if (!ctx || !params || !ref || !cur || !feat_offsets || !K || !p_c_in_ref || !ref_uv || !cur_uv || !q_rc || !p_rc || !status)
        return FTK_ERR_INVALID_ARGUMENT;
    if (n_pairs <= 0) return SetError(ctx, FTK_ERR_EMPTY_INPUT, "no frame pairs");
    if (ref->view.levels != cur->view.levels)  // direct_method_tracker.cpp:45
        return SetError(ctx, FTK_ERR_LEVEL_MISMATCH, "ref has %d levels, cur has %d", ref->view.levels, cur->view.levels);
    if (ref->view.rows[0] != cur->view.rows[0] || ref->view.cols[0] != cur->view.cols[0])
        return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "ref and cur pyramids differ in image size");
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    std::vector<int32_t> h_offsets(n_pairs + 1);
    if (on_device) {
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(h_offsets.data(), feat_offsets, sizeof(int32_t) * (n_pairs + 1), cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        memcpy(h_offsets.data(), feat_offsets, sizeof(int32_t) * (n_pairs + 1));
    }
    if (h_offsets[0] != 0) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "feat_offsets[0] must be 0");
    std::vector<int32_t> h_ri, h_ci;  // the pair -> image maps index device memory inside the kernel: validated for device callers too
    if (on_device && ref_image) {
        h_ri.resize(n_pairs);
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(h_ri.data(), ref_image, sizeof(int32_t) * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (on_device && cur_image) {
        h_ci.resize(n_pairs);
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(h_ci.data(), cur_image, sizeof(int32_t) * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (on_device && (ref_image || cur_image)) FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    const int32_t *ri_map = on_device ? (ref_image ? h_ri.data() : nullptr) : ref_image;
    const int32_t *ci_map = on_device ? (cur_image ? h_ci.data() : nullptr) : cur_image;
    for (int p = 0; p < n_pairs; ++p) {
        if (h_offsets[p + 1] < h_offsets[p]) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "feat_offsets must be non-decreasing");
        if (h_offsets[p + 1] == h_offsets[p]) return SetError(ctx, FTK_ERR_EMPTY_INPUT, "pair %d has no features", p);  // :44
        const int ri = ri_map ? ri_map[p] : p, ci = ci_map ? ci_map[p] : p;
        if (ri < 0 || ri >= ref->view.n_images || ci < 0 || ci >= cur->view.n_images)
            return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "pair %d references image %d/%d outside the pyramid batches", p, ri, ci);
    }
    const int n = h_offsets[n_pairs];
    const bool has_prediction = !(flags & FTK_FLAG_NO_PREDICTION), has_status = !(flags & FTK_FLAG_NO_STATUS);
    const int32_t *d_offsets = nullptr, *d_ri = nullptr, *d_ci = nullptr;
    const float *d_K = nullptr, *d_points = nullptr;
    const float2 *d_ref_uv = nullptr;
    float2 *d_cur_uv = reinterpret_cast<float2 *>(cur_uv);
    float *d_q = q_rc, *d_p = p_rc;
    uint8_t *d_status = status;
    if (int rc = Stage(ctx, ctx->d_offsets, feat_offsets, n_pairs + 1, on_device, &d_offsets)) return rc;
    if (ref_image)
        if (int rc = Stage(ctx, ctx->d_ref_img, ref_image, n_pairs, on_device, &d_ri)) return rc;
    if (cur_image)
        if (int rc = Stage(ctx, ctx->d_cur_img, cur_image, n_pairs, on_device, &d_ci)) return rc;
    if (int rc = Stage(ctx, ctx->d_dm_K, K, static_cast<size_t>(n_pairs) * 4, on_device, &d_K)) return rc;
    if (int rc = Stage(ctx, ctx->d_dm_points, p_c_in_ref, static_cast<size_t>(n) * 3, on_device, &d_points)) return rc;
    if (int rc = Stage(ctx, ctx->d_ref_uv, reinterpret_cast<const float2 *>(ref_uv), n, on_device, &d_ref_uv)) return rc;
    if (!on_device) {
        if (int rc = EnsureDevice(ctx, ctx->d_cur_uv, sizeof(float2) * n)) return rc;
        if (int rc = EnsureDevice(ctx, ctx->d_status, n)) return rc;
        if (int rc = EnsureDevice(ctx, ctx->d_dm_q, sizeof(float) * 4 * n_pairs)) return rc;
        if (int rc = EnsureDevice(ctx, ctx->d_dm_p, sizeof(float) * 3 * n_pairs)) return rc;
        d_cur_uv = static_cast<float2 *>(ctx->d_cur_uv.ptr);
        d_status = static_cast<uint8_t *>(ctx->d_status.ptr);
        d_q = static_cast<float *>(ctx->d_dm_q.ptr);
        d_p = static_cast<float *>(ctx->d_dm_p.ptr);
        if (has_prediction) FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(d_cur_uv, cur_uv, sizeof(float2) * n, cudaMemcpyHostToDevice, ctx->stream));
        if (has_status) FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(d_status, status, n, cudaMemcpyHostToDevice, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(d_q, q_rc, sizeof(float) * 4 * n_pairs, cudaMemcpyHostToDevice, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(d_p, p_rc, sizeof(float) * 3 * n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (int rc = ftk::LaunchDirectMethod(ctx, *params, ref->view, cur->view, n_pairs, d_ri, d_ci, d_offsets, d_K, d_points, d_ref_uv, d_cur_uv, d_q, d_p,
                                         d_status, n, has_prediction, has_status))
        return rc;
    if (!on_device) {
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(cur_uv, d_cur_uv, sizeof(float2) * n, cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(status, d_status, n, cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(q_rc, d_q, sizeof(float) * 4 * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(p_rc, d_p, sizeof(float) * 3 * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return FTK_OK;
}

int ftk_dense_flow_track(ftk_context *ctx, const ftk_dense_flow_params *params, const ftk_pyramid *ref, const ftk_pyramid *cur, int32_t ref_image,
                         int32_t cur_image, float *flow_row, float *flow_col, uint32_t flags) {
    if (!ctx || !params || !ref || !cur || !flow_row || !flow_col) return FTK_ERR_INVALID_ARGUMENT;
    if (ref->view.levels != cur->view.levels)  // dense_optical_flow.cpp:40
        return SetError(ctx, FTK_ERR_LEVEL_MISMATCH, "ref has %d levels, cur has %d", ref->view.levels, cur->view.levels);
    if (ref->view.rows[0] != cur->view.rows[0] || ref->view.cols[0] != cur->view.cols[0])
        return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "ref and cur pyramids differ in image size");
    if (ref_image < 0 || ref_image >= ref->view.n_images || cur_image < 0 || cur_image >= cur->view.n_images)
        return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "image %d/%d outside the pyramid batches", ref_image, cur_image);
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS, single = flags & FTK_FLAG_SINGLE_LEVEL;
    const bool initial = single && !(flags & FTK_FLAG_NO_PREDICTION);
    const size_t n0 = static_cast<size_t>(ref->view.rows[0]) * ref->view.cols[0];
    float *d_r = flow_row, *d_c = flow_col;
    if (!on_device) {
        if (int rc = EnsureDevice(ctx, ctx->d_flow, sizeof(float) * 2 * n0)) return rc;
        d_r = static_cast<float *>(ctx->d_flow.ptr);
        d_c = d_r + n0;
        if (initial) {
            FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(d_r, flow_row, sizeof(float) * n0, cudaMemcpyHostToDevice, ctx->stream));
            FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(d_c, flow_col, sizeof(float) * n0, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    if (int rc = ftk::LaunchDenseFlow(ctx, *params, ref->view, cur->view, ref_image, cur_image, single, initial, d_r, d_c)) return rc;
    if (!on_device) {
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(flow_row, d_r, sizeof(float) * n0, cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(flow_col, d_c, sizeof(float) * n0, cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return FTK_OK;
}

int ftk_detect_response(ftk_context *ctx, const ftk_detector_params *params, const ftk_pyramid *pyr, int32_t image, float *response, uint32_t flags) {
    if (!ctx || !params || !pyr || !response) return FTK_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    const size_t n = static_cast<size_t>(pyr->view.rows[0]) * pyr->view.cols[0];
    float *d_response = response;
    if (!on_device) {
        if (int rc = EnsureDevice(ctx, ctx->d_det_response, sizeof(float) * n)) return rc;
        d_response = static_cast<float *>(ctx->d_det_response.ptr);
    }
    if (int rc = ftk::LaunchDetectResponse(ctx, *params, pyr->view, image, d_response)) return rc;
    if (!on_device) {
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(response, d_response, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return FTK_OK;
}

// Shared body of ftk_detect_features / ftk_detect_features_batch: out_uv [count][needed][2], out_response [count][needed], n_out [count] (host).
static int DetectImages(ftk_context *ctx, const ftk_detector_params *params, const ftk_pyramid *pyr, int32_t first, int32_t count, const float *existing_uv,
                        int32_t n_existing, int32_t needed, float *out_uv, float *out_response, int32_t *n_out, uint32_t flags) {
    if (!ctx || !params || !pyr || !n_out || count < 1 || n_existing < 0 || needed < 0 || (n_existing > 0 && !existing_uv) || (needed > 0 && !out_uv))
        return FTK_ERR_INVALID_ARGUMENT;
    for (int32_t i = 0; i < count; ++i) n_out[i] = 0;
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    const float2 *d_existing = nullptr;
    if (int rc = Stage(ctx, ctx->d_ref_uv, reinterpret_cast<const float2 *>(existing_uv), static_cast<size_t>(n_existing), on_device, &d_existing)) return rc;
    const size_t slots = static_cast<size_t>(count) * static_cast<size_t>(needed);
    // staging: [n_out (count ints)] then, for host callers, [uv (slots float2)] [response (slots floats)]
    const size_t head = (sizeof(int32_t) * count + 15) / 16 * 16;
    if (int rc = EnsureDevice(ctx, ctx->d_det_out, head + (on_device ? 0 : sizeof(float) * 3 * slots) + 16)) return rc;
    int32_t *d_n_out = static_cast<int32_t *>(ctx->d_det_out.ptr);
    float2 *d_uv = reinterpret_cast<float2 *>(out_uv);
    float *d_resp = out_response;
    if (!on_device) {
        d_uv = reinterpret_cast<float2 *>(static_cast<char *>(ctx->d_det_out.ptr) + head);
        d_resp = reinterpret_cast<float *>(d_uv + slots);
        if (slots) FTK_CUDA_CHECK(ctx, cudaMemsetAsync(d_uv, 0, sizeof(float) * 3 * slots, ctx->stream));  // slots past n_out[i] read as zeros
    }
    if (int rc = ftk::LaunchDetectFeatures(ctx, *params, pyr->view, first, count, d_existing, n_existing, needed, d_uv, d_resp, d_n_out)) return rc;
    FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(n_out, d_n_out, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, ctx->stream));
    if (!on_device && slots) {
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(out_uv, d_uv, sizeof(float2) * slots, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_response) FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(out_response, d_resp, sizeof(float) * slots, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int32_t i = 0; i < count; ++i)
        if (n_out[i] < 0) return SetError(ctx, FTK_ERR_CUDA, "feature selection of image %d took more features than can exist", first + i);
    return FTK_OK;
}

int ftk_detect_features(ftk_context *ctx, const ftk_detector_params *params, const ftk_pyramid *pyr, int32_t image, const float *existing_uv,
                        int32_t n_existing, int32_t needed, float *out_uv, float *out_response, int32_t *n_out, uint32_t flags) {
    return DetectImages(ctx, params, pyr, image, 1, existing_uv, n_existing, needed, out_uv, out_response, n_out, flags);
}

int ftk_detect_features_batch(ftk_context *ctx, const ftk_detector_params *params, const ftk_pyramid *pyr, int32_t first_image, int32_t n_images,
                              int32_t needed, float *out_uv, float *out_response, int32_t *n_out, uint32_t flags) {
    return DetectImages(ctx, params, pyr, first_image, n_images, nullptr, 0, needed, out_uv, out_response, n_out, flags);
}

// Shared body of ftk_describe_brief / ftk_describe_brief_batch; feat_offsets (HOST, count + 1 entries) null = one image.
static int DescribeImages(ftk_context *ctx, const ftk_pyramid *pyr, int32_t first, int32_t count, const int32_t *feat_offsets, const float *uv, int32_t n,
                          const int8_t *pattern, int32_t n_bits, int32_t half_patch, uint32_t *desc, uint8_t *valid, uint32_t flags) {
    if (!ctx || !pyr || !pattern || n < 0 || (n > 0 && (!uv || !desc))) return FTK_ERR_INVALID_ARGUMENT;
    if (n_bits <= 0 || n_bits % 32 != 0 || n_bits > 1024) return SetError(ctx, FTK_ERR_UNSUPPORTED, "BRIEF length %d is not a multiple of 32 in 32..1024", n_bits);
    if (half_patch < 0 || half_patch > 127) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "BRIEF half patch %d", half_patch);
    for (int32_t k = 0; k < 4 * n_bits; ++k)
        if (pattern[k] < -half_patch || pattern[k] > half_patch) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "BRIEF pair %d leaves the +-%d patch", k / 4, half_patch);
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    const int words = n_bits / 32;
    const char4 *d_pattern = nullptr;
    if (int rc = Stage(ctx, ctx->d_det_pattern, reinterpret_cast<const char4 *>(pattern), static_cast<size_t>(n_bits), false, &d_pattern)) return rc;
    const float2 *d_uv = nullptr;
    if (int rc = Stage(ctx, ctx->d_ref_uv, reinterpret_cast<const float2 *>(uv), static_cast<size_t>(n), on_device, &d_uv)) return rc;
    int *d_feat_image = nullptr;
    if (feat_offsets && n > 0) {
        const int32_t *d_off = nullptr;
        if (int rc = Stage(ctx, ctx->d_offsets, feat_offsets, static_cast<size_t>(count) + 1, false, &d_off)) return rc;
        if (int rc = EnsureDevice(ctx, ctx->d_feat_pair, sizeof(int) * static_cast<size_t>(n))) return rc;
        d_feat_image = static_cast<int *>(ctx->d_feat_pair.ptr);
        if (int rc = ftk::LaunchFeaturePairs(ctx, d_off, count, n, d_feat_image)) return rc;
    }
    uint32_t *d_desc = desc;
    uint8_t *d_valid = valid;
    if (!on_device) {
        if (int rc = EnsureDevice(ctx, ctx->d_desc_ref, sizeof(uint32_t) * static_cast<size_t>(n ? n : 1) * words)) return rc;
        if (int rc = EnsureDevice(ctx, ctx->d_status, static_cast<size_t>(n ? n : 1))) return rc;
        d_desc = static_cast<uint32_t *>(ctx->d_desc_ref.ptr);
        d_valid = static_cast<uint8_t *>(ctx->d_status.ptr);
    }
    if (int rc = ftk::LaunchDescribeBrief(ctx, pyr->view, first, count, d_feat_image, d_uv, n, d_pattern, n_bits, half_patch, d_desc, d_valid)) return rc;
    if (!on_device && n > 0) {
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(desc, d_desc, sizeof(uint32_t) * static_cast<size_t>(n) * words, cudaMemcpyDeviceToHost, ctx->stream));
        if (valid) FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(valid, d_valid, static_cast<size_t>(n), cudaMemcpyDeviceToHost, ctx->stream));
        FTK_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return FTK_OK;
}

int ftk_describe_brief(ftk_context *ctx, const ftk_pyramid *pyr, int32_t image, const float *uv, int32_t n, const int8_t *pattern, int32_t n_bits,
                       int32_t half_patch, uint32_t *desc, uint8_t *valid, uint32_t flags) {
    return DescribeImages(ctx, pyr, image, 1, nullptr, uv, n, pattern, n_bits, half_patch, desc, valid, flags);
}

int ftk_describe_brief_batch(ftk_context *ctx, const ftk_pyramid *pyr, int32_t first_image, int32_t n_images, const int32_t *feat_offsets, const float *uv,
                             const int8_t *pattern, int32_t n_bits, int32_t half_patch, uint32_t *desc, uint8_t *valid, uint32_t flags) {
    if (!feat_offsets || n_images < 1) return FTK_ERR_INVALID_ARGUMENT;
    for (int32_t i = 0; i < n_images; ++i)
        if (feat_offsets[i + 1] < feat_offsets[i] || feat_offsets[0] != 0) return ctx ? SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "feature offsets must start at 0 and not decrease") : FTK_ERR_INVALID_ARGUMENT;
    return DescribeImages(ctx, pyr, first_image, n_images, feat_offsets, uv, feat_offsets[n_images], pattern, n_bits, half_patch, desc, valid, flags);
}

int ftk_match_mutual_scores(ftk_context *ctx, const float *scores, int32_t n_ref, int32_t n_cur, float min_score, int32_t *idx, uint32_t flags) {
    if (!ctx || !idx || n_ref < 0) return FTK_ERR_INVALID_ARGUMENT;
    if (n_cur <= 0) return SetError(ctx, FTK_ERR_EMPTY_INPUT, "the score matrix has no columns");
    if (n_ref > 0 && !scores) return FTK_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    const float *d_scores = nullptr;
    if (int rc = Stage(ctx, ctx->d_desc_ref, scores, static_cast<size_t>(n_ref) * n_cur, on_device, &d_scores)) return rc;
    int *d_idx = nullptr;
    if (int rc = PrepareIndex(ctx, idx, n_ref, flags | FTK_FLAG_NO_INDEX_INPUT, &d_idx)) return rc;  // every entry is written
    if (int rc = ftk::LaunchMutualScores(ctx, d_scores, n_ref, n_cur, min_score, d_idx)) return rc;
    return FinishIndex(ctx, idx, n_ref, flags, d_idx);
}

int ftk_match_cross_check(ftk_context *ctx, int32_t *idx_ref_to_cur, int32_t n_ref, const int32_t *idx_cur_to_ref, int32_t n_cur, uint32_t flags) {
    if (!ctx || !idx_ref_to_cur || n_ref < 0 || n_cur < 0 || (n_cur > 0 && !idx_cur_to_ref)) return FTK_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    int *d_fwd = nullptr;
    if (int rc = PrepareIndex(ctx, idx_ref_to_cur, n_ref, flags & ~FTK_FLAG_NO_INDEX_INPUT, &d_fwd)) return rc;
    const int32_t *d_bwd = nullptr;
    if (int rc = Stage(ctx, ctx->d_work2, idx_cur_to_ref, static_cast<size_t>(n_cur), on_device, &d_bwd)) return rc;
    if (int rc = ftk::LaunchCrossCheck(ctx, d_fwd, n_ref, d_bwd, n_cur)) return rc;
    return FinishIndex(ctx, idx_ref_to_cur, n_ref, flags, d_fwd);
}

int ftk_match_cosine_force(ftk_context *ctx, const float *ref, int32_t n_ref, const float *cur, int32_t n_cur, int32_t dim, float max_dist,
                           int32_t *idx, uint32_t flags) {
    if (!ctx || !idx || n_ref < 0 || dim <= 0) return FTK_ERR_INVALID_ARGUMENT;
    if (n_cur <= 0) return SetError(ctx, FTK_ERR_EMPTY_INPUT, "descriptors_cur is empty");
    if ((n_ref > 0 && !ref) || !cur) return FTK_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    const float *d_ref = nullptr, *d_cur = nullptr;
    if (int rc = Stage(ctx, ctx->d_desc_ref, ref, static_cast<size_t>(n_ref) * dim, on_device, &d_ref)) return rc;
    if (int rc = Stage(ctx, ctx->d_desc_cur, cur, static_cast<size_t>(n_cur) * dim, on_device, &d_cur)) return rc;
    int *d_idx = nullptr;
    if (int rc = PrepareIndex(ctx, idx, n_ref, flags, &d_idx, true)) return rc;
    if (int rc = ftk::LaunchCosineForce(ctx, d_ref, n_ref, d_cur, n_cur, dim, max_dist, d_idx, (flags & FTK_FLAG_NO_INDEX_INPUT) != 0)) return rc;
    return FinishIndex(ctx, idx, n_ref, flags, d_idx);
}

int ftk_match_cosine_nearby(ftk_context *ctx, const float *ref, int32_t n_ref, const float *cur, int32_t n_cur, int32_t dim, const float *pred_uv,
                            const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist, int32_t *idx, uint32_t flags) {
    if (!ctx || !idx || n_ref < 0 || dim <= 0) return FTK_ERR_INVALID_ARGUMENT;
    if (n_cur <= 0) return SetError(ctx, FTK_ERR_EMPTY_INPUT, "descriptors_cur is empty");
    if ((n_ref > 0 && (!ref || !pred_uv)) || !cur || !cur_uv) return FTK_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    const bool on_device = flags & FTK_FLAG_DEVICE_POINTERS;
    const float *d_ref = nullptr, *d_cur = nullptr;
    const float2 *d_pred = nullptr, *d_pos = nullptr;
    if (int rc = Stage(ctx, ctx->d_desc_ref, ref, static_cast<size_t>(n_ref) * dim, on_device, &d_ref)) return rc;
    if (int rc = Stage(ctx, ctx->d_desc_cur, cur, static_cast<size_t>(n_cur) * dim, on_device, &d_cur)) return rc;
    if (int rc = Stage(ctx, ctx->d_pred_uv, reinterpret_cast<const float2 *>(pred_uv), n_ref, on_device, &d_pred)) return rc;
    if (int rc = Stage(ctx, ctx->d_pos_cur, reinterpret_cast<const float2 *>(cur_uv), n_cur, on_device, &d_pos)) return rc;
    int *d_idx = nullptr;
    if (int rc = PrepareIndex(ctx, idx, n_ref, flags, &d_idx, true)) return rc;
    if (int rc = ftk::LaunchCosineNearby(ctx, d_ref, n_ref, d_cur, n_cur, dim, d_pred, d_pos, max_drow, max_dcol, max_dist, d_idx,
                                         (flags & FTK_FLAG_NO_INDEX_INPUT) != 0))
        return rc;
    return FinishIndex(ctx, idx, n_ref, flags, d_idx);
}

int ftk_last_cosine_exact_scan_items(ftk_context *ctx) {
    if (!ctx) return FTK_ERR_INVALID_ARGUMENT;
    if (!ctx->d_last_scan_items) return -1;
    DeviceGuard guard(ctx->device);
    int n = 0;
    if (cudaMemcpyAsync(&n, ctx->d_last_scan_items, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    return n;
}

// descriptor_matcher.h:135-157 FillMatchedPixelByPairIndices (negligible host work in the reference; kept on the host).
int ftk_fill_matched(const int32_t *idx, int32_t n_ref, const float *cur_uv, int32_t n_cur, float *matched_uv, uint8_t *status, int32_t status_valid) {
    if (n_ref < 0 || (n_ref > 0 && (!idx || !matched_uv || !status))) return FTK_ERR_INVALID_ARGUMENT;
    if (!status_valid) memset(status, FTK_STATUS_NOT_TRACKED, n_ref);
    for (int32_t i = 0; i < n_ref; ++i) {
        if (status[i] > FTK_STATUS_TRACKED) continue;
        const int32_t j = idx[i];
        if (j >= 0 && j < n_cur) {
            matched_uv[2 * i] = cur_uv[2 * j];
            matched_uv[2 * i + 1] = cur_uv[2 * j + 1];
            status[i] = FTK_STATUS_TRACKED;
        } else {
            status[i] = FTK_STATUS_LARGE_RESIDUAL;
        }
    }
    return FTK_OK;
}

}  // extern "C"
