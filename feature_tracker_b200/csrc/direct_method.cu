// K9: direct-method pose tracker.
//
// Replaces DirectMethod::TrackFeatures (camera-frame overload, src/direct_method_tracker/direct_method_tracker.cpp:41-95) and
// TrackAllFeaturesDirect (:115-192): a 6-DoF Gauss-Newton alignment of the current frame against the reference frame in which
// ALL features of a frame pair feed ONE 6x6 normal equation.  kInverse / kFast are empty upstream (:107-113, :194-199) and
// therefore do nothing here either.
//
// Mapping: one CTA per frame pair (a batch of pairs fills the GPU; a single pair is latency bound by construction, see below).
// Per Gauss-Newton iteration
//   1. feature pass   - thread per feature: project with the current pose, write cur_uv (the reference's side effect, :143),
//                       the 2x6 Jacobian (:146-150) and a validity flag to a scratch row;
//   2. pixel pass     - warps 1..7 evaluate tiles of 224 patch pixels (feature-major, row-major inside the patch: the
//                       reference's loop order): gradient stencil + residual -> the 21 + 6 per-pixel terms of H and b, written
//                       to a double-buffered shared tile;  warp 0 folds the previous tile: lane k adds term k of the 224
//                       pixels in order into accumulator k.  The 27 sums are therefore the reference's sequential sums, bit
//                       for bit, while the evaluation of the next tile overlaps the (serial, 4-cycle-per-add) fold;
//   3. solve          - warp 0: LDLT<6> + pose update (:178-187), convergence flags through shared memory.
// All arithmetic is the reference's own sequence of fp32 operations (no FMA contraction; quaternion / camera semantics as
// frozen in oracle/shim/basic_type.h and camera_pinhole.h).
#include <cfloat>

#include "klt_device.cuh"

namespace ftk {

namespace {

constexpr int kDmThreads = 256;
constexpr int kDmProducerWarps = kDmThreads / 32 - 1;
constexpr int kDmTile = kDmProducerWarps * 32;  // pixels per tile
constexpr int kDmChains = 27;
constexpr int kDmStride = kDmTile + 4;
constexpr int kDmScratch = 16;  // floats per feature: J[12], valid, pad
constexpr float kDmZeroFloat = 1e-6f;  // kZeroFloat

struct DmArgs {
    ftk_direct_params p;
    PyramidView ref, cur;
    int n_pairs;
    const int *ref_image, *cur_image;  // may be null (image = pair)
    const int *feat_offsets;
    const float4 *K;        // per pair (fx, fy, cx, cy)
    const float *points;    // n x 3
    const float2 *ref_uv;
    float2 *cur_uv;         // in/out
    float4 *q_rc;           // per pair (w, x, y, z), in/out
    float *p_rc;            // per pair x 3, in/out
    uint8_t *status;        // in/out
    float *scratch;         // n x kDmScratch
    int has_prediction, has_status;
};

__device__ __forceinline__ void QuatRotate(const float (&q)[4], const float (&v)[3], float (&out)[3]) {
    const float qw = q[0], qx = q[1], qy = q[2], qz = q[3];
    float ux = fsub(fmul(qy, v[2]), fmul(qz, v[1])), uy = fsub(fmul(qz, v[0]), fmul(qx, v[2])), uz = fsub(fmul(qx, v[1]), fmul(qy, v[0]));
    ux = fadd(ux, ux), uy = fadd(uy, uy), uz = fadd(uz, uz);
    out[0] = fadd(fadd(v[0], fmul(qw, ux)), fsub(fmul(qy, uz), fmul(qz, uy)));
    out[1] = fadd(fadd(v[1], fmul(qw, uy)), fsub(fmul(qz, ux), fmul(qx, uz)));
    out[2] = fadd(fadd(v[2], fmul(qw, uz)), fsub(fmul(qx, uy), fmul(qy, ux)));
}
__device__ __forceinline__ float QuatSquaredNorm(const float (&q)[4]) {
    return fadd(fadd(fadd(fmul(q[1], q[1]), fmul(q[2], q[2])), fmul(q[3], q[3])), fmul(q[0], q[0]));
}
__device__ __forceinline__ void QuatNormalize(float (&q)[4]) {
    const float n = __fsqrt_rn(QuatSquaredNorm(q));
    q[0] = fdiv(q[0], n), q[1] = fdiv(q[1], n), q[2] = fdiv(q[2], n), q[3] = fdiv(q[3], n);
}

struct DmShared {
    float term[2][kDmChains * kDmStride];
    float q[4], q_inv[4], p[3];
    float K[4];  // this level's scaled intrinsics
    int stop;    // the level's iteration loop ends
    LdltShared<6> ldlt;  // normal equations / factors of the cooperative 6 x 6 solve
};

__global__ void __launch_bounds__(kDmThreads) DirectMethodKernel(DmArgs a) {
    extern __shared__ __align__(16) unsigned char dm_smem_raw[];
    DmShared &sm = *reinterpret_cast<DmShared *>(dm_smem_raw);
    const int pair = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int f0 = a.feat_offsets[pair], f1 = a.feat_offsets[pair + 1];
    const int n = f1 - f0;
    const int ref_image = a.ref_image ? a.ref_image[pair] : pair;
    const int cur_image = a.cur_image ? a.cur_image[pair] : pair;
    const int levels = a.ref.levels;
    const int hr = a.p.patch_row_half, hc = a.p.patch_col_half;
    const int pc = 2 * hc + 1, P = (2 * hr + 1) * pc;
    const int n_track = min(n, static_cast<int>(min(a.p.max_track_points, 0x7FFFFFFFu)));
    const float2 *ref_uv = a.ref_uv + f0;
    float2 *cur_uv = a.cur_uv + f0;
    const float *points = a.points + 3 * static_cast<size_t>(f0);
    float *scratch = a.scratch + static_cast<size_t>(f0) * kDmScratch;

    // :48-50: no prediction -> cur = ref
    if (!a.has_prediction)
        for (int i = tid; i < n; i += kDmThreads) cur_uv[i] = ref_uv[i];
    if (tid == 0) {
        const float4 q = a.q_rc[pair];
        sm.q[0] = q.x, sm.q[1] = q.y, sm.q[2] = q.z, sm.q[3] = q.w;
        sm.p[0] = a.p_rc[3 * pair], sm.p[1] = a.p_rc[3 * pair + 1], sm.p[2] = a.p_rc[3 * pair + 2];
    }
    __syncthreads();

    const float scale = static_cast<float>(1 << (levels - 1));
    const float4 K0 = a.K[pair];
    const bool active = a.p.method == 1 && n_track > 0;  // kDirect; the other methods are empty upstream

    for (int level = levels - 1; level > -1 && active; --level) {
        const Img ref = LevelImage(a.ref, ref_image, level), cur = LevelImage(a.cur, cur_image, level);
        // K / scale, then doubled once per finer level (:59, :76-78): doubling is exact, so this is the same value
        const float up = static_cast<float>(1 << (levels - 1 - level));
        if (tid == 0) {
            sm.K[0] = fmul(fdiv(K0.x, scale), up), sm.K[1] = fmul(fdiv(K0.y, scale), up);
            sm.K[2] = fmul(fdiv(K0.z, scale), up), sm.K[3] = fmul(fdiv(K0.w, scale), up);
            sm.stop = 0;
        }
        __syncthreads();
        const float fx = sm.K[0], fy = sm.K[1], cx = sm.K[2], cy = sm.K[3];
        const long long total_px = static_cast<long long>(n_track) * P;
        const int n_tiles = static_cast<int>((total_px + kDmTile - 1) / kDmTile);

        for (uint32_t iter = 0; iter < a.p.max_iteration; ++iter) {
            // ---- 1. feature pass (:128-150) ----
            if (tid == 0) {
                const float n2 = QuatSquaredNorm(sm.q);
                if (n2 > 0.0f) {
                    sm.q_inv[0] = fdiv(sm.q[0], n2), sm.q_inv[1] = fdiv(-sm.q[1], n2), sm.q_inv[2] = fdiv(-sm.q[2], n2), sm.q_inv[3] = fdiv(-sm.q[3], n2);
                } else {
                    sm.q_inv[0] = sm.q_inv[1] = sm.q_inv[2] = sm.q_inv[3] = 0.0f;
                }
            }
            __syncthreads();
            {
                const float qi[4] = {sm.q_inv[0], sm.q_inv[1], sm.q_inv[2], sm.q_inv[3]};
                const float pr0 = sm.p[0], pr1 = sm.p[1], pr2 = sm.p[2];
                for (int i = tid; i < n_track; i += kDmThreads) {
                    const float x = points[3 * i], y = points[3 * i + 1], z = points[3 * i + 2];
                    float *s = scratch + static_cast<size_t>(i) * kDmScratch;
                    float valid = 0.0f;
                    if (!(z < kDmZeroFloat)) {
                        const float d[3] = {fsub(x, pr0), fsub(y, pr1), fsub(z, pr2)};
                        float pcur[3];
                        QuatRotate(qi, d, pcur);
                        if (!(pcur[2] < kDmZeroFloat)) {
                            const float nx = fdiv(pcur[0], pcur[2]), ny = fdiv(pcur[1], pcur[2]);
                            cur_uv[i] = make_float2(fadd(fmul(fx, nx), cx), fadd(fmul(fy, ny), cy));
                            const float z_inv = fdiv(1.0f, z);
                            const float z2_inv = fmul(z_inv, z_inv);
                            s[0] = fmul(fx, z_inv);
                            s[1] = 0.0f;
                            s[2] = fmul(fmul(-fx, x), z2_inv);
                            s[3] = fmul(fmul(fmul(-fx, x), y), z2_inv);
                            s[4] = fadd(fx, fmul(fmul(fmul(fx, x), x), z2_inv));
                            s[5] = fmul(fmul(-fx, y), z_inv);
                            s[6] = 0.0f;
                            s[7] = fmul(fy, z_inv);
                            s[8] = fmul(fmul(-fy, y), z2_inv);
                            s[9] = fsub(-fy, fmul(fmul(fmul(fy, y), y), z2_inv));
                            s[10] = fmul(fmul(fmul(fy, x), y), z2_inv);
                            s[11] = fmul(fmul(fy, x), z_inv);
                            valid = 1.0f;
                        }
                    }
                    s[12] = valid;
                }
            }
            __syncthreads();

            // ---- 2. pixel pass: producers one tile ahead of the folding warp ----
            float acc = 0.0f;
            // producer state: my pixel of tile 0 is g = ptid; (feature, pixel-in-patch) advance by kDmTile per tile
            const int ptid = tid - 32;
            int feat = ptid >= 0 ? ptid / P : 0, k = ptid >= 0 ? ptid - (ptid / P) * P : 0;
            const int step_f = kDmTile / P, step_k = kDmTile - step_f * P;
            for (int t = 0; t <= n_tiles; ++t) {
                if (warp > 0 && t < n_tiles) {
                    float *tb = sm.term[t & 1];
                    float term[kDmChains];
                    float gx = 0.0f, gy = 0.0f, residual = 0.0f;
                    float J[12];
#pragma unroll
                    for (int q = 0; q < 12; ++q) J[q] = 0.0f;
                    if (feat < n_track) {
                        const float *s = scratch + static_cast<size_t>(feat) * kDmScratch;
                        if (s[12] != 0.0f) {
                            const int prow = k / pc, pcol = k - prow * pc;
                            const float fdrow = static_cast<float>(prow - hr), fdcol = static_cast<float>(pcol - hc);
                            // scaled reference position: ref / scale, doubled per finer level (:56-58, :73-75)
                            const float2 r = ref_uv[feat];
                            const float2 c = cur_uv[feat];
                            const float row_i = fadd(fdrow, fmul(fdiv(r.y, scale), up)), col_i = fadd(fdcol, fmul(fdiv(r.x, scale), up));
                            const float row_j = fadd(fdrow, c.y), col_j = fadd(fdcol, c.x);
                            float v0, v1, v2, v3, v4, v5;
                            if (PxStencil5(cur, row_j, col_j, &v0, &v1, &v2, &v3, &v5) && PxChecked(ref, row_i, col_i, &v4)) {
                                gx = fmul(fsub(v1, v0), 0.5f);
                                gy = fmul(fsub(v3, v2), 0.5f);
                                residual = fsub(v5, v4);
                                const float4 j0 = *reinterpret_cast<const float4 *>(s), j1 = *reinterpret_cast<const float4 *>(s + 4),
                                             j2 = *reinterpret_cast<const float4 *>(s + 8);
                                J[0] = j0.x, J[1] = j0.y, J[2] = j0.z, J[3] = j0.w, J[4] = j1.x, J[5] = j1.y;
                                J[6] = j1.z, J[7] = j1.w, J[8] = j2.x, J[9] = j2.y, J[10] = j2.z, J[11] = j2.w;
                            }
                        }
                    }
                    // invalid pixels run the same products on zeros: every term is +-0, a no-op in the chains
                    float jac[6];
#pragma unroll
                    for (int q = 0; q < 6; ++q) jac[q] = fadd(fmul(gx, J[q]), fmul(gy, J[6 + q]));
                    int q = 0;
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int c = r; c < 6; ++c) term[q++] = fmul(jac[r], jac[c]);
#pragma unroll
                    for (int r = 0; r < 6; ++r) term[21 + r] = fmul(residual, jac[r]);
#pragma unroll
                    for (int c = 0; c < kDmChains; ++c) tb[c * kDmStride + ptid] = term[c];
                    feat += step_f, k += step_k;
                    if (k >= P) k -= P, ++feat;
                }
                if (warp == 0 && t > 0 && lane < kDmChains) {
                    const float4 *t4 = reinterpret_cast<const float4 *>(sm.term[(t - 1) & 1] + lane * kDmStride);
#pragma unroll 8
                    for (int q = 0; q < kDmTile / 4; ++q) {
                        const float4 v = t4[q];
                        acc = fadd(acc, v.x);
                        acc = fadd(acc, v.y);
                        acc = fadd(acc, v.z);
                        acc = fadd(acc, v.w);
                    }
                }
                __syncthreads();
            }

            // ---- 3. solve + pose update (:178-187), warp 0 ----
            if (warp == 0) {
                // chain lanes store the normal equations (lower-triangle position a[col][row], bias b) into the shared LDLT scratch;
                // the 6 x 6 system is solved cooperatively (klt_device.cuh LdltFactorShared: lane i owns row i; no local memory)
                float dx[6];
                {
                    int q = 0;
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int c = r; c < 6; ++c) {
                            if (lane == q) sm.ldlt.a[c * 6 + r] = acc;
                            ++q;
                        }
                    if (lane >= 21 && lane < 27) sm.ldlt.b[lane - 21] = acc;
                    __syncwarp();
                }
                const Group<32> g32;
                LdltFactorShared<6, 32>(g32, sm.ldlt);
                LdltSolveShared<6, 32>(g32, sm.ldlt, dx);
                if (lane == 0) {
                    bool any_nan = false;
#pragma unroll
                    for (int r = 0; r < 6; ++r) any_nan = any_nan || dx[r] != dx[r];
                    if (any_nan) {
                        sm.stop = 1;  // BREAK_IF(isnan) before the update
                    } else {
                        sm.p[0] = fadd(sm.p[0], dx[0]), sm.p[1] = fadd(sm.p[1], dx[1]), sm.p[2] = fadd(sm.p[2], dx[2]);
                        float dq[4] = {1.0f, fmul(dx[3], 0.5f), fmul(dx[4], 0.5f), fmul(dx[5], 0.5f)};
                        QuatNormalize(dq);
                        const float (&b4)[4] = sm.q;
                        float qn[4];
                        qn[0] = fsub(fsub(fsub(fmul(dq[0], b4[0]), fmul(dq[1], b4[1])), fmul(dq[2], b4[2])), fmul(dq[3], b4[3]));
                        qn[1] = fsub(fadd(fadd(fmul(dq[0], b4[1]), fmul(dq[1], b4[0])), fmul(dq[2], b4[3])), fmul(dq[3], b4[2]));
                        qn[2] = fsub(fadd(fadd(fmul(dq[0], b4[2]), fmul(dq[2], b4[0])), fmul(dq[3], b4[1])), fmul(dq[1], b4[3]));
                        qn[3] = fsub(fadd(fadd(fmul(dq[0], b4[3]), fmul(dq[3], b4[0])), fmul(dq[1], b4[2])), fmul(dq[2], b4[1]));
                        QuatNormalize(qn);
                        sm.q[0] = qn[0], sm.q[1] = qn[1], sm.q[2] = qn[2], sm.q[3] = qn[3];
                        float sq = fmul(dx[0], dx[0]);
#pragma unroll
                        for (int r = 1; r < 6; ++r) sq = fadd(sq, fmul(dx[r], dx[r]));
                        if (sq < a.p.max_converge_step) sm.stop = 1;
                    }
                }
            }
            __syncthreads();
            if (sm.stop) break;
        }
        __syncthreads();  // everyone has read sm.stop before the next level resets it
    }

    // :81-91: status defaults to kTracked, then the outside test on the reference pyramid's level-0 size
    const float max_x = static_cast<float>(a.ref.cols[0] - 1), max_y = static_cast<float>(a.ref.rows[0] - 1);
    for (int i = tid; i < n; i += kDmThreads) {
        uint8_t st = a.has_status ? a.status[f0 + i] : static_cast<uint8_t>(FTK_STATUS_TRACKED);
        const float2 c = cur_uv[i];
        if (c.x < 0.0f || c.x > max_x || c.y < 0.0f || c.y > max_y) st = FTK_STATUS_OUTSIDE;
        a.status[f0 + i] = st;
    }
    if (tid == 0) {
        a.q_rc[pair] = make_float4(sm.q[0], sm.q[1], sm.q[2], sm.q[3]);
        a.p_rc[3 * pair] = sm.p[0], a.p_rc[3 * pair + 1] = sm.p[1], a.p_rc[3 * pair + 2] = sm.p[2];
    }
}

}  // namespace

int LaunchDirectMethod(ftk_context *ctx, const ftk_direct_params &p, const PyramidView &ref, const PyramidView &cur, int n_pairs, const int *d_ref_image,
                       const int *d_cur_image, const int *d_offsets, const float *d_K, const float *d_points, const float2 *d_ref_uv, float2 *d_cur_uv,
                       float *d_q_rc, float *d_p_rc, uint8_t *d_status, int n_features, bool has_prediction, bool has_status) {
    if (p.patch_row_half < 0 || p.patch_col_half < 0) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "negative patch half size");
    if (int rc = EnsureDevice(ctx, ctx->d_work0, sizeof(float) * kDmScratch * static_cast<size_t>(n_features))) return rc;
    DmArgs a{};
    a.p = p;
    a.ref = ref;
    a.cur = cur;
    a.n_pairs = n_pairs;
    a.ref_image = d_ref_image;
    a.cur_image = d_cur_image;
    a.feat_offsets = d_offsets;
    a.K = reinterpret_cast<const float4 *>(d_K);
    a.points = d_points;
    a.ref_uv = d_ref_uv;
    a.cur_uv = d_cur_uv;
    a.q_rc = reinterpret_cast<float4 *>(d_q_rc);
    a.p_rc = d_p_rc;
    a.status = d_status;
    a.scratch = static_cast<float *>(ctx->d_work0.ptr);
    a.has_prediction = has_prediction ? 1 : 0;
    a.has_status = has_status ? 1 : 0;
    const size_t smem = sizeof(DmShared);
    FTK_CUDA_CHECK(ctx, cudaFuncSetAttribute(DirectMethodKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    DirectMethodKernel<<<n_pairs, kDmThreads, smem, ctx->stream>>>(a);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

}  // namespace ftk
