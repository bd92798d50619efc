// K2, north-star configuration (basic KLT, kInverse, 13x13 / 15x15): the fast path of klt_basic_fastpath.cu with the
// sequential sums POOLED across the eight features of a CTA.
//
// Why: the normal-equation sums must be added in the reference's pixel order (bit-exact parity), i.e. 225 dependent FADDs per
// accumulator.  A warp instruction advances one add of every accumulator that lives in the warp; with two features per warp only
// 4-6 of the 32 lanes did useful work in the per-row folds, and those folds (16 FADD + 4 LDS.128 + 2 barriers per patch row)
// were ~25 % of all issued instructions.  Here the four warps of a CTA first write the per-pixel terms of the whole patch to
// shared memory (no fold in the row loop), then ONE warp folds the chains of all eight features at once (lane = feature x
// chain), and lanes 0-7 of that warp also run the eight 2x2 LDLT solves and status updates in parallel.  Same operations, same
// order per accumulator => same bits; about 1.4x fewer instructions per feature.
//
// The CTA moves in lock step (two __syncthreads per Gauss-Newton iteration); warps whose two features have converged skip their
// sampling pass.  Four CTAs per SM keep the issue slots busy while one CTA's fold warp walks its 240-add dependency chain.
#include "klt_fast_common.cuh"

namespace ftk {

namespace {

using namespace fastk;

constexpr int kFeat = kGroupsPerBlock;  // 8 features per CTA
constexpr int kWarpsPerCta = kThreads / 32;

template <int PR>
struct FeatSmem {
    static constexpr int kTerms = PR * kG;        // one term per (patch row, lane)
    static constexpr int kChainStride = kTerms + 4;  // float4-aligned; +4 spreads the fold lanes over the banks
    Entry rows[3 * PR];                           // setup: {R0, Rm, Rp} per patch row; iteration: the first PR entries
    float fx[PR][kG], fy[PR][kG], iref[PR][kG];   // per-level reference gradients / centre samples
    float term[3][kChainStride];                  // setup: 3 Hessian chains; iteration: 2 bias chains
    unsigned okbits[kG];                          // per column: this iteration's validity bits (masked Hessian)
    float cur_x, cur_y;
    float h[3];                                   // Hessian over the level's reference-valid mask
    int status, running, valid, mask_changed;
    static constexpr int kRawWords = 4 * 3 * PR + 3 * PR * kG + 3 * kChainStride + kG + 9;
    static constexpr int kPad = ((16 - kRawWords % 32) + 32) % 32;  // feature stride = 16 (mod 32) words: the two groups of a warp use different banks
    float pad[kPad == 0 ? 32 : kPad];
};

// Sequential fold of one chain (n4 float4 groups) -- the reference's pixel order.
template <bool SUBTRACT>
__device__ __forceinline__ float FoldChain(const float *chain, int n4) {
    const float4 *t4 = reinterpret_cast<const float4 *>(chain);
    float acc = 0.0f;
#pragma unroll 4
    for (int q = 0; q < n4; ++q) {
        const float4 v = t4[q];
        if (SUBTRACT) {
            acc = fsub(acc, v.x);
            acc = fsub(acc, v.y);
            acc = fsub(acc, v.z);
            acc = fsub(acc, v.w);
        } else {
            acc = fadd(acc, v.x);
            acc = fadd(acc, v.y);
            acc = fadd(acc, v.z);
            acc = fadd(acc, v.w);
        }
    }
    return acc;
}

// Hessian chain k (0: fx*fx, 1: fx*fy, 2: fy*fy) over the pixels of this iteration's validity mask, straight from the hoisted gradients.
template <int PR>
__device__ __forceinline__ float MaskedHessianChain(const FeatSmem<PR> &t, int k) {
    float acc = 0.0f;
#pragma unroll 1
    for (int r = 0; r < PR; ++r) {
#pragma unroll 4
        for (int c = 0; c < kG; ++c) {
            const bool ok = (t.okbits[c] >> r) & 1u;
            const float fx = t.fx[r][c], fy = t.fy[r][c];
            const float p = fmul(k == 2 ? fy : fx, k == 0 ? fx : fy);
            acc = fadd(acc, ok ? p : 0.0f);
        }
    }
    return acc;
}

// Phase 1 of the per-level setup: five reference samples per pixel -> hoisted fx, fy, I_ref and the 3 Hessian terms of every pixel.
template <int PR, bool REGULAR>
__device__ __forceinline__ void SetupTerms(const Img &ref, FeatSmem<PR> &sm, int lane, const Entry &C0, const Entry &Cm, const Entry &Cp, unsigned okbits) {
    float s0[4], s1[4], s2[4], s3[4];
    const uint8_t *colp = ref.p + Clamp(Cm.base, 0, ref.cols - 3);
    if (REGULAR) {
        const int rr = sm.rows[1].base;
        const uint8_t *p = colp + Clamp(rr, 0, ref.rows) * ref.pitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) s0[i] = LoadPx(p + i);
        p = colp + Clamp(rr + 1, 0, ref.rows) * ref.pitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) s1[i] = LoadPx(p + i);
        p = colp + Clamp(rr + 2, 0, ref.rows) * ref.pitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) s2[i] = LoadPx(p + i);
    }
#pragma unroll 1
    for (int r = 0; r < PR; ++r) {
        const Entry R0 = sm.rows[3 * r], Rm = sm.rows[3 * r + 1], Rp = sm.rows[3 * r + 2];
        float v0, v1, v2, v3, v4;
        if (REGULAR) {
            const uint8_t *p = colp + R0.off;
#pragma unroll
            for (int i = 0; i < 4; ++i) s3[i] = LoadPx(p + i);
            v0 = Bilerp(R0, Cm, s1[0], s1[1], s2[0], s2[1]);
            v1 = Bilerp(R0, Cp, s1[2], s1[3], s2[2], s2[3]);
            v2 = Bilerp(Rm, C0, s0[1], s0[2], s1[1], s1[2]);
            v3 = Bilerp(Rp, C0, s2[1], s2[2], s3[1], s3[2]);
            v4 = Bilerp(R0, C0, s1[1], s1[2], s2[1], s2[2]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                s0[i] = s1[i];
                s1[i] = s2[i];
                s2[i] = s3[i];
            }
        } else {
            v0 = SampleDirect(ref, R0, Cm);
            v1 = SampleDirect(ref, R0, Cp);
            v2 = SampleDirect(ref, Rm, C0);
            v3 = SampleDirect(ref, Rp, C0);
            v4 = SampleDirect(ref, R0, C0);
        }
        const bool ok = (okbits >> r) & 1u;
        const float fx = fsub(v1, v0), fy = fsub(v3, v2);
        sm.fx[r][lane] = fx;
        sm.fy[r][lane] = fy;
        sm.iref[r][lane] = v4;
        sm.term[0][r * kG + lane] = ok ? fmul(fx, fx) : 0.0f;
        sm.term[1][r * kG + lane] = ok ? fmul(fx, fy) : 0.0f;
        sm.term[2][r * kG + lane] = ok ? fmul(fy, fy) : 0.0f;
    }
}

// Phase 1 of one iteration: current-image sample, residual, the 2 bias terms of every pixel (the fold subtracts them).
template <int PR, bool REGULAR>
__device__ __forceinline__ void IterateTerms(const Img &cur, FeatSmem<PR> &sm, int lane, const Entry &Cj, unsigned okbits) {
    const uint8_t *colp = cur.p + Clamp(Cj.base, 0, cur.cols - 1);
    float top0 = 0.0f, top1 = 0.0f;
    if (REGULAR) {
        const uint8_t *p = colp + Clamp(sm.rows[0].base, 0, cur.rows) * cur.pitch;
        top0 = LoadPx(p);
        top1 = LoadPx(p + 1);
    }
#pragma unroll 1
    for (int r = 0; r < PR; ++r) {
        const Entry Rj = sm.rows[r];
        float v5;
        if (REGULAR) {
            const uint8_t *p = colp + Rj.off;
            const float bot0 = LoadPx(p), bot1 = LoadPx(p + 1);
            v5 = Bilerp(Rj, Cj, top0, top1, bot0, bot1);
            top0 = bot0;
            top1 = bot1;
        } else {
            v5 = SampleDirect(cur, Rj, Cj);
        }
        const bool ok = (okbits >> r) & 1u;
        const float ft = fsub(v5, sm.iref[r][lane]);
        sm.term[0][r * kG + lane] = ok ? fmul(sm.fx[r][lane], ft) : 0.0f;
        sm.term[1][r * kG + lane] = ok ? fmul(sm.fy[r][lane], ft) : 0.0f;
    }
}

template <int PR, int PC>
__global__ void __launch_bounds__(kThreads) BasicInversePooledKernel(KltLaunch a) {
    static_assert(PC <= kG && PR <= kG, "patch must fit one 16-lane group");
    constexpr int HR = PR / 2, HC = PC / 2;
    constexpr unsigned kRowMask = (1u << PR) - 1u;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FeatSmem<PR> *feats = reinterpret_cast<FeatSmem<PR> *>(smem_raw);
    __shared__ int cta_any_running;

    Lanes g;
    g.lane = threadIdx.x & (kG - 1);
    g.base = (threadIdx.x & 31) - g.lane;
    g.mask = 0xFFFFu << g.base;
    const int lane = g.lane;
    const int warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
    const int group_in_block = threadIdx.x / kG;
    const int f_raw = blockIdx.x * kFeat + group_in_block;
    const bool exists = f_raw < a.n_features;
    const int f = exists ? f_raw : a.n_features - 1;  // idle groups shadow the last feature and write nothing
    FeatSmem<PR> &sm = feats[group_in_block];
    const bool col_active = lane < PC;
    const float dcol = static_cast<float>(lane - HC);

    const int pair = a.feat_pair[f];
    const int local = f - a.feat_offsets[pair];
    const float2 ref_uv = a.ref_uv[f];
    float2 cur_uv = a.has_prediction ? a.cur_uv[f] : ref_uv;
    uint8_t status = a.has_status ? a.status[f] : static_cast<uint8_t>(FTK_STATUS_NOT_TRACKED);
    // basic_klt.cpp:9,12,15: only the first kMaxTrackPointsNumber features, never re-track failed ones
    const bool tracked = exists && static_cast<uint32_t>(local) < a.p.max_track_points && status <= FTK_STATUS_TRACKED;

    if (__syncthreads_or(tracked)) {  // CTA-uniform
        const int ref_image = a.ref_image ? a.ref_image[pair] : pair;
        const int cur_image = a.cur_image ? a.cur_image[pair] : pair;
        const int levels = a.single_level ? 1 : a.ref.levels;  // TrackSingleLevel (basic_klt.cpp:59-86) == one level, scale 1
        const float scale = static_cast<float>(1 << (levels - 1));
        float ref_x = fdiv(ref_uv.x, scale), ref_y = fdiv(ref_uv.y, scale);
        float cur_x = fdiv(cur_uv.x, scale), cur_y = fdiv(cur_uv.y, scale);
        int st = status;

        for (int level = levels - 1; level > -1; --level) {
            const Img ref = LevelImage(a.ref, ref_image, level), cur = LevelImage(a.cur, cur_image, level);
            const int cur_rows = a.cur.rows[level], cur_cols = a.cur.cols[level];  // identical for every feature of the CTA

            // ================= per-level setup =================
            unsigned refmask = 0;
            if (__any_sync(kFull, tracked)) {
                const float col_i = fadd(dcol, ref_x);
                bool c0_ok, cm_ok, cp_ok;
                const Entry C0 = MakeEntry(col_i, ref.cols, &c0_ok);
                const Entry Cm = MakeEntry(fsub(col_i, 1.0f), ref.cols, &cm_ok);
                const Entry Cp = MakeEntry(fadd(col_i, 1.0f), ref.cols, &cp_ok);
                bool rows_ok = false, regular = Cm.base == C0.base - 1 && Cp.base == C0.base + 1 && ref.cols >= 4 && ref.rows >= 4;
                int my_base = 0;
                if (lane < PR) {
                    const float row_i = fadd(static_cast<float>(lane - HR), ref_y);
                    bool r0_ok, rm_ok, rp_ok;
                    Entry R0 = MakeEntry(row_i, ref.rows, &r0_ok);
                    const Entry Rm = MakeEntry(fsub(row_i, 1.0f), ref.rows, &rm_ok);
                    const Entry Rp = MakeEntry(fadd(row_i, 1.0f), ref.rows, &rp_ok);
                    R0.off = Clamp(Rp.base + 1, 0, ref.rows) * ref.pitch;
                    sm.rows[3 * lane + 0] = R0;
                    sm.rows[3 * lane + 1] = Rm;
                    sm.rows[3 * lane + 2] = Rp;
                    rows_ok = r0_ok && rm_ok && rp_ok;
                    regular = regular && Rm.base == R0.base - 1 && Rp.base == R0.base + 1;
                    my_base = R0.base;
                }
                const int next_base = __shfl_down_sync(kFull, my_base, 1);
                if (lane + 1 < PR) regular = regular && next_base == my_base + 1;
                __syncwarp();
                const unsigned rows_ok_bits = g.bits(rows_ok) & kRowMask;
                refmask = (c0_ok && cm_ok && cp_ok && col_active) ? rows_ok_bits : 0u;
                if (__all_sync(kFull, regular)) SetupTerms<PR, true>(ref, sm, lane, C0, Cm, Cp, refmask);
                else SetupTerms<PR, false>(ref, sm, lane, C0, Cm, Cp, refmask);
            }
            if (lane == 0) {
                sm.cur_x = cur_x;
                sm.cur_y = cur_y;
                sm.status = st;
                sm.running = tracked ? 1 : 0;
            }
            __syncthreads();
            if (warp == 0 && lane32 < 3 * kFeat) {  // lane = feature x chain: the eight features' Hessians in one pass
                FeatSmem<PR> &t = feats[lane32 / 3];
                t.h[lane32 % 3] = FoldChain<false>(t.term[lane32 % 3], PR * kG / 4);
            }
            if (threadIdx.x == 0) cta_any_running = 1;
            __syncthreads();

            // ================= Gauss-Newton iterations (basic_klt.cpp:88-116), CTA in lock step =================
            bool running = tracked;
            for (uint32_t iter = 0; iter < a.p.max_iteration; ++iter) {
                // cta_any_running was written by warp 0 before the barrier that ended the previous iteration and is rewritten only
                // after this iteration's phase-1 barrier, so this read needs no barrier of its own.
                if (iter > 0 && !cta_any_running) break;
                // ---- phase 1: warps with a running feature sample the current image and write the bias terms ----
                if (__any_sync(kFull, running)) {
                    bool cj_ok;
                    const Entry Cj = MakeEntry(fadd(dcol, cur_x), cur.cols, &cj_ok);
                    bool row_ok = false, regular = true;
                    int my_base = 0;
                    if (lane < PR) {
                        Entry Rj = MakeEntry(fadd(static_cast<float>(lane - HR), cur_y), cur.rows, &row_ok);
                        Rj.off = Clamp(Rj.base + 1, 0, cur.rows) * cur.pitch;
                        sm.rows[lane] = Rj;
                        my_base = Rj.base;
                    }
                    const int next_base = __shfl_down_sync(kFull, my_base, 1);
                    if (lane + 1 < PR) regular = next_base == my_base + 1;
                    __syncwarp();
                    const unsigned rows_ok_bits = g.bits(row_ok);
                    const unsigned okbits = (cj_ok && col_active) ? (refmask & rows_ok_bits) : 0u;
                    const int valid = g.sum(__popc(okbits));
                    const bool mask_changed = g.any(okbits != refmask);
                    sm.okbits[lane] = okbits;
                    if (lane == 0) {
                        sm.valid = valid;
                        sm.mask_changed = (mask_changed && running) ? 1 : 0;
                    }
                    if (__all_sync(kFull, regular)) IterateTerms<PR, true>(cur, sm, lane, Cj, okbits);
                    else IterateTerms<PR, false>(cur, sm, lane, Cj, okbits);
                }
                __syncthreads();
                // ---- phase 2: one warp folds all chains and solves the eight 2x2 systems ----
                if (warp == 0) {
                    float acc = 0.0f;
                    if (lane32 < 2 * kFeat) acc = FoldChain<true>(feats[lane32 >> 1].term[lane32 & 1], PR * kG / 4);
                    float hacc = 0.0f;
                    const bool need_h = lane32 < 3 * kFeat && feats[lane32 / 3].mask_changed != 0 && feats[lane32 / 3].running != 0;
                    if (__any_sync(kFull, need_h)) {
                        // some reference-valid pixel left the current image: that feature's Hessian runs over fewer pixels this iteration
                        if (need_h) hacc = MaskedHessianChain<PR>(feats[lane32 / 3], lane32 % 3);
                    }
                    const int fl = lane32 & (kFeat - 1);
                    const float b0 = __shfl_sync(kFull, acc, 2 * fl), b1 = __shfl_sync(kFull, acc, 2 * fl + 1);
                    const float m00 = __shfl_sync(kFull, hacc, 3 * fl), m01 = __shfl_sync(kFull, hacc, 3 * fl + 1), m11 = __shfl_sync(kFull, hacc, 3 * fl + 2);
                    bool still = false;
                    if (lane32 < kFeat) {
                        FeatSmem<PR> &t = feats[lane32];
                        if (t.running) {
                            still = true;
                            if (t.valid == 0) {
                                still = false;  // BREAK_IF(ConstructIncrementalFunction(...) == 0)
                            } else {
                                const bool masked = t.mask_changed != 0;
                                const float h00 = masked ? m00 : t.h[0], h01 = masked ? m01 : t.h[1], h11 = masked ? m11 : t.h[2];
                                const float A[2][2] = {{h00, h01}, {h01, h11}};
                                const float b[2] = {b0, b1};
                                float v[2];
                                LdltSolve<2>(A, b, v);
                                if (v[0] != v[0] || v[1] != v[1]) {
                                    t.status = FTK_STATUS_NUMERIC_ERROR;
                                    still = false;
                                } else {
                                    const float nx = fadd(t.cur_x, v[0]), ny = fadd(t.cur_y, v[1]);
                                    t.cur_x = nx;
                                    t.cur_y = ny;
                                    if (nx < 0.0f || nx > static_cast<float>(cur_cols - 1) || ny < 0.0f || ny > static_cast<float>(cur_rows - 1)) {
                                        t.status = FTK_STATUS_OUTSIDE;
                                        still = false;
                                    } else if (fadd(fmul(v[0], v[0]), fmul(v[1], v[1])) < a.p.max_converge_step) {
                                        t.status = FTK_STATUS_TRACKED;
                                        still = false;
                                    }
                                }
                            }
                            t.running = still ? 1 : 0;
                        }
                    }
                    const bool any_still = __any_sync(kFull, still);
                    if (lane32 == 0) cta_any_running = any_still ? 1 : 0;
                }
                __syncthreads();
                cur_x = sm.cur_x;
                cur_y = sm.cur_y;
                st = sm.status;
                running = sm.running != 0;
            }
            __syncthreads();  // nobody is still reading this level's shared state

            if (level == 0) break;
            ref_x = fmul(ref_x, 2.0f), ref_y = fmul(ref_y, 2.0f);
            cur_x = fmul(cur_x, 2.0f), cur_y = fmul(cur_y, 2.0f);
        }
        if (tracked) {
            status = static_cast<uint8_t>(st);
            cur_uv = make_float2(cur_x, cur_y);
            const Img cur0 = LevelImage(a.cur, cur_image, 0);
            if (IsOutside(cur0, cur_uv.x, cur_uv.y)) status = FTK_STATUS_OUTSIDE;  // basic_klt.cpp:49-53
        }
    }
    if (exists && lane == 0) {
        a.cur_uv[f] = cur_uv;
        a.status[f] = status;
    }
}

template <int PR, int PC>
int Launch(ftk_context *ctx, const KltLaunch &a) {
    const size_t smem = sizeof(FeatSmem<PR>) * kFeat;
    auto kernel = BasicInversePooledKernel<PR, PC>;
    static bool attr_set = false;
    if (!attr_set) {
        FTK_CUDA_CHECK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        attr_set = true;
    }
    const int blocks = (a.n_features + kFeat - 1) / kFeat;
    kernel<<<blocks, kThreads, smem, ctx->stream>>>(a);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

}  // namespace

// Returns FTK_ERR_UNSUPPORTED when no specialisation covers the configuration.
int LaunchKltBasicPooled(ftk_context *ctx, const KltLaunch &a) {
    if (a.p.variant != FTK_VARIANT_BASIC || a.p.method != FTK_METHOD_INVERSE) return FTK_ERR_UNSUPPORTED;
    if (a.p.patch_row_half == 7 && a.p.patch_col_half == 7) return Launch<15, 15>(ctx, a);
    if (a.p.patch_row_half == 6 && a.p.patch_col_half == 6) return Launch<13, 13>(ctx, a);
    return FTK_ERR_UNSUPPORTED;
}

}  // namespace ftk
