// K2 fast path: basic (2-DoF) KLT, kInverse method, patches up to 16 columns wide (13x13, 15x15 instantiated).
// Same results, bit for bit, as the generic kernel in klt.cu and as the reference
// (src/optical_flow_tracker/basic_klt/optical_flow_basic_klt.cpp:7-181); only the work decomposition differs.
//
// Mapping on sm_100a
//   * 16 lanes per feature, lane = patch column; two features per warp; the row loop is fully unrolled.
//   * The reference recomputes, every Gauss-Newton iteration, five bilinear samples of the REFERENCE image per pixel
//     (:127-134).  They do not depend on the iteration, so they are evaluated once per pyramid level and kept in
//     registers (fx, fy, I_ref per row).  The 2x2 Hessian over the "all reference samples valid" mask is also built
//     once per level; it is rebuilt only in iterations whose current-image validity mask differs (border features),
//     which reproduces the reference's per-iteration Hessian exactly because the summands and their order are equal.
//   * floor / fraction / bounds test of a sample position are separable in row and column: each lane keeps its column
//     entries in registers and the row entries of the level live in a small shared table, so a bilinear sample costs
//     4 weight products + 4 products + 3 sums.  Every product and sum is the same fp32 operation the reference
//     performs, in the same order.
//   * Pixel bytes are fetched once into a register strip that rolls down the patch (2 new bytes per pixel for the
//     current image, 4 for the reference image) whenever the integer bases advance regularly, which is the case unless
//     a coordinate rounds across an integer or the patch touches the image border; otherwise the group falls back to
//     loading every sample's four bytes directly.
//   * Normal-equation sums use the chain scheme of klt_device.cuh: lane k of a group folds the k-th term of the 16
//     pixels of a row in pixel order, so the sums equal the reference's sequential sums bit for bit.
#include "klt_fast_common.cuh"

namespace ftk {

namespace {

using namespace fastk;

template <int PR, int PC>
struct GroupSmem {
    float term[5 * (kG + 4)];
    Entry rows[3 * PR];  // setup: {R0, Rm, Rp} per patch row; iteration: the first PR entries
    float fx[PR][kG], fy[PR][kG], iref[PR][kG];  // per-level reference gradients / samples, [patch row][lane]
    static constexpr int kRawWords = 5 * (kG + 4) + 4 * 3 * PR + 3 * PR * kG;
    // pad the group stride to 16 (mod 32) words so the two groups of a warp hit different banks
    static constexpr int kPad = ((16 - kRawWords % 32) + 32) % 32;
    float pad[kPad == 0 ? 32 : kPad];
};

// Lane k < K of each group folds the k-th term of the group's 16 pixels in pixel order (see klt_device.cuh Chain).
// SUBTRACT: acc -= term (the reference's `bias -= ...`), which equals acc += (-term) bit for bit.
template <int K, bool SUBTRACT>
__device__ __forceinline__ void Fold(float *term, const Lanes &g, float &acc) {
    __syncwarp();
    if (g.lane < K) {
        const float4 *t4 = reinterpret_cast<const float4 *>(term + g.lane * (kG + 4));
#pragma unroll
        for (int q = 0; q < kG / 4; ++q) {
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
    }
    __syncwarp();
}

// Reference samples of one pyramid level -> fx, fy, I_ref per patch row (shared), 3 Hessian chains over the pixels of `okbits`.
// The row loops are deliberately NOT unrolled: the fully unrolled kernel overflowed the instruction cache
// (ncu: 12 "no_instruction" stall cycles per issued instruction).
template <int PR, bool REGULAR, typename Smem>
__device__ __forceinline__ void SetupRows(const Img &ref, Smem &sm, const Lanes &g, const Entry &C0, const Entry &Cm, const Entry &Cp,
                                          unsigned okbits, float &acc) {
    // strip rows s0..s3: pixels (rb + j, cb + i), rb = base(Rm of the current patch row), cb = base(Cm).  Rows / columns
    // outside the image are clamped to addressable ones: they only feed pixels whose validity bit is clear.
    float s0[4], s1[4], s2[4], s3[4];
    const uint8_t *colp = ref.p + Clamp(Cm.base, 0, ref.cols - 3);
    if (REGULAR) {
        const int rr = sm.rows[1].base;  // image row of strip row s0
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
#pragma unroll 3
    for (int r = 0; r < PR; ++r) {
        const Entry R0 = sm.rows[3 * r], Rm = sm.rows[3 * r + 1], Rp = sm.rows[3 * r + 2];
        float v0, v1, v2, v3, v4;
        if (REGULAR) {
            const uint8_t *p = colp + R0.off;  // the new bottom row of the strip: base(Rp) + 1, clamped (precomputed)
#pragma unroll
            for (int i = 0; i < 4; ++i) s3[i] = LoadPx(p + i);
            v0 = Bilerp(R0, Cm, s1[0], s1[1], s2[0], s2[1]);  // (row_i, col_i - 1)
            v1 = Bilerp(R0, Cp, s1[2], s1[3], s2[2], s2[3]);  // (row_i, col_i + 1)
            v2 = Bilerp(Rm, C0, s0[1], s0[2], s1[1], s1[2]);  // (row_i - 1, col_i)
            v3 = Bilerp(Rp, C0, s2[1], s2[2], s3[1], s3[2]);  // (row_i + 1, col_i)
            v4 = Bilerp(R0, C0, s1[1], s1[2], s2[1], s2[2]);  // (row_i, col_i)
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
        sm.fx[r][g.lane] = fx;
        sm.fy[r][g.lane] = fy;
        sm.iref[r][g.lane] = v4;
        sm.term[0 * (kG + 4) + g.lane] = ok ? fmul(fx, fx) : 0.0f;
        sm.term[1 * (kG + 4) + g.lane] = ok ? fmul(fx, fy) : 0.0f;
        sm.term[2 * (kG + 4) + g.lane] = ok ? fmul(fy, fy) : 0.0f;
        Fold<3, false>(sm.term, g, acc);
    }
}

// One Gauss-Newton iteration's pass over the patch: current-image sample, residual, 2 bias chains (acc -= g * ft).
template <int PR, bool REGULAR, typename Smem>
__device__ __forceinline__ void IterateRows(const Img &cur, Smem &sm, const Lanes &g, const Entry &Cj, unsigned okbits, float &acc) {
    const uint8_t *colp = cur.p + Clamp(Cj.base, 0, cur.cols - 1);
    float top0 = 0.0f, top1 = 0.0f;
    if (REGULAR) {
        const uint8_t *p = colp + Clamp(sm.rows[0].base, 0, cur.rows) * cur.pitch;
        top0 = LoadPx(p);
        top1 = LoadPx(p + 1);
    }
#pragma unroll 15
    for (int r = 0; r < PR; ++r) {
        const Entry Rj = sm.rows[r];
        float v5;
        if (REGULAR) {
            const uint8_t *p = colp + Rj.off;  // row base + 1, clamped (precomputed)
            const float bot0 = LoadPx(p), bot1 = LoadPx(p + 1);
            v5 = Bilerp(Rj, Cj, top0, top1, bot0, bot1);
            top0 = bot0;
            top1 = bot1;
        } else {
            v5 = SampleDirect(cur, Rj, Cj);
        }
        const bool ok = (okbits >> r) & 1u;
        const float ft = fsub(v5, sm.iref[r][g.lane]);
        sm.term[0 * (kG + 4) + g.lane] = ok ? fmul(sm.fx[r][g.lane], ft) : 0.0f;
        sm.term[1 * (kG + 4) + g.lane] = ok ? fmul(sm.fy[r][g.lane], ft) : 0.0f;
        Fold<2, true>(sm.term, g, acc);
    }
}

template <int PR, int PC>
__global__ void __launch_bounds__(kThreads) BasicInverseFastKernel(KltLaunch a) {
    static_assert(PC <= kG && PR <= kG, "patch must fit one 16-lane group (lane = column; lane r also builds row r's table entry)");
    constexpr int HR = PR / 2, HC = PC / 2;
    __shared__ GroupSmem<PR, PC> smem_all[kGroupsPerBlock];

    Lanes g;
    g.lane = threadIdx.x & (kG - 1);
    g.base = (threadIdx.x & 31) - g.lane;
    g.mask = 0xFFFFu << g.base;
    const int group_in_block = threadIdx.x / kG;
    const int f_raw = blockIdx.x * kGroupsPerBlock + group_in_block;
    const bool exists = f_raw < a.n_features;
    const int f = exists ? f_raw : a.n_features - 1;  // idle groups shadow the last feature and write nothing
    GroupSmem<PR, PC> &sm = smem_all[group_in_block];
    float *term = sm.term;

    const int lane = g.lane;
    const bool col_active = lane < PC;
    const float dcol = static_cast<float>(lane - HC);
    constexpr unsigned kRowMask = (1u << PR) - 1u;

    const int pair = a.feat_pair ? a.feat_pair[f] : 0;  // null: a single frame pair
    const int local = f - a.feat_offsets[pair];
    const float2 ref_uv = a.ref_uv[f];
    float2 cur_uv = a.has_prediction ? a.cur_uv[f] : ref_uv;
    uint8_t status = a.has_status ? a.status[f] : static_cast<uint8_t>(FTK_STATUS_NOT_TRACKED);
    // basic_klt.cpp:9,12,15: only the first kMaxTrackPointsNumber features, never re-track failed ones
    const bool tracked = exists && static_cast<uint32_t>(local) < a.p.max_track_points && status <= FTK_STATUS_TRACKED;

    if (__any_sync(kFull, tracked)) {
        const int ref_image = a.ref_image ? a.ref_image[pair] : pair;
        const int cur_image = a.cur_image ? a.cur_image[pair] : pair;
        const int levels = a.single_level ? 1 : a.ref.levels;  // TrackSingleLevel (basic_klt.cpp:59-86) == one level, scale 1
        const float scale = static_cast<float>(1 << (levels - 1));
        float ref_x = fdiv(ref_uv.x, scale), ref_y = fdiv(ref_uv.y, scale);
        float cur_x = fdiv(cur_uv.x, scale), cur_y = fdiv(cur_uv.y, scale);

        for (int level = levels - 1; level > -1; --level) {
            const Img ref = LevelImage(a.ref, ref_image, level), cur = LevelImage(a.cur, cur_image, level);

            // ================= per-level setup: reference samples, gradients, full-mask Hessian =================
            unsigned refmask = 0;  // bit r: all five reference samples of my pixel in patch row r are inside the image
            float acc = 0.0f;
            {
                const float col_i = fadd(dcol, ref_x);
                bool c0_ok, cm_ok, cp_ok;
                const Entry C0 = MakeEntry(col_i, ref.cols, &c0_ok);
                const Entry Cm = MakeEntry(fsub(col_i, 1.0f), ref.cols, &cm_ok);
                const Entry Cp = MakeEntry(fadd(col_i, 1.0f), ref.cols, &cp_ok);
                // lane r builds the three row entries of patch row r
                bool rows_ok = false, regular = Cm.base == C0.base - 1 && Cp.base == C0.base + 1 && ref.cols >= 4 && ref.rows >= 4;
                int my_base = 0;
                __syncwarp();
                if (lane < PR) {
                    const float row_i = fadd(static_cast<float>(lane - HR), ref_y);
                    bool r0_ok, rm_ok, rp_ok;
                    Entry R0 = MakeEntry(row_i, ref.rows, &r0_ok);
                    const Entry Rm = MakeEntry(fsub(row_i, 1.0f), ref.rows, &rm_ok);
                    const Entry Rp = MakeEntry(fadd(row_i, 1.0f), ref.rows, &rp_ok);
                    R0.off = Clamp(Rp.base + 1, 0, ref.rows) * ref.pitch;  // the strip's new bottom row while processing this patch row
                    sm.rows[3 * lane + 0] = R0;
                    sm.rows[3 * lane + 1] = Rm;
                    sm.rows[3 * lane + 2] = Rp;
                    rows_ok = r0_ok && rm_ok && rp_ok;
                    regular = regular && Rm.base == R0.base - 1 && Rp.base == R0.base + 1;
                    my_base = R0.base;
                }
                // Do the integer bases advance regularly from one patch row to the next?
                const int next_base = __shfl_down_sync(kFull, my_base, 1);
                if (lane + 1 < PR) regular = regular && next_base == my_base + 1;
                __syncwarp();
                const unsigned rows_ok_bits = g.bits(rows_ok) & kRowMask;
                refmask = (c0_ok && cm_ok && cp_ok && col_active) ? rows_ok_bits : 0u;
                if (__all_sync(kFull, regular)) SetupRows<PR, true>(ref, sm, g, C0, Cm, Cp, refmask, acc);
                else SetupRows<PR, false>(ref, sm, g, C0, Cm, Cp, refmask, acc);
            }
            const float hfull00 = g.get(acc, 0), hfull01 = g.get(acc, 1), hfull11 = g.get(acc, 2);

            // ================= Gauss-Newton iterations (basic_klt.cpp:88-116) =================
            bool running = tracked;
            for (uint32_t iter = 0; iter < a.p.max_iteration && __any_sync(kFull, running); ++iter) {
                bool cj_ok;
                const Entry Cj = MakeEntry(fadd(dcol, cur_x), cur.cols, &cj_ok);
                bool row_ok = false, regular = true;
                int my_base = 0;
                __syncwarp();
                if (lane < PR) {
                    Entry Rj = MakeEntry(fadd(static_cast<float>(lane - HR), cur_y), cur.rows, &row_ok);
                    Rj.off = Clamp(Rj.base + 1, 0, cur.rows) * cur.pitch;
                    sm.rows[lane] = Rj;
                    my_base = Rj.base;
                }
                const int next_base = __shfl_down_sync(kFull, my_base, 1);
                if (lane + 1 < PR) regular = next_base == my_base + 1;
                __syncwarp();
                // pixels that pass all six bounds tests this iteration; their count is the reference's num_of_valid_pixel
                const unsigned rows_ok_bits = g.bits(row_ok);  // every lane votes (full-warp ballot)
                const unsigned okbits = (cj_ok && col_active) ? (refmask & rows_ok_bits) : 0u;
                const int valid = g.sum(__popc(okbits));

                acc = 0.0f;
                if (__all_sync(kFull, regular)) IterateRows<PR, true>(cur, sm, g, Cj, okbits, acc);
                else IterateRows<PR, false>(cur, sm, g, Cj, okbits, acc);
                const float b[2] = {g.get(acc, 0), g.get(acc, 1)};

                float h00 = hfull00, h01 = hfull01, h11 = hfull11;
                const bool mask_changed = g.any(okbits != refmask);
                if (__any_sync(kFull, mask_changed && running)) {
                    // Some reference-valid pixel left the current image: this iteration's Hessian runs over fewer pixels.
                    acc = 0.0f;
#pragma unroll 1
                    for (int r = 0; r < PR; ++r) {
                        const bool ok = (okbits >> r) & 1u;
                        const float fx = sm.fx[r][lane], fy = sm.fy[r][lane];
                        term[0 * (kG + 4) + lane] = ok ? fmul(fx, fx) : 0.0f;
                        term[1 * (kG + 4) + lane] = ok ? fmul(fx, fy) : 0.0f;
                        term[2 * (kG + 4) + lane] = ok ? fmul(fy, fy) : 0.0f;
                        Fold<3, false>(term, g, acc);
                    }
                    const float n00 = g.get(acc, 0), n01 = g.get(acc, 1), n11 = g.get(acc, 2);
                    if (mask_changed) h00 = n00, h01 = n01, h11 = n11;
                }

                if (running) {
                    if (valid == 0) {
                        running = false;  // BREAK_IF(ConstructIncrementalFunction(...) == 0)
                    } else {
                        const float A[2][2] = {{h00, h01}, {h01, h11}};
                        float v[2];
                        LdltSolve<2>(A, b, v);
                        if (v[0] != v[0] || v[1] != v[1]) {
                            status = FTK_STATUS_NUMERIC_ERROR;
                            running = false;
                        } else {
                            cur_x = fadd(cur_x, v[0]);
                            cur_y = fadd(cur_y, v[1]);
                            if (IsOutside(cur, cur_x, cur_y)) {
                                status = FTK_STATUS_OUTSIDE;
                                running = false;
                            } else if (fadd(fmul(v[0], v[0]), fmul(v[1], v[1])) < a.p.max_converge_step) {
                                status = FTK_STATUS_TRACKED;
                                running = false;
                            }
                        }
                    }
                }
            }

            if (level == 0) break;
            ref_x = fmul(ref_x, 2.0f), ref_y = fmul(ref_y, 2.0f);
            cur_x = fmul(cur_x, 2.0f), cur_y = fmul(cur_y, 2.0f);
        }
        if (tracked) {
            cur_uv = make_float2(cur_x, cur_y);
            const Img cur0 = LevelImage(a.cur, cur_image, 0);
            if (IsOutside(cur0, cur_uv.x, cur_uv.y)) status = FTK_STATUS_OUTSIDE;  // basic_klt.cpp:49-53
        }
    }
    if (exists && g.lane == 0) {
        a.cur_uv[f] = cur_uv;
        a.status[f] = status;
    }
}

// =====================================================================================================================
// kDirect method (basic_klt.cpp:151-177): the gradient comes from the CURRENT image at the moving position, so the five-sample
// pass of SetupRows runs every iteration on `cur`; only the reference centre sample (and its bounds bit) is per-level work.
// =====================================================================================================================

// Reference centre sample of every patch pixel -> sm.iref (per level).
template <int PR, bool REGULAR, typename Smem>
__device__ __forceinline__ void RefCentreRows(const Img &ref, Smem &sm, const Lanes &g, const Entry &C0) {
    const uint8_t *colp = ref.p + Clamp(C0.base, 0, ref.cols - 1);
    float top0 = 0.0f, top1 = 0.0f;
    if (REGULAR) {
        const uint8_t *p = colp + Clamp(sm.rows[0].base, 0, ref.rows) * ref.pitch;
        top0 = LoadPx(p);
        top1 = LoadPx(p + 1);
    }
#pragma unroll 5
    for (int r = 0; r < PR; ++r) {
        const Entry R0 = sm.rows[r];
        float v;
        if (REGULAR) {
            const uint8_t *p = colp + R0.off;
            const float bot0 = LoadPx(p), bot1 = LoadPx(p + 1);
            v = Bilerp(R0, C0, top0, top1, bot0, bot1);
            top0 = bot0;
            top1 = bot1;
        } else {
            v = SampleDirect(ref, R0, C0);
        }
        sm.iref[r][g.lane] = v;
    }
}

// One iteration: five current-image samples per pixel, 3 Hessian + 2 bias chains (bias terms stored negated).
template <int PR, bool REGULAR, typename Smem>
__device__ __forceinline__ void DirectRows(const Img &cur, Smem &sm, const Lanes &g, const Entry &C0, const Entry &Cm, const Entry &Cp, unsigned okbits,
                                           float &acc) {
    float s0[4], s1[4], s2[4], s3[4];
    const uint8_t *colp = cur.p + Clamp(Cm.base, 0, cur.cols - 3);
    if (REGULAR) {
        const int rr = sm.rows[1].base;
        const uint8_t *p = colp + Clamp(rr, 0, cur.rows) * cur.pitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) s0[i] = LoadPx(p + i);
        p = colp + Clamp(rr + 1, 0, cur.rows) * cur.pitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) s1[i] = LoadPx(p + i);
        p = colp + Clamp(rr + 2, 0, cur.rows) * cur.pitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) s2[i] = LoadPx(p + i);
    }
#pragma unroll 3
    for (int r = 0; r < PR; ++r) {
        const Entry R0 = sm.rows[3 * r], Rm = sm.rows[3 * r + 1], Rp = sm.rows[3 * r + 2];
        float v0, v1, v2, v3, v5;
        if (REGULAR) {
            const uint8_t *p = colp + R0.off;
#pragma unroll
            for (int i = 0; i < 4; ++i) s3[i] = LoadPx(p + i);
            v0 = Bilerp(R0, Cm, s1[0], s1[1], s2[0], s2[1]);
            v1 = Bilerp(R0, Cp, s1[2], s1[3], s2[2], s2[3]);
            v2 = Bilerp(Rm, C0, s0[1], s0[2], s1[1], s1[2]);
            v3 = Bilerp(Rp, C0, s2[1], s2[2], s3[1], s3[2]);
            v5 = Bilerp(R0, C0, s1[1], s1[2], s2[1], s2[2]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                s0[i] = s1[i];
                s1[i] = s2[i];
                s2[i] = s3[i];
            }
        } else {
            v0 = SampleDirect(cur, R0, Cm);
            v1 = SampleDirect(cur, R0, Cp);
            v2 = SampleDirect(cur, Rm, C0);
            v3 = SampleDirect(cur, Rp, C0);
            v5 = SampleDirect(cur, R0, C0);
        }
        const bool ok = (okbits >> r) & 1u;
        const float fx = fsub(v1, v0), fy = fsub(v3, v2), ft = fsub(v5, sm.iref[r][g.lane]);
        sm.term[0 * (kG + 4) + g.lane] = ok ? fmul(fx, fx) : 0.0f;
        sm.term[1 * (kG + 4) + g.lane] = ok ? fmul(fx, fy) : 0.0f;
        sm.term[2 * (kG + 4) + g.lane] = ok ? fmul(fy, fy) : 0.0f;
        sm.term[3 * (kG + 4) + g.lane] = ok ? -fmul(fx, ft) : 0.0f;
        sm.term[4 * (kG + 4) + g.lane] = ok ? -fmul(fy, ft) : 0.0f;
        Fold<5, false>(sm.term, g, acc);
    }
}

template <int PR, int PC>
__global__ void __launch_bounds__(kThreads) BasicDirectFastKernel(KltLaunch a) {
    static_assert(PC <= kG && PR <= kG, "patch must fit one 16-lane group");
    constexpr int HR = PR / 2, HC = PC / 2;
    __shared__ GroupSmem<PR, PC> smem_all[kGroupsPerBlock];

    Lanes g;
    g.lane = threadIdx.x & (kG - 1);
    g.base = (threadIdx.x & 31) - g.lane;
    g.mask = 0xFFFFu << g.base;
    const int group_in_block = threadIdx.x / kG;
    const int f_raw = blockIdx.x * kGroupsPerBlock + group_in_block;
    const bool exists = f_raw < a.n_features;
    const int f = exists ? f_raw : a.n_features - 1;
    GroupSmem<PR, PC> &sm = smem_all[group_in_block];
    const int lane = g.lane;
    const bool col_active = lane < PC;
    const float dcol = static_cast<float>(lane - HC);
    constexpr unsigned kRowMask = (1u << PR) - 1u;

    const int pair = a.feat_pair ? a.feat_pair[f] : 0;  // null: a single frame pair
    const int local = f - a.feat_offsets[pair];
    const float2 ref_uv = a.ref_uv[f];
    float2 cur_uv = a.has_prediction ? a.cur_uv[f] : ref_uv;
    uint8_t status = a.has_status ? a.status[f] : static_cast<uint8_t>(FTK_STATUS_NOT_TRACKED);
    const bool tracked = exists && static_cast<uint32_t>(local) < a.p.max_track_points && status <= FTK_STATUS_TRACKED;

    if (__any_sync(kFull, tracked)) {
        const int ref_image = a.ref_image ? a.ref_image[pair] : pair;
        const int cur_image = a.cur_image ? a.cur_image[pair] : pair;
        const int levels = a.single_level ? 1 : a.ref.levels;
        const float scale = static_cast<float>(1 << (levels - 1));
        float ref_x = fdiv(ref_uv.x, scale), ref_y = fdiv(ref_uv.y, scale);
        float cur_x = fdiv(cur_uv.x, scale), cur_y = fdiv(cur_uv.y, scale);

        for (int level = levels - 1; level > -1; --level) {
            const Img ref = LevelImage(a.ref, ref_image, level), cur = LevelImage(a.cur, cur_image, level);

            // ---- per level: reference centre samples and their bounds bits ----
            unsigned ref_rows_ok;
            bool ref_col_ok;
            {
                const Entry C0 = MakeEntry(fadd(dcol, ref_x), ref.cols, &ref_col_ok);
                bool row_ok = false, regular = true;
                int my_base = 0;
                __syncwarp();
                if (lane < PR) {
                    Entry R0 = MakeEntry(fadd(static_cast<float>(lane - HR), ref_y), ref.rows, &row_ok);
                    R0.off = Clamp(R0.base + 1, 0, ref.rows) * ref.pitch;
                    sm.rows[lane] = R0;
                    my_base = R0.base;
                }
                const int next_base = __shfl_down_sync(kFull, my_base, 1);
                if (lane + 1 < PR) regular = next_base == my_base + 1;
                __syncwarp();
                ref_rows_ok = g.bits(row_ok) & kRowMask;
                if (__all_sync(kFull, regular)) RefCentreRows<PR, true>(ref, sm, g, C0);
                else RefCentreRows<PR, false>(ref, sm, g, C0);
            }

            // ---- Gauss-Newton iterations (basic_klt.cpp:88-116 with the kDirect branch of :151-177) ----
            bool running = tracked;
            for (uint32_t iter = 0; iter < a.p.max_iteration && __any_sync(kFull, running); ++iter) {
                const float col_j = fadd(dcol, cur_x);
                bool c0_ok, cm_ok, cp_ok;
                const Entry C0 = MakeEntry(col_j, cur.cols, &c0_ok);
                const Entry Cm = MakeEntry(fsub(col_j, 1.0f), cur.cols, &cm_ok);
                const Entry Cp = MakeEntry(fadd(col_j, 1.0f), cur.cols, &cp_ok);
                bool rows_ok = false, regular = Cm.base == C0.base - 1 && Cp.base == C0.base + 1 && cur.cols >= 4 && cur.rows >= 4;
                int my_base = 0;
                __syncwarp();
                if (lane < PR) {
                    const float row_j = fadd(static_cast<float>(lane - HR), cur_y);
                    bool r0_ok, rm_ok, rp_ok;
                    Entry R0 = MakeEntry(row_j, cur.rows, &r0_ok);
                    const Entry Rm = MakeEntry(fsub(row_j, 1.0f), cur.rows, &rm_ok);
                    const Entry Rp = MakeEntry(fadd(row_j, 1.0f), cur.rows, &rp_ok);
                    R0.off = Clamp(Rp.base + 1, 0, cur.rows) * cur.pitch;
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
                const unsigned row_bits = g.bits(rows_ok) & ref_rows_ok;
                const bool lane_gate = ref_col_ok && c0_ok && cm_ok && cp_ok && col_active;
                const unsigned okbits = lane_gate ? row_bits : 0u;
                const int valid = __popc(row_bits) * __popc(g.bits(lane_gate));  // the six bounds tests are separable in row and column

                float acc = 0.0f;
                if (__all_sync(kFull, regular)) DirectRows<PR, true>(cur, sm, g, C0, Cm, Cp, okbits, acc);
                else DirectRows<PR, false>(cur, sm, g, C0, Cm, Cp, okbits, acc);
                const float h00 = g.get(acc, 0), h01 = g.get(acc, 1), h11 = g.get(acc, 2);
                const float b[2] = {g.get(acc, 3), g.get(acc, 4)};

                if (running) {
                    if (valid == 0) {
                        running = false;
                    } else {
                        const float A[2][2] = {{h00, h01}, {h01, h11}};
                        float v[2];
                        LdltSolve<2>(A, b, v);
                        if (v[0] != v[0] || v[1] != v[1]) {
                            status = FTK_STATUS_NUMERIC_ERROR;
                            running = false;
                        } else {
                            cur_x = fadd(cur_x, v[0]);
                            cur_y = fadd(cur_y, v[1]);
                            if (IsOutside(cur, cur_x, cur_y)) {
                                status = FTK_STATUS_OUTSIDE;
                                running = false;
                            } else if (fadd(fmul(v[0], v[0]), fmul(v[1], v[1])) < a.p.max_converge_step) {
                                status = FTK_STATUS_TRACKED;
                                running = false;
                            }
                        }
                    }
                }
            }

            if (level == 0) break;
            ref_x = fmul(ref_x, 2.0f), ref_y = fmul(ref_y, 2.0f);
            cur_x = fmul(cur_x, 2.0f), cur_y = fmul(cur_y, 2.0f);
        }
        if (tracked) {
            cur_uv = make_float2(cur_x, cur_y);
            const Img cur0 = LevelImage(a.cur, cur_image, 0);
            if (IsOutside(cur0, cur_uv.x, cur_uv.y)) status = FTK_STATUS_OUTSIDE;
        }
    }
    if (exists && g.lane == 0) {
        a.cur_uv[f] = cur_uv;
        a.status[f] = status;
    }
}

template <int PR, int PC>
int LaunchDirect(ftk_context *ctx, const KltLaunch &a) {
    const int blocks = (a.n_features + kGroupsPerBlock - 1) / kGroupsPerBlock;
    BasicDirectFastKernel<PR, PC><<<blocks, kThreads, 0, ctx->stream>>>(a);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

// =====================================================================================================================
// kFast method (basic_klt_fast.cpp:7-195 + optical_flow.cpp:49-102), patches up to 14 columns wide (the reference's default
// 13x13 instantiated).  The reference's fast method is integer-aligned by construction: the extended reference patch
// ((2h+3)^2 samples) and every iteration's current patch use ONE set of bilinear weights over a window at floor(position),
// and a sample is valid iff 0 <= row <= rows-2 && 0 <= col <= cols-2 -- separable in row and column.  Mapping: lane = column of
// the extended patch (15 of 16 lanes for 13x13), two features per warp, rolling 2-byte strips, row validity as bit masks,
// gradients from the neighbouring lanes (shuffles) and the previous / next extended rows (registers).
// =====================================================================================================================
template <int PR, int PC>
struct FastSmem {
    static constexpr int ER = PR + 2;
    float term[5 * (kG + 4)];
    float ex[ER][kG];            // extended reference patch, [row][lane]
    float dx[PR][kG], dy[PR][kG];  // reference gradients of patch pixel (row, lane - 1)
    static constexpr int kRawWords = 5 * (kG + 4) + ER * kG + 2 * PR * kG;
    static constexpr int kPad = ((16 - kRawWords % 32) + 32) % 32;
    float pad[kPad == 0 ? 32 : kPad];
};

// One weight set for a whole integer-aligned window (optical_flow.cpp:53-60, basic_klt_fast.cpp:109-116).
struct WindowWeights {
    float tl, tr, bl, br;
    int int_row, int_col;
};
__device__ __forceinline__ WindowWeights MakeWindowWeights(float x, float y) {
    WindowWeights w;
    const float fr = floorf(y), fc = floorf(x);
    const float dr = fsub(y, fr), dc = fsub(x, fc);
    w.tl = fmul(fsub(1.0f, dr), fsub(1.0f, dc));
    w.tr = fmul(fsub(1.0f, dr), dc);
    w.bl = fmul(dr, fsub(1.0f, dc));
    w.br = fmul(dr, dc);
    w.int_row = min(max(static_cast<int>(fr), -(1 << 24)), 1 << 24);
    w.int_col = min(max(static_cast<int>(fc), -(1 << 24)), 1 << 24);
    return w;
}
__device__ __forceinline__ float WindowSample(const WindowWeights &w, float p00, float p01, float p10, float p11) {
    return fadd(fadd(fadd(fmul(w.tl, p00), fmul(w.tr, p01)), fmul(w.bl, p10)), fmul(w.br, p11));
}
// bit r = (0 <= first + r <= limit) for r in [0, n)
__device__ __forceinline__ unsigned RangeBits(int first, int n, int limit) {
    unsigned bits = 0;
#pragma unroll 1
    for (int r = 0; r < n; ++r) bits |= (first + r >= 0 && first + r <= limit) ? (1u << r) : 0u;
    return bits;
}

template <int PR, int PC>
__global__ void __launch_bounds__(kThreads) BasicFastMethodKernel(KltLaunch a) {
    constexpr int ER = PR + 2, EC = PC + 2;
    static_assert(EC <= kG && ER <= 32, "extended patch must fit one 16-lane group");
    __shared__ FastSmem<PR, PC> smem_all[kGroupsPerBlock];

    Lanes g;
    g.lane = threadIdx.x & (kG - 1);
    g.base = (threadIdx.x & 31) - g.lane;
    g.mask = 0xFFFFu << g.base;
    const int group_in_block = threadIdx.x / kG;
    const int f_raw = blockIdx.x * kGroupsPerBlock + group_in_block;
    const bool exists = f_raw < a.n_features;
    const int f = exists ? f_raw : a.n_features - 1;
    FastSmem<PR, PC> &sm = smem_all[group_in_block];
    float *term = sm.term;
    const int lane = g.lane;

    const int pair = a.feat_pair ? a.feat_pair[f] : 0;  // null: a single frame pair
    const int local = f - a.feat_offsets[pair];
    const float2 ref_uv = a.ref_uv[f];
    float2 cur_uv = a.has_prediction ? a.cur_uv[f] : ref_uv;
    uint8_t status = a.has_status ? a.status[f] : static_cast<uint8_t>(FTK_STATUS_NOT_TRACKED);
    const bool tracked = exists && static_cast<uint32_t>(local) < a.p.max_track_points && status <= FTK_STATUS_TRACKED;

    if (__any_sync(kFull, tracked)) {
        const int ref_image = a.ref_image ? a.ref_image[pair] : pair;
        const int cur_image = a.cur_image ? a.cur_image[pair] : pair;
        const int levels = a.single_level ? 1 : a.ref.levels;
        const float scale = static_cast<float>(1 << (levels - 1));
        float ref_x = fdiv(ref_uv.x, scale), ref_y = fdiv(ref_uv.y, scale);
        float cur_x = fdiv(cur_uv.x, scale), cur_y = fdiv(cur_uv.y, scale);

        for (int level = levels - 1; level > -1; --level) {
            const Img ref = LevelImage(a.ref, ref_image, level), cur = LevelImage(a.cur, cur_image, level);

            // ============ ExtractExtendPatchInReferenceImage + PrecomputeJacobianAndHessian ============
            const WindowWeights wr = MakeWindowWeights(ref_x, ref_y);
            const int ex_row0 = wr.int_row - ER / 2, ex_col0 = wr.int_col - EC / 2;
            const unsigned ex_rows_ok = RangeBits(ex_row0, ER, ref.rows - 2);                   // bit er
            const int my_col = ex_col0 + lane;
            const bool ex_col_ok = lane < EC && my_col >= 0 && my_col <= ref.cols - 2;
            const unsigned ex_cols_ok = g.bits(ex_col_ok);                                       // bit lane
            const int ex_valid = __popc(ex_rows_ok) * __popc(ex_cols_ok);                        // valid_pixel_num
            // gradient of patch column (lane - 1) needs the extended columns lane - 1, lane, lane + 1
            const bool grad_cols_ok = lane >= 1 && lane <= PC && ((ex_cols_ok >> (lane - 1)) & 7u) == 7u;

            float acc = 0.0f;
            {
                const uint8_t *colp = ref.p + Clamp(my_col, 0, ref.cols - 1);
                const uint8_t *p = colp + Clamp(ex_row0, 0, ref.rows) * ref.pitch;
                float top0 = LoadPx(p), top1 = LoadPx(p + 1);
                float e_prev2 = 0.0f, e_prev1 = 0.0f;  // extended rows er - 2, er - 1 of my column
#pragma unroll 3
                for (int er = 0; er < ER; ++er) {
                    p = colp + Clamp(ex_row0 + er + 1, 0, ref.rows) * ref.pitch;
                    const float bot0 = LoadPx(p), bot1 = LoadPx(p + 1);
                    const bool ok = ex_col_ok && ((ex_rows_ok >> er) & 1u);
                    const float e = ok ? WindowSample(wr, top0, top1, bot0, bot1) : 0.0f;
                    top0 = bot0;
                    top1 = bot1;
                    sm.ex[er][lane] = e;
                    // horizontal neighbours of the PREVIOUS extended row (the centre row of patch row er - 2)
                    const float left = __shfl_sync(kFull, e_prev1, (threadIdx.x & 31) - 1);
                    const float right = __shfl_sync(kFull, e_prev1, (threadIdx.x & 31) + 1);
                    if (er >= 2) {
                        const int pr = er - 2;
                        const bool gok = grad_cols_ok && ((ex_rows_ok >> pr) & 7u) == 7u;
                        const float dx = gok ? fsub(right, left) : 0.0f;
                        const float dy = gok ? fsub(e, e_prev2) : 0.0f;
                        sm.dx[pr][lane] = dx;
                        sm.dy[pr][lane] = dy;
                        term[0 * (kG + 4) + lane] = gok ? fmul(dx, dx) : 0.0f;
                        term[1 * (kG + 4) + lane] = gok ? fmul(dx, dy) : 0.0f;
                        term[2 * (kG + 4) + lane] = gok ? fmul(dy, dy) : 0.0f;
                        Fold<3, false>(term, g, acc);
                    }
                    e_prev2 = e_prev1;
                    e_prev1 = e;
                }
            }
            const float h00 = g.get(acc, 0), h01 = g.get(acc, 1), h11 = g.get(acc, 2);
            const float A[2][2] = {{h00, h01}, {h01, h11}};
            LdltFactors<2> factors;
            LdltFactor<2>(A, factors);

            // ============ iterations (basic_klt_fast.cpp:29-61) ============
            bool running = tracked;
            if (running && ex_valid == 0) {  // :16-19
                status = FTK_STATUS_OUTSIDE;
                running = false;
            }
            if (running) status = FTK_STATUS_LARGE_RESIDUAL;
            float last_squared_step = INFINITY;
            uint32_t large_step_cnt = 0;
            for (uint32_t iter = 0; iter < a.p.max_iteration && __any_sync(kFull, running); ++iter) {
                // ComputeBias: integer window at floor(cur), one weight set; my patch column is lane - 1
                const WindowWeights wc = MakeWindowWeights(cur_x, cur_y);
                const int row0 = wc.int_row - PR / 2, col = wc.int_col - PC / 2 + lane - 1;
                const unsigned rows_ok = RangeBits(row0, PR, cur.rows - 2) & (ex_rows_ok >> 1);   // cur row valid && centre row of ex valid
                const bool col_ok = lane >= 1 && lane <= PC && col >= 0 && col <= cur.cols - 2 && ex_col_ok;
                const int valid = __popc(rows_ok) * __popc(g.bits(col_ok));
                const uint8_t *colp = cur.p + Clamp(col, 0, cur.cols - 1);
                const uint8_t *p = colp + Clamp(row0, 0, cur.rows) * cur.pitch;
                float top0 = LoadPx(p), top1 = LoadPx(p + 1);
                acc = 0.0f;
#pragma unroll
                for (int pr = 0; pr < PR; ++pr) {
                    p = colp + Clamp(row0 + pr + 1, 0, cur.rows) * cur.pitch;
                    const float bot0 = LoadPx(p), bot1 = LoadPx(p + 1);
                    const float cur_value = WindowSample(wc, top0, top1, bot0, bot1);
                    top0 = bot0;
                    top1 = bot1;
                    const bool ok = col_ok && ((rows_ok >> pr) & 1u);
                    const float dt = fsub(cur_value, sm.ex[pr + 1][lane]);
                    term[0 * (kG + 4) + lane] = ok ? fmul(sm.dx[pr][lane], dt) : 0.0f;
                    term[1 * (kG + 4) + lane] = ok ? fmul(sm.dy[pr][lane], dt) : 0.0f;
                    Fold<2, true>(term, g, acc);
                }
                const float b[2] = {g.get(acc, 0), g.get(acc, 1)};
                if (running) {
                    if (valid == 0) {
                        running = false;  // BREAK_IF(ComputeBias(...) == 0)
                    } else {
                        float v[2];
                        LdltSolveFactored<2>(factors, b, v);
                        if (v[0] != v[0] || v[1] != v[1]) {
                            status = FTK_STATUS_NUMERIC_ERROR;
                            running = false;
                        } else {
                            cur_x = fadd(cur_x, v[0]);
                            cur_y = fadd(cur_y, v[1]);
                            const float squared_step = fadd(fmul(v[0], v[0]), fmul(v[1], v[1]));
                            if (squared_step < last_squared_step) {
                                last_squared_step = squared_step;
                                large_step_cnt = 0;
                            } else {
                                ++large_step_cnt;
                                if (large_step_cnt >= a.p.max_tolerance_large_step) running = false;
                            }
                            if (running && squared_step < a.p.max_converge_step) {
                                status = FTK_STATUS_TRACKED;
                                running = false;
                            }
                        }
                    }
                }
            }

            if (level == 0) break;
            ref_x = fmul(ref_x, 2.0f), ref_y = fmul(ref_y, 2.0f);
            cur_x = fmul(cur_x, 2.0f), cur_y = fmul(cur_y, 2.0f);
        }
        if (tracked) {
            cur_uv = make_float2(cur_x, cur_y);
            const Img cur0 = LevelImage(a.cur, cur_image, 0);
            if (IsOutside(cur0, cur_uv.x, cur_uv.y)) status = FTK_STATUS_OUTSIDE;  // basic_klt.cpp:49-53
        }
    }
    if (exists && g.lane == 0) {
        a.cur_uv[f] = cur_uv;
        a.status[f] = status;
    }
}

template <int PR, int PC>
int LaunchFastMethod(ftk_context *ctx, const KltLaunch &a) {
    const int blocks = (a.n_features + kGroupsPerBlock - 1) / kGroupsPerBlock;
    BasicFastMethodKernel<PR, PC><<<blocks, kThreads, 0, ctx->stream>>>(a);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

template <int PR, int PC>
int Launch(ftk_context *ctx, const KltLaunch &a) {
    const int blocks = (a.n_features + kGroupsPerBlock - 1) / kGroupsPerBlock;
    BasicInverseFastKernel<PR, PC><<<blocks, kThreads, 0, ctx->stream>>>(a);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

}  // namespace

// Returns FTK_ERR_UNSUPPORTED when no specialisation covers the configuration (the caller then uses the generic kernel).
int LaunchKltBasicFastPath(ftk_context *ctx, const KltLaunch &a) {
    if (a.p.variant != FTK_VARIANT_BASIC) return FTK_ERR_UNSUPPORTED;
    if (a.p.method >= FTK_METHOD_FAST || a.p.method < 0) {  // kFast, and kSse / kNeon which take the reference's `default:` branch
        if (a.p.patch_row_half == 6 && a.p.patch_col_half == 6) return LaunchFastMethod<13, 13>(ctx, a);
        if (a.p.patch_row_half == 5 && a.p.patch_col_half == 5) return LaunchFastMethod<11, 11>(ctx, a);
        if (a.p.patch_row_half == 4 && a.p.patch_col_half == 4) return LaunchFastMethod<9, 9>(ctx, a);
        return FTK_ERR_UNSUPPORTED;
    }
    if (a.p.method == FTK_METHOD_DIRECT) {
        if (a.p.patch_row_half == 7 && a.p.patch_col_half == 7) return LaunchDirect<15, 15>(ctx, a);
        if (a.p.patch_row_half == 6 && a.p.patch_col_half == 6) return LaunchDirect<13, 13>(ctx, a);
        return FTK_ERR_UNSUPPORTED;
    }
    if (a.p.patch_row_half == 7 && a.p.patch_col_half == 7) return Launch<15, 15>(ctx, a);
    if (a.p.patch_row_half == 6 && a.p.patch_col_half == 6) return Launch<13, 13>(ctx, a);
    return FTK_ERR_UNSUPPORTED;
}

}  // namespace ftk
