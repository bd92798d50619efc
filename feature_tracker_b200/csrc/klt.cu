// K2-K4: pyramidal sparse Lucas-Kanade tracking on sm_100a -- basic (2-DoF), affine (6-DoF) and LSSD (SE(2) + mean
// normalisation) trackers, each with the reference's kInverse / kDirect / kFast methods.
//
// Replaces (paths relative to the reference's src/optical_flow_tracker/):
//   basic_klt/optical_flow_basic_klt.cpp:7-181, basic_klt/optical_flow_basic_klt_fast.cpp:7-195,
//   affine_klt/optical_flow_affine_klt.cpp:6-273, affine_klt/optical_flow_affine_klt_fast.cpp:7-188,
//   lssd_klt/optical_flow_lssd_klt.cpp:7-250, lssd_klt/optical_flow_lssd_klt_fast.cpp:7-229,
//   optical_flow.cpp:49-102 (ExtractExtendPatchInReferenceImage).
//
// Mapping: a group of G lanes (G = 8 for the 2-DoF tracker, 32 for affine / LSSD) owns one feature for its whole
// coarse-to-fine life; level and Gauss-Newton loops stay inside the kernel.  Patch pixels are processed in
// row-major chunks of G: phase 1 evaluates the per-pixel terms of the normal equations on G lanes in parallel
// (each term is a chain of correctly rounded fp32 operations identical to the reference's), phase 2 lets lane k
// fold the chunk's k-th term into accumulator k in pixel order.  The sums are therefore bit-identical to the
// sequential CPU loops while K accumulators x (32/G) features advance per warp instruction.
// This file is the generic ("literal") implementation for every variant/method and any patch size; the
// specialised kernels for the head-line configurations live in klt_basic_fastpath.cu.
#include "klt_device.cuh"

namespace ftk {

namespace {

constexpr int kInverse = FTK_METHOD_INVERSE;
constexpr int kDirect = FTK_METHOD_DIRECT;
constexpr int kFast = FTK_METHOD_FAST;

// Patch geometry (optical_flow.cpp:104-124 PrepareForTracking).
struct Geometry {
    int hr, hc;          // half sizes
    int pr, pc, psize;   // patch
    int er, ec, esize;   // extended patch (+1 border)
};

// Per-group shared-memory carve-up (sizes decided on the host, see GroupSmemBytes()).
struct Scratch {
    float *term;
    float *ex;         // extended ref patch (fast methods)
    float *dx, *dy;    // ref gradients (fast methods)
    float *curp;       // cur patch (lssd fast)
    float *hoist;      // lssd inverse: per-level reference gradients / samples + this iteration's cur samples: 4 x p_floats
    Ldlt6Shared *ldlt; // affine: normal equations + factors of the cooperative 6x6 LDLT
    LdltShared<3> *ldlt3;  // lssd: the same for its 3x3 system (aliases the same scratch region)
    uint8_t *exv;      // ex patch validity
    uint8_t *curv;     // cur patch validity (lssd fast)
};

struct SmemLayout {
    int term_floats, ex_floats, p_floats, curp_floats, hoist_floats, ldlt_floats, exv_bytes, curv_bytes, total_bytes;
};

__host__ __device__ inline int RoundUp(int v, int m) { return (v + m - 1) / m * m; }

__host__ __device__ inline SmemLayout MakeLayout(int variant, int method, int G, const Geometry &geo) {
    SmemLayout l{};
    int kmax = variant == FTK_VARIANT_BASIC ? 5 : (variant == FTK_VARIANT_AFFINE ? 27 : 9);
    if (variant == FTK_VARIANT_AFFINE && method == kFast) kmax = G == 16 ? 8 : 21;  // 21 Hessian chains per level (in three parts on 16 lanes), 6 per iteration
    l.term_floats = RoundUp(kmax * (G + 4), 4);
    // paired chain layout of the 16-lane kDirect / kInverse trackers (kFast folds 21 + 6 plain chains: the smaller buffer buys two more CTAs per SM)
    if (variant == FTK_VARIANT_AFFINE && G == 16 && method != kFast) l.term_floats = RoundUp(G * (2 * G + 4), 4) > l.term_floats ? RoundUp(G * (2 * G + 4), 4) : l.term_floats;
    if (method == kFast) {
        l.ex_floats = RoundUp(geo.esize, 4);
        l.p_floats = RoundUp(geo.psize, 4);
        l.curp_floats = variant == FTK_VARIANT_LSSD ? l.p_floats : 0;
        l.exv_bytes = RoundUp(geo.esize, 16);
        l.curv_bytes = variant == FTK_VARIANT_LSSD ? RoundUp(geo.psize, 16) : 0;
    }
    int pair_floats = 2 * l.p_floats;  // dx, dy of the fast methods
    if (variant == FTK_VARIANT_LSSD && method == kInverse) {
        pair_floats = 0;
        l.hoist_floats = 4 * RoundUp(geo.psize, 4);
    }
    if (variant == FTK_VARIANT_AFFINE && method != kFast) {
        pair_floats = 0;
        l.hoist_floats = (method == kInverse ? 3 : 1) * RoundUp(geo.psize, 4);  // per level: [fx, fy,] reference centre sample
    }
    l.ldlt_floats = variant == FTK_VARIANT_AFFINE ? RoundUp(kLdlt6Floats, 4) : (variant == FTK_VARIANT_LSSD ? RoundUp(kLdlt3Floats, 4) : 0);
    l.total_bytes = 4 * (l.term_floats + l.ex_floats + pair_floats + l.curp_floats + l.hoist_floats + l.ldlt_floats) + l.exv_bytes + l.curv_bytes;
    l.total_bytes = RoundUp(l.total_bytes, 16);
    // Groups of one warp must start on different banks: make the group stride (in words) congruent to G modulo 32.
    if (G < 32)
        while ((l.total_bytes / 4) % 32 != G % 32) l.total_bytes += 16;
    return l;
}

__device__ __forceinline__ Scratch CarveScratch(unsigned char *base, const SmemLayout &l) {
    Scratch s;
    float *f = reinterpret_cast<float *>(base);
    s.term = f;
    f += l.term_floats;
    s.ex = f;
    f += l.ex_floats;
    s.dx = f;
    f += l.p_floats;
    s.dy = f;
    f += l.p_floats;
    s.curp = f;
    f += l.curp_floats;
    s.hoist = f;
    f += l.hoist_floats;
    s.ldlt = reinterpret_cast<Ldlt6Shared *>(f);
    s.ldlt3 = reinterpret_cast<LdltShared<3> *>(f);
    f += l.ldlt_floats;
    uint8_t *b = reinterpret_cast<uint8_t *>(f);
    s.exv = b;
    b += l.exv_bytes;
    s.curv = b;
    return s;
}

// Position of a lane's pixel inside the patch while the row-major pixel index advances by G per chunk: replaces k / pc and
// k % pc (a ~20-instruction integer division each) in the per-iteration loops.
struct PatchWalk {
    int row, col;          // k / pc, k % pc for k = chunk * G + lane
    int step_row, step_col, pc;
    __device__ __forceinline__ void next() {
        col += step_col;
        row += step_row;
        if (col >= pc) col -= pc, ++row;
    }
};

// Everything a group needs while tracking one feature.
template <int G>
struct Ctx {
    PatchWalk walk;  // chunk 0 of this lane
    Group<G> g;
    Chain<G> ch;
    Scratch s;
    Geometry geo;
    const ftk_klt_params *p;
    int ldlt_dst0, ldlt_dst1;  // affine: float offsets inside Ldlt6Shared of chain `lane` and chain `G + lane` (-1: none)
};

// ---- shared pieces of the fast methods ---------------------------------------------------------------------------

// optical_flow.cpp:49-102 ExtractExtendPatchInReferenceImage: (2h+3)^2 integer-aligned window at floor(ref) with ONE
// set of bilinear weights; valid iff 0<=row<=rows-2 && 0<=col<=cols-2.  Returns the valid count (group-uniform).
template <int G>
__device__ int ExtractExRefPatch(Ctx<G> &c, const Img &ref, float ref_x, float ref_y) {
    const float int_row = floorf(ref_y), int_col = floorf(ref_x);
    const float dec_row = fsub(ref_y, int_row), dec_col = fsub(ref_x, int_col);
    const float w_tl = fmul(fsub(1.0f, dec_row), fsub(1.0f, dec_col));
    const float w_tr = fmul(fsub(1.0f, dec_row), dec_col);
    const float w_bl = fmul(dec_row, fsub(1.0f, dec_col));
    const float w_br = fmul(dec_row, dec_col);
    const int min_row = static_cast<int>(int_row) - c.geo.er / 2;
    const int min_col = static_cast<int>(int_col) - c.geo.ec / 2;
    int valid = 0;
    if (G >= 16 && c.geo.ec <= 16) {
        // Row-structured walk for the usual patch widths (the reference's default 13 + 2 = 15 columns): 16 lanes per extended row, G / 16
        // rows per step; a lane's column never changes, so a step costs one address increment, the row test and the four loads (the
        // generic walk below spends ~130 instructions per 32 samples, most of them index arithmetic).  Same values, same validity.
        constexpr int kRowsPerStep = G >= 16 ? G / 16 : 1;
        const int lcol = c.g.lane & 15, lrow = c.g.lane >> 4;
        const int col = min_col + lcol;
        const bool col_ok = lcol < c.geo.ec && !(col < 0 || col > ref.cols - 2);
        // running pointers: image row of this lane, its slot in the extended patch
        const uint8_t *p = ref.p + min(max(col, 0), ref.cols - 1) + static_cast<long long>(min_row + lrow) * ref.pitch;
        const long long p_step = static_cast<long long>(kRowsPerStep) * ref.pitch;
        float *ex = c.s.ex + lrow * c.geo.ec + lcol;
        uint8_t *exv = c.s.exv + lrow * c.geo.ec + lcol;
        const int e_step = kRowsPerStep * c.geo.ec;
        int row = min_row + lrow, er = lrow;
        for (int r0 = 0; r0 < c.geo.er; r0 += kRowsPerStep) {
            const bool in_patch = er < c.geo.er && lcol < c.geo.ec;
            const bool ok = in_patch && col_ok && !(row < 0 || row > ref.rows - 2);
            float v = 0.0f;
            if (ok)  // invalid lanes never touch memory: their address may lie outside the image
                v = fadd(fadd(fadd(fmul(w_tl, PxToFloat(p)), fmul(w_tr, PxToFloat(p + 1))), fmul(w_bl, PxToFloat(p + ref.pitch))),
                         fmul(w_br, PxToFloat(p + ref.pitch + 1)));
            if (in_patch) {
                *ex = v;
                *exv = ok ? 1 : 0;
            }
            valid += c.g.count(ok);
            p += p_step, ex += e_step, exv += e_step, row += kRowsPerStep, er += kRowsPerStep;
        }
        c.g.sync();
        return valid;
    }
    PatchWalk w;  // over the extended patch (two divisions per call instead of two per chunk)
    w.pc = c.geo.ec;
    w.row = c.g.lane / c.geo.ec;
    w.col = c.g.lane - w.row * c.geo.ec;
    w.step_row = G / c.geo.ec;
    w.step_col = G - w.step_row * c.geo.ec;
    for (int base = 0; base < c.geo.esize; base += G, w.next()) {
        const int e = base + c.g.lane;
        bool ok = false;
        if (e < c.geo.esize) {
            const int row = min_row + w.row, col = min_col + w.col;
            ok = !(row < 0 || row > ref.rows - 2 || col < 0 || col > ref.cols - 2);
            float v = 0.0f;
            if (ok) {
                v = fadd(fadd(fadd(fmul(w_tl, PxI(ref, row, col)), fmul(w_tr, PxI(ref, row, col + 1))), fmul(w_bl, PxI(ref, row + 1, col))),
                         fmul(w_br, PxI(ref, row + 1, col + 1)));
            }
            c.s.ex[e] = v;
            c.s.exv[e] = ok ? 1 : 0;
        }
        valid += c.g.count(ok);
    }
    c.g.sync();
    return valid;
}

// Gradient of the extended patch at interior pixel k (basic_klt_fast.cpp:71-94 and the affine / lssd twins).
template <int G>
__device__ __forceinline__ bool ExGradient(const Ctx<G> &c, int row, int col, float *dx, float *dy) {
    const int e = (row + 1) * c.geo.ec + col + 1;
    const int l = e - 1, r = e + 1, u = e - c.geo.ec, d = e + c.geo.ec;
    // The four 0 / 1 flags are read and combined unconditionally: `a && b && c && d` on shared-memory bytes costs a divergent branch per
    // flag (29 instructions here instead of 8).  The samples of invalid neighbours are 0 and always addressable (interior e).
    const unsigned ok = c.s.exv[l] & c.s.exv[r] & c.s.exv[u] & c.s.exv[d];
    const float gx = fsub(c.s.ex[r], c.s.ex[l]), gy = fsub(c.s.ex[d], c.s.ex[u]);
    *dx = ok ? gx : 0.0f;
    *dy = ok ? gy : 0.0f;
    return ok != 0u;
}

// Step bookkeeping shared by the fast trackers (basic_klt_fast.cpp:48-60, affine_klt_fast.cpp:55-67,
// lssd_klt_fast.cpp:101-112).  Returns true when the iteration loop must stop.
__device__ __forceinline__ bool FastStepCheck(const ftk_klt_params &p, float squared_step, float &last_squared_step, uint32_t &large_step_cnt,
                                              uint8_t &status) {
    if (squared_step < last_squared_step) {
        last_squared_step = squared_step;
        large_step_cnt = 0;
    } else {
        ++large_step_cnt;
        if (large_step_cnt >= p.max_tolerance_large_step) return true;
    }
    if (squared_step < p.max_converge_step) {
        status = FTK_STATUS_TRACKED;
        return true;
    }
    return false;
}

__device__ __forceinline__ bool IsNan(float v) { return v != v; }

// Sequential sum (row-major) of the interior of a rows x cols float array in shared memory: one chain.
template <int G>
__device__ float InteriorSum(Ctx<G> &c, const float *vals, int rows, int cols) {
    const int ir = rows - 2, ic = cols - 2;
    const int n = ir > 0 && ic > 0 ? ir * ic : 0;
    c.ch.reset();
    for (int base = 0; base < n; base += G) {
        const int q = base + c.g.lane;
        float v = 0.0f;
        if (q < n) v = vals[(q / ic + 1) * cols + q % ic + 1];
        c.ch.put(c.g.lane, 0, v);
        c.ch.template fold<1>(c.g);
    }
    return c.g.get(c.ch.acc, 0);
}

// ===================================================================================================================
// BASIC KLT
// ===================================================================================================================
struct BasicState {
    float cur_x, cur_y;
};

// basic_klt.cpp:118-181 ConstructIncrementalFunction.  H = {h00, h01, h11}.
template <int METHOD, int G>
__device__ int BasicConstruct(Ctx<G> &c, const Img &ref, const Img &cur, float ref_x, float ref_y, float cur_x, float cur_y, float (&H)[3],
                              float (&b)[2]) {
    int valid = 0;
    c.ch.reset();
    PatchWalk w = c.walk;
    for (int base = 0; base < c.geo.psize; base += G) {
        const int k = base + c.g.lane;
        float t[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        bool ok = false;
        if (k < c.geo.psize) {
            const int drow = w.row - c.geo.hr, dcol = w.col - c.geo.hc;
            const float row_i = fadd(static_cast<float>(drow), ref_y), col_i = fadd(static_cast<float>(dcol), ref_x);
            const float row_j = fadd(static_cast<float>(drow), cur_y), col_j = fadd(static_cast<float>(dcol), cur_x);
            const Img &gi = METHOD == kInverse ? ref : cur;
            const float gr = METHOD == kInverse ? row_i : row_j, gc = METHOD == kInverse ? col_i : col_j;
            float v0, v1, v2, v3, v4, v5;
            ok = PxChecked(gi, gr, fsub(gc, 1.0f), &v0) && PxChecked(gi, gr, fadd(gc, 1.0f), &v1) && PxChecked(gi, fsub(gr, 1.0f), gc, &v2) &&
                 PxChecked(gi, fadd(gr, 1.0f), gc, &v3) && PxChecked(ref, row_i, col_i, &v4) && PxChecked(cur, row_j, col_j, &v5);
            if (ok) {
                const float fx = fsub(v1, v0), fy = fsub(v3, v2), ft = fsub(v5, v4);
                t[0] = fmul(fx, fx);
                t[1] = fmul(fx, fy);
                t[2] = fmul(fy, fy);
                t[3] = -fmul(fx, ft);
                t[4] = -fmul(fy, ft);
            }
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) c.ch.put(c.g.lane, q, t[q]);
        valid += c.g.count(ok);
        c.ch.template fold<5>(c.g);
        w.next();
    }
    H[0] = c.g.get(c.ch.acc, 0);
    H[1] = c.g.get(c.ch.acc, 1);
    H[2] = c.g.get(c.ch.acc, 2);
    b[0] = c.g.get(c.ch.acc, 3);
    b[1] = c.g.get(c.ch.acc, 4);
    return valid;
}

// basic_klt.cpp:88-116 TrackOneFeature.
template <int METHOD, int G>
__device__ void BasicTrackOne(Ctx<G> &c, const Img &ref, const Img &cur, float ref_x, float ref_y, BasicState &s, uint8_t &status, bool done = false) {
    // `done`: see AffineTrackOne
    for (uint32_t iter = 0; iter < c.p->max_iteration; ++iter) {
        if (c.g.all(done)) break;
        float H[3], b[2];
        if (BasicConstruct<METHOD, G>(c, ref, cur, ref_x, ref_y, s.cur_x, s.cur_y, H, b) == 0) done = true;  // `break` in the reference
        if (done) continue;
        const float A[2][2] = {{H[0], H[1]}, {H[1], H[2]}};
        float v[2];
        LdltSolve<2>(A, b, v);
        if (IsNan(v[0]) || IsNan(v[1])) {
            status = FTK_STATUS_NUMERIC_ERROR;
            done = true;
            continue;
        }
        s.cur_x = fadd(s.cur_x, v[0]);
        s.cur_y = fadd(s.cur_y, v[1]);
        if (IsOutside(cur, s.cur_x, s.cur_y)) {
            status = FTK_STATUS_OUTSIDE;
            done = true;
            continue;
        }
        if (fadd(fmul(v[0], v[0]), fmul(v[1], v[1])) < c.p->max_converge_step) {
            status = FTK_STATUS_TRACKED;
            done = true;
        }
    }
}

// basic_klt_fast.cpp:7-62 TrackOneFeatureFast (+ :64-99 PrecomputeJacobianAndHessian, :101-195 ComputeBias).
template <int G>
__device__ void BasicTrackOneFast(Ctx<G> &c, const Img &ref, const Img &cur, float ref_x, float ref_y, BasicState &s, uint8_t &status,
                                  bool done = false) {  // `done`: see AffineTrackOne
    const int valid_ref = ExtractExRefPatch(c, ref, ref_x, ref_y);
    if (!done && valid_ref == 0) {
        status = FTK_STATUS_OUTSIDE;
        done = true;  // `return` in the reference
    }
    if (c.g.all(done)) return;
    // gradients + Hessian (3 chains)
    c.ch.reset();
    PatchWalk wg = c.walk;
    for (int base = 0; base < c.geo.psize; base += G, wg.next()) {
        const int k = base + c.g.lane;
        float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
        if (k < c.geo.psize) {
            float dx, dy;
            if (ExGradient(c, wg.row, wg.col, &dx, &dy)) {
                t0 = fmul(dx, dx);
                t1 = fmul(dx, dy);
                t2 = fmul(dy, dy);
            }
            c.s.dx[k] = dx;
            c.s.dy[k] = dy;
        }
        c.ch.put(c.g.lane, 0, t0);
        c.ch.put(c.g.lane, 1, t1);
        c.ch.put(c.g.lane, 2, t2);
        c.ch.template fold<3>(c.g);
    }
    const float h00 = c.g.get(c.ch.acc, 0), h01 = c.g.get(c.ch.acc, 1), h11 = c.g.get(c.ch.acc, 2);
    const float A[2][2] = {{h00, h01}, {h01, h11}};
    LdltFactors<2> factors;
    LdltFactor<2>(A, factors);

    if (!done) status = FTK_STATUS_LARGE_RESIDUAL;
    float last_squared_step = INFINITY;
    uint32_t large_step_cnt = 0;
    for (uint32_t iter = 0; iter < c.p->max_iteration; ++iter) {
        if (c.g.all(done)) break;
        // ComputeBias: integer-aligned window at floor(cur), one weight set.
        const float int_row = floorf(s.cur_y), int_col = floorf(s.cur_x);
        const float dec_row = fsub(s.cur_y, int_row), dec_col = fsub(s.cur_x, int_col);
        const float w_tl = fmul(fsub(1.0f, dec_row), fsub(1.0f, dec_col));
        const float w_tr = fmul(fsub(1.0f, dec_row), dec_col);
        const float w_bl = fmul(dec_row, fsub(1.0f, dec_col));
        const float w_br = fmul(dec_row, dec_col);
        const int min_row = static_cast<int>(int_row) - c.geo.pr / 2;
        const int min_col = static_cast<int>(int_col) - c.geo.pc / 2;
        int valid = 0;
        c.ch.reset();
        PatchWalk w = c.walk;
        for (int base = 0; base < c.geo.psize; base += G) {
            const int k = base + c.g.lane;
            float t0 = 0.0f, t1 = 0.0f;
            bool ok = false;
            if (k < c.geo.psize) {
                const int prow = w.row, pcol = w.col;
                const int row = min_row + prow, col = min_col + pcol;
                const int e = (prow + 1) * c.geo.ec + pcol + 1;
                ok = !(row < 0 || row > cur.rows - 2 || col < 0 || col > cur.cols - 2) & (c.s.exv[e] != 0);  // flag read unconditionally: no branch
                if (ok) {
                    const float cur_value = fadd(fadd(fadd(fmul(w_tl, PxI(cur, row, col)), fmul(w_tr, PxI(cur, row, col + 1))), fmul(w_bl, PxI(cur, row + 1, col))),
                                                 fmul(w_br, PxI(cur, row + 1, col + 1)));
                    const float dt = fsub(cur_value, c.s.ex[e]);
                    t0 = -fmul(c.s.dx[k], dt);
                    t1 = -fmul(c.s.dy[k], dt);
                }
            }
            c.ch.put(c.g.lane, 0, t0);
            c.ch.put(c.g.lane, 1, t1);
            valid += c.g.count(ok);
            c.ch.template fold<2>(c.g);
            w.next();
        }
        const float b[2] = {c.g.get(c.ch.acc, 0), c.g.get(c.ch.acc, 1)};
        if (valid == 0) done = true;  // `break` in the reference
        if (done) continue;
        float v[2];
        LdltSolveFactored<2>(factors, b, v);
        if (IsNan(v[0]) || IsNan(v[1])) {
            status = FTK_STATUS_NUMERIC_ERROR;
            done = true;
            continue;
        }
        s.cur_x = fadd(s.cur_x, v[0]);
        s.cur_y = fadd(s.cur_y, v[1]);
        const float squared_step = fadd(fmul(v[0], v[0]), fmul(v[1], v[1]));
        if (FastStepCheck(*c.p, squared_step, last_squared_step, large_step_cnt, status)) done = true;
    }
}

// ===================================================================================================================
// AFFINE KLT.  affine = {a00, a01, a10, a11} row-major.
// ===================================================================================================================
struct AffineState {
    float cur_x, cur_y;
    float a[4];
};

// Upper-triangle index pairs of the 21 accumulated Hessian entries, in the reference's order.
__device__ const unsigned char kAffRow[21] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 5};
__device__ const unsigned char kAffCol[21] = {0, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 2, 3, 4, 5, 3, 4, 5, 4, 5, 5};

// The 21 Hessian terms of one pixel (affine_klt.cpp:157-187; entry (3,4) accumulates yy*dxdy -- reproduced).
__device__ __forceinline__ void AffineHessianTerms(float x, float y, float dx, float dy, float (&t)[27]) {
    const float xx = fmul(x, x), yy = fmul(y, y), dxdx = fmul(dx, dx), dydy = fmul(dy, dy), xy = fmul(x, y), dxdy = fmul(dx, dy);
    t[0] = fmul(xx, dxdx);
    t[1] = fmul(xx, dxdy);
    t[2] = fmul(xy, dxdx);
    t[3] = fmul(xy, dxdy);
    t[4] = fmul(x, dxdx);
    t[5] = fmul(x, dxdy);
    t[6] = fmul(xx, dydy);
    t[7] = fmul(xy, dxdy);
    t[8] = fmul(xy, dydy);
    t[9] = fmul(x, dxdy);
    t[10] = fmul(x, dydy);
    t[11] = fmul(yy, dxdx);
    t[12] = fmul(yy, dxdy);
    t[13] = fmul(y, dxdx);
    t[14] = fmul(y, dxdy);
    t[15] = fmul(yy, dydy);
    t[16] = fmul(yy, dxdy);
    t[17] = fmul(y, dydy);
    t[18] = dxdx;
    t[19] = dxdy;
    t[20] = dydy;
}

// The 6 bias terms of one pixel, negated because the reference subtracts them (affine_klt.cpp:189-194).
__device__ __forceinline__ void AffineBiasTerms(float x, float y, float dx, float dy, float dt, float *t) {
    // (-a) * b == -(a * b) bit for bit (signed zeros included): the negation rides on an operand instead of costing an instruction
    const float ndt = -dt;
    t[0] = fmul(fmul(ndt, x), dx);
    t[1] = fmul(fmul(ndt, x), dy);
    t[2] = fmul(fmul(ndt, y), dx);
    t[3] = fmul(fmul(ndt, y), dy);
    t[4] = fmul(ndt, dx);
    t[5] = fmul(ndt, dy);
}

// Where chain q of the affine normal equations goes inside Ldlt6Shared (as a float offset): Hessian entry (row, col) with
// row <= col is stored at the lower-triangle position a[col][row]; chains 21..26 are the right-hand side b[0..5].
__device__ __forceinline__ int AffineChainSlot(int q) {
    if (q < 0 || q >= 27) return -1;
    if (q >= 21) return 36 + 8 + (q - 21);
    return kAffCol[q] * 6 + kAffRow[q];
}

// Chain lanes store their accumulators into the LDLT scratch (n_chains = 27: Hessian + bias, 21: Hessian, 6: bias only, which then
// sit in chains 0..5).
template <int G>
__device__ __forceinline__ void AffineScatterChains(Ctx<G> &c, int n_chains) {
    float *dst = reinterpret_cast<float *>(c.s.ldlt);
    if (n_chains == 6) {
        if (c.g.lane < 6) c.s.ldlt->b[c.g.lane] = c.ch.acc;
    } else {
        if (c.g.lane < n_chains && c.ldlt_dst0 >= 0) dst[c.ldlt_dst0] = c.ch.acc;
        if (G + c.g.lane < n_chains && c.ldlt_dst1 >= 0) dst[c.ldlt_dst1] = c.ch.acc_hi;
    }
    c.g.sync();
}

// The reference samples of affine_klt.cpp:131-273 do not depend on the iteration: the centre sample I_ref(row_i, col_i)
// (both methods) and, for kInverse, the gradient stencil of the reference image are evaluated once per level into the
// group's shared scratch.  Bit q of the result = this lane's pixel of chunk q has all its reference samples inside.
template <int METHOD, int G>
__device__ unsigned long long AffineHoistRef(Ctx<G> &c, const Img &ref, float ref_x, float ref_y) {
    const int pf = RoundUp(c.geo.psize, 4);
    float *hv4 = c.s.hoist, *hfx = c.s.hoist + pf, *hfy = c.s.hoist + 2 * pf;  // hfx / hfy exist for kInverse only
    unsigned long long ref_bits = 0ull;
    PatchWalk w = c.walk;
    int chunk = 0;
    c.g.sync();  // the previous level has finished reading the scratch
    for (int base = 0; base < c.geo.psize; base += G, ++chunk) {
        const int k = base + c.g.lane;
        if (k < c.geo.psize) {
            const float row_i = fadd(static_cast<float>(w.row - c.geo.hr), ref_y), col_i = fadd(static_cast<float>(w.col - c.geo.hc), ref_x);
            if constexpr (METHOD == kInverse) {
                float v0, v1, v2, v3, v4;
                if (PxStencil5(ref, row_i, col_i, &v0, &v1, &v2, &v3, &v4)) {
                    hfx[k] = fsub(v1, v0);
                    hfy[k] = fsub(v3, v2);
                    hv4[k] = v4;
                    ref_bits |= 1ull << chunk;
                }
            } else {
                float v4;
                if (PxChecked(ref, row_i, col_i, &v4)) {
                    hv4[k] = v4;
                    ref_bits |= 1ull << chunk;
                }
            }
        }
        w.next();
    }
    return ref_bits;  // each lane reads back only what it wrote: no barrier needed
}

// affine_klt.cpp:131-273 ConstructIncrementalFunction: 21 Hessian + 6 bias chains, on top of the hoisted reference samples.
template <int METHOD, int G>
__device__ int AffineConstruct(Ctx<G> &c, const Img &cur, const AffineState &s, unsigned long long ref_bits) {
    static_assert(2 * G >= 27, "affine needs 27 chains, at most two per lane");
    const int pf = RoundUp(c.geo.psize, 4);
    const float *hv4 = c.s.hoist, *hfx = c.s.hoist + pf, *hfy = c.s.hoist + 2 * pf;
    int valid = 0;
    c.ch.reset();
    PatchWalk w = c.walk;
    int chunk = 0;
    for (int base = 0; base < c.geo.psize; base += G, ++chunk) {
        const int k = base + c.g.lane;
        float t[27];
        bool ok = false;
        float tx = 0.0f, ty = 0.0f, tdx = 0.0f, tdy = 0.0f, tdt = 0.0f;  // invalid pixels: all 27 terms become +-0 (no-ops in the chains)
        if ((ref_bits >> chunk) & 1ull) {  // implies k < psize
            const float fdrow = static_cast<float>(w.row - c.geo.hr), fdcol = static_cast<float>(w.col - c.geo.hc);
            const float ax = fadd(fmul(s.a[0], fdcol), fmul(s.a[1], fdrow));
            const float ay = fadd(fmul(s.a[2], fdcol), fmul(s.a[3], fdrow));
            const float row_j = fadd(ay, s.cur_y), col_j = fadd(ax, s.cur_x);
            float dx, dy, v5;
            if constexpr (METHOD == kDirect) {
                float v0, v1, v2, v3;
                ok = PxStencil5(cur, row_j, col_j, &v0, &v1, &v2, &v3, &v5);
                dx = fsub(v1, v0), dy = fsub(v3, v2);
            } else {
                ok = PxChecked(cur, row_j, col_j, &v5);
                dx = hfx[k], dy = hfy[k];
            }
            if (ok) tx = col_j, ty = row_j, tdx = dx, tdy = dy, tdt = fsub(v5, hv4[k]);
        }
        AffineHessianTerms(tx, ty, tdx, tdy, t);
        AffineBiasTerms(tx, ty, tdx, tdy, tdt, &t[21]);
        valid += c.g.count(ok);
        if constexpr (G == 16) {
            // 27 chains on 16 lanes: chain k and chain 16 + k as one 64-bit term, one packed add per pixel (see Chain::fold_pairs)
#pragma unroll
            for (int q = 0; q < 16; ++q) c.ch.put_pair(c.g.lane, q, t[q], q + 16 < 27 ? t[q + 16] : 0.0f);
            c.ch.fold_pairs(c.g);
        } else {
#pragma unroll
            for (int q = 0; q < 27; ++q) c.ch.put(c.g.lane, q, t[q]);
            c.ch.template fold_wide<27>(c.g);
        }
        w.next();
    }
    AffineScatterChains(c, 27);  // Hessian (lower triangle) and bias -> the LDLT scratch
    return valid;
}

__device__ __forceinline__ void AffineApply(AffineState &s, const float (&z)[6], float v0, float v1) {
    s.cur_x = fadd(s.cur_x, v0);
    s.cur_y = fadd(s.cur_y, v1);
    s.a[0] = fadd(s.a[0], z[0]);  // col(0) += z.head<2>()
    s.a[2] = fadd(s.a[2], z[1]);
    s.a[1] = fadd(s.a[1], z[2]);  // col(1) += z.segment<2>(2)
    s.a[3] = fadd(s.a[3], z[3]);
}

// affine_klt.cpp:93-129 TrackOneFeature.
// `done` on entry: this group only keeps its partner company (a feature that is not tracked, or a group without a feature).  The loop
// runs until every group that shares the warp-level operations (Group::mask) has left it: a finished group keeps executing the same
// instructions on its frozen state and discards the results, so barriers and votes stay warp-uniform.
template <int METHOD, int G>
__device__ void AffineTrackOne(Ctx<G> &c, const Img &ref, const Img &cur, float ref_x, float ref_y, AffineState &s, uint8_t &status, bool done = false) {
    const unsigned long long ref_bits = AffineHoistRef<METHOD, G>(c, ref, ref_x, ref_y);
    for (uint32_t iter = 0; iter < c.p->max_iteration; ++iter) {
        if (c.g.all(done)) break;
        float z[6];
        if (AffineConstruct<METHOD, G>(c, cur, s, ref_bits) == 0) done = true;  // `break` in the reference
        if (c.g.all(done)) break;
        Ldlt6FactorShared(c.g, *c.s.ldlt);  // hessian.ldlt().solve(bias), affine_klt.cpp:103
        Ldlt6SolveShared(c.g, *c.s.ldlt, z);
        if (done) continue;
        const float v0 = fadd(fadd(fmul(z[0], s.cur_x), fmul(z[2], s.cur_y)), z[4]);
        const float v1 = fadd(fadd(fmul(z[1], s.cur_x), fmul(z[3], s.cur_y)), z[5]);
        if (IsNan(v0) || IsNan(v1)) {
            status = FTK_STATUS_NUMERIC_ERROR;
            done = true;
            continue;
        }
        AffineApply(s, z, v0, v1);
        if (IsOutside(cur, s.cur_x, s.cur_y)) {
            status = FTK_STATUS_OUTSIDE;
            done = true;
            continue;
        }
        if (fadd(fmul(v0, v0), fmul(v1, v1)) < c.p->max_converge_step) {
            status = FTK_STATUS_TRACKED;
            done = true;
        }
    }
}

// affine_klt_fast.cpp:7-69 TrackOneFeatureFast (+ :71-138 PrecomputeJacobianAndHessian, :140-188 ComputeBias).
template <int G>
__device__ void AffineTrackOneFast(Ctx<G> &c, const Img &ref, const Img &cur, float ref_x, float ref_y, AffineState &s, uint8_t &status,
                                   bool done = false) {  // `done`: see AffineTrackOne
    const int valid_ref = ExtractExRefPatch(c, ref, ref_x, ref_y);
    if (!done && valid_ref == 0) {
        status = FTK_STATUS_OUTSIDE;
        done = true;  // `return` in the reference
    }
    if (c.g.all(done)) return;
    // Hessian at the level-entry cur position: 21 chains of which (1,2), (1,4), (3,4) are overwritten by copies.
    c.ch.reset();
    PatchWalk w = c.walk;
    for (int base = 0; base < c.geo.psize; base += G) {
        const int k = base + c.g.lane;
        float t[27];
        // Pixels without a gradient (or past the patch) run the same products on zeros: every term is then +0, and adding
        // +0 leaves a chain unchanged -- no zero-fill, no divergent branch around the 21 products.
        float dx = 0.0f, dy = 0.0f, x = 0.0f, y = 0.0f;
        if (k < c.geo.psize) {
            if (ExGradient(c, w.row, w.col, &dx, &dy)) {
                x = fadd(static_cast<float>(w.col - c.geo.hc), s.cur_x);
                y = fadd(static_cast<float>(w.row - c.geo.hr), s.cur_y);
            }
            c.s.dx[k] = dx;
            c.s.dy[k] = dy;
        }
        AffineHessianTerms(x, y, dx, dy, t);
        if constexpr (G == 16) {
            // 21 chains through an 8-chain buffer in three parts (chains 0..7, 8..15, 16..20): shared memory per feature decides how many
            // warps an SM holds, and this pass runs once per level
#pragma unroll
            for (int q = 0; q < 8; ++q) c.ch.put(c.g.lane, q, t[q]);
            c.ch.template fold_at<8, 0>(c.g);
#pragma unroll
            for (int q = 8; q < 16; ++q) c.ch.put(c.g.lane, q - 8, t[q]);
            c.ch.template fold_at<8, 8>(c.g);
#pragma unroll
            for (int q = 16; q < 21; ++q) c.ch.put(c.g.lane, q - 16, t[q]);
            c.ch.template fold_hi<5>(c.g);
        } else {
#pragma unroll
            for (int q = 0; q < 21; ++q) c.ch.put(c.g.lane, q, t[q]);
            c.ch.template fold_wide<21>(c.g);
        }
        w.next();
    }
    AffineScatterChains(c, 21);
    if (c.g.lane == 0) {
        // affine_klt_fast.cpp:127-135: (1,2), (1,4), (3,4) are copies of (0,3), (0,5), (2,3); lower-triangle positions a[col][row]
        float *a = c.s.ldlt->a;
        a[2 * 6 + 1] = a[3 * 6 + 0];
        a[4 * 6 + 1] = a[5 * 6 + 0];
        a[4 * 6 + 3] = a[3 * 6 + 2];
    }
    c.g.sync();
    Ldlt6FactorShared(c.g, *c.s.ldlt);  // the Hessian is fixed for the level: factorise once (identical factors every iteration)

    float last_squared_step = INFINITY;
    uint32_t large_step_cnt = 0;
    if (!done) status = FTK_STATUS_LARGE_RESIDUAL;
    for (uint32_t iter = 0; iter < c.p->max_iteration; ++iter) {
        if (c.g.all(done)) break;
        int valid = 0;
        c.ch.reset();
        PatchWalk w = c.walk;
        for (int base = 0; base < c.geo.psize; base += G) {
            const int k = base + c.g.lane;
            float t[6];
            bool ok = false;
            float bx = 0.0f, by = 0.0f, bdx = 0.0f, bdy = 0.0f, bdt = 0.0f;  // invalid pixels: products of zeros, i.e. -0 terms (no-ops)
            if (k < c.geo.psize) {
                const int prow = w.row, pcol = w.col;
                const int drow = prow - c.geo.hr, dcol = pcol - c.geo.hc;
                const float ax = fadd(fmul(s.a[0], static_cast<float>(dcol)), fmul(s.a[1], static_cast<float>(drow)));
                const float ay = fadd(fmul(s.a[2], static_cast<float>(dcol)), fmul(s.a[3], static_cast<float>(drow)));
                const float row_c = fadd(ay, s.cur_y), col_c = fadd(ax, s.cur_x);
                const int e = (prow + 1) * c.geo.ec + pcol + 1;
                float cur_value;
                const bool ref_ok = c.s.exv[e] != 0;  // read before the sample so that `&&` does not become a second branch
                ok = PxChecked(cur, row_c, col_c, &cur_value) & ref_ok;
                if (ok) {
                    bdt = fsub(cur_value, c.s.ex[e]);
                    bx = col_c, by = row_c, bdx = c.s.dx[k], bdy = c.s.dy[k];
                }
            }
            AffineBiasTerms(bx, by, bdx, bdy, bdt, t);
#pragma unroll
            for (int q = 0; q < 6; ++q) c.ch.put(c.g.lane, q, t[q]);
            valid += c.g.count(ok);
            c.ch.template fold<6>(c.g);
            w.next();
        }
        if (valid == 0) done = true;  // `break` in the reference
        if (c.g.all(done)) break;
        float z[6];
        AffineScatterChains(c, 6);
        Ldlt6SolveShared(c.g, *c.s.ldlt, z);
        if (done) continue;
        bool any_nan = false;
#pragma unroll
        for (int q = 0; q < 6; ++q) any_nan = any_nan || IsNan(z[q]);
        if (any_nan) {
            status = FTK_STATUS_NUMERIC_ERROR;
            done = true;
            continue;
        }
        const float v0 = fadd(fadd(fmul(z[0], s.cur_x), fmul(z[2], s.cur_y)), z[4]);
        const float v1 = fadd(fadd(fmul(z[1], s.cur_x), fmul(z[3], s.cur_y)), z[5]);
        AffineApply(s, z, v0, v1);
        const float squared_step = fadd(fmul(v0, v0), fmul(v1, v1));
        if (FastStepCheck(*c.p, squared_step, last_squared_step, large_step_cnt, status)) done = true;
    }
}

// ===================================================================================================================
// LSSD KLT.  R = {r00, r01, r10, r11} row-major, t = {tx, ty}.
// ===================================================================================================================
struct LssdState {
    float R[4];
    float t[2];
};

// lssd_klt.cpp:113-117 / lssd_klt_fast.cpp:95-99: R *= [1 -th; th 1]; R /= ||R.col(0)||; t += v.tail<2>().
__device__ __forceinline__ void LssdUpdate(LssdState &s, const float (&v)[3]) {
    const float th = v[0];
    const float n00 = fadd(fmul(s.R[0], 1.0f), fmul(s.R[1], th));
    const float n01 = fadd(fmul(s.R[0], -th), fmul(s.R[1], 1.0f));
    const float n10 = fadd(fmul(s.R[2], 1.0f), fmul(s.R[3], th));
    const float n11 = fadd(fmul(s.R[2], -th), fmul(s.R[3], 1.0f));
    const float norm = __fsqrt_rn(fadd(fmul(n00, n00), fmul(n10, n10)));
    s.R[0] = fdiv(n00, norm);
    s.R[1] = fdiv(n01, norm);
    s.R[2] = fdiv(n10, norm);
    s.R[3] = fdiv(n11, norm);
    s.t[0] = fadd(s.t[0], v[1]);
    s.t[1] = fadd(s.t[1], v[2]);
}

__device__ __forceinline__ void LssdWarp(const LssdState &s, float col_i, float row_i, float *col_j, float *row_j) {
    *col_j = fadd(fadd(fmul(s.R[0], col_i), fmul(s.R[1], row_i)), s.t[0]);
    *row_j = fadd(fadd(fmul(s.R[2], col_i), fmul(s.R[3], row_i)), s.t[1]);
}

// 6 unique Hessian products + 3 negated bias products of one pixel (H += J^T J, b -= J^T r).
__device__ __forceinline__ void LssdTerms(const float (&J)[3], float residual, float (&t)[9]) {
    t[0] = fmul(J[0], J[0]);
    t[1] = fmul(J[0], J[1]);
    t[2] = fmul(J[0], J[2]);
    t[3] = fmul(J[1], J[1]);
    t[4] = fmul(J[1], J[2]);
    t[5] = fmul(J[2], J[2]);
    const float nres = -residual;  // (-a) * b == -(a * b) bit for bit
    t[6] = fmul(J[0], nres);
    t[7] = fmul(J[1], nres);
    t[8] = fmul(J[2], nres);
}

// Chains 0..5 = Hessian (0,0) (0,1) (0,2) (1,1) (1,2) (2,2), chains 6..8 = bias: the owning lanes store them into the LDLT scratch
// (lower-triangle position a[col][row]), then the 3 x 3 system is solved cooperatively (hessian.ldlt().solve(bias)).
template <int G>
__device__ __forceinline__ void LssdSolve(Ctx<G> &c, float (&v)[3]) {
    LdltShared<3> &s = *c.s.ldlt3;
    const int lane = c.g.lane;
    if (lane < 6) {
        const int slot = lane == 0 ? 0 : (lane == 1 ? 3 : (lane == 2 ? 6 : (lane == 3 ? 4 : (lane == 4 ? 7 : 8))));
        s.a[slot] = c.ch.acc;
    } else if (lane < 9) {
        s.b[lane - 6] = c.ch.acc;
    }
    c.g.sync();
    LdltFactorShared<3, G>(c.g, s);
    LdltSolveShared<3, G>(c.g, s, v);
}

// Second pass of lssd_klt.cpp:127-250 (Jacobians, residuals, 9 chains) over the pixels of `ok_bits`.  FAST: both patch means are
// in the shared-divisor range (always, for real images), so every division is the 3-instruction form.
template <int METHOD, int G, bool FAST>
__device__ __forceinline__ void LssdSecondPass(Ctx<G> &c, const Img &ref, const Img &cur, float ref_x, float ref_y, const LssdState &s, unsigned long long ok_bits,
                                               const SharedDivisor &by_ref, const SharedDivisor &by_cur) {
    c.ch.reset();
    int chunk = 0;
    PatchWalk w = c.walk;
    for (int base = 0; base < c.geo.psize; base += G, ++chunk) {
        float t[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) t[q] = 0.0f;
        if ((ok_bits >> chunk) & 1ull) {
            const int drow = w.row - c.geo.hr, dcol = w.col - c.geo.hc;
            const float row_i = fadd(static_cast<float>(drow), ref_y), col_i = fadd(static_cast<float>(dcol), ref_x);
            float row_j, col_j;
            LssdWarp(s, col_i, row_i, &col_j, &row_j);
            const Img &gi = METHOD == kInverse ? ref : cur;
            const float gr = METHOD == kInverse ? row_i : row_j, gc = METHOD == kInverse ? col_i : col_j;
            const float v0 = PxF(gi, gr, fsub(gc, 1.0f));
            const float v1 = PxF(gi, gr, fadd(gc, 1.0f));
            const float v2 = PxF(gi, fsub(gr, 1.0f), gc);
            const float v3 = PxF(gi, fadd(gr, 1.0f), gc);
            const float v4 = PxF(ref, row_i, col_i);
            const float v5 = PxF(cur, row_j, col_j);
            const SharedDivisor &by_avg = METHOD == kInverse ? by_ref : by_cur;
            const float jp0 = DivideBy<FAST>(by_avg, fsub(v1, v0)), jp1 = DivideBy<FAST>(by_avg, fsub(v3, v2));
            const float s00 = fadd(fmul(s.R[0], -row_i), fmul(s.R[1], col_i));
            const float s10 = fadd(fmul(s.R[2], -row_i), fmul(s.R[3], col_i));
            float J[3];
            J[0] = fadd(fmul(jp0, s00), fmul(jp1, s10));
            J[1] = fadd(fmul(jp0, 1.0f), fmul(jp1, 0.0f));
            J[2] = fadd(fmul(jp0, 0.0f), fmul(jp1, 1.0f));
            const float residual = fsub(DivideBy<FAST>(by_cur, v5), DivideBy<FAST>(by_ref, v4));
            LssdTerms(J, residual, t);
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) c.ch.put(c.g.lane, q, t[q]);
        c.ch.template fold<9>(c.g);
        w.next();
    }
}

// lssd_klt.cpp:127-250 ConstructIncrementalFunction: pass 1 = validity + patch means (2 chains), pass 2 = 9 chains.
template <int METHOD, int G>
__device__ int LssdConstruct(Ctx<G> &c, const Img &ref, const Img &cur, float ref_x, float ref_y, const LssdState &s) {
    int valid = 0;
    unsigned long long ok_bits = 0ull;  // bit q: this lane's pixel of chunk q is valid (psize <= 64 * G, checked on the host)
    c.ch.reset();
    int chunk = 0;
    PatchWalk w = c.walk;
    for (int base = 0; base < c.geo.psize; base += G, ++chunk) {
        const int k = base + c.g.lane;
        float t0 = 0.0f, t1 = 0.0f;
        bool ok = false;
        if (k < c.geo.psize) {
            const int drow = w.row - c.geo.hr, dcol = w.col - c.geo.hc;
            const float row_i = fadd(static_cast<float>(drow), ref_y), col_i = fadd(static_cast<float>(dcol), ref_x);
            float row_j, col_j;
            LssdWarp(s, col_i, row_i, &col_j, &row_j);
            const Img &gi = METHOD == kInverse ? ref : cur;
            const float gr = METHOD == kInverse ? row_i : row_j, gc = METHOD == kInverse ? col_i : col_j;
            float v4, v5;
            ok = PxInside(gi, gr, fsub(gc, 1.0f)) && PxInside(gi, gr, fadd(gc, 1.0f)) && PxInside(gi, fsub(gr, 1.0f), gc) &&
                 PxInside(gi, fadd(gr, 1.0f), gc) && PxChecked(ref, row_i, col_i, &v4) && PxChecked(cur, row_j, col_j, &v5);
            if (ok) {
                t0 = v4;
                t1 = v5;
                ok_bits |= 1ull << chunk;
            }
        }
        c.ch.put(c.g.lane, 0, t0);
        c.ch.put(c.g.lane, 1, t1);
        valid += c.g.count(ok);
        c.ch.template fold<2>(c.g);
        w.next();
    }
    const float ref_avg = fdiv(c.g.get(c.ch.acc, 0), static_cast<float>(valid));
    const float cur_avg = fdiv(c.g.get(c.ch.acc, 1), static_cast<float>(valid));
    const SharedDivisor by_ref = MakeSharedDivisor(ref_avg), by_cur = MakeSharedDivisor(cur_avg);  // every division below is by one of the two patch means
    if (by_ref.fast && by_cur.fast) LssdSecondPass<METHOD, G, true>(c, ref, cur, ref_x, ref_y, s, ok_bits, by_ref, by_cur);
    else LssdSecondPass<METHOD, G, false>(c, ref, cur, ref_x, ref_y, s, ok_bits, by_ref, by_cur);
    return valid;
}

// kInverse only: the five reference samples of a pixel (lssd_klt.cpp:152-155, 202-206) do not depend on the iteration.  They are
// evaluated once per level: gradients fx, fy and the centre sample go to shared memory, the "all five inside the image" bit to
// a per-lane mask (bit q = this lane's pixel of chunk q).  Same values, same order of use as the reference's per-iteration
// recomputation.
template <int G>
__device__ unsigned long long LssdHoistRef(Ctx<G> &c, const Img &ref, float ref_x, float ref_y) {
    const int pf = RoundUp(c.geo.psize, 4);
    float *hfx = c.s.hoist, *hfy = c.s.hoist + pf, *hv4 = c.s.hoist + 2 * pf;
    unsigned long long ref_bits = 0ull;
    int chunk = 0;
    PatchWalk w = c.walk;
    for (int base = 0; base < c.geo.psize; base += G, ++chunk) {
        const int k = base + c.g.lane;
        if (k < c.geo.psize) {
            const int drow = w.row - c.geo.hr, dcol = w.col - c.geo.hc;
            const float row_i = fadd(static_cast<float>(drow), ref_y), col_i = fadd(static_cast<float>(dcol), ref_x);
            float v0, v1, v2, v3, v4;
            if (PxStencil5(ref, row_i, col_i, &v0, &v1, &v2, &v3, &v4)) {
                hfx[k] = fsub(v1, v0);
                hfy[k] = fsub(v3, v2);
                hv4[k] = v4;
                ref_bits |= 1ull << chunk;
            }
        }
        w.next();
    }
    return ref_bits;
}

// Second pass of the hoisted kInverse form: Jacobians and residuals from the per-level fx / fy / v4 and this iteration's v5.
template <int G, bool FAST>
__device__ __forceinline__ void LssdSecondPassHoisted(Ctx<G> &c, float ref_x, float ref_y, const LssdState &s, unsigned long long ok_bits,
                                                      const SharedDivisor &by_ref, const SharedDivisor &by_cur) {
    const int pf = RoundUp(c.geo.psize, 4);
    const float *hfx = c.s.hoist, *hfy = c.s.hoist + pf, *hv4 = c.s.hoist + 2 * pf, *hv5 = c.s.hoist + 3 * pf;
    c.ch.reset();
    int chunk = 0;
    PatchWalk w = c.walk;
    for (int base = 0; base < c.geo.psize; base += G, ++chunk) {
        const int k = base + c.g.lane;
        float t[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) t[q] = 0.0f;
        if ((ok_bits >> chunk) & 1ull) {
            const int drow = w.row - c.geo.hr, dcol = w.col - c.geo.hc;
            const float row_i = fadd(static_cast<float>(drow), ref_y), col_i = fadd(static_cast<float>(dcol), ref_x);
            const float jp0 = DivideBy<FAST>(by_ref, hfx[k]), jp1 = DivideBy<FAST>(by_ref, hfy[k]);
            const float s00 = fadd(fmul(s.R[0], -row_i), fmul(s.R[1], col_i));
            const float s10 = fadd(fmul(s.R[2], -row_i), fmul(s.R[3], col_i));
            float J[3];
            J[0] = fadd(fmul(jp0, s00), fmul(jp1, s10));
            J[1] = fadd(fmul(jp0, 1.0f), fmul(jp1, 0.0f));
            J[2] = fadd(fmul(jp0, 0.0f), fmul(jp1, 1.0f));
            const float residual = fsub(DivideBy<FAST>(by_cur, hv5[k]), DivideBy<FAST>(by_ref, hv4[k]));
            LssdTerms(J, residual, t);
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) c.ch.put(c.g.lane, q, t[q]);
        c.ch.template fold<9>(c.g);
        w.next();
    }
}

// lssd_klt.cpp:127-250 ConstructIncrementalFunction, kInverse, on top of the hoisted reference samples.
template <int G>
__device__ int LssdConstructHoisted(Ctx<G> &c, const Img &cur, float ref_x, float ref_y, const LssdState &s, unsigned long long ref_bits, bool done) {
    const int pf = RoundUp(c.geo.psize, 4);
    const float *hfx = c.s.hoist, *hfy = c.s.hoist + pf, *hv4 = c.s.hoist + 2 * pf;
    float *hv5 = c.s.hoist + 3 * pf;
    int valid = 0;
    unsigned long long ok_bits = 0ull;
    c.ch.reset();
    int chunk = 0;
    PatchWalk w = c.walk;
    for (int base = 0; base < c.geo.psize; base += G, ++chunk) {
        const int k = base + c.g.lane;
        float t0 = 0.0f, t1 = 0.0f;
        bool ok = false;
        if ((ref_bits >> chunk) & 1ull) {
            const int drow = w.row - c.geo.hr, dcol = w.col - c.geo.hc;
            const float row_i = fadd(static_cast<float>(drow), ref_y), col_i = fadd(static_cast<float>(dcol), ref_x);
            float row_j, col_j, v5;
            LssdWarp(s, col_i, row_i, &col_j, &row_j);
            ok = PxChecked(cur, row_j, col_j, &v5);
            if (ok) {
                t0 = hv4[k];
                t1 = v5;
                hv5[k] = v5;
                ok_bits |= 1ull << chunk;
            }
        }
        c.ch.put(c.g.lane, 0, t0);
        c.ch.put(c.g.lane, 1, t1);
        valid += c.g.count(ok);
        c.ch.template fold<2>(c.g);
        w.next();
    }
    const float ref_avg = fdiv(c.g.get(c.ch.acc, 0), static_cast<float>(valid));
    const float cur_avg = fdiv(c.g.get(c.ch.acc, 1), static_cast<float>(valid));
    const SharedDivisor by_ref = MakeSharedDivisor(ref_avg), by_cur = MakeSharedDivisor(cur_avg);  // every division below is by one of the two patch means
    // One choice for every group that shares the warp-level operations; a finished group (results discarded) never forces the general form.
    if (c.g.all(done || (by_ref.fast && by_cur.fast))) LssdSecondPassHoisted<G, true>(c, ref_x, ref_y, s, ok_bits, by_ref, by_cur);
    else LssdSecondPassHoisted<G, false>(c, ref_x, ref_y, s, ok_bits, by_ref, by_cur);
    return valid;
}

// lssd_klt.cpp:96-125 TrackOneFeature.
template <int METHOD, int G>
__device__ void LssdTrackOne(Ctx<G> &c, const Img &ref, const Img &cur, float ref_x, float ref_y, LssdState &s, uint8_t &status, bool done = false) {
    // `done`: see AffineTrackOne
    unsigned long long ref_bits = 0ull;
    if constexpr (METHOD == kInverse) ref_bits = LssdHoistRef<G>(c, ref, ref_x, ref_y);
    for (uint32_t iter = 0; iter < c.p->max_iteration; ++iter) {
        if (c.g.all(done)) break;
        float v[3];
        int valid;
        if constexpr (METHOD == kInverse) valid = LssdConstructHoisted<G>(c, cur, ref_x, ref_y, s, ref_bits, done);
        else valid = LssdConstruct<METHOD, G>(c, ref, cur, ref_x, ref_y, s);
        if (valid == 0) done = true;  // `break` in the reference
        if (c.g.all(done)) break;
        LssdSolve(c, v);
        if (done) continue;
        if (IsNan(v[0]) || IsNan(v[1]) || IsNan(v[2])) {
            status = FTK_STATUS_NUMERIC_ERROR;
            done = true;
            continue;
        }
        LssdUpdate(s, v);
        if (fadd(fadd(fmul(v[0], v[0]), fmul(v[1], v[1])), fmul(v[2], v[2])) < c.p->max_converge_step) {
            status = FTK_STATUS_TRACKED;
            done = true;
        }
    }
}

// lssd_klt_fast.cpp:7-114 TrackOneFeatureFast (+ :116-143, :145-195, :197-229).
template <int G>
__device__ void LssdTrackOneFast(Ctx<G> &c, const Img &ref, const Img &cur, float ref_x, float ref_y, LssdState &s, uint8_t &status) {
    const int valid_ref = ExtractExRefPatch(c, ref, ref_x, ref_y);
    if (valid_ref == 0) {
        status = FTK_STATUS_OUTSIDE;
        return;
    }
    PatchWalk wg = c.walk;
    for (int k = c.g.lane; k < c.geo.psize; k += G, wg.next()) {
        float dx, dy;
        ExGradient(c, wg.row, wg.col, &dx, &dy);
        c.s.dx[k] = dx;
        c.s.dy[k] = dy;
    }
    c.g.sync();
    if (c.p->consider_patch_luminance) {
        // :27-46: interior sum of the extended patch divided by the WHOLE extended patch's valid count.
        const float ref_avg = fdiv(InteriorSum(c, c.s.ex, c.geo.er, c.geo.ec), static_cast<float>(valid_ref));
        for (int k = c.g.lane; k < c.geo.psize; k += G) {
            c.s.dx[k] = fdiv(c.s.dx[k], ref_avg);
            c.s.dy[k] = fdiv(c.s.dy[k], ref_avg);
        }
        for (int e = c.g.lane; e < c.geo.esize; e += G) c.s.ex[e] = fdiv(c.s.ex[e], ref_avg);
        c.g.sync();
    }

    status = FTK_STATUS_LARGE_RESIDUAL;
    float last_squared_step = INFINITY;
    uint32_t large_step_cnt = 0;
    for (uint32_t iter = 0; iter < c.p->max_iteration; ++iter) {
        // ExtractPatchInCurrentImage: the "inside" test truncates and uses a +-patch_rows/cols margin.
        float cx, cy;
        LssdWarp(s, ref_x, ref_y, &cx, &cy);
        const int min_row = static_cast<int>(cy) - c.geo.pr, min_col = static_cast<int>(cx) - c.geo.pc;
        const int max_row = min_row + c.geo.pr * 2, max_col = min_col + c.geo.pc * 2;
        const bool partly_outside = min_row < 0 || max_row > cur.rows - 2 || min_col < 0 || max_col > cur.cols - 2;
        int valid_cur = 0;
        PatchWalk w = c.walk;
        for (int base = 0; base < c.geo.psize; base += G) {
            const int k = base + c.g.lane;
            bool ok = false;
            if (k < c.geo.psize) {
                const int drow = w.row - c.geo.hr, dcol = w.col - c.geo.hc;
                const float row_i = fadd(static_cast<float>(drow), ref_y), col_i = fadd(static_cast<float>(dcol), ref_x);
                float row_j, col_j;
                LssdWarp(s, col_i, row_i, &col_j, &row_j);
                float value = 0.0f;
                if (partly_outside) {
                    ok = PxChecked(cur, row_j, col_j, &value);
                    if (!ok) value = 0.0f;
                } else {
                    value = PxFUnchecked(cur, row_j, col_j);
                    ok = true;
                }
                c.s.curp[k] = value;
                c.s.curv[k] = ok ? 1 : 0;
            }
            valid_cur += c.g.count(ok);
            w.next();
        }
        c.g.sync();
        if (valid_cur == 0) break;
        if (c.p->consider_patch_luminance) {
            // :65-78: interior sum of the cur patch divided by the whole patch's valid count.
            const float cur_avg = fdiv(InteriorSum(c, c.s.curp, c.geo.pr, c.geo.pc), static_cast<float>(valid_cur));
            for (int k = c.g.lane; k < c.geo.psize; k += G) c.s.curp[k] = fdiv(c.s.curp[k], cur_avg);
            c.g.sync();
        }

        // ComputeHessianAndBias: 9 chains.
        int valid = 0;
        c.ch.reset();
        w = c.walk;
        for (int base = 0; base < c.geo.psize; base += G) {
            const int k = base + c.g.lane;
            float t[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) t[q] = 0.0f;
            bool ok = false;
            if (k < c.geo.psize) {
                const int prow = w.row, pcol = w.col;
                const float row_i = fadd(static_cast<float>(w.row - c.geo.hr), ref_y), col_i = fadd(static_cast<float>(w.col - c.geo.hc), ref_x);
                const int e = (prow + 1) * c.geo.ec + pcol + 1;
                ok = (c.s.exv[e] & c.s.curv[k]) != 0;
                if (ok) {
                    const float s0 = fadd(fmul(s.R[0], -row_i), fmul(s.R[1], col_i));
                    const float s1 = fadd(fmul(s.R[2], -row_i), fmul(s.R[3], col_i));
                    float J[3];
                    J[0] = fadd(fmul(c.s.dx[k], s0), fmul(c.s.dy[k], s1));
                    J[1] = c.s.dx[k];
                    J[2] = c.s.dy[k];
                    LssdTerms(J, fsub(c.s.curp[k], c.s.ex[e]), t);
                }
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) c.ch.put(c.g.lane, q, t[q]);
            valid += c.g.count(ok);
            c.ch.template fold<9>(c.g);
            w.next();
        }
        if (valid == 0) break;
        float v[3];
        LssdSolve(c, v);
        if (IsNan(v[0]) || IsNan(v[1]) || IsNan(v[2])) {
            status = FTK_STATUS_NUMERIC_ERROR;
            break;
        }
        LssdUpdate(s, v);
        const float squared_step = fadd(fadd(fmul(v[0], v[0]), fmul(v[1], v[1])), fmul(v[2], v[2]));
        if (FastStepCheck(*c.p, squared_step, last_squared_step, large_step_cnt, status)) break;
    }
}

// ===================================================================================================================
// Kernel: one group per feature; TrackMultipleLevel / TrackSingleLevel of the three subclasses.
// ===================================================================================================================
template <int VARIANT, int METHOD, int G>
__global__ void __launch_bounds__(128, VARIANT == FTK_VARIANT_AFFINE && METHOD == FTK_METHOD_FAST ? 8 : 7) KltKernel(KltLaunch a, SmemLayout layout) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // Several features per warp (affine and small-patch LSSD kInverse on 16 lanes, basic on 8): the groups of a warp run one instruction stream (see Group / AffineTrackOne), so a group without a
    // feature shadows the last one and writes nothing instead of leaving.
    constexpr bool kWholeWarp = ((VARIANT == FTK_VARIANT_AFFINE || (VARIANT == FTK_VARIANT_LSSD && METHOD == kInverse)) && G == 16) ||
                                (VARIANT == FTK_VARIANT_BASIC && G == 8);
    Ctx<G> c{PatchWalk{}, Group<G>(kWholeWarp)};
    const int groups_per_block = blockDim.x / G;
    const int group_in_block = threadIdx.x / G;
    const int f_raw = blockIdx.x * groups_per_block + group_in_block;
    const bool exists = f_raw < a.n_features;
    if (!kWholeWarp && !exists) return;
    const int f = exists ? f_raw : a.n_features - 1;

    c.s = CarveScratch(smem_raw + static_cast<size_t>(group_in_block) * layout.total_bytes, layout);
    c.ch.term = c.s.term;
    c.p = &a.p;
    c.geo.hr = a.p.patch_row_half;
    c.geo.hc = a.p.patch_col_half;
    c.geo.pr = 2 * c.geo.hr + 1;
    c.geo.pc = 2 * c.geo.hc + 1;
    c.geo.psize = c.geo.pr * c.geo.pc;
    c.geo.er = c.geo.pr + 2;
    c.geo.ec = c.geo.pc + 2;
    c.geo.esize = c.geo.er * c.geo.ec;
    c.walk.pc = c.geo.pc;
    c.walk.row = c.g.lane / c.geo.pc;
    c.walk.col = c.g.lane - c.walk.row * c.geo.pc;
    c.walk.step_row = G / c.geo.pc;
    c.walk.step_col = G - c.walk.step_row * c.geo.pc;
    c.ldlt_dst0 = c.ldlt_dst1 = -1;
    if constexpr (VARIANT == FTK_VARIANT_AFFINE) {
        c.ldlt_dst0 = AffineChainSlot(c.g.lane);
        c.ldlt_dst1 = AffineChainSlot(G + c.g.lane);
    }

    const int pair = a.feat_pair ? a.feat_pair[f] : 0;  // null: a single frame pair
    const int local = f - a.feat_offsets[pair];
    const float2 ref_uv = a.ref_uv[f];
    float2 cur_uv = a.has_prediction ? a.cur_uv[f] : ref_uv;  // optical_flow.cpp:12-14
    uint8_t status = a.has_status ? a.status[f] : static_cast<uint8_t>(FTK_STATUS_NOT_TRACKED);  // :17-19

    // basic_klt.cpp:9,12,15: only the first kMaxTrackPointsNumber features, never re-track failed ones.
    const bool tracked = exists && static_cast<uint32_t>(local) < a.p.max_track_points && status <= FTK_STATUS_TRACKED;
    if (kWholeWarp ? !c.g.all(!tracked) : tracked) {
        const int ref_image = a.ref_image ? a.ref_image[pair] : pair;
        const int cur_image = a.cur_image ? a.cur_image[pair] : pair;
        const Img cur0 = LevelImage(a.cur, cur_image, 0);
        if (a.single_level) {
            const Img ref0 = LevelImage(a.ref, ref_image, 0);
            if constexpr (VARIANT == FTK_VARIANT_BASIC) {
                // basic_klt.cpp:59-86
                BasicState s{cur_uv.x, cur_uv.y};
                if constexpr (METHOD == kFast) BasicTrackOneFast<G>(c, ref0, cur0, ref_uv.x, ref_uv.y, s, status, !tracked);
                else BasicTrackOne<METHOD, G>(c, ref0, cur0, ref_uv.x, ref_uv.y, s, status, !tracked);
                if (tracked) cur_uv = make_float2(s.cur_x, s.cur_y);
            } else if constexpr (VARIANT == FTK_VARIANT_AFFINE) {
                // affine_klt.cpp:61-91: starts from predict_affine_
                {
                    AffineState s{cur_uv.x, cur_uv.y, {a.p.predict[0], a.p.predict[1], a.p.predict[2], a.p.predict[3]}};
                    if constexpr (METHOD == kFast) AffineTrackOneFast<G>(c, ref0, cur0, ref_uv.x, ref_uv.y, s, status, !tracked);
                    else AffineTrackOne<METHOD, G>(c, ref0, cur0, ref_uv.x, ref_uv.y, s, status, !tracked);
                    if (tracked) cur_uv = make_float2(s.cur_x, s.cur_y);
                }
            } else {
                // lssd_klt.cpp:63-94: the result is never written back to cur_pixel_uv (reference quirk, kept).
                const float *P = a.p.predict;
                LssdState s;
                s.R[0] = P[0], s.R[1] = P[1], s.R[2] = P[2], s.R[3] = P[3];
                s.t[0] = fsub(cur_uv.x, fadd(fmul(P[0], ref_uv.x), fmul(P[1], ref_uv.y)));
                s.t[1] = fsub(cur_uv.y, fadd(fmul(P[2], ref_uv.x), fmul(P[3], ref_uv.y)));
                if constexpr (METHOD == kFast) LssdTrackOneFast<G>(c, ref0, cur0, ref_uv.x, ref_uv.y, s, status);
                else LssdTrackOne<METHOD, G>(c, ref0, cur0, ref_uv.x, ref_uv.y, s, status, !tracked);
            }
        } else {
            const int levels = a.ref.levels;
            const float scale = static_cast<float>(1 << (levels - 1));
            float sref_x = fdiv(ref_uv.x, scale), sref_y = fdiv(ref_uv.y, scale);
            const float scur_x = fdiv(cur_uv.x, scale), scur_y = fdiv(cur_uv.y, scale);
            if constexpr (VARIANT == FTK_VARIANT_BASIC) {
                // basic_klt.cpp:7-57
                BasicState s{scur_x, scur_y};
                for (int l = levels - 1; l > -1; --l) {
                    const Img ref = LevelImage(a.ref, ref_image, l), cur = LevelImage(a.cur, cur_image, l);
                    if constexpr (METHOD == kFast) BasicTrackOneFast<G>(c, ref, cur, sref_x, sref_y, s, status, !tracked);
                    else BasicTrackOne<METHOD, G>(c, ref, cur, sref_x, sref_y, s, status, !tracked);
                    if (l == 0) break;
                    sref_x = fmul(sref_x, 2.0f), sref_y = fmul(sref_y, 2.0f);
                    s.cur_x = fmul(s.cur_x, 2.0f), s.cur_y = fmul(s.cur_y, 2.0f);
                }
                if (tracked) cur_uv = make_float2(s.cur_x, s.cur_y);
            } else if constexpr (VARIANT == FTK_VARIANT_AFFINE) {
                // affine_klt.cpp:6-59: affine starts at identity and is carried across levels un-scaled.
                {
                    AffineState s{scur_x, scur_y, {1.0f, 0.0f, 0.0f, 1.0f}};
                    for (int l = levels - 1; l > -1; --l) {
                        const Img ref = LevelImage(a.ref, ref_image, l), cur = LevelImage(a.cur, cur_image, l);
                        if constexpr (METHOD == kFast) AffineTrackOneFast<G>(c, ref, cur, sref_x, sref_y, s, status, !tracked);
                        else AffineTrackOne<METHOD, G>(c, ref, cur, sref_x, sref_y, s, status, !tracked);
                        if (l == 0) break;
                        sref_x = fmul(sref_x, 2.0f), sref_y = fmul(sref_y, 2.0f);
                        s.cur_x = fmul(s.cur_x, 2.0f), s.cur_y = fmul(s.cur_y, 2.0f);
                    }
                    if (tracked) cur_uv = make_float2(s.cur_x, s.cur_y);
                }
            } else {
                // lssd_klt.cpp:7-61
                const float *P = a.p.predict;
                LssdState s;
                s.R[0] = P[0], s.R[1] = P[1], s.R[2] = P[2], s.R[3] = P[3];
                s.t[0] = fsub(scur_x, fadd(fmul(P[0], sref_x), fmul(P[1], sref_y)));
                s.t[1] = fsub(scur_y, fadd(fmul(P[2], sref_x), fmul(P[3], sref_y)));
                for (int l = levels - 1; l > -1; --l) {
                    const Img ref = LevelImage(a.ref, ref_image, l), cur = LevelImage(a.cur, cur_image, l);
                    if constexpr (METHOD == kFast) LssdTrackOneFast<G>(c, ref, cur, sref_x, sref_y, s, status);
                    else LssdTrackOne<METHOD, G>(c, ref, cur, sref_x, sref_y, s, status, !tracked);
                    if (l == 0) break;
                    sref_x = fmul(sref_x, 2.0f), sref_y = fmul(sref_y, 2.0f);
                    s.t[0] = fmul(s.t[0], 2.0f), s.t[1] = fmul(s.t[1], 2.0f);
                }
                if (tracked) {
                    cur_uv.x = fadd(fadd(fmul(s.R[0], ref_uv.x), fmul(s.R[1], ref_uv.y)), s.t[0]);
                    cur_uv.y = fadd(fadd(fmul(s.R[2], ref_uv.x), fmul(s.R[3], ref_uv.y)), s.t[1]);
                }
            }
        }
        // final "outside" test on level-0 size (basic_klt.cpp:49-53 and twins)
        if (tracked && IsOutside(cur0, cur_uv.x, cur_uv.y)) status = FTK_STATUS_OUTSIDE;
    }
    if (c.g.lane == 0 && exists) {
        a.cur_uv[f] = cur_uv;
        a.status[f] = status;
    }
}

__global__ void FeaturePairKernel(const int *offsets, int n_pairs, int n_features, int *feat_pair) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_features) return;
    int lo = 0, hi = n_pairs - 1;  // last pair whose offset <= f
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (offsets[mid] <= f) lo = mid;
        else hi = mid - 1;
    }
    feat_pair[f] = lo;
}

template <int VARIANT, int METHOD, int G>
int LaunchOne(ftk_context *ctx, const KltLaunch &a, const Geometry &geo, int threads = 128) {
    const SmemLayout layout = MakeLayout(VARIANT, METHOD, G, geo);
    const int groups_per_block = threads / G;
    const size_t smem = static_cast<size_t>(layout.total_bytes) * groups_per_block;
    if (smem > 227 * 1024) return SetError(ctx, FTK_ERR_UNSUPPORTED, "patch %dx%d needs %zu bytes of shared memory per block", geo.pr, geo.pc, smem);
    auto kernel = KltKernel<VARIANT, METHOD, G>;
    if (smem > 48 * 1024) FTK_CUDA_CHECK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int blocks = (a.n_features + groups_per_block - 1) / groups_per_block;
    kernel<<<blocks, threads, smem, ctx->stream>>>(a, layout);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

template <int VARIANT, int G>
int LaunchMethod(ftk_context *ctx, const KltLaunch &a, const Geometry &geo) {
    switch (a.p.method) {
        case kInverse: return LaunchOne<VARIANT, kInverse, G>(ctx, a, geo);
        case kDirect: return LaunchOne<VARIANT, kDirect, G>(ctx, a, geo);
        default: return LaunchOne<VARIANT, kFast, G>(ctx, a, geo);  // kFast, kSse, kNeon (basic_klt.cpp:31-34 `default:`)
    }
}

}  // namespace

int LaunchFeaturePairs(ftk_context *ctx, const int *d_offsets, int n_pairs, int n_features, int *d_feat_pair) {
    const int threads = 256;
    FeaturePairKernel<<<(n_features + threads - 1) / threads, threads, 0, ctx->stream>>>(d_offsets, n_pairs, n_features, d_feat_pair);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

static int LaunchKltTrackImpl(ftk_context *ctx, const KltLaunch &a);

int LaunchKltTrack(ftk_context *ctx, const KltLaunch &a) {
    ProfBegin(ctx);
    const int rc = LaunchKltTrackImpl(ctx, a);
    ProfEnd(ctx);
    return rc;
}

static int LaunchKltTrackImpl(ftk_context *ctx, const KltLaunch &a) {
    Geometry geo;
    geo.hr = a.p.patch_row_half;
    geo.hc = a.p.patch_col_half;
    geo.pr = 2 * geo.hr + 1;
    geo.pc = 2 * geo.hc + 1;
    geo.psize = geo.pr * geo.pc;
    geo.er = geo.pr + 2;
    geo.ec = geo.pc + 2;
    geo.esize = geo.er * geo.ec;
    if (geo.hr < 0 || geo.hc < 0) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "negative patch half size");
    if (ctx->use_fast_paths) {
        const int rc = LaunchKltBasicFastPath(ctx, a);
        if (rc != FTK_ERR_UNSUPPORTED) return rc;
    }
    switch (a.p.variant) {
        case FTK_VARIANT_BASIC:
            if (geo.psize <= 8 * 64) return LaunchMethod<FTK_VARIANT_BASIC, 8>(ctx, a, geo);
            if (geo.psize <= 32 * 64) return LaunchMethod<FTK_VARIANT_BASIC, 32>(ctx, a, geo);
            break;
        case FTK_VARIANT_AFFINE:
            // 16 lanes per feature: two features share a warp's fold / LDLT instructions and 13x13 patches fill 11 chunks of 16 to 96 %.  Both
            // groups run one instruction stream (Group whole-warp mode), which is what makes this pay: with per-group masks the uniformity
            // checks around every barrier / vote made kFast and kInverse slower on 16 lanes than on 32 (43.8 vs 35.9 ms per 2 M features).
            // Measured per 200 k features: kDirect (16 lanes before, per-group masks) 3.71 -> 3.33 ms; kFast 32 -> 16 lanes 3.37 -> 3.11 ms, kInverse
            // 3.91 -> 3.41 ms.
            // kFast was then limited by shared memory (5.2 KB per feature: 20 warps per SM): its term buffer now holds 8 chains instead of the 16
            // pairs kDirect needs (the 21 Hessian chains of a level pass through it in three parts) = 3.4 KB per feature, 32 warps per SM with
            // 128-thread CTAs: 3.11 -> 2.77 ms.  kDirect / kInverse are register limited; 64 / 32 threads per CTA balance best for them
            // (128 / 64 / 32 threads: kDirect 3.36 / 3.33 / 3.45 ms, kInverse 3.47 / 3.48 / 3.41 ms).
            if (geo.psize <= 16 * 64) {
                switch (a.p.method) {
                    case kDirect: return LaunchOne<FTK_VARIANT_AFFINE, kDirect, 16>(ctx, a, geo, 64);
                    case kInverse: return LaunchOne<FTK_VARIANT_AFFINE, kInverse, 16>(ctx, a, geo, 32);
                    default: return LaunchOne<FTK_VARIANT_AFFINE, kFast, 16>(ctx, a, geo, 128);
                }
            }
            if (geo.psize <= 32 * 64) return LaunchMethod<FTK_VARIANT_AFFINE, 32>(ctx, a, geo);
            break;
        case FTK_VARIANT_LSSD:
            // kInverse on 16 lanes per feature (whole-warp mode, one warp per CTA) while the hoisted samples of two features per warp still
            // leave enough warps on an SM: per 200 k features 13x13 6.15 -> 5.58 ms, 15x15 7.14 -> 6.88 ms, but 17x17 8.28 -> 8.75 ms and
            // 21x21 (BASELINE configs[2]) 14.2 -> 18.2 ms (7 KB of hoisted samples per feature: 13 warps per SM)
            if (a.p.method == kInverse && geo.psize <= 15 * 15) return LaunchOne<FTK_VARIANT_LSSD, kInverse, 16>(ctx, a, geo, 32);
            // One warp per CTA: a CTA's shared memory is released when its slowest feature finishes, and iteration counts differ a lot between
            // features (21x21, 200 k features: 128 / 96 / 64 / 32 threads per CTA = 14.15 / 13.84 / 13.44 / 12.93 ms)
            if (a.p.method == kInverse && geo.psize <= 32 * 64) return LaunchOne<FTK_VARIANT_LSSD, kInverse, 32>(ctx, a, geo, 32);
            if (geo.psize <= 32 * 64) return LaunchMethod<FTK_VARIANT_LSSD, 32>(ctx, a, geo);
            break;
        default:
            return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "unknown tracker variant %d", a.p.variant);
    }
    return SetError(ctx, FTK_ERR_UNSUPPORTED, "patch of %d pixels is larger than this build supports (2048)", geo.psize);
}

namespace {

// Forward-backward consistency test (ftk_klt_params::forward_backward_max_error): see include/ftk_c.h.
__global__ void ForwardBackwardCheckKernel(int n_features, const float2 *ref_uv, const float2 *back_uv, const uint8_t *back_status, uint8_t *status,
                                           float max_sq) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_features || status[f] != FTK_STATUS_TRACKED) return;
    const float dx = fsub(back_uv[f].x, ref_uv[f].x), dy = fsub(back_uv[f].y, ref_uv[f].y);
    const float d2 = fadd(fmul(dx, dx), fmul(dy, dy));
    if (!(back_status[f] == FTK_STATUS_TRACKED && d2 <= max_sq)) status[f] = FTK_STATUS_LARGE_RESIDUAL;
}

}  // namespace

int LaunchKltTrackChecked(ftk_context *ctx, const KltLaunch &a, size_t scratch_offset) {
    if (int rc = LaunchKltTrack(ctx, a)) return rc;
    const float max_error = a.p.forward_backward_max_error;
    if (!(max_error > 0.0f)) return FTK_OK;
    const size_t n = static_cast<size_t>(a.n_features);
    if (int rc = EnsureDevice(ctx, ctx->d_back_uv, sizeof(float2) * (scratch_offset + n))) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_back_status, scratch_offset + n)) return rc;
    float2 *back_uv = static_cast<float2 *>(ctx->d_back_uv.ptr) + scratch_offset;
    uint8_t *back_status = static_cast<uint8_t *>(ctx->d_back_status.ptr) + scratch_offset;
    // TrackFeatures(cur_pyramid, ref_pyramid, ref := forward result, cur := reference position (prediction), status := forward status)
    FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(back_uv, a.ref_uv, sizeof(float2) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(back_status, a.status, n, cudaMemcpyDeviceToDevice, ctx->stream));
    KltLaunch b = a;
    b.ref = a.cur;
    b.cur = a.ref;
    b.ref_image = a.cur_image;
    b.cur_image = a.ref_image;
    b.ref_uv = a.cur_uv;
    b.cur_uv = back_uv;
    b.status = back_status;
    b.has_prediction = 1;
    b.has_status = 1;
    b.p.predict[0] = 1.0f, b.p.predict[1] = 0.0f, b.p.predict[2] = 0.0f, b.p.predict[3] = 1.0f;
    if (int rc = LaunchKltTrack(ctx, b)) return rc;
    const int threads = 256;
    ForwardBackwardCheckKernel<<<(a.n_features + threads - 1) / threads, threads, 0, ctx->stream>>>(a.n_features, a.ref_uv, back_uv, back_status, a.status,
                                                                                                   max_error * max_error);
    ++ctx->launches;
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    return FTK_OK;
}

}  // namespace ftk
