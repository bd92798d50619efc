/* ORACLE (test infrastructure -- see ftk_oracle.h for the rules and the pinning statement).
 *
 * Plain-C sequential restatement of the reference hot path.  Reference paths are relative to
 * /root/reference/src/ ; "OF/" abbreviates optical_flow_tracker/.
 *
 * Arithmetic contract: IEEE-754 binary32, round-to-nearest, NO fused multiply-add (build with
 * -ffp-contract=off), every sum evaluated left to right in row-major pixel order -- exactly what the
 * reference does when compiled with its own flags (CMakeLists.txt:6).  */
#include "ftk_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { ST_NOT_TRACKED = 0, ST_TRACKED = 1, ST_LARGE_RESIDUAL = 2, ST_OUTSIDE = 3, ST_NUMERIC_ERROR = 4 }; /* feature_tracker.h:8-14 */
enum { M_INVERSE = 0, M_DIRECT = 1, M_FAST = 2 };                                                          /* OF/optical_flow.h:12-18 */

/* ------------------------------------------------------------------------------------------------------------
 * GrayImage (external Slam_Utility type; semantics frozen in oracle/shim/datatype_image.h)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t *d;
    int32_t rows, cols;
} image_t;

static inline float px_i(const image_t *im, int32_t row, int32_t col) { return (float)im->d[row * im->cols + col]; }

/* Diagnostics only (tools/fuzz_parity.py): how many unchecked samples addressed a base pixel outside the image.  The
 * reference has undefined behaviour there (lssd_klt_fast.cpp:182-193 once a track diverges); results are untouched. */
static long long g_outside_reads = 0;
long long ftko_outside_reads(int32_t reset) {
    const long long n = g_outside_reads;
    if (reset) g_outside_reads = 0;
    return n;
}

/* Unchecked bilinear sample: base pixel by truncation, fractions by floor, 4 weighted terms left to right. */
static inline float px_f(const image_t *im, float row, float col) {
    if (!(row >= 0.0f && col >= 0.0f && (int32_t)row <= im->rows - 1 && (int32_t)col <= im->cols - 1)) ++g_outside_reads;
    const uint8_t *v = &im->d[(int32_t)row * im->cols + (int32_t)col];
    const float sr = row - floorf(row);
    const float sc = col - floorf(col);
    const float ir = 1.0f - sr;
    const float ic = 1.0f - sc;
    return ic * ir * (float)v[0] + sc * ir * (float)v[1] + ic * sr * (float)v[im->cols] + sc * sr * (float)v[im->cols + 1];
}

/* Checked sample: fails iff the position lies outside [0, cols-1] x [0, rows-1]. */
static inline int px_checked(const image_t *im, float row, float col, float *out) {
    if (col < 0 || row < 0 || col > (float)(im->cols - 1) || row > (float)(im->rows - 1)) return 0;
    *out = px_f(im, row, col);
    return 1;
}

/* ------------------------------------------------------------------------------------------------------------
 * LDLT solve (external Eigen; restated in oracle/shim/basic_type.h).  n <= 6.
 * Call sites: OF/basic_klt/optical_flow_basic_klt.cpp:97, ..._fast.cpp:39, OF/affine_klt/...:103, ..._fast.cpp:41,
 * OF/lssd_klt/...:107, ..._fast.cpp:88.
 * ---------------------------------------------------------------------------------------------------------- */
static void swapf(float *p, float *q) {
    const float t = *p;
    *p = *q;
    *q = t;
}

static void ldlt_solve(int n, const float *A, const float *b, float *x) {
    float a[6][6];
    int tr[6];
    float temp[6];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) a[i][j] = A[i * n + j];

    int whole_diagonal_zero = 0;
    for (int k = 0; k < n && !whole_diagonal_zero; ++k) {
        int p = k;
        float best = fabsf(a[k][k]);
        for (int i = k + 1; i < n; ++i) {
            const float v = fabsf(a[i][i]);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        tr[k] = p;
        if (p != k) {
            for (int j = 0; j < k; ++j) swapf(&a[k][j], &a[p][j]);
            for (int i = p + 1; i < n; ++i) swapf(&a[i][k], &a[i][p]);
            swapf(&a[k][k], &a[p][p]);
            for (int i = k + 1; i < p; ++i) swapf(&a[i][k], &a[p][i]);
        }
        if (k > 0) {
            for (int j = 0; j < k; ++j) temp[j] = a[j][j] * a[k][j];
            float s = a[k][0] * temp[0];
            for (int j = 1; j < k; ++j) s = s + a[k][j] * temp[j];
            a[k][k] = a[k][k] - s;
            for (int i = k + 1; i < n; ++i) {
                float t = a[i][0] * temp[0];
                for (int j = 1; j < k; ++j) t = t + a[i][j] * temp[j];
                a[i][k] = a[i][k] - t;
            }
        }
        const float akk = a[k][k];
        const int pivot_ok = fabsf(akk) > 0.0f;
        if (k == 0 && !pivot_ok) {
            for (int j = 0; j < n; ++j) tr[j] = j;
            whole_diagonal_zero = 1;
            break;
        }
        if (pivot_ok)
            for (int i = k + 1; i < n; ++i) a[i][k] = a[i][k] / akk;
    }

    for (int i = 0; i < n; ++i) x[i] = b[i];
    for (int k = 0; k < n; ++k) swapf(&x[k], &x[tr[k]]);
    for (int i = 1; i < n; ++i) {
        float s = a[i][0] * x[0];
        for (int j = 1; j < i; ++j) s = s + a[i][j] * x[j];
        x[i] = x[i] - s;
    }
    for (int i = 0; i < n; ++i) x[i] = (fabsf(a[i][i]) > FLT_MIN) ? x[i] / a[i][i] : 0.0f;
    for (int i = n - 2; i >= 0; --i) {
        float s = a[i + 1][i] * x[i + 1];
        for (int j = i + 2; j < n; ++j) s = s + a[j][i] * x[j];
        x[i] = x[i] - s;
    }
    for (int k = n - 1; k >= 0; --k) swapf(&x[k], &x[tr[k]]);
}

void ftko_ldlt_solve(int32_t n, const float *a, const float *b, float *x) { ldlt_solve(n, a, b, x); }

/* ------------------------------------------------------------------------------------------------------------
 * Image pyramid (external Slam_Utility ImagePyramid::CreateImagePyramid; shim: datatype_image_pyramid.h).
 * Call sites: test/test_optical_flow.cpp:70-71.
 * ---------------------------------------------------------------------------------------------------------- */
int ftko_pyramid_build(const uint8_t *image, int32_t rows, int32_t cols, int32_t levels, uint8_t *out) {
    const uint8_t *src = image;
    int32_t sr = rows, sc = cols;
    for (int32_t l = 1; l < levels; ++l) {
        const int32_t dr = sr >> 1, dc = sc >> 1;
        for (int32_t r = 0; r < dr; ++r) {
            for (int32_t c = 0; c < dc; ++c) {
                const uint16_t sum = (uint16_t)((uint16_t)src[(2 * r) * sc + 2 * c] + (uint16_t)src[(2 * r + 1) * sc + 2 * c] +
                                                (uint16_t)src[(2 * r) * sc + 2 * c + 1] + (uint16_t)src[(2 * r + 1) * sc + 2 * c + 1]);
                out[r * dc + c] = (uint8_t)(sum >> 2);
            }
        }
        src = out;
        out += (size_t)dr * dc;
        sr = dr;
        sc = dc;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------------------
 * Tracker context: options + patch geometry (OF/optical_flow.cpp:104-124 PrepareForTracking) + scratch.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    ftko_klt_params o;
    int32_t patch_rows, patch_cols, patch_size;
    int32_t ex_rows, ex_cols, ex_size;
    float *ex_patch;        /* ex_size */
    uint8_t *ex_valid;      /* ex_size */
    float *dx, *dy;         /* patch_size */
    float *cur_patch;       /* patch_size */
    uint8_t *cur_valid;     /* patch_size */
    int32_t *pixel_valid;   /* patch_size (lssd two-pass mask) */
    int32_t iterations;     /* trace: GN iterations executed for the current feature */
} tracker_t;

static int tracker_init(tracker_t *t, const ftko_klt_params *p) {
    memset(t, 0, sizeof(*t));
    t->o = *p;
    t->patch_rows = (p->patch_row_half << 1) + 1;
    t->patch_cols = (p->patch_col_half << 1) + 1;
    t->patch_size = t->patch_rows * t->patch_cols;
    t->ex_rows = t->patch_rows + 2;
    t->ex_cols = t->patch_cols + 2;
    t->ex_size = t->ex_rows * t->ex_cols;
    if (t->patch_size <= 0) return 0;
    t->ex_patch = (float *)malloc(sizeof(float) * t->ex_size);
    t->ex_valid = (uint8_t *)malloc(t->ex_size);
    t->dx = (float *)malloc(sizeof(float) * t->patch_size);
    t->dy = (float *)malloc(sizeof(float) * t->patch_size);
    t->cur_patch = (float *)malloc(sizeof(float) * t->patch_size);
    t->cur_valid = (uint8_t *)malloc(t->patch_size);
    t->pixel_valid = (int32_t *)malloc(sizeof(int32_t) * t->patch_size);
    return 1;
}

static void tracker_free(tracker_t *t) {
    free(t->ex_patch);
    free(t->ex_valid);
    free(t->dx);
    free(t->dy);
    free(t->cur_patch);
    free(t->cur_valid);
    free(t->pixel_valid);
}

static inline int is_outside(const image_t *im, float x, float y) { return x < 0 || x > (float)(im->cols - 1) || y < 0 || y > (float)(im->rows - 1); }

/* OF/optical_flow.cpp:49-102 ExtractExtendPatchInReferenceImage: (2h+3)^2 integer-aligned window at floor(ref),
 * ONE set of bilinear weights, pixel valid iff 0<=row<=rows-2 && 0<=col<=cols-2.  Returns the valid count. */
static uint32_t extract_ex_ref_patch(tracker_t *t, const image_t *ref, float ref_x, float ref_y) {
    const float int_row = floorf(ref_y), int_col = floorf(ref_x);
    const float dec_row = ref_y - int_row, dec_col = ref_x - int_col;
    const float w_tl = (1.0f - dec_row) * (1.0f - dec_col);
    const float w_tr = (1.0f - dec_row) * dec_col;
    const float w_bl = dec_row * (1.0f - dec_col);
    const float w_br = dec_row * dec_col;
    const int32_t min_row = (int32_t)int_row - t->ex_rows / 2;
    const int32_t min_col = (int32_t)int_col - t->ex_cols / 2;
    uint32_t valid = 0;
    int k = 0;
    for (int32_t row = min_row; row < min_row + t->ex_rows; ++row) {
        for (int32_t col = min_col; col < min_col + t->ex_cols; ++col, ++k) {
            if (row < 0 || row > ref->rows - 2 || col < 0 || col > ref->cols - 2) {
                t->ex_valid[k] = 0;
                t->ex_patch[k] = 0.0f;
            } else {
                t->ex_valid[k] = 1;
                t->ex_patch[k] = w_tl * px_i(ref, row, col) + w_tr * px_i(ref, row, col + 1) + w_bl * px_i(ref, row + 1, col) + w_br * px_i(ref, row + 1, col + 1);
                ++valid;
            }
        }
    }
    return valid;
}

/* Gradient of the extended patch at interior pixel (row, col): valid only if the 4 neighbours are
 * (OF/basic_klt/optical_flow_basic_klt_fast.cpp:71-94, affine_klt_fast.cpp:78-91, lssd_klt_fast.cpp:122-141). */
static inline int ex_gradient(const tracker_t *t, int32_t row, int32_t col, float *dx, float *dy) {
    const int32_t e = (row + 1) * t->ex_cols + col + 1;
    const int32_t l = e - 1, r = e + 1, u = e - t->ex_cols, d = e + t->ex_cols;
    if (t->ex_valid[l] && t->ex_valid[r] && t->ex_valid[u] && t->ex_valid[d]) {
        *dx = t->ex_patch[r] - t->ex_patch[l];
        *dy = t->ex_patch[d] - t->ex_patch[u];
        return 1;
    }
    *dx = 0.0f;
    *dy = 0.0f;
    return 0;
}

/* ============================================================================================================
 * BASIC KLT
 * ========================================================================================================== */

/* OF/basic_klt/optical_flow_basic_klt.cpp:118-181 ConstructIncrementalFunction. H = {h00, h01, h11}. */
static int32_t basic_construct(const tracker_t *t, const image_t *ref, const image_t *cur, float ref_x, float ref_y, float cur_x, float cur_y, float *H,
                               float *b) {
    float v[6];
    int32_t valid = 0;
    const int inverse = t->o.method == M_INVERSE;
    for (int32_t drow = -t->o.patch_row_half; drow <= t->o.patch_row_half; ++drow) {
        for (int32_t dcol = -t->o.patch_col_half; dcol <= t->o.patch_col_half; ++dcol) {
            const float row_i = (float)drow + ref_y, col_i = (float)dcol + ref_x;
            const float row_j = (float)drow + cur_y, col_j = (float)dcol + cur_x;
            const image_t *g = inverse ? ref : cur; /* image the gradient is taken from */
            const float gr = inverse ? row_i : row_j, gc = inverse ? col_i : col_j;
            if (px_checked(g, gr, gc - 1.0f, &v[0]) && px_checked(g, gr, gc + 1.0f, &v[1]) && px_checked(g, gr - 1.0f, gc, &v[2]) &&
                px_checked(g, gr + 1.0f, gc, &v[3]) && px_checked(ref, row_i, col_i, &v[4]) && px_checked(cur, row_j, col_j, &v[5])) {
                const float fx = v[1] - v[0], fy = v[3] - v[2], ft = v[5] - v[4];
                H[0] += fx * fx;
                H[2] += fy * fy;
                H[1] += fx * fy;
                b[0] -= fx * ft;
                b[1] -= fy * ft;
                ++valid;
            }
        }
    }
    return valid;
}

/* OF/basic_klt/optical_flow_basic_klt.cpp:88-116 TrackOneFeature. */
static void basic_track_one(tracker_t *t, const image_t *ref, const image_t *cur, float ref_x, float ref_y, float *cur_x, float *cur_y, uint8_t *status) {
    for (uint32_t iter = 0; iter < t->o.max_iteration; ++iter) {
        float H[3] = {0.0f, 0.0f, 0.0f}, b[2] = {0.0f, 0.0f};
        ++t->iterations;
        if (basic_construct(t, ref, cur, ref_x, ref_y, *cur_x, *cur_y, H, b) == 0) break;
        const float A[4] = {H[0], H[1], H[1], H[2]};
        float v[2];
        ldlt_solve(2, A, b, v);
        if (isnan(v[0]) || isnan(v[1])) {
            *status = ST_NUMERIC_ERROR;
            break;
        }
        *cur_x += v[0];
        *cur_y += v[1];
        if (is_outside(cur, *cur_x, *cur_y)) {
            *status = ST_OUTSIDE;
            break;
        }
        if (v[0] * v[0] + v[1] * v[1] < t->o.max_converge_step) {
            *status = ST_TRACKED;
            break;
        }
    }
}

/* OF/basic_klt/optical_flow_basic_klt_fast.cpp:64-99 PrecomputeJacobianAndHessian. */
static void basic_fast_precompute(tracker_t *t, float *H) {
    H[0] = H[1] = H[2] = 0.0f;
    for (int32_t row = 0; row < t->patch_rows; ++row) {
        for (int32_t col = 0; col < t->patch_cols; ++col) {
            float dx, dy;
            if (ex_gradient(t, row, col, &dx, &dy)) {
                H[0] += dx * dx;
                H[1] += dx * dy;
                H[2] += dy * dy;
            }
            t->dx[row * t->patch_cols + col] = dx;
            t->dy[row * t->patch_cols + col] = dy;
        }
    }
}

/* OF/basic_klt/optical_flow_basic_klt_fast.cpp:101-195 ComputeBias: integer-aligned (2h+1)^2 window at
 * floor(cur) with ONE weight set; a pixel contributes iff it is inside [0,rows-2]x[0,cols-2] of cur and valid
 * in the extended ref patch.  (The reference's "totally inside" branch is the same loop without the cur test,
 * which is vacuous there.) */
static int32_t basic_fast_bias(const tracker_t *t, const image_t *cur, float cur_x, float cur_y, float *b) {
    b[0] = b[1] = 0.0f;
    const float int_row = floorf(cur_y), int_col = floorf(cur_x);
    const float dec_row = cur_y - int_row, dec_col = cur_x - int_col;
    const float w_tl = (1.0f - dec_row) * (1.0f - dec_col);
    const float w_tr = (1.0f - dec_row) * dec_col;
    const float w_bl = dec_row * (1.0f - dec_col);
    const float w_br = dec_row * dec_col;
    const int32_t min_row = (int32_t)int_row - t->patch_rows / 2;
    const int32_t min_col = (int32_t)int_col - t->patch_cols / 2;
    int32_t valid = 0;
    for (int32_t row = min_row; row < min_row + t->patch_rows; ++row) {
        for (int32_t col = min_col; col < min_col + t->patch_cols; ++col) {
            if (row < 0 || row > cur->rows - 2 || col < 0 || col > cur->cols - 2) continue;
            const int32_t pr = row - min_row, pc = col - min_col;
            const int32_t e = (pr + 1) * t->ex_cols + pc + 1;
            if (!t->ex_valid[e]) continue;
            const float cur_value = w_tl * px_i(cur, row, col) + w_tr * px_i(cur, row, col + 1) + w_bl * px_i(cur, row + 1, col) + w_br * px_i(cur, row + 1, col + 1);
            const float dt = cur_value - t->ex_patch[e];
            const int32_t k = pr * t->patch_cols + pc;
            b[0] -= t->dx[k] * dt;
            b[1] -= t->dy[k] * dt;
            ++valid;
        }
    }
    return valid;
}

/* Shared tail of every fast tracker: step bookkeeping (basic_klt_fast.cpp:48-60, affine_klt_fast.cpp:55-67,
 * lssd_klt_fast.cpp:101-112).  Returns 1 when the iteration loop must stop. */
static int fast_step_check(const tracker_t *t, float squared_step, float *last_squared_step, uint32_t *large_step_cnt, uint8_t *status) {
    if (squared_step < *last_squared_step) {
        *last_squared_step = squared_step;
        *large_step_cnt = 0;
    } else {
        ++*large_step_cnt;
        if (*large_step_cnt >= t->o.max_tolerance_large_step) return 1;
    }
    if (squared_step < t->o.max_converge_step) {
        *status = ST_TRACKED;
        return 1;
    }
    return 0;
}

/* OF/basic_klt/optical_flow_basic_klt_fast.cpp:7-62 TrackOneFeatureFast. */
static void basic_track_one_fast(tracker_t *t, const image_t *ref, const image_t *cur, float ref_x, float ref_y, float *cur_x, float *cur_y,
                                 uint8_t *status) {
    if (extract_ex_ref_patch(t, ref, ref_x, ref_y) == 0) {
        *status = ST_OUTSIDE;
        return;
    }
    float H[3];
    basic_fast_precompute(t, H);
    const float A[4] = {H[0], H[1], H[1], H[2]};
    *status = ST_LARGE_RESIDUAL;
    float last_squared_step = INFINITY;
    uint32_t large_step_cnt = 0;
    for (uint32_t iter = 0; iter < t->o.max_iteration; ++iter) {
        float b[2], v[2];
        ++t->iterations;
        if (basic_fast_bias(t, cur, *cur_x, *cur_y, b) == 0) break;
        ldlt_solve(2, A, b, v);
        if (isnan(v[0]) || isnan(v[1])) {
            *status = ST_NUMERIC_ERROR;
            break;
        }
        *cur_x += v[0];
        *cur_y += v[1];
        const float squared_step = v[0] * v[0] + v[1] * v[1];
        if (fast_step_check(t, squared_step, &last_squared_step, &large_step_cnt, status)) break;
    }
}

/* ============================================================================================================
 * AFFINE KLT.  affine = {a00, a01, a10, a11} row-major.
 * ========================================================================================================== */

static const int kAffIdx[21][2] = {{0, 0}, {0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {1, 1}, {1, 2}, {1, 3}, {1, 4}, {1, 5},
                                   {2, 2}, {2, 3}, {2, 4}, {2, 5}, {3, 3}, {3, 4}, {3, 5}, {4, 4}, {4, 5}, {5, 5}};

/* OF/affine_klt/optical_flow_affine_klt.cpp:131-273 ConstructIncrementalFunction (H is 6x6 row-major).
 * Note the reference's (3,4) entry accumulates yy*dxdy (:183,245) -- reproduced. */
static int32_t affine_construct(const tracker_t *t, const image_t *ref, const image_t *cur, float ref_x, float ref_y, float cur_x, float cur_y,
                                const float *affine, float *H, float *b) {
    float v[6];
    float u[21];
    int32_t valid = 0;
    for (int i = 0; i < 21; ++i) u[i] = 0.0f;
    for (int i = 0; i < 6; ++i) b[i] = 0.0f;
    const int direct = t->o.method == M_DIRECT;
    for (int32_t drow = -t->o.patch_row_half; drow <= t->o.patch_row_half; ++drow) {
        for (int32_t dcol = -t->o.patch_col_half; dcol <= t->o.patch_col_half; ++dcol) {
            const float row_i = (float)drow + ref_y, col_i = (float)dcol + ref_x;
            const float ax = affine[0] * (float)dcol + affine[1] * (float)drow;
            const float ay = affine[2] * (float)dcol + affine[3] * (float)drow;
            const float row_j = ay + cur_y, col_j = ax + cur_x;
            const image_t *g = direct ? cur : ref;
            const float gr = direct ? row_j : row_i, gc = direct ? col_j : col_i;
            if (px_checked(g, gr, gc - 1.0f, &v[0]) && px_checked(g, gr, gc + 1.0f, &v[1]) && px_checked(g, gr - 1.0f, gc, &v[2]) &&
                px_checked(g, gr + 1.0f, gc, &v[3]) && px_checked(ref, row_i, col_i, &v[4]) && px_checked(cur, row_j, col_j, &v[5])) {
                const float dx = v[1] - v[0], dy = v[3] - v[2], dt = v[5] - v[4];
                const float x = col_j, y = row_j;
                const float xx = x * x, yy = y * y, dxdx = dx * dx, dydy = dy * dy, xy = x * y, dxdy = dx * dy;
                u[0] += xx * dxdx;  /* (0,0) */
                u[1] += xx * dxdy;  /* (0,1) */
                u[2] += xy * dxdx;  /* (0,2) */
                u[3] += xy * dxdy;  /* (0,3) */
                u[4] += x * dxdx;   /* (0,4) */
                u[5] += x * dxdy;   /* (0,5) */
                u[6] += xx * dydy;  /* (1,1) */
                u[7] += xy * dxdy;  /* (1,2) */
                u[8] += xy * dydy;  /* (1,3) */
                u[9] += x * dxdy;   /* (1,4) */
                u[10] += x * dydy;  /* (1,5) */
                u[11] += yy * dxdx; /* (2,2) */
                u[12] += yy * dxdy; /* (2,3) */
                u[13] += y * dxdx;  /* (2,4) */
                u[14] += y * dxdy;  /* (2,5) */
                u[15] += yy * dydy; /* (3,3) */
                u[16] += yy * dxdy; /* (3,4) sic */
                u[17] += y * dydy;  /* (3,5) */
                u[18] += dxdx;      /* (4,4) */
                u[19] += dxdy;      /* (4,5) */
                u[20] += dydy;      /* (5,5) */
                b[0] -= dt * x * dx;
                b[1] -= dt * x * dy;
                b[2] -= dt * y * dx;
                b[3] -= dt * y * dy;
                b[4] -= dt * dx;
                b[5] -= dt * dy;
                ++valid;
            }
        }
    }
    for (int i = 0; i < 21; ++i) {
        H[kAffIdx[i][0] * 6 + kAffIdx[i][1]] = u[i];
        H[kAffIdx[i][1] * 6 + kAffIdx[i][0]] = u[i];
    }
    return valid;
}

/* OF/affine_klt/optical_flow_affine_klt.cpp:93-129 TrackOneFeature. */
static void affine_track_one(tracker_t *t, const image_t *ref, const image_t *cur, float ref_x, float ref_y, float *cur_x, float *cur_y, float *affine,
                             uint8_t *status) {
    float H[36], b[6], z[6];
    for (uint32_t iter = 0; iter < t->o.max_iteration; ++iter) {
        ++t->iterations;
        if (affine_construct(t, ref, cur, ref_x, ref_y, *cur_x, *cur_y, affine, H, b) == 0) break;
        ldlt_solve(6, H, b, z);
        const float v0 = z[0] * *cur_x + z[2] * *cur_y + z[4];
        const float v1 = z[1] * *cur_x + z[3] * *cur_y + z[5];
        if (isnan(v0) || isnan(v1)) {
            *status = ST_NUMERIC_ERROR;
            break;
        }
        *cur_x += v0;
        *cur_y += v1;
        affine[0] += z[0]; /* col(0) += z.head<2>() */
        affine[2] += z[1];
        affine[1] += z[2]; /* col(1) += z.segment<2>(2) */
        affine[3] += z[3];
        if (is_outside(cur, *cur_x, *cur_y)) {
            *status = ST_OUTSIDE;
            break;
        }
        if (v0 * v0 + v1 * v1 < t->o.max_converge_step) {
            *status = ST_TRACKED;
            break;
        }
    }
}

/* OF/affine_klt/optical_flow_affine_klt_fast.cpp:71-138 PrecomputeJacobianAndHessian: 18 accumulated entries,
 * (1,2)=(0,3), (1,4)=(0,5), (3,4)=(2,3) copied afterwards; x,y use cur_pixel_uv frozen at level entry. */
static void affine_fast_precompute(tracker_t *t, float cur_x, float cur_y, float *H) {
    float u[21];
    for (int i = 0; i < 21; ++i) u[i] = 0.0f;
    for (int32_t row = 0; row < t->patch_rows; ++row) {
        for (int32_t col = 0; col < t->patch_cols; ++col) {
            float dx, dy;
            if (ex_gradient(t, row, col, &dx, &dy)) {
                const float x = (float)(col - t->o.patch_col_half) + cur_x;
                const float y = (float)(row - t->o.patch_row_half) + cur_y;
                const float xx = x * x, yy = y * y, xy = x * y, dxdx = dx * dx, dydy = dy * dy, dxdy = dx * dy;
                u[0] += xx * dxdx;
                u[1] += xx * dxdy;
                u[2] += xy * dxdx;
                u[3] += xy * dxdy;
                u[4] += x * dxdx;
                u[5] += x * dxdy;
                u[6] += xx * dydy;
                u[8] += xy * dydy;
                u[10] += x * dydy;
                u[11] += yy * dxdx;
                u[12] += yy * dxdy;
                u[13] += y * dxdx;
                u[14] += y * dxdy;
                u[15] += yy * dydy;
                u[17] += y * dydy;
                u[18] += dxdx;
                u[19] += dxdy;
                u[20] += dydy;
            }
            t->dx[row * t->patch_cols + col] = dx;
            t->dy[row * t->patch_cols + col] = dy;
        }
    }
    u[7] = u[3];   /* (1,2) = (0,3) */
    u[9] = u[5];   /* (1,4) = (0,5) */
    u[16] = u[12]; /* (3,4) = (2,3) */
    for (int i = 0; i < 21; ++i) {
        H[kAffIdx[i][0] * 6 + kAffIdx[i][1]] = u[i];
        H[kAffIdx[i][1] * 6 + kAffIdx[i][0]] = u[i];
    }
}

/* OF/affine_klt/optical_flow_affine_klt_fast.cpp:140-188 ComputeBias: one CHECKED bilinear sample per pixel at the
 * affinely warped position; the count only includes pixels also valid in the extended ref patch. */
static int32_t affine_fast_bias(const tracker_t *t, const image_t *cur, float cur_x, float cur_y, const float *affine, float *b) {
    int32_t valid = 0;
    for (int i = 0; i < 6; ++i) b[i] = 0.0f;
    for (int32_t drow = -t->o.patch_row_half; drow <= t->o.patch_row_half; ++drow) {
        for (int32_t dcol = -t->o.patch_col_half; dcol <= t->o.patch_col_half; ++dcol) {
            const float ax = affine[0] * (float)dcol + affine[1] * (float)drow;
            const float ay = affine[2] * (float)dcol + affine[3] * (float)drow;
            const float row_c = ay + cur_y, col_c = ax + cur_x;
            float cur_value = 0.0f;
            if (!px_checked(cur, row_c, col_c, &cur_value)) continue;
            const int32_t er = drow + t->o.patch_row_half + 1, ec = dcol + t->o.patch_col_half + 1;
            const int32_t e = er * t->ex_cols + ec;
            if (!t->ex_valid[e]) continue;
            const float dt = cur_value - t->ex_patch[e];
            const int32_t k = (er - 1) * t->patch_cols + (ec - 1);
            const float dx = t->dx[k], dy = t->dy[k];
            b[0] -= dt * col_c * dx;
            b[1] -= dt * col_c * dy;
            b[2] -= dt * row_c * dx;
            b[3] -= dt * row_c * dy;
            b[4] -= dt * dx;
            b[5] -= dt * dy;
            ++valid;
        }
    }
    return valid;
}

/* OF/affine_klt/optical_flow_affine_klt_fast.cpp:7-69 TrackOneFeatureFast. */
static void affine_track_one_fast(tracker_t *t, const image_t *ref, const image_t *cur, float ref_x, float ref_y, float *cur_x, float *cur_y,
                                  float *affine, uint8_t *status) {
    if (extract_ex_ref_patch(t, ref, ref_x, ref_y) == 0) {
        *status = ST_OUTSIDE;
        return;
    }
    float H[36], b[6], z[6];
    affine_fast_precompute(t, *cur_x, *cur_y, H);
    float last_squared_step = INFINITY;
    uint32_t large_step_cnt = 0;
    *status = ST_LARGE_RESIDUAL;
    for (uint32_t iter = 0; iter < t->o.max_iteration; ++iter) {
        ++t->iterations;
        if (affine_fast_bias(t, cur, *cur_x, *cur_y, affine, b) == 0) break;
        ldlt_solve(6, H, b, z);
        int any_nan = 0;
        for (int i = 0; i < 6; ++i) any_nan |= isnan(z[i]) ? 1 : 0;
        if (any_nan) {
            *status = ST_NUMERIC_ERROR;
            break;
        }
        const float v0 = z[0] * *cur_x + z[2] * *cur_y + z[4];
        const float v1 = z[1] * *cur_x + z[3] * *cur_y + z[5];
        *cur_x += v0;
        *cur_y += v1;
        affine[0] += z[0];
        affine[2] += z[1];
        affine[1] += z[2];
        affine[3] += z[3];
        const float squared_step = v0 * v0 + v1 * v1;
        if (fast_step_check(t, squared_step, &last_squared_step, &large_step_cnt, status)) break;
    }
}

/* ============================================================================================================
 * LSSD KLT.  R = {r00, r01, r10, r11} row-major, t = {tx, ty}.
 * ========================================================================================================== */

/* Shared SE(2) update (OF/lssd_klt/optical_flow_lssd_klt.cpp:113-117, ..._fast.cpp:95-99):
 * R *= [1 -th; th 1];  R /= ||R.col(0)||;  t += v.tail<2>(). */
static void lssd_update(float *R, float *tr, const float *v) {
    const float th = v[0];
    const float n00 = R[0] * 1.0f + R[1] * th;
    const float n01 = R[0] * (-th) + R[1] * 1.0f;
    const float n10 = R[2] * 1.0f + R[3] * th;
    const float n11 = R[2] * (-th) + R[3] * 1.0f;
    const float norm = sqrtf(n00 * n00 + n10 * n10);
    R[0] = n00 / norm;
    R[1] = n01 / norm;
    R[2] = n10 / norm;
    R[3] = n11 / norm;
    tr[0] += v[1];
    tr[1] += v[2];
}

/* OF/lssd_klt/optical_flow_lssd_klt.cpp:127-250 ConstructIncrementalFunction: pass 1 = validity mask + patch means
 * (checked samples), pass 2 = Jacobians/residuals with unchecked samples.  H is 3x3 row-major. */
static int32_t lssd_construct(tracker_t *t, const image_t *ref, const image_t *cur, float ref_x, float ref_y, const float *R, const float *tr, float *H,
                              float *b) {
    float v[6];
    int32_t valid = 0;
    float ref_avg = 0.0f, cur_avg = 0.0f;
    const int inverse = t->o.method == M_INVERSE;
    int k = 0;
    for (int32_t drow = -t->o.patch_row_half; drow <= t->o.patch_row_half; ++drow) {
        for (int32_t dcol = -t->o.patch_col_half; dcol <= t->o.patch_col_half; ++dcol, ++k) {
            const float row_i = (float)drow + ref_y, col_i = (float)dcol + ref_x;
            const float col_j = (R[0] * col_i + R[1] * row_i) + tr[0];
            const float row_j = (R[2] * col_i + R[3] * row_i) + tr[1];
            const image_t *g = inverse ? ref : cur;
            const float gr = inverse ? row_i : row_j, gc = inverse ? col_i : col_j;
            if (px_checked(g, gr, gc - 1.0f, &v[0]) && px_checked(g, gr, gc + 1.0f, &v[1]) && px_checked(g, gr - 1.0f, gc, &v[2]) &&
                px_checked(g, gr + 1.0f, gc, &v[3]) && px_checked(ref, row_i, col_i, &v[4]) && px_checked(cur, row_j, col_j, &v[5])) {
                ref_avg += v[4];
                cur_avg += v[5];
                ++valid;
                t->pixel_valid[k] = 1;
            } else {
                t->pixel_valid[k] = 0;
            }
        }
    }
    ref_avg /= (float)valid;
    cur_avg /= (float)valid;

    k = 0;
    for (int32_t drow = -t->o.patch_row_half; drow <= t->o.patch_row_half; ++drow) {
        for (int32_t dcol = -t->o.patch_col_half; dcol <= t->o.patch_col_half; ++dcol, ++k) {
            if (!t->pixel_valid[k]) continue;
            const float row_i = (float)drow + ref_y, col_i = (float)dcol + ref_x;
            const float col_j = (R[0] * col_i + R[1] * row_i) + tr[0];
            const float row_j = (R[2] * col_i + R[3] * row_i) + tr[1];
            const image_t *g = inverse ? ref : cur;
            const float gr = inverse ? row_i : row_j, gc = inverse ? col_i : col_j;
            v[0] = px_f(g, gr, gc - 1.0f);
            v[1] = px_f(g, gr, gc + 1.0f);
            v[2] = px_f(g, gr - 1.0f, gc);
            v[3] = px_f(g, gr + 1.0f, gc);
            v[4] = px_f(ref, row_i, col_i);
            v[5] = px_f(cur, row_j, col_j);
            const float avg = inverse ? ref_avg : cur_avg;
            const float jp0 = (v[1] - v[0]) / avg, jp1 = (v[3] - v[2]) / avg;
            /* jacobian_se2 = [R*(-row_i, col_i) | I2] */
            const float s00 = R[0] * (-row_i) + R[1] * col_i;
            const float s10 = R[2] * (-row_i) + R[3] * col_i;
            float J[3];
            J[0] = jp0 * s00 + jp1 * s10;
            J[1] = jp0 * 1.0f + jp1 * 0.0f;
            J[2] = jp0 * 0.0f + jp1 * 1.0f;
            const float residual = v[5] / cur_avg - v[4] / ref_avg;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) H[i * 3 + j] += J[i] * J[j];
            for (int i = 0; i < 3; ++i) b[i] -= J[i] * residual;
        }
    }
    return valid;
}

/* OF/lssd_klt/optical_flow_lssd_klt.cpp:96-125 TrackOneFeature. */
static void lssd_track_one(tracker_t *t, const image_t *ref, const image_t *cur, float ref_x, float ref_y, float *R, float *tr, uint8_t *status) {
    for (uint32_t iter = 0; iter < t->o.max_iteration; ++iter) {
        float H[9] = {0}, b[3] = {0}, v[3];
        ++t->iterations;
        if (lssd_construct(t, ref, cur, ref_x, ref_y, R, tr, H, b) == 0) break;
        ldlt_solve(3, H, b, v);
        if (isnan(v[0]) || isnan(v[1]) || isnan(v[2])) {
            *status = ST_NUMERIC_ERROR;
            break;
        }
        lssd_update(R, tr, v);
        if (v[0] * v[0] + v[1] * v[1] + v[2] * v[2] < t->o.max_converge_step) {
            *status = ST_TRACKED;
            break;
        }
    }
}

/* OF/lssd_klt/optical_flow_lssd_klt_fast.cpp:145-195 ExtractPatchInCurrentImage.  The "inside" test truncates
 * (static_cast<int32_t>) and uses a +-patch_rows/cols margin; inside it samples unchecked. */
static uint32_t lssd_extract_cur_patch(tracker_t *t, const image_t *cur, float ref_x, float ref_y, const float *R, const float *tr) {
    const float cx = (R[0] * ref_x + R[1] * ref_y) + tr[0];
    const float cy = (R[2] * ref_x + R[3] * ref_y) + tr[1];
    const int32_t min_row = (int32_t)cy - t->patch_rows;
    const int32_t min_col = (int32_t)cx - t->patch_cols;
    const int32_t max_row = min_row + t->patch_rows * 2;
    const int32_t max_col = min_col + t->patch_cols * 2;
    const int partly_outside = min_row < 0 || max_row > cur->rows - 2 || min_col < 0 || max_col > cur->cols - 2;
    uint32_t valid = 0;
    int k = 0;
    for (int32_t drow = -t->o.patch_row_half; drow <= t->o.patch_row_half; ++drow) {
        for (int32_t dcol = -t->o.patch_col_half; dcol <= t->o.patch_col_half; ++dcol, ++k) {
            const float row_i = (float)drow + ref_y, col_i = (float)dcol + ref_x;
            const float col_j = (R[0] * col_i + R[1] * row_i) + tr[0];
            const float row_j = (R[2] * col_i + R[3] * row_i) + tr[1];
            if (partly_outside) {
                float value;
                if (px_checked(cur, row_j, col_j, &value)) {
                    t->cur_patch[k] = value;
                    t->cur_valid[k] = 1;
                    ++valid;
                } else {
                    t->cur_patch[k] = 0.0f;
                    t->cur_valid[k] = 0;
                }
            } else {
                t->cur_patch[k] = px_f(cur, row_j, col_j);
                t->cur_valid[k] = 1;
                ++valid;
            }
        }
    }
    return valid;
}

/* OF/lssd_klt/optical_flow_lssd_klt_fast.cpp:197-229 ComputeHessianAndBias. */
static int32_t lssd_fast_hessian_bias(const tracker_t *t, float ref_x, float ref_y, const float *R, float *H, float *b) {
    int32_t valid = 0;
    int k = 0;
    for (int32_t drow = -t->o.patch_row_half; drow <= t->o.patch_row_half; ++drow) {
        for (int32_t dcol = -t->o.patch_col_half; dcol <= t->o.patch_col_half; ++dcol, ++k) {
            const float row_i = (float)drow + ref_y, col_i = (float)dcol + ref_x;
            const int32_t e = (drow + t->o.patch_row_half + 1) * t->ex_cols + dcol + t->o.patch_col_half + 1;
            if (!(t->ex_valid[e] && t->cur_valid[k])) continue;
            const float s0 = R[0] * (-row_i) + R[1] * col_i;
            const float s1 = R[2] * (-row_i) + R[3] * col_i;
            float J[3];
            J[0] = t->dx[k] * s0 + t->dy[k] * s1;
            J[1] = t->dx[k];
            J[2] = t->dy[k];
            const float residual = t->cur_patch[k] - t->ex_patch[e];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) H[i * 3 + j] += J[i] * J[j];
            for (int i = 0; i < 3; ++i) b[i] -= J[i] * residual;
            ++valid;
        }
    }
    return valid;
}

/* OF/lssd_klt/optical_flow_lssd_klt_fast.cpp:7-114 TrackOneFeatureFast. */
static void lssd_track_one_fast(tracker_t *t, const image_t *ref, const image_t *cur, float ref_x, float ref_y, float *R, float *tr, uint8_t *status) {
    const uint32_t valid_ref = extract_ex_ref_patch(t, ref, ref_x, ref_y);
    if (valid_ref == 0) {
        *status = ST_OUTSIDE;
        return;
    }
    /* :116-143 PrecomputeJacobian */
    for (int32_t row = 0; row < t->patch_rows; ++row)
        for (int32_t col = 0; col < t->patch_cols; ++col) ex_gradient(t, row, col, &t->dx[row * t->patch_cols + col], &t->dy[row * t->patch_cols + col]);

    if (t->o.consider_patch_luminance) {
        /* :27-46: mean over the INTERIOR of the extended patch divided by the WHOLE extended patch's valid count. */
        float ref_avg = 0.0f;
        for (int32_t row = 1; row < t->ex_rows - 1; ++row)
            for (int32_t col = 1; col < t->ex_cols - 1; ++col) ref_avg += t->ex_patch[row * t->ex_cols + col];
        ref_avg /= (float)valid_ref;
        for (int32_t k = 0; k < t->patch_size; ++k) t->dx[k] /= ref_avg;
        for (int32_t k = 0; k < t->patch_size; ++k) t->dy[k] /= ref_avg;
        for (int32_t k = 0; k < t->ex_size; ++k) t->ex_patch[k] /= ref_avg;
    }

    *status = ST_LARGE_RESIDUAL;
    float last_squared_step = INFINITY;
    uint32_t large_step_cnt = 0;
    for (uint32_t iter = 0; iter < t->o.max_iteration; ++iter) {
        ++t->iterations;
        const uint32_t valid_cur = lssd_extract_cur_patch(t, cur, ref_x, ref_y, R, tr);
        if (valid_cur == 0) break;
        if (t->o.consider_patch_luminance) {
            /* :65-78: mean over the interior of the cur patch divided by the whole patch's valid count. */
            float cur_avg = 0.0f;
            for (int32_t row = 1; row < t->patch_rows - 1; ++row)
                for (int32_t col = 1; col < t->patch_cols - 1; ++col) cur_avg += t->cur_patch[row * t->patch_cols + col];
            cur_avg /= (float)valid_cur;
            for (int32_t k = 0; k < t->patch_size; ++k) t->cur_patch[k] /= cur_avg;
        }
        float H[9] = {0}, b[3] = {0}, v[3];
        if (lssd_fast_hessian_bias(t, ref_x, ref_y, R, H, b) == 0) break;
        ldlt_solve(3, H, b, v);
        if (isnan(v[0]) || isnan(v[1]) || isnan(v[2])) {
            *status = ST_NUMERIC_ERROR;
            break;
        }
        lssd_update(R, tr, v);
        const float squared_step = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
        if (fast_step_check(t, squared_step, &last_squared_step, &large_step_cnt, status)) break;
    }
}

/* ============================================================================================================
 * Per-variant level drivers (TrackMultipleLevel / TrackSingleLevel) and the public entry (TrackFeatures).
 * ========================================================================================================== */

static void track_feature_multi(tracker_t *t, int32_t levels, const image_t *ref_pyr, const image_t *cur_pyr, const float *ref_uv, float *cur_uv,
                                uint8_t *status) {
    const float scale = (float)(1 << (levels - 1));
    const int fast = !(t->o.method == M_INVERSE || t->o.method == M_DIRECT);
    float sref_x = ref_uv[0] / scale, sref_y = ref_uv[1] / scale;
    float scur_x = cur_uv[0] / scale, scur_y = cur_uv[1] / scale;
    if (t->o.variant == 0) {
        /* OF/basic_klt/optical_flow_basic_klt.cpp:7-57 */
        for (int32_t l = levels - 1; l > -1; --l) {
            if (!fast) basic_track_one(t, &ref_pyr[l], &cur_pyr[l], sref_x, sref_y, &scur_x, &scur_y, status);
            else basic_track_one_fast(t, &ref_pyr[l], &cur_pyr[l], sref_x, sref_y, &scur_x, &scur_y, status);
            if (!l) {
                cur_uv[0] = scur_x;
                cur_uv[1] = scur_y;
                break;
            }
            sref_x *= 2.0f, sref_y *= 2.0f, scur_x *= 2.0f, scur_y *= 2.0f;
        }
    } else if (t->o.variant == 1) {
        /* OF/affine_klt/optical_flow_affine_klt.cpp:6-59: affine starts at identity, carried across levels un-scaled */
        float affine[4] = {1.0f, 0.0f, 0.0f, 1.0f};
        for (int32_t l = levels - 1; l > -1; --l) {
            if (!fast) affine_track_one(t, &ref_pyr[l], &cur_pyr[l], sref_x, sref_y, &scur_x, &scur_y, affine, status);
            else affine_track_one_fast(t, &ref_pyr[l], &cur_pyr[l], sref_x, sref_y, &scur_x, &scur_y, affine, status);
            if (!l) {
                cur_uv[0] = scur_x;
                cur_uv[1] = scur_y;
                break;
            }
            sref_x *= 2.0f, sref_y *= 2.0f, scur_x *= 2.0f, scur_y *= 2.0f;
        }
    } else {
        /* OF/lssd_klt/optical_flow_lssd_klt.cpp:7-61 */
        const float *P = t->o.predict;
        float R[4] = {P[0], P[1], P[2], P[3]};
        float tr[2] = {scur_x - (P[0] * sref_x + P[1] * sref_y), scur_y - (P[2] * sref_x + P[3] * sref_y)};
        for (int32_t l = levels - 1; l > -1; --l) {
            if (!fast) lssd_track_one(t, &ref_pyr[l], &cur_pyr[l], sref_x, sref_y, R, tr, status);
            else lssd_track_one_fast(t, &ref_pyr[l], &cur_pyr[l], sref_x, sref_y, R, tr, status);
            if (!l) {
                cur_uv[0] = (R[0] * ref_uv[0] + R[1] * ref_uv[1]) + tr[0];
                cur_uv[1] = (R[2] * ref_uv[0] + R[3] * ref_uv[1]) + tr[1];
                break;
            }
            sref_x *= 2.0f, sref_y *= 2.0f;
            tr[0] *= 2.0f, tr[1] *= 2.0f;
        }
    }
    if (is_outside(&cur_pyr[0], cur_uv[0], cur_uv[1])) *status = ST_OUTSIDE;
}

static void track_feature_single(tracker_t *t, const image_t *ref, const image_t *cur, const float *ref_uv, float *cur_uv, uint8_t *status) {
    const int fast = !(t->o.method == M_INVERSE || t->o.method == M_DIRECT);
    if (t->o.variant == 0) {
        /* OF/basic_klt/optical_flow_basic_klt.cpp:59-86 */
        if (!fast) basic_track_one(t, ref, cur, ref_uv[0], ref_uv[1], &cur_uv[0], &cur_uv[1], status);
        else basic_track_one_fast(t, ref, cur, ref_uv[0], ref_uv[1], &cur_uv[0], &cur_uv[1], status);
    } else if (t->o.variant == 1) {
        /* OF/affine_klt/optical_flow_affine_klt.cpp:61-91: starts from predict_affine_ */
        float affine[4] = {t->o.predict[0], t->o.predict[1], t->o.predict[2], t->o.predict[3]};
        if (!fast) affine_track_one(t, ref, cur, ref_uv[0], ref_uv[1], &cur_uv[0], &cur_uv[1], affine, status);
        else affine_track_one_fast(t, ref, cur, ref_uv[0], ref_uv[1], &cur_uv[0], &cur_uv[1], affine, status);
    } else {
        /* OF/lssd_klt/optical_flow_lssd_klt.cpp:63-94: the result is never written back to cur_pixel_uv (quirk) */
        const float *P = t->o.predict;
        float R[4] = {P[0], P[1], P[2], P[3]};
        float tr[2] = {cur_uv[0] - (P[0] * ref_uv[0] + P[1] * ref_uv[1]), cur_uv[1] - (P[2] * ref_uv[0] + P[3] * ref_uv[1])};
        if (!fast) lssd_track_one(t, ref, cur, ref_uv[0], ref_uv[1], R, tr, status);
        else lssd_track_one_fast(t, ref, cur, ref_uv[0], ref_uv[1], R, tr, status);
    }
    if (is_outside(cur, cur_uv[0], cur_uv[1])) *status = ST_OUTSIDE;
}

/* OF/optical_flow.cpp:6-47 TrackFeatures (both overloads). */
int ftko_klt_track_traced(const ftko_klt_params *params, int32_t levels, const uint8_t *const *ref_levels, const uint8_t *const *cur_levels,
                          const int32_t *rows, const int32_t *cols, int32_t n, const float *ref_uv, float *cur_uv, int32_t cur_uv_count,
                          uint8_t *status, int32_t status_count, int32_t single_level, int32_t *iterations) {
    if (n <= 0) return 0;                   /* :8, :30 ref_pixel_uv.empty() */
    if (levels < 1 || levels > 16) return 0; /* pyramids of equal depth are implied by the shared `levels` (:9) */
    if (params->variant < 0 || params->variant > 2) return 0;
    if (cur_uv_count != n) memcpy(cur_uv, ref_uv, sizeof(float) * 2 * (size_t)n); /* :12-14 */
    if (status_count != n) memset(status, ST_NOT_TRACKED, (size_t)n);              /* :17-19 */
    if (iterations) memset(iterations, 0, sizeof(int32_t) * (size_t)n);

    tracker_t t;
    if (!tracker_init(&t, params)) return 0;

    /* private padded copies: the sampler reads the +1 neighbour with weight 0 on the last row/col */
    image_t ref_pyr[16], cur_pyr[16];
    uint8_t *store[32];
    for (int32_t l = 0; l < levels; ++l) {
        const size_t npx = (size_t)rows[l] * cols[l];
        store[2 * l] = (uint8_t *)calloc(npx + cols[l] + 2, 1);
        store[2 * l + 1] = (uint8_t *)calloc(npx + cols[l] + 2, 1);
        memcpy(store[2 * l], ref_levels[l], npx);
        memcpy(store[2 * l + 1], cur_levels[l], npx);
        ref_pyr[l].d = store[2 * l], ref_pyr[l].rows = rows[l], ref_pyr[l].cols = cols[l];
        cur_pyr[l].d = store[2 * l + 1], cur_pyr[l].rows = rows[l], cur_pyr[l].cols = cols[l];
    }

    const uint32_t max_id = (uint32_t)n < params->max_track_points ? (uint32_t)n : params->max_track_points;
    for (uint32_t i = 0; i < max_id; ++i) {
        if (status[i] > ST_TRACKED) continue; /* never re-track failed features */
        t.iterations = 0;
        if (single_level) track_feature_single(&t, &ref_pyr[0], &cur_pyr[0], &ref_uv[2 * i], &cur_uv[2 * i], &status[i]);
        else track_feature_multi(&t, levels, ref_pyr, cur_pyr, &ref_uv[2 * i], &cur_uv[2 * i], &status[i]);
        if (iterations) iterations[i] = t.iterations;
    }

    for (int32_t l = 0; l < 2 * levels; ++l) free(store[l]);
    tracker_free(&t);
    return 1;
}

int ftko_klt_track(const ftko_klt_params *params, int32_t levels, const uint8_t *const *ref_levels, const uint8_t *const *cur_levels,
                   const int32_t *rows, const int32_t *cols, int32_t n, const float *ref_uv, float *cur_uv, int32_t cur_uv_count, uint8_t *status,
                   int32_t status_count, int32_t single_level) {
    return ftko_klt_track_traced(params, levels, ref_levels, cur_levels, rows, cols, n, ref_uv, cur_uv, cur_uv_count, status, status_count, single_level,
                                 NULL);
}

/* The reference demo's timed region: CreateImagePyramid x2 + TrackFeatures (test/test_optical_flow.cpp:69-73). */
int ftko_pyramid_and_track(const ftko_klt_params *params, int32_t levels, const uint8_t *ref_image, const uint8_t *cur_image, int32_t rows,
                           int32_t cols, int32_t n, const float *ref_uv, float *cur_uv, uint8_t *status) {
    if (levels < 1 || levels > 16) return 0;
    const size_t npx = (size_t)rows * cols;
    uint8_t *ref_buf = (uint8_t *)malloc(npx + 1), *cur_buf = (uint8_t *)malloc(npx + 1);
    ftko_pyramid_build(ref_image, rows, cols, levels, ref_buf);
    ftko_pyramid_build(cur_image, rows, cols, levels, cur_buf);
    const uint8_t *rl[16], *cl[16];
    int32_t lr[16], lc[16];
    rl[0] = ref_image, cl[0] = cur_image, lr[0] = rows, lc[0] = cols;
    size_t off = 0;
    for (int32_t l = 1; l < levels; ++l) {
        lr[l] = lr[l - 1] >> 1, lc[l] = lc[l - 1] >> 1;
        rl[l] = ref_buf + off, cl[l] = cur_buf + off;
        off += (size_t)lr[l] * lc[l];
    }
    const int ok = ftko_klt_track(params, levels, rl, cl, lr, lc, n, ref_uv, cur_uv, 0, status, 0, 0);
    free(ref_buf);
    free(cur_buf);
    return ok;
}

/* ============================================================================================================
 * DESCRIPTOR MATCHING  (descriptor_matcher/descriptor_matcher.h)
 * ========================================================================================================== */

/* test/test_descriptor_matcher_brief.cpp:33-45: count of differing elements, as float. */
static float brief_distance(const uint8_t *a, const uint8_t *b, int32_t len) {
    if (len == 0) return (float)2147483647;
    int32_t d = 0;
    for (int32_t k = 0; k < len; ++k) d += (a[k] != b[k]) ? 1 : 0;
    return (float)d;
}

/* test/test_descriptor_matcher_superpoint.cpp:32-34 / test_descriptor_matcher_disk.cpp:32-34:
 * 0.5 - dot/|a|/|b|*0.5, dot and norms summed sequentially k = 0..dim-1. */
static float seq_dot(const float *a, const float *b, int32_t n) {
    float s = a[0] * b[0];
    for (int32_t k = 1; k < n; ++k) s = s + a[k] * b[k];
    return s;
}
static float cosine_distance(const float *a, const float *b, int32_t dim) {
    return 0.5f - seq_dot(a, b, dim) / sqrtf(seq_dot(a, a, dim)) / sqrtf(seq_dot(b, b, dim)) * 0.5f;
}

typedef struct {
    int kind; /* 0 brief, 1 cosine */
    const void *ref, *cur;
    int32_t len;
} desc_set_t;

static float pair_distance(const desc_set_t *s, int32_t i, int32_t j) {
    if (s->kind == 0) return brief_distance((const uint8_t *)s->ref + (size_t)i * s->len, (const uint8_t *)s->cur + (size_t)j * s->len, s->len);
    return cosine_distance((const float *)s->ref + (size_t)i * s->len, (const float *)s->cur + (size_t)j * s->len, s->len);
}

/* descriptor_matcher.h:55-79 ForceMatch (index version). */
static int force_match(const desc_set_t *s, int32_t n_ref, int32_t n_cur, float max_dist, int32_t *idx, int32_t idx_count) {
    if (n_cur <= 0) return 0;
    if (idx_count != n_ref)
        for (int32_t i = 0; i < n_ref; ++i) idx[i] = -1;
    for (int32_t i = 0; i < n_ref; ++i) {
        float min_distance = max_dist;
        for (int32_t j = 0; j < n_cur; ++j) {
            const float d = pair_distance(s, i, j);
            if (d < min_distance && d < max_dist) {
                min_distance = d;
                idx[i] = j;
            }
        }
    }
    return 1;
}

/* descriptor_matcher.h:90-124 NearbyMatch (index version): window gate, strict <, break on d == 0. */
static int nearby_match(const desc_set_t *s, int32_t n_ref, int32_t n_cur, const float *pred_uv, const float *cur_uv, int32_t max_drow, int32_t max_dcol,
                        float max_dist, int32_t *idx, int32_t idx_count) {
    if (n_cur <= 0) return 0;
    if (idx_count != n_ref)
        for (int32_t i = 0; i < n_ref; ++i) idx[i] = -1;
    for (int32_t i = 0; i < n_ref; ++i) {
        float min_distance = max_dist;
        for (int32_t j = 0; j < n_cur; ++j) {
            if (fabsf(pred_uv[2 * i] - cur_uv[2 * j]) > (float)max_dcol || fabsf(pred_uv[2 * i + 1] - cur_uv[2 * j + 1]) > (float)max_drow) continue;
            const float d = pair_distance(s, i, j);
            if (d < min_distance && d < max_dist) {
                min_distance = d;
                idx[i] = j;
            }
            if (d == 0) break;
        }
    }
    return 1;
}

/* descriptor_matcher.h:135-157 FillMatchedPixelByPairIndices. */
static void fill_matched(const int32_t *idx, int32_t n_ref, const float *cur_uv, int32_t n_cur, float *matched_uv, uint8_t *status, int32_t status_count) {
    if (status_count != n_ref) memset(status, ST_NOT_TRACKED, (size_t)n_ref);
    for (int32_t i = 0; i < n_ref; ++i) {
        if (status[i] > ST_TRACKED) continue;
        const int32_t j = idx[i];
        if (j >= 0 && j < n_cur) {
            matched_uv[2 * i] = cur_uv[2 * j];
            matched_uv[2 * i + 1] = cur_uv[2 * j + 1];
            status[i] = ST_TRACKED;
        } else {
            status[i] = ST_LARGE_RESIDUAL;
        }
    }
}

int ftko_match_brief_force(const uint8_t *ref_bits, int32_t n_ref, const uint8_t *cur_bits, int32_t n_cur, int32_t len, float max_dist, int32_t *idx,
                           int32_t idx_count) {
    const desc_set_t s = {0, ref_bits, cur_bits, len};
    return force_match(&s, n_ref, n_cur, max_dist, idx, idx_count);
}

int ftko_match_brief_nearby(const uint8_t *ref_bits, int32_t n_ref, const uint8_t *cur_bits, int32_t n_cur, int32_t len, const float *pred_uv,
                            const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist, int32_t *idx, int32_t idx_count) {
    const desc_set_t s = {0, ref_bits, cur_bits, len};
    return nearby_match(&s, n_ref, n_cur, pred_uv, cur_uv, max_drow, max_dcol, max_dist, idx, idx_count);
}

int ftko_match_cosine_force(const float *ref, int32_t n_ref, const float *cur, int32_t n_cur, int32_t dim, float max_dist, int32_t *idx,
                            int32_t idx_count) {
    const desc_set_t s = {1, ref, cur, dim};
    return force_match(&s, n_ref, n_cur, max_dist, idx, idx_count);
}

int ftko_match_cosine_nearby(const float *ref, int32_t n_ref, const float *cur, int32_t n_cur, int32_t dim, const float *pred_uv, const float *cur_uv,
                             int32_t max_drow, int32_t max_dcol, float max_dist, int32_t *idx, int32_t idx_count) {
    const desc_set_t s = {1, ref, cur, dim};
    return nearby_match(&s, n_ref, n_cur, pred_uv, cur_uv, max_drow, max_dcol, max_dist, idx, idx_count);
}

/* descriptor_matcher.h:126-133 NearbyMatch (uv + status version): a fresh index vector, then the fill. */
int ftko_match_brief_nearby_uv(const uint8_t *ref_bits, int32_t n_ref, const uint8_t *cur_bits, int32_t n_cur, int32_t len, const float *pred_uv,
                               const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist, float *matched_uv, uint8_t *status,
                               int32_t status_count) {
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_ref > 0 ? n_ref : 1));
    const desc_set_t s = {0, ref_bits, cur_bits, len};
    const int ok = nearby_match(&s, n_ref, n_cur, pred_uv, cur_uv, max_drow, max_dcol, max_dist, idx, -1);
    if (ok) fill_matched(idx, n_ref, cur_uv, n_cur, matched_uv, status, status_count);
    free(idx);
    return ok;
}

/* descriptor_matcher.h:81-88 ForceMatch (uv + status version). */
int ftko_match_brief_force_uv(const uint8_t *ref_bits, int32_t n_ref, const uint8_t *cur_bits, int32_t n_cur, int32_t len, const float *cur_uv,
                              float max_dist, float *matched_uv, uint8_t *status, int32_t status_count) {
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_ref > 0 ? n_ref : 1));
    const desc_set_t s = {0, ref_bits, cur_bits, len};
    const int ok = force_match(&s, n_ref, n_cur, max_dist, idx, -1);
    if (ok) fill_matched(idx, n_ref, cur_uv, n_cur, matched_uv, status, status_count);
    free(idx);
    return ok;
}

/* src/nn_feature_matcher/nn_feature_matcher.cpp:180-216, the score-matrix branch of NNFeatureMatcher::Match after the network
 * has run (the network itself needs ONNX Runtime and is outside the path; so is a compiled-in-place check of these lines --
 * this restatement is unpinned).  scores is the n_ref x n_cur row-major matrix; idx[i] = matched column or -1
 * (status kTracked / kLargeResidual in the reference). */
int ftko_mutual_scores(const float *scores, int32_t n_ref, int32_t n_cur, float min_score, int32_t *idx) {
    if (n_cur <= 0) return 0;
    int32_t *max_scores_in_cols_index = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_cur);
    for (int32_t j = 0; j < n_cur; ++j) { /* :187-198 */
        int32_t max_score_index = 0;
        float max_score = n_ref > 0 ? scores[j] : 0.0f;
        for (int32_t i = 1; i < n_ref; ++i) {
            if (scores[(size_t)i * n_cur + j] > max_score) {
                max_score = scores[(size_t)i * n_cur + j];
                max_score_index = i;
            }
        }
        max_scores_in_cols_index[j] = max_score_index;
    }
    for (int32_t idx_ref = 0; idx_ref < n_ref; ++idx_ref) { /* :200-214 */
        const float *row = scores + (size_t)idx_ref * n_cur;
        int32_t max_score_index = 0;
        float max_score = row[0];
        for (int32_t j = 1; j < n_cur; ++j) {
            if (row[j] > max_score) {
                max_score = row[j];
                max_score_index = j;
            }
        }
        idx[idx_ref] = -1;
        if (max_score < min_score) continue;
        if (max_scores_in_cols_index[max_score_index] != idx_ref) continue;
        idx[idx_ref] = max_score_index;
    }
    free(max_scores_in_cols_index);
    return 1;
}

/* ------------------------------------------------------------------------------------------------------------
 * Direct-method pose tracker: src/direct_method_tracker/direct_method_tracker.cpp.
 * Quaternion / camera arithmetic is external (Eigen::Quaternionf, Sensor_Model CameraPinhole); its semantics are frozen in
 * oracle/shim/basic_type.h (Quat) and oracle/shim/camera_pinhole.h and restated here.  q = (w, x, y, z).
 * ---------------------------------------------------------------------------------------------------------- */
static void quat_rotate(const float *q, const float *v, float *out) { /* Quat::operator*(Vec3) */
    const float qw = q[0], qx = q[1], qy = q[2], qz = q[3];
    float ux = qy * v[2] - qz * v[1], uy = qz * v[0] - qx * v[2], uz = qx * v[1] - qy * v[0];
    ux = ux + ux, uy = uy + uy, uz = uz + uz;
    out[0] = v[0] + qw * ux + (qy * uz - qz * uy);
    out[1] = v[1] + qw * uy + (qz * ux - qx * uz);
    out[2] = v[2] + qw * uz + (qx * uy - qy * ux);
}
static void quat_inverse(const float *q, float *out) {
    const float n2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0];
    if (n2 > 0.0f) out[0] = q[0] / n2, out[1] = -q[1] / n2, out[2] = -q[2] / n2, out[3] = -q[3] / n2;
    else out[0] = out[1] = out[2] = out[3] = 0.0f;
}
static void quat_normalize(float *q) {
    const float n = sqrtf(q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0]);
    q[0] = q[0] / n, q[1] = q[1] / n, q[2] = q[2] / n, q[3] = q[3] / n;
}
static void quat_mul(const float *a, const float *b, float *out) {
    const float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    const float x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    const float y = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    const float z = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
    out[0] = w, out[1] = x, out[2] = y, out[3] = z;
}

#define FTKO_ZERO_FLOAT 1e-6f /* kZeroFloat, oracle/shim/basic_type.h */

/* direct_method_tracker.cpp:115-192 TrackAllFeaturesDirect on one level. */
static void direct_track_level(const ftko_direct_params *o, const image_t *ref, const image_t *cur, const float *K, int32_t n, const float *p_c_in_ref,
                               const float *ref_uv, float *cur_uv, float *q_rc, float *p_rc) {
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    for (uint32_t iter = 0; iter < o->max_iteration; ++iter) {
        float H[36], b[6];
        for (int k = 0; k < 36; ++k) H[k] = 0.0f;
        for (int k = 0; k < 6; ++k) b[k] = 0.0f;
        const uint32_t max_id = (uint32_t)n < o->max_track_points ? (uint32_t)n : o->max_track_points;
        float q_inv[4];
        quat_inverse(q_rc, q_inv);
        for (uint32_t i = 0; i < max_id; ++i) {
            const float *pr = &p_c_in_ref[3 * i];
            if (pr[2] < FTKO_ZERO_FLOAT) continue; /* :129 */
            const float x = pr[0], y = pr[1], z = pr[2];
            const float z_inv = 1.0f / z;
            const float z2_inv = z_inv * z_inv;
            const float d[3] = {pr[0] - p_rc[0], pr[1] - p_rc[1], pr[2] - p_rc[2]};
            float pc[3];
            quat_rotate(q_inv, d, pc); /* :139 */
            if (pc[2] < FTKO_ZERO_FLOAT) continue; /* :140 */
            const float nx = pc[0] / pc[2], ny = pc[1] / pc[2]; /* :142 */
            cur_uv[2 * i] = fx * nx + cx;                       /* :143 */
            cur_uv[2 * i + 1] = fy * ny + cy;
            /* :146-150, row-major 2x6; every product chain runs left to right */
            const float J[12] = {fx * z_inv, 0.0f, -fx * x * z2_inv, -fx * x * y * z2_inv, fx + fx * x * x * z2_inv, -fx * y * z_inv,
                                 0.0f, fy * z_inv, -fy * y * z2_inv, -fy - fy * y * y * z2_inv, fy * x * y * z2_inv, fy * x * z_inv};
            for (int32_t drow = -o->patch_row_half; drow <= o->patch_row_half; ++drow) {
                for (int32_t dcol = -o->patch_col_half; dcol <= o->patch_col_half; ++dcol) {
                    const float row_i = (float)drow + ref_uv[2 * i + 1], col_i = (float)dcol + ref_uv[2 * i];
                    const float row_j = (float)drow + cur_uv[2 * i + 1], col_j = (float)dcol + cur_uv[2 * i];
                    float t[6];
                    if (px_checked(cur, row_j, col_j - 1.0f, &t[0]) && px_checked(cur, row_j, col_j + 1.0f, &t[1]) &&
                        px_checked(cur, row_j - 1.0f, col_j, &t[2]) && px_checked(cur, row_j + 1.0f, col_j, &t[3]) &&
                        px_checked(ref, row_i, col_i, &t[4]) && px_checked(cur, row_j, col_j, &t[5])) {
                        const float gx = (t[1] - t[0]) * 0.5f, gy = (t[3] - t[2]) * 0.5f; /* :165 */
                        const float residual = t[5] - t[4];
                        float jac[6];
                        for (int k = 0; k < 6; ++k) jac[k] = gx * J[k] + gy * J[6 + k]; /* :169 */
                        for (int r = 0; r < 6; ++r)
                            for (int c = 0; c < 6; ++c) H[6 * r + c] = H[6 * r + c] + jac[r] * jac[c]; /* :170 */
                        for (int k = 0; k < 6; ++k) b[k] = b[k] + residual * jac[k];                  /* :171 */
                    }
                }
            }
        }
        float dx[6];
        ldlt_solve(6, H, b, dx); /* :178 */
        int any_nan = 0;
        for (int k = 0; k < 6; ++k) any_nan = any_nan || isnan(dx[k]);
        if (any_nan) break;
        p_rc[0] = p_rc[0] + dx[0], p_rc[1] = p_rc[1] + dx[1], p_rc[2] = p_rc[2] + dx[2]; /* :182 */
        float dq[4] = {1.0f, dx[3] * 0.5f, dx[4] * 0.5f, dx[5] * 0.5f}, qn[4];
        quat_normalize(dq);
        quat_mul(dq, q_rc, qn); /* :183 */
        q_rc[0] = qn[0], q_rc[1] = qn[1], q_rc[2] = qn[2], q_rc[3] = qn[3];
        quat_normalize(q_rc); /* :184 */
        float sq = dx[0] * dx[0];
        for (int k = 1; k < 6; ++k) sq = sq + dx[k] * dx[k];
        if (sq < o->max_converge_step) break; /* :187 */
    }
}

/* direct_method_tracker.cpp:41-95 TrackFeatures (camera-frame overload). */
int ftko_direct_method_track(const ftko_direct_params *params, int32_t levels, const uint8_t *const *ref_levels, const uint8_t *const *cur_levels,
                             const int32_t *rows, const int32_t *cols, const float *K, int32_t n, const float *p_c_in_ref, const float *ref_uv,
                             float *cur_uv, int32_t cur_uv_count, float *q_rc, float *p_rc, uint8_t *status, int32_t status_count) {
    if (n <= 0) return 0;                    /* :44 */
    if (levels < 1 || levels > 16) return 0; /* :45 equal depth is implied by the shared `levels` */
    if (cur_uv_count != n) memcpy(cur_uv, ref_uv, sizeof(float) * 2 * (size_t)n); /* :48-50 */
    image_t ref_pyr[16], cur_pyr[16];
    uint8_t *store[32];
    for (int32_t l = 0; l < levels; ++l) {
        const size_t npx = (size_t)rows[l] * cols[l];
        store[2 * l] = (uint8_t *)calloc(npx + cols[l] + 2, 1);
        store[2 * l + 1] = (uint8_t *)calloc(npx + cols[l] + 2, 1);
        memcpy(store[2 * l], ref_levels[l], npx);
        memcpy(store[2 * l + 1], cur_levels[l], npx);
        ref_pyr[l].d = store[2 * l], ref_pyr[l].rows = rows[l], ref_pyr[l].cols = cols[l];
        cur_pyr[l].d = store[2 * l + 1], cur_pyr[l].rows = rows[l], cur_pyr[l].cols = cols[l];
    }
    const float scale = (float)(1 << (levels - 1));
    float *scaled_ref = (float *)malloc(sizeof(float) * 2 * (size_t)n);
    for (int32_t i = 0; i < 2 * n; ++i) scaled_ref[i] = ref_uv[i] / scale; /* :56-58 */
    float sK[4] = {K[0] / scale, K[1] / scale, K[2] / scale, K[3] / scale};
    for (int32_t l = levels - 1; l > -1; --l) {
        if (params->method == 1) direct_track_level(params, &ref_pyr[l], &cur_pyr[l], sK, n, p_c_in_ref, scaled_ref, cur_uv, q_rc, p_rc);
        /* kInverse (:107-113) and kFast (:194-199) are empty upstream: they return true without touching anything */
        if (l == 0) break;
        for (int32_t i = 0; i < 2 * n; ++i) scaled_ref[i] = scaled_ref[i] * 2.0f;
        for (int k = 0; k < 4; ++k) sK[k] = sK[k] * 2.0f;
    }
    if (status_count != n) memset(status, ST_TRACKED, (size_t)n); /* :82-84 */
    for (int32_t i = 0; i < n; ++i) {                             /* :86-91 (bounds of the REFERENCE pyramid's level 0) */
        if (cur_uv[2 * i] < 0 || cur_uv[2 * i] > (float)(cols[0] - 1) || cur_uv[2 * i + 1] < 0 || cur_uv[2 * i + 1] > (float)(rows[0] - 1))
            status[i] = ST_OUTSIDE;
    }
    free(scaled_ref);
    for (int32_t l = 0; l < 2 * levels; ++l) free(store[l]);
    return 1;
}

/* ------------------------------------------------------------------------------------------------------------
 * Dense optical flow, Gunnar Farneback: src/dense_optical_flow_tracker/dense_optical_flow.cpp.
 * External: slam_utility::Utility::Interpolate (frozen in oracle/shim/slam_basic_math.h), Eigen 2x2 inverse (basic_type.h).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    float w[32 * 32]; /* normalised Gaussian kernel, (2h+1)^2 entries */
    float k2, k4, k22;
    int32_t half, size;
} dof_kernel_t;

static void dof_init_kernel(dof_kernel_t *g, int32_t half) { /* :87-137 */
    g->half = half;
    g->size = 2 * half + 1;
    g->k2 = g->k4 = g->k22 = 0.0f;
    for (int32_t i = 0; i < g->size * g->size; ++i) g->w[i] = 0.0f;
    if (half == 0) {
        g->w[0] = 1.0f;
        return; /* the moments stay 0 */
    }
    const float sigma = 1.0f;
    const float sigma2 = sigma * sigma;
    float sum = 0.0f;
    for (int32_t row = 0; row < g->size; ++row)
        for (int32_t col = 0; col < g->size; ++col) {
            const int32_t dr = row - half, dc = col - half;
            g->w[row * g->size + col] = expf(-0.5f * (float)(dr * dr + dc * dc) / sigma2);
            sum += g->w[row * g->size + col];
        }
    for (int32_t i = 0; i < g->size * g->size; ++i) g->w[i] = g->w[i] / sum;
    for (int32_t row = 0; row < g->size; ++row)
        for (int32_t col = 0; col < g->size; ++col) {
            const int32_t dr = row - half, dc = col - half;
            const float w = g->w[row * g->size + col];
            g->k2 += w * (float)dr * (float)dr;
            g->k4 += w * (float)dr * (float)dr * (float)dr * (float)dr;
            g->k22 += w * (float)dr * (float)dr * (float)dc * (float)dc;
        }
}

/* :139-195: six Gaussian-weighted moment maps, planes S0, Srow, Scol, Srowcol, Srowrow, Scolcol of rows*cols floats */
static void dof_moments(const dof_kernel_t *g, const image_t *im, float *S) {
    const int32_t rows = im->rows, cols = im->cols, h = g->half;
    const size_t n = (size_t)rows * cols;
    for (int32_t row = 0; row < rows; ++row)
        for (int32_t col = 0; col < cols; ++col) {
            float s0 = 0.0f, sr = 0.0f, sc = 0.0f, src = 0.0f, srr = 0.0f, scc = 0.0f;
            for (int32_t dr = -h; dr <= h; ++dr)
                for (int32_t dc = -h; dc <= h; ++dc) {
                    int32_t r = row + dr, c = col + dc;
                    if (r < 0) r = 0;
                    else if (r >= rows) r = rows - 1;
                    if (c < 0) c = 0;
                    else if (c >= cols) c = cols - 1;
                    const float w = g->w[(dr + h) * g->size + dc + h];
                    const float val = (float)im->d[r * cols + c];
                    s0 += val * w;
                    sr += (float)dr * val * w;
                    sc += (float)dc * val * w;
                    src += (float)(dr * dc) * val * w;
                    srr += (float)(dr * dr) * val * w;
                    scc += (float)(dc * dc) * val * w;
                }
            const size_t i = (size_t)row * cols + col;
            S[i] = s0, S[n + i] = sr, S[2 * n + i] = sc, S[3 * n + i] = src, S[4 * n + i] = srr, S[5 * n + i] = scc;
        }
}

static float dof_interpolate(const float *m, int32_t rows, int32_t cols, float row, float col) { /* shim slam_basic_math.h */
    const float max_r = (float)(rows - 1), max_c = (float)(cols - 1);
    const float r = row < 0.0f ? 0.0f : (row > max_r ? max_r : row);
    const float c = col < 0.0f ? 0.0f : (col > max_c ? max_c : col);
    const float fr = floorf(r), fc = floorf(c);
    const int32_t r0 = (int32_t)fr, c0 = (int32_t)fc;
    const int32_t r1 = r0 + 1 < rows ? r0 + 1 : rows - 1, c1 = c0 + 1 < cols ? c0 + 1 : cols - 1;
    const float dr = r - fr, dc = c - fc;
    const float ir = 1.0f - dr, ic = 1.0f - dc;
    return ir * ic * m[r0 * cols + c0] + ir * dc * m[r0 * cols + c1] + dr * ic * m[r1 * cols + c0] + dr * dc * m[r1 * cols + c1];
}

/* :259-313 / :315-345: polynomial-expansion coefficients from the six moments */
static void dof_coefficients(const dof_kernel_t *g, float S0, float Sr, float Sc, float Src, float Srr, float Scc, float *A, float *b) {
    const float D = g->k4 - g->k2 * g->k2;
    const float E = g->k22 - g->k2 * g->k2;
    const float inv_D_plus_E = 1.0f / (D + E + 1e-6f);
    const float inv_D_minus_E = 1.0f / (D - E + 1e-6f);
    const float term1 = (Srr + Scc - 2.0f * g->k2 * S0) * inv_D_plus_E;
    const float term2 = (Srr - Scc) * inv_D_minus_E;
    const float a = 0.5f * (term1 + term2);
    const float b_coeff = 0.5f * (term1 - term2);
    const float c_coeff = Src / (g->k22 + 1e-6f);
    A[0] = a;
    A[1] = 0.5f * c_coeff;
    A[2] = A[1];
    A[3] = b_coeff;
    b[0] = Sr / (g->k2 + 1e-6f);
    b[1] = Sc / (g->k2 + 1e-6f);
}

/* :197-257 ComputeFlowByPixel */
static void dof_flow_pixel(const ftko_dense_flow_params *o, const dof_kernel_t *g, const float *S_ref, const float *S_cur, int32_t rows, int32_t cols,
                           int32_t row, int32_t col, float *flow_r, float *flow_c) {
    const size_t n = (size_t)rows * cols, i = (size_t)row * cols + col;
    float A1[4], b1[2];
    dof_coefficients(g, S_ref[i], S_ref[n + i], S_ref[2 * n + i], S_ref[3 * n + i], S_ref[4 * n + i], S_ref[5 * n + i], A1, b1);
    for (int32_t iter = 0; iter < o->max_iteration; ++iter) {
        const float sample_r = (float)row + flow_r[i], sample_c = (float)col + flow_c[i];
        float A2[4], b2[2];
        dof_coefficients(g, dof_interpolate(S_cur, rows, cols, sample_r, sample_c), dof_interpolate(S_cur + n, rows, cols, sample_r, sample_c),
                         dof_interpolate(S_cur + 2 * n, rows, cols, sample_r, sample_c), dof_interpolate(S_cur + 3 * n, rows, cols, sample_r, sample_c),
                         dof_interpolate(S_cur + 4 * n, rows, cols, sample_r, sample_c), dof_interpolate(S_cur + 5 * n, rows, cols, sample_r, sample_c), A2,
                         b2);
        float M[4], bd[2];
        for (int k = 0; k < 4; ++k) M[k] = ((A1[k] + A2[k]) * 0.5f) * 2.0f; /* A_avg = (A1 + A2) * 0.5; M = A_avg * 2 */
        bd[0] = b1[0] - b2[0], bd[1] = b1[1] - b2[1];
        /* MtM = M^T * M, Mtb = M^T * b_diff: r(i,j) = Mt(i,0) M(0,j) + Mt(i,1) M(1,j) */
        const float m00 = M[0] * M[0] + M[2] * M[2], m01 = M[0] * M[1] + M[2] * M[3];
        const float m10 = M[1] * M[0] + M[3] * M[2], m11 = M[1] * M[1] + M[3] * M[3];
        const float t0 = M[0] * bd[0] + M[2] * bd[1], t1 = M[1] * bd[0] + M[3] * bd[1];
        const float lambda = 0.1f * (m00 + m11) + 1.0f;
        /* H = MtM + Identity * lambda */
        const float h00 = m00 + 1.0f * lambda, h01 = m01 + 0.0f * lambda, h10 = m10 + 0.0f * lambda, h11 = m11 + 1.0f * lambda;
        const float invdet = 1.0f / (h00 * h11 - h10 * h01);
        const float i00 = h11 * invdet, i10 = -h10 * invdet, i01 = -h01 * invdet, i11 = h00 * invdet;
        float d0 = i00 * t0 + i01 * t1, d1 = i10 * t0 + i11 * t1;
        const float step_norm = sqrtf(d0 * d0 + d1 * d1);
        if (step_norm > o->max_delta_flow_step) {
            const float sc = o->max_delta_flow_step / step_norm;
            d0 = d0 * sc, d1 = d1 * sc;
        }
        flow_r[i] += d0;
        flow_c[i] += d1;
        if (d0 * d0 + d1 * d1 < o->max_converge_step) break;
    }
}

static int cmp_float(const void *a, const void *b) {
    const float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* :347-371 SmoothFlow: 3x3 median with clamped borders */
static void dof_median(float *flow, int32_t rows, int32_t cols) {
    float *out = (float *)malloc(sizeof(float) * (size_t)rows * cols);
    for (int32_t r = 0; r < rows; ++r)
        for (int32_t c = 0; c < cols; ++c) {
            float w[9];
            int k = 0;
            for (int32_t dr = -1; dr <= 1; ++dr)
                for (int32_t dc = -1; dc <= 1; ++dc) {
                    int32_t nr = r + dr, nc = c + dc;
                    nr = nr < 0 ? 0 : (nr > rows - 1 ? rows - 1 : nr);
                    nc = nc < 0 ? 0 : (nc > cols - 1 ? cols - 1 : nc);
                    w[k++] = flow[nr * cols + nc];
                }
            qsort(w, 9, sizeof(float), cmp_float);
            out[r * cols + c] = w[4];
        }
    memcpy(flow, out, sizeof(float) * (size_t)rows * cols);
    free(out);
}

/* :7-33 Track(GrayImage, GrayImage, flow) */
static void dof_track_level(const ftko_dense_flow_params *o, const dof_kernel_t *g, const image_t *ref, const image_t *cur, float *flow_r, float *flow_c) {
    const int32_t rows = ref->rows, cols = ref->cols;
    const size_t n = (size_t)rows * cols;
    float *S_ref = (float *)malloc(sizeof(float) * 6 * n), *S_cur = (float *)malloc(sizeof(float) * 6 * n);
    dof_moments(g, ref, S_ref);
    dof_moments(g, cur, S_cur);
    for (int32_t row = 0; row < rows; ++row)
        for (int32_t col = 0; col < cols; ++col) dof_flow_pixel(o, g, S_ref, S_cur, rows, cols, row, col, flow_r, flow_c);
    dof_median(flow_r, rows, cols);
    dof_median(flow_c, rows, cols);
    free(S_ref);
    free(S_cur);
}

int ftko_dense_flow_track(const ftko_dense_flow_params *params, int32_t levels, const uint8_t *const *ref_levels, const uint8_t *const *cur_levels,
                          const int32_t *rows, const int32_t *cols, int32_t single_level, int32_t flow_valid, float *flow_row, float *flow_col) {
    if (levels < 1 || levels > 16 || params->half_patch_size < 0 || params->half_patch_size > 15) return 0;
    dof_kernel_t g;
    dof_init_kernel(&g, params->half_patch_size);
    image_t ref, cur;
    if (single_level) { /* :7-33; the moment maps of `cur` must match `ref`: same size is implied by the shared rows / cols */
        ref.d = ref_levels[0], ref.rows = rows[0], ref.cols = cols[0];
        cur.d = cur_levels[0], cur.rows = rows[0], cur.cols = cols[0];
        if (!flow_valid) {
            memset(flow_row, 0, sizeof(float) * (size_t)rows[0] * cols[0]); /* :18-23 */
            memset(flow_col, 0, sizeof(float) * (size_t)rows[0] * cols[0]);
        }
        dof_track_level(params, &g, &ref, &cur, flow_row, flow_col);
        return 1;
    }
    /* :35-85 coarse to fine */
    const int32_t top = levels - 1;
    size_t n = (size_t)rows[top] * cols[top];
    float *fr = (float *)calloc(n, sizeof(float)), *fc = (float *)calloc(n, sizeof(float));
    for (int32_t l = top; l >= 0; --l) {
        ref.d = ref_levels[l], ref.rows = rows[l], ref.cols = cols[l];
        cur.d = cur_levels[l], cur.rows = rows[l], cur.cols = cols[l];
        dof_track_level(params, &g, &ref, &cur, fr, fc);
        if (l == 0) break;
        const int32_t nr = rows[l - 1], nc = cols[l - 1];
        float *ur = (float *)malloc(sizeof(float) * (size_t)nr * nc), *uc = (float *)malloc(sizeof(float) * (size_t)nr * nc);
        for (int32_t r = 0; r < nr; ++r)
            for (int32_t c = 0; c < nc; ++c) {
                const float frow = (float)r * 0.5f, fcol = (float)c * 0.5f; /* :72-73 */
                ur[(size_t)r * nc + c] = dof_interpolate(fr, rows[l], cols[l], frow, fcol) * 2.0f;
                uc[(size_t)r * nc + c] = dof_interpolate(fc, rows[l], cols[l], frow, fcol) * 2.0f;
            }
        free(fr);
        free(fc);
        fr = ur, fc = uc;
    }
    memcpy(flow_row, fr, sizeof(float) * (size_t)rows[0] * cols[0]);
    memcpy(flow_col, fc, sizeof(float) * (size_t)rows[0] * cols[0]);
    free(fr);
    free(fc);
    return 1;
}

/* ------------------------------------------------------------------------------------------------------------
 * Feature detection + BRIEF description -- PARITY UNPINNED.
 * The reference calls feature_detector::FeaturePointHarrisDetector::DetectGoodFeatures and feature_detector::BriefDescriptor::
 * Compute (test/test_descriptor_matcher_brief.cpp:59-76, test/test_optical_flow.cpp:60-66), which live in the sibling
 * repository Feature_Detector (CMakeLists.txt:24-29); that repository is not vendored, not pinned and not present, and the
 * reference holds no fixtures of detector output.  What is frozen here is the published algorithm behind those names, under
 * the option names the call sites use:
 *   response  Harris & Stephens 1988 / Shi & Tomasi 1994 on the (2h+1)^2 box-summed structure tensor of central differences
 *             (gx = I(r,c+1)-I(r,c-1), gy = I(r+1,c)-I(r-1,c); integer sums, then a = Sxx/(4n), b = Sxy/(4n), c = Syy/(4n));
 *             Harris = a c - b b - k (a+c)^2, Shi-Tomasi = ((a+c) - sqrt((a-c)^2 + 4 b b)) / 2, fp32, no contraction;
 *             defined where the whole window and its gradients are inside the image, -inf elsewhere.
 *   selection candidates = pixels with response >= kMinValidResponse, visited by falling response (ties: row-major index);
 *             a candidate is taken unless an already taken or pre-existing feature lies within |drow| < kMinFeatureDistance
 *             and |dcol| < kMinFeatureDistance (pre-existing features count at their truncated pixel, and only when inside
 *             the image); stops after `needed` features.  Output (x = col, y = row).
 *   BRIEF     Calonder et al. 2010: bit k = I(p + a_k) < I(p + b_k) for a fixed list of offset pairs inside +-kHalfPatchSize,
 *             on the raw image at the truncated feature position; a feature whose patch leaves the image gets an all-zero
 *             descriptor and valid = 0.  Word w holds pairs 32w .. 32w+31, least significant bit first (the packed layout of
 *             ftko_match_brief_* / ftk_match_hamming256).  The pair list is an INPUT, so a caller that owns the upstream list
 *             can pass it; ftko_brief_pattern is the default list (xorshift32, documented below).
 * ---------------------------------------------------------------------------------------------------------- */
static float detector_response_at(const ftko_detector_params *p, const uint8_t *image, int32_t cols, int32_t r, int32_t c) {
    const int32_t h = p->half_patch;
    int32_t sxx = 0, syy = 0, sxy = 0;
    for (int32_t dr = -h; dr <= h; ++dr) {
        for (int32_t dc = -h; dc <= h; ++dc) {
            const uint8_t *q = image + (size_t)(r + dr) * cols + (c + dc);
            const int32_t gx = (int32_t)q[1] - (int32_t)q[-1], gy = (int32_t)q[cols] - (int32_t)q[-cols];
            sxx += gx * gx, syy += gy * gy, sxy += gx * gy;
        }
    }
    const float inv = 1.0f / (float)(4 * (2 * h + 1) * (2 * h + 1));
    const float a = (float)sxx * inv, b = (float)sxy * inv, cc = (float)syy * inv;
    if (p->kind == 0) {
        const float det = a * cc - b * b, tr = a + cc;
        return det - p->harris_k * (tr * tr);
    }
    const float d = a - cc;
    const float disc = d * d + 4.0f * (b * b);
    return 0.5f * ((a + cc) - sqrtf(disc));
}

int ftko_detect_response(const ftko_detector_params *params, const uint8_t *image, int32_t rows, int32_t cols, float *response) {
    if (params->half_patch < 1 || params->half_patch > 3 || rows <= 0 || cols <= 0) return 0;
    const int32_t m = params->half_patch + 1;
    for (int32_t r = 0; r < rows; ++r)
        for (int32_t c = 0; c < cols; ++c)
            response[(size_t)r * cols + c] =
                (r >= m && r < rows - m && c >= m && c < cols - m) ? detector_response_at(params, image, cols, r, c) : -INFINITY;
    return 1;
}

typedef struct {
    float response;
    int32_t index;
} detector_candidate_t;

static int detector_candidate_cmp(const void *pa, const void *pb) {
    const detector_candidate_t *a = (const detector_candidate_t *)pa, *b = (const detector_candidate_t *)pb;
    if (a->response != b->response) return a->response > b->response ? -1 : 1;
    return a->index < b->index ? -1 : (a->index > b->index ? 1 : 0);
}

static void detector_mask(uint8_t *mask, int32_t rows, int32_t cols, int32_t r, int32_t c, int32_t dist) {
    for (int32_t rr = r - (dist - 1); rr <= r + (dist - 1); ++rr)
        for (int32_t qc = c - (dist - 1); qc <= c + (dist - 1); ++qc)
            if (rr >= 0 && rr < rows && qc >= 0 && qc < cols) mask[(size_t)rr * cols + qc] = 1;
}

/* Returns the number of features written (<= needed), -1 on bad parameters. */
int ftko_detect_features(const ftko_detector_params *params, const uint8_t *image, int32_t rows, int32_t cols, const float *existing_uv,
                         int32_t n_existing, int32_t needed, float *out_uv, float *out_response) {
    const size_t n = (size_t)rows * cols;
    float *response = (float *)malloc(sizeof(float) * (n ? n : 1));
    if (!ftko_detect_response(params, image, rows, cols, response)) {
        free(response);
        return -1;
    }
    uint8_t *mask = (uint8_t *)calloc(n, 1);
    detector_candidate_t *cand = (detector_candidate_t *)malloc(sizeof(detector_candidate_t) * n);
    const int32_t dist = params->min_distance;
    for (int32_t i = 0; i < n_existing; ++i) { /* (int) casts truncate, like the reference's pixel look-ups */
        const float x = existing_uv[2 * i], y = existing_uv[2 * i + 1];
        if (!(x >= 0.0f && y >= 0.0f && x < (float)cols && y < (float)rows)) continue; /* outside the image (or NaN): masks nothing */
        if (dist > 0) detector_mask(mask, rows, cols, (int32_t)y, (int32_t)x, dist);
    }
    size_t n_cand = 0;
    for (size_t i = 0; i < n; ++i)
        if (response[i] >= params->min_response) cand[n_cand].response = response[i], cand[n_cand].index = (int32_t)i, ++n_cand;
    qsort(cand, n_cand, sizeof(detector_candidate_t), detector_candidate_cmp);
    int32_t n_out = 0;
    for (size_t k = 0; k < n_cand && n_out < needed; ++k) {
        if (mask[cand[k].index]) continue;
        const int32_t r = cand[k].index / cols, c = cand[k].index % cols;
        out_uv[2 * n_out] = (float)c, out_uv[2 * n_out + 1] = (float)r;
        if (out_response) out_response[n_out] = cand[k].response;
        ++n_out;
        if (dist > 0) detector_mask(mask, rows, cols, r, c, dist);
    }
    free(cand);
    free(mask);
    free(response);
    return n_out;
}

/* Default pair list: xorshift32 (13, 17, 5) seeded with `seed` (0 -> 0x9E3779B9); every coordinate = draw % (2 half + 1) - half,
 * in the order (drow_a, dcol_a, drow_b, dcol_b) per pair; a pair with a == b is redrawn. */
void ftko_brief_pattern(int32_t n_bits, int32_t half_patch, uint32_t seed, int8_t *pattern) {
    uint32_t x = seed ? seed : 0x9E3779B9u;
    const uint32_t span = (uint32_t)(2 * half_patch + 1);
    for (int32_t k = 0; k < n_bits; ++k) {
        int8_t v[4];
        do {
            for (int j = 0; j < 4; ++j) {
                x ^= x << 13, x ^= x >> 17, x ^= x << 5;
                v[j] = (int8_t)((int32_t)(x % span) - half_patch);
            }
        } while (half_patch > 0 && v[0] == v[2] && v[1] == v[3]);
        for (int j = 0; j < 4; ++j) pattern[4 * k + j] = v[j];
    }
}

int ftko_describe_brief(const uint8_t *image, int32_t rows, int32_t cols, const float *uv, int32_t n, const int8_t *pattern, int32_t n_bits,
                        int32_t half_patch, uint32_t *desc, uint8_t *valid) {
    if (n_bits <= 0 || n_bits % 32 != 0 || half_patch < 0) return 0;
    const int32_t words = n_bits / 32;
    for (int32_t i = 0; i < n; ++i) {
        uint32_t *d = desc + (size_t)i * words;
        for (int32_t w = 0; w < words; ++w) d[w] = 0;
        const float x = uv[2 * i], y = uv[2 * i + 1];
        /* the float tests keep NaN and out-of-range positions away from the int casts */
        const int ok = x >= (float)half_patch && y >= (float)half_patch && x < (float)(cols - half_patch) && y < (float)(rows - half_patch);
        if (valid) valid[i] = (uint8_t)ok;
        if (!ok) continue;
        const int32_t c = (int32_t)x, r = (int32_t)y;
        for (int32_t k = 0; k < n_bits; ++k) {
            const int8_t *q = pattern + 4 * k;
            const uint8_t va = image[(size_t)(r + q[0]) * cols + (c + q[1])], vb = image[(size_t)(r + q[2]) * cols + (c + q[3])];
            if (va < vb) d[k >> 5] |= 1u << (k & 31);
        }
    }
    return 1;
}
