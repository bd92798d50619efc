// K10: dense optical flow, Gunnar Farneback's polynomial expansion.
//
// Replaces DenseOpticalFlow::Track (src/dense_optical_flow_tracker/dense_optical_flow.cpp:7-85) and its helpers: Gaussian
// weighted moment maps (:139-195), the per-pixel Gauss-Newton refinement (:197-257, :259-345), the 3x3 median (:347-371)
// and the 2x bilinear up-sampling between pyramid levels (:63-79).  Every pixel is independent, so one thread owns one pixel
// and runs the reference's own sequence of fp32 operations (no FMA contraction): the flow is bit-identical.
//   MomentsKernel  : 6 moment planes of one image, (2h+1)^2 clamped taps per pixel, taps in the reference's order
//   FlowKernel     : A1 / b1 from the reference planes, <= kMaxIteration bilinear look-ups of the 6 current planes
//   MedianKernel   : exact median of the 3x3 clamped neighbourhood (19-exchange network)
//   UpsampleKernel : flow of the next finer level = 2 * bilinear(flow, r / 2, c / 2)
// HBM traffic per level is ~100 B per pixel (6 + 6 planes written once, read a few times from L2); the kernels are latency /
// L2 bound at 752x480 and the whole 4-level pass is a few hundred microseconds.
#include <cmath>

#include "klt_device.cuh"

namespace ftk {

namespace {

constexpr int kDofMaxHalf = 7;
constexpr int kDofThreads = 256;

struct DofKernel {
    float w[(2 * kDofMaxHalf + 1) * (2 * kDofMaxHalf + 1)];
    float k2, k4, k22;
    int half, size;
};

// :87-137 on the host (expf of the C library, as the reference).
DofKernel MakeKernel(int half) {
    DofKernel g{};
    g.half = half;
    g.size = 2 * half + 1;
    if (half == 0) {
        g.w[0] = 1.0f;
        return g;
    }
    const float sigma = 1.0f;
    const float sigma2 = sigma * sigma;
    volatile float sum = 0.0f;  // volatile: plain sequential fp32 sums, no re-association by the host compiler
    for (int row = 0; row < g.size; ++row)
        for (int col = 0; col < g.size; ++col) {
            const int dr = row - half, dc = col - half;
            g.w[row * g.size + col] = expf(-0.5f * static_cast<float>(dr * dr + dc * dc) / sigma2);
            sum = sum + g.w[row * g.size + col];
        }
    for (int i = 0; i < g.size * g.size; ++i) g.w[i] = g.w[i] / sum;
    volatile float k2 = 0.0f, k4 = 0.0f, k22 = 0.0f;
    for (int row = 0; row < g.size; ++row)
        for (int col = 0; col < g.size; ++col) {
            const float dr = static_cast<float>(row - half), dc = static_cast<float>(col - half);
            const float w = g.w[row * g.size + col];
            volatile float t = w * dr;
            t = t * dr;
            k2 = k2 + t;
            t = t * dr;
            t = t * dr;
            k4 = k4 + t;
            volatile float u = w * dr;
            u = u * dr;
            u = u * dc;
            u = u * dc;
            k22 = k22 + u;
        }
    g.k2 = k2, g.k4 = k4, g.k22 = k22;
    return g;
}

// :139-195.  S = 6 planes of n floats: S0, Srow, Scol, Srowcol, Srowrow, Scolcol.
__global__ void __launch_bounds__(kDofThreads) DofMomentsKernel(Img im, DofKernel g, float *__restrict__ S) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= im.cols) return;
    const int h = g.half;
    float s0 = 0.0f, sr = 0.0f, sc = 0.0f, src = 0.0f, srr = 0.0f, scc = 0.0f;
    for (int dr = -h; dr <= h; ++dr) {
        const int r = min(max(row + dr, 0), im.rows - 1);
        const uint8_t *line = im.p + r * im.pitch;
        for (int dc = -h; dc <= h; ++dc) {
            const int c = min(max(col + dc, 0), im.cols - 1);
            const float w = g.w[(dr + h) * g.size + dc + h];
            const float val = static_cast<float>(__ldg(line + c));
            s0 = fadd(s0, fmul(val, w));
            sr = fadd(sr, fmul(fmul(static_cast<float>(dr), val), w));
            sc = fadd(sc, fmul(fmul(static_cast<float>(dc), val), w));
            src = fadd(src, fmul(fmul(static_cast<float>(dr * dc), val), w));
            srr = fadd(srr, fmul(fmul(static_cast<float>(dr * dr), val), w));
            scc = fadd(scc, fmul(fmul(static_cast<float>(dc * dc), val), w));
        }
    }
    const size_t n = static_cast<size_t>(im.rows) * im.cols, i = static_cast<size_t>(row) * im.cols + col;
    S[i] = s0, S[n + i] = sr, S[2 * n + i] = sc, S[3 * n + i] = src, S[4 * n + i] = srr, S[5 * n + i] = scc;
}

// slam_utility::Utility::Interpolate as frozen in oracle/shim/slam_basic_math.h; the six planes share position and weights.
struct DofTap {
    int i00, i01, i10, i11;
    float w00, w01, w10, w11;
};
__device__ __forceinline__ DofTap MakeTap(int rows, int cols, float row, float col) {
    const float max_r = static_cast<float>(rows - 1), max_c = static_cast<float>(cols - 1);
    const float r = row < 0.0f ? 0.0f : (row > max_r ? max_r : row);
    const float c = col < 0.0f ? 0.0f : (col > max_c ? max_c : col);
    const float fr = floorf(r), fc = floorf(c);
    const int r0 = static_cast<int>(fr), c0 = static_cast<int>(fc);
    const int r1 = r0 + 1 < rows ? r0 + 1 : rows - 1, c1 = c0 + 1 < cols ? c0 + 1 : cols - 1;
    const float dr = fsub(r, fr), dc = fsub(c, fc);
    const float ir = fsub(1.0f, dr), ic = fsub(1.0f, dc);
    DofTap t;
    t.i00 = r0 * cols + c0, t.i01 = r0 * cols + c1, t.i10 = r1 * cols + c0, t.i11 = r1 * cols + c1;
    t.w00 = fmul(ir, ic), t.w01 = fmul(ir, dc), t.w10 = fmul(dr, ic), t.w11 = fmul(dr, dc);
    return t;
}
__device__ __forceinline__ float Lookup(const float *__restrict__ m, const DofTap &t) {
    return fadd(fadd(fadd(fmul(t.w00, __ldg(m + t.i00)), fmul(t.w01, __ldg(m + t.i01))), fmul(t.w10, __ldg(m + t.i10))), fmul(t.w11, __ldg(m + t.i11)));
}

struct DofConst {  // :294-297: the same for every pixel
    float inv_D_plus_E, inv_D_minus_E, two_k2, k22_eps, k2_eps;
};

// :259-313 / :315-345
__device__ __forceinline__ void Coefficients(const DofConst &k, float S0, float Sr, float Sc, float Src, float Srr, float Scc, float (&A)[4], float (&b)[2]) {
    const float term1 = fmul(fsub(fadd(Srr, Scc), fmul(k.two_k2, S0)), k.inv_D_plus_E);
    const float term2 = fmul(fsub(Srr, Scc), k.inv_D_minus_E);
    const float a = fmul(0.5f, fadd(term1, term2));
    const float b_coeff = fmul(0.5f, fsub(term1, term2));
    const float c_coeff = fdiv(Src, k.k22_eps);
    A[0] = a;
    A[1] = fmul(0.5f, c_coeff);
    A[2] = A[1];
    A[3] = b_coeff;
    b[0] = fdiv(Sr, k.k2_eps);
    b[1] = fdiv(Sc, k.k2_eps);
}

// :197-257 for every pixel.
__global__ void __launch_bounds__(kDofThreads) DofFlowKernel(const float *__restrict__ S_ref, const float *__restrict__ S_cur, int rows, int cols, DofConst k,
                                                           int max_iteration, float max_step, float converge, float *__restrict__ flow_r,
                                                           float *__restrict__ flow_c) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= cols) return;
    const size_t n = static_cast<size_t>(rows) * cols, i = static_cast<size_t>(row) * cols + col;
    float A1[4], b1[2];
    Coefficients(k, S_ref[i], S_ref[n + i], S_ref[2 * n + i], S_ref[3 * n + i], S_ref[4 * n + i], S_ref[5 * n + i], A1, b1);
    float fr = flow_r[i], fc = flow_c[i];
    for (int iter = 0; iter < max_iteration; ++iter) {
        const DofTap t = MakeTap(rows, cols, fadd(static_cast<float>(row), fr), fadd(static_cast<float>(col), fc));
        float A2[4], b2[2];
        Coefficients(k, Lookup(S_cur, t), Lookup(S_cur + n, t), Lookup(S_cur + 2 * n, t), Lookup(S_cur + 3 * n, t), Lookup(S_cur + 4 * n, t),
                     Lookup(S_cur + 5 * n, t), A2, b2);
        float M[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) M[q] = fmul(fmul(fadd(A1[q], A2[q]), 0.5f), 2.0f);
        const float bd0 = fsub(b1[0], b2[0]), bd1 = fsub(b1[1], b2[1]);
        const float m00 = fadd(fmul(M[0], M[0]), fmul(M[2], M[2])), m01 = fadd(fmul(M[0], M[1]), fmul(M[2], M[3]));
        const float m10 = fadd(fmul(M[1], M[0]), fmul(M[3], M[2])), m11 = fadd(fmul(M[1], M[1]), fmul(M[3], M[3]));
        const float t0 = fadd(fmul(M[0], bd0), fmul(M[2], bd1)), t1 = fadd(fmul(M[1], bd0), fmul(M[3], bd1));
        const float lambda = fadd(fmul(0.1f, fadd(m00, m11)), 1.0f);
        const float h00 = fadd(m00, fmul(1.0f, lambda)), h01 = fadd(m01, fmul(0.0f, lambda)), h10 = fadd(m10, fmul(0.0f, lambda)),
                    h11 = fadd(m11, fmul(1.0f, lambda));
        const float invdet = fdiv(1.0f, fsub(fmul(h00, h11), fmul(h10, h01)));
        const float i00 = fmul(h11, invdet), i10 = fmul(-h10, invdet), i01 = fmul(-h01, invdet), i11 = fmul(h00, invdet);
        float d0 = fadd(fmul(i00, t0), fmul(i01, t1)), d1 = fadd(fmul(i10, t0), fmul(i11, t1));
        const float step_norm = __fsqrt_rn(fadd(fmul(d0, d0), fmul(d1, d1)));
        if (step_norm > max_step) {
            const float sc = fdiv(max_step, step_norm);
            d0 = fmul(d0, sc), d1 = fmul(d1, sc);
        }
        fr = fadd(fr, d0);
        fc = fadd(fc, d1);
        if (fadd(fmul(d0, d0), fmul(d1, d1)) < converge) break;
    }
    flow_r[i] = fr;
    flow_c[i] = fc;
}

__device__ __forceinline__ void Sort2(float &a, float &b) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo, b = hi;
}
// :347-371 (one flow component)
__global__ void __launch_bounds__(kDofThreads) DofMedianKernel(const float *__restrict__ in, int rows, int cols, float *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= cols) return;
    float v[9];
    int q = 0;
#pragma unroll
    for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
        for (int dc = -1; dc <= 1; ++dc) {
            const int nr = min(max(r + dr, 0), rows - 1), nc = min(max(c + dc, 0), cols - 1);
            v[q++] = __ldg(in + static_cast<size_t>(nr) * cols + nc);
        }
    // 19-exchange median-of-9 network (Paeth)
    Sort2(v[1], v[2]); Sort2(v[4], v[5]); Sort2(v[7], v[8]);
    Sort2(v[0], v[1]); Sort2(v[3], v[4]); Sort2(v[6], v[7]);
    Sort2(v[1], v[2]); Sort2(v[4], v[5]); Sort2(v[7], v[8]);
    Sort2(v[0], v[3]); Sort2(v[5], v[8]); Sort2(v[4], v[7]);
    Sort2(v[3], v[6]); Sort2(v[1], v[4]); Sort2(v[2], v[5]);
    Sort2(v[4], v[7]); Sort2(v[4], v[2]); Sort2(v[6], v[4]);
    Sort2(v[4], v[2]);
    out[static_cast<size_t>(r) * cols + c] = v[4];
}

// :63-79
__global__ void __launch_bounds__(kDofThreads) DofUpsampleKernel(const float *__restrict__ in_r, const float *__restrict__ in_c, int rows, int cols,
                                                               int out_rows, int out_cols, float *__restrict__ out_r, float *__restrict__ out_c) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= out_cols) return;
    const DofTap t = MakeTap(rows, cols, fmul(static_cast<float>(r), 0.5f), fmul(static_cast<float>(c), 0.5f));
    const size_t i = static_cast<size_t>(r) * out_cols + c;
    out_r[i] = fmul(Lookup(in_r, t), 2.0f);
    out_c[i] = fmul(Lookup(in_c, t), 2.0f);
}

}  // namespace

// d_flow_r / d_flow_c: rows[0] x cols[0] device floats (in/out for the single-level overload with use_initial_flow).
int LaunchDenseFlow(ftk_context *ctx, const ftk_dense_flow_params &p, const PyramidView &ref, const PyramidView &cur, int ref_image, int cur_image,
                    bool single_level, bool use_initial_flow, float *d_flow_r, float *d_flow_c) {
    if (p.half_patch_size < 0) return SetError(ctx, FTK_ERR_INVALID_ARGUMENT, "kHalfPatchSize < 0");  // :88
    if (p.half_patch_size > kDofMaxHalf) return SetError(ctx, FTK_ERR_UNSUPPORTED, "kHalfPatchSize > %d", kDofMaxHalf);
    const DofKernel g = MakeKernel(p.half_patch_size);
    DofConst k;
    {
        volatile float k2sq = g.k2 * g.k2;
        volatile float D = g.k4 - k2sq, E = g.k22 - k2sq;
        volatile float dpe = D + E, dme = D - E;
        dpe = dpe + 1e-6f, dme = dme + 1e-6f;
        k.inv_D_plus_E = 1.0f / dpe;
        k.inv_D_minus_E = 1.0f / dme;
        k.two_k2 = 2.0f * g.k2;
        k.k22_eps = g.k22 + 1e-6f;
        k.k2_eps = g.k2 + 1e-6f;
    }
    const int levels = single_level ? 1 : ref.levels;
    const size_t n0 = static_cast<size_t>(ref.rows[0]) * ref.cols[0];
    // scratch: 2 x 6 moment planes + 2 x 2 flow planes (ping / pong), all sized for level 0
    if (int rc = EnsureDevice(ctx, ctx->d_work0, sizeof(float) * 12 * n0)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_work1, sizeof(float) * 4 * n0)) return rc;
    float *S_ref = static_cast<float *>(ctx->d_work0.ptr), *S_cur = S_ref + 6 * n0;
    float *fa_r = static_cast<float *>(ctx->d_work1.ptr), *fa_c = fa_r + n0, *fb_r = fa_c + n0, *fb_c = fb_r + n0;
    cudaStream_t st = ctx->stream;

    const int top = levels - 1;
    const size_t n_top = static_cast<size_t>(ref.rows[top]) * ref.cols[top];
    if (single_level && use_initial_flow) {
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(fa_r, d_flow_r, sizeof(float) * n0, cudaMemcpyDeviceToDevice, st));
        FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(fa_c, d_flow_c, sizeof(float) * n0, cudaMemcpyDeviceToDevice, st));
    } else {
        FTK_CUDA_CHECK(ctx, cudaMemsetAsync(fa_r, 0, sizeof(float) * n_top, st));  // :18-23, :44-45
        FTK_CUDA_CHECK(ctx, cudaMemsetAsync(fa_c, 0, sizeof(float) * n_top, st));
    }
    for (int level = top; level >= 0; --level) {
        Img ri, ci;
        ri.p = ref.base[level] + ref_image * ref.image_stride[level], ri.rows = ref.rows[level], ri.cols = ref.cols[level], ri.pitch = ref.pitch[level];
        ci.p = cur.base[level] + cur_image * cur.image_stride[level], ci.rows = cur.rows[level], ci.cols = cur.cols[level], ci.pitch = cur.pitch[level];
        const int rows = ri.rows, cols = ri.cols;
        const dim3 grid((cols + kDofThreads - 1) / kDofThreads, rows);
        DofMomentsKernel<<<grid, kDofThreads, 0, st>>>(ri, g, S_ref);
        DofMomentsKernel<<<grid, kDofThreads, 0, st>>>(ci, g, S_cur);
        DofFlowKernel<<<grid, kDofThreads, 0, st>>>(S_ref, S_cur, rows, cols, k, p.max_iteration, p.max_delta_flow_step, p.max_converge_step, fa_r, fa_c);
        DofMedianKernel<<<grid, kDofThreads, 0, st>>>(fa_r, rows, cols, fb_r);
        DofMedianKernel<<<grid, kDofThreads, 0, st>>>(fa_c, rows, cols, fb_c);
        ctx->launches += 5;
        FTK_CUDA_CHECK(ctx, cudaGetLastError());
        if (level == 0) break;
        const int nr = ref.rows[level - 1], nc = ref.cols[level - 1];
        const dim3 ugrid((nc + kDofThreads - 1) / kDofThreads, nr);
        DofUpsampleKernel<<<ugrid, kDofThreads, 0, st>>>(fb_r, fb_c, rows, cols, nr, nc, fa_r, fa_c);
        ++ctx->launches;
        FTK_CUDA_CHECK(ctx, cudaGetLastError());
    }
    FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(d_flow_r, fb_r, sizeof(float) * n0, cudaMemcpyDeviceToDevice, st));
    FTK_CUDA_CHECK(ctx, cudaMemcpyAsync(d_flow_c, fb_c, sizeof(float) * n0, cudaMemcpyDeviceToDevice, st));
    return FTK_OK;
}

}  // namespace ftk
