// ORACLE (test infrastructure).  C entry points around the UNMODIFIED reference hot-path sources, which the
// Makefile compiles where they lie under /root/reference against oracle/shim/.  Nothing here is product code;
// only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load the resulting
// oracle/_ref/libftk_ref.so.
#include <cstring>
#include <memory>
#include <vector>

#include "descriptor_matcher.h"
#include "dense_optical_flow.h"
#include "direct_method_tracker.h"
#include "nn_feature_matcher.h"
#include "optical_flow_affine_klt.h"
#include "optical_flow_basic_klt.h"
#include "optical_flow_lssd_klt.h"

#include "ftk_oracle_types.h"

namespace {

// One padded private copy per level: the reference's bilinear sampler reads the +1 neighbour with weight 0
// when a coordinate sits exactly on the last row/col, i.e. up to cols+1 bytes past
// the image.
struct PaddedLevels {
    std::vector<std::vector<uint8_t>> store;
    std::vector<uint8_t *> ptr;
    void Adopt(int levels, const uint8_t *const *data, const int32_t *rows, const int32_t *cols) {
        store.resize(levels);
        ptr.resize(levels);
        for (int i = 0; i < levels; ++i) {
            const size_t n = static_cast<size_t>(rows[i]) * cols[i];
            store[i].assign(n + cols[i] + 2, 0);
            std::memcpy(store[i].data(), data[i], n);
            ptr[i] = store[i].data();
        }
    }
};

void ApplyOptions(feature_tracker::OpticalFlow &klt, const ftko_klt_params &p) {
    klt.options().kMaxTrackPointsNumber = p.max_track_points;
    klt.options().kMaxIteration = p.max_iteration;
    klt.options().kMaxToleranceLargeStep = p.max_tolerance_large_step;
    klt.options().kPatchRowHalfSize = p.patch_row_half;
    klt.options().kPatchColHalfSize = p.patch_col_half;
    klt.options().kMaxConvergeStep = p.max_converge_step;
    klt.options().kMethod = static_cast<feature_tracker::OpticalFlowMethod>(p.method);
}

std::unique_ptr<feature_tracker::OpticalFlow> MakeTracker(const ftko_klt_params &p) {
    Mat2 predict;
    predict << p.predict[0], p.predict[1], p.predict[2], p.predict[3];
    std::unique_ptr<feature_tracker::OpticalFlow> klt;
    if (p.variant == 0) {
        klt.reset(new feature_tracker::OpticalFlowBasicKlt());
    } else if (p.variant == 1) {
        auto *t = new feature_tracker::OpticalFlowAffineKlt();
        t->predict_affine() = predict;
        klt.reset(t);
    } else if (p.variant == 2) {
        auto *t = new feature_tracker::OpticalFlowLssdKlt();
        t->predict_R_cr() = predict;
        t->consider_patch_luminance() = p.consider_patch_luminance != 0;
        klt.reset(t);
    } else {
        return nullptr;
    }
    ApplyOptions(*klt, p);
    return klt;
}

// BRIEF distance: number of differing elements (test/test_descriptor_matcher_brief.cpp:33-45).
using BriefType = std::vector<uint8_t>;
class BriefMatcher: public feature_tracker::DescriptorMatcher<BriefType> {
    float ComputeDistance(const BriefType &a, const BriefType &b) override {
        if (a.empty() || b.empty()) return static_cast<float>(kMaxInt32);
        int32_t n = 0;
        for (uint32_t k = 0; k < a.size(); ++k) n += (a[k] != b[k]) ? 1 : 0;
        return static_cast<float>(n);
    }
};

// Float descriptor distance 0.5 - 0.5*cos (test/test_descriptor_matcher_superpoint.cpp:32-34, disk:32-34).
// Dynamic-length float vector with the same sequential dot()/norm() as shim::Mat.
struct FloatDesc {
    const float *p = nullptr;
    int32_t n = 0;
    float dot(const FloatDesc &o) const {
        float s = p[0] * o.p[0];
        for (int32_t k = 1; k < n; ++k) s = s + p[k] * o.p[k];
        return s;
    }
    float norm() const { return std::sqrt(dot(*this)); }
};
class CosineMatcher: public feature_tracker::DescriptorMatcher<FloatDesc> {
    float ComputeDistance(const FloatDesc &a, const FloatDesc &b) override { return 0.5f - a.dot(b) / a.norm() / b.norm() * 0.5f; }
};

template <typename T> void ApplyMatcherOptions(T &m, int32_t max_drow, int32_t max_dcol, float max_dist) {
    m.options().kMaxValidPredictRowDistance = max_drow;
    m.options().kMaxValidPredictColDistance = max_dcol;
    m.options().kMaxValidDescriptorDistance = max_dist;
}

std::vector<BriefType> UnpackBrief(const uint8_t *bits, int32_t n, int32_t len) {
    std::vector<BriefType> out(n);
    for (int32_t i = 0; i < n; ++i) out[i].assign(bits + static_cast<size_t>(i) * len, bits + static_cast<size_t>(i + 1) * len);
    return out;
}
std::vector<FloatDesc> WrapFloat(const float *d, int32_t n, int32_t dim) {
    std::vector<FloatDesc> out(n);
    for (int32_t i = 0; i < n; ++i) out[i] = FloatDesc{d + static_cast<size_t>(i) * dim, dim};
    return out;
}
std::vector<Vec2> WrapUv(const float *uv, int32_t n) {
    std::vector<Vec2> out(n);
    for (int32_t i = 0; i < n; ++i) out[i] = Vec2(uv[2 * i], uv[2 * i + 1]);
    return out;
}

}  // namespace


extern "C" {

// Levels 1..levels-1 are written packed into `out` (sum of (rows>>i)*(cols>>i) bytes).
int ftkref_pyramid_build(const uint8_t *image, int32_t rows, int32_t cols, int32_t levels, uint8_t *out) {
    std::vector<uint8_t> raw(image, image + static_cast<size_t>(rows) * cols);
    ImagePyramid pyr;
    pyr.SetPyramidBuff(out, false);
    pyr.SetRawImage(raw.data(), rows, cols);
    return pyr.CreateImagePyramid(static_cast<uint32_t>(levels)) ? 1 : 0;
}

// cur_uv_count / status_count: the SIZES of the caller's vectors on entry (reference semantics,
// optical_flow.cpp:12-19: a size different from n means "no prediction" / "all kNotTracked").
// single_level != 0 selects the GrayImage overload (optical_flow.cpp:28-47) on level 0.
int ftkref_klt_track(const ftko_klt_params *params, int32_t levels, const uint8_t *const *ref_levels, const uint8_t *const *cur_levels,
                     const int32_t *rows, const int32_t *cols, int32_t n, const float *ref_uv, float *cur_uv, int32_t cur_uv_count, uint8_t *status,
                     int32_t status_count, int32_t single_level) {
    auto klt = MakeTracker(*params);
    if (!klt) return 0;
    PaddedLevels ref_store, cur_store;
    ref_store.Adopt(levels, ref_levels, rows, cols);
    cur_store.Adopt(levels, cur_levels, rows, cols);

    std::vector<Vec2> ref_vec = WrapUv(ref_uv, n);
    std::vector<Vec2> cur_vec = WrapUv(cur_uv, cur_uv_count);
    std::vector<uint8_t> status_vec(status, status + status_count);

    bool ok = false;
    if (single_level) {
        GrayImage ref_image(ref_store.ptr[0], rows[0], cols[0]);
        GrayImage cur_image(cur_store.ptr[0], rows[0], cols[0]);
        ok = klt->TrackFeatures(ref_image, cur_image, ref_vec, cur_vec, status_vec);
    } else {
        ImagePyramid ref_pyr, cur_pyr;
        ref_pyr.SetLevels(levels, ref_store.ptr.data(), rows, cols);
        cur_pyr.SetLevels(levels, cur_store.ptr.data(), rows, cols);
        ok = klt->TrackFeatures(ref_pyr, cur_pyr, ref_vec, cur_vec, status_vec);
    }
    if (!ok) return 0;
    for (int32_t i = 0; i < n; ++i) {
        cur_uv[2 * i] = cur_vec[i].x();
        cur_uv[2 * i + 1] = cur_vec[i].y();
        status[i] = status_vec[i];
    }
    return 1;
}

// The reference demo's timed region (test/test_optical_flow.cpp:69-73): CreateImagePyramid x2 + TrackFeatures,
// from raw level-0 images.  `scratch` must hold 2*rows*cols + 2*cols + 8 bytes.
int ftkref_pyramid_and_track(const ftko_klt_params *params, int32_t levels, const uint8_t *ref_image, const uint8_t *cur_image, int32_t rows,
                             int32_t cols, int32_t n, const float *ref_uv, float *cur_uv, uint8_t *status) {
    auto klt = MakeTracker(*params);
    if (!klt) return 0;
    const size_t n_px = static_cast<size_t>(rows) * cols;
    std::vector<uint8_t> ref0(n_px + cols + 2, 0), cur0(n_px + cols + 2, 0), ref_buf(n_px + cols + 2, 0), cur_buf(n_px + cols + 2, 0);
    std::memcpy(ref0.data(), ref_image, n_px);
    std::memcpy(cur0.data(), cur_image, n_px);
    ImagePyramid ref_pyr, cur_pyr;
    ref_pyr.SetPyramidBuff(ref_buf.data(), false);
    cur_pyr.SetPyramidBuff(cur_buf.data(), false);
    ref_pyr.SetRawImage(ref0.data(), rows, cols);
    cur_pyr.SetRawImage(cur0.data(), rows, cols);
    std::vector<Vec2> ref_vec = WrapUv(ref_uv, n);
    std::vector<Vec2> cur_vec;
    std::vector<uint8_t> status_vec;
    ref_pyr.CreateImagePyramid(levels);
    cur_pyr.CreateImagePyramid(levels);
    if (!klt->TrackFeatures(ref_pyr, cur_pyr, ref_vec, cur_vec, status_vec)) return 0;
    for (int32_t i = 0; i < n; ++i) {
        cur_uv[2 * i] = cur_vec[i].x();
        cur_uv[2 * i + 1] = cur_vec[i].y();
        status[i] = status_vec[i];
    }
    return 1;
}

// BRIEF descriptors arrive unpacked: one byte (0/1) per element, `len` elements per descriptor.
// idx_count is the SIZE of the caller's index vector on entry (descriptor_matcher.h:60-62,98-100).
int ftkref_match_brief_force(const uint8_t *ref_bits, int32_t n_ref, const uint8_t *cur_bits, int32_t n_cur, int32_t len, float max_dist, int32_t *idx,
                             int32_t idx_count) {
    BriefMatcher m;
    ApplyMatcherOptions(m, 0, 0, max_dist);
    std::vector<int32_t> out(idx, idx + idx_count);
    if (!m.ForceMatch(UnpackBrief(ref_bits, n_ref, len), UnpackBrief(cur_bits, n_cur, len), out)) return 0;
    std::memcpy(idx, out.data(), sizeof(int32_t) * n_ref);
    return 1;
}

int ftkref_match_brief_nearby(const uint8_t *ref_bits, int32_t n_ref, const uint8_t *cur_bits, int32_t n_cur, int32_t len, const float *pred_uv,
                              const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist, int32_t *idx, int32_t idx_count) {
    BriefMatcher m;
    ApplyMatcherOptions(m, max_drow, max_dcol, max_dist);
    std::vector<int32_t> out(idx, idx + idx_count);
    if (!m.NearbyMatch(UnpackBrief(ref_bits, n_ref, len), UnpackBrief(cur_bits, n_cur, len), WrapUv(pred_uv, n_ref), WrapUv(cur_uv, n_cur), out)) return 0;
    std::memcpy(idx, out.data(), sizeof(int32_t) * n_ref);
    return 1;
}

int ftkref_match_cosine_force(const float *ref, int32_t n_ref, const float *cur, int32_t n_cur, int32_t dim, float max_dist, int32_t *idx,
                              int32_t idx_count) {
    CosineMatcher m;
    ApplyMatcherOptions(m, 0, 0, max_dist);
    std::vector<int32_t> out(idx, idx + idx_count);
    if (!m.ForceMatch(WrapFloat(ref, n_ref, dim), WrapFloat(cur, n_cur, dim), out)) return 0;
    std::memcpy(idx, out.data(), sizeof(int32_t) * n_ref);
    return 1;
}

int ftkref_match_cosine_nearby(const float *ref, int32_t n_ref, const float *cur, int32_t n_cur, int32_t dim, const float *pred_uv, const float *cur_uv,
                               int32_t max_drow, int32_t max_dcol, float max_dist, int32_t *idx, int32_t idx_count) {
    CosineMatcher m;
    ApplyMatcherOptions(m, max_drow, max_dcol, max_dist);
    std::vector<int32_t> out(idx, idx + idx_count);
    if (!m.NearbyMatch(WrapFloat(ref, n_ref, dim), WrapFloat(cur, n_cur, dim), WrapUv(pred_uv, n_ref), WrapUv(cur_uv, n_cur), out)) return 0;
    std::memcpy(idx, out.data(), sizeof(int32_t) * n_ref);
    return 1;
}

// The matched-uv + status overloads (descriptor_matcher.h:81-88,126-133) exercise FillMatchedPixelByPairIndices.
int ftkref_match_brief_nearby_uv(const uint8_t *ref_bits, int32_t n_ref, const uint8_t *cur_bits, int32_t n_cur, int32_t len, const float *pred_uv,
                                 const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist, float *matched_uv, uint8_t *status,
                                 int32_t status_count) {
    BriefMatcher m;
    ApplyMatcherOptions(m, max_drow, max_dcol, max_dist);
    std::vector<Vec2> matched;
    std::vector<uint8_t> st(status, status + status_count);
    if (!m.NearbyMatch(UnpackBrief(ref_bits, n_ref, len), UnpackBrief(cur_bits, n_cur, len), WrapUv(pred_uv, n_ref), WrapUv(cur_uv, n_cur), matched, st))
        return 0;
    for (int32_t i = 0; i < n_ref; ++i) {
        matched_uv[2 * i] = matched[i].x();
        matched_uv[2 * i + 1] = matched[i].y();
        status[i] = st[i];
    }
    return 1;
}

int ftkref_match_brief_force_uv(const uint8_t *ref_bits, int32_t n_ref, const uint8_t *cur_bits, int32_t n_cur, int32_t len, const float *cur_uv,
                                float max_dist, float *matched_uv, uint8_t *status, int32_t status_count) {
    BriefMatcher m;
    ApplyMatcherOptions(m, 0, 0, max_dist);
    std::vector<Vec2> matched;
    std::vector<uint8_t> st(status, status + status_count);
    if (!m.ForceMatch(UnpackBrief(ref_bits, n_ref, len), UnpackBrief(cur_bits, n_cur, len), WrapUv(cur_uv, n_cur), matched, st)) return 0;
    for (int32_t i = 0; i < n_ref; ++i) {
        matched_uv[2 * i] = matched[i].x();
        matched_uv[2 * i + 1] = matched[i].y();
        status[i] = st[i];
    }
    return 1;
}

// DirectMethod::TrackFeatures, camera-frame overload (direct_method_tracker.cpp:41-95).  p_c_in_ref: n x 3; q_rc = (w, x, y, z)
// and p_rc in/out; cur_uv_count / status_count = the sizes of the caller's vectors on entry.
int ftkref_direct_method_track(const ftko_direct_params *params, int32_t levels, const uint8_t *const *ref_levels, const uint8_t *const *cur_levels,
                               const int32_t *rows, const int32_t *cols, const float *K, int32_t n, const float *p_c_in_ref, const float *ref_uv,
                               float *cur_uv, int32_t cur_uv_count, float *q_rc, float *p_rc, uint8_t *status, int32_t status_count) {
    feature_tracker::DirectMethod solver;
    solver.options().kMaxTrackPointsNumber = params->max_track_points;
    solver.options().kMaxIteration = params->max_iteration;
    solver.options().kPatchRowHalfSize = params->patch_row_half;
    solver.options().kPatchColHalfSize = params->patch_col_half;
    solver.options().kMaxConvergeStep = params->max_converge_step;
    solver.options().kMaxConvergeResidual = params->max_converge_residual;
    solver.options().kMethod = static_cast<feature_tracker::DirectMethodMethod>(params->method);
    PaddedLevels ref_store, cur_store;
    ref_store.Adopt(levels, ref_levels, rows, cols);
    cur_store.Adopt(levels, cur_levels, rows, cols);
    ImagePyramid ref_pyr, cur_pyr;
    ref_pyr.SetLevels(levels, ref_store.ptr.data(), rows, cols);
    cur_pyr.SetLevels(levels, cur_store.ptr.data(), rows, cols);
    std::vector<Vec2> ref_vec = WrapUv(ref_uv, n);
    std::vector<Vec2> cur_vec = WrapUv(cur_uv, cur_uv_count);
    std::vector<uint8_t> status_vec(status, status + status_count);
    std::vector<Vec3> points(n);
    for (int32_t i = 0; i < n; ++i) points[i] << p_c_in_ref[3 * i], p_c_in_ref[3 * i + 1], p_c_in_ref[3 * i + 2];
    const std::array<float, 4> Kc = {K[0], K[1], K[2], K[3]};
    Quat q(q_rc[0], q_rc[1], q_rc[2], q_rc[3]);
    Vec3 p;
    p << p_rc[0], p_rc[1], p_rc[2];
    if (!solver.TrackFeatures(ref_pyr, cur_pyr, Kc, points, ref_vec, cur_vec, q, p, status_vec)) return 0;
    for (int32_t i = 0; i < n; ++i) {
        cur_uv[2 * i] = cur_vec[i].x();
        cur_uv[2 * i + 1] = cur_vec[i].y();
        status[i] = status_vec[i];
    }
    q_rc[0] = q.w(), q_rc[1] = q.x(), q_rc[2] = q.y(), q_rc[3] = q.z();
    p_rc[0] = p(0), p_rc[1] = p(1), p_rc[2] = p(2);
    return 1;
}

// The world-frame overload (direct_method_tracker.cpp:8-39): ref pose, world points, current pose in/out.
int ftkref_direct_method_track_world(const ftko_direct_params *params, int32_t levels, const uint8_t *const *ref_levels, const uint8_t *const *cur_levels,
                                     const int32_t *rows, const int32_t *cols, const float *K, const float *ref_q_wc, const float *ref_p_wc, int32_t n,
                                     const float *p_w, const float *ref_uv, float *cur_uv, int32_t cur_uv_count, float *cur_q_wc, float *cur_p_wc,
                                     uint8_t *status, int32_t status_count) {
    feature_tracker::DirectMethod solver;
    solver.options().kMaxTrackPointsNumber = params->max_track_points;
    solver.options().kMaxIteration = params->max_iteration;
    solver.options().kPatchRowHalfSize = params->patch_row_half;
    solver.options().kPatchColHalfSize = params->patch_col_half;
    solver.options().kMaxConvergeStep = params->max_converge_step;
    solver.options().kMaxConvergeResidual = params->max_converge_residual;
    solver.options().kMethod = static_cast<feature_tracker::DirectMethodMethod>(params->method);
    PaddedLevels ref_store, cur_store;
    ref_store.Adopt(levels, ref_levels, rows, cols);
    cur_store.Adopt(levels, cur_levels, rows, cols);
    ImagePyramid ref_pyr, cur_pyr;
    ref_pyr.SetLevels(levels, ref_store.ptr.data(), rows, cols);
    cur_pyr.SetLevels(levels, cur_store.ptr.data(), rows, cols);
    std::vector<Vec2> ref_vec = WrapUv(ref_uv, n);
    std::vector<Vec2> cur_vec = WrapUv(cur_uv, cur_uv_count);
    std::vector<uint8_t> status_vec(status, status + status_count);
    std::vector<Vec3> points(n);
    for (int32_t i = 0; i < n; ++i) points[i] << p_w[3 * i], p_w[3 * i + 1], p_w[3 * i + 2];
    const std::array<float, 4> Kc = {K[0], K[1], K[2], K[3]};
    const Quat q_ref(ref_q_wc[0], ref_q_wc[1], ref_q_wc[2], ref_q_wc[3]);
    Quat q_cur(cur_q_wc[0], cur_q_wc[1], cur_q_wc[2], cur_q_wc[3]);
    Vec3 p_ref, p_cur;
    p_ref << ref_p_wc[0], ref_p_wc[1], ref_p_wc[2];
    p_cur << cur_p_wc[0], cur_p_wc[1], cur_p_wc[2];
    if (!solver.TrackFeatures(ref_pyr, cur_pyr, Kc, q_ref, p_ref, points, ref_vec, cur_vec, q_cur, p_cur, status_vec)) return 0;
    for (int32_t i = 0; i < n; ++i) {
        cur_uv[2 * i] = cur_vec[i].x();
        cur_uv[2 * i + 1] = cur_vec[i].y();
        status[i] = status_vec[i];
    }
    cur_q_wc[0] = q_cur.w(), cur_q_wc[1] = q_cur.x(), cur_q_wc[2] = q_cur.y(), cur_q_wc[3] = q_cur.z();
    cur_p_wc[0] = p_cur(0), cur_p_wc[1] = p_cur(1), cur_p_wc[2] = p_cur(2);
    return 1;
}

// DenseOpticalFlow::Track (dense_optical_flow.cpp:7-85).  single_level != 0: the GrayImage overload on level 0, with
// flow_row / flow_col as in/out (flow_valid == 0 = the caller's matrices have the wrong size, i.e. start from zero);
// otherwise the pyramid overload (the flow is an output only).  flow_* are rows[0] x cols[0] row-major.
int ftkref_dense_flow_track(const ftko_dense_flow_params *params, int32_t levels, const uint8_t *const *ref_levels, const uint8_t *const *cur_levels,
                            const int32_t *rows, const int32_t *cols, int32_t single_level, int32_t flow_valid, float *flow_row, float *flow_col) {
    feature_tracker::DenseOpticalFlow solver;
    solver.options().kMaxIteration = params->max_iteration;
    solver.options().kHalfPatchSize = params->half_patch_size;
    solver.options().kMaxConvergeStep = params->max_converge_step;
    solver.options().kMaxDeltaFlowStep = params->max_delta_flow_step;
    PaddedLevels ref_store, cur_store;
    ref_store.Adopt(levels, ref_levels, rows, cols);
    cur_store.Adopt(levels, cur_levels, rows, cols);
    std::array<Mat, 2> flow;
    const size_t n0 = static_cast<size_t>(rows[0]) * cols[0];
    bool ok;
    if (single_level) {
        if (flow_valid) {
            flow[0].resize(rows[0], cols[0]);
            flow[1].resize(rows[0], cols[0]);
            std::copy(flow_row, flow_row + n0, flow[0].v.begin());
            std::copy(flow_col, flow_col + n0, flow[1].v.begin());
        }
        GrayImage ref_image(ref_store.ptr[0], rows[0], cols[0]);
        GrayImage cur_image(cur_store.ptr[0], rows[0], cols[0]);
        ok = solver.Track(ref_image, cur_image, flow);
    } else {
        ImagePyramid ref_pyr, cur_pyr;
        ref_pyr.SetLevels(levels, ref_store.ptr.data(), rows, cols);
        cur_pyr.SetLevels(levels, cur_store.ptr.data(), rows, cols);
        ok = solver.Track(ref_pyr, cur_pyr, flow);
    }
    if (!ok) return 0;
    std::copy(flow[0].v.begin(), flow[0].v.end(), flow_row);
    std::copy(flow[1].v.begin(), flow[1].v.end(), flow_col);
    return 1;
}

// NNFeatureMatcher::Match, score-matrix branch (nn_feature_matcher.cpp:150-219), run by the reference's own code on an injected
// network output: `scores` (n_ref x n_cur row-major) is handed to the stub session of oracle/shim/onnx_run_time.h, Match() then does
// its column / row arg-max, kMinValidMatchScore gate and mutual check.  idx[i] = the current index Match() assigned to reference
// feature i (recovered from matched_pixel_uv_cur: current feature j sits at pixel (j, j)), -1 where status stayed kLargeResidual.
// n_ref <= n_cur is required (the reference sizes matched_pixel_uv_cur by pixel_uv_cur and indexes it by idx_ref, :158,215).
int ftkref_nn_match_scores(const float *scores, int32_t n_ref, int32_t n_cur, float min_score, int32_t *idx) {
    if (n_ref <= 0 || n_cur <= 0 || n_ref > n_cur) return 0;
    feature_tracker::NNFeatureMatcher matcher;
    matcher.options().kMinValidMatchScore = min_score;
    matcher.options().kMaxNumberOfMatches = 4;
    matcher.options().kModelType = feature_tracker::NNFeatureMatcher::ModelType::kLightglueForSuperpointScoreMat;
    shim::NextOutputs().clear();
    if (!matcher.Initialize()) return 0;
    Ort::Value out;
    out.rows = n_ref;
    out.cols = n_cur;
    out.f.assign(scores, scores + static_cast<size_t>(n_ref) * n_cur);
    shim::NextOutputs().clear();
    shim::NextOutputs().push_back(out);
    std::vector<SuperpointDescriptorType> desc_ref(n_ref), desc_cur(n_cur);
    std::vector<Vec2> uv_ref(n_ref), uv_cur(n_cur), matched;
    for (int32_t j = 0; j < n_cur; ++j) uv_cur[j] = Vec2(static_cast<float>(j), static_cast<float>(j));
    std::vector<uint8_t> status;
    if (!matcher.Match(desc_ref, desc_cur, uv_ref, uv_cur, matched, status)) return 0;
    for (int32_t i = 0; i < n_ref; ++i)
        idx[i] = status[i] == static_cast<uint8_t>(feature_tracker::TrackStatus::kTracked) ? static_cast<int32_t>(matched[i].x()) : -1;
    return 1;
}

// The "matches" branch of the same function (:160-178; the fused LightGlue models return index pairs + scores): matches = n x 2
// int64 (idx_ref, idx_cur).  Outputs as above.
int ftkref_nn_match_pairs(const int64_t *matches, int32_t n_matches, int32_t n_ref, int32_t n_cur, int32_t *idx) {
    if (n_ref <= 0 || n_cur <= 0 || n_ref > n_cur) return 0;
    feature_tracker::NNFeatureMatcher matcher;
    matcher.options().kMaxNumberOfMatches = 4;
    shim::NextOutputs().clear();
    if (!matcher.Initialize()) return 0;
    Ort::Value pairs, mscores;
    pairs.rows = n_matches, pairs.cols = 2;
    pairs.i64.assign(matches, matches + static_cast<size_t>(n_matches) * 2);
    mscores.rows = n_matches, mscores.cols = 1;
    mscores.f.assign(static_cast<size_t>(n_matches), 1.0f);
    shim::NextOutputs() = {pairs, mscores};
    std::vector<SuperpointDescriptorType> desc_ref(n_ref), desc_cur(n_cur);
    std::vector<Vec2> uv_ref(n_ref), uv_cur(n_cur), matched;
    for (int32_t j = 0; j < n_cur; ++j) uv_cur[j] = Vec2(static_cast<float>(j), static_cast<float>(j));
    std::vector<uint8_t> status;
    if (!matcher.Match(desc_ref, desc_cur, uv_ref, uv_cur, matched, status)) return 0;
    for (int32_t i = 0; i < n_ref; ++i)
        idx[i] = status[i] == static_cast<uint8_t>(feature_tracker::TrackStatus::kTracked) ? static_cast<int32_t>(matched[i].x()) : -1;
    return 1;
}

}  // extern "C"
