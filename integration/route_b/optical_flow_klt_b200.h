// INTEGRATION Route B -- a B200 backend INSIDE the reference's own class hierarchy.
//
// This is the file a Feature_Tracker maintainer adds to the reference tree (e.g. src/optical_flow_tracker/b200/).  It is
// compiled against the REFERENCE's headers (optical_flow.h and the three subclass headers, Slam_Utility's basic_type.h /
// datatype_image*.h) plus include/ftk_c.h, and linked with libftk_b200.so.  The seam is the pair of private pure virtuals
// OpticalFlow::TrackMultipleLevel / TrackSingleLevel (src/optical_flow_tracker/optical_flow.h:83-86): the subclasses below
// override them, so OpticalFlow::TrackFeatures (optical_flow.cpp:6-47) -- input normalisation, PrepareForTracking, dispatch --
// and every public accessor (options(), predict_affine(), predict_R_cr(), consider_patch_luminance()) stay the reference's.
//
//   feature_tracker::OpticalFlowBasicKltB200 klt;      // instead of OpticalFlowBasicKlt
//   klt.options().kMethod = ...;                       // unchanged application code from here on
//   klt.TrackFeatures(ref_pyramid, cur_pyramid, ref_pixel_uv, cur_pixel_uv, status);
//
// The pyramids are the application's own (Slam_Utility ImagePyramid): every level is uploaded verbatim, so the result does not
// depend on how the application built them.  tests/test_route_b.py builds this header against the reference headers where
// they lie and checks the three subclasses, multi- and single-level, bit for bit against the reference's CPU classes.
#ifndef FEATURE_TRACKER_OPTICAL_FLOW_KLT_B200_H_
#define FEATURE_TRACKER_OPTICAL_FLOW_KLT_B200_H_

#include <cstring>
#include <vector>

#include "ftk_c.h"
#include "optical_flow_affine_klt.h"
#include "optical_flow_basic_klt.h"
#include "optical_flow_lssd_klt.h"

namespace feature_tracker {

namespace b200_detail {

inline void FillExtras(const OpticalFlowBasicKlt &, ftk_klt_params &p) { p.variant = FTK_VARIANT_BASIC; }
inline void FillExtras(const OpticalFlowAffineKlt &t, ftk_klt_params &p) {
    p.variant = FTK_VARIANT_AFFINE;
    const Mat2 &m = t.predict_affine();
    p.predict[0] = m(0, 0), p.predict[1] = m(0, 1), p.predict[2] = m(1, 0), p.predict[3] = m(1, 1);
}
inline void FillExtras(const OpticalFlowLssdKlt &t, ftk_klt_params &p) {
    p.variant = FTK_VARIANT_LSSD;
    const Mat2 &m = t.predict_R_cr();
    p.predict[0] = m(0, 0), p.predict[1] = m(0, 1), p.predict[2] = m(1, 0), p.predict[3] = m(1, 1);
    p.consider_patch_luminance = t.consider_patch_luminance() ? 1 : 0;
}

// One GPU context + two device pyramids (ref, cur), re-created when the image size or level count changes.
class DeviceSide {
public:
    DeviceSide() = default;
    ~DeviceSide() {
        Release();
        if (ctx_) ftk_destroy(ctx_);
    }
    DeviceSide(const DeviceSide &) = delete;
    DeviceSide &operator=(const DeviceSide &) = delete;

    ftk_context *ctx() {
        if (!ctx_ && ftk_create(0, &ctx_) != FTK_OK) ctx_ = nullptr;  // no B200 => every TrackFeatures returns false (there is no CPU fallback here)
        return ctx_;
    }
    bool Ensure(int32_t rows, int32_t cols, int32_t levels) {
        if (!ctx()) return false;
        if (ref_ && rows == rows_ && cols == cols_ && levels == levels_) return true;
        Release();
        if (ftk_pyramid_create(ctx_, rows, cols, levels, 1, &ref_) != FTK_OK || ftk_pyramid_create(ctx_, rows, cols, levels, 1, &cur_) != FTK_OK) {
            Release();
            return false;
        }
        rows_ = rows, cols_ = cols, levels_ = levels;
        return true;
    }
    ftk_pyramid *ref() { return ref_; }
    ftk_pyramid *cur() { return cur_; }

private:
    void Release() {
        if (ref_) ftk_pyramid_destroy(ctx_, ref_);
        if (cur_) ftk_pyramid_destroy(ctx_, cur_);
        ref_ = cur_ = nullptr;
        rows_ = cols_ = levels_ = 0;
    }
    ftk_context *ctx_ = nullptr;
    ftk_pyramid *ref_ = nullptr, *cur_ = nullptr;
    int32_t rows_ = 0, cols_ = 0, levels_ = 0;
};

}  // namespace b200_detail

template <typename Base>
class OpticalFlowKltB200 : public Base {
public:
    OpticalFlowKltB200() : Base() {}
    virtual ~OpticalFlowKltB200() = default;
    virtual std::string OpticalFlowMethodName() const override { return Base::OpticalFlowMethodName() + " (B200)"; }

private:
    // optical_flow.h:83-84.  Called by OpticalFlow::TrackFeatures after it normalised cur_pixel_uv / status (optical_flow.cpp:12-19).
    virtual bool TrackMultipleLevel(const ImagePyramid &ref_pyramid, const ImagePyramid &cur_pyramid, const std::vector<Vec2> &ref_pixel_uv,
                                    std::vector<Vec2> &cur_pixel_uv, std::vector<uint8_t> &status) override {
        const int32_t levels = static_cast<int32_t>(ref_pyramid.level());
        const GrayImage &r0 = ref_pyramid.GetImageConst(0), &c0 = cur_pyramid.GetImageConst(0);
        if (levels < 1 || r0.rows() != c0.rows() || r0.cols() != c0.cols()) return false;
        if (!gpu_.Ensure(r0.rows(), r0.cols(), levels)) return false;
        for (int32_t l = 0; l < levels; ++l) {
            if (ftk_pyramid_set_level(gpu_.ctx(), gpu_.ref(), 0, l, ref_pyramid.GetImageConst(l).data()) != FTK_OK) return false;
            if (ftk_pyramid_set_level(gpu_.ctx(), gpu_.cur(), 0, l, cur_pyramid.GetImageConst(l).data()) != FTK_OK) return false;
        }
        return Run(ref_pixel_uv, cur_pixel_uv, status, 0u);
    }
    // optical_flow.h:85-86
    virtual bool TrackSingleLevel(const GrayImage &ref_image, const GrayImage &cur_image, const std::vector<Vec2> &ref_pixel_uv,
                                  std::vector<Vec2> &cur_pixel_uv, std::vector<uint8_t> &status) override {
        if (ref_image.rows() != cur_image.rows() || ref_image.cols() != cur_image.cols()) return false;
        if (!gpu_.Ensure(ref_image.rows(), ref_image.cols(), 1)) return false;
        if (ftk_pyramid_set_level(gpu_.ctx(), gpu_.ref(), 0, 0, ref_image.data()) != FTK_OK) return false;
        if (ftk_pyramid_set_level(gpu_.ctx(), gpu_.cur(), 0, 0, cur_image.data()) != FTK_OK) return false;
        return Run(ref_pixel_uv, cur_pixel_uv, status, FTK_FLAG_SINGLE_LEVEL);
    }

    bool Run(const std::vector<Vec2> &ref_pixel_uv, std::vector<Vec2> &cur_pixel_uv, std::vector<uint8_t> &status, uint32_t flags) {
        ftk_klt_params p;
        ftk_klt_params_default(&p);
        const OpticalFlowOptions &o = this->options();
        p.method = static_cast<int32_t>(o.kMethod);
        p.max_track_points = o.kMaxTrackPointsNumber;
        p.max_iteration = o.kMaxIteration;
        p.max_tolerance_large_step = o.kMaxToleranceLargeStep;
        p.patch_row_half = o.kPatchRowHalfSize;
        p.patch_col_half = o.kPatchColHalfSize;
        p.max_converge_step = o.kMaxConvergeStep;
        b200_detail::FillExtras(static_cast<const Base &>(*this), p);
        const int32_t n = static_cast<int32_t>(ref_pixel_uv.size());
        // Vec2 -> interleaved (x, y) floats.  With Eigen::Vector2f the vectors already have that layout and the two copies can be
        // replaced by ref_pixel_uv[0].data() / cur_pixel_uv[0].data(); the copy keeps this file independent of Vec2's layout.
        ref_flat_.resize(2 * static_cast<size_t>(n));
        cur_flat_.resize(2 * static_cast<size_t>(n));
        for (int32_t i = 0; i < n; ++i) {
            ref_flat_[2 * i] = ref_pixel_uv[i].x(), ref_flat_[2 * i + 1] = ref_pixel_uv[i].y();
            cur_flat_[2 * i] = cur_pixel_uv[i].x(), cur_flat_[2 * i + 1] = cur_pixel_uv[i].y();
        }
        const int32_t offsets[2] = {0, n}, image0 = 0;
        if (ftk_klt_track(gpu_.ctx(), &p, gpu_.ref(), gpu_.cur(), 1, &image0, &image0, offsets, ref_flat_.data(), cur_flat_.data(), status.data(), flags) != FTK_OK)
            return false;
        for (int32_t i = 0; i < n; ++i) cur_pixel_uv[i].x() = cur_flat_[2 * i], cur_pixel_uv[i].y() = cur_flat_[2 * i + 1];
        return true;
    }

    b200_detail::DeviceSide gpu_;
    std::vector<float> ref_flat_, cur_flat_;
};

using OpticalFlowBasicKltB200 = OpticalFlowKltB200<OpticalFlowBasicKlt>;
using OpticalFlowAffineKltB200 = OpticalFlowKltB200<OpticalFlowAffineKlt>;
using OpticalFlowLssdKltB200 = OpticalFlowKltB200<OpticalFlowLssdKlt>;

}  // namespace feature_tracker

#endif
