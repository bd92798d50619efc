// C++ facade over the C ABI (include/ftk_c.h) with the reference's class names, option structs and method signatures
// (namespace feature_tracker; src/feature_tracker.h, src/optical_flow_tracker/optical_flow.h and the three KLT subclasses,
// src/descriptor_matcher/descriptor_matcher.h), so application code that holds `OpticalFlowBasicKlt klt;` or a
// `DescriptorMatcher<T>` subclass recompiles against this header and runs on the B200 instead of the CPU loops.
//
// Differences a maintainer has to know about (all dictated by where the data lives):
//   * ImagePyramid is device resident: SetRawImage() + CreateImagePyramid() upload level 0 and build the levels on the GPU.
//   * DescriptorMatcher<T> keeps the private virtual ComputeDistance (descriptor_matcher.h:45), so the reference's subclasses
//     (test/test_descriptor_matcher_brief.cpp:27-46) compile unchanged, but it is NOT called per pair (one virtual call per pair
//     cannot feed a GPU): the distance the GPU evaluates is selected by DescriptorTraits<T> -- Hamming for element-wise boolean
//     containers (BriefType), 0.5 - 0.5 * cos for float vectors (Superpoint / Disk descriptors), exactly the demos' bodies.  Every
//     match call evaluates the subclass's override on a few (ref, cur) pairs and FAILS LOUDLY (std::logic_error) when it
//     disagrees with that metric, so a custom distance can never be silently replaced.
//   * Vec2 is any type with x() / y() accessors and a (float, float) constructor (Eigen::Vector2f qualifies); define
//     FTK_VEC2_TYPE before including this header to use the application's own type.
// Header only; link with libftk_b200.so.
#ifndef FEATURE_TRACKER_B200_FEATURE_TRACKER_H_
#define FEATURE_TRACKER_B200_FEATURE_TRACKER_H_

#include <array>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "ftk_c.h"

#ifndef FTK_VEC2_TYPE
namespace feature_tracker {
struct Vec2f {
    float v[2];
    Vec2f() : v{0.0f, 0.0f} {}
    Vec2f(float x, float y) : v{x, y} {}
    float &x() { return v[0]; }
    float &y() { return v[1]; }
    const float &x() const { return v[0]; }
    const float &y() const { return v[1]; }
};
}  // namespace feature_tracker
#define FTK_VEC2_TYPE ::feature_tracker::Vec2f
#endif

namespace feature_tracker {

using Vec2 = FTK_VEC2_TYPE;

// src/feature_tracker.h:8-14
enum class TrackStatus : uint8_t {
    kNotTracked = 0,
    kTracked = 1,
    kLargeResidual = 2,
    kOutside = 3,
    kNumericError = 4,
};

// One GPU context shared by the facade objects of a thread (the reference objects are not re-entrant either).
class Device {
public:
    static ftk_context *Get(int device = 0) {
        static thread_local std::unique_ptr<Device> instance;
        if (!instance || instance->device_ != device) instance.reset(new Device(device));
        return instance->ctx_;
    }
    ~Device() { ftk_destroy(ctx_); }

private:
    explicit Device(int device) : device_(device) {
        if (ftk_create(device, &ctx_) != FTK_OK) throw std::runtime_error("feature_tracker_b200: no usable B200 (sm_100) device; there is no CPU fallback");
    }
    int device_ = 0;
    ftk_context *ctx_ = nullptr;
};

// Host image view with the accessors of Slam_Utility's GrayImage the reference's call sites use (data / rows / cols;
// test/test_optical_flow.cpp:45-52).  The GrayImage overloads below are templates over any type with these three accessors,
// so an application passes its own Slam_Utility GrayImage directly; this class only serves applications that have none.
class GrayImage {
public:
    GrayImage() = default;
    GrayImage(uint8_t *data, int32_t rows, int32_t cols) : data_(data), rows_(rows), cols_(cols) {}
    void SetImage(uint8_t *data, int32_t rows, int32_t cols) { data_ = data, rows_ = rows, cols_ = cols; }
    uint8_t *data() const { return data_; }
    int32_t rows() const { return rows_; }
    int32_t cols() const { return cols_; }

private:
    uint8_t *data_ = nullptr;
    int32_t rows_ = 0, cols_ = 0;
};

// Device-resident stand-in for Slam_Utility's ImagePyramid (call sites: test/test_optical_flow.cpp:49-53,70-71).
class ImagePyramid {
public:
    ImagePyramid() = default;
    ~ImagePyramid() { Release(); }
    ImagePyramid(const ImagePyramid &) = delete;
    ImagePyramid &operator=(const ImagePyramid &) = delete;

    void SetPyramidBuff(uint8_t *, bool) {}  // levels live in HBM; kept for source compatibility
    void SetRawImage(const uint8_t *data, int32_t rows, int32_t cols) {
        raw_ = data;
        rows_ = rows;
        cols_ = cols;
    }
    bool CreateImagePyramid(uint32_t level) {
        if (raw_ == nullptr || level == 0) return false;
        ftk_context *ctx = Device::Get();
        if (pyr_ == nullptr || ftk_pyramid_levels(pyr_) != static_cast<int32_t>(level) || built_rows_ != rows_ || built_cols_ != cols_) {
            Release();
            if (ftk_pyramid_create(ctx, rows_, cols_, static_cast<int32_t>(level), 1, &pyr_) != FTK_OK) return false;
            built_rows_ = rows_;
            built_cols_ = cols_;
        }
        for (auto &h : host_levels_) h.clear();
        if (ftk_pyramid_set_images(ctx, pyr_, 0, 1, raw_, 0) != FTK_OK) return false;
        return ftk_pyramid_build(ctx, pyr_, 0, 1) == FTK_OK && ftk_synchronize(ctx) == FTK_OK;
    }
    uint32_t level() const { return pyr_ ? static_cast<uint32_t>(ftk_pyramid_levels(pyr_)) : 0u; }
    int32_t rows() const { return rows_; }
    int32_t cols() const { return cols_; }
    const ftk_pyramid *handle() const { return pyr_; }
    // ImagePyramid::GetImageConst(i) of Slam_Utility (callers: basic_klt.cpp:22-23): level i as a host image.  The levels live in
    // HBM; level i is copied back on first use after each CreateImagePyramid and cached (level 0 is the caller's own buffer).
    const GrayImage &GetImageConst(uint32_t i) const {
        if (i >= kPyramidMaxLevel || pyr_ == nullptr || i >= level()) throw std::out_of_range("ImagePyramid::GetImageConst: no such level");
        if (i == 0) {
            views_[0].SetImage(const_cast<uint8_t *>(raw_), rows_, cols_);
        } else if (host_levels_[i].empty()) {
            const int32_t r = rows_ >> i, c = cols_ >> i;
            host_levels_[i].resize(static_cast<size_t>(r) * c);
            if (ftk_pyramid_get_level(Device::Get(), pyr_, 0, static_cast<int32_t>(i), host_levels_[i].data()) != FTK_OK)
                throw std::runtime_error("ImagePyramid::GetImageConst: device read failed");
            views_[i].SetImage(host_levels_[i].data(), r, c);
        }
        return views_[i];
    }
    static constexpr uint32_t kPyramidMaxLevel = 10;

private:
    void Release() {
        if (pyr_) ftk_pyramid_destroy(Device::Get(), pyr_);
        pyr_ = nullptr;
    }
    mutable std::array<std::vector<uint8_t>, kPyramidMaxLevel> host_levels_;
    mutable std::array<GrayImage, kPyramidMaxLevel> views_;
    const uint8_t *raw_ = nullptr;
    int32_t rows_ = 0, cols_ = 0, built_rows_ = 0, built_cols_ = 0;
    ftk_pyramid *pyr_ = nullptr;
};

// A batch of same-sized frames, device resident (not in the reference, whose callers loop over frame pairs: the batch entry points
// below -- OpticalFlow::TrackFeaturesBatch, DescriptorMatcher<T>::MatchPairs -- are how the GPU is kept busy).
class ImagePyramidBatch {
public:
    ImagePyramidBatch() = default;
    ~ImagePyramidBatch() { Release(); }
    ImagePyramidBatch(const ImagePyramidBatch &) = delete;
    ImagePyramidBatch &operator=(const ImagePyramidBatch &) = delete;

    // `images`: n_images tightly packed rows x cols frames in host memory (pinned memory makes the upload asynchronous).
    bool CreateImagePyramids(const uint8_t *images, int32_t rows, int32_t cols, int32_t n_images, uint32_t level) {
        if (images == nullptr || level == 0 || n_images <= 0) return false;
        ftk_context *ctx = Device::Get();
        if (pyr_ == nullptr || ftk_pyramid_levels(pyr_) != static_cast<int32_t>(level) || rows != rows_ || cols != cols_ || ftk_pyramid_images(pyr_) != n_images) {
            Release();
            if (ftk_pyramid_create(ctx, rows, cols, static_cast<int32_t>(level), n_images, &pyr_) != FTK_OK) return false;
            rows_ = rows, cols_ = cols;
        }
        if (ftk_pyramid_set_images(ctx, pyr_, 0, n_images, images, 0) != FTK_OK) return false;
        return ftk_pyramid_build(ctx, pyr_, 0, n_images) == FTK_OK && ftk_synchronize(ctx) == FTK_OK;
    }
    uint32_t level() const { return pyr_ ? static_cast<uint32_t>(ftk_pyramid_levels(pyr_)) : 0u; }
    int32_t images() const { return pyr_ ? ftk_pyramid_images(pyr_) : 0; }
    const ftk_pyramid *handle() const { return pyr_; }

private:
    void Release() {
        if (pyr_) ftk_pyramid_destroy(Device::Get(), pyr_);
        pyr_ = nullptr;
    }
    int32_t rows_ = 0, cols_ = 0;
    ftk_pyramid *pyr_ = nullptr;
};

// src/optical_flow_tracker/optical_flow.h:12-18
enum class OpticalFlowMethod : uint8_t {
    kInverse = 0,
    kDirect = 1,
    kFast = 2,
    kSse = 3,
    kNeon = 4,
};

// src/optical_flow_tracker/optical_flow.h:20-28
struct OpticalFlowOptions {
    uint32_t kMaxTrackPointsNumber = 500;
    uint32_t kMaxIteration = 15;
    uint32_t kMaxToleranceLargeStep = 3;
    int32_t kPatchRowHalfSize = 6;
    int32_t kPatchColHalfSize = 6;
    float kMaxConvergeStep = 4e-2f;
    OpticalFlowMethod kMethod = OpticalFlowMethod::kFast;
};

// src/optical_flow_tracker/optical_flow.h:30-104
class OpticalFlow {
public:
    OpticalFlow() = default;
    virtual ~OpticalFlow() = default;
    virtual std::string OpticalFlowMethodName() const { return "None"; }

    // optical_flow.cpp:6-26
    bool TrackFeatures(const ImagePyramid &ref_pyramid, const ImagePyramid &cur_pyramid, const std::vector<Vec2> &ref_pixel_uv,
                       std::vector<Vec2> &cur_pixel_uv, std::vector<uint8_t> &status) {
        if (ref_pixel_uv.empty()) return false;
        if (cur_pyramid.level() != ref_pyramid.level()) return false;
        return Track(ref_pyramid, cur_pyramid, ref_pixel_uv, cur_pixel_uv, status, 0u);
    }
    // optical_flow.cpp:28-47, optical_flow.h:41-42: TrackFeatures(const GrayImage &, const GrayImage &, ...) -- TrackSingleLevel on
    // two host images.  A template over the image type (data() / rows() / cols()), so Slam_Utility's GrayImage is accepted as is.
    template <typename GrayImageT, typename = decltype(std::declval<const GrayImageT &>().data()), typename = decltype(std::declval<const GrayImageT &>().rows())>
    bool TrackFeatures(const GrayImageT &ref_image, const GrayImageT &cur_image, const std::vector<Vec2> &ref_pixel_uv, std::vector<Vec2> &cur_pixel_uv,
                       std::vector<uint8_t> &status) {
        if (ref_pixel_uv.empty()) return false;                                       // :30
        if (ref_image.data() == nullptr || cur_image.data() == nullptr) return false;  // :31-32
        if (ref_image.rows() != cur_image.rows() || ref_image.cols() != cur_image.cols()) return false;  // one device batch holds equal-sized images
        single_ref_.SetRawImage(ref_image.data(), ref_image.rows(), ref_image.cols());
        single_cur_.SetRawImage(cur_image.data(), cur_image.rows(), cur_image.cols());
        if (!single_ref_.CreateImagePyramid(1) || !single_cur_.CreateImagePyramid(1)) return false;
        return Track(single_ref_, single_cur_, ref_pixel_uv, cur_pixel_uv, status, FTK_FLAG_SINGLE_LEVEL);
    }
    // The same on level 0 of two device pyramids (no upload).
    bool TrackFeaturesSingleLevel(const ImagePyramid &ref_image, const ImagePyramid &cur_image, const std::vector<Vec2> &ref_pixel_uv,
                                  std::vector<Vec2> &cur_pixel_uv, std::vector<uint8_t> &status) {
        if (ref_pixel_uv.empty()) return false;
        return Track(ref_image, cur_image, ref_pixel_uv, cur_pixel_uv, status, FTK_FLAG_SINGLE_LEVEL);
    }

    // Many frame pairs in one call (ftk_klt_track): pair p tracks features [feat_offsets[p], feat_offsets[p + 1]) from frame
    // ref_image[p] to frame cur_image[p] of the batch, each with exactly the semantics of TrackFeatures above (kMaxTrackPointsNumber
    // applies per pair).  Equal to looping over the pairs; one launch instead of one call per pair.
    bool TrackFeaturesBatch(const ImagePyramidBatch &pyramids, const std::vector<int32_t> &ref_image, const std::vector<int32_t> &cur_image,
                            const std::vector<int32_t> &feat_offsets, const std::vector<Vec2> &ref_pixel_uv, std::vector<Vec2> &cur_pixel_uv,
                            std::vector<uint8_t> &status) {
        if (ref_pixel_uv.empty() || !pyramids.handle()) return false;
        const size_t n_pairs = ref_image.size();
        if (n_pairs == 0 || cur_image.size() != n_pairs || feat_offsets.size() != n_pairs + 1 || static_cast<size_t>(feat_offsets.back()) != ref_pixel_uv.size()) return false;
        ftk_klt_params p = Params();
        uint32_t flags = 0;
        const size_t n = ref_pixel_uv.size();
        std::vector<float> ref_flat(2 * n), cur_flat(2 * n, 0.0f);
        for (size_t i = 0; i < n; ++i) ref_flat[2 * i] = ref_pixel_uv[i].x(), ref_flat[2 * i + 1] = ref_pixel_uv[i].y();
        if (cur_pixel_uv.size() == n) {
            for (size_t i = 0; i < n; ++i) cur_flat[2 * i] = cur_pixel_uv[i].x(), cur_flat[2 * i + 1] = cur_pixel_uv[i].y();
        } else {
            flags |= FTK_FLAG_NO_PREDICTION;
        }
        if (status.size() != n) {
            flags |= FTK_FLAG_NO_STATUS;
            status.assign(n, static_cast<uint8_t>(TrackStatus::kNotTracked));
        }
        if (ftk_klt_track(Device::Get(), &p, pyramids.handle(), pyramids.handle(), static_cast<int32_t>(n_pairs), ref_image.data(), cur_image.data(), feat_offsets.data(),
                          ref_flat.data(), cur_flat.data(), status.data(), flags) != FTK_OK)
            return false;
        cur_pixel_uv.resize(n);
        for (size_t i = 0; i < n; ++i) cur_pixel_uv[i] = Vec2(cur_flat[2 * i], cur_flat[2 * i + 1]);
        return true;
    }

    OpticalFlowOptions &options() { return options_; }
    const OpticalFlowOptions &options() const { return options_; }

    // Not in the reference: > 0 adds the forward-backward consistency pass of ftk_klt_params::forward_backward_max_error.
    float &forward_backward_max_error() { return forward_backward_max_error_; }
    const float &forward_backward_max_error() const { return forward_backward_max_error_; }

protected:
    virtual void FillParams(ftk_klt_params &p) const = 0;

private:
    ftk_klt_params Params() const {
        ftk_klt_params p;
        ftk_klt_params_default(&p);
        p.max_track_points = options_.kMaxTrackPointsNumber;
        p.max_iteration = options_.kMaxIteration;
        p.max_tolerance_large_step = options_.kMaxToleranceLargeStep;
        p.patch_row_half = options_.kPatchRowHalfSize;
        p.patch_col_half = options_.kPatchColHalfSize;
        p.max_converge_step = options_.kMaxConvergeStep;
        p.method = static_cast<int32_t>(options_.kMethod);
        p.forward_backward_max_error = forward_backward_max_error_;
        FillParams(p);
        return p;
    }
    bool Track(const ImagePyramid &ref, const ImagePyramid &cur, const std::vector<Vec2> &ref_uv, std::vector<Vec2> &cur_uv, std::vector<uint8_t> &status,
               uint32_t flags) {
        if (!ref.handle() || !cur.handle()) return false;
        const int32_t n = static_cast<int32_t>(ref_uv.size());
        ftk_klt_params p = Params();
        std::vector<float> ref_flat(2 * static_cast<size_t>(n)), cur_flat(2 * static_cast<size_t>(n), 0.0f);
        for (int32_t i = 0; i < n; ++i) ref_flat[2 * i] = ref_uv[i].x(), ref_flat[2 * i + 1] = ref_uv[i].y();
        if (cur_uv.size() == ref_uv.size()) {
            for (int32_t i = 0; i < n; ++i) cur_flat[2 * i] = cur_uv[i].x(), cur_flat[2 * i + 1] = cur_uv[i].y();
        } else {
            flags |= FTK_FLAG_NO_PREDICTION;  // optical_flow.cpp:12-14
        }
        if (status.size() != ref_uv.size()) {
            flags |= FTK_FLAG_NO_STATUS;  // optical_flow.cpp:17-19
            status.assign(ref_uv.size(), static_cast<uint8_t>(TrackStatus::kNotTracked));
        }
        const int32_t offsets[2] = {0, n};
        const int32_t image0 = 0;
        const int rc = ftk_klt_track(Device::Get(), &p, ref.handle(), cur.handle(), 1, &image0, &image0, offsets, ref_flat.data(), cur_flat.data(), status.data(), flags);
        if (rc != FTK_OK) return false;
        cur_uv.resize(ref_uv.size());
        for (int32_t i = 0; i < n; ++i) cur_uv[i] = Vec2(cur_flat[2 * i], cur_flat[2 * i + 1]);
        return true;
    }
    OpticalFlowOptions options_;
    float forward_backward_max_error_ = 0.0f;
    ImagePyramid single_ref_, single_cur_;  // 1-level device images of the GrayImage overload
};

// basic_klt/optical_flow_basic_klt.h:9-41
class OpticalFlowBasicKlt : public OpticalFlow {
public:
    std::string OpticalFlowMethodName() const override { return "Basic-Klt"; }

protected:
    void FillParams(ftk_klt_params &p) const override { p.variant = FTK_VARIANT_BASIC; }
};

// Row-major 2x2 float matrix with Eigen-like (i, j) access for the prediction accessors.
struct Mat2f {
    float m[4] = {1.0f, 0.0f, 0.0f, 1.0f};
    float &operator()(int i, int j) { return m[2 * i + j]; }
    const float &operator()(int i, int j) const { return m[2 * i + j]; }
};

// affine_klt/optical_flow_affine_klt.h:9-52
class OpticalFlowAffineKlt : public OpticalFlow {
public:
    std::string OpticalFlowMethodName() const override { return "Affine-Klt"; }
    Mat2f &predict_affine() { return predict_affine_; }
    const Mat2f &predict_affine() const { return predict_affine_; }

protected:
    void FillParams(ftk_klt_params &p) const override {
        p.variant = FTK_VARIANT_AFFINE;
        std::memcpy(p.predict, predict_affine_.m, sizeof(p.predict));
    }

private:
    Mat2f predict_affine_;
};

// lssd_klt/optical_flow_lssd_klt.h:9-56
class OpticalFlowLssdKlt : public OpticalFlow {
public:
    std::string OpticalFlowMethodName() const override { return "Lssd-Klt"; }
    Mat2f &predict_R_cr() { return predict_R_cr_; }
    const Mat2f &predict_R_cr() const { return predict_R_cr_; }
    bool &consider_patch_luminance() { return consider_patch_luminance_; }
    const bool &consider_patch_luminance() const { return consider_patch_luminance_; }

protected:
    void FillParams(ftk_klt_params &p) const override {
        p.variant = FTK_VARIANT_LSSD;
        std::memcpy(p.predict, predict_R_cr_.m, sizeof(p.predict));
        p.consider_patch_luminance = consider_patch_luminance_ ? 1 : 0;
    }

private:
    Mat2f predict_R_cr_;
    bool consider_patch_luminance_ = false;
};

// ---- descriptor matching (src/descriptor_matcher/descriptor_matcher.h) --------------------------------------------

// How a descriptor type is flattened for the GPU.  kBinary: element-wise boolean container -> packed little-endian
// bits, Hamming distance (test/test_descriptor_matcher_brief.cpp:33-45).  Otherwise: float vector, 0.5 - 0.5 * cos
// (test/test_descriptor_matcher_superpoint.cpp:32-34).
template <typename DescriptorType, typename Enable = void>
struct DescriptorTraits {
    static constexpr bool kBinary = false;
    static size_t Size(const DescriptorType &d) { return static_cast<size_t>(d.size()); }
    static float At(const DescriptorType &d, size_t k) { return static_cast<float>(d[k]); }
};
template <typename DescriptorType>
struct DescriptorTraits<DescriptorType, typename std::enable_if<std::is_integral<typename DescriptorType::value_type>::value>::type> {
    static constexpr bool kBinary = true;
    static size_t Size(const DescriptorType &d) { return static_cast<size_t>(d.size()); }
    static bool At(const DescriptorType &d, size_t k) { return d[k] != 0; }
};

template <typename DescriptorType>
class DescriptorMatcher {
public:
    // descriptor_matcher.h:16-20
    struct Options {
        int32_t kMaxValidPredictRowDistance = 40;
        int32_t kMaxValidPredictColDistance = 40;
        float kMaxValidDescriptorDistance = 0.0f;
    };

    DescriptorMatcher() = default;
    virtual ~DescriptorMatcher() = default;

    // descriptor_matcher.h:55-79
    bool ForceMatch(const std::vector<DescriptorType> &descriptors_ref, const std::vector<DescriptorType> &descriptors_cur,
                    std::vector<int32_t> &index_pairs_in_cur) {
        if (descriptors_cur.empty()) return false;
        uint32_t flags = PrepareIndex(descriptors_ref.size(), index_pairs_in_cur);
        return Run(descriptors_ref, descriptors_cur, nullptr, nullptr, index_pairs_in_cur, flags);
    }
    // descriptor_matcher.h:81-88
    bool ForceMatch(const std::vector<DescriptorType> &descriptors_ref, const std::vector<DescriptorType> &descriptors_cur,
                    const std::vector<Vec2> &pixel_uv_cur, std::vector<Vec2> &matched_pixel_uv_cur, std::vector<uint8_t> &status) {
        std::vector<int32_t> index_pairs_in_cur;
        if (!ForceMatch(descriptors_ref, descriptors_cur, index_pairs_in_cur)) return false;
        return FillMatchedPixelByPairIndices(index_pairs_in_cur, pixel_uv_cur, matched_pixel_uv_cur, status);
    }
    // descriptor_matcher.h:90-124
    bool NearbyMatch(const std::vector<DescriptorType> &descriptors_ref, const std::vector<DescriptorType> &descriptors_cur,
                     const std::vector<Vec2> &pixel_uv_pred_in_cur, const std::vector<Vec2> &pixel_uv_cur, std::vector<int32_t> &index_pairs_in_cur) {
        if (descriptors_cur.empty()) return false;
        if (descriptors_ref.size() != pixel_uv_pred_in_cur.size()) return false;
        if (descriptors_cur.size() != pixel_uv_cur.size()) return false;
        uint32_t flags = PrepareIndex(descriptors_ref.size(), index_pairs_in_cur);
        return Run(descriptors_ref, descriptors_cur, &pixel_uv_pred_in_cur, &pixel_uv_cur, index_pairs_in_cur, flags);
    }
    // descriptor_matcher.h:126-133
    bool NearbyMatch(const std::vector<DescriptorType> &descriptors_ref, const std::vector<DescriptorType> &descriptors_cur,
                     const std::vector<Vec2> &pixel_uv_pred_in_cur, const std::vector<Vec2> &pixel_uv_cur, std::vector<Vec2> &matched_pixel_uv_cur,
                     std::vector<uint8_t> &status) {
        std::vector<int32_t> index_pairs_in_cur;
        if (!NearbyMatch(descriptors_ref, descriptors_cur, pixel_uv_pred_in_cur, pixel_uv_cur, index_pairs_in_cur)) return false;
        return FillMatchedPixelByPairIndices(index_pairs_in_cur, pixel_uv_cur, matched_pixel_uv_cur, status);
    }

    // Many independent ForceMatch (pixel_uv_* == nullptr) or NearbyMatch problems in one call (ftk_match_hamming_pairs /
    // ftk_match_cosine_pairs): pair p matches descriptors_ref[ref_offsets[p] .. ref_offsets[p + 1]) against
    // descriptors_cur[cur_offsets[p] .. cur_offsets[p + 1]); index_pairs_in_cur[i] is the index INSIDE the pair's current block.
    bool MatchPairs(const std::vector<DescriptorType> &descriptors_ref, const std::vector<int32_t> &ref_offsets, const std::vector<DescriptorType> &descriptors_cur,
                    const std::vector<int32_t> &cur_offsets, const std::vector<Vec2> *pixel_uv_pred_in_cur, const std::vector<Vec2> *pixel_uv_cur,
                    std::vector<int32_t> &index_pairs_in_cur) {
        using Traits = DescriptorTraits<DescriptorType>;
        if (descriptors_cur.empty() || ref_offsets.size() < 2 || ref_offsets.size() != cur_offsets.size()) return false;
        if (static_cast<size_t>(ref_offsets.back()) != descriptors_ref.size() || static_cast<size_t>(cur_offsets.back()) != descriptors_cur.size()) return false;
        if ((pixel_uv_pred_in_cur == nullptr) != (pixel_uv_cur == nullptr)) return false;
        if (pixel_uv_pred_in_cur && (pixel_uv_pred_in_cur->size() != descriptors_ref.size() || pixel_uv_cur->size() != descriptors_cur.size())) return false;
        const size_t len = Traits::Size(descriptors_cur[0]);
        for (const auto &d : descriptors_ref)
            if (Traits::Size(d) != len) return false;
        for (const auto &d : descriptors_cur)
            if (Traits::Size(d) != len) return false;
        if (len == 0) return false;
        CheckOverride(descriptors_ref, descriptors_cur);
        const uint32_t flags = PrepareIndex(descriptors_ref.size(), index_pairs_in_cur);
        std::vector<float> pred_flat, pos_flat;
        if (pixel_uv_pred_in_cur) pred_flat = Flatten(*pixel_uv_pred_in_cur), pos_flat = Flatten(*pixel_uv_cur);
        const float *pred = pixel_uv_pred_in_cur ? pred_flat.data() : nullptr, *pos = pixel_uv_pred_in_cur ? pos_flat.data() : nullptr;
        const int32_t n_pairs = static_cast<int32_t>(ref_offsets.size()) - 1;
        ftk_context *ctx = Device::Get();
        if constexpr (Traits::kBinary) {
            const int32_t words = static_cast<int32_t>((len + 31) / 32);
            const std::vector<uint32_t> r = PackBits(descriptors_ref, len, words), c = PackBits(descriptors_cur, len, words);
            return ftk_match_hamming_pairs(ctx, r.data(), c.data(), words, n_pairs, ref_offsets.data(), cur_offsets.data(), pred, pos, options_.kMaxValidPredictRowDistance,
                                           options_.kMaxValidPredictColDistance, options_.kMaxValidDescriptorDistance, index_pairs_in_cur.data(), flags) == FTK_OK;
        } else {
            const int32_t dim = static_cast<int32_t>(len);
            const std::vector<float> r = PackFloats(descriptors_ref, dim), c = PackFloats(descriptors_cur, dim);
            return ftk_match_cosine_pairs(ctx, r.data(), c.data(), dim, n_pairs, ref_offsets.data(), cur_offsets.data(), pred, pos, options_.kMaxValidPredictRowDistance,
                                          options_.kMaxValidPredictColDistance, options_.kMaxValidDescriptorDistance, index_pairs_in_cur.data(), flags) == FTK_OK;
        }
    }

    Options &options() { return options_; }
    const Options &options() const { return options_; }

private:
    static std::vector<uint32_t> PackBits(const std::vector<DescriptorType> &set, size_t len, int32_t words) {
        using Traits = DescriptorTraits<DescriptorType>;
        std::vector<uint32_t> out(set.size() * static_cast<size_t>(words), 0u);
        for (size_t i = 0; i < set.size(); ++i)
            for (size_t k = 0; k < len; ++k)
                if (Traits::At(set[i], k)) out[i * words + k / 32] |= 1u << (k % 32);
        return out;
    }
    static std::vector<float> PackFloats(const std::vector<DescriptorType> &set, int32_t dim) {
        using Traits = DescriptorTraits<DescriptorType>;
        std::vector<float> out(set.size() * static_cast<size_t>(dim));
        for (size_t i = 0; i < set.size(); ++i)
            for (int32_t k = 0; k < dim; ++k) out[i * dim + k] = static_cast<float>(Traits::At(set[i], k));
        return out;
    }
    // descriptor_matcher.h:45 declares `virtual float ComputeDistance(ref, cur) = 0` and the reference's subclasses override it
    // (test/test_descriptor_matcher_brief.cpp:33, test_descriptor_matcher_superpoint.cpp:32).  Kept so those subclasses compile
    // unchanged; the default is the metric the GPU evaluates.  It is never called per pair: Run() spot-checks the override.
    virtual float ComputeDistance(const DescriptorType &descriptor_ref, const DescriptorType &descriptor_cur) {
        return TraitsDistance(descriptor_ref, descriptor_cur);
    }
    // The metric of DescriptorTraits<T> on the host, in the arithmetic of the demos' ComputeDistance bodies.
    static float TraitsDistance(const DescriptorType &a, const DescriptorType &b) {
        using Traits = DescriptorTraits<DescriptorType>;
        const size_t n = Traits::Size(a);
        if constexpr (Traits::kBinary) {
            if (n == 0 || Traits::Size(b) == 0) return 2147483647.0f;  // kMaxInt32 (test_descriptor_matcher_brief.cpp:34-36)
            int32_t d = 0;
            for (size_t k = 0; k < n; ++k) d += Traits::At(a, k) != Traits::At(b, k) ? 1 : 0;
            return static_cast<float>(d);
        } else {
            float dot = 0.0f, na = 0.0f, nb = 0.0f;
            for (size_t k = 0; k < n; ++k) {
                dot += Traits::At(a, k) * Traits::At(b, k);
                na += Traits::At(a, k) * Traits::At(a, k);
                nb += Traits::At(b, k) * Traits::At(b, k);
            }
            return 0.5f - dot / std::sqrt(na) / std::sqrt(nb) * 0.5f;  // test_descriptor_matcher_superpoint.cpp:33
        }
    }
    // A subclass's ComputeDistance must be the metric the kernels implement: checked on a few pairs of every call.
    void CheckOverride(const std::vector<DescriptorType> &ref, const std::vector<DescriptorType> &cur) {
        using Traits = DescriptorTraits<DescriptorType>;
        const size_t probes = 4;
        for (size_t q = 0; q < probes && !ref.empty(); ++q) {
            const DescriptorType &a = ref[(q * 7919u) % ref.size()], &b = cur[(q * 104729u + q) % cur.size()];
            const float user = ComputeDistance(a, b), gpu = TraitsDistance(a, b);
            const bool same = Traits::kBinary ? user == gpu : (std::fabs(user - gpu) <= 1e-4f || (user != user && gpu != gpu));
            if (!same)
                throw std::logic_error("feature_tracker_b200: the ComputeDistance override of this DescriptorMatcher subclass is not the metric the GPU "
                                       "kernels evaluate for its descriptor type (Hamming for boolean containers, 0.5 - 0.5*cos for float vectors); "
                                       "got " + std::to_string(user) + ", GPU metric gives " + std::to_string(gpu));
        }
    }
    static uint32_t PrepareIndex(size_t n_ref, std::vector<int32_t> &idx) {
        if (idx.size() == n_ref) return 0u;
        idx.assign(n_ref, -1);  // descriptor_matcher.h:60-62, 98-100
        return FTK_FLAG_NO_INDEX_INPUT;
    }
    static std::vector<float> Flatten(const std::vector<Vec2> &uv) {
        std::vector<float> out(2 * uv.size());
        for (size_t i = 0; i < uv.size(); ++i) out[2 * i] = uv[i].x(), out[2 * i + 1] = uv[i].y();
        return out;
    }
    bool Run(const std::vector<DescriptorType> &ref, const std::vector<DescriptorType> &cur, const std::vector<Vec2> *pred, const std::vector<Vec2> *pos,
             std::vector<int32_t> &idx, uint32_t flags) {
        using Traits = DescriptorTraits<DescriptorType>;
        ftk_context *ctx = Device::Get();
        const int32_t n_ref = static_cast<int32_t>(ref.size()), n_cur = static_cast<int32_t>(cur.size());
        const size_t len = Traits::Size(cur[0]);
        // every descriptor must have the length of cur[0]: the reference's distances index element k of both operands
        // (test_descriptor_matcher_brief.cpp:39-41), which is undefined for ragged sets -- refused here instead.
        for (const auto &d : ref)
            if (Traits::Size(d) != len) return false;
        for (const auto &d : cur)
            if (Traits::Size(d) != len) return false;
        if (len == 0) return false;
        CheckOverride(ref, cur);
        std::vector<float> pred_flat, pos_flat;
        if (pred) pred_flat = Flatten(*pred), pos_flat = Flatten(*pos);
        int rc;
        if constexpr (Traits::kBinary) {
            const int32_t words = static_cast<int32_t>((len + 31) / 32);
            const std::vector<uint32_t> r = PackBits(ref, len, words), c = PackBits(cur, len, words);
            rc = pred ? ftk_match_hamming_nearby(ctx, r.data(), n_ref, c.data(), n_cur, words, pred_flat.data(), pos_flat.data(),
                                                 options_.kMaxValidPredictRowDistance, options_.kMaxValidPredictColDistance,
                                                 options_.kMaxValidDescriptorDistance, idx.data(), flags)
                      : ftk_match_hamming_force(ctx, r.data(), n_ref, c.data(), n_cur, words, options_.kMaxValidDescriptorDistance, idx.data(), flags);
        } else {
            const int32_t dim = static_cast<int32_t>(len);
            const std::vector<float> r = PackFloats(ref, dim), c = PackFloats(cur, dim);
            rc = pred ? ftk_match_cosine_nearby(ctx, r.data(), n_ref, c.data(), n_cur, dim, pred_flat.data(), pos_flat.data(),
                                                options_.kMaxValidPredictRowDistance, options_.kMaxValidPredictColDistance,
                                                options_.kMaxValidDescriptorDistance, idx.data(), flags)
                      : ftk_match_cosine_force(ctx, r.data(), n_ref, c.data(), n_cur, dim, options_.kMaxValidDescriptorDistance, idx.data(), flags);
        }
        return rc == FTK_OK;
    }
    // descriptor_matcher.h:135-157
    bool FillMatchedPixelByPairIndices(const std::vector<int32_t> &idx, const std::vector<Vec2> &pixel_uv_cur, std::vector<Vec2> &matched, std::vector<uint8_t> &status) {
        int32_t valid = 1;
        if (idx.size() != status.size()) {
            status.assign(idx.size(), static_cast<uint8_t>(TrackStatus::kNotTracked));
            valid = 0;
        }
        const std::vector<float> pos = Flatten(pixel_uv_cur);
        std::vector<float> out(2 * idx.size(), 0.0f);
        for (size_t i = 0; i < matched.size() && i < idx.size(); ++i) out[2 * i] = matched[i].x(), out[2 * i + 1] = matched[i].y();
        if (ftk_fill_matched(idx.data(), static_cast<int32_t>(idx.size()), pos.data(), static_cast<int32_t>(pixel_uv_cur.size()), out.data(), status.data(), valid) != FTK_OK)
            return false;
        matched.resize(idx.size());
        for (size_t i = 0; i < idx.size(); ++i) matched[i] = Vec2(out[2 * i], out[2 * i + 1]);
        return true;
    }

    Options options_;
};

// src/direct_method_tracker/direct_method_tracker.h:14-77.  Quaternions are passed as (w, x, y, z) arrays and 3-vectors as
// std::array<float, 3> (the reference's Quat / Vec3 are Eigen types of the absent Slam_Utility; an application that has them
// converts with q.w(), q.x(), ... / v.data()).  Only the camera-frame overload is offered: the world-frame one
// (direct_method_tracker.cpp:8-39) is eight lines of quaternion algebra around it.
enum DirectMethodMethod : uint8_t { kDirectMethodInverse = 0, kDirectMethodDirect = 1, kDirectMethodFast = 2 };

struct DirectMethodOptions {
    uint32_t kMaxTrackPointsNumber = 500;
    uint32_t kMaxIteration = 15;
    int32_t kPatchRowHalfSize = 6;
    int32_t kPatchColHalfSize = 6;
    float kMaxConvergeStep = 1e-6f;
    float kMaxConvergeResidual = 2.0f;
    DirectMethodMethod kMethod = kDirectMethodDirect;
};

class DirectMethod {
public:
    DirectMethod() = default;
    virtual ~DirectMethod() = default;
    DirectMethod(const DirectMethod &) = delete;

    // direct_method_tracker.cpp:41-95
    bool TrackFeatures(const ImagePyramid &ref_pyramid, const ImagePyramid &cur_pyramid, const std::array<float, 4> &K,
                       const std::vector<std::array<float, 3>> &p_c_in_ref, const std::vector<Vec2> &ref_pixel_uv, std::vector<Vec2> &cur_pixel_uv,
                       std::array<float, 4> &q_rc, std::array<float, 3> &p_rc, std::vector<uint8_t> &status) {
        if (ref_pixel_uv.empty()) return false;
        if (cur_pyramid.level() != ref_pyramid.level()) return false;
        if (!ref_pyramid.handle() || !cur_pyramid.handle() || p_c_in_ref.size() != ref_pixel_uv.size()) return false;
        const int32_t n = static_cast<int32_t>(ref_pixel_uv.size());
        ftk_direct_params p;
        ftk_direct_params_default(&p);
        p.max_track_points = options_.kMaxTrackPointsNumber;
        p.max_iteration = options_.kMaxIteration;
        p.patch_row_half = options_.kPatchRowHalfSize;
        p.patch_col_half = options_.kPatchColHalfSize;
        p.max_converge_step = options_.kMaxConvergeStep;
        p.max_converge_residual = options_.kMaxConvergeResidual;
        p.method = static_cast<int32_t>(options_.kMethod);
        uint32_t flags = 0;
        std::vector<float> ref_flat(2 * static_cast<size_t>(n)), cur_flat(2 * static_cast<size_t>(n), 0.0f), pts(3 * static_cast<size_t>(n));
        for (int32_t i = 0; i < n; ++i) {
            ref_flat[2 * i] = ref_pixel_uv[i].x(), ref_flat[2 * i + 1] = ref_pixel_uv[i].y();
            pts[3 * i] = p_c_in_ref[i][0], pts[3 * i + 1] = p_c_in_ref[i][1], pts[3 * i + 2] = p_c_in_ref[i][2];
        }
        if (cur_pixel_uv.size() == ref_pixel_uv.size()) {
            for (int32_t i = 0; i < n; ++i) cur_flat[2 * i] = cur_pixel_uv[i].x(), cur_flat[2 * i + 1] = cur_pixel_uv[i].y();
        } else {
            flags |= FTK_FLAG_NO_PREDICTION;  // :48-50
        }
        if (status.size() != ref_pixel_uv.size()) {
            flags |= FTK_FLAG_NO_STATUS;  // :82-84
            status.assign(ref_pixel_uv.size(), static_cast<uint8_t>(TrackStatus::kTracked));
        }
        const int32_t offsets[2] = {0, n};
        const int32_t image0 = 0;
        const int rc = ftk_direct_method_track(Device::Get(), &p, ref_pyramid.handle(), cur_pyramid.handle(), 1, &image0, &image0, offsets, K.data(), pts.data(),
                                               ref_flat.data(), cur_flat.data(), q_rc.data(), p_rc.data(), status.data(), flags);
        if (rc != FTK_OK) return false;
        cur_pixel_uv.resize(ref_pixel_uv.size());
        for (int32_t i = 0; i < n; ++i) cur_pixel_uv[i] = Vec2(cur_flat[2 * i], cur_flat[2 * i + 1]);
        return true;
    }

    DirectMethodOptions &options() { return options_; }
    const DirectMethodOptions &options() const { return options_; }

private:
    DirectMethodOptions options_;
};

// src/dense_optical_flow_tracker/dense_optical_flow.h:12-65 (Gunnar Farneback).  The reference's flow type is Slam_Utility's
// `Mat` (Eigen::MatrixXf, absent here): the facade hands the two flow components back as rows x cols row-major float vectors.
class DenseOpticalFlow {
public:
    struct Options {
        int32_t kMaxIteration = 10;
        int32_t kHalfPatchSize = 2;
        float kMaxConvergeStep = 1e-6f;
        float kMaxDeltaFlowStep = 1.0f;
    };
    DenseOpticalFlow() = default;
    virtual ~DenseOpticalFlow() = default;

    // dense_optical_flow.cpp:35-85: coarse to fine over the pyramids; flow_rc[0] = row component, flow_rc[1] = column component
    bool Track(const ImagePyramid &ref_pyramid, const ImagePyramid &cur_pyramid, std::array<std::vector<float>, 2> &flow_rc) {
        return Run(ref_pyramid, cur_pyramid, flow_rc, 0u);
    }
    // dense_optical_flow.cpp:7-33 (the GrayImage overload): level 0 only; flow_rc of the image's size is the initial flow
    bool TrackSingleLevel(const ImagePyramid &ref_image, const ImagePyramid &cur_image, std::array<std::vector<float>, 2> &flow_rc) {
        return Run(ref_image, cur_image, flow_rc, FTK_FLAG_SINGLE_LEVEL);
    }
    std::string OpticalFlowMethodName() const { return "Gunnar Farneback"; }
    Options &options() { return options_; }
    const Options &options() const { return options_; }

private:
    bool Run(const ImagePyramid &ref, const ImagePyramid &cur, std::array<std::vector<float>, 2> &flow_rc, uint32_t flags) {
        if (!ref.handle() || !cur.handle()) return false;  // :9-10, :38-39
        if (ref.level() != cur.level()) return false;       // :40
        const size_t n = static_cast<size_t>(ref.rows()) * ref.cols();
        if (flow_rc[0].size() != n || flow_rc[1].size() != n) {  // :18-23
            flow_rc[0].assign(n, 0.0f);
            flow_rc[1].assign(n, 0.0f);
            flags |= FTK_FLAG_NO_PREDICTION;
        }
        ftk_dense_flow_params p;
        p.max_iteration = options_.kMaxIteration;
        p.half_patch_size = options_.kHalfPatchSize;
        p.max_converge_step = options_.kMaxConvergeStep;
        p.max_delta_flow_step = options_.kMaxDeltaFlowStep;
        return ftk_dense_flow_track(Device::Get(), &p, ref.handle(), cur.handle(), 0, 0, flow_rc[0].data(), flow_rc[1].data(), flags) == FTK_OK;
    }
    Options options_;
};

}  // namespace feature_tracker

// The callers' upstream step (test/test_descriptor_matcher_brief.cpp:59-76, test/test_optical_flow.cpp:60-66).  The reference takes
// these classes from the sibling repository Feature_Detector, which is not in its tree: names and option names follow the call
// sites, the arithmetic is the published algorithm documented in ftk_c.h (parity unpinned).  `image` is the device pyramid the
// trackers use (its level 0 is searched), so detection costs no second upload.
namespace feature_detector {

using Vec2 = ::feature_tracker::Vec2;
using BriefType = std::vector<uint8_t>;  // element-wise boolean container, as DescriptorMatcher<BriefType> expects

template <int Kind>
class FeaturePointDetector {
public:
    struct Options {
        float kMinValidResponse = 40.0f;
        int32_t kMinFeatureDistance = 20;
        int32_t kHalfPatchSize = 1;
        float kHarrisK = 0.04f;
    };
    // `features` is in/out like the reference's: what it holds on entry is kept and blocks its neighbourhood; new features are
    // appended in falling response order until `needed_feature_num` are held.
    bool DetectGoodFeatures(const ::feature_tracker::ImagePyramid &image, const uint32_t needed_feature_num, std::vector<Vec2> &features) {
        if (!image.handle()) return false;
        const size_t n_old = features.size();
        if (n_old >= needed_feature_num) return true;
        const int32_t want = static_cast<int32_t>(needed_feature_num - n_old);
        std::vector<float> existing(2 * n_old), found(2 * static_cast<size_t>(want));
        for (size_t i = 0; i < n_old; ++i) existing[2 * i] = features[i].x(), existing[2 * i + 1] = features[i].y();
        ftk_detector_params p;
        p.kind = Kind;
        p.half_patch = options_.kHalfPatchSize;
        p.harris_k = options_.kHarrisK;
        p.min_response = options_.kMinValidResponse;
        p.min_distance = options_.kMinFeatureDistance;
        int32_t n_new = 0;
        if (ftk_detect_features(::feature_tracker::Device::Get(), &p, image.handle(), 0, n_old ? existing.data() : nullptr, static_cast<int32_t>(n_old), want,
                                found.data(), nullptr, &n_new, 0u) != FTK_OK)
            return false;
        for (int32_t i = 0; i < n_new; ++i) features.emplace_back(found[2 * i], found[2 * i + 1]);
        return true;
    }
    Options &options() { return options_; }
    const Options &options() const { return options_; }

private:
    Options options_;
};
using FeaturePointHarrisDetector = FeaturePointDetector<FTK_DETECTOR_HARRIS>;
using FeaturePointShiTomasDetector = FeaturePointDetector<FTK_DETECTOR_SHI_TOMASI>;

class BriefDescriptor {
public:
    struct Options {
        int32_t kLength = 256;
        int32_t kHalfPatchSize = 8;
        uint32_t kPatternSeed = 0;
    };
    // descriptors[i][k] = bit k of feature i; a feature whose patch leaves the image gets an all-zero descriptor (the C ABI's
    // convention; ftk_describe_brief also reports a validity flag).
    bool Compute(const ::feature_tracker::ImagePyramid &image, const std::vector<Vec2> &features, std::vector<BriefType> &descriptors) {
        descriptors.clear();
        if (!image.handle() || options_.kLength <= 0 || options_.kLength % 32 != 0) return false;
        const size_t n = features.size(), words = static_cast<size_t>(options_.kLength / 32);
        std::vector<int8_t> pattern(4 * static_cast<size_t>(options_.kLength));
        ftk_brief_pattern_default(options_.kLength, options_.kHalfPatchSize, options_.kPatternSeed, pattern.data());
        std::vector<float> uv(2 * n);
        for (size_t i = 0; i < n; ++i) uv[2 * i] = features[i].x(), uv[2 * i + 1] = features[i].y();
        std::vector<uint32_t> packed(n * words);
        if (ftk_describe_brief(::feature_tracker::Device::Get(), image.handle(), 0, uv.data(), static_cast<int32_t>(n), pattern.data(), options_.kLength,
                               options_.kHalfPatchSize, packed.data(), nullptr, 0u) != FTK_OK)
            return false;
        descriptors.assign(n, BriefType(static_cast<size_t>(options_.kLength), 0));
        for (size_t i = 0; i < n; ++i) {
            for (int32_t k = 0; k < options_.kLength; ++k) descriptors[i][k] = (packed[i * words + (k >> 5)] >> (k & 31)) & 1u;
        }
        return true;
    }
    Options &options() { return options_; }
    const Options &options() const { return options_; }

private:
    Options options_;
};

}  // namespace feature_detector

#endif
