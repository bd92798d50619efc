// Exercises the C++ facade (include/feature_tracker_b200/feature_tracker.h) the way the reference's demos use its classes
// (test/test_optical_flow.cpp:41-82, test/test_descriptor_matcher_brief.cpp:48-99): read a binary fixture written by the
// python test, run the trackers / matchers, dump the results for comparison with the oracle.
//   facade_test <in.bin> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "feature_tracker_b200/feature_tracker.h"

using namespace feature_tracker;

template <typename T> static std::vector<T> ReadVec(FILE *f, size_t n) {
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) exit(3);
    return v;
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *in = fopen(argv[1], "rb");
    FILE *out = fopen(argv[2], "wb");
    if (!in || !out) return 2;
    int32_t hdr[8];
    if (fread(hdr, sizeof(int32_t), 8, in) != 8) return 3;
    const int32_t rows = hdr[0], cols = hdr[1], levels = hdr[2], n = hdr[3], n_ref = hdr[4], n_cur = hdr[5], bits = hdr[6];
    std::vector<uint8_t> ref_img = ReadVec<uint8_t>(in, size_t(rows) * cols), cur_img = ReadVec<uint8_t>(in, size_t(rows) * cols);
    std::vector<float> uv = ReadVec<float>(in, size_t(2) * n);
    std::vector<uint8_t> ref_bits = ReadVec<uint8_t>(in, size_t(n_ref) * bits), cur_bits = ReadVec<uint8_t>(in, size_t(n_cur) * bits);
    std::vector<float> pred = ReadVec<float>(in, size_t(2) * n_ref), pos = ReadVec<float>(in, size_t(2) * n_cur);

    ImagePyramid ref_pyramid, cur_pyramid;
    ref_pyramid.SetRawImage(ref_img.data(), rows, cols);
    cur_pyramid.SetRawImage(cur_img.data(), rows, cols);
    if (!ref_pyramid.CreateImagePyramid(levels) || !cur_pyramid.CreateImagePyramid(levels)) return 4;

    std::vector<Vec2> ref_pixel_uv(n);
    for (int i = 0; i < n; ++i) ref_pixel_uv[i] = Vec2(uv[2 * i], uv[2 * i + 1]);

    auto dump_track = [&](OpticalFlow &klt) {
        std::vector<Vec2> cur_pixel_uv;   // empty: no prediction
        std::vector<uint8_t> status;      // empty: all kNotTracked
        const bool ok = klt.TrackFeatures(ref_pyramid, cur_pyramid, ref_pixel_uv, cur_pixel_uv, status);
        int32_t okv = ok ? 1 : 0;
        fwrite(&okv, sizeof(okv), 1, out);
        for (int i = 0; i < n; ++i) {
            const float xy[2] = {cur_pixel_uv[i].x(), cur_pixel_uv[i].y()};
            fwrite(xy, sizeof(float), 2, out);
        }
        fwrite(status.data(), 1, n, out);
    };
    OpticalFlowBasicKlt basic;  // reference defaults: kFast, half 6
    dump_track(basic);
    basic.options().kMethod = OpticalFlowMethod::kInverse;
    basic.options().kPatchRowHalfSize = basic.options().kPatchColHalfSize = 7;
    dump_track(basic);
    OpticalFlowAffineKlt affine;
    dump_track(affine);
    OpticalFlowLssdKlt lssd;
    lssd.consider_patch_luminance() = false;
    dump_track(lssd);
    // empty input and level mismatch return false (optical_flow.cpp:8-9)
    {
        std::vector<Vec2> none, cur_uv;
        std::vector<uint8_t> st;
        ImagePyramid other;
        other.SetRawImage(cur_img.data(), rows, cols);
        other.CreateImagePyramid(levels > 1 ? levels - 1 : levels + 1);
        int32_t flags[2] = {basic.TrackFeatures(ref_pyramid, cur_pyramid, none, cur_uv, st) ? 1 : 0,
                            basic.TrackFeatures(ref_pyramid, other, ref_pixel_uv, cur_uv, st) ? 1 : 0};
        fwrite(flags, sizeof(int32_t), 2, out);
    }

    // BRIEF matching, as in test_descriptor_matcher_brief.cpp: BriefType = element-wise boolean container
    using BriefType = std::vector<uint8_t>;
    class BriefMatcher : public DescriptorMatcher<BriefType> {};
    std::vector<BriefType> ref_desp(n_ref), cur_desp(n_cur);
    for (int i = 0; i < n_ref; ++i) ref_desp[i].assign(ref_bits.begin() + size_t(i) * bits, ref_bits.begin() + size_t(i + 1) * bits);
    for (int j = 0; j < n_cur; ++j) cur_desp[j].assign(cur_bits.begin() + size_t(j) * bits, cur_bits.begin() + size_t(j + 1) * bits);
    std::vector<Vec2> pred_uv(n_ref), cur_uv(n_cur);
    for (int i = 0; i < n_ref; ++i) pred_uv[i] = Vec2(pred[2 * i], pred[2 * i + 1]);
    for (int j = 0; j < n_cur; ++j) cur_uv[j] = Vec2(pos[2 * j], pos[2 * j + 1]);
    BriefMatcher matcher;
    matcher.options().kMaxValidPredictRowDistance = 50;
    matcher.options().kMaxValidPredictColDistance = 50;
    matcher.options().kMaxValidDescriptorDistance = 60;
    std::vector<int32_t> idx;
    int32_t okv = matcher.ForceMatch(ref_desp, cur_desp, idx) ? 1 : 0;
    fwrite(&okv, sizeof(okv), 1, out);
    fwrite(idx.data(), sizeof(int32_t), n_ref, out);
    std::vector<Vec2> matched;
    std::vector<uint8_t> status;
    okv = matcher.NearbyMatch(ref_desp, cur_desp, pred_uv, cur_uv, matched, status) ? 1 : 0;
    fwrite(&okv, sizeof(okv), 1, out);
    for (int i = 0; i < n_ref; ++i) {
        const float xy[2] = {matched[i].x(), matched[i].y()};
        fwrite(xy, sizeof(float), 2, out);
    }
    fwrite(status.data(), 1, n_ref, out);

    // Direct-method pose tracker, as in test_direct_method.cpp: the KLT features read as points of a plane 5 m in front of the camera
    {
        const std::array<float, 4> K = {400.0f, 400.0f, cols / 2.0f, rows / 2.0f};
        std::vector<std::array<float, 3>> p_c_in_ref(n);
        for (int i = 0; i < n; ++i) p_c_in_ref[i] = {(uv[2 * i] - K[2]) / K[0] * 5.0f, (uv[2 * i + 1] - K[3]) / K[1] * 5.0f, 5.0f};
        DirectMethod solver;
        std::vector<Vec2> cur_pixel_uv;
        std::vector<uint8_t> st;
        std::array<float, 4> q_rc = {1.0f, 0.0f, 0.0f, 0.0f};
        std::array<float, 3> p_rc = {0.0f, 0.0f, 0.0f};
        okv = solver.TrackFeatures(ref_pyramid, cur_pyramid, K, p_c_in_ref, ref_pixel_uv, cur_pixel_uv, q_rc, p_rc, st) ? 1 : 0;
        fwrite(&okv, sizeof(okv), 1, out);
        fwrite(q_rc.data(), sizeof(float), 4, out);
        fwrite(p_rc.data(), sizeof(float), 3, out);
        for (int i = 0; i < n; ++i) {
            const float xy[2] = {cur_pixel_uv[i].x(), cur_pixel_uv[i].y()};
            fwrite(xy, sizeof(float), 2, out);
        }
        fwrite(st.data(), 1, n, out);
    }
    // Dense optical flow, as in test_dense_optical_flow.cpp
    {
        DenseOpticalFlow dense;
        std::array<std::vector<float>, 2> flow_rc;
        okv = dense.Track(ref_pyramid, cur_pyramid, flow_rc) ? 1 : 0;
        fwrite(&okv, sizeof(okv), 1, out);
        fwrite(flow_rc[0].data(), sizeof(float), flow_rc[0].size(), out);
        fwrite(flow_rc[1].data(), sizeof(float), flow_rc[1].size(), out);
    }
    // Detect -> describe -> match on the same device pyramids, as in test_descriptor_matcher_brief.cpp:57-95
    {
        feature_detector::FeaturePointHarrisDetector detector;
        detector.options().kMinFeatureDistance = 20;
        detector.options().kMinValidResponse = 40.0f;
        std::vector<Vec2> ref_features, cur_features;
        okv = (detector.DetectGoodFeatures(ref_pyramid, 150, ref_features) && detector.DetectGoodFeatures(cur_pyramid, 150, cur_features)) ? 1 : 0;
        feature_detector::BriefDescriptor descriptor;
        descriptor.options().kLength = 256;
        descriptor.options().kHalfPatchSize = 8;
        std::vector<feature_detector::BriefType> rd, cd;
        okv = (okv && descriptor.Compute(ref_pyramid, ref_features, rd) && descriptor.Compute(cur_pyramid, cur_features, cd)) ? 1 : 0;
        class DemoMatcher : public DescriptorMatcher<feature_detector::BriefType> {} demo;
        demo.options().kMaxValidPredictRowDistance = 50;
        demo.options().kMaxValidPredictColDistance = 50;
        demo.options().kMaxValidDescriptorDistance = 60;
        std::vector<int32_t> pairs;
        okv = (okv && demo.NearbyMatch(rd, cd, ref_features, cur_features, pairs)) ? 1 : 0;
        fwrite(&okv, sizeof(okv), 1, out);
        const int32_t counts[2] = {static_cast<int32_t>(ref_features.size()), static_cast<int32_t>(cur_features.size())};
        fwrite(counts, sizeof(int32_t), 2, out);
        for (const auto *set : {&ref_features, &cur_features})
            for (const Vec2 &f : *set) {
                const float xy[2] = {f.x(), f.y()};
                fwrite(xy, sizeof(float), 2, out);
            }
        fwrite(pairs.data(), sizeof(int32_t), pairs.size(), out);
    }
    fclose(in);
    fclose(out);
    return 0;
}
