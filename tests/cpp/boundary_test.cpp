// Boundary checks of the C++ facade against the reference's own class interfaces:
//   * a DescriptorMatcher<T> subclass that overrides the private virtual ComputeDistance, written the way the reference's demos
//     write theirs (test/test_descriptor_matcher_brief.cpp:27-46, test_descriptor_matcher_superpoint.cpp:24-36), compiles and runs;
//   * an override that is NOT the metric the GPU evaluates is refused loudly (std::logic_error);
//   * ragged descriptor sets are refused (return false);
//   * OpticalFlow::TrackFeatures(const GrayImage &, const GrayImage &, ...) (optical_flow.h:41-42) and
//     ImagePyramid::GetImageConst(i).
//   boundary_test <in.bin> <out.bin>      (fixture layout: see tests/test_facade.py)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "feature_tracker_b200/feature_tracker.h"

using namespace feature_tracker;
using BriefType = std::vector<uint8_t>;
using FloatDescriptor = std::vector<float>;
constexpr int32_t kMaxInt32 = 2147483647;

class BriefMatcher : public DescriptorMatcher<BriefType> {
public:
    BriefMatcher() : DescriptorMatcher<BriefType>() {}
    virtual ~BriefMatcher() = default;
    virtual float ComputeDistance(const BriefType &descriptor_ref, const BriefType &descriptor_cur) override {
        if (descriptor_ref.empty() || descriptor_cur.empty()) return kMaxInt32;
        int32_t differing = 0;
        for (uint32_t i = 0; i < descriptor_ref.size(); ++i) differing += descriptor_ref[i] != descriptor_cur[i];
        return static_cast<float>(differing);
    }
};

class CosineMatcher : public DescriptorMatcher<FloatDescriptor> {
public:
    virtual float ComputeDistance(const FloatDescriptor &a, const FloatDescriptor &b) override {
        float dot = 0.0f, na = 0.0f, nb = 0.0f;
        for (size_t k = 0; k < a.size(); ++k) dot += a[k] * b[k], na += a[k] * a[k], nb += b[k] * b[k];
        return 0.5f - dot / std::sqrt(na) / std::sqrt(nb) * 0.5f;
    }
};

// An L1 distance: a legitimate ComputeDistance for the CPU loops, but not what the kernels evaluate.
class L1Matcher : public DescriptorMatcher<FloatDescriptor> {
public:
    virtual float ComputeDistance(const FloatDescriptor &a, const FloatDescriptor &b) override {
        float d = 0.0f;
        for (size_t k = 0; k < a.size(); ++k) d += std::fabs(a[k] - b[k]);
        return d;
    }
};

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
    const int32_t rows = hdr[0], cols = hdr[1], levels = hdr[2], n = hdr[3], n_ref = hdr[4], n_cur = hdr[5], bits = hdr[6], dim = hdr[7];
    std::vector<uint8_t> ref_img = ReadVec<uint8_t>(in, size_t(rows) * cols), cur_img = ReadVec<uint8_t>(in, size_t(rows) * cols);
    std::vector<float> uv = ReadVec<float>(in, size_t(2) * n);
    std::vector<uint8_t> ref_bits = ReadVec<uint8_t>(in, size_t(n_ref) * bits), cur_bits = ReadVec<uint8_t>(in, size_t(n_cur) * bits);
    std::vector<float> ref_f = ReadVec<float>(in, size_t(n_ref) * dim), cur_f = ReadVec<float>(in, size_t(n_cur) * dim);

    // ---- matcher subclasses with ComputeDistance overrides ----
    std::vector<BriefType> rb(n_ref), cb(n_cur);
    for (int i = 0; i < n_ref; ++i) rb[i].assign(ref_bits.begin() + size_t(i) * bits, ref_bits.begin() + size_t(i + 1) * bits);
    for (int j = 0; j < n_cur; ++j) cb[j].assign(cur_bits.begin() + size_t(j) * bits, cur_bits.begin() + size_t(j + 1) * bits);
    std::vector<FloatDescriptor> rf(n_ref), cf(n_cur);
    for (int i = 0; i < n_ref; ++i) rf[i].assign(ref_f.begin() + size_t(i) * dim, ref_f.begin() + size_t(i + 1) * dim);
    for (int j = 0; j < n_cur; ++j) cf[j].assign(cur_f.begin() + size_t(j) * dim, cur_f.begin() + size_t(j + 1) * dim);

    BriefMatcher brief;
    brief.options().kMaxValidDescriptorDistance = 60;
    std::vector<int32_t> idx;
    int32_t okv = brief.ForceMatch(rb, cb, idx) ? 1 : 0;
    fwrite(&okv, sizeof(okv), 1, out);
    fwrite(idx.data(), sizeof(int32_t), n_ref, out);

    CosineMatcher cosine;
    cosine.options().kMaxValidDescriptorDistance = 0.1f;
    idx.clear();
    okv = cosine.ForceMatch(rf, cf, idx) ? 1 : 0;
    fwrite(&okv, sizeof(okv), 1, out);
    fwrite(idx.data(), sizeof(int32_t), n_ref, out);

    L1Matcher l1;
    l1.options().kMaxValidDescriptorDistance = 0.1f;
    int32_t refused = 0;
    try {
        idx.clear();
        l1.ForceMatch(rf, cf, idx);
    } catch (const std::logic_error &) {
        refused = 1;
    }
    fwrite(&refused, sizeof(refused), 1, out);

    std::vector<BriefType> ragged = rb;
    ragged[n_ref / 2].resize(bits / 2);
    idx.clear();
    int32_t ragged_ok = brief.ForceMatch(ragged, cb, idx) ? 1 : 0;
    fwrite(&ragged_ok, sizeof(ragged_ok), 1, out);

    // ---- TrackFeatures(const GrayImage &, const GrayImage &, ...) ----
    GrayImage ref_image(ref_img.data(), rows, cols), cur_image(cur_img.data(), rows, cols);
    std::vector<Vec2> ref_pixel_uv(n), cur_pixel_uv;
    for (int i = 0; i < n; ++i) ref_pixel_uv[i] = Vec2(uv[2 * i], uv[2 * i + 1]);
    std::vector<uint8_t> status;
    OpticalFlowBasicKlt klt;
    klt.options().kMethod = OpticalFlowMethod::kInverse;
    okv = klt.TrackFeatures(ref_image, cur_image, ref_pixel_uv, cur_pixel_uv, status) ? 1 : 0;
    fwrite(&okv, sizeof(okv), 1, out);
    for (int i = 0; i < n; ++i) {
        const float xy[2] = {cur_pixel_uv[i].x(), cur_pixel_uv[i].y()};
        fwrite(xy, sizeof(float), 2, out);
    }
    fwrite(status.data(), 1, n, out);

    // ---- ImagePyramid::GetImageConst(i) ----
    ImagePyramid pyramid;
    pyramid.SetRawImage(ref_img.data(), rows, cols);
    if (!pyramid.CreateImagePyramid(levels)) return 4;
    for (int l = 0; l < levels; ++l) {
        const GrayImage &im = pyramid.GetImageConst(l);
        const int32_t shape[2] = {im.rows(), im.cols()};
        fwrite(shape, sizeof(int32_t), 2, out);
        fwrite(im.data(), 1, size_t(im.rows()) * im.cols(), out);
    }
    // ---- batch entry points == loops over the single-pair calls ----
    {
        std::vector<uint8_t> frames(4 * size_t(rows) * cols);  // frames 0, 1 = ref, cur; 2, 3 = cur, ref (the reverse pair)
        const size_t plane = size_t(rows) * cols;
        memcpy(&frames[0], ref_img.data(), plane), memcpy(&frames[plane], cur_img.data(), plane);
        memcpy(&frames[2 * plane], cur_img.data(), plane), memcpy(&frames[3 * plane], ref_img.data(), plane);
        ImagePyramidBatch batch;
        if (!batch.CreateImagePyramids(frames.data(), rows, cols, 4, levels)) return 5;
        const int half_n = n / 2;
        std::vector<Vec2> uv2(ref_pixel_uv.begin(), ref_pixel_uv.end());
        uv2.insert(uv2.end(), ref_pixel_uv.begin(), ref_pixel_uv.begin() + half_n);
        const std::vector<int32_t> ref_idx = {0, 2}, cur_idx = {1, 3}, offsets = {0, n, n + half_n};
        std::vector<Vec2> batch_uv;
        std::vector<uint8_t> batch_st;
        OpticalFlowAffineKlt affine;
        int32_t batch_ok = affine.TrackFeaturesBatch(batch, ref_idx, cur_idx, offsets, uv2, batch_uv, batch_st) ? 1 : 0;
        ImagePyramid pa, pb;
        pa.SetRawImage(ref_img.data(), rows, cols), pb.SetRawImage(cur_img.data(), rows, cols);
        pa.CreateImagePyramid(levels), pb.CreateImagePyramid(levels);
        std::vector<Vec2> uv_fwd, uv_bwd, first_half(ref_pixel_uv.begin(), ref_pixel_uv.begin() + half_n);
        std::vector<uint8_t> st_fwd, st_bwd;
        affine.TrackFeatures(pa, pb, ref_pixel_uv, uv_fwd, st_fwd);
        affine.TrackFeatures(pb, pa, first_half, uv_bwd, st_bwd);
        int32_t differ = 0;
        for (int i = 0; i < n; ++i) differ += memcmp(&batch_uv[i], &uv_fwd[i], sizeof(Vec2)) != 0 || batch_st[i] != st_fwd[i];
        for (int i = 0; i < half_n; ++i) differ += memcmp(&batch_uv[n + i], &uv_bwd[i], sizeof(Vec2)) != 0 || batch_st[n + i] != st_bwd[i];
        fwrite(&batch_ok, sizeof(batch_ok), 1, out);
        fwrite(&differ, sizeof(differ), 1, out);

        // MatchPairs (binary and float) against per-pair ForceMatch
        const int r_split = n_ref / 3, c_split = n_cur / 2;
        const std::vector<int32_t> ro = {0, r_split, n_ref}, co = {0, c_split, n_cur};
        std::vector<int32_t> pairs_idx, fpairs_idx;
        int32_t pairs_ok = brief.MatchPairs(rb, ro, cb, co, nullptr, nullptr, pairs_idx) ? 1 : 0;
        pairs_ok = (pairs_ok && cosine.MatchPairs(rf, ro, cf, co, nullptr, nullptr, fpairs_idx)) ? 1 : 0;
        int32_t pairs_differ = 0;
        for (int q = 0; q < 2; ++q) {
            std::vector<BriefType> r1(rb.begin() + ro[q], rb.begin() + ro[q + 1]), c1(cb.begin() + co[q], cb.begin() + co[q + 1]);
            std::vector<FloatDescriptor> r2(rf.begin() + ro[q], rf.begin() + ro[q + 1]), c2(cf.begin() + co[q], cf.begin() + co[q + 1]);
            std::vector<int32_t> i1, i2;
            brief.ForceMatch(r1, c1, i1);
            cosine.ForceMatch(r2, c2, i2);
            for (size_t i = 0; i < i1.size(); ++i) pairs_differ += i1[i] != pairs_idx[ro[q] + i] || i2[i] != fpairs_idx[ro[q] + i];
        }
        fwrite(&pairs_ok, sizeof(pairs_ok), 1, out);
        fwrite(&pairs_differ, sizeof(pairs_differ), 1, out);
    }
    fclose(in);
    fclose(out);
    return 0;
}
