// BASELINE configs[0] (C1) latency probe: ONE 752x480 frame pair, 200 features, 4-level pyramid, basic KLT kInverse -- the only
// shape the reference itself runs (test/test_optical_flow.cpp:64-73: CreateImagePyramid x2 + TrackFeatures inside its timer).
// A call this small is bound by launches, copies and synchronisation, not by kernel throughput, so it is timed on the host clock
// over many repetitions, through the two reference-facing routes:
//   facade  : ImagePyramid::CreateImagePyramid x2 + OpticalFlowBasicKlt::TrackFeatures (include/feature_tracker_b200/feature_tracker.h),
//             std::vector<Vec2> in / out, ordinary (pageable) host memory -- what an application recompiled against the facade pays;
//   one_call: ftk_track_image_pairs with n_pairs = 1 on pinned host buffers (ftk_alloc_pinned) -- the C ABI's fused entry point.
// Usage: c1_latency <fixture.bin> [reps]    fixture: int32 rows, cols, levels, n, half; u8 ref[rows*cols], cur[rows*cols]; f32 uv[2n]
// Prints one JSON line.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "feature_tracker_b200/feature_tracker.h"

using namespace feature_tracker;
using Clock = std::chrono::steady_clock;

static double Median(std::vector<double> v) {
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}
static double Percentile(std::vector<double> v, double q) {
    std::sort(v.begin(), v.end());
    return v[std::min(v.size() - 1, static_cast<size_t>(q * v.size()))];
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *in = fopen(argv[1], "rb");
    if (!in) return 2;
    const int reps = argc > 2 ? atoi(argv[2]) : 300;
    int32_t hdr[5];
    if (fread(hdr, sizeof(int32_t), 5, in) != 5) return 3;
    const int32_t rows = hdr[0], cols = hdr[1], levels = hdr[2], n = hdr[3], half = hdr[4];
    const size_t plane = static_cast<size_t>(rows) * cols;
    std::vector<uint8_t> ref_img(plane), cur_img(plane);
    std::vector<float> uv(2 * static_cast<size_t>(n));
    if (fread(ref_img.data(), 1, plane, in) != plane || fread(cur_img.data(), 1, plane, in) != plane || fread(uv.data(), sizeof(float), uv.size(), in) != uv.size())
        return 3;
    fclose(in);
    ftk_context *ctx = Device::Get();

    // ---- route 1: the facade, as test/test_optical_flow.cpp:64-73 uses the reference's classes ----
    std::vector<Vec2> ref_pixel_uv(n), cur_pixel_uv;
    for (int i = 0; i < n; ++i) ref_pixel_uv[i] = Vec2(uv[2 * i], uv[2 * i + 1]);
    std::vector<uint8_t> status;
    ImagePyramid ref_pyramid, cur_pyramid;
    ref_pyramid.SetRawImage(ref_img.data(), rows, cols);
    cur_pyramid.SetRawImage(cur_img.data(), rows, cols);
    OpticalFlowBasicKlt klt;
    klt.options().kPatchRowHalfSize = klt.options().kPatchColHalfSize = half;
    klt.options().kMethod = OpticalFlowMethod::kInverse;
    std::vector<double> t_facade, t_track_only;
    uint64_t launches_facade = 0;
    int tracked = 0;
    for (int r = 0; r < reps + 20; ++r) {
        cur_pixel_uv.clear();
        status.clear();
        const uint64_t l0 = ftk_kernel_launches(ctx);
        const auto t0 = Clock::now();
        ref_pyramid.CreateImagePyramid(levels);
        cur_pyramid.CreateImagePyramid(levels);
        const auto t1 = Clock::now();
        if (!klt.TrackFeatures(ref_pyramid, cur_pyramid, ref_pixel_uv, cur_pixel_uv, status)) return 4;
        const auto t2 = Clock::now();
        if (r >= 20) {
            t_facade.push_back(std::chrono::duration<double, std::micro>(t2 - t0).count());
            t_track_only.push_back(std::chrono::duration<double, std::micro>(t2 - t1).count());
        }
        launches_facade = ftk_kernel_launches(ctx) - l0;
    }
    for (uint8_t s : status) tracked += s == 1;

    // ---- route 2: one C-ABI call on pinned host buffers ----
    uint8_t *pin_img = nullptr;
    float *pin_ref = nullptr, *pin_cur = nullptr;
    uint8_t *pin_st = nullptr;
    if (ftk_alloc_pinned(2 * plane, reinterpret_cast<void **>(&pin_img)) != FTK_OK || ftk_alloc_pinned(sizeof(float) * 2 * n, reinterpret_cast<void **>(&pin_ref)) != FTK_OK ||
        ftk_alloc_pinned(sizeof(float) * 2 * n, reinterpret_cast<void **>(&pin_cur)) != FTK_OK || ftk_alloc_pinned(n, reinterpret_cast<void **>(&pin_st)) != FTK_OK)
        return 5;
    memcpy(pin_img, ref_img.data(), plane);
    memcpy(pin_img + plane, cur_img.data(), plane);
    memcpy(pin_ref, uv.data(), sizeof(float) * 2 * n);
    ftk_klt_params p;
    ftk_klt_params_default(&p);
    p.variant = FTK_VARIANT_BASIC, p.method = FTK_METHOD_INVERSE, p.patch_row_half = p.patch_col_half = half;
    const int32_t offsets[2] = {0, n};
    std::vector<double> t_one;
    uint64_t launches_one = 0;
    for (int r = 0; r < reps + 20; ++r) {
        const uint64_t l0 = ftk_kernel_launches(ctx);
        const auto t0 = Clock::now();
        if (ftk_track_image_pairs(ctx, &p, rows, cols, levels, 1, pin_img, pin_img + plane, offsets, pin_ref, pin_cur, pin_st, FTK_FLAG_NO_PREDICTION | FTK_FLAG_NO_STATUS) != FTK_OK)
            return 6;
        const auto t1 = Clock::now();
        if (r >= 20) t_one.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
        launches_one = ftk_kernel_launches(ctx) - l0;
    }
    // both routes must agree bit for bit
    int differ = 0;
    for (int i = 0; i < n; ++i) {
        const float xy[2] = {cur_pixel_uv[i].x(), cur_pixel_uv[i].y()};
        differ += memcmp(xy, pin_cur + 2 * i, sizeof(xy)) != 0 || status[i] != pin_st[i];
    }
    printf("{\"reps\": %d, \"features\": %d, \"tracked\": %d, \"routes_differ\": %d, "
           "\"facade_us\": {\"median\": %.1f, \"p10\": %.1f, \"p90\": %.1f, \"track_call_only_median\": %.1f, \"kernel_launches\": %llu}, "
           "\"one_call_us\": {\"median\": %.1f, \"p10\": %.1f, \"p90\": %.1f, \"kernel_launches\": %llu}}\n",
           reps, n, tracked, differ, Median(t_facade), Percentile(t_facade, 0.1), Percentile(t_facade, 0.9), Median(t_track_only),
           static_cast<unsigned long long>(launches_facade), Median(t_one), Percentile(t_one, 0.1), Percentile(t_one, 0.9),
           static_cast<unsigned long long>(launches_one));
    ftk_free_pinned(pin_img), ftk_free_pinned(pin_ref), ftk_free_pinned(pin_cur), ftk_free_pinned(pin_st);
    return 0;
}
