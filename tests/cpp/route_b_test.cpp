// Test driver of INTEGRATION Route B (integration/route_b/optical_flow_klt_b200.h): the B200 subclasses of the reference's own
// tracker classes next to the reference's CPU classes, on the same inputs, through the same public OpticalFlow::TrackFeatures.
// Built by `make -C oracle route_b` against the reference headers + sources where they lie (oracle/_ref/route_b_test).
//   route_b_test <in.bin> <out.bin>
// in : int32 rows, cols, levels, n; u8 ref[rows*cols], cur[rows*cols]; f32 uv[2n]
// out: for variant in (basic, affine, lssd), method in (inverse, direct, fast), mode in (multi, single):
//      int32 ok_cpu, ok_gpu; f32 cpu_uv[2n]; u8 cpu_status[n]; f32 gpu_uv[2n]; u8 gpu_status[n]
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "optical_flow_klt_b200.h"

using namespace feature_tracker;

template <typename T> static std::vector<T> ReadVec(FILE *f, size_t n) {
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) exit(3);
    return v;
}

static void Dump(FILE *out, const std::vector<Vec2> &uv, const std::vector<uint8_t> &st) {
    for (const Vec2 &p : uv) {
        const float xy[2] = {p.x(), p.y()};
        fwrite(xy, sizeof(float), 2, out);
    }
    fwrite(st.data(), 1, st.size(), out);
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *in = fopen(argv[1], "rb");
    FILE *out = fopen(argv[2], "wb");
    if (!in || !out) return 2;
    int32_t hdr[4];
    if (fread(hdr, sizeof(int32_t), 4, in) != 4) return 3;
    const int32_t rows = hdr[0], cols = hdr[1], levels = hdr[2], n = hdr[3];
    const size_t plane = size_t(rows) * cols;
    // + one padded row: the reference's sampler reads the weight-0 "+1" neighbour past the last row
    std::vector<uint8_t> ref_img(plane + cols + 2, 0), cur_img(plane + cols + 2, 0);
    if (fread(ref_img.data(), 1, plane, in) != plane || fread(cur_img.data(), 1, plane, in) != plane) return 3;
    std::vector<float> uv = ReadVec<float>(in, size_t(2) * n);
    std::vector<Vec2> ref_pixel_uv(n);
    for (int i = 0; i < n; ++i) ref_pixel_uv[i] = Vec2(uv[2 * i], uv[2 * i + 1]);

    // test/test_optical_flow.cpp:45-53: pyramids over the caller's image + a caller-owned buffer
    std::vector<uint8_t> ref_buf(plane + 4096), cur_buf(plane + 4096);
    ImagePyramid ref_pyramid, cur_pyramid;
    ref_pyramid.SetPyramidBuff(ref_buf.data(), false);
    cur_pyramid.SetPyramidBuff(cur_buf.data(), false);
    ref_pyramid.SetRawImage(ref_img.data(), rows, cols);
    cur_pyramid.SetRawImage(cur_img.data(), rows, cols);
    if (!ref_pyramid.CreateImagePyramid(levels) || !cur_pyramid.CreateImagePyramid(levels)) return 4;
    const GrayImage ref_image(ref_img.data(), rows, cols), cur_image(cur_img.data(), rows, cols);

    for (int variant = 0; variant < 3; ++variant) {
        for (int method = 0; method < 3; ++method) {
            std::unique_ptr<OpticalFlow> cpu, gpu;
            if (variant == 0) cpu.reset(new OpticalFlowBasicKlt()), gpu.reset(new OpticalFlowBasicKltB200());
            if (variant == 1) cpu.reset(new OpticalFlowAffineKlt()), gpu.reset(new OpticalFlowAffineKltB200());
            if (variant == 2) cpu.reset(new OpticalFlowLssdKlt()), gpu.reset(new OpticalFlowLssdKltB200());
            for (OpticalFlow *t : {cpu.get(), gpu.get()}) {
                t->options().kMethod = static_cast<OpticalFlowMethod>(method);
                t->options().kPatchRowHalfSize = 6;
                t->options().kPatchColHalfSize = 6;
                t->options().kMaxTrackPointsNumber = 1000;
            }
            for (int single = 0; single < 2; ++single) {
                std::vector<Vec2> cpu_uv, gpu_uv;  // empty: no prediction
                std::vector<uint8_t> cpu_st, gpu_st;
                const bool ok_cpu = single ? cpu->TrackFeatures(ref_image, cur_image, ref_pixel_uv, cpu_uv, cpu_st)
                                           : cpu->TrackFeatures(ref_pyramid, cur_pyramid, ref_pixel_uv, cpu_uv, cpu_st);
                const bool ok_gpu = single ? gpu->TrackFeatures(ref_image, cur_image, ref_pixel_uv, gpu_uv, gpu_st)
                                           : gpu->TrackFeatures(ref_pyramid, cur_pyramid, ref_pixel_uv, gpu_uv, gpu_st);
                const int32_t oks[2] = {ok_cpu ? 1 : 0, ok_gpu ? 1 : 0};
                fwrite(oks, sizeof(int32_t), 2, out);
                cpu_uv.resize(n), gpu_uv.resize(n), cpu_st.resize(n), gpu_st.resize(n);
                Dump(out, cpu_uv, cpu_st);
                Dump(out, gpu_uv, gpu_st);
            }
        }
    }
    fclose(in);
    fclose(out);
    return 0;
}
