// K1: image pyramid construction (replaces ImagePyramid::CreateImagePyramid, an external Slam_Utility routine
// whose call sites are test/test_optical_flow.cpp:70-71; semantics frozen in oracle/shim/datatype_image_pyramid.h:
// level i+1 pixel = (p(2r,2c) + p(2r+1,2c) + p(2r,2c+1) + p(2r+1,2c+1)) >> 2, level sizes rows>>1, cols>>1).
//
// B200 mapping: HBM-bound integer/byte kernel.  One thread owns an 8x8 block of the source level and produces the
// 4x4 / 2x2 / 1x1 outputs of up to three finer-to-coarser levels in registers, so a 4-level pyramid reads level 0
// exactly once and writes levels 1..3 once (479 400 algorithmic bytes per 752x480 image).  A warp reads 256
// contiguous bytes per source row (8-byte vector loads) and writes 128 / 64 / 32 contiguous bytes per output row.
// Pixel sums use packed 16-bit lanes (two pixels per 32-bit register).
#include "ftk_internal.h"

namespace ftk {

namespace {

struct PyramidPass {
    const uint8_t *src;
    uint8_t *dst[3];
    long long src_stride, dst_stride[3];
    int src_rows, src_cols, src_pitch;
    int dst_rows[3], dst_cols[3], dst_pitch[3];
    int n_down;  // 1..3
    int blocks_x, blocks_y;  // 8x8 source blocks per image
};

// Horizontal pair sums of four bytes: returns (b0 + b1) | (b2 + b3) << 16.
__device__ __forceinline__ uint32_t pair_sums(uint32_t w) { return (w & 0x00FF00FFu) + ((w >> 8) & 0x00FF00FFu); }

// Two vertically adjacent words (8 source pixels) -> two output pixels packed into the low 16 bits.
__device__ __forceinline__ uint32_t down2(uint32_t top, uint32_t bottom) {
    const uint32_t s = ((pair_sums(top) + pair_sums(bottom)) >> 2) & 0x00FF00FFu;
    return (s | (s >> 8)) & 0xFFFFu;
}

__global__ void __launch_bounds__(256) pyramid_kernel(PyramidPass a) {
    const int bx = blockIdx.x * blockDim.x + threadIdx.x;
    const int by = blockIdx.y * blockDim.y + threadIdx.y;
    const int image = blockIdx.z;
    if (bx >= a.blocks_x || by >= a.blocks_y) return;

    const uint8_t *src = a.src + image * a.src_stride;
    const int col0 = bx * 8, row0 = by * 8;

    // 8 rows x 8 bytes of the source level.
    uint2 px[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = row0 + i;
        px[i] = (r < a.src_rows && col0 < a.src_pitch) ? __ldg(reinterpret_cast<const uint2 *>(src + (long long)r * a.src_pitch + col0)) : make_uint2(0u, 0u);
    }

    // Level +1: 4 rows x 4 pixels.
    uint32_t l1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) l1[i] = down2(px[2 * i].x, px[2 * i + 1].x) | (down2(px[2 * i].y, px[2 * i + 1].y) << 16);
    {
        uint8_t *dst = a.dst[0] + image * a.dst_stride[0];
        const int c = bx * 4;
        if (c < a.dst_pitch[0]) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = by * 4 + i;
                if (r < a.dst_rows[0]) *reinterpret_cast<uint32_t *>(dst + (long long)r * a.dst_pitch[0] + c) = l1[i];
            }
        }
    }
    if (a.n_down < 2) return;

    // Level +2: 2 rows x 2 pixels.
    uint32_t l2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) l2[i] = down2(l1[2 * i], l1[2 * i + 1]);
    {
        uint8_t *dst = a.dst[1] + image * a.dst_stride[1];
        const int c = bx * 2;
        if (c < a.dst_pitch[1]) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = by * 2 + i;
                if (r < a.dst_rows[1]) *reinterpret_cast<uint16_t *>(dst + (long long)r * a.dst_pitch[1] + c) = static_cast<uint16_t>(l2[i]);
            }
        }
    }
    if (a.n_down < 3) return;

    // Level +3: 1 pixel.
    {
        const uint32_t s = (l2[0] & 0xFFu) + ((l2[0] >> 8) & 0xFFu) + (l2[1] & 0xFFu) + ((l2[1] >> 8) & 0xFFu);
        uint8_t *dst = a.dst[2] + image * a.dst_stride[2];
        if (bx < a.dst_pitch[2] && by < a.dst_rows[2]) dst[(long long)by * a.dst_pitch[2] + bx] = static_cast<uint8_t>(s >> 2);
    }
}

}  // namespace

int LaunchPyramidBuild(ftk_context *ctx, ftk_pyramid *pyr, int first, int count) {
    const PyramidView &v = pyr->view;
    for (int s = 0; s + 1 < v.levels; s += 3) {
        PyramidPass a{};
        a.n_down = (v.levels - 1 - s) < 3 ? (v.levels - 1 - s) : 3;
        a.src = v.base[s] + first * v.image_stride[s];
        a.src_stride = v.image_stride[s];
        a.src_rows = v.rows[s];
        a.src_cols = v.cols[s];
        a.src_pitch = v.pitch[s];
        for (int d = 0; d < a.n_down; ++d) {
            const int l = s + 1 + d;
            a.dst[d] = const_cast<uint8_t *>(v.base[l]) + first * v.image_stride[l];
            a.dst_stride[d] = v.image_stride[l];
            a.dst_rows[d] = v.rows[l];
            a.dst_cols[d] = v.cols[l];
            a.dst_pitch[d] = v.pitch[l];
        }
        a.blocks_x = (v.cols[s] + 7) / 8;
        a.blocks_y = (v.rows[s] + 7) / 8;
        if (a.blocks_x == 0 || a.blocks_y == 0) break;
        const dim3 block(32, 8);
        const dim3 grid((a.blocks_x + block.x - 1) / block.x, (a.blocks_y + block.y - 1) / block.y, count);
        pyramid_kernel<<<grid, block, 0, ctx->stream>>>(a);
        ++ctx->launches;
        FTK_CUDA_CHECK(ctx, cudaGetLastError());
    }
    return FTK_OK;
}

}  // namespace ftk
