// K7: force matching of dense float descriptors (SuperPoint / DISK, up to 256-d) on the 5th-generation tensor cores.
//
// Replaces DescriptorMatcher<T>::ForceMatch (src/descriptor_matcher/descriptor_matcher.h:55-79) with the cosine
// ComputeDistance of test/test_descriptor_matcher_superpoint.cpp:32-34 (d = 0.5 - dot / |a| / |b| * 0.5).
//
// The result is still EXACT (same indices as the reference's fp32 evaluation, lowest j on ties):
//   1. NormPrepKernel  : exact fp32 norms (the reference's sequential sum) and unit-normalised BF16 copies of both descriptor sets
//                        (K padded to a multiple of 64); eight lanes per row, broadcast chain sum.
//   2. CosineTcKernel  : C = A_hat * B_hat^T with tcgen05.mma (BF16 in, FP32 accumulate in TMEM); operands arrive by TMA
//                        (128B-swizzled boxes).  Persistent: one CTA per SM walks a contiguous range of (reference tile, split of
//                        the current set) work items; the 128 x K reference tile stays resident in shared memory while 256-column
//                        tiles of the current set stream through a 5-stage mbarrier pipeline; two TMEM accumulators.  The score
//                        matrix is never written: the epilogue warps read the accumulator with tcgen05.ld and keep a
//                        running top-2 (largest approximate dot, lowest j first) per reference row and split in registers.
//   3. RerankKernel    : |approximate dot - exact dot| <= kEpsDot for unit vectors, so only candidates within
//                        2 * kEpsDot of the row's best approximate dot can be the exact arg-min.  Those (usually one)
//                        are re-evaluated with the reference's own sequential fp32 arithmetic; rows whose top-2 are
//                        both inside the margin (a third candidate could hide) go to an exact scan (ExactScanKernel).
// Kernels 2-4 are launched with programmatic stream serialization (ftk_internal.h: GridDepWait); nothing else is queued per call.
// Warp roles in CosineTcKernel (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 =
// epilogue (one TMEM lane quarter each).
#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>

#include "ftk_internal.h"

namespace ftk {

namespace {

constexpr int kTileM = 128;      // reference rows per CTA (TMEM lanes)
constexpr int kTileN = 256;      // current descriptors per MMA tile (TMEM columns); N = 256 keeps the smem operand reads (A 4 KB + B 8 KB per
                                 // UMMA) under the 128 B/clk shared-memory bandwidth, N = 128 would sit exactly on it
constexpr int kKBlock = 64;      // BF16 elements per 128-byte swizzle row
constexpr int kMaxKBlocks = 4;   // K <= 256
constexpr int kStages = 5;       // pipeline stage = one K block (64 wide) of a 256-row tile of the current set
constexpr int kBoxBytesA = kTileM * kKBlock * 2;  // 16 KiB: one TMA box of the reference tile (128 rows x 128 B)
constexpr int kBoxBytesB = kTileN * kKBlock * 2;  // 32 KiB: one TMA box of the current set (256 rows x 128 B)
constexpr int kTcThreads = 192;
constexpr int kTmemCols = 512;   // two 256-column fp32 accumulators (all of TMEM)
constexpr int kSlotBytesA = kMaxKBlocks * kBoxBytesA;  // 64 KiB
constexpr size_t kTcSmemBytes = 1024 + static_cast<size_t>(kSlotBytesA) + static_cast<size_t>(kStages) * kBoxBytesB + 256;

// |dot(bf16(a_hat), bf16(b_hat)) - exact dot of the unit vectors| <= 2 * 2^-9 + 2^-18 (Cauchy-Schwarz), plus fp32
// accumulation slack on both sides.
constexpr float kEpsDot = 0.0041f;

constexpr unsigned long long kNoKey64 = 0xFFFFFFFFFFFFFFFFull;

struct __align__(16) Top2 {
    float b1;
    int j1;
    float b2;
    int j2;
};

// ---- PTX wrappers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t SmemU32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void MbarInit(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemU32(bar)), "r"(count));
}
__device__ __forceinline__ void MbarExpectTx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemU32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarArrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(SmemU32(bar)) : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(SmemU32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void TmaLoad2D(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(SmemU32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(SmemU32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void UmmaBf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void UmmaCommit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(SmemU32(bar)) : "memory");
}
__device__ __forceinline__ void TcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void TcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Asynchronous: pair with TmemLoadWait() before the registers are read.
__device__ __forceinline__ void TmemLoad32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}
// The registers of every tcgen05.ld issued before this point are valid afterwards.  `r` is threaded through the statement as
// in/out operands so that no use of the loaded values can be scheduled ahead of the wait.
__device__ __forceinline__ void TmemLoadWait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                   "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),
                   "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),
                   "+r"(r[31])
                 :
                 : "memory");
}

// Shared-memory matrix descriptor for a K-major, 128B-swizzled operand tile ([rows][64 bf16], 8-row groups 1024 B
// apart): start address >> 4, LBO = 1 (unused for swizzled K-major), SBO = 1024 B, version 1 (Blackwell), SWIZZLE_128B.
__device__ __forceinline__ uint64_t MakeSmemDesc(const void *tile) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((SmemU32(tile) >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// kind::f16 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kTileN >> 3) << 17) | (static_cast<uint32_t>(kTileM >> 4) << 24);

// ---- kernels -----------------------------------------------------------------------------------------------------

// A descriptor whose norm is zero, not finite or far outside the range in which the unit-normalised BF16 copy bounds the dot-product
// error (kEpsDot) cannot be screened by the tensor-core pass.  The reference still assigns such pairs a distance (e.g. a sum of
// squares that overflows gives dot / inf = 0, i.e. d = 0.5; an underflowed norm gives +-inf): they are evaluated with the exact
// kernel arithmetic instead -- an abnormal CURRENT descriptor becomes an extra exact candidate of every reference row (list
// `abn_cur`), an abnormal REFERENCE row is scanned exactly against the whole current set.  Their BF16 rows are zero.
__device__ __forceinline__ bool AbnormalNorm(float nrm) { return !(nrm >= 1e-12f && nrm <= 1e12f); }

// ---- eight lanes per descriptor row -------------------------------------------------------------------------------------------
// NormPrepKernel and RerankKernel need fp32 sums in the reference's scalar order (k ascending, one rounding per add, no FMA): a
// dependent chain of `dim` adds per row.  Eight lanes share a row: lane l of the group holds the runs of kRun = 16 consecutive
// elements k = 128 q + 16 l + e (q < 2, e < 16; dim <= 256), so a group reads a row as whole 512-byte pieces and every load of a row
// is in flight at once.  The chain then walks the lanes: the owner of the next run adds it to the running sum, which is broadcast to
// the group (16 hand-overs for 256 elements).  No shared memory, no block-wide barrier; the other warps of the SM hide the latency.
constexpr int kRowLanes = 8;
constexpr int kRun = 16;
constexpr int kRowBlocks = kMaxKBlocks * kKBlock / (kRowLanes * kRun);  // blocks of 128 elements per row (2)

// v[q][e] = row[128 q + 16 l + e] for indices below `bound` (0 elsewhere; bound = 0: nothing is read).  VEC: dim % 4 == 0 and a
// 16-byte aligned base, i.e. every float4 of the row is aligned and lies wholly inside or outside the row.
template <bool VEC>
__device__ __forceinline__ void GroupLoadRow(const float *row, int bound, int l, float (&v)[kRowBlocks][kRun]) {
#pragma unroll
    for (int q = 0; q < kRowBlocks; ++q) {
#pragma unroll
        for (int h = 0; h < kRun / 4; ++h) {
            const int k0 = kRowLanes * kRun * q + kRun * l + 4 * h;
            if (VEC) {
                const float4 t = k0 < bound ? __ldg(reinterpret_cast<const float4 *>(row + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[q][4 * h] = t.x, v[q][4 * h + 1] = t.y, v[q][4 * h + 2] = t.z, v[q][4 * h + 3] = t.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[q][4 * h + e] = k0 + e < bound ? __ldg(row + k0 + e) : 0.0f;
            }
        }
    }
}

// s = (..((x[0] + x[1]) + x[2]) + ...) over k < dim with x distributed as above; every lane of the group returns s.  dim >= 1.
// EVERY lane adds its own run to the running sum at every hand-over and the broadcast keeps the owner's result: no divergent branch
// and no predicates around the adds (the other lanes' sums are discarded garbage).
__device__ __forceinline__ float GroupChainSum(const float (&v)[kRowBlocks][kRun], int dim, int lane) {
    const int base = lane & ~(kRowLanes - 1);
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < kRowBlocks; ++q) {
#pragma unroll
        for (int hop = 0; hop < kRowLanes; ++hop) {
            const int k0 = kRowLanes * kRun * q + kRun * hop;
            if (k0 + kRun <= dim) {  // uniform: a whole run
#pragma unroll
                for (int e = 0; e < kRun; ++e) s = k0 + e == 0 ? v[q][e] : __fadd_rn(s, v[q][e]);
                s = __shfl_sync(0xFFFFFFFFu, s, base | hop);
            } else if (k0 < dim) {  // uniform: the row ends inside this run
#pragma unroll
                for (int e = 0; e < kRun; ++e) {
                    if (k0 + e == 0)
                        s = v[q][e];
                    else if (k0 + e < dim)
                        s = __fadd_rn(s, v[q][e]);
                }
                s = __shfl_sync(0xFFFFFFFFu, s, base | hop);
            }
        }
    }
    return s;
}

// norm[i] = sqrt(sequential fp32 dot(a, a)) -- the reference's evaluation order (oracle/shim: k ascending, no FMA) -- and the
// unit-normalised BF16 copy, for BOTH descriptor sets in one launch (blocks [0, ref_blocks) = reference rows, the rest = current
// rows).  Global memory is read once (the row stays in registers for the BF16 pass) and written in 16-byte pieces.
constexpr int kPrepThreads = 128;
constexpr int kPrepRows = kPrepThreads / kRowLanes;
template <bool VEC>
__global__ void __launch_bounds__(kPrepThreads) NormPrepKernel(const float *ref, int n_ref, const float *cur, int n_cur, int ref_blocks, int dim, int k_pad,
                                                              float *ref_norm, float *cur_norm, __nv_bfloat16 *ref_unit, __nv_bfloat16 *cur_unit,
                                                              int *counters, int *next_counters, int *abn_cur) {
    GridDepLaunchDependents();
    if (blockIdx.x == 0 && threadIdx.x == 0) next_counters[0] = next_counters[1] = 0;  // the other set, for the next call
    const bool is_cur = static_cast<int>(blockIdx.x) >= ref_blocks;
    const float *desc = is_cur ? cur : ref;
    const int n = is_cur ? n_cur : n_ref;
    const int lane = threadIdx.x & 31, l = lane & (kRowLanes - 1);
    const int row = (static_cast<int>(blockIdx.x) - (is_cur ? ref_blocks : 0)) * kPrepRows + (threadIdx.x / kRowLanes);
    const bool live = row < n;
    float v[kRowBlocks][kRun];
    GroupLoadRow<VEC>(desc + static_cast<size_t>(live ? row : 0) * dim, live ? dim : 0, l, v);
    float sq[kRowBlocks][kRun];
#pragma unroll
    for (int q = 0; q < kRowBlocks; ++q)
#pragma unroll
        for (int e = 0; e < kRun; ++e) sq[q][e] = __fmul_rn(v[q][e], v[q][e]);
    const float nrm = __fsqrt_rn(GroupChainSum(sq, dim, lane));
    if (!live) return;
    const bool abnormal = AbnormalNorm(nrm);
    if (l == 0) {
        (is_cur ? cur_norm : ref_norm)[row] = nrm;
        if (abnormal && is_cur) abn_cur[atomicAdd(&counters[1], 1)] = row;
    }
    // The BF16 copy only feeds the screening GEMM, whose error margin (kEpsDot) has room for the 1-ulp difference between
    // x * (1 / norm) and x / norm: one multiply per element instead of a division.  Abnormal rows become zero BF16 rows.
    const float inv = abnormal ? 0.0f : 1.0f / nrm;
    __nv_bfloat16 *dst = (is_cur ? cur_unit : ref_unit) + static_cast<size_t>(row) * k_pad;
#pragma unroll
    for (int q = 0; q < kRowBlocks; ++q) {
        const int k0 = kRowLanes * kRun * q + kRun * l;
        if (k0 < k_pad) {  // columns past dim hold 0 (loaded as 0); k_pad is a multiple of 64, so the whole run lies inside the row
            __align__(16) __nv_bfloat162 o[kRun / 2];
#pragma unroll
            for (int e = 0; e < kRun / 2; ++e) o[e] = __floats2bfloat162_rn(abnormal ? 0.0f : v[q][2 * e] * inv, abnormal ? 0.0f : v[q][2 * e + 1] * inv);
#pragma unroll
            for (int h = 0; h < kRun / 8; ++h) reinterpret_cast<uint4 *>(dst + k0)[h] = reinterpret_cast<const uint4 *>(o)[h];
        }
    }
}

// Persistent: one CTA per SM.  The work items (reference tile m, split of the current set), numbered m-major, are dealt out as one
// contiguous range per CTA, so a CTA changes its reference tile at most a few times in the whole kernel and the rest of shared memory
// can go to a deep pipeline of current-set stages.  The three roles run the same item loop and carry their pipeline state (stage /
// accumulator / reference-tile parities) across items: the TMA stream, the MMA issue and the epilogue of consecutive items overlap, and
// every CTA gets the same number of items to within one.
__global__ void __launch_bounds__(kTcThreads, 1)
CosineTcKernel(const __grid_constant__ CUtensorMap map_ref, const __grid_constant__ CUtensorMap map_cur, int n_ref, int n_cur, int k_blocks,
               int tiles_per_split, int n_tiles, int n_splits, int n_items, Top2 *__restrict__ out, int n_ref_pad, float floor_dot) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem_a + kSlotBytesA;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_b + kStages * kBoxBytesB);
    uint64_t *bar_a_full = &bars[0];
    uint64_t *bar_a_empty = &bars[1];
    uint64_t *bar_full = &bars[2];                     // [kStages]
    uint64_t *bar_empty = bar_full + kStages;          // [kStages]
    uint64_t *bar_acc_full = bar_empty + kStages;      // [2]
    uint64_t *bar_acc_empty = bar_acc_full + 2;        // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int item_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * n_items / gridDim.x);
    const int item_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * n_items / gridDim.x);

    if (threadIdx.x == 0) {
        MbarInit(bar_a_full, 1);
        MbarInit(bar_a_empty, 1);
        for (int s = 0; s < kStages; ++s) {
            MbarInit(&bar_full[s], 1);
            MbarInit(&bar_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            MbarInit(&bar_acc_full[a], 1);
            MbarInit(&bar_acc_empty[a], 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemU32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    TcFenceBefore();
    __syncthreads();
    TcFenceAfter();
    const uint32_t tmem_base = *tmem_slot;
    GridDepLaunchDependents();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            GridDepWait();  // the BF16 unit descriptors come from NormPrepKernel; nothing else in this kernel reads its output
            int stage = 0, m_loaded = -1;
            uint32_t phase = 0, ref_phase = 0;
            for (int item = item_begin; item < item_end; ++item) {
                const int m_tile = item / n_splits;
                if (m_tile != m_loaded) {
                    MbarWait(bar_a_empty, ref_phase ^ 1u);  // the MMAs on the previous reference tile have retired
                    MbarExpectTx(bar_a_full, static_cast<uint32_t>(k_blocks) * kBoxBytesA);
                    for (int kb = 0; kb < k_blocks; ++kb) TmaLoad2D(smem_a + kb * kBoxBytesA, &map_ref, bar_a_full, kb * kKBlock, m_tile * kTileM);
                    ref_phase ^= 1u;
                    m_loaded = m_tile;
                }
                const int t_begin = (item % n_splits) * tiles_per_split, t_end = min(n_tiles, t_begin + tiles_per_split);
                for (int t = t_begin; t < t_end; ++t) {
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        MbarWait(&bar_empty[stage], phase ^ 1u);
                        MbarExpectTx(&bar_full[stage], kBoxBytesB);
                        TmaLoad2D(smem_b + stage * kBoxBytesB, &map_cur, &bar_full[stage], kb * kKBlock, t * kTileN);
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            int stage = 0, acc = 0, m_loaded = -1;
            uint32_t phase = 0, acc_phase = 0, ref_phase = 0;
            for (int item = item_begin; item < item_end; ++item) {
                const int m_tile = item / n_splits;
                const int t_begin = (item % n_splits) * tiles_per_split, t_end = min(n_tiles, t_begin + tiles_per_split);
                if (m_tile != m_loaded) {
                    MbarWait(bar_a_full, ref_phase);
                    TcFenceAfter();
                    ref_phase ^= 1u;
                    m_loaded = m_tile;
                }
                for (int t = t_begin; t < t_end; ++t) {
                    MbarWait(&bar_acc_empty[acc], acc_phase ^ 1u);
                    TcFenceAfter();
                    const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kTileN);
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        MbarWait(&bar_full[stage], phase);
                        TcFenceAfter();
                        const uint64_t adesc = MakeSmemDesc(smem_a + kb * kBoxBytesA);
                        const uint64_t bdesc = MakeSmemDesc(smem_b + stage * kBoxBytesB);
#pragma unroll
                        for (int k = 0; k < kKBlock / 16; ++k) {
                            // each UMMA consumes K = 16 BF16 = 32 bytes of the 128-byte swizzle row: advance the start address by 32 B
                            UmmaBf16(tmem_d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), kInstrDesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        UmmaCommit(&bar_empty[stage]);  // the stage's operands may be overwritten once these MMAs retire
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    UmmaCommit(&bar_acc_full[acc]);  // the accumulator is complete
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1u;
                }
                // the reference tile may be replaced once the MMAs of its last item retire
                if (item + 1 < item_end && (item + 1) / n_splits != m_tile) UmmaCommit(bar_a_empty);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: running top-2 per reference row =====================
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        // A match needs distance < max_dist, i.e. a cosine above 1 - 2 * max_dist; approximate dots below `floor_dot` (that bound
        // minus the BF16 error margin) cannot belong to a match and never enter the top-2.  With the usual tight thresholds the
        // costly "new best in this chunk" path (taken by the whole warp when any of its 32 rows sees a new maximum) then runs
        // for real candidates only instead of for every running maximum of the random background.
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = item_begin; item < item_end; ++item) {
        const int split = item % n_splits, row = (item / n_splits) * kTileM + quarter * 32 + lane;
        const int t_begin = split * tiles_per_split, t_end = min(n_tiles, t_begin + tiles_per_split);
        float b1 = floor_dot, b2 = -INFINITY;
        int j1 = -1;
        for (int t = t_begin; t < t_end; ++t) {
            MbarWait(&bar_acc_full[acc], acc_phase);
            TcFenceAfter();
            const int n0 = t * kTileN;
            const bool partial = n0 + kTileN > n_cur;
            // The TMEM read of chunk c + 1 is in flight while chunk c is reduced (two register buffers).
            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * kTileN);
            auto reduce_chunk = [&](uint32_t (&r)[32], int chunk) {
                const int c0 = n0 + chunk * 32;
                if (partial) {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (c0 + c >= n_cur) r[c] = 0xFF800000u;  // -inf: zero-filled columns past the end never win
                }
                // chunk maximum with a shallow tree (fmaxf drops NaN)
                float m4[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    m4[q] = fmaxf(fmaxf(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1])), fmaxf(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
                const float m = fmaxf(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])), fmaxf(fmaxf(m4[4], m4[5]), fmaxf(m4[6], m4[7])));
                if (m > b1) {
                    // A new best lives in this chunk (rare per lane): locate it (lowest column on ties) and the chunk's runner-up.
                    // Only the VALUE of the overall second best is needed later (rows whose two best are both inside the
                    // error margin are re-scanned exactly), so no index is kept for it.
                    float cb1 = -INFINITY, cb2 = -INFINITY;
                    int cj = 0;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float v = __uint_as_float(r[c]);
                        if (v > cb1) {
                            cb2 = cb1;
                            cb1 = v;
                            cj = c;
                        } else {
                            cb2 = fmaxf(cb2, v);
                        }
                    }
                    b2 = fmaxf(b1, cb2);  // old best and the chunk's runner-up compete for second place (old b2 <= old b1)
                    b1 = cb1;
                    j1 = c0 + cj;
                } else {
                    b2 = fmaxf(b2, m);  // m <= b1: the chunk can only improve the second place
                }
            };
            uint32_t ra[32], rb[32];
            TmemLoad32(tbase, ra);
            TmemLoadWait(ra);
#pragma unroll 1
            for (int chunk = 0; chunk < kTileN / 32; chunk += 2) {
                TmemLoad32(tbase + static_cast<uint32_t>((chunk + 1) * 32), rb);  // in flight while ra is reduced
                reduce_chunk(ra, chunk);
                TmemLoadWait(rb);
                if (chunk + 2 < kTileN / 32) TmemLoad32(tbase + static_cast<uint32_t>((chunk + 2) * 32), ra);  // in flight while rb is reduced
                reduce_chunk(rb, chunk + 1);
                if (chunk + 2 < kTileN / 32) TmemLoadWait(ra);
            }
            TcFenceBefore();
            __syncwarp();
            if (lane == 0) MbarArrive(&bar_acc_empty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
        if (row < n_ref) {
            Top2 o;
            o.b1 = j1 >= 0 ? b1 : -INFINITY;  // no candidate above the floor in this split
            o.j1 = j1, o.b2 = b2, o.j2 = b2 > -INFINITY ? 0 : -1;  // j2 only says whether a second candidate exists
            out[static_cast<size_t>(split) * n_ref_pad + row] = o;
        }
        }
    }

    TcFenceBefore();
    __syncthreads();
    if (warp == 1) {
        TcFenceAfter();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// Order-preserving key for a non-NaN float distance (with -0 canonicalised to +0).
__device__ __forceinline__ unsigned FloatKey(float d) {
    d = d + 0.0f;
    const unsigned b = __float_as_uint(d);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float KeyFloat(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }

// The reference's distance, bit for bit: sequential dot (k ascending, no FMA), 0.5 - dot / |a| / |b| * 0.5.
__device__ __forceinline__ float ExactDistance(const float *a, const float *b, int dim, float na, float nb) {
    float s = __fmul_rn(__ldg(a), __ldg(b));
#pragma unroll 8
    for (int k = 1; k < dim; ++k) s = __fadd_rn(s, __fmul_rn(__ldg(a + k), __ldg(b + k)));
    return __fsub_rn(0.5f, __fmul_rn(__fdiv_rn(__fdiv_rn(s, na), nb), 0.5f));
}

// Exact re-evaluation of the candidates inside the error margin.  Eight lanes own a reference row (see GroupLoadRow):
//   1. lane l reads the top-2 of column splits l and l + 8; the group forms the row's best approximate dot, flags the splits whose
//      best lies inside the margin (candidates) and marks "crowded" splits (both of its top-2 inside the margin: a third candidate
//      could hide) for the exact scan;
//   2. per candidate (almost always exactly one) the group loads both descriptors, forms the 256 products a[k] * b[k] -- each one
//      correctly rounded multiply, the reference's -- and adds them in ascending k (GroupChainSum);
//   3. abnormal current descriptors (see AbnormalNorm) are exact candidates of every row;
//   4. a row without scan work is FINISHED here (idx written when its best distance < max_dist); a row with crowded splits, or an
//      abnormal reference row, parks its key in best[] and queues (row, split mask) for ExactScanKernel, which finishes it.
constexpr int kRerankThreads = 128;
constexpr int kRerankRows = kRerankThreads / kRowLanes;
constexpr int kMaxSplits = 2 * kRowLanes;
template <bool VEC>
__global__ void __launch_bounds__(kRerankThreads, 5) RerankKernel(const float *ref, int n_ref, const float *cur, int dim, const float *ref_norm,
                                                              const float *cur_norm, const Top2 *top, int n_splits, int n_ref_pad,
                                                              unsigned long long *best, int2 *work, int *counters, const int *abn_cur, float max_dist,
                                                              int fill_unmatched, int *idx) {
    const int lane = threadIdx.x & 31, l = lane & (kRowLanes - 1), base = lane & ~(kRowLanes - 1);
    const int i = blockIdx.x * kRerankRows + threadIdx.x / kRowLanes;
    const bool live = i < n_ref;
    const unsigned all_splits = (1u << n_splits) - 1u;  // n_splits <= kMaxSplits = 16
    GridDepLaunchDependents();
    GridDepWait();
    const int n_abn_all = counters[1];

    // ---- 1. screening ----
    Top2 t0, t1;
    t0.b1 = t1.b1 = -INFINITY, t0.j1 = t1.j1 = -1, t0.b2 = t1.b2 = -INFINITY, t0.j2 = t1.j2 = -1;
    float na = 1.0f;
    if (live) {
        if (l < n_splits) t0 = top[static_cast<size_t>(l) * n_ref_pad + i];
        if (l + kRowLanes < n_splits) t1 = top[static_cast<size_t>(l + kRowLanes) * n_ref_pad + i];
        na = ref_norm[i];
    }
    float gmax = fmaxf(t0.b1, t1.b1);
#pragma unroll
    for (int o = kRowLanes / 2; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, o));
    const float thr = gmax - 2.0f * kEpsDot;
    const bool trusted = live && !AbnormalNorm(na) && gmax > -INFINITY;  // an abnormal row: nothing the tensor-core pass said about it counts
    const bool cand0 = trusted && t0.j1 >= 0 && t0.b1 >= thr, crowded0 = cand0 && t0.j2 >= 0 && t0.b2 >= thr;
    const bool cand1 = trusted && t1.j1 >= 0 && t1.b1 >= thr, crowded1 = cand1 && t1.j2 >= 0 && t1.b2 >= thr;
    const unsigned todo0 = __ballot_sync(0xFFFFFFFFu, cand0 && !crowded0), todo1 = __ballot_sync(0xFFFFFFFFu, cand1 && !crowded1);
    const unsigned scan0 = __ballot_sync(0xFFFFFFFFu, crowded0), scan1 = __ballot_sync(0xFFFFFFFFu, crowded1);
    unsigned todo = ((todo0 >> base) & 0xFFu) | (((todo1 >> base) & 0xFFu) << kRowLanes);  // the row's candidate splits
    const unsigned scan = !live ? 0u : AbnormalNorm(na) ? all_splits : ((scan0 >> base) & 0xFFu) | (((scan1 >> base) & 0xFFu) << kRowLanes);

    // ---- 2. exact distances of the candidates ----
    unsigned long long key = kNoKey64;
    while (__any_sync(0xFFFFFFFFu, todo != 0u)) {
        const bool has = todo != 0u;
        const int split = has ? __ffs(static_cast<int>(todo)) - 1 : 0;
        todo &= todo - 1u;
        const int src = base | (split & (kRowLanes - 1));
        const int j_lo = __shfl_sync(0xFFFFFFFFu, t0.j1, src), j_hi = __shfl_sync(0xFFFFFFFFu, t1.j1, src);
        const int j = has ? (split < kRowLanes ? j_lo : j_hi) : 0;
        float a[kRowBlocks][kRun], b[kRowBlocks][kRun];
        GroupLoadRow<VEC>(ref + static_cast<size_t>(has ? i : 0) * dim, has ? dim : 0, l, a);
        GroupLoadRow<VEC>(cur + static_cast<size_t>(j) * dim, has ? dim : 0, l, b);
        const float nb = has ? cur_norm[j] : 1.0f;
#pragma unroll
        for (int q = 0; q < kRowBlocks; ++q)
#pragma unroll
            for (int e = 0; e < kRun; ++e) a[q][e] = __fmul_rn(a[q][e], b[q][e]);
        const float dot = GroupChainSum(a, dim, lane);
        const float d = __fsub_rn(0.5f, __fmul_rn(__fdiv_rn(__fdiv_rn(dot, na), nb), 0.5f));
        if (has && d == d) {
            const unsigned long long k64 = (static_cast<unsigned long long>(FloatKey(d)) << 32) | static_cast<unsigned>(j);
            key = k64 < key ? k64 : key;
        }
    }
    if (l != 0 || !live) return;
    // ---- 3. abnormal current descriptors: exact candidates of every row (normally none) ----
    const int n_abn = scan == all_splits ? 0 : n_abn_all;  // a full scan covers them anyway
    for (int q = 0; q < n_abn; ++q) {
        const int j = abn_cur[q];
        const float d = ExactDistance(ref + static_cast<size_t>(i) * dim, cur + static_cast<size_t>(j) * dim, dim, na, cur_norm[j]);
        if (d == d) {
            const unsigned long long k64 = (static_cast<unsigned long long>(FloatKey(d)) << 32) | static_cast<unsigned>(j);
            key = k64 < key ? k64 : key;
        }
    }
    // ---- 4. finish the row, or hand it to the exact scan ----
    if (scan != 0u) {
        best[i] = key;
        work[atomicAdd(&counters[0], 1)] = make_int2(i, static_cast<int>(scan));
    } else if (key != kNoKey64 && KeyFloat(static_cast<unsigned>(key >> 32)) < max_dist) {
        idx[i] = static_cast<int>(key & 0xFFFFFFFFull);
    } else if (fill_unmatched) {
        idx[i] = -1;  // otherwise the caller's entry stays (descriptor_matcher.h: only matched rows are assigned)
    }
}

// One block per queued row: exact scan of the column ranges of the flagged splits, merged with the row's parked key; writes idx.
__global__ void __launch_bounds__(128) ExactScanKernel(const float *ref, const float *cur, int n_cur, int dim, const float *ref_norm, const float *cur_norm,
                                                      const int2 *work, const int *counters, int cols_per_split, int n_splits,
                                                      const unsigned long long *best, float max_dist, int fill_unmatched, int *idx) {
    __shared__ unsigned long long s_key[4];
    GridDepWait();
    const int n = counters[0];
    for (int w = blockIdx.x; w < n; w += gridDim.x) {
        const int i = work[w].x;
        const unsigned mask = static_cast<unsigned>(work[w].y);
        const float *a = ref + static_cast<size_t>(i) * dim;
        const float na = ref_norm[i];
        unsigned long long key = kNoKey64;
        for (int s = 0; s < n_splits; ++s) {
            if (!((mask >> s) & 1u)) continue;
            const int j_begin = s * cols_per_split, j_end = min(n_cur, j_begin + cols_per_split);
            for (int j = j_begin + threadIdx.x; j < j_end; j += blockDim.x) {
                const float d = ExactDistance(a, cur + static_cast<size_t>(j) * dim, dim, na, cur_norm[j]);
                if (d == d) {
                    const unsigned long long k = (static_cast<unsigned long long>(FloatKey(d)) << 32) | static_cast<unsigned>(j);
                    key = k < key ? k : key;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
            key = other < key ? other : key;
        }
        if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = key;
        __syncthreads();
        if (threadIdx.x == 0) {
            key = best[i];
            for (int q = 0; q < 4; ++q) key = s_key[q] < key ? s_key[q] : key;
            if (key != kNoKey64 && KeyFloat(static_cast<unsigned>(key >> 32)) < max_dist)
                idx[i] = static_cast<int>(key & 0xFFFFFFFFull);
            else if (fill_unmatched)
                idx[i] = -1;
        }
        __syncthreads();
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn GetEncodeTiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows][k_pad] BF16, K-major; box = 64 elements x box_rows rows, 128B swizzle, out-of-range rows read as zero.
bool MakeMap(CUtensorMap *map, const __nv_bfloat16 *base, int rows, int k_pad, int box_rows) {
    EncodeTiledFn encode = GetEncodeTiled();
    if (!encode) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_pad), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(k_pad) * 2};
    const cuuint32_t box[2] = {kKBlock, static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t elem[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16 *>(base), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int Blocks(int n, int threads) { return (n + threads - 1) / threads; }

}  // namespace

// Returns FTK_ERR_UNSUPPORTED when the tensor-core path does not cover the shape (dim > 256): the caller then runs the
// exact CUDA-core kernel of match.cu.  Four launches: NormPrepKernel (both sets) -> CosineTcKernel -> RerankKernel (finishes the
// rows) -> ExactScanKernel (finishes the few rows the screening could not decide; its blocks exit at once when there are none).
int LaunchCosineForceTensor(ftk_context *ctx, const float *d_ref, int n_ref, const float *d_cur, int n_cur, int dim, float max_dist, int *d_idx,
                            bool fill_unmatched) {
    if (dim > kMaxKBlocks * kKBlock) return FTK_ERR_UNSUPPORTED;
    if (n_ref == 0) return FTK_OK;
    cudaStream_t st = ctx->stream;
    const int k_blocks = (dim + kKBlock - 1) / kKBlock, k_pad = k_blocks * kKBlock;
    const int m_tiles = (n_ref + kTileM - 1) / kTileM, n_tiles = (n_cur + kTileN - 1) / kTileN;
    // Work items = (reference tile, split of the current set), walked by one persistent CTA per SM.  Pick the split count that
    // gives every CTA the fewest current-set tiles in total (ties: fewer splits = fewer partial top-2 records to merge).
    int splits = 1;
    {
        double best_cost = 0.0;
        const int max_splits = n_tiles < kMaxSplits ? n_tiles : kMaxSplits;
        for (int sp = 1; sp <= max_splits; ++sp) {
            const int per = (n_tiles + sp - 1) / sp, real = (n_tiles + per - 1) / per;
            const long long items = static_cast<long long>(m_tiles) * real;
            const long long per_cta = (items + ctx->sm_count - 1) / ctx->sm_count;
            // time ~ items per CTA * (tiles per item + a fraction of a tile for the item switch)
            const double cost = static_cast<double>(per_cta) * (per + 0.1);
            if (sp == 1 || cost < best_cost * 0.99) {
                best_cost = cost;
                splits = real;
            }
        }
    }
    const int tiles_per_split = (n_tiles + splits - 1) / splits;
    splits = (n_tiles + tiles_per_split - 1) / tiles_per_split;
    const int n_ref_pad = m_tiles * kTileM;

    // scratch: norms | unit bf16 copies | top-2 | counters, parked keys, work list, abnormal-column list
    const size_t bytes_norm = sizeof(float) * (static_cast<size_t>(n_ref) + n_cur);
    const size_t bytes_unit = sizeof(__nv_bfloat16) * static_cast<size_t>(k_pad) * (static_cast<size_t>(n_ref) + n_cur);
    if (int rc = EnsureDevice(ctx, ctx->d_work3, bytes_norm + 256)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_work1, bytes_unit + 1024)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_work2, sizeof(Top2) * static_cast<size_t>(splits) * n_ref_pad + 256)) return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_work0, (sizeof(unsigned long long) + sizeof(int2)) * static_cast<size_t>(n_ref) + sizeof(int) * static_cast<size_t>(n_cur) + 256))
        return rc;
    if (int rc = EnsureDevice(ctx, ctx->d_cos_counters, 32)) return rc;
    float *ref_norm = static_cast<float *>(ctx->d_work3.ptr), *cur_norm = ref_norm + n_ref;
    __nv_bfloat16 *ref_unit = static_cast<__nv_bfloat16 *>(ctx->d_work1.ptr);
    __nv_bfloat16 *cur_unit = ref_unit + static_cast<size_t>(n_ref) * k_pad;
    Top2 *top = static_cast<Top2 *>(ctx->d_work2.ptr);
    // [0] rows queued for the exact scan, [1] abnormal current descriptors; two sets, see ftk_context::d_cos_counters
    if (!ctx->cos_counters_clean) FTK_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_cos_counters.ptr, 0, 32, st));
    ctx->cos_counters_clean = false;  // until this call's launches are all queued
    int *counters = static_cast<int *>(ctx->d_cos_counters.ptr) + 4 * (ctx->cos_calls & 1u);
    int *next_counters = static_cast<int *>(ctx->d_cos_counters.ptr) + 4 * ((ctx->cos_calls + 1u) & 1u);
    ++ctx->cos_calls;
    unsigned long long *best = static_cast<unsigned long long *>(ctx->d_work0.ptr);
    int2 *work = reinterpret_cast<int2 *>(best + n_ref);
    int *abn_cur = reinterpret_cast<int *>(work + n_ref);

    CUtensorMap map_ref, map_cur;
    if (!MakeMap(&map_ref, ref_unit, n_ref, k_pad, kTileM) || !MakeMap(&map_cur, cur_unit, n_cur, k_pad, kTileN))
        return SetError(ctx, FTK_ERR_CUDA, "cuTensorMapEncodeTiled failed for the descriptor matrices");

    // float4 row loads need 16-byte aligned rows: dim % 4 == 0 and aligned bases (device pointers may come from the caller)
    const bool vec = dim % 4 == 0 && (reinterpret_cast<uintptr_t>(d_ref) | reinterpret_cast<uintptr_t>(d_cur)) % 16 == 0;
    const int ref_blocks = Blocks(n_ref, kPrepRows);
    // An ordinary launch: only the later kernels of the call start early (a first kernel that did would chain the early launches across
    // calls and let this call's CTAs take whatever slots the previous call frees first; measured gain of allowing it: < 0.3 us).
    (vec ? NormPrepKernel<true> : NormPrepKernel<false>)<<<ref_blocks + Blocks(n_cur, kPrepRows), kPrepThreads, 0, st>>>(
        d_ref, n_ref, d_cur, n_cur, ref_blocks, dim, k_pad, ref_norm, cur_norm, ref_unit, cur_unit, counters, next_counters, abn_cur);

    // cos > 1 - 2 * max_dist is necessary for distance < max_dist; 3 * kEpsDot covers the BF16 dot error and the fp32 rounding of the
    // distance formula.  NaN / huge thresholds give NaN / -inf floors: nothing or everything passes, as in the reference.
    const float floor_dot = 1.0f - 2.0f * max_dist - 3.0f * kEpsDot;
    // The opt-in is per device and the ABI allows one process to hold contexts on several GPUs: set it before every launch (cheap).
    FTK_CUDA_CHECK(ctx, cudaFuncSetAttribute(CosineTcKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kTcSmemBytes)));
    const int n_items = m_tiles * splits;
    ProfBegin(ctx);
    FTK_CUDA_CHECK(ctx, LaunchDependent(CosineTcKernel, n_items < ctx->sm_count ? n_items : ctx->sm_count, kTcThreads, kTcSmemBytes, st, map_ref, map_cur, n_ref,
                                        n_cur, k_blocks, tiles_per_split, n_tiles, splits, n_items, top, n_ref_pad, floor_dot));
    ProfEnd(ctx);
    FTK_CUDA_CHECK(ctx, LaunchDependent(vec ? RerankKernel<true> : RerankKernel<false>, Blocks(n_ref, kRerankRows), kRerankThreads, 0, st, d_ref, n_ref, d_cur, dim,
                                        ref_norm, cur_norm, top, splits, n_ref_pad, best, work, counters, abn_cur, max_dist, fill_unmatched ? 1 : 0, d_idx));
    FTK_CUDA_CHECK(ctx, LaunchDependent(ExactScanKernel, ctx->sm_count * 2, 128, 0, st, d_ref, d_cur, n_cur, dim, ref_norm, cur_norm, work, counters,
                                        tiles_per_split * kTileN, splits, best, max_dist, fill_unmatched ? 1 : 0, d_idx));
    FTK_CUDA_CHECK(ctx, cudaGetLastError());
    ctx->cos_counters_clean = true;
    ctx->launches += 4;
    ctx->d_last_scan_items = counters;
    return FTK_OK;
}

}  // namespace ftk
