// Instruction-throughput microbenchmarks for the pipes the hot kernels lean on (POPC for Hamming; I2F / I2FP / FRND /
// F2I conversions and FADD / FMUL for KLT sampling).  Prints warp-instructions per clock per SM.  Used to pin the
// "integer-pipe peak" denominator of the Hamming roofline (SURVEY 8(d) asks for a measured POPC rate on B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kUnroll = 8;

#define BENCH_KERNEL(name, body)                                                     \
    __global__ void name(unsigned *out, unsigned seed) {                              \
        unsigned a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3; \
        unsigned a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;              \
        float f0 = a0 * 0.37f, f1 = a1 * 0.11f, f2 = a2 * 0.23f, f3 = a3 * 0.31f;      \
        float f4 = a4 * 0.41f, f5 = a5 * 0.43f, f6 = a6 * 0.47f, f7 = a7 * 0.53f;      \
        for (int i = 0; i < kIters; ++i) {                                            \
            body                                                                      \
        }                                                                             \
        out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + \
            __float_as_uint(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7);                   \
    }

#define R8(op) op(0) op(1) op(2) op(3) op(4) op(5) op(6) op(7)

#define POPC_OP(k) a##k = __popc(a##k) + a##k;  // POPC + IADD per step
BENCH_KERNEL(k_popc_add, R8(POPC_OP))
#define IADD_OP(k) a##k = a##k + seed;
BENCH_KERNEL(k_iadd, R8(IADD_OP))
#define XOR_POPC_OP(k) a##k += __popc(a##k ^ seed);
BENCH_KERNEL(k_xor_popc_add, R8(XOR_POPC_OP))
#define I2F_U8_OP(k) f##k += (float)(unsigned char)(a##k + i);
BENCH_KERNEL(k_i2f_u8_fadd, R8(I2F_U8_OP))
#define I2FP_S32_OP(k) f##k += (float)(int)(a##k + i);
BENCH_KERNEL(k_i2f_s32_fadd, R8(I2FP_S32_OP))
#define MAGIC_U8_OP(k) f##k += __uint_as_float(0x4B000000u | ((a##k + i) & 255u)) - 8388608.0f;
BENCH_KERNEL(k_magic_u8_fadd, R8(MAGIC_U8_OP))
#define FADD_OP(k) f##k = f##k + 1.25f;
BENCH_KERNEL(k_fadd, R8(FADD_OP))
#define FMUL_OP(k) f##k = __fmul_rn(f##k, 1.0000001f);
BENCH_KERNEL(k_fmul, R8(FMUL_OP))
#define FRND_OP(k) f##k = floorf(f##k) + 0.5f;
BENCH_KERNEL(k_floor_fadd, R8(FRND_OP))
#define F2I_OP(k) a##k += (int)f##k; f##k += 0.5f;
BENCH_KERNEL(k_f2i_iadd, R8(F2I_OP))
#define LOP_OP(k) a##k = (a##k ^ seed) | (a##k >> 3);
BENCH_KERNEL(k_lop_shf, R8(LOP_OP))
#define SHFL_OP(k) a##k = __shfl_xor_sync(0xffffffffu, a##k, 1);
BENCH_KERNEL(k_shfl, R8(SHFL_OP))

template <typename K>
void run(const char *name, K kernel, int ops_per_step, int sm_count, double clock_hz, unsigned *out) {
    const int blocks = sm_count * 8, threads = 256;
    kernel<<<blocks, threads>>>(out, 1u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kernel<<<blocks, threads>>>(out, 2u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double steps = double(blocks) * threads / 32.0 * kIters * kUnroll;  // warp-level steps
    const double per_clk_sm = steps / (ms * 1e-3) / clock_hz / sm_count;
    printf("%-18s %8.3f ms  %7.3f warp-steps/clk/SM  (%d instr per step => %7.3f warp-instr/clk/SM)\n", name, ms, per_clk_sm, ops_per_step,
           per_clk_sm * ops_per_step);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    const double hz = clock_khz * 1e3;
    printf("device %s, %d SMs, max clock %.0f MHz (rates below assume the max clock)\n", p.name, p.multiProcessorCount, hz / 1e6);
    unsigned *out;
    cudaMalloc(&out, sizeof(unsigned) * p.multiProcessorCount * 8 * 256);
    run("popc+iadd", k_popc_add, 2, p.multiProcessorCount, hz, out);
    run("iadd", k_iadd, 1, p.multiProcessorCount, hz, out);
    run("xor+popc+iadd", k_xor_popc_add, 3, p.multiProcessorCount, hz, out);
    run("i2f.u8+fadd", k_i2f_u8_fadd, 3, p.multiProcessorCount, hz, out);
    run("i2f.s32+fadd", k_i2f_s32_fadd, 3, p.multiProcessorCount, hz, out);
    run("magic.u8+fadd", k_magic_u8_fadd, 5, p.multiProcessorCount, hz, out);
    run("fadd", k_fadd, 1, p.multiProcessorCount, hz, out);
    run("fmul", k_fmul, 1, p.multiProcessorCount, hz, out);
    run("floor+fadd", k_floor_fadd, 2, p.multiProcessorCount, hz, out);
    run("f2i+iadd+fadd", k_f2i_iadd, 3, p.multiProcessorCount, hz, out);
    run("lop+shf", k_lop_shf, 2, p.multiProcessorCount, hz, out);
    run("shfl", k_shfl, 1, p.multiProcessorCount, hz, out);
    return 0;
}
