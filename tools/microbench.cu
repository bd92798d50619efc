// Instruction-throughput microbenchmarks for the pipes the hot kernels lean on (POPC for Hamming; I2F / I2FP / FRND /
// F2I conversions and FADD / FMUL for KLT sampling).  Prints warp-instructions per clock per SM.  Used to pin the
// "integer-pipe peak" denominator of the Hamming roofline (SURVEY 8(d) asks for a measured POPC rate on B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <string>
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
// packed FP32 add (Blackwell FADD2): two independent adds per lane and instruction
__device__ __forceinline__ void Fadd2(float &a, float &b, float c, float d) {
    unsigned long long x, y;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(c), "f"(d));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x));
}
#define FADD2_OP(k) Fadd2(f##k, g##k, 1.25f, 0.75f);
__global__ void k_fadd2(unsigned *out, unsigned seed) {
    unsigned a0 = threadIdx.x + seed;
    float f0 = a0 * 0.37f, f1 = a0 * 0.11f, f2 = a0 * 0.23f, f3 = a0 * 0.31f, f4 = a0 * 0.41f, f5 = a0 * 0.43f, f6 = a0 * 0.47f, f7 = a0 * 0.53f;
    float g0 = f0 + 1, g1 = f1 + 1, g2 = f2 + 1, g3 = f3 + 1, g4 = f4 + 1, g5 = f5 + 1, g6 = f6 + 1, g7 = f7 + 1;
    for (int i = 0; i < kIters; ++i) {
        R8(FADD2_OP)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + g0 + g1 + g2 + g3 + g4 + g5 + g6 + g7);
}
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

struct Rate {
    double ms, steps_per_s, per_clk_sm;
};

template <typename K>
Rate measure(K kernel, int sm_count, double clock_hz, unsigned *out) {
    const int blocks = sm_count * 8, threads = 256;
    kernel<<<blocks, threads>>>(out, 1u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {  // best of 5: the first launches may still see the clocks ramping
        cudaEventRecord(e0);
        kernel<<<blocks, threads>>>(out, 2u + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double steps = double(blocks) * threads / 32.0 * kIters * kUnroll;  // warp-level steps
    Rate r;
    r.ms = best;
    r.steps_per_s = steps / (best * 1e-3);
    r.per_clk_sm = r.steps_per_s / clock_hz / sm_count;
    return r;
}

template <typename K>
void run(const char *name, K kernel, int ops_per_step, int sm_count, double clock_hz, unsigned *out) {
    const Rate r = measure(kernel, sm_count, clock_hz, out);
    printf("%-18s %8.3f ms  %7.3f warp-steps/clk/SM  (%d instr per step => %7.3f warp-instr/clk/SM)\n", name, r.ms, r.per_clk_sm, ops_per_step,
           r.per_clk_sm * ops_per_step);
}

// microbench            : the table above (profiles/*_microbench_pipe_rates.txt)
// microbench --json [d] : one JSON line with the FADD / FMUL / POPC issue rates of device d, for bench.py's roofline denominators
int main(int argc, char **argv) {
    const bool json = argc > 1 && std::string(argv[1]) == "--json";
    const int device = (json && argc > 2) ? atoi(argv[2]) : 0;
    if (cudaSetDevice(device) != cudaSuccess) return 1;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, device);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, device);
    const double hz = clock_khz * 1e3;
    unsigned *out;
    cudaMalloc(&out, sizeof(unsigned) * p.multiProcessorCount * 8 * 256);
    const int sms = p.multiProcessorCount;
    if (json) {
        const Rate fadd = measure(k_fadd, sms, hz, out), fmul = measure(k_fmul, sms, hz, out), popc = measure(k_popc_add, sms, hz, out);
        // one flop / one POPC per lane and step; rates in operations per second are clock independent (timed), the per-clock figures
        // assume the maximum clock
        printf("{\"device\": \"%s\", \"sm_count\": %d, \"max_clock_mhz\": %.0f, "
               "\"fadd\": {\"flops_per_s\": %.6e, \"warp_instr_per_clk_sm\": %.4f}, "
               "\"fmul\": {\"flops_per_s\": %.6e, \"warp_instr_per_clk_sm\": %.4f}, "
               "\"popc\": {\"ops_per_s\": %.6e, \"warp_instr_per_clk_sm\": %.4f}}\n",
               p.name, sms, hz / 1e6, fadd.steps_per_s * 32.0, fadd.per_clk_sm, fmul.steps_per_s * 32.0, fmul.per_clk_sm, popc.steps_per_s * 32.0,
               popc.per_clk_sm);
        return 0;
    }
    printf("device %s, %d SMs, max clock %.0f MHz (rates below assume the max clock)\n", p.name, sms, hz / 1e6);
    run("popc+iadd", k_popc_add, 2, sms, hz, out);
    run("iadd", k_iadd, 1, sms, hz, out);
    run("xor+popc+iadd", k_xor_popc_add, 3, sms, hz, out);
    run("i2f.u8+fadd", k_i2f_u8_fadd, 3, sms, hz, out);
    run("i2f.s32+fadd", k_i2f_s32_fadd, 3, sms, hz, out);
    run("magic.u8+fadd", k_magic_u8_fadd, 5, sms, hz, out);
    run("fadd", k_fadd, 1, sms, hz, out);
    run("fmul", k_fmul, 1, sms, hz, out);
    run("fadd2 (f32x2)", k_fadd2, 1, sms, hz, out);
    run("floor+fadd", k_floor_fadd, 2, sms, hz, out);
    run("f2i+iadd+fadd", k_f2i_iadd, 3, sms, hz, out);
    run("lop+shf", k_lop_shf, 2, sms, hz, out);
    run("shfl", k_shfl, 1, sms, hz, out);
    return 0;
}
