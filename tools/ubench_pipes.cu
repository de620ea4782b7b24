// Instruction-throughput microbenchmark for the integer datapath choices of the
// FIR kernels (sm_100a): dp2a (IDP.2A), IMAD, FFMA, packed min/max (VIMNMX),
// PRMT, and mixes — lanes per clock per SM.  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ubench_pipes tools/ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kChains = 8;

__device__ __forceinline__ int dp2a(uint32_t a, uint32_t b, int c)
{
    int d;
    asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int MODE>
__global__ void bench(int *out, uint32_t a, uint32_t b, float fa)
{
    int acc[kChains], acc2[kChains];
    float facc[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        acc[i] = threadIdx.x + i;
        acc2[i] = threadIdx.x * 3 + i;
        facc[i] = (float)(threadIdx.x + i);
    }
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (MODE == 0) acc[i] = dp2a(a, b + i, acc[i]);
            if (MODE == 1) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b + i));
            if (MODE == 2) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(facc[i]) : "f"(fa), "f"(fa));
            if (MODE == 3) asm volatile("vmin2.s32.s32.s32 %0, %0, %1, %0;" : "+r"(acc[i]) : "r"(a));   // packed min (may be emulated)
            if (MODE == 4) asm volatile("prmt.b32 %0, %0, %1, 0x6420;" : "+r"(acc[i]) : "r"(a));
            if (MODE == 5) {                                   // dp2a + fma interleaved 1:1
                acc[i] = dp2a(a, b + i, acc[i]);
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(facc[i]) : "f"(fa), "f"(fa));
            }
            if (MODE == 6) {                                   // dp2a + imad 1:1
                acc[i] = dp2a(a, b + i, acc[i]);
                asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc2[i]) : "r"(a), "r"(b + i));
            }
            if (MODE == 7) {                                   // dp2a + prmt 1:1 (fma pipe + alu pipe)
                acc[i] = dp2a(a, b + i, acc[i]);
                asm volatile("prmt.b32 %0, %0, %1, 0x6420;" : "+r"(acc2[i]) : "r"(a));
            }
            if (MODE == 8) acc[i] = __vimin_s16x2_relu(acc[i], a);
            if (MODE == 9) {                                   // imad + fma 1:1
                asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b + i));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(facc[i]) : "f"(fa), "f"(fa));
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += acc[i] + acc2[i] + (int)facc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int ops_per_iter, int sms, int *out, double clock_hz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 4, threads = 512;
    bench<MODE><<<blocks, threads>>>(out, 0x00030002u, 0x01020304u, 1.0001f);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, threads>>>(out, 0x00030002u, 0x01020304u, 1.0001f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * kIters * kChains * ops_per_iter;
    printf("{\"op\": \"%s\", \"ms\": %.3f, \"Tops_per_s\": %.2f, \"lanes_per_clk_per_sm_at_%.0fMHz\": %.1f}\n", name, ms,
           ops / (ms * 1e-3) / 1e12, clock_hz / 1e6, ops / (ms * 1e-3) / sms / clock_hz);
}

int main()
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double hz = khz * 1e3;
    int *out;
    cudaMalloc(&out, sizeof(int) * sms * 4 * 512);
    run<0>("dp2a", 1, sms, out, hz);
    run<1>("imad", 1, sms, out, hz);
    run<2>("ffma", 1, sms, out, hz);
    run<3>("vmin2 (ptx)", 1, sms, out, hz);
    run<8>("vimin_s16x2_relu", 1, sms, out, hz);
    run<4>("prmt", 1, sms, out, hz);
    run<5>("dp2a+ffma", 2, sms, out, hz);
    run<6>("dp2a+imad", 2, sms, out, hz);
    run<7>("dp2a+prmt", 2, sms, out, hz);
    run<9>("imad+ffma", 2, sms, out, hz);
    return 0;
}
