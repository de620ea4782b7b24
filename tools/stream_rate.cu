// Block rate of the device-resident streaming objects from a C caller (no Python per-call cost): a 2^28-sample int16
// signal in blocks of 2^16 ... 2^28 samples through firgpu_stream_process, against one causal call.
//   nvcc -O2 -std=c++17 -I include -o tools/_bin/stream_rate tools/stream_rate.cu -L warmup-fir-filter_b200/firgpu -lfirgpu -Xlinker -rpath -Xlinker '$ORIGIN/../../warmup-fir-filter_b200/firgpu'
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "firgpu.h"

#define CK(x) do { int rc__ = (x); if (rc__) { printf("%s -> %d (%s)\n", #x, rc__, firgpu_last_error()); return 1; } } while (0)

int main()
{
    const int64_t n = 1ll << 28;
    int16_t *x, *y;
    if (cudaMalloc(&x, n * 2) || cudaMalloc(&y, n * 2)) return 1;
    cudaMemset(x, 1, n * 2);
    cudaStream_t st;
    cudaStreamCreate(&st);
    const int32_t hq[5] = {2048, 8192, 12288, 8192, 2048};          // simple_lp, Q1.15
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int w = 0; w < 3; ++w) CK(firgpu_fir1d_fixed_ex(x, y, FIRGPU_I16, FIRGPU_I16, 1, n, n, n, hq, 5, 1, 15, 32, 0, 0, 0, 0, 255, 0, 255, st));
    cudaEventRecord(a, st);
    for (int r = 0; r < 10; ++r) CK(firgpu_fir1d_fixed_ex(x, y, FIRGPU_I16, FIRGPU_I16, 1, n, n, n, hq, 5, 1, 15, 32, 0, 0, 0, 0, 255, 0, 255, st));
    cudaEventRecord(b, st);
    cudaStreamSynchronize(st);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("{\"case\": \"one causal call, 2^28 int16 samples\", \"ms\": %.4f, \"gsamples_s\": %.1f}\n", ms / 10, n / (ms / 10 * 1e-3) / 1e9);
    for (int lg : {16, 18, 20, 22, 24, 26, 28}) {
        const int64_t blk = 1ll << lg;
        firgpu_stream_t s;
        CK(firgpu_stream_create_fixed(&s, FIRGPU_I16, FIRGPU_I16, hq, 5, 1, 15, 32, 0, 255, 0, 255, -1));
        for (int64_t o = 0; o < n; o += blk) CK(firgpu_stream_process(s, x + o, y + o, blk, st));       // warm-up pass
        cudaStreamSynchronize(st);
        const int reps = lg < 20 ? 1 : 3;
        const auto t0 = std::chrono::steady_clock::now();
        cudaEventRecord(a, st);
        for (int r = 0; r < reps; ++r)
            for (int64_t o = 0; o < n; o += blk) CK(firgpu_stream_process(s, x + o, y + o, blk, st));
        cudaEventRecord(b, st);
        const double host_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        cudaStreamSynchronize(st);
        cudaEventElapsedTime(&ms, a, b);
        printf("{\"case\": \"stream object, blocks of 2^%d samples\", \"ms_per_2p28\": %.4f, \"gsamples_s\": %.1f, \"us_per_block_device\": %.2f, "
               "\"us_per_block_host_submit\": %.2f, \"launches_per_block\": %.2f}\n",
               lg, ms / reps, n / (ms / reps * 1e-3) / 1e9, ms * 1e3 / reps / (double)(n / blk), host_s * 1e6 / reps / (double)(n / blk),
               0.0);
        firgpu_stream_destroy(s);
    }
    return 0;
}
