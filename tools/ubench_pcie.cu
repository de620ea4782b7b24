// PCIe ceiling for the host pipeline: pinned host <-> device copies, one direction at a time and both together
// (two streams), for several chunk sizes.  Build: nvcc -O2 -o tools/_bin/ubench_pcie tools/ubench_pcie.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main()
{
    const size_t total = 1ull << 30;
    char *hin, *hout, *din, *dout;
    CK(cudaHostAlloc((void **)&hin, total, cudaHostAllocDefault));
    CK(cudaHostAlloc((void **)&hout, total, cudaHostAllocDefault));
    CK(cudaMalloc((void **)&din, total));
    CK(cudaMalloc((void **)&dout, total));
    for (size_t i = 0; i < total; i += 4096) hin[i] = (char)i, hout[i] = 0;
    cudaStream_t s1, s2;
    CK(cudaStreamCreate(&s1));
    CK(cudaStreamCreate(&s2));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    // misaligned H2D (source 4 bytes before a 32 MiB boundary, destination 12 bytes into its buffer): what a segment
    // copy that starts at its 2-sample halo looks like
    for (int mis = 0; mis < 2; ++mis) {
        const size_t chunk = 32ull << 20;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(a, 0));
            CK(cudaStreamWaitEvent(s1, a, 0));
            CK(cudaStreamWaitEvent(s2, a, 0));
            for (size_t off = chunk; off + chunk < total; off += chunk) {
                CK(cudaMemcpyAsync(din + off + (mis ? 12 : 0), hin + off - (mis ? 4 : 0), chunk - 256, cudaMemcpyHostToDevice, s1));
                CK(cudaMemcpyAsync(hout + off, dout + off, chunk, cudaMemcpyDeviceToHost, s2));
            }
            CK(cudaEventRecord(b, s1));
            CK(cudaStreamWaitEvent(0, b, 0));
            CK(cudaEventRecord(b, s2));
            CK(cudaStreamWaitEvent(0, b, 0));
            CK(cudaEventRecord(b, 0));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            if (ms < best) best = ms;
        }
        printf("{\"chunk_mib\": 32, \"mode\": \"both, h2d %s\", \"ms\": %.3f, \"gb_s_per_direction\": %.2f}\n",
               mis ? "misaligned" : "aligned", best, (total - 2 * chunk) / (best * 1e-3) / 1e9);
    }
    for (size_t chunk : {32ull << 20}) {
        for (int mode = 0; mode < 3; ++mode) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(a, 0));
                CK(cudaStreamWaitEvent(s1, a, 0));
                CK(cudaStreamWaitEvent(s2, a, 0));
                for (size_t off = 0; off < total; off += chunk) {
                    if (mode != 1) CK(cudaMemcpyAsync(din + off, hin + off, chunk, cudaMemcpyHostToDevice, s1));
                    if (mode != 0) CK(cudaMemcpyAsync(hout + off, dout + off, chunk, cudaMemcpyDeviceToHost, s2));
                }
                CK(cudaEventRecord(b, s1));
                CK(cudaStreamWaitEvent(0, b, 0));
                CK(cudaEventRecord(b, s2));
                CK(cudaStreamWaitEvent(0, b, 0));
                CK(cudaEventRecord(b, 0));
                CK(cudaDeviceSynchronize());
                float ms;
                CK(cudaEventElapsedTime(&ms, a, b));
                if (ms < best) best = ms;
            }
            printf("{\"chunk_mib\": %zu, \"mode\": \"%s\", \"ms\": %.3f, \"gb_s_per_direction\": %.2f}\n", chunk >> 20,
                   mode == 0 ? "h2d" : mode == 1 ? "d2h" : "both", best, total / (best * 1e-3) / 1e9);
        }
    }
    return 0;
}
