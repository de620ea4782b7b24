// Tuning harness for the staged 1-D kernel (not part of the library): builds
// ONE configuration of the tunables (-DFIRGPU_K_THREADS/SUB/STAGES/MINBLOCKS,
// -DFIRGPU_STORE_CS) and times the headline shapes over a range of grid sizes,
// next to a plain vectorised copy and cudaMemcpyAsync as HBM references.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I warmup-fir-filter_b200/csrc ... tools/sweep_staged.cu
#include "fir1d_staged.cuh"

#include <cstdio>
#include <vector>

namespace firgpu {
int set_error(int code, const char *, ...) { return code; }
int check_cuda(cudaError_t e, const char *what)
{
    if (e != cudaSuccess) fprintf(stderr, "CUDA error %s in %s\n", cudaGetErrorString(e), what);
    return (int)e;
}
void note_launch(int) {}
}  // namespace firgpu

using namespace firgpu;
using namespace firgpu::staged;

__global__ void fill_kernel(int16_t *x, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] = (int16_t)((i * 2654435761u >> 7) % 300 - 20);
}

__global__ void store_kernel(uint4 *__restrict__ b, size_t n)
{
    const uint4 v = make_uint4(1u, 2u, 3u, 4u);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = v;
}

__global__ void load_kernel(const uint4 *__restrict__ a, uint4 *__restrict__ sink, size_t n)
{
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = a[i];
        acc.x ^= v.x;
        acc.y ^= v.y;
        acc.z ^= v.z;
        acc.w ^= v.w;
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[0] = acc;      // never true in practice; keeps the loads alive
}

__global__ void copy_kernel(const uint4 *__restrict__ a, uint4 *__restrict__ b, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

static uint32_t pack_pair(int a, int b) { return ((uint32_t)a & 0xffffu) | ((uint32_t)b << 16); }
static int tap_at(const int *t, int n, int d)
{
    const int hi = n / 2, lo = n - 1 - hi;
    return (d < -lo || d > hi) ? 0 : t[hi - d];
}

template <typename InT, typename OutT>
static void run_case(const char *name, const void *x, void *y, size_t n_scalars, int bytes_per_sample, int sms, int reps)
{
    using C = typename TuneFor<InT, OutT>::type;
    constexpr int kThreads = C::kThreads, kSub = C::kSub, kStages = C::kStages, kSmemBytes = C::kSmemBytes, kTileBytes = C::kTileBytes;
    auto kern = fir1d_staged_kernel<C, InT, OutT, 5, 1, true, 0, 1>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, kSmemBytes);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    StagedParams p{};
    p.x = (const unsigned char *)x;
    p.y = (unsigned char *)y;
    p.rows = 1;
    p.n_filters = 1;
    p.cols = (long long)n_scalars;
    p.ld_x = p.ld_y = p.cols;
    const long long ts = kTileBytes / (int)sizeof(InT);
    p.tiles_per_row = (p.cols + ts - 1) / ts;
    p.total_tiles = p.tiles_per_row;
    p.shift = 16;
    p.bias = 32768;
    p.x_bulk_ok = 1;
    p.y_vec_ok = 1;
    const int taps[5] = {4096, 16384, 24576, 16384, 4096};      // simple_lp Q1.15, pre-scaled by 2
    for (int k = 0; k < kMaxPairs; ++k) {
        p.cp_even[0][k] = pack_pair(tap_at(taps, 5, -2 + 2 * k), tap_at(taps, 5, -2 + 2 * k + 1));
        p.cp_odd[0][k] = pack_pair(tap_at(taps, 5, -3 + 2 * k), tap_at(taps, 5, -2 + 2 * k));
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    std::vector<int> grids;
    for (int c = 1; c <= occ; ++c) grids.push_back(sms * c);
    for (int g : grids) {
        for (int w = 0; w < 3; ++w) kern<<<g, kThreads, kSmemBytes>>>(p);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) kern<<<g, kThreads, kSmemBytes>>>(p);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= reps;
        printf("{\"case\": \"%s\", \"threads\": %d, \"sub\": %d, \"stages\": %d, \"minblocks\": %d, \"store_cs\": %d, \"regs\": %d, \"occ\": %d, "
               "\"ctas_per_sm\": %d, \"ms\": %.4f, \"gb_s\": %.1f, \"err\": \"%s\"}\n",
               name, kThreads, kSub, kStages, FIRGPU_K_MINBLOCKS, FIRGPU_STORE_CS, fa.numRegs, occ, g / sms, ms,
               n_scalars * (double)bytes_per_sample / (ms * 1e-3) / 1e9, cudaGetErrorString(err));
        fflush(stdout);
    }
}

int main(int argc, char **argv)
{
    const size_t n = (size_t)1 << 28;
    const int reps = argc > 1 ? atoi(argv[1]) : 20;
    const bool refs = argc > 2;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int16_t *x, *y;
    cudaMalloc(&x, n * 2);
    cudaMalloc(&y, n * 2);
    fill_kernel<<<sms * 8, 256>>>(x, n);
    cudaDeviceSynchronize();
    if (refs) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int g : {sms * 2, sms * 4, sms * 8, sms * 16, sms * 32}) {
            for (int w = 0; w < 3; ++w) copy_kernel<<<g, 256>>>((const uint4 *)x, (uint4 *)y, n / 8);
            cudaEventRecord(e0);
            for (int r = 0; r < reps; ++r) copy_kernel<<<g, 256>>>((const uint4 *)x, (uint4 *)y, n / 8);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            ms /= reps;
            printf("{\"case\": \"copy_kernel uint4 grid-stride\", \"grid\": %d, \"ms\": %.4f, \"gb_s\": %.1f}\n", g, ms, n * 4.0 / (ms * 1e-3) / 1e9);
        }
        for (int g : {sms * 8, sms * 32}) {
            cudaEventRecord(e0);
            for (int r = 0; r < reps; ++r) store_kernel<<<g, 256>>>((uint4 *)y, n / 8);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            ms /= reps;
            printf("{\"case\": \"store-only kernel uint4\", \"grid\": %d, \"ms\": %.4f, \"gb_s\": %.1f}\n", g, ms, n * 2.0 / (ms * 1e-3) / 1e9);
            cudaEventRecord(e0);
            for (int r = 0; r < reps; ++r) load_kernel<<<g, 256>>>((const uint4 *)x, (uint4 *)y, n / 8);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            cudaEventElapsedTime(&ms, e0, e1);
            ms /= reps;
            printf("{\"case\": \"load-only kernel uint4\", \"grid\": %d, \"ms\": %.4f, \"gb_s\": %.1f}\n", g, ms, n * 2.0 / (ms * 1e-3) / 1e9);
        }
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) cudaMemcpyAsync(y, x, n * 2, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= reps;
        printf("{\"case\": \"cudaMemcpyAsync D2D\", \"ms\": %.4f, \"gb_s\": %.1f}\n", ms, n * 4.0 / (ms * 1e-3) / 1e9);
    }
    run_case<int16_t, int16_t>("i16->i16 5tap 2^28", x, y, n, 4, sms, reps);
    run_case<int16_t, uint8_t>("i16->u8 5tap 2^28", x, y, n, 3, sms, reps);
    run_case<uint8_t, uint8_t>("u8->u8 5tap 2^29", x, y, n * 2, 2, sms, reps);
    return 0;
}
