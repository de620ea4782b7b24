// Tuning harness for the tiled 2-D kernel (not part of the library): one build
// per tunable set (-DFIRGPU_T2_THREADS/STAGES/RPS5/RPS3/MINBLOCKS), times the
// config-4 shapes (16384^2 int16) for 1..occupancy CTAs per SM.
#include "fir2d_tiled.cu"

#include <cstdio>

#define FIRGPU_TILED_T int16_t
#define FIRGPU_TILED_NAME launch_tiled_i16
#include "fir2d_tiled_inst.cu"

namespace firgpu {
int set_error(int code, const char *, ...) { return code; }
int check_cuda(cudaError_t e, const char *what)
{
    if (e != cudaSuccess) fprintf(stderr, "CUDA error %s in %s\n", cudaGetErrorString(e), what);
    return (int)e;
}
void note_launch(int) {}
namespace tiled {
int launch_tiled_u8(const CUtensorMap &, const TiledParams &, int, int, int, bool, cudaStream_t) { return 0; }
int launch_tiled_i16u8(const CUtensorMap &, const TiledParams &, int, int, int, bool, cudaStream_t) { return 0; }
}
}  // namespace firgpu

using namespace firgpu;

__global__ void fill_kernel(int16_t *x, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] = (int16_t)((i * 2654435761u >> 7) % 300 - 20);
}

int main(int argc, char **argv)
{
    const int reps = argc > 1 ? atoi(argv[1]) : 10;
    const long long n = 16384;
    int16_t *x, *y;
    cudaMalloc(&x, n * n * 2);
    cudaMalloc(&y, n * n * 2);
    fill_kernel<<<148 * 8, 256>>>(x, n * n);
    cudaDeviceSynchronize();
    struct Case { const char *name; int kh, kw; int32_t taps[25]; };
    Case cases[3] = {{"3x3 box (general)", 3, 3, {455, 455, 455, 455, 455, 455, 455, 455, 455}},
                     {"5x5 lp (rank-1)", 5, 5, {}},
                     {"5x5 non-separable", 5, 5, {}}};
    const int lp[5] = {1, 4, 6, 4, 1};
    for (int m = 0; m < 5; ++m)
        for (int k = 0; k < 5; ++k) {
            cases[1].taps[m * 5 + k] = 16 * lp[m] * lp[k];
            cases[2].taps[m * 5 + k] = 16 * lp[m] * lp[k] + ((m * 5 + k) % 3 == 0 ? 7 : -3);
        }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (auto &c : cases) {
        for (int ctas = 1; ctas <= 6; ++ctas) {
            char buf[8];
            snprintf(buf, sizeof buf, "%d", ctas);
            setenv("FIRGPU_CTAS_PER_SM", buf, 1);
            Fir2dArgs a{x, y, FIRGPU_I16, FIRGPU_I16, 1, n, n, n, n, n * n, n * n, c.kh, c.kw, 12, 32, 0, 0, 0};
            int rc = 0;
            for (int w = 0; w < 2; ++w) rc = try_launch_fir2d_fixed_tiled(a, c.taps);
            cudaEventRecord(e0);
            for (int r = 0; r < reps; ++r) try_launch_fir2d_fixed_tiled(a, c.taps);
            cudaEventRecord(e1);
            cudaError_t err = cudaDeviceSynchronize();
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            ms /= reps;
            printf("{\"case\": \"%s\", \"threads\": %d, \"stages\": %d, \"rps5\": %d, \"rps3\": %d, \"minblocks\": %d, \"warpspec\": %d, \"ctas_per_sm_cap\": %d, "
                   "\"rc\": %d, \"ms\": %.4f, \"gb_s\": %.1f, \"err\": \"%s\"}\n",
                   c.name, FIRGPU_T2_THREADS, FIRGPU_T2_STAGES, FIRGPU_T2_RPS5, FIRGPU_T2_RPS3, FIRGPU_T2_MINBLOCKS, FIRGPU_T2_WARPSPEC, ctas, rc, ms,
                   n * n * 4.0 / (ms * 1e-3) / 1e9, cudaGetErrorString(err));
            fflush(stdout);
        }
    }
    return 0;
}
