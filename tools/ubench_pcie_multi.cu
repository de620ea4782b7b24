// Host <-> device copy ceiling of ONE box with 1, 2, 4, 8 GPUs copying at once, both directions together: what bounds
// the end-to-end (host buffers in, host buffers out) figure of bench.py when several ranks share the host's memory
// and PCIe root complexes.  One host thread per GPU, pinned buffers, 32 MiB chunks on two streams per GPU.
//   nvcc -O2 -std=c++17 -o tools/_bin/ubench_pcie_multi tools/ubench_pcie_multi.cu -lpthread
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

static std::atomic<int> g_ready{0}, g_go{0};

struct Result { double seconds = 0; int err = 0; };

static void worker(int dev, size_t total, int passes, int n_workers, Result *res)
{
    auto fail = [&](cudaError_t e) { if (e != cudaSuccess && !res->err) { res->err = (int)e; fprintf(stderr, "dev %d: %s\n", dev, cudaGetErrorString(e)); } };
    fail(cudaSetDevice(dev));
    char *hin = nullptr, *hout = nullptr, *din = nullptr, *dout = nullptr;
    fail(cudaHostAlloc((void **)&hin, total, cudaHostAllocDefault));
    fail(cudaHostAlloc((void **)&hout, total, cudaHostAllocDefault));
    fail(cudaMalloc((void **)&din, total));
    fail(cudaMalloc((void **)&dout, total));
    if (res->err) { g_ready++; return; }
    for (size_t i = 0; i < total; i += 4096) hin[i] = (char)i, hout[i] = 0;
    cudaStream_t s1, s2;
    fail(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    fail(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    const size_t chunk = 32ull << 20;
    for (size_t off = 0; off < total; off += chunk) {          // warm-up
        fail(cudaMemcpyAsync(din + off, hin + off, chunk, cudaMemcpyHostToDevice, s1));
        fail(cudaMemcpyAsync(hout + off, dout + off, chunk, cudaMemcpyDeviceToHost, s2));
    }
    fail(cudaDeviceSynchronize());
    g_ready++;
    while (g_go.load() == 0) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    for (int p = 0; p < passes; ++p)
        for (size_t off = 0; off < total; off += chunk) {
            fail(cudaMemcpyAsync(din + off, hin + off, chunk, cudaMemcpyHostToDevice, s1));
            fail(cudaMemcpyAsync(hout + off, dout + off, chunk, cudaMemcpyDeviceToHost, s2));
        }
    fail(cudaStreamSynchronize(s1));
    fail(cudaStreamSynchronize(s2));
    res->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    cudaFreeHost(hin);
    cudaFreeHost(hout);
    cudaFree(din);
    cudaFree(dout);
    (void)n_workers;
}

int main()
{
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) { printf("no devices\n"); return 1; }
    const size_t total = 512ull << 20;
    const int passes = 6;
    for (int n : {1, 2, 4, 8}) {
        if (n > n_dev) break;
        g_ready = 0;
        g_go = 0;
        std::vector<Result> res(n);
        std::vector<std::thread> pool;
        for (int d = 0; d < n; ++d) pool.emplace_back(worker, d, total, passes, n, &res[d]);
        while (g_ready.load() < n) std::this_thread::yield();
        g_go = 1;
        for (auto &t : pool) t.join();
        double slowest = 0;
        int err = 0;
        for (auto &r : res) { slowest = r.seconds > slowest ? r.seconds : slowest; err |= r.err; }
        const double per_dir = (double)n * passes * total / slowest / 1e9;
        printf("{\"gpus\": %d, \"mode\": \"h2d + d2h together, pinned, 32 MiB chunks\", \"seconds\": %.4f, \"gb_s_per_direction_total\": %.1f, "
               "\"gb_s_per_direction_per_gpu\": %.1f, \"err\": %d}\n", n, slowest, per_dir, per_dir / n, err);
        fflush(stdout);
    }
    return 0;
}
