// Host staging-copy microbenchmark: threaded memcpy pageable -> pageable and
// pageable -> pinned (cudaHostAlloc), to size the pageable leg of the host pipeline.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
static void ser(char *d, const char *s, size_t rb, size_t r0, size_t r1) { for (size_t r = r0; r < r1; ++r) memcpy(d + r * rb, s + r * rb, rb); }
static double run(char *dst, const char *src, size_t total, size_t chunk, int threads, bool spawn)
{
    const size_t rb = 1u << 20;
    auto t0 = std::chrono::steady_clock::now();
    for (size_t c = 0; c < total / chunk; ++c) {
        char *d = dst + (c % 4) * chunk;
        const char *s = src + c * chunk;
        const size_t nr = chunk / rb;
        std::vector<std::thread> pool;
        if (spawn)
            for (int t = 1; t < threads; ++t) pool.emplace_back(ser, d, s, rb, nr * t / threads, nr * (t + 1) / threads);
        ser(d, s, rb, 0, spawn ? nr / threads : nr);
        for (auto &th : pool) th.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
int main()
{
    const size_t total = 512u << 20, chunk = 32u << 20;
    char *src = (char *)malloc(total), *dst = (char *)malloc(chunk * 4), *pin = nullptr;
    cudaHostAlloc((void **)&pin, chunk * 4, cudaHostAllocDefault);
    memset(src, 1, total);
    memset(dst, 2, chunk * 4);
    memset(pin, 3, chunk * 4);
    for (int threads : {1, 2, 4, 8, 16}) {
        run(dst, src, total, chunk, threads, threads > 1);
        const double a = run(dst, src, total, chunk, threads, threads > 1);
        run(pin, src, total, chunk, threads, threads > 1);
        const double b = run(pin, src, total, chunk, threads, threads > 1);
        printf("{\"threads\": %d, \"pageable_to_pageable_gb_s\": %.1f, \"pageable_to_pinned_gb_s\": %.1f}\n", threads, total / a / 1e9, total / b / 1e9);
    }
    // thread spawn + join cost in this process (CUDA context loaded)
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < 200; ++i) {
        std::thread th([] {});
        th.join();
    }
    printf("{\"thread_spawn_join_us\": %.1f}\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 200 * 1e6);
    return 0;
}
