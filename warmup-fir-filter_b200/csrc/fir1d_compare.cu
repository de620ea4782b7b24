// Fused fixed-vs-ideal comparison: the reduction the reference's compare
// reports perform on two output FILES (fir_1d/sim/vector/
// gen_3tap_compare_report.py:67-112, `_compute_metrics`) computed here in ONE
// pass over the input, so neither the uint8 fixed output nor the 8-byte-per-
// sample float64 ideal output ever touches HBM or disk.
//
// Per sample the kernel evaluates exactly the two models (same arithmetic as
// fir1d_generic.cu: int64 accumulate + wrap/round/shift/saturate, and unfused
// float64 multiply/add in ascending tap order), forms diff = fixed - ideal in
// float64, and reduces:
//   [0] sum |diff|   [1] sum diff^2   [2] sum diff   [3] max |diff|
//   [4] #(fixed == 0)   [5] #(fixed == 255)   [6] #(ideal < 0 or ideal > 255)
// Sums are float64 (block tree + one atomicAdd per block), so they agree with
// NumPy's pairwise float64 sums to ~1e-12 relative, not bit for bit.
#include "common.cuh"

namespace firgpu {
namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <typename InT>
__global__ void __launch_bounds__(kBlock)
fir1d_compare_kernel(const InT *__restrict__ x, int64_t rows, int64_t cols, int64_t ld_x, InlineTaps hq,
                     InlineTapsF64 h, int n_taps, int frac_bits, int acc_bits, double *__restrict__ sums,
                     unsigned long long *__restrict__ counts)
{
    const int centre = n_taps / 2;
    double s_abs = 0.0, s_sq = 0.0, s_diff = 0.0, m_abs = 0.0;
    unsigned long long c_lo = 0, c_hi = 0, c_clip = 0;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const InT *xr = x + r * ld_x;
        for (int64_t n = (int64_t)blockIdx.x * kBlock + threadIdx.x; n < cols; n += (int64_t)gridDim.x * kBlock) {
            long long acc = 0;
            double ideal = 0.0;
            for (int k = 0; k < n_taps; ++k) {
                const int64_t src = n - (k - centre);
                if (src >= 0 && src < cols) {
                    const int s = prep_sample(xr[src]);
                    acc += (long long)s * (long long)hq.v[k];
                    ideal = __dadd_rn(ideal, __dmul_rn(h.v[k], (double)s));
                }
            }
            const int fixed = post_accumulate(acc, frac_bits, acc_bits);
            const double diff = (double)fixed - ideal;
            const double ad = fabs(diff);
            s_abs += ad;
            s_sq += diff * diff;
            s_diff += diff;
            m_abs = fmax(m_abs, ad);
            c_lo += fixed == 0;
            c_hi += fixed == kPixelMax;
            c_clip += (ideal < 0.0) | (ideal > 255.0);
        }
    }
    __shared__ double sh[4][kBlock / 32];
    __shared__ unsigned long long shc[3][kBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s_abs = warp_sum(s_abs);
    s_sq = warp_sum(s_sq);
    s_diff = warp_sum(s_diff);
    m_abs = warp_max(m_abs);
    c_lo = __reduce_add_sync(0xffffffffu, (unsigned)c_lo);         // per-thread counts stay far below 2^32
    c_hi = __reduce_add_sync(0xffffffffu, (unsigned)c_hi);
    c_clip = __reduce_add_sync(0xffffffffu, (unsigned)c_clip);
    if (lane == 0) {
        sh[0][warp] = s_abs;
        sh[1][warp] = s_sq;
        sh[2][warp] = s_diff;
        sh[3][warp] = m_abs;
        shc[0][warp] = c_lo;
        shc[1][warp] = c_hi;
        shc[2][warp] = c_clip;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0, m = 0;
        unsigned long long l = 0, hcount = 0, cl = 0;
        for (int w = 0; w < kBlock / 32; ++w) {
            a += sh[0][w];
            b += sh[1][w];
            c += sh[2][w];
            m = fmax(m, sh[3][w]);
            l += shc[0][w];
            hcount += shc[1][w];
            cl += shc[2][w];
        }
        atomicAdd(&sums[0], a);
        atomicAdd(&sums[1], b);
        atomicAdd(&sums[2], c);
        // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long *>(&sums[3]), (unsigned long long)__double_as_longlong(m));
        atomicAdd(&counts[0], l);
        atomicAdd(&counts[1], hcount);
        atomicAdd(&counts[2], cl);
    }
}

template <typename InT>
int run(const void *x, int64_t rows, int64_t cols, int64_t ld_x, const InlineTaps &hq, const InlineTapsF64 &h, int n_taps,
        int frac_bits, int acc_bits, double *d_sums, unsigned long long *d_counts, cudaStream_t s)
{
    int64_t gx = (cols + kBlock - 1) / kBlock;
    if (gx > 148 * 8) gx = 148 * 8;
    int64_t gy = rows;
    const int64_t cap = (148 * 16 + gx - 1) / gx;
    if (gy > cap) gy = cap;
    if (gy > 65535) gy = 65535;
    fir1d_compare_kernel<InT><<<dim3((unsigned)gx, (unsigned)gy), kBlock, 0, s>>>((const InT *)x, rows, cols, ld_x, hq, h, n_taps,
                                                                                frac_bits, acc_bits, d_sums, d_counts);
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_GENERIC);
    return 0;
}

}  // namespace

// x: DEVICE pointer.  out[7] (HOST): the seven reductions listed at the top.
// Synchronises the stream before returning.
int launch_fir1d_compare(const void *x, int in_dtype, int64_t rows, int64_t cols, int64_t ld_x, const int32_t *hq_host,
                         const double *h_host, int n_taps, int frac_bits, int acc_bits, double *out, cudaStream_t s)
{
    if (n_taps > kInlineTaps) return set_error(FIRGPU_E_TAPS, "firgpu_fir1d_compare: n_taps=%d exceeds %d", n_taps, kInlineTaps);
    InlineTaps hq{};
    InlineTapsF64 h{};
    for (int k = 0; k < n_taps; ++k) {
        hq.v[k] = hq_host[k];
        h.v[k] = h_host[k];
    }
    double *d_sums = nullptr;
    FIRGPU_CUDA(cudaMallocAsync((void **)&d_sums, 8 * sizeof(double), s));
    FIRGPU_CUDA(cudaMemsetAsync(d_sums, 0, 8 * sizeof(double), s));
    unsigned long long *d_counts = reinterpret_cast<unsigned long long *>(d_sums + 4);
    int rc;
    switch (in_dtype) {
    case FIRGPU_U8: rc = run<uint8_t>(x, rows, cols, ld_x, hq, h, n_taps, frac_bits, acc_bits, d_sums, d_counts, s); break;
    case FIRGPU_I16: rc = run<int16_t>(x, rows, cols, ld_x, hq, h, n_taps, frac_bits, acc_bits, d_sums, d_counts, s); break;
    case FIRGPU_I32: rc = run<int32_t>(x, rows, cols, ld_x, hq, h, n_taps, frac_bits, acc_bits, d_sums, d_counts, s); break;
    case FIRGPU_F32: rc = run<float>(x, rows, cols, ld_x, hq, h, n_taps, frac_bits, acc_bits, d_sums, d_counts, s); break;
    case FIRGPU_F64: rc = run<double>(x, rows, cols, ld_x, hq, h, n_taps, frac_bits, acc_bits, d_sums, d_counts, s); break;
    default: rc = set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_compare: unsupported in_dtype %d", in_dtype);
    }
    if (rc) {
        cudaFreeAsync(d_sums, s);
        return rc;
    }
    double host[8];
    FIRGPU_CUDA(cudaMemcpyAsync(host, d_sums, sizeof(host), cudaMemcpyDeviceToHost, s));
    FIRGPU_CUDA(cudaFreeAsync(d_sums, s));
    FIRGPU_CUDA(cudaStreamSynchronize(s));
    unsigned long long cnt[3];
    memcpy(cnt, &host[4], sizeof(cnt));
    out[0] = host[0];
    out[1] = host[1];
    out[2] = host[2];
    out[3] = host[3];
    out[4] = (double)cnt[0];
    out[5] = (double)cnt[1];
    out[6] = (double)cnt[2];
    return 0;
}

}  // namespace firgpu
