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
#include "fastplan.h"
#include "ideal_common.cuh"

#include <cstdlib>

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

__device__ __forceinline__ void block_reduce_and_publish(double s_abs, double s_sq, double s_diff, double m_abs, unsigned long long c_lo,
                                                         unsigned long long c_hi, unsigned long long c_clip,
                                                         double *__restrict__ sums, unsigned long long *__restrict__ counts)
{
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

// Register form for uint8 / int16 pixels and up to nine taps whose fixed-point sum provably fits int32
// (csrc/fastplan.h): the construction of fir1d_ideal_fast.cu — a lane takes two adjacent outputs from the few sample
// pairs its window covers (global loads, L1-backed), converts each sample once to float64 with the exact 2^52 + v bit
// trick, evaluates the float64 model (unfused, ascending taps) and the integer model (int32 multiply-adds, the
// host-proved no-wrap shortcut of the fast kernels), and folds the difference into the seven reductions.
// Bound: the fp64 pipe (about 22 float64 operations per sample), not HBM (1-2 B per sample).
template <typename InT, int L>
__global__ void __launch_bounds__(kBlock)
fir1d_compare_fast_kernel(const InT *__restrict__ x, int64_t rows, int64_t cols, int64_t ld_x, InlineTaps hq, InlineTapsF64 h,
                          int frac_bits, int bias, int64_t tiles_per_row, int64_t total_tiles, double *__restrict__ sums,
                          unsigned long long *__restrict__ counts)
{
    constexpr int kWarps = kBlock / 32, kGroups = 4, kTile = 64 * kGroups;
    constexpr int hi = L / 2, lo = L - 1 - hi;
    constexpr int PL = (lo + 1) / 2, PR = (hi + 1) / 2, NP = PL + 1 + PR;
    using P2 = typename PairOf<InT>::type;
    const int lane = threadIdx.x & 31;
    double s_abs = 0.0, s_sq = 0.0, s_diff = 0.0, m_abs = 0.0;
    unsigned c_lo = 0, c_hi = 0, c_clip = 0;

    auto fold = [&](const int (&v)[2 * NP], const double (&vd)[2 * NP], int e) {   // output base + e from the window samples
        int acc = 0;
        double ideal = 0.0;
#pragma unroll
        for (int k = 0; k < L; ++k) {
            acc += hq.v[k] * v[2 * PL + hi - k + e];
            ideal = __dadd_rn(ideal, __dmul_rn(h.v[k], vd[2 * PL + hi - k + e]));
        }
        const int fixed = min(max((acc + bias) >> frac_bits, 0), kPixelMax);
        const double diff = pixel_to<double>(fixed) - ideal;
        const double ad = fabs(diff);
        s_abs += ad;
        s_sq = fma(diff, diff, s_sq);
        s_diff += diff;
        m_abs = fmax(m_abs, ad);
        c_lo += fixed == 0;
        c_hi += fixed == kPixelMax;
        c_clip += (ideal < 0.0) | (ideal > 255.0);
    };

    const int64_t step = (int64_t)gridDim.x * kWarps;
    const int64_t step_rows = step / tiles_per_row, step_cols = step - step_rows * tiles_per_row;
    int64_t tile = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    int64_t row = tile / tiles_per_row, tcol = tile - row * tiles_per_row;
    for (; tile < total_tiles; tile += step) {
        const int64_t col0 = tcol * kTile;
        const InT *xr = x + row * ld_x;
        const bool x_pairs = (reinterpret_cast<uintptr_t>(xr) & (2 * sizeof(InT) - 1)) == 0;
        const bool interior = col0 - 2 * PL >= 0 && col0 + kTile + 2 * PR <= cols && x_pairs;
#pragma unroll
        for (int j = 0; j < kGroups; ++j) {
            const int64_t n = col0 + j * 64 + 2 * lane;
            if (n >= cols) break;
            int v[2 * NP];
            if (interior) {
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const P2 r = *reinterpret_cast<const P2 *>(xr + n + 2 * (p - PL));
                    if constexpr (sizeof(InT) == 1) {
                        v[2 * p] = r & 0xff;
                        v[2 * p + 1] = r >> 8;
                    } else {
                        v[2 * p] = prep_sample((int16_t)(r & 0xffffu));
                        v[2 * p + 1] = prep_sample((int16_t)(r >> 16));
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 2 * NP; ++i) {
                    const int64_t g = n - 2 * PL + i;
                    v[i] = (g >= 0 && g < cols) ? prep_sample(xr[g]) : 0;
                }
            }
            double vd[2 * NP];                               // each sample converted once for both outputs
#pragma unroll
            for (int i = 0; i < 2 * NP; ++i) vd[i] = pixel_to<double>(v[i]);
            fold(v, vd, 0);
            if (n + 1 < cols) fold(v, vd, 1);
        }
        row += step_rows;
        tcol += step_cols;
        if (tcol >= tiles_per_row) {
            tcol -= tiles_per_row;
            ++row;
        }
    }
    block_reduce_and_publish(s_abs, s_sq, s_diff, m_abs, c_lo, c_hi, c_clip, sums, counts);
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
    block_reduce_and_publish(s_abs, s_sq, s_diff, m_abs, c_lo, c_hi, c_clip, sums, counts);
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

template <typename InT, int L>
int run_fast(const void *x, int64_t rows, int64_t cols, int64_t ld_x, const InlineTaps &hq, const InlineTapsF64 &h, int frac_bits,
             int bias, double *d_sums, unsigned long long *d_counts, cudaStream_t s)
{
    constexpr int kTile = 64 * 4, kWarps = kBlock / 32;
    const int64_t tiles_per_row = (cols + kTile - 1) / kTile, total = tiles_per_row * rows;
    int64_t grid = (total + kWarps - 1) / kWarps;
    if (grid > 148 * 8) grid = 148 * 8;                      // persistent: per-thread partial sums, one publish per CTA
    fir1d_compare_fast_kernel<InT, L><<<(unsigned)grid, kBlock, 0, s>>>((const InT *)x, rows, cols, ld_x, hq, h, frac_bits, bias,
                                                                       tiles_per_row, total, d_sums, d_counts);
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_STAGED);
    return 0;
}

template <typename InT>
int run_fast_taps(int n_taps, const void *x, int64_t rows, int64_t cols, int64_t ld_x, const InlineTaps &hq, const InlineTapsF64 &h,
                  int frac_bits, int bias, double *d_sums, unsigned long long *d_counts, cudaStream_t s)
{
    switch (n_taps) {
    case 1: return run_fast<InT, 1>(x, rows, cols, ld_x, hq, h, frac_bits, bias, d_sums, d_counts, s);
    case 2: return run_fast<InT, 2>(x, rows, cols, ld_x, hq, h, frac_bits, bias, d_sums, d_counts, s);
    case 3: return run_fast<InT, 3>(x, rows, cols, ld_x, hq, h, frac_bits, bias, d_sums, d_counts, s);
    case 4: return run_fast<InT, 4>(x, rows, cols, ld_x, hq, h, frac_bits, bias, d_sums, d_counts, s);
    case 5: return run_fast<InT, 5>(x, rows, cols, ld_x, hq, h, frac_bits, bias, d_sums, d_counts, s);
    case 6: return run_fast<InT, 6>(x, rows, cols, ld_x, hq, h, frac_bits, bias, d_sums, d_counts, s);
    case 7: return run_fast<InT, 7>(x, rows, cols, ld_x, hq, h, frac_bits, bias, d_sums, d_counts, s);
    case 8: return run_fast<InT, 8>(x, rows, cols, ld_x, hq, h, frac_bits, bias, d_sums, d_counts, s);
    default: return run_fast<InT, 9>(x, rows, cols, ld_x, hq, h, frac_bits, bias, d_sums, d_counts, s);
    }
}

}  // namespace

// x: DEVICE pointer.  d_sums: DEVICE accumulator of 8 doubles (4 float64 sums / max + 3 uint64 counters + pad),
// zeroed by the caller; the kernels ADD to it, so several pieces of one case may be accumulated (on any streams).
int launch_fir1d_compare_into(const void *x, int in_dtype, int64_t rows, int64_t cols, int64_t ld_x, const int32_t *hq_host,
                              const double *h_host, int n_taps, int frac_bits, int acc_bits, double *d_sums, cudaStream_t s)
{
    if (n_taps > kInlineTaps) return set_error(FIRGPU_E_TAPS, "firgpu_fir1d_compare: n_taps=%d exceeds %d", n_taps, kInlineTaps);
    InlineTaps hq{};
    InlineTapsF64 h{};
    for (int k = 0; k < n_taps; ++k) {
        hq.v[k] = hq_host[k];
        h.v[k] = h_host[k];
    }
    unsigned long long *d_counts = reinterpret_cast<unsigned long long *>(d_sums + 4);
    // int32 shortcut: the wrap is provably the identity and acc + bias fits int32 (same proof as the fast filter kernels)
    const FastPlan plan = n_taps <= 9 ? plan_fast(hq_host, n_taps, frac_bits, acc_bits, 0x7fffffff, false) : FastPlan{};
    if (plan.ok && !force_generic() && (in_dtype == FIRGPU_U8 || in_dtype == FIRGPU_I16)) {
        const int bias = 1 << (frac_bits - 1);
        return in_dtype == FIRGPU_U8 ? run_fast_taps<uint8_t>(n_taps, x, rows, cols, ld_x, hq, h, frac_bits, bias, d_sums, d_counts, s)
                                     : run_fast_taps<int16_t>(n_taps, x, rows, cols, ld_x, hq, h, frac_bits, bias, d_sums, d_counts, s);
    }
    switch (in_dtype) {
    case FIRGPU_U8: return run<uint8_t>(x, rows, cols, ld_x, hq, h, n_taps, frac_bits, acc_bits, d_sums, d_counts, s);
    case FIRGPU_I16: return run<int16_t>(x, rows, cols, ld_x, hq, h, n_taps, frac_bits, acc_bits, d_sums, d_counts, s);
    case FIRGPU_I32: return run<int32_t>(x, rows, cols, ld_x, hq, h, n_taps, frac_bits, acc_bits, d_sums, d_counts, s);
    case FIRGPU_F32: return run<float>(x, rows, cols, ld_x, hq, h, n_taps, frac_bits, acc_bits, d_sums, d_counts, s);
    case FIRGPU_F64: return run<double>(x, rows, cols, ld_x, hq, h, n_taps, frac_bits, acc_bits, d_sums, d_counts, s);
    default: return set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_compare: unsupported in_dtype %d", in_dtype);
    }
}

// The accumulator as read back from the device -> the seven reductions of firgpu_fir1d_compare.
void compare_unpack(const double (&host)[8], double *out)
{
    unsigned long long cnt[3];
    memcpy(cnt, &host[4], sizeof(cnt));
    out[0] = host[0];
    out[1] = host[1];
    out[2] = host[2];
    out[3] = host[3];
    out[4] = (double)cnt[0];
    out[5] = (double)cnt[1];
    out[6] = (double)cnt[2];
}

// x: DEVICE pointer.  out[7] (HOST): the seven reductions listed at the top.
// Synchronises the stream before returning.
int launch_fir1d_compare(const void *x, int in_dtype, int64_t rows, int64_t cols, int64_t ld_x, const int32_t *hq_host,
                         const double *h_host, int n_taps, int frac_bits, int acc_bits, double *out, cudaStream_t s)
{
    double *d_sums = nullptr;
    if (int rc0 = scratch_alloc((void **)&d_sums, 8 * sizeof(double), s)) return rc0;
    FIRGPU_CUDA(cudaMemsetAsync(d_sums, 0, 8 * sizeof(double), s));
    int rc = launch_fir1d_compare_into(x, in_dtype, rows, cols, ld_x, hq_host, h_host, n_taps, frac_bits, acc_bits, d_sums, s);
    if (rc) {
        cudaFreeAsync(d_sums, s);
        return rc;
    }
    double host[8];
    FIRGPU_CUDA(cudaMemcpyAsync(host, d_sums, sizeof(host), cudaMemcpyDeviceToHost, s));
    FIRGPU_CUDA(cudaFreeAsync(d_sums, s));
    FIRGPU_CUDA(cudaStreamSynchronize(s));
    compare_unpack(host, out);
    return 0;
}

}  // namespace firgpu
