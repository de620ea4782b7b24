// Generic 1-D kernels: any tap count <= FIRGPU_MAX_TAPS, any acc_bits, any
// alignment / pitch, every sample container.  One thread per output sample,
// int64 accumulator (fixed) or unfused float64 multiply + add (ideal).
//
// These are the correctness anchor of the library and the path for shapes the
// staged kernel (fir1d_staged.cu) does not take.  They are coalesced but not
// tuned: each thread re-reads its taps-wide window through L1.
//
// Reference arithmetic: fir_1d/model/python/fir_1d_fixed_ref.py:95-128 and
// fir_1d/model/python/fir_1d_ref.py:55-63.
#include "common.cuh"

namespace firgpu {
namespace {

constexpr int kBlock = 256;

template <typename InT, typename OutT>
__global__ void __launch_bounds__(kBlock)
fir1d_fixed_generic_kernel(const InT *__restrict__ x, OutT *__restrict__ y,
                           int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                           InlineTaps inline_taps, const int32_t *__restrict__ long_taps,
                           int n_taps, int dilation, int frac_bits, int acc_bits,
                           int64_t halo_left, int64_t halo_right, int centre, int in_lo, int in_hi, int out_lo, int out_hi)
{
    const int64_t lo = -halo_left, hi = cols + halo_right;   // readable index range of a row
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const InT *xr = x + r * ld_x;
        OutT *yr = y + r * ld_y;
        for (int64_t n = (int64_t)blockIdx.x * kBlock + threadIdx.x; n < cols;
             n += (int64_t)gridDim.x * kBlock) {
            long long acc = 0;
            for (int k = 0; k < n_taps; ++k) {
                const int64_t src = n - (int64_t)(k - centre) * dilation;
                if (src >= lo && src < hi) {
                    const int32_t tap = long_taps ? long_taps[k] : inline_taps.v[k];
                    acc += (long long)prep_sample_range(xr[src], in_lo, in_hi) * (long long)tap;
                }
            }
            yr[n] = (OutT)post_accumulate(acc, frac_bits, acc_bits, out_lo, out_hi);
        }
    }
}

// OutT double: bit-identical float64 (no FMA).  OutT float: float32 datapath.
template <typename InT, typename OutT>
__global__ void __launch_bounds__(kBlock)
fir1d_ideal_kernel(const InT *__restrict__ x, OutT *__restrict__ y,
                   int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                   InlineTapsF64 inline_taps, const double *__restrict__ long_taps,
                   int n_taps, int dilation, int64_t halo_left, int64_t halo_right, int centre, int in_lo, int in_hi, int raw)
{
    const int64_t lo = -halo_left, hi = cols + halo_right;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const InT *xr = x + r * ld_x;
        OutT *yr = y + r * ld_y;
        for (int64_t n = (int64_t)blockIdx.x * kBlock + threadIdx.x; n < cols;
             n += (int64_t)gridDim.x * kBlock) {
            OutT acc = (OutT)0;
            for (int k = 0; k < n_taps; ++k) {
                const int64_t src = n - (int64_t)(k - centre) * dilation;
                if (src >= lo && src < hi) {
                    const double tap = long_taps ? long_taps[k] : inline_taps.v[k];
                    // raw: the C++ Fir1D model filters the samples as given (fir_1d/model/cpp/fir_1d_ref.cpp:37-63)
                    const double s = raw ? (double)xr[src] : (double)prep_sample_range(xr[src], in_lo, in_hi);
                    if constexpr (sizeof(OutT) == 8) {
                        acc = __dadd_rn(acc, __dmul_rn(tap, s));
                    } else {
                        acc = fmaf((float)tap, (float)s, acc);
                    }
                }
            }
            yr[n] = acc;
        }
    }
}

dim3 grid_for(int64_t rows, int64_t cols)
{
    int64_t gx = (cols + kBlock - 1) / kBlock;
    if (gx > 148 * 32) gx = 148 * 32;
    if (gx < 1) gx = 1;
    int64_t gy = rows < 1 ? 1 : rows;
    const int64_t cap = (148 * 64 + gx - 1) / gx;
    if (gy > cap) gy = cap < 1 ? 1 : cap;
    if (gy > 65535) gy = 65535;
    return dim3((unsigned)gx, (unsigned)gy, 1);
}

template <typename InT, typename OutT>
int run_fixed(const Fir1dArgs &a, const int32_t *hq_host)
{
    InlineTaps it{};
    int32_t *dev_taps = nullptr;
    if (a.n_taps <= kInlineTaps) {
        for (int k = 0; k < a.n_taps; ++k) it.v[k] = hq_host[k];
    } else {
        int rc = upload_taps(hq_host, a.n_taps, a.stream, &dev_taps);
        if (rc) return rc;
    }
    fir1d_fixed_generic_kernel<InT, OutT><<<grid_for(a.rows, a.cols), kBlock, 0, a.stream>>>(
        (const InT *)a.x, (OutT *)a.y, a.rows, a.cols, a.ld_x, a.ld_y, it, dev_taps, a.n_taps,
        a.dilation, a.frac_bits, a.acc_bits, a.halo_left, a.halo_right, a.centre < 0 ? a.n_taps / 2 : a.centre, a.in_lo, a.in_hi,
        a.out_lo, a.out_hi);
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_GENERIC);
    if (dev_taps) FIRGPU_CUDA(cudaFreeAsync(dev_taps, a.stream));
    return 0;
}

template <typename InT, typename OutT>
int run_ideal(const Fir1dArgs &a, const double *h_host)
{
    InlineTapsF64 it{};
    double *dev_taps = nullptr;
    if (a.n_taps <= kInlineTaps) {
        for (int k = 0; k < a.n_taps; ++k) it.v[k] = h_host[k];
    } else {
        int rc = upload_taps(h_host, a.n_taps, a.stream, &dev_taps);
        if (rc) return rc;
    }
    fir1d_ideal_kernel<InT, OutT><<<grid_for(a.rows, a.cols), kBlock, 0, a.stream>>>(
        (const InT *)a.x, (OutT *)a.y, a.rows, a.cols, a.ld_x, a.ld_y, it, dev_taps, a.n_taps,
        a.dilation, a.halo_left, a.halo_right, a.centre < 0 ? a.n_taps / 2 : a.centre, a.in_lo, a.in_hi, a.raw);
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_GENERIC);
    if (dev_taps) FIRGPU_CUDA(cudaFreeAsync(dev_taps, a.stream));
    return 0;
}

template <typename OutT>
int dispatch_fixed_in(const Fir1dArgs &a, const int32_t *hq)
{
    switch (a.in_dtype) {
    case FIRGPU_U8: return run_fixed<uint8_t, OutT>(a, hq);
    case FIRGPU_I16: return run_fixed<int16_t, OutT>(a, hq);
    case FIRGPU_I32: return run_fixed<int32_t, OutT>(a, hq);
    case FIRGPU_F32: return run_fixed<float, OutT>(a, hq);
    case FIRGPU_F64: return run_fixed<double, OutT>(a, hq);
    default: return set_error(FIRGPU_E_DTYPE, "fir1d_fixed: unsupported in_dtype %d", a.in_dtype);
    }
}

template <typename OutT>
int dispatch_ideal_in(const Fir1dArgs &a, const double *h)
{
    switch (a.in_dtype) {
    case FIRGPU_U8: return run_ideal<uint8_t, OutT>(a, h);
    case FIRGPU_I16: return run_ideal<int16_t, OutT>(a, h);
    case FIRGPU_I32: return run_ideal<int32_t, OutT>(a, h);
    case FIRGPU_F32: return run_ideal<float, OutT>(a, h);
    case FIRGPU_F64: return run_ideal<double, OutT>(a, h);
    default: return set_error(FIRGPU_E_DTYPE, "fir1d_ideal: unsupported in_dtype %d", a.in_dtype);
    }
}

}  // namespace

int launch_fir1d_fixed_generic(const Fir1dArgs &a, const int32_t *hq_host)
{
    switch (a.out_dtype) {
    case FIRGPU_U8: return dispatch_fixed_in<uint8_t>(a, hq_host);
    case FIRGPU_I16: return dispatch_fixed_in<int16_t>(a, hq_host);
    default: return set_error(FIRGPU_E_DTYPE, "fir1d_fixed: out_dtype must be U8 or I16, got %d", a.out_dtype);
    }
}

int launch_fir1d_ideal(const Fir1dArgs &a, const double *h_host)
{
    const int fast = try_launch_fir1d_ideal_fast(a, h_host);
    if (fast == 1) return 0;
    if (fast != 0) return fast;
    switch (a.out_dtype) {
    case FIRGPU_F64: return dispatch_ideal_in<double>(a, h_host);
    case FIRGPU_F32: return dispatch_ideal_in<float>(a, h_host);
    default: return set_error(FIRGPU_E_DTYPE, "fir1d_ideal: out_dtype must be F64 or F32, got %d", a.out_dtype);
    }
}

}  // namespace firgpu
