// Generic 2-D kernels: any kh x kw <= FIRGPU_MAX_TAPS, any acc_bits, any
// pitch, every sample container.  One thread per output pixel.
//
// The reference ships no fir_2d; the arithmetic is the 1-D model's
// (fir_1d/model/python/fir_1d_fixed_ref.py:79-81,110-126) around the course
// formula y[i,j] = sum_m sum_n h[m,n] x[i-m+kh/2, j-n+kw/2]
// (fir_1d/docs/DSP_2025_2_2nd_assignment_v4.pdf p.14), one accumulator and
// one wrap/round/saturate per pixel.  See DESIGN.md "2-D specification".
#include "common.cuh"

namespace firgpu {
namespace {

constexpr int kBlock = 256;

template <typename InT, typename OutT>
__global__ void __launch_bounds__(kBlock)
fir2d_fixed_generic_kernel(const InT *__restrict__ x, OutT *__restrict__ y,
                           int64_t n_img, int64_t height, int64_t width,
                           int64_t ld_x, int64_t ld_y, int64_t img_stride_x, int64_t img_stride_y,
                           InlineTaps inline_taps, const int32_t *__restrict__ long_taps,
                           int kh, int kw, int frac_bits, int acc_bits,
                           int64_t halo_top, int64_t halo_bottom)
{
    const int ch = kh / 2, cw = kw / 2;
    const int64_t row_lo = -halo_top, row_hi = height + halo_bottom;
    const int64_t total_rows = n_img * height;
    for (int64_t gr = blockIdx.y; gr < total_rows; gr += gridDim.y) {
        const int64_t b = gr / height, i = gr - b * height;
        const InT *xb = x + b * img_stride_x;
        OutT *yr = y + b * img_stride_y + i * ld_y;
        for (int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x; j < width;
             j += (int64_t)gridDim.x * kBlock) {
            long long acc = 0;
            for (int m = 0; m < kh; ++m) {
                const int64_t si = i - m + ch;
                if (si < row_lo || si >= row_hi) continue;
                const InT *xr = xb + si * ld_x;
                for (int n = 0; n < kw; ++n) {
                    const int64_t sj = j - n + cw;
                    if (sj >= 0 && sj < width) {
                        const int t = m * kw + n;
                        const int32_t tap = long_taps ? long_taps[t] : inline_taps.v[t];
                        acc += (long long)prep_sample(xr[sj]) * (long long)tap;
                    }
                }
            }
            yr[j] = (OutT)post_accumulate(acc, frac_bits, acc_bits);
        }
    }
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(kBlock)
fir2d_ideal_kernel(const InT *__restrict__ x, OutT *__restrict__ y,
                   int64_t n_img, int64_t height, int64_t width,
                   int64_t ld_x, int64_t ld_y, int64_t img_stride_x, int64_t img_stride_y,
                   InlineTapsF64 inline_taps, const double *__restrict__ long_taps,
                   int kh, int kw, int64_t halo_top, int64_t halo_bottom)
{
    const int ch = kh / 2, cw = kw / 2;
    const int64_t row_lo = -halo_top, row_hi = height + halo_bottom;
    const int64_t total_rows = n_img * height;
    for (int64_t gr = blockIdx.y; gr < total_rows; gr += gridDim.y) {
        const int64_t b = gr / height, i = gr - b * height;
        const InT *xb = x + b * img_stride_x;
        OutT *yr = y + b * img_stride_y + i * ld_y;
        for (int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x; j < width;
             j += (int64_t)gridDim.x * kBlock) {
            OutT acc = (OutT)0;
            for (int m = 0; m < kh; ++m) {
                const int64_t si = i - m + ch;
                if (si < row_lo || si >= row_hi) continue;
                const InT *xr = xb + si * ld_x;
                for (int n = 0; n < kw; ++n) {
                    const int64_t sj = j - n + cw;
                    if (sj >= 0 && sj < width) {
                        const int t = m * kw + n;
                        const double tap = long_taps ? long_taps[t] : inline_taps.v[t];
                        const int s = prep_sample(xr[sj]);
                        if constexpr (sizeof(OutT) == 8) {
                            acc = __dadd_rn(acc, __dmul_rn(tap, (double)s));
                        } else {
                            acc = fmaf((float)tap, (float)s, acc);
                        }
                    }
                }
            }
            yr[j] = acc;
        }
    }
}

dim3 grid_for(int64_t total_rows, int64_t width)
{
    int64_t gx = (width + kBlock - 1) / kBlock;
    if (gx > 148 * 8) gx = 148 * 8;
    if (gx < 1) gx = 1;
    int64_t gy = total_rows < 1 ? 1 : total_rows;
    if (gy > 65535) gy = 65535;
    return dim3((unsigned)gx, (unsigned)gy, 1);
}

template <typename InT, typename OutT>
int run_fixed(const Fir2dArgs &a, const int32_t *hq_host)
{
    const int nt = a.kh * a.kw;
    InlineTaps it{};
    int32_t *dev_taps = nullptr;
    if (nt <= kInlineTaps) {
        for (int k = 0; k < nt; ++k) it.v[k] = hq_host[k];
    } else {
        int rc = upload_taps(hq_host, nt, a.stream, &dev_taps);
        if (rc) return rc;
    }
    fir2d_fixed_generic_kernel<InT, OutT><<<grid_for(a.n_img * a.height, a.width), kBlock, 0, a.stream>>>(
        (const InT *)a.x, (OutT *)a.y, a.n_img, a.height, a.width, a.ld_x, a.ld_y, a.img_stride_x,
        a.img_stride_y, it, dev_taps, a.kh, a.kw, a.frac_bits, a.acc_bits, a.halo_top, a.halo_bottom);
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_GENERIC);
    if (dev_taps) FIRGPU_CUDA(cudaFreeAsync(dev_taps, a.stream));
    return 0;
}

template <typename InT, typename OutT>
int run_ideal(const Fir2dArgs &a, const double *h_host)
{
    const int nt = a.kh * a.kw;
    InlineTapsF64 it{};
    double *dev_taps = nullptr;
    if (nt <= kInlineTaps) {
        for (int k = 0; k < nt; ++k) it.v[k] = h_host[k];
    } else {
        int rc = upload_taps(h_host, nt, a.stream, &dev_taps);
        if (rc) return rc;
    }
    fir2d_ideal_kernel<InT, OutT><<<grid_for(a.n_img * a.height, a.width), kBlock, 0, a.stream>>>(
        (const InT *)a.x, (OutT *)a.y, a.n_img, a.height, a.width, a.ld_x, a.ld_y, a.img_stride_x,
        a.img_stride_y, it, dev_taps, a.kh, a.kw, a.halo_top, a.halo_bottom);
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_GENERIC);
    if (dev_taps) FIRGPU_CUDA(cudaFreeAsync(dev_taps, a.stream));
    return 0;
}

template <typename OutT>
int dispatch_fixed_in(const Fir2dArgs &a, const int32_t *hq)
{
    switch (a.in_dtype) {
    case FIRGPU_U8: return run_fixed<uint8_t, OutT>(a, hq);
    case FIRGPU_I16: return run_fixed<int16_t, OutT>(a, hq);
    case FIRGPU_I32: return run_fixed<int32_t, OutT>(a, hq);
    case FIRGPU_F32: return run_fixed<float, OutT>(a, hq);
    case FIRGPU_F64: return run_fixed<double, OutT>(a, hq);
    default: return set_error(FIRGPU_E_DTYPE, "fir2d_fixed: unsupported in_dtype %d", a.in_dtype);
    }
}

template <typename OutT>
int dispatch_ideal_in(const Fir2dArgs &a, const double *h)
{
    switch (a.in_dtype) {
    case FIRGPU_U8: return run_ideal<uint8_t, OutT>(a, h);
    case FIRGPU_I16: return run_ideal<int16_t, OutT>(a, h);
    case FIRGPU_I32: return run_ideal<int32_t, OutT>(a, h);
    case FIRGPU_F32: return run_ideal<float, OutT>(a, h);
    case FIRGPU_F64: return run_ideal<double, OutT>(a, h);
    default: return set_error(FIRGPU_E_DTYPE, "fir2d_ideal: unsupported in_dtype %d", a.in_dtype);
    }
}

}  // namespace

int launch_fir2d_fixed_generic(const Fir2dArgs &a, const int32_t *hq_host)
{
    switch (a.out_dtype) {
    case FIRGPU_U8: return dispatch_fixed_in<uint8_t>(a, hq_host);
    case FIRGPU_I16: return dispatch_fixed_in<int16_t>(a, hq_host);
    default: return set_error(FIRGPU_E_DTYPE, "fir2d_fixed: out_dtype must be U8 or I16, got %d", a.out_dtype);
    }
}

int launch_fir2d_ideal(const Fir2dArgs &a, const double *h_host)
{
    const int fast = try_launch_fir2d_ideal_fast(a, h_host);
    if (fast == 1) return 0;
    if (fast != 0) return fast;
    switch (a.out_dtype) {
    case FIRGPU_F64: return dispatch_ideal_in<double>(a, h_host);
    case FIRGPU_F32: return dispatch_ideal_in<float>(a, h_host);
    default: return set_error(FIRGPU_E_DTYPE, "fir2d_ideal: out_dtype must be F64 or F32, got %d", a.out_dtype);
    }
}

}  // namespace firgpu
