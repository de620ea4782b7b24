/*
 * Plain-C restatement of the reference FIR golden models.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * "--impl reference" legs may load this library, and only as the checker or
 * as the CPU baseline that is reported next to the GPU number.  The product
 * is the CUDA library declared in include/firgpu.h; it has no CPU fallback.
 *
 * Parity status: 1-D fixed and ideal are PINNED (reference KATs and golden
 * vectors made by the unmodified reference, tests/golden/).  Interleaved I/Q
 * (dilation 2) is pinned by composition.  fir_2d is PARITY UNPINNED: the
 * reference contains no fir_2d code; the 2-D functions are our specification
 * (see oracle/fir_oracle_py.py).
 *
 * All file:line citations are relative to /root/reference.  The sample loop
 * mirrors the reference's structure on purpose (one accumulator per output,
 * bounds test per tap) — it is meant to be read next to
 * fir_1d/model/python/fir_1d_fixed_ref.py:95-128, not to be fast.
 *
 * Exactness: the reference accumulates in unbounded Python ints.  Here the
 * accumulator is int64; with |x| <= 255, |hq| <= 2**31 and L <= 1024 the sum
 * stays below 2**50, so int64 never overflows and the explicit wrap to
 * acc_bits is the only wrap that happens, exactly as in the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define FIRO_MAX_TAPS 1024

enum { FIRO_U8 = 0, FIRO_I16 = 1, FIRO_I32 = 2, FIRO_F32 = 3, FIRO_F64 = 4 };

/* fir_1d/model/python/fir_1d_ref.py:35-41 — floor(s + 0.5) in float64, then
 * clamp to [0, 255].  For integer containers the rounding is the identity. */
static inline int64_t prep_from_double(double s)
{
    double r = floor(s + 0.5);
    if (r < 0.0) return 0;
    if (r > 255.0) return 255;
    return (int64_t)r;
}

static inline int64_t load_sample(const void *base, int dtype, int64_t idx)
{
    switch (dtype) {
    case FIRO_U8:  return ((const uint8_t *)base)[idx];
    case FIRO_I16: { int64_t v = ((const int16_t *)base)[idx]; return v < 0 ? 0 : v > 255 ? 255 : v; }
    case FIRO_I32: { int64_t v = ((const int32_t *)base)[idx]; return v < 0 ? 0 : v > 255 ? 255 : v; }
    case FIRO_F32: {
        /* float32 containers: the reference's `sample + 0.5` is a float32 add (NumPy 2 scalar promotion,
         * fir_1d_ref.py:37); volatile keeps the compiler from evaluating it in a wider type */
        volatile float sum = ((const float *)base)[idx] + 0.5f;
        double r = floor((double)sum);
        if (r < 0.0) return 0;
        if (r > 255.0) return 255;
        return (int64_t)r;
    }
    default:       return prep_from_double(((const double *)base)[idx]);
    }
}

/* fir_1d_fixed_ref.py:110-126 — mask to acc_bits, restore the sign from the
 * top kept bit, add half an LSB WITHOUT wrapping again, floor-shift, clamp. */
static inline int64_t post_accumulate(int64_t acc, int frac_bits, int acc_bits)
{
    if (acc_bits < 63) {
        uint64_t m = ((uint64_t)acc) & ((1ULL << acc_bits) - 1ULL);
        acc = (int64_t)m;
        if (m >> (acc_bits - 1)) acc -= (int64_t)(1ULL << acc_bits);
    }
    int64_t v;
    if (frac_bits > 62) {
        v = 0;                      /* 0 <= acc + 2**(frac-1) < 2**frac */
    } else {
        acc += (int64_t)1 << (frac_bits - 1);
        v = acc >> frac_bits;       /* arithmetic shift == Python's floor */
    }
    return v < 0 ? 0 : v > 255 ? 255 : v;
}

static inline void store_out(void *base, int dtype, int64_t idx, int64_t v)
{
    if (dtype == FIRO_U8) ((uint8_t *)base)[idx] = (uint8_t)v;
    else                  ((int16_t *)base)[idx] = (int16_t)v;
}

/* rows independent signals of cols samples each (gen_fixed_output.py:44-59);
 * ld_* are row pitches in ELEMENTS.  y[n] = post(sum_k hq[k] * xs[n - (k -
 * L/2) * dilation]) with zeros outside the row (fir_1d_fixed_ref.py:95-107). */
int firo_fir1d_fixed(const void *x, void *y, int in_dtype, int out_dtype,
                     int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                     const int32_t *hq, int n_taps, int dilation,
                     int frac_bits, int acc_bits, int n_threads)
{
    if (n_taps < 1 || n_taps > FIRO_MAX_TAPS || frac_bits < 1 || acc_bits < 1 || dilation < 1)
        return -1;
    const int centre = n_taps / 2;
    const size_t in_size = in_dtype == FIRO_U8 ? 1 : in_dtype == FIRO_I16 ? 2 : in_dtype == FIRO_F64 ? 8 : 4;
    const size_t out_size = out_dtype == FIRO_U8 ? 1 : 2;
    (void)n_threads;
#pragma omp parallel for collapse(2) schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
    for (int64_t r = 0; r < rows; ++r) {
        for (int64_t blk = 0; blk < (cols + 65535) / 65536; ++blk) {
            const char *xr = (const char *)x + (size_t)r * (size_t)ld_x * in_size;
            char *yr = (char *)y + (size_t)r * (size_t)ld_y * out_size;
            int64_t n_end = (blk + 1) * 65536 < cols ? (blk + 1) * 65536 : cols;
            for (int64_t n = blk * 65536; n < n_end; ++n) {
                int64_t acc = 0;
                for (int k = 0; k < n_taps; ++k) {
                    int64_t src = n - (int64_t)(k - centre) * dilation;
                    if (src >= 0 && src < cols)
                        acc += load_sample(xr, in_dtype, src) * (int64_t)hq[k];
                }
                store_out(yr, out_dtype, n, post_accumulate(acc, frac_bits, acc_bits));
            }
        }
    }
    return 0;
}

/* fir_1d/model/python/fir_1d_ref.py:43-65 — float64, ascending k, separate
 * multiply and add (compile with -ffp-contract=off), no clamp on the output. */
int firo_fir1d_ideal(const void *x, double *y, int in_dtype,
                     int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                     const double *h, int n_taps, int dilation, int n_threads)
{
    if (n_taps < 1 || n_taps > FIRO_MAX_TAPS || dilation < 1) return -1;
    const int centre = n_taps / 2;
    const size_t in_size = in_dtype == FIRO_U8 ? 1 : in_dtype == FIRO_I16 ? 2 : in_dtype == FIRO_F64 ? 8 : 4;
    (void)n_threads;
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
    for (int64_t r = 0; r < rows; ++r) {
        const char *xr = (const char *)x + (size_t)r * (size_t)ld_x * in_size;
        double *yr = y + (size_t)r * (size_t)ld_y;
        for (int64_t n = 0; n < cols; ++n) {
            double acc = 0.0;
            for (int k = 0; k < n_taps; ++k) {
                int64_t src = n - (int64_t)(k - centre) * dilation;
                if (src >= 0 && src < cols) {
                    double prod = h[k] * (double)load_sample(xr, in_dtype, src);
                    acc = acc + prod;
                }
            }
            yr[n] = acc;
        }
    }
    return 0;
}

/* 2-D (our spec, parity unpinned): images of height x width, batch of n_img
 * separated by img_stride elements; one accumulator per pixel over the whole
 * kh x kw window, zero outside the image, then the 1-D post stage. */
int firo_fir2d_fixed(const void *x, void *y, int in_dtype, int out_dtype,
                     int64_t n_img, int64_t height, int64_t width,
                     int64_t ld_x, int64_t ld_y, int64_t img_stride_x, int64_t img_stride_y,
                     const int32_t *hq, int kh, int kw,
                     int frac_bits, int acc_bits, int n_threads)
{
    if (kh < 1 || kw < 1 || kh * kw > FIRO_MAX_TAPS || frac_bits < 1 || acc_bits < 1) return -1;
    const int ch = kh / 2, cw = kw / 2;
    const size_t in_size = in_dtype == FIRO_U8 ? 1 : in_dtype == FIRO_I16 ? 2 : in_dtype == FIRO_F64 ? 8 : 4;
    const size_t out_size = out_dtype == FIRO_U8 ? 1 : 2;
    (void)n_threads;
#pragma omp parallel for collapse(2) schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
    for (int64_t b = 0; b < n_img; ++b) {
        for (int64_t i = 0; i < height; ++i) {
            const char *xb = (const char *)x + (size_t)b * (size_t)img_stride_x * in_size;
            char *yrow = (char *)y + ((size_t)b * (size_t)img_stride_y + (size_t)i * (size_t)ld_y) * out_size;
            for (int64_t j = 0; j < width; ++j) {
                int64_t acc = 0;
                for (int m = 0; m < kh; ++m) {
                    int64_t si = i - m + ch;
                    if (si < 0 || si >= height) continue;
                    for (int n = 0; n < kw; ++n) {
                        int64_t sj = j - n + cw;
                        if (sj >= 0 && sj < width)
                            acc += load_sample(xb, in_dtype, si * ld_x + sj) * (int64_t)hq[m * kw + n];
                    }
                }
                store_out(yrow, out_dtype, j, post_accumulate(acc, frac_bits, acc_bits));
            }
        }
    }
    return 0;
}

int firo_fir2d_ideal(const void *x, double *y, int in_dtype,
                     int64_t n_img, int64_t height, int64_t width,
                     int64_t ld_x, int64_t ld_y, int64_t img_stride_x, int64_t img_stride_y,
                     const double *h, int kh, int kw, int n_threads)
{
    if (kh < 1 || kw < 1 || kh * kw > FIRO_MAX_TAPS) return -1;
    const int ch = kh / 2, cw = kw / 2;
    const size_t in_size = in_dtype == FIRO_U8 ? 1 : in_dtype == FIRO_I16 ? 2 : in_dtype == FIRO_F64 ? 8 : 4;
    (void)n_threads;
#pragma omp parallel for collapse(2) schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
    for (int64_t b = 0; b < n_img; ++b) {
        for (int64_t i = 0; i < height; ++i) {
            const char *xb = (const char *)x + (size_t)b * (size_t)img_stride_x * in_size;
            double *yrow = y + (size_t)b * (size_t)img_stride_y + (size_t)i * (size_t)ld_y;
            for (int64_t j = 0; j < width; ++j) {
                double acc = 0.0;
                for (int m = 0; m < kh; ++m) {
                    int64_t si = i - m + ch;
                    if (si < 0 || si >= height) continue;
                    for (int n = 0; n < kw; ++n) {
                        int64_t sj = j - n + cw;
                        if (sj >= 0 && sj < width) {
                            double prod = h[m * kw + n] * (double)load_sample(xb, in_dtype, si * ld_x + sj);
                            acc = acc + prod;
                        }
                    }
                }
                yrow[j] = acc;
            }
        }
    }
    return 0;
}

int firo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
