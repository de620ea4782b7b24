// Shared declarations for the firgpu CUDA library (sm_100a only).
//
// Arithmetic contract (bit-exact with the reference, paths relative to the
// reference root):
//   prep    floor(x + 0.5) in float64, clamp [0,255]   fir_1d/model/python/fir_1d_ref.py:35-41
//   taps    int32, quantised by the caller              fir_1d/model/python/fir_1d_fixed_ref.py:79-81
//   sum     one accumulator per output, all taps        fir_1d_fixed_ref.py:95-107
//   post    wrap(acc_bits) -> +2^(frac-1) -> >> frac -> clamp [0,255]   fir_1d_fixed_ref.py:110-126
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/firgpu.h"

namespace firgpu {

constexpr int kPixelMax = 255;
constexpr int kInlineTaps = 32;   // taps passed by value in the launch parameters

struct InlineTaps {
    int32_t v[kInlineTaps];
};
struct InlineTapsF64 {
    double v[kInlineTaps];
};

// ---- error plumbing (api.cu) ------------------------------------------------
int set_error(int code, const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
void note_launch(int path);

// FIRGPU_FORCE_GENERIC=1 routes every call to the one-thread-per-output kernels (tests of the seams between kernel
// families); read once per process.
inline bool force_generic()
{
    static const bool on = [] {
        const char *e = getenv("FIRGPU_FORCE_GENERIC");
        return e && e[0] == '1';
    }();
    return on;
}

#define FIRGPU_CUDA(call)                                                  \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) return ::firgpu::check_cuda(e__, #call);  \
    } while (0)

inline int dtype_size(int dt)
{
    switch (dt) {
    case FIRGPU_U8: return 1;
    case FIRGPU_I16: return 2;
    case FIRGPU_I32: return 4;
    case FIRGPU_F32: return 4;
    case FIRGPU_F64: return 8;
    default: return 0;
    }
}

// ---- device helpers ---------------------------------------------------------
// Input container -> integer pixel in [0,255].
__device__ __forceinline__ int prep_sample(uint8_t v) { return v; }
__device__ __forceinline__ int prep_sample(int16_t v) { return min(max((int)v, 0), kPixelMax); }
__device__ __forceinline__ int prep_sample(int32_t v) { return min(max(v, 0), kPixelMax); }
// float64 / float32: the reference rounds half up as floor(x + 0.5) with a float64 add, then clamps
// (fir_1d_ref.py:35-41).  The conversion instructions that spelling needs (F2F / FRND / F2I on float64) run at a
// fraction of the integer rate, so the same value is computed from the bit pattern: for 0 <= x < 256,
// floor(x + 0.5) = (floor(2x) + 1) >> 1 and floor(2x) is a right shift of the mantissa.  The float64 add is exact
// on that range (0.5 is a multiple of ulp(x), and a carry into the next binade cannot reach an integer) with ONE
// exception, the largest double below 0.5: 0.5 - 2^-54 + 0.5 is a tie that rounds to 1.0, so the reference
// returns 1 there and so does this.  Negative inputs (x + 0.5 < 0.5) give 0 after the clamp; x >= 256 gives 255.
// NaN / inf never reach the device: the host wrapper rejects them like _validate_x does (fir_1d_ref.py:27-33).
__device__ __forceinline__ int prep_sample(double v)
{
    const uint32_t hi = (uint32_t)__double2hiint(v), lo = (uint32_t)__double2loint(v);
    const int e = (int)((hi >> 20) & 0x7ffu);
    const uint32_t m = (hi & 0xfffffu) | 0x100000u;            // mantissa >> 32 with the implicit one
    const uint32_t q = m >> min(max(1042 - e, 0), 31);         // floor(2x) for x < 256
    int r = min((int)((q + 1u) >> 1), kPixelMax);
    if (e >= 1031) r = kPixelMax;                              // x >= 256
    if (hi == 0x3fdfffffu && lo == 0xffffffffu) r = 1;         // the float64 tie described above
    if (hi >> 31) r = 0;                                       // x < 0 or -0.0
    return r;
}
// float32 containers: the reference evaluates `sample + 0.5` on the array's own scalars (NumPy 2: np.float32 + 0.5 stays
// float32, fir_1d_ref.py:37), so the add is a FLOAT32 add.  Like the float64 one it is exact on [0, 256) except for one
// value, the largest float32 below 0.5 (0x3effffff): 0.5 - 2^-25 + 0.5 is a tie that rounds to 1.0, and the reference
// returns 1 for it (a Python list of the same values is float64 arithmetic and returns 0: the container decides).
__device__ __forceinline__ int prep_sample(float v)
{
    // floor(2x) with one saturating conversion (2x is exact; +-inf-sized products saturate), then (q + 1) >> 1.
    const int q = min(__float2int_rd(v * 2.0f), 2 * kPixelMax + 1);
    int r = min(max((q + 1) >> 1, 0), kPixelMax);
    if (__float_as_uint(v) == 0x3effffffu) r = 1;              // the float32 tie described above
    return r;
}

// Parametrised forms (extended entry points): sample clamp range [lo, hi] and
// a raw mode for the floating-point model; (0, 255) reproduces the reference.
__device__ __forceinline__ int prep_sample_range(uint8_t v, int lo, int hi) { return min(max((int)v, lo), hi); }
__device__ __forceinline__ int prep_sample_range(int16_t v, int lo, int hi) { return min(max((int)v, lo), hi); }
__device__ __forceinline__ int prep_sample_range(int32_t v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int prep_sample_range(double v, int lo, int hi)
{
    double r = floor(__dadd_rn(v, 0.5));
    r = fmin(fmax(r, (double)lo), (double)hi);
    return (int)r;
}
__device__ __forceinline__ int prep_sample_range(float v, int lo, int hi)
{
    double r = floor((double)__fadd_rn(v, 0.5f));              // the reference's add is a float32 add for float32 containers
    r = fmin(fmax(r, (double)lo), (double)hi);
    return (int)r;
}

// wrap -> bias -> shift -> saturate on a 64-bit accumulator.  |acc| < 2^50 for
// every supported tap count, so int64 itself never overflows and the only
// wrap is the explicit one the reference performs.
__device__ __forceinline__ int post_accumulate(long long acc, int frac_bits, int acc_bits, int out_lo = 0, int out_hi = kPixelMax)
{
    if (acc_bits < 64) {
        unsigned long long m = (unsigned long long)acc & ((1ULL << acc_bits) - 1ULL);
        if (m >> (acc_bits - 1)) m -= (1ULL << acc_bits);
        acc = (long long)m;
    }
    long long v;
    if (frac_bits > 62) {
        v = 0;   // 0 <= acc + 2^(frac-1) < 2^frac for |acc| < 2^61
    } else {
        acc += 1LL << (frac_bits - 1);
        v = acc >> frac_bits;
    }
    return (int)(v < out_lo ? out_lo : (v > out_hi ? out_hi : v));
}

template <typename T> struct DtypeOf;
template <> struct DtypeOf<uint8_t> { static constexpr int value = FIRGPU_U8; };
template <> struct DtypeOf<int16_t> { static constexpr int value = FIRGPU_I16; };
template <> struct DtypeOf<int32_t> { static constexpr int value = FIRGPU_I32; };
template <> struct DtypeOf<float> { static constexpr int value = FIRGPU_F32; };
template <> struct DtypeOf<double> { static constexpr int value = FIRGPU_F64; };

// ---- launchers implemented per translation unit -----------------------------
struct Fir1dArgs {
    const void *x;
    void *y;
    int in_dtype, out_dtype;
    int64_t rows, cols, ld_x, ld_y;
    int n_taps, dilation, frac_bits, acc_bits;
    int64_t halo_left, halo_right;
    cudaStream_t stream;
    // extended entry points (defaults = the reference model)
    int centre = -1;                 // taps ahead of the output position; -1 = n_taps / 2
    int in_lo = 0, in_hi = 255;      // sample clamp range
    int out_lo = 0, out_hi = 255;    // output saturation range
    int raw = 0;                     // ideal model only: use samples as they are (no round / clamp)
    // streaming (rows == 1, fast kernels only): the halo_left scalars in front of x live in `hist`, not in memory before
    // x; the kernel writes the row's last halo_left scalars to `hist_out` (a different buffer)
    const void *hist = nullptr;
    void *hist_out = nullptr;
    int64_t period = 0;              // staged kernel only: this one row is really rows of `period` scalars, each zero-padded
                                     // on its own (a dense row batch filtered in one flattened pass)
    bool is_reference_model() const
    {
        return (centre < 0 || centre == n_taps / 2) && in_lo == 0 && in_hi == 255 && out_lo == 0 && out_hi == 255 && !raw;
    }
};

struct Fir2dArgs {
    const void *x;
    void *y;
    int in_dtype, out_dtype;
    int64_t n_img, height, width, ld_x, ld_y, img_stride_x, img_stride_y;
    int kh, kw, frac_bits, acc_bits;
    int64_t halo_top, halo_bottom;
    cudaStream_t stream;
    int x_row_padded = 0;            // uint8 rows are followed by one readable zero byte (re-pitched copies of odd widths)
};

// fir1d_generic.cu
int launch_fir1d_fixed_generic(const Fir1dArgs &a, const int32_t *hq_host);
int launch_fir1d_ideal(const Fir1dArgs &a, const double *h_host);
// fir1d_ideal_fast.cu — 1 = served, 0 = not eligible
int try_launch_fir1d_ideal_fast(const Fir1dArgs &a, const double *h_host);
// fir1d_staged.cu — returns 1 when the call was served, 0 when the shape /
// taps are outside the staged kernel's domain (caller falls to generic), <0 / >0 on error.
int try_launch_fir1d_fixed_staged(const Fir1dArgs &a, const int32_t *hq_host);
// the body of firgpu_fir1d_fixed_ex after its argument checks; fast_only: do not fall back to the generic kernel
// (returns 1 when a fast kernel served the call, 0 when none applies, else an error code)
int fir1d_fixed_ex_device(const Fir1dArgs &a, const int32_t *hq, bool fast_only);
// host-only: the datapath that would serve these taps (firgpu_fir1d_fixed_plan)
int plan_fir1d_fixed(const int32_t *hq_host, int n_taps, int frac_bits, int acc_bits, int *path, int *post, int *mac);
// wide datapath of the extended entry point (signed int16 ranges); 1 = served, 0 = not eligible
int try_launch_fir1d_fixed_staged_wide(const Fir1dArgs &a, const int32_t *hq_host);
// n_filters tap sets ([n_filters][n_taps]) over one pass of the input; outputs filter_stride_y elements apart.
int try_launch_fir1d_fixed_staged_multi(const Fir1dArgs &a, const int32_t *hq_host, int n_filters, int64_t filter_stride_y);
// prep_u8.cu — int32 / float32 / float64 containers -> uint8 pixels (round half up, clamp)
// same-type copy of rows x cols elements (uint8 / int16) to a new pitch, zero-filled up to it (pitches in elements,
// ld_y * element size a multiple of 4, y 4-byte aligned)
int launch_repitch(const void *x, int dtype, void *y, int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y, cudaStream_t s);
int launch_prep_u8(const void *x, int in_dtype, uint8_t *y, int64_t rows, int64_t span, int64_t ld_x, int64_t ld_y,
                   int64_t before, int64_t shift, cudaStream_t s);
// fir1d_complex.cu — y[i] = (i even) ? a[i] : b[i] over n scalars of es bytes
int launch_select_even_odd(const void *a, const void *b, void *y, int64_t n, int es, cudaStream_t s);
// fir1d_compare.cu
int launch_fir1d_compare(const void *x, int in_dtype, int64_t rows, int64_t cols, int64_t ld_x, const int32_t *hq_host,
                         const double *h_host, int n_taps, int frac_bits, int acc_bits, double *out, cudaStream_t s);
// accumulating form (d_sums: 8 zeroed doubles on the device) + unpacking of the accumulator read back to the host
int launch_fir1d_compare_into(const void *x, int in_dtype, int64_t rows, int64_t cols, int64_t ld_x, const int32_t *hq_host,
                              const double *h_host, int n_taps, int frac_bits, int acc_bits, double *d_sums, cudaStream_t s);
void compare_unpack(const double (&host)[8], double *out);
// fir2d.cu
int launch_fir2d_fixed_generic(const Fir2dArgs &a, const int32_t *hq_host);
int launch_fir2d_ideal(const Fir2dArgs &a, const double *h_host);
// fir2d_ideal_fast.cu — 1 = served, 0 = not eligible
int try_launch_fir2d_ideal_fast(const Fir2dArgs &a, const double *h_host);
int try_launch_fir2d_fixed_tiled(const Fir2dArgs &a, const int32_t *hq_host);
bool fir2d_tiled_window_ok(int kh, int kw);      // window sizes the tensor-map kernel serves (even sizes by zero padding)

// api.cu — stream-ordered scratch memory from the library's own pool (keeps its pages between calls); free with
// cudaFreeAsync on the same stream
int scratch_alloc(void **ptr, size_t bytes, cudaStream_t stream);

// Device copy of a long tap vector, freed on the stream after the launch.
template <typename T>
int upload_taps(const T *host, int n, cudaStream_t s, T **dev)
{
    if (int rc = scratch_alloc((void **)dev, sizeof(T) * (size_t)n, s)) return rc;
    FIRGPU_CUDA(cudaMemcpyAsync(*dev, host, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, s));
    return 0;
}

}  // namespace firgpu
