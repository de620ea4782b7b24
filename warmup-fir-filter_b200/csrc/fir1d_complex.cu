// Complex taps on interleaved I/Q (firgpu_fir1d_fixed_complex) — the step SURVEY 8(f)4 lists after streaming.
//
// The reference's MyComplex has complex + complex and complex * double only (lib/mycomplex.h:10-41), so it pins how a
// real-scalar product and a sum act per component but defines no complex product: the specification is ours,
// parity unpinned except for imaginary taps of zero, where it is the
// real-tap I/Q filter (dilation 2) that IS pinned by composition.
//
// Over the interleaved scalar stream s[2n] = Re z[n], s[2n+1] = Im z[n] the two components are two REAL filters of
// 2L taps, one accumulator and one rounding each:
//   Re y[n] = y_A[2n],    A: taps g[2k] = -hq_im[k], g[2k+1] = hq_re[k], centre 2c + 1
//   Im y[n] = y_B[2n+1],  B: taps g[2k] =  hq_re[k], g[2k+1] = hq_im[k], centre 2c          (c = L // 2)
// Both run on the ordinary fixed-point kernels (firgpu_fir1d_fixed_ex's dispatch: staged / wide / generic) into scratch
// rows; the kernel below keeps the even scalars of A and the odd scalars of B.
#include "common.cuh"

namespace firgpu {
namespace {

// y[i] = (i even) ? a[i] : b[i], 16 bytes per thread: a 32-bit lane mask for 16-bit elements, a byte mask for 8-bit ones
template <int ES>
__global__ void __launch_bounds__(256) select_even_odd_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b, uint4 *__restrict__ y,
                                                              int64_t n_vec, const unsigned char *a_tail, const unsigned char *b_tail,
                                                              unsigned char *y_tail, int tail_elems)
{
    constexpr uint32_t keep_a = ES == 2 ? 0x0000ffffu : 0x00ff00ffu;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 va = a[i], vb = b[i];
        uint4 r;
        r.x = (va.x & keep_a) | (vb.x & ~keep_a);
        r.y = (va.y & keep_a) | (vb.y & ~keep_a);
        r.z = (va.z & keep_a) | (vb.z & ~keep_a);
        r.w = (va.w & keep_a) | (vb.w & ~keep_a);
        y[i] = r;
    }
    if (blockIdx.x == 0)
        for (int e = threadIdx.x; e < tail_elems * ES; e += blockDim.x) {
            const int elem = e / ES;
            y_tail[e] = (elem & 1) ? b_tail[e] : a_tail[e];
        }
}

}  // namespace

// a, b, y: n scalars of ES bytes each, all 16-byte aligned (scratch rows and a caller row at an even scalar offset).
int launch_select_even_odd(const void *a, const void *b, void *y, int64_t n, int es, cudaStream_t s)
{
    const int per_vec = 16 / es;
    const bool aligned = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)y) & 15) == 0;
    const int64_t n_vec = aligned ? n / per_vec : 0;
    const int64_t done = n_vec * per_vec;
    if (n - done > (1 << 30)) return set_error(FIRGPU_E_SHAPE, "fir1d_fixed_complex: unaligned output rows are not supported");
    int64_t grid = (n_vec + 255) / 256;
    grid = grid < 1 ? 1 : (grid > 148 * 16 ? 148 * 16 : grid);
    const unsigned char *at = (const unsigned char *)a + done * es, *bt = (const unsigned char *)b + done * es;
    unsigned char *yt = (unsigned char *)y + done * es;
    if (es == 2)
        select_even_odd_kernel<2><<<(unsigned)grid, 256, 0, s>>>((const uint4 *)a, (const uint4 *)b, (uint4 *)y, n_vec, at, bt, yt, (int)(n - done));
    else
        select_even_odd_kernel<1><<<(unsigned)grid, 256, 0, s>>>((const uint4 *)a, (const uint4 *)b, (uint4 *)y, n_vec, at, bt, yt, (int)(n - done));
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_STAGED);
    return 0;
}

}  // namespace firgpu
