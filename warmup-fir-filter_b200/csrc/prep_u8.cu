// Sample containers the fast kernels do not read (int32, float32, float64):
// one streaming pass converts them to the reference's uint8 pixels —
// round half up as floor(x + 0.5) in float64, clamp to [0, 255]
// (fir_1d/model/python/fir_1d_ref.py:35-41) — so that the filtering itself can
// run on the staged / tiled kernels instead of the one-thread-per-output ones.
#include "common.cuh"

#include <cstdlib>
#include <type_traits>

namespace firgpu {
namespace {

constexpr int kPrepThreads = 256;

__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d)
{
    return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
}
// four consecutive samples from a 16-byte aligned address -> four uint8 pixels
__device__ __forceinline__ uint32_t quad(const float *p)
{
    const float4 v = *reinterpret_cast<const float4 *>(p);
    return pack4(prep_sample(v.x), prep_sample(v.y), prep_sample(v.z), prep_sample(v.w));
}
__device__ __forceinline__ uint32_t quad(const int32_t *p)
{
    const int4 v = *reinterpret_cast<const int4 *>(p);
    return pack4(prep_sample(v.x), prep_sample(v.y), prep_sample(v.z), prep_sample(v.w));
}
__device__ __forceinline__ uint32_t quad(const int16_t *p)        // 8-byte aligned is enough here
{
    const short4 v = *reinterpret_cast<const short4 *>(p);
    return pack4(prep_sample(v.x), prep_sample(v.y), prep_sample(v.z), prep_sample(v.w));
}
__device__ __forceinline__ uint32_t quad(const double *p)
{
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    return pack4(prep_sample(a.x), prep_sample(a.y), prep_sample(b.x), prep_sample(b.y));
}

template <typename InT, int kPrepQuads>                          // 4-sample groups per thread per block step
__global__ void __launch_bounds__(kPrepThreads)
prep_u8_kernel(const InT *__restrict__ x, uint8_t *__restrict__ y, int64_t rows, int64_t span, int64_t ld_x, int64_t ld_y,
               int64_t before, int64_t shift)
{
    // Row r of the temporary (16-byte aligned, pitch ld_y a multiple of 16) holds, from byte `shift` on, the `span`
    // samples that start `before` elements in front of the row pointer; bytes outside that window are written as 0.
    // A thread turns four consecutive samples (one 16-byte load; two for float64) into one 32-bit store, so a warp
    // reads 512 contiguous bytes and writes 128.  Block steps that touch a row end or start from an unaligned
    // source take the checked scalar path.
    constexpr int kPrepBlock = kPrepThreads * kPrepQuads * 4;     // temporary bytes per block step
    const int tid = threadIdx.x;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const InT *xr = x + r * ld_x - before - shift;        // xr[j] is the sample for temporary byte j
        uint8_t *yr = y + r * ld_y;
        const bool aligned = (reinterpret_cast<uintptr_t>(xr) & 15) == 0;
        for (int64_t j0 = (int64_t)blockIdx.x * kPrepBlock; j0 < ld_y; j0 += (int64_t)gridDim.x * kPrepBlock) {
            if (aligned && j0 >= shift && j0 + kPrepBlock - shift <= span) {
                uint32_t w[kPrepQuads];
#pragma unroll
                for (int q = 0; q < kPrepQuads; ++q) w[q] = quad(xr + j0 + (q * kPrepThreads + tid) * 4);
#pragma unroll
                for (int q = 0; q < kPrepQuads; ++q) *reinterpret_cast<uint32_t *>(yr + j0 + (q * kPrepThreads + tid) * 4) = w[q];
            } else {
                for (int q = 0; q < kPrepQuads; ++q) {
                    const int64_t j = j0 + (q * kPrepThreads + tid) * 4;
                    if (j >= ld_y) break;
                    uint32_t w = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int64_t i = j + b - shift;
                        if (i >= 0 && i < span) w |= (uint32_t)prep_sample(xr[j + b]) << (8 * b);
                    }
                    *reinterpret_cast<uint32_t *>(yr + j) = w;
                }
            }
        }
    }
}

// Same-type copy to a new pitch: a thread writes one aligned 32-bit word of the destination from element loads of
// the source (rows of an odd pitch start at every alignment); elements past `cols` are written as 0.
template <typename T>
__global__ void __launch_bounds__(kPrepThreads)
repitch_kernel(const T *__restrict__ x, T *__restrict__ y, int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y)
{
    constexpr int PER = 4 / (int)sizeof(T);
    const int64_t words = ld_y / PER;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const T *xr = x + r * ld_x;
        uint32_t *yr = reinterpret_cast<uint32_t *>(y + r * ld_y);
        for (int64_t w = (int64_t)blockIdx.x * kPrepThreads + threadIdx.x; w < words; w += (int64_t)gridDim.x * kPrepThreads) {
            uint32_t v = 0;
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int64_t j = w * PER + q;
                if (j < cols) v |= (uint32_t)(typename std::make_unsigned<T>::type)xr[j] << (8 * (int)sizeof(T) * q);
            }
            yr[w] = v;
        }
    }
}

template <typename InT, int Q>
int run_prep(const void *x, uint8_t *y, int64_t rows, int64_t span, int64_t ld_x, int64_t ld_y, int64_t before, int64_t shift,
             int ctas_per_sm, cudaStream_t s)
{
    constexpr int kPrepBlock = kPrepThreads * Q * 4;
    int64_t gx = (ld_y + kPrepBlock - 1) / kPrepBlock;
    if (gx > 148 * ctas_per_sm) gx = 148 * ctas_per_sm;
    int64_t gy = rows;
    const int64_t cap = (148 * ctas_per_sm + gx - 1) / gx;
    if (gy > cap) gy = cap;
    prep_u8_kernel<InT, Q><<<dim3((unsigned)gx, (unsigned)gy), kPrepThreads, 0, s>>>((const InT *)x, y, rows, span, ld_x, ld_y, before, shift);
    return 0;
}

template <typename InT>
int run_prep_q(const void *x, uint8_t *y, int64_t rows, int64_t span, int64_t ld_x, int64_t ld_y, int64_t before, int64_t shift,
               cudaStream_t s)
{
    static int q = -1, c = -1;
    if (q < 0) {
        const char *e = getenv("FIRGPU_PREP_QUADS");
        q = e ? atoi(e) : 8;
        e = getenv("FIRGPU_PREP_CTAS");
        c = e ? atoi(e) : 4;
    }
    switch (q) {
    case 1: return run_prep<InT, 1>(x, y, rows, span, ld_x, ld_y, before, shift, c, s);
    case 2: return run_prep<InT, 2>(x, y, rows, span, ld_x, ld_y, before, shift, c, s);
    case 4: return run_prep<InT, 4>(x, y, rows, span, ld_x, ld_y, before, shift, c, s);
    default: return run_prep<InT, 8>(x, y, rows, span, ld_x, ld_y, before, shift, c, s);
    }
}

}  // namespace

int launch_repitch(const void *x, int dtype, void *y, int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y, cudaStream_t s)
{
    const int es = dtype == FIRGPU_U8 ? 1 : 2;
    int64_t gx = (ld_y * es / 4 + kPrepThreads - 1) / kPrepThreads;
    if (gx > 148 * 8) gx = 148 * 8;
    if (gx < 1) gx = 1;
    int64_t gy = rows;
    const int64_t cap = (148 * 16 + gx - 1) / gx;
    if (gy > cap) gy = cap;
    const dim3 grid((unsigned)gx, (unsigned)gy);
    if (dtype == FIRGPU_U8) repitch_kernel<uint8_t><<<grid, kPrepThreads, 0, s>>>((const uint8_t *)x, (uint8_t *)y, rows, cols, ld_x, ld_y);
    else if (dtype == FIRGPU_I16) repitch_kernel<int16_t><<<grid, kPrepThreads, 0, s>>>((const int16_t *)x, (int16_t *)y, rows, cols, ld_x, ld_y);
    else return set_error(FIRGPU_E_DTYPE, "repitch: unsupported dtype %d", dtype);
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_STAGED);
    return 0;
}

// y: 16-byte aligned temporary of rows x ld_y bytes (ld_y a multiple of 16).  Sample k of the window (k = 0 is the
// element `before` in front of the row pointer, `span` samples in all) lands at byte shift + k of its row.
int launch_prep_u8(const void *x, int in_dtype, uint8_t *y, int64_t rows, int64_t span, int64_t ld_x, int64_t ld_y,
                   int64_t before, int64_t shift, cudaStream_t s)
{
    switch (in_dtype) {
    case FIRGPU_I16: run_prep_q<int16_t>(x, y, rows, span, ld_x, ld_y, before, shift, s); break;
    case FIRGPU_I32: run_prep_q<int32_t>(x, y, rows, span, ld_x, ld_y, before, shift, s); break;
    case FIRGPU_F32: run_prep_q<float>(x, y, rows, span, ld_x, ld_y, before, shift, s); break;
    case FIRGPU_F64: run_prep_q<double>(x, y, rows, span, ld_x, ld_y, before, shift, s); break;
    default: return set_error(FIRGPU_E_DTYPE, "prep_u8: unsupported in_dtype %d", in_dtype);
    }
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_STAGED);
    return 0;
}

}  // namespace firgpu
