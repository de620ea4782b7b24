// Fast path of the floating-point 2-D model (DESIGN.md §2 "fir_2d": the 2-D
// counterpart of fir_1d/model/python/fir_1d_ref.py:43-65) for uint8 / int16
// pixels, 3x3 and 5x5 windows, float64 (bit-identical to the generic kernel and
// the scalar specification) or float32 (tolerance) output.
//
// Same construction as fir1d_ideal_fast.cu: registers only, a lane produces one
// 16-byte store worth of adjacent outputs of one row (2 doubles / 4 floats) and
// reads the sample pairs of its KH x (KW + G - 1) window as 2- or 4-byte global
// loads.  Horizontally neighbouring lanes and vertically neighbouring warps
// overlap, so all but the first touch of a sector hit L1 / L2 (an image row is
// re-read by the KH output rows around it; DRAM sees it once).  Accumulation
// order is the model's: window rows m ascending, columns n ascending, unfused
// __dmul_rn / __dadd_rn from +0.0; samples outside the image enter as +0.0,
// which never changes such a sum (see fir1d_ideal_fast.cu).
//
// Bound: the fp64 pipe for 5x5 (50 multiply/adds + 15 conversions per pixel at
// 64 lanes/clk/SM ~ 280 Gpixel/s), HBM (9 B/pixel) for 3x3.
#include "common.cuh"
#include "ideal_common.cuh"

#include <cstdlib>

namespace firgpu {
namespace {

constexpr int kThreads2I = 256;

template <int N> struct Taps2F64 {
    double v[N];
};

template <typename InT, typename F, int KH, int KW, int kGroups>
__global__ void __launch_bounds__(kThreads2I)
fir2d_ideal_fast_kernel(const InT *__restrict__ x, F *__restrict__ y, int64_t height, int64_t width, int64_t ld_x, int64_t ld_y,
                        int64_t img_stride_x, int64_t img_stride_y, int64_t tiles_per_row, int64_t total_tiles,
                        Taps2F64<KH * KW> h, int64_t halo_top, int64_t halo_bottom)
{
    constexpr int kWarps = kThreads2I / 32;
    constexpr int OP = 16 / (2 * (int)sizeof(F));        // output pairs per group: one 16-byte store
    constexpr int G = 2 * OP;                            // outputs per group
    constexpr int kTileI = 32 * G * kGroups;             // outputs (of one row) per warp tile
    constexpr int ch = KH / 2, cw = KW / 2, lo = KW - 1 - cw;
    constexpr int PL = (lo + 1) / 2, PR = (cw + 1) / 2;  // neighbour pairs a window reaches into, left / right
    constexpr int NP = PL + OP + PR;                     // sample pairs per window row
    using P2 = typename PairOf<InT>::type;
    const int lane = threadIdx.x & 31;
    const int64_t row_lo = -halo_top, row_hi = height + halo_bottom;

    auto widen = [](P2 r, F &w0, F &w1) {
        if constexpr (sizeof(InT) == 1) {
            w0 = pixel_to<F>(r & 0xff);
            w1 = pixel_to<F>(r >> 8);
        } else {
            w0 = pixel_to<F>(prep_sample((int16_t)(r & 0xffffu)));
            w1 = pixel_to<F>(prep_sample((int16_t)(r >> 16)));
        }
    };
    // window row m: w[i] is sample base - 2 PL + i of image row i - m + ch; output base + e reads w[2 PL + cw - n + e]
    auto accumulate_row = [&](int m, const F (&w)[2 * NP], F (&a)[G]) {
#pragma unroll
        for (int n = 0; n < KW; ++n) {
#pragma unroll
            for (int e = 0; e < G; ++e) {
                if constexpr (sizeof(F) == 8) a[e] = __dadd_rn(a[e], __dmul_rn(h.v[m * KW + n], w[2 * PL + cw - n + e]));
                else a[e] = fmaf((float)h.v[m * KW + n], w[2 * PL + cw - n + e], a[e]);
            }
        }
    };

    // one tile per warp, tiles in address order (see fir1d_ideal_fast.cu)
    const int64_t tile = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (tile >= total_tiles) return;
    const int64_t gr = tile / tiles_per_row, tcol = tile - gr * tiles_per_row;
    const int64_t b = gr / height, i = gr - b * height;
    const InT *xb = x + b * img_stride_x;
    F *yr = y + b * img_stride_y + i * ld_y;
    const int pad = (int)((reinterpret_cast<uintptr_t>(yr) / sizeof(F)) % G);   // shift that keeps the 16-byte stores aligned
    const int64_t col0 = tcol * kTileI - pad;
    // pair loads need every window row's first pair aligned: an even pitch keeps the rows' alignment equal
    const bool x_pairs = (ld_x & 1) == 0 && (reinterpret_cast<uintptr_t>(xb + i * ld_x + col0) & (2 * sizeof(InT) - 1)) == 0;
    // interior: every window row and column of every output of the tile is inside the image (no checks at all)
    const bool interior = col0 - 2 * PL >= 0 && col0 + kTileI + 2 * PR <= width && x_pairs &&
                          i + ch - (KH - 1) >= row_lo && i + ch < row_hi;

    if (interior) {
#pragma unroll 2
        for (int j = 0; j < kGroups; ++j) {
            const int64_t n0 = col0 + (j * 32 + lane) * G;
            P2 raw[KH][NP];
#pragma unroll
            for (int m = 0; m < KH; ++m)
#pragma unroll
                for (int p = 0; p < NP; ++p) raw[m][p] = *reinterpret_cast<const P2 *>(xb + (i - m + ch) * ld_x + n0 + 2 * (p - PL));
            F a[G];
#pragma unroll
            for (int e = 0; e < G; ++e) a[e] = (F)0;
#pragma unroll
            for (int m = 0; m < KH; ++m) {
                F w[2 * NP];
#pragma unroll
                for (int p = 0; p < NP; ++p) widen(raw[m][p], w[2 * p], w[2 * p + 1]);
                accumulate_row(m, w, a);
            }
            if constexpr (sizeof(F) == 8) *reinterpret_cast<double2 *>(yr + n0) = make_double2(a[0], a[1]);
            else *reinterpret_cast<float4 *>(yr + n0) = make_float4(a[0], a[1], a[2], a[3]);
        }
    } else {
        // image edges (left / right / top / bottom), odd pitches: every window row checks its row, every sample pair its column range
        for (int j = 0; j < kGroups; ++j) {
            const int64_t n0 = col0 + (j * 32 + lane) * G;
            if (n0 >= width) break;
            if (n0 + G <= 0) continue;
            F a[G];
#pragma unroll
            for (int e = 0; e < G; ++e) a[e] = (F)0;
#pragma unroll
            for (int m = 0; m < KH; ++m) {
                const int64_t si = i - m + ch;
                if (si < row_lo || si >= row_hi) continue;
                const InT *xr = xb + si * ld_x;
                F w[2 * NP];
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const int64_t g = n0 - 2 * PL + 2 * p;
                    if (x_pairs && g >= 0 && g + 1 < width) {
                        widen(*reinterpret_cast<const P2 *>(xr + g), w[2 * p], w[2 * p + 1]);
                    } else {
                        w[2 * p] = (g >= 0 && g < width) ? pixel_to<F>(prep_sample(xr[g])) : (F)0;
                        w[2 * p + 1] = (g + 1 >= 0 && g + 1 < width) ? pixel_to<F>(prep_sample(xr[g + 1])) : (F)0;
                    }
                }
                accumulate_row(m, w, a);
            }
            if (n0 >= 0 && n0 + G <= width) {
                if constexpr (sizeof(F) == 8) *reinterpret_cast<double2 *>(yr + n0) = make_double2(a[0], a[1]);
                else *reinterpret_cast<float4 *>(yr + n0) = make_float4(a[0], a[1], a[2], a[3]);
            } else {
#pragma unroll
                for (int e = 0; e < G; ++e)
                    if (n0 + e >= 0 && n0 + e < width) yr[n0 + e] = a[e];
            }
        }
    }
}

template <typename InT, typename F, int KH, int KW>
int run(const Fir2dArgs &a, const double *h_host)
{
    constexpr int kGroups = 4;
    constexpr int kWarps = kThreads2I / 32;
    constexpr int kGroup = 16 / (int)sizeof(F);
    constexpr int kTileI = 32 * kGroup * kGroups;
    Taps2F64<KH * KW> h;
    for (int k = 0; k < KH * KW; ++k) h.v[k] = h_host[k];
    const int64_t tiles_per_row = (a.width + kGroup - 1 + kTileI - 1) / kTileI;   // room for the per-row alignment shift
    const int64_t total = tiles_per_row * a.height * a.n_img;
    const int64_t grid = (total + kWarps - 1) / kWarps;
    if (grid > 0x7fffffffLL) return 0;
    fir2d_ideal_fast_kernel<InT, F, KH, KW, kGroups><<<(unsigned)grid, kThreads2I, 0, a.stream>>>(
        (const InT *)a.x, (F *)a.y, a.height, a.width, a.ld_x, a.ld_y, a.img_stride_x, a.img_stride_y, tiles_per_row, total, h,
        a.halo_top, a.halo_bottom);
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_STAGED);
    return 1;
}

template <typename InT, typename F>
int by_window(const Fir2dArgs &a, const double *h)
{
    if (a.kh == 3 && a.kw == 3) return run<InT, F, 3, 3>(a, h);
    if (a.kh == 5 && a.kw == 5) return run<InT, F, 5, 5>(a, h);
    if (a.kh == 3 && a.kw == 5) return run<InT, F, 3, 5>(a, h);
    if (a.kh == 5 && a.kw == 3) return run<InT, F, 5, 3>(a, h);
    return 0;
}

}  // namespace

// 1 = served, 0 = not eligible (caller uses the generic kernel), else error.
int try_launch_fir2d_ideal_fast(const Fir2dArgs &a, const double *h_host)
{
    if (force_generic()) return 0;
    if (a.in_dtype == FIRGPU_U8) {
        if (a.out_dtype == FIRGPU_F64) return by_window<uint8_t, double>(a, h_host);
        if (a.out_dtype == FIRGPU_F32) return by_window<uint8_t, float>(a, h_host);
    } else if (a.in_dtype == FIRGPU_I16) {
        if (a.out_dtype == FIRGPU_F64) return by_window<int16_t, double>(a, h_host);
        if (a.out_dtype == FIRGPU_F32) return by_window<int16_t, float>(a, h_host);
    }
    return 0;
}

}  // namespace firgpu
