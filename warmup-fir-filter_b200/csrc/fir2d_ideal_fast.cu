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
// 64 lanes/clk/SM ~ 280 Gpixel/s), HBM (9 B/pixel) for 3x3.  Images whose rows
// are aligned alike take fir2d_ideal_rows_kernel below (four output rows per
// lane: 4 conversions per pixel); this kernel serves the rest.
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


// Several output rows per lane (R), 32-byte output groups: a lane owns G = 4 float64 / 8 float32 adjacent columns of R
// consecutive output rows and walks the R + KH - 1 image rows under them ONCE, from the bottom one up — so that every
// output still receives its window rows in ascending m — converting each row's G + 4 samples once and feeding every output
// row it belongs to.  Conversions per pixel: (R + KH - 1) (G + 4) / (R G) = 4 for 5x5 float64 at R = 4, against 15 for the
// one-row form above; loads 3 aligned chunks per image row, one 256-bit store per output row.  Needs the rows' chunks
// aligned alike (pitches multiples of the group, x and y at the same element offset modulo the group); other images take
// the one-row kernel.
constexpr int kThreadsRows = 128;                     // upper bound; 64 measured equal, 256 (with the looser register bound) 4-6 % slower

template <typename InT, typename F, int KH, int KW, int R>
__global__ void __launch_bounds__(kThreadsRows)
fir2d_ideal_rows_kernel(const InT *__restrict__ x, F *__restrict__ y, int64_t height, int64_t width, int64_t ld_x, int64_t ld_y,
                        int64_t img_stride_x, int64_t img_stride_y, uint32_t tiles_per_row, uint32_t row_blocks, uint32_t total_tiles,
                        Taps2F64<KH * KW> h, int64_t halo_top, int64_t halo_bottom)
{
    const int kWarps = (int)(blockDim.x >> 5);
    constexpr int G = 32 / (int)sizeof(F);               // outputs per group: one 256-bit store
    constexpr int ch = KH / 2, cw = KW / 2, lo = KW - 1 - cw;
    static_assert(lo <= 4 && cw <= 4, "one chunk either side");
    constexpr int NW = G + 8;                            // converted samples per image row: n0 - 4 ... n0 + G + 3
    constexpr int WPC = (int)sizeof(InT);                // 32-bit words per 4-sample chunk
    constexpr int NCH = G / 4 + 2;
    const int lane = threadIdx.x & 31;
    const int64_t row_lo = -halo_top, row_hi = height + halo_bottom;

    const uint32_t tile = blockIdx.x * (uint32_t)kWarps + (threadIdx.x >> 5);
    if (tile >= total_tiles) return;
    const uint32_t gr = tile / tiles_per_row, tcol = tile - gr * tiles_per_row;
    const uint32_t b = gr / row_blocks, rb = gr - b * row_blocks;
    const int64_t i0 = (int64_t)rb * R;
    const InT *xb = x + b * img_stride_x;
    F *yb = y + b * img_stride_y;
    const int pad = (int)((reinterpret_cast<uintptr_t>(yb) / sizeof(F)) % G);   // the same for every row: ld_y is a multiple of G
    const int64_t n0 = (int64_t)tcol * (32 * G) - pad + lane * G;
    if (n0 >= width || n0 + G <= 0) return;
    const bool fast = n0 - 4 >= 0 && n0 + G + 4 <= width;                     // all chunks inside the row

    F a[R][G];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int e = 0; e < G; ++e) a[r][e] = (F)0;

    uint32_t raw[R + KH - 1][NCH * WPC];
    if (fast) {
#pragma unroll
        for (int t = 0; t < R + KH - 1; ++t) {
            const int64_t s = i0 + R - 1 + ch - t;
            if (s < row_lo || s >= row_hi) continue;
            const InT *q = xb + s * ld_x + n0 - 4;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if constexpr (sizeof(InT) == 1) raw[t][c] = *reinterpret_cast<const uint32_t *>(q + 4 * c);
                else {
                    const uint2 v = *reinterpret_cast<const uint2 *>(q + 4 * c);
                    raw[t][2 * c] = v.x, raw[t][2 * c + 1] = v.y;
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < R + KH - 1; ++t) {
        const int64_t s = i0 + R - 1 + ch - t;                                 // image rows from the bottom up
        if (s < row_lo || s >= row_hi) continue;                               // a row of +0.0 samples changes no sum
        F w[NW];
        if (fast) {
#pragma unroll
            for (int c = 4 - lo; c < 4 + G + cw; ++c) {
                if constexpr (sizeof(InT) == 1 && sizeof(F) == 4)
                    w[c] = __fadd_rn(__uint_as_float(__byte_perm(raw[t][c >> 2], 0x4b000000u, 0x7650u + (c & 3))), -8388608.0f);
                else if constexpr (sizeof(InT) == 1)
                    w[c] = pixel_to<F>((int)__byte_perm(raw[t][c >> 2], 0u, 0x4440u + (c & 3)));
                else
                    w[c] = pixel_to<F>(prep_sample((int16_t)(raw[t][c >> 1] >> (16 * (c & 1)))));
            }
        } else {
            const InT *xr = xb + s * ld_x;
#pragma unroll
            for (int c = 4 - lo; c < 4 + G + cw; ++c) {
                const int64_t g = n0 - 4 + c;
                w[c] = (g >= 0 && g < width) ? pixel_to<F>(prep_sample(xr[g])) : (F)0;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int m = r - (R - 1) + t;                                     // window row of output row i0 + r that image row s is
            if (m < 0 || m >= KH) continue;
#pragma unroll
            for (int n = 0; n < KW; ++n) {
#pragma unroll
                for (int e = 0; e < G; ++e) {
                    if constexpr (sizeof(F) == 8) a[r][e] = __dadd_rn(a[r][e], __dmul_rn(h.v[m * KW + n], w[4 + cw - n + e]));
                    else a[r][e] = fmaf((float)h.v[m * KW + n], w[4 + cw - n + e], a[r][e]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (i0 + r >= height) break;
        F *dst = yb + (i0 + r) * ld_y + n0;
        if (n0 >= 0 && n0 + G <= width) {
            if constexpr (sizeof(F) == 8)
                asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "d"(a[r][0]), "d"(a[r][1]), "d"(a[r][2]), "d"(a[r][3]) : "memory");
            else
                asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(a[r][0]), "f"(a[r][1]), "f"(a[r][2]),
                             "f"(a[r][3]), "f"(a[r][4]), "f"(a[r][5]), "f"(a[r][6]), "f"(a[r][7]) : "memory");
        } else {
#pragma unroll
            for (int e = 0; e < G; ++e)
                if (n0 + e >= 0 && n0 + e < width) dst[e] = a[r][e];
        }
    }
}

template <typename InT, typename F, int KH, int KW, int R>
int run_rows(const Fir2dArgs &a, const double *h_host, int threads)
{
    constexpr int G = 32 / (int)sizeof(F);
    const uintptr_t xa = reinterpret_cast<uintptr_t>(a.x), ya = reinterpret_cast<uintptr_t>(a.y);
    const bool aligned = xa % sizeof(InT) == 0 && ya % sizeof(F) == 0 && (xa / sizeof(InT)) % G == (ya / sizeof(F)) % G &&
                         a.ld_x % G == 0 && a.ld_y % G == 0 && (a.n_img == 1 || (a.img_stride_x % G == 0 && a.img_stride_y % G == 0));
    if (!aligned) return 0;
    Taps2F64<KH * KW> h;
    for (int k = 0; k < KH * KW; ++k) h.v[k] = h_host[k];
    const int64_t tiles_per_row = (a.width + G - 1 + 32 * G - 1) / (32 * G);      // room for the alignment shift
    const int64_t row_blocks = (a.height + R - 1) / R;
    const int64_t total = tiles_per_row * row_blocks * a.n_img;
    if (total >= 0xffffffffLL) return 0;
    const int kWarps = threads / 32;
    fir2d_ideal_rows_kernel<InT, F, KH, KW, R><<<(unsigned)((total + kWarps - 1) / kWarps), threads, 0, a.stream>>>(
        (const InT *)a.x, (F *)a.y, a.height, a.width, a.ld_x, a.ld_y, a.img_stride_x, a.img_stride_y, (uint32_t)tiles_per_row,
        (uint32_t)row_blocks, (uint32_t)total, h, a.halo_top, a.halo_bottom);
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_STAGED);
    return 1;
}

template <typename InT, typename F, int KH, int KW>
int run(const Fir2dArgs &a, const double *h_host)
{
    static const bool one_row = getenv("FIRGPU_IDEAL2D_ONE_ROW") != nullptr;       // A/B knob: the one-row-per-lane kernel
    if (!one_row) {
        static const int threads = getenv("FIRGPU_IDEAL2D_THREADS") ? atoi(getenv("FIRGPU_IDEAL2D_THREADS")) : 128;
        // rows per lane 2 / 4 / 8 on 8192^2 uint8: 3x3 float64 0.61 / 0.69-0.72 / 0.69, 5x5 float64 0.31 / 0.35 / 0.30-0.33,
        // float32 0.58 / 0.61 / 0.52 and 0.34 / 0.34 / 0.26 (profiles/r2_sweep_ideal2d_rows_threads.txt): only 4 is built
        const int served = run_rows<InT, F, KH, KW, 4>(a, h_host, threads == 64 ? 64 : 128);
        if (served != 0) return served;
    }
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
