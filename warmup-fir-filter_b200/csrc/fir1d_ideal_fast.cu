// Fast path of the ideal (floating-point) 1-D model for uint8 / int16 inputs and
// float64 (bit-identical) or float32 (tolerance) output: fir_1d/model/python/fir_1d_ref.py:43-65, batched as
// fir_1d/sim/vector/gen_ideal_output.py:37-50.
//
// Bound: HBM, 9-10 bytes per sample (1-2 in, 8 out) — the output dominates.
// Registers only: every lane produces a group of adjacent outputs and reads the
// samples its window covers straight from global memory — neighbouring lanes
// overlap, so all but the first touch of a sector hit L1.  Two forms: 32-byte
// groups (4 float64 / 8 float32 outputs, ONE 256-bit store per lane — a warp
// writes 1 KiB of whole sectors per instruction — window from aligned
// 4-sample chunks: 3 loads per group) wherever input and output rows share
// their alignment, else 16-byte groups with 2- or 4-byte pair loads.
// An earlier version converted a tile to float64 once into shared memory and
// filtered from 128-bit shared loads; ncu showed it bound by the L1/shared
// data pipe (86 % of peak: 24 B of shared loads + 8 B of 4-way conflicting
// shared stores + 8 B of global stores per sample) at 0.63 of the HBM peak.
// Converting per window instead costs (L + 1) / 2 conversions per output, so the
// cast is the exact "2^52 + v" bit trick — one add on the fp64 pipe, which has
// the headroom, instead of an I2F on the quarter-rate conversion unit.
// Round/clamp is the identity on integers, so conversion = clamp + cast.
// Arithmetic is the reference's: taps in ascending k,
// __dmul_rn then __dadd_rn (no FMA contraction), accumulator starting at +0.0
// — bit-identical to the Python model.  Zero padding outside the row enters
// as +0.0 samples, which adds +-0.0 to a +0.0-started sum and never
// changes it (x + (+-0.0) == x for every x, and +0.0 + (-0.0) == +0.0), so the
// padded form is bit-identical to the reference's "skip taps outside the row".
#include "common.cuh"
#include "ideal_common.cuh"

#include <cstdlib>

namespace firgpu {
namespace {

constexpr int kThreadsI = 256;
constexpr int kMaxFastIdealTaps = 9;

template <int L> struct TapsF64 {
    double v[L];
};

template <typename F> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

// CAUSAL = false: the Python model (output n reads samples n + L/2 ... n - (L-1-L/2), rounded and clamped pixels).
// CAUSAL = true, InT = double: the C++ Fir1D model (fir_1d/model/cpp/fir_1d_ref.cpp:16-63): y[n] = sum h[k] x[n-k] on the
// samples as they are.
// DIL = 2: interleaved I/Q with real taps (lib/mycomplex.cpp:12-22) — output n reads scalars n + 2 (hi - k); the window of
// a group is still one contiguous run of scalars, so only its reach and the tap stride change.
// GM = 2 (float32 output of uint8 / int16 samples, DIL = 1): a lane's group is two 16-byte stores (8 adjacent outputs) and its
// window comes from aligned 4-sample chunks — 3 loads and 12 conversions per 8 outputs instead of 8 and 16.  The launcher
// picks it only where every row's chunks are aligned.
template <typename InT, typename F, int L, int kGroups, bool CAUSAL, int DIL, int GM = 1>   // kGroups output groups per lane per tile
__global__ void __launch_bounds__(kThreadsI)
fir1d_ideal_fast_kernel(const InT *__restrict__ x, F *__restrict__ y, int64_t rows, int64_t cols, int64_t ld_x,
                        int64_t ld_y, int64_t tiles_per_row, int64_t total_tiles, TapsF64<L> h, int64_t halo_left,
                        int64_t halo_right, const InT *__restrict__ hist, InT *__restrict__ hist_out)
{
    // hist (streaming, rows == 1): the halo_left samples in front of column 0 come from this buffer instead of the
    // memory before x; the warp of the last tile copies the row's tail to hist_out for the next block.
    const int kWarps = (int)(blockDim.x >> 5);           // block size is a launch-time choice (<= kThreadsI)
    constexpr int OP = GM * 16 / (2 * (int)sizeof(F));   // output pairs per group: GM 16-byte stores (2 doubles / 4 floats each)
    constexpr int G = 2 * OP;                            // outputs per group
    constexpr int kTileI = 32 * G * kGroups;             // outputs per warp tile
    constexpr int hi = CAUSAL ? 0 : L / 2, lo = L - 1 - hi;
    constexpr int PL = DIL == 1 ? (lo + 1) / 2 : lo, PR = DIL == 1 ? (hi + 1) / 2 : hi;  // neighbour pairs a window reaches into, left / right
    constexpr int NP = PL + OP + PR;                     // sample pairs per window: samples base - 2 PL ... base + G - 1 + 2 PR
    constexpr int CL = (2 * PL + 3) / 4, CR = (2 * PR + 3) / 4;   // wide groups: whole 4-sample chunks either side of a group
    using P2 = typename PairOf<InT>::type;
    const int lane = threadIdx.x & 31;
    const int64_t vlo = -halo_left, vhi = cols + halo_right;
    const int64_t plo = hist ? 0 : vlo;                  // lowest column that may be loaded from x itself

    // w[i] is sample base - 2 PL + i; output base + e reads samples base + e + hi - k, k = 0 .. L-1
    auto filter_group = [&](const F (&w)[2 * NP], F (&a)[G]) {
#pragma unroll
        for (int e = 0; e < G; ++e) a[e] = (F)0;
#pragma unroll
        for (int k = 0; k < L; ++k) {
#pragma unroll
            for (int e = 0; e < G; ++e) {
                if constexpr (sizeof(F) == 8) a[e] = __dadd_rn(a[e], __dmul_rn(h.v[k], w[2 * PL + DIL * (hi - k) + e]));   // the reference's unfused float64
                else a[e] = fmaf((float)h.v[k], w[2 * PL + DIL * (hi - k) + e], a[e]);                                    // float32 datapath (same as the generic kernel's)
            }
        }
    };

    auto one = [](InT v) -> F {
        if constexpr (sizeof(InT) == 8) return (F)v;       // raw samples
        else return pixel_to<F>(prep_sample(v));
    };
    auto widen = [](P2 r, F &w0, F &w1) {
        if constexpr (sizeof(InT) == 8) {
            w0 = (F)r.x;
            w1 = (F)r.y;
        } else if constexpr (sizeof(InT) == 1) {
            w0 = pixel_to<F>(r & 0xff);
            w1 = pixel_to<F>(r >> 8);
        } else {
            w0 = pixel_to<F>(prep_sample((int16_t)(r & 0xffffu)));
            w1 = pixel_to<F>(prep_sample((int16_t)(r >> 16)));
        }
    };

    // One tile per warp; the tile index fits 32 bits (2^32 tiles would be 2^41 outputs), so locating the tile costs
    // one 32-bit division — skipped for a single row.  (A persistent grid-stride form measured 8 % slower, and its
    // 64-bit tile arithmetic was a third of all instructions executed: 150 per tile against 12 per output.)
    // Loads of a tile are all issued before its arithmetic; other resident warps cover their latency.
    // Every row's tile grid is shifted left by `pad` outputs so that group stores are 16-byte aligned whatever the
    // row pitch (odd image widths); the first tile then starts at a negative column and skips those outputs.
    const uint32_t tile = blockIdx.x * (uint32_t)kWarps + (threadIdx.x >> 5);
    if (tile >= (uint32_t)total_tiles) return;
    const uint32_t tpr = (uint32_t)tiles_per_row;
    const uint32_t row32 = rows == 1 ? 0u : tile / tpr;
    const int64_t row = row32, tcol = tile - row32 * tpr;
    {
        const InT *xr = x + row * ld_x;
        F *yr = y + row * ld_y;
        const int pad = (int)((reinterpret_cast<uintptr_t>(yr) / sizeof(F)) % G);
        const int64_t col0 = tcol * kTileI - pad;
        const bool x_pairs = (reinterpret_cast<uintptr_t>(xr + col0) & (2 * sizeof(InT) - 1)) == 0;
        auto store_group = [](F *dst, const F (&a)[G]) {
            if constexpr (sizeof(F) == 8 && GM == 1) *reinterpret_cast<double2 *>(dst) = make_double2(a[0], a[1]);
            else if constexpr (sizeof(F) == 8)
                asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]) : "memory");
            else if constexpr (GM == 1) *reinterpret_cast<float4 *>(dst) = make_float4(a[0], a[1], a[2], a[3]);
            else {
                // one 256-bit store per lane (STG.E.ENL2.256): a warp writes 1 KiB of whole sectors in one instruction.
                // Two 128-bit stores at a 32-byte lane stride half-fill every sector twice: 1.00-1.05 Tsamples/s whatever
                // the input type, against 1.12 for the 4-output form.
                asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]),
                             "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]) : "memory");
            }
        };
        bool interior = col0 >= 0 && col0 - 2 * PL >= plo && col0 + kTileI + 2 * PR <= vhi && col0 + kTileI <= cols && x_pairs;
        if constexpr (GM == 2) {
            static_assert(!CAUSAL && sizeof(InT) <= 2, "wide groups: pixels");
            // whole chunks either side of the tile
            interior = interior && col0 - 4 * CL >= plo && col0 + kTileI + 4 * CR <= vhi && (reinterpret_cast<uintptr_t>(xr + col0) & (G * sizeof(InT) - 1)) == 0;
        }
        if (GM == 2 && interior) {
            if constexpr (GM == 2) {
                constexpr int WPC = (int)sizeof(InT);                    // 32-bit words per 4-sample chunk: 1 (uint8) / 2 (int16)
                constexpr int NCH = CL + G / 4 + CR;                     // chunks: samples base - 4 CL ... base + G + 4 CR - 1
                uint32_t raw[kGroups][NCH * WPC];
#pragma unroll
                for (int j = 0; j < kGroups; ++j) {
                    const InT *b = xr + col0 + (j * 32 + lane) * G;
                    auto chunk = [&](int c) {
                        if constexpr (sizeof(InT) == 1) raw[j][c] = *reinterpret_cast<const uint32_t *>(b - 4 * CL + 4 * c);
                        else {
                            const uint2 v = *reinterpret_cast<const uint2 *>(b - 4 * CL + 4 * c);
                            raw[j][2 * c] = v.x, raw[j][2 * c + 1] = v.y;
                        }
                    };
#pragma unroll
                    for (int c = 0; c < CL; ++c) chunk(c);
                    if constexpr (G == 8) {                              // the two middle chunks in one load
                        if constexpr (sizeof(InT) == 1) {
                            const uint2 m = *reinterpret_cast<const uint2 *>(b);
                            raw[j][CL] = m.x, raw[j][CL + 1] = m.y;
                        } else {
                            const uint4 m = *reinterpret_cast<const uint4 *>(b);
                            raw[j][2 * CL] = m.x, raw[j][2 * CL + 1] = m.y, raw[j][2 * CL + 2] = m.z, raw[j][2 * CL + 3] = m.w;
                        }
                    } else chunk(CL);
#pragma unroll
                    for (int c = CL + G / 4; c < NCH; ++c) chunk(c);
                }
#pragma unroll
                for (int j = 0; j < kGroups; ++j) {
                    F w[2 * NP];
#pragma unroll
                    for (int i = 0; i < 2 * NP; ++i) {
                        constexpr int kFirst = 4 * CL - 2 * PL;          // w[0] is sample base - 2 PL = chunk sample kFirst
                        const int c = kFirst + i;
                        if constexpr (sizeof(InT) == 1 && sizeof(F) == 4) // byte -> 2^23 + byte in one permute, then the exact subtraction
                            w[i] = __fadd_rn(__uint_as_float(__byte_perm(raw[j][c >> 2], 0x4b000000u, 0x7650u + (c & 3))), -8388608.0f);
                        else if constexpr (sizeof(InT) == 1)
                            w[i] = pixel_to<F>((int)__byte_perm(raw[j][c >> 2], 0u, 0x4440u + (c & 3)));
                        else
                            w[i] = pixel_to<F>(prep_sample((int16_t)(raw[j][c >> 1] >> (16 * (c & 1)))));
                    }
                    F a[G];
                    filter_group(w, a);
                    store_group(yr + col0 + (j * 32 + lane) * G, a);
                }
            }
        } else if (GM == 1 && interior) {
            P2 raw[kGroups][NP];
#pragma unroll
            for (int j = 0; j < kGroups; ++j)
#pragma unroll
                for (int p = 0; p < NP; ++p)
                    raw[j][p] = *reinterpret_cast<const P2 *>(xr + col0 + (j * 32 + lane) * G + 2 * (p - PL));
#pragma unroll
            for (int j = 0; j < kGroups; ++j) {
                F w[2 * NP];
#pragma unroll
                for (int p = 0; p < NP; ++p) widen(raw[j][p], w[2 * p], w[2 * p + 1]);
                F a[G];
                filter_group(w, a);
                store_group(yr + col0 + (j * 32 + lane) * G, a);
            }
        } else {
            // row ends (two tiles per row), shard seams, rows whose samples are not pair-aligned: every sample pair
            // checks its range (zero padding outside [vlo, vhi)), every output its column
            for (int j = 0; j < kGroups; ++j) {
                const int64_t n = col0 + (j * 32 + lane) * G;
                if (n >= cols) break;
                if (n + G <= 0) continue;
                F w[2 * NP];
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const int64_t g = n - 2 * PL + 2 * p;
                    if (x_pairs && g >= plo && g + 1 < vhi) {
                        widen(*reinterpret_cast<const P2 *>(xr + g), w[2 * p], w[2 * p + 1]);
                    } else {
                        auto at = [&](int64_t c) -> F {
                            if (c < vlo || c >= vhi) return (F)0;
                            return one((c < 0 && hist) ? hist[c - vlo] : xr[c]);
                        };
                        w[2 * p] = at(g);
                        w[2 * p + 1] = at(g + 1);
                    }
                }
                F a[G];
                filter_group(w, a);
                if (n >= 0 && n + G <= cols) {
                    store_group(yr + n, a);
                } else {
#pragma unroll
                    for (int e = 0; e < G; ++e)
                        if (n + e >= 0 && n + e < cols) yr[n + e] = a[e];
                }
            }
        }
        if (hist_out && tile + 1 == (uint32_t)total_tiles)
            for (int64_t j = lane; j < halo_left; j += 32) hist_out[j] = xr[cols - halo_left + j];
    }
}

template <typename InT, typename F, int L, int kGroups, bool CAUSAL, int DIL = 1, int GM = 1>
int run_groups(const Fir1dArgs &a, const double *h_host)
{
    static int threads = -1;
    if (threads < 0) {
        const char *e = getenv("FIRGPU_IDEAL_THREADS");
        // measured 64 / 128 / 256: 128 is best or equal everywhere (2999 x 4499 rows 0.64 vs 0.56 at 256); wide groups: 64
        threads = e ? atoi(e) : (GM == 2 ? 64 : 128);
        if (threads < 32 || threads > kThreadsI || (threads & 31)) threads = 128;
    }
    const int kWarps = threads / 32;
    constexpr int kTileI = 32 * GM * (16 / (int)sizeof(F)) * kGroups;
    TapsF64<L> h;
    for (int k = 0; k < L; ++k) h.v[k] = h_host[k];
    constexpr int kGroup = GM * 16 / (int)sizeof(F);
    const int64_t tiles_per_row = (a.cols + kGroup - 1 + kTileI - 1) / kTileI;   // room for the per-row alignment shift
    const int64_t total = tiles_per_row * a.rows;
    // One tile per warp, CTAs in address order: concurrently running CTAs then write one advancing front, which
    // the write-dominated stream needs.
    if (total >= 0xffffffffLL) return 0;                  // (2^41 outputs: not a real case; the generic kernel takes it)
    const int64_t grid = (total + kWarps - 1) / kWarps;
    if (grid > 0x7fffffffLL) return 0;
    if (a.hist && (a.rows != 1 || a.halo_left < 1 || a.cols < a.halo_left)) return 0;
    fir1d_ideal_fast_kernel<InT, F, L, kGroups, CAUSAL, DIL, GM><<<(unsigned)grid, threads, 0, a.stream>>>(
        (const InT *)a.x, (F *)a.y, a.rows, a.cols, a.ld_x, a.ld_y, tiles_per_row, total, h, a.halo_left, a.halo_right,
        (const InT *)a.hist, (InT *)a.hist_out);
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_STAGED);
    return 1;
}

// Wide groups (GM = 2) need every row's 4-sample chunks aligned: x and y at the same element offset modulo the group.
template <typename InT, typename F>
bool chunks_aligned(const Fir1dArgs &a)
{
    constexpr int G = 32 / (int)sizeof(F);
    const uintptr_t x = reinterpret_cast<uintptr_t>(a.x), y = reinterpret_cast<uintptr_t>(a.y);
    return x % sizeof(InT) == 0 && y % sizeof(F) == 0 && (x / sizeof(InT)) % G == (y / sizeof(F)) % G &&
           (a.rows == 1 || (a.ld_x - a.ld_y) % G == 0);
}

template <typename InT, int L>
int run(const Fir1dArgs &a, const double *h_host)
{
    // groups per lane per tile: 512-output warp tiles measured best for both output types (sweep in profiles/README.md)
    if constexpr (sizeof(InT) == 8) return run_groups<InT, double, L, 8, true>(a, h_host);
    else if (a.dilation == 2) {                         // wider windows (L pairs per group): fewer groups in flight per lane
        // wide groups here too (2^26 complex int16 samples, 5 taps: float64 0.84 -> 1.02 of the HBM peak, float32 0.66 -> 0.91;
        // 2 groups per lane instead of 4: 0.92 / 0.79)
        static const int wide_iq = getenv("FIRGPU_IDEAL_IQ_WIDE") ? atoi(getenv("FIRGPU_IDEAL_IQ_WIDE")) : 4;   // 0: A/B, the 16-byte form
        if (a.out_dtype == FIRGPU_F32) {
            if (wide_iq && chunks_aligned<InT, float>(a)) {
                if (wide_iq == 4) return run_groups<InT, float, L, 2, false, 2, 2>(a, h_host);
                return run_groups<InT, float, L, 1, false, 2, 2>(a, h_host);
            }
            return run_groups<InT, float, L, 2, false, 2>(a, h_host);
        }
        if (wide_iq && chunks_aligned<InT, double>(a)) {
            if (wide_iq == 4) return run_groups<InT, double, L, 4, false, 2, 2>(a, h_host);
            return run_groups<InT, double, L, 2, false, 2, 2>(a, h_host);
        }
        return run_groups<InT, double, L, 4, false, 2>(a, h_host);
    } else if (a.out_dtype == FIRGPU_F32) {
        // 1024-output warp tiles for float32 output (twice the arithmetic per byte written of the float64 form: more
        // independent work per lane pays): 2 / 4 / 8 / 16 groups per lane 0.73 / 0.78 / 0.83 / 0.69 of the HBM peak
        static const bool four = getenv("FIRGPU_IDEAL_F32_GROUPS4") != nullptr;        // A/B knob: the round-1 tile
        if (four) return run_groups<InT, float, L, 4, false>(a, h_host);
        // Wide groups (8 outputs = one 256-bit store per lane) wherever the chunks are aligned: uint8 2^26 samples 0.86 -> 0.99
        // of the HBM peak, int16 0.88 -> 0.97, 8192 x 16384 rows 0.82 -> 0.98 (profiles/README.md R2.9b).  FIRGPU_IDEAL_F32_WIDE=0
        // keeps the 16-byte form for A/B.
        static const int wide = getenv("FIRGPU_IDEAL_F32_WIDE") ? atoi(getenv("FIRGPU_IDEAL_F32_WIDE")) : -1;
        if (wide != 0 && chunks_aligned<InT, float>(a)) {
            if (wide == 4 || (wide < 0 && sizeof(InT) == 2 && a.cols >= (1 << 20))) return run_groups<InT, float, L, 4, false, 1, 2>(a, h_host);
            return run_groups<InT, float, L, 2, false, 1, 2>(a, h_host);
        }
        return run_groups<InT, float, L, 8, false>(a, h_host);
    }
    else {
        // float64 output, wide groups (4 outputs = one 256-bit store per lane), 2 / 4 / 8 groups per lane: 1.01 / 1.00 / 0.96 of
        // the HBM peak on 2^26 samples against 0.97 for the 16-byte form, rows of the 2999 x 4499 image 0.81 / 0.80 / 0.65 (0.69)
        static const int wide64 = getenv("FIRGPU_IDEAL_F64_WIDE") ? atoi(getenv("FIRGPU_IDEAL_F64_WIDE")) : 2;   // 0: A/B, the 16-byte form
        if (wide64 != 0 && chunks_aligned<InT, double>(a)) {
            if (wide64 == 4) return run_groups<InT, double, L, 4, false, 1, 2>(a, h_host);
            return run_groups<InT, double, L, 2, false, 1, 2>(a, h_host);
        }
        return run_groups<InT, double, L, 8, false>(a, h_host);
    }
}

template <typename InT>
int by_taps(const Fir1dArgs &a, const double *h)
{
    switch (a.n_taps) {
    case 1: return run<InT, 1>(a, h);
    case 2: return run<InT, 2>(a, h);
    case 3: return run<InT, 3>(a, h);
    case 4: return run<InT, 4>(a, h);
    case 5: return run<InT, 5>(a, h);
    case 6: return run<InT, 6>(a, h);
    case 7: return run<InT, 7>(a, h);
    case 8: return run<InT, 8>(a, h);
    case 9: return run<InT, 9>(a, h);
    default: return 0;
    }
}

}  // namespace

// 1 = served, 0 = not eligible (caller uses the generic kernel), else error.
int try_launch_fir1d_ideal_fast(const Fir1dArgs &a, const double *h_host)
{
    if ((a.out_dtype != FIRGPU_F64 && a.out_dtype != FIRGPU_F32) || (a.dilation != 1 && a.dilation != 2) || a.n_taps > kMaxFastIdealTaps) return 0;
    if (force_generic()) return 0;
    // the C++ Fir1D model: causal, raw float64 samples, float64 output
    if (a.raw && a.centre == 0 && a.in_dtype == FIRGPU_F64 && a.out_dtype == FIRGPU_F64 && a.dilation == 1) return by_taps<double>(a, h_host);
    if (!a.is_reference_model()) return 0;
    if (a.in_dtype == FIRGPU_U8) return by_taps<uint8_t>(a, h_host);
    if (a.in_dtype == FIRGPU_I16) return by_taps<int16_t>(a, h_host);
    return 0;
}

}  // namespace firgpu
