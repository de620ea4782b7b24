// Tiled 2-D fixed-point FIR for sm_100a (device code and launch templates):
// the fast path behind firgpu_fir2d_fixed for windows of up to 7 x 7 taps on uint8 /
// int16 images.  The reference has no fir_2d; arithmetic = the 1-D model's
// pre/post-processing (fir_1d/model/python/fir_1d_fixed_ref.py:79-81,110-126)
// around y[i,j] = sum_m sum_n hq[m][n] * xs[i-m+kh/2][j-n+kw/2].
//
// Design (DESIGN.md "fir2d_tiled"):
//   * a CUtensorMap over [width, rows, images] describes the batch; TMA
//     (cp.async.bulk.tensor.3d -> UTMALDG) stages boxes of 256 columns x BH
//     rows; out-of-bounds box elements are ZERO-FILLED by the TMA unit, which
//     is exactly the reference's zero padding at the four image borders — the
//     kernel has no border code at all.  Row-band shards expose their halo
//     rows by shifting the map's origin (halo_top_valid / halo_bottom_valid).
//   * a CTA owns a vertical strip (2 KiB of columns) of one row band and
//     streams down it, kStages2 TMA stages in flight; each thread owns 16
//     bytes of columns and keeps the last KH input rows in registers as
//     byte-packed samples, so every staged row is read from shared memory once
//     and nothing is re-read vertically inside a band;
//   * per output row: KH x ceil((KW+1)/2) dp2a per pixel (general kernels), or
//     — when the integer kernel is exactly rank-1, hq = a (x) b, proved on the
//     host — one dp2a row pass per new input row plus KH IMADs per pixel;
//   * post stage (rounding bias, shift-by-16 as a byte permute, packed
//     saturate) and the exactness proof are shared with the 1-D kernel.
#pragma once
#include "fir1d_staged.cuh"

#include <cuda.h>

#include <algorithm>

namespace firgpu {
namespace tiled {

using staged::dp2a_hi;
using staged::dp2a_lo;
using staged::lds_words;
using staged::mbar_expect_tx;
using staged::mbar_init;
using staged::mbar_wait;

// The producer lane's wait for a free stage.  A bare try_wait loop returned every ~15 ns: the three producer warps of an SM
// issued 16 % of the kernel's warp-instructions and took 14 % of its stall samples
// (profiles/r2_ncu_fir2d_5x5_general_source_hotspots.txt).  With a suspend-time hint and a back-off between polls they
// stay out of the schedulers' way.  Measured effect on throughput: none (general 5x5 0.578 before and after, 3x3 0.91) —
// the consumers were not waiting for issue slots — kept because it costs nothing.
__device__ __forceinline__ void mbar_wait_patient(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (done) break;
        __nanosleep(128);
    }
}
using staged::post_pair;
using staged::post_pair_mode;
using staged::POST_HI16;
using staged::POST_SHIFT;
using staged::POST_WRAP;
using staged::smem_u32;
using staged::store_v4;

// Tunables (overridable only by tools/sweep_tiled.cu).
#ifndef FIRGPU_T2_THREADS
#define FIRGPU_T2_THREADS 128
#endif
#ifndef FIRGPU_T2_STAGES
#define FIRGPU_T2_STAGES 2
#endif
#ifndef FIRGPU_T2_RPS5
#define FIRGPU_T2_RPS5 2          // stage rows = KH * RPS for KH >= 4
#endif
#ifndef FIRGPU_T2_RPS3
#define FIRGPU_T2_RPS3 3          // ... for KH == 3
#endif
#ifndef FIRGPU_T2_MINBLOCKS
#define FIRGPU_T2_MINBLOCKS 3
#endif
#ifndef FIRGPU_T2_FFMA
#define FIRGPU_T2_FFMA 0          // 1: also build the general 5x5 kernel with 1 / 2 kernel rows on FFMA (FIRGPU_T2_FFMA_ROWS selects
                                  // them at run time).  Bit-exact but measured SLOWER (0.57 -> 0.50 / 0.48 of the HBM peak,
                                  // profiles/r2_ab_fir2d_5x5_ffma_rows.jsonl), so the shipped library does not carry them.
#endif
#ifndef FIRGPU_T2_WARPSPEC
#define FIRGPU_T2_WARPSPEC 1      // 1: dedicated TMA producer warp + per-stage empty barriers; 0: __syncthreads per stage
#endif
constexpr int kThreads2 = FIRGPU_T2_THREADS;
constexpr int kStages2 = FIRGPU_T2_STAGES;
constexpr int kBlock2 = kThreads2 + (FIRGPU_T2_WARPSPEC ? 32 : 0);   // consumer threads (+ one producer warp)
constexpr int kBandTarget = 64;                  // output rows per band, about
#ifndef FIRGPU_T2_RPS7
#define FIRGPU_T2_RPS7 1          // ... for KH == 7 (measured: one row group per stage 972 / 435 Gpixel/s rank-1 / general 7x7, two 758 / 358)
#endif
inline constexpr int rows_per_stage(int kh) { return kh == 1 ? 8 : kh * (kh >= 6 ? FIRGPU_T2_RPS7 : kh >= 4 ? FIRGPU_T2_RPS5 : FIRGPU_T2_RPS3); }
constexpr int kStripBytes = kThreads2 * 16;      // staged bytes per row of a strip
constexpr int kBoxElems = 256;                   // TMA box width limit (elements)
constexpr int kMapElemBytes = 2;                 // the tensor map always describes 16-bit elements: int16 pixels, or PAIRS of
                                                 // uint8 pixels (even widths), so that a box row is 512 bytes for both
constexpr int kMaxKH = 7;
constexpr int kMaxPairs2 = 4;                    // KW <= 7

struct TiledParams {
    unsigned char *y;
    long long ld_y, img_stride_y;                // elements
    int width, height, n_img;
    int n_strips, n_bands, band_rows, stages_per_item;
    int halo_top;                                // map row of image row 0
    int shift, bias, y_vec_ok;
    int total_items;
    // WRAP post (live acc_bits wrap, see staged::post_pair_mode): same fields as StagedParams
    int wrap_sh, wrap_split, wrap_mul;
    uint32_t post_bias, frac_mask;
    uint32_t cp_even[kMaxKH][kMaxPairs2];        // general: dp2a tap pairs of kernel row m
    uint32_t cp_odd[kMaxKH][kMaxPairs2];
    uint32_t xp_even[2], xp_odd[2];              // general 5x5: the odd tap of kernel rows (0,1) and (2,3) paired ACROSS rows
    uint32_t rp_even[kMaxPairs2], rp_odd[kMaxPairs2];   // separable: row-pass pairs (vector b)
    int32_t col_taps[kMaxKH];                    // separable: column taps (vector a, pre-scaled)
    int col_sym;                                 // separable: a[m] == a[KH-1-m] for every m (host-checked)
    // general 5x5, FR > 0: the newest FR kernel rows run as FFMA on the other FMA pipe (see fir2d_tiled_kernel)
    int ffma_rows;                               // FR chosen by the host (0: all 25 taps on dp2a)
    float ftap[2][5];                            // UNSCALED integer taps of kernel rows 0 / 1 by sample offset (-2 .. +2)
    float fbase;                                 // 2^23 + 255 * sum of the negative taps' magnitudes
    uint32_t acc0;                               // accumulator start: bias - (bits(fbase) << fsh)  (mod 2^32)
    int fsh;                                     // the dp2a taps' pre-scale (16 - frac_bits in the shift-by-16 form, else 0)
};

template <typename T, int KH, int KW> struct Geo2 {
    using G1 = staged::Geo<T, KW, 1>;            // horizontal geometry = the 1-D kernel's
    static constexpr int ES = sizeof(T);
    static constexpr int BH = rows_per_stage(KH);   // rows per TMA stage (multiple of KH)
    static constexpr int RPS = BH / KH;
    static constexpr int BOXB = kBoxElems * kMapElemBytes;   // bytes per box row
    static constexpr int NBX = kStripBytes / BOXB;
    static constexpr int STAGE = BH * kStripBytes;
    static constexpr int PADS = 16 / ES;         // columns of pad staged on each side
    static constexpr int OUTW = kStripBytes / ES - 2 * PADS;    // output columns per strip
    static constexpr int ACTIVE = OUTW / G1::SC; // threads that own output columns
    static constexpr int SMEM = kStages2 * STAGE + 2 * kStages2 * 8;
    static constexpr int UP = KH - 1 - KH / 2;   // input rows above the output row
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// Byte offset inside a stage (row 0) of logical byte b of the strip row: the
// stage is NBX boxes, each [BH][BOXB].
template <typename G> __device__ __forceinline__ uint32_t stage_offset(uint32_t b)
{
    return (b / G::BOXB) * (G::BH * G::BOXB) + (b % G::BOXB);
}

// acc[i] += taps (x) B for one row of byte-packed samples (same pairing rule
// as staged::mac_stream).
template <typename G1, int L>
__device__ __forceinline__ void mac_row(const uint32_t (&cpe)[kMaxPairs2], const uint32_t (&cpo)[kMaxPairs2],
                                        const uint32_t (&B)[G1::NB], int (&acc)[G1::S])
{
#pragma unroll
    for (int i = 0; i < G1::S; ++i) {
        int a = acc[i];
        const int s0 = G1::HL4 + i - G1::lo;
        if ((s0 & 1) == 0) {
#pragma unroll
            for (int k = 0; k < G1::NPE; ++k) {
                const int q = s0 / 2 + k;
                if ((q >> 1) < G1::NB) a = (q & 1) ? dp2a_hi(cpe[k], B[q >> 1], a) : dp2a_lo(cpe[k], B[q >> 1], a);
            }
        } else {
#pragma unroll
            for (int k = 0; k < G1::NPO; ++k) {
                const int q = (s0 - 1) / 2 + k;
                if ((q >> 1) < G1::NB) a = (q & 1) ? dp2a_hi(cpo[k], B[q >> 1], a) : dp2a_lo(cpo[k], B[q >> 1], a);
            }
        }
        acc[i] = a;
    }
}

// The same without the tap that has no partner inside its kernel row (5-wide rows: the last tap of a window that
// starts on an even sample, the first tap of one that starts on an odd sample): two dp2a per pixel instead of three.
template <typename G1, int L>
__device__ __forceinline__ void mac_row_pairs(const uint32_t (&cpe)[kMaxPairs2], const uint32_t (&cpo)[kMaxPairs2],
                                              const uint32_t (&B)[G1::NB], int (&acc)[G1::S])
{
    static_assert(L == 5, "cross-row pairing is built for 5-wide kernel rows");
#pragma unroll
    for (int i = 0; i < G1::S; ++i) {
        int a = acc[i];
        const int s0 = G1::HL4 + i - G1::lo;
        if ((s0 & 1) == 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int q = s0 / 2 + k;
                a = (q & 1) ? dp2a_hi(cpe[k], B[q >> 1], a) : dp2a_lo(cpe[k], B[q >> 1], a);
            }
        } else {
#pragma unroll
            for (int k = 1; k < 3; ++k) {
                const int q = (s0 - 1) / 2 + k;
                a = (q & 1) ? dp2a_hi(cpo[k], B[q >> 1], a) : dp2a_lo(cpo[k], B[q >> 1], a);
            }
        }
        acc[i] = a;
    }
}

// ... and those left-over taps of TWO vertically adjacent kernel rows in one dp2a: V holds the two input rows'
// bytes interleaved (V[2q] = [x.b0, y.b0, x.b1, y.b1], V[2q+1] = [x.b2, y.b2, x.b3, y.b3] of byte-quad q, x = the newer
// row), `even` / `odd` the two rows' left-over taps for windows starting on an even / odd sample.  25 taps then cost
// 13 dp2a per pixel instead of 15 — the dp2a pipe is what bounds a general 5x5 kernel.
template <typename G1>
__device__ __forceinline__ void mac_cross(uint32_t even, uint32_t odd, const uint32_t (&V)[2 * G1::NB], int (&acc)[G1::S])
{
#pragma unroll
    for (int i = 0; i < G1::S; ++i) {
        const int s0 = G1::HL4 + i - G1::lo;
        const bool ev = (s0 & 1) == 0;
        const int sg = ev ? s0 + 4 : s0;                 // the sample that left-over tap multiplies
        const uint32_t x = V[2 * (sg >> 2) + ((sg & 3) >> 1)];
        acc[i] = (sg & 1) ? dp2a_hi(ev ? even : odd, x, acc[i]) : dp2a_lo(ev ? even : odd, x, acc[i]);
    }
}

// FR (general 5x5 only; an experiment kept for the record, compiled with -DFIRGPU_T2_FFMA=1 only): the dp2a pipe ("heavy"
// FMA pipe, 64 lanes / clk / SM, shared with IMAD) is the busiest unit of this kernel (70 %), and FFMA issues on the
// other FMA pipe alongside it (profiles/r1_ubench_pipes.jsonl: dp2a + ffma 119.5 lanes / clk / SM).
// The newest FR input rows (kernel rows 0 .. FR-1) are therefore accumulated in float, EXACTLY: samples are integers
// 0 .. 255, taps integers, and the host admits FR > 0 only when 255 * sum|taps of those rows| < 2^23, so that every
// partial sum, started at fbase = 2^23 + 255 * sum|negative taps|, is an integer in [2^23, 2^24) — representable, hence
// every FFMA result is exact whatever the order — and the float's low 23 bits ARE the integer sum.  Its bits are added
// (shifted by the dp2a taps' pre-scale) to the int32 accumulator, whose start value carries minus the constant part.
// Kernel rows FR .. 4 stay on dp2a: 10 per pixel (FR = 1) or 8 (FR = 2) instead of 13.
// Measured (profiles/r2_ab_fir2d_5x5_ffma_rows.jsonl): bit-equal outputs, 15 - 20 % SLOWER — 3 dp2a per pixel become
// 5 FFMA + 1.5 byte-to-float conversions (PRMT + FADD) + 1 add, and the kernel is not bound by the pipe's throughput but
// by how many instructions three warps per scheduler get issued.  FR = 0 is what ships.
template <typename T, int KH, int KW, int POST, bool SEP, typename OutT = T, int FR = 0, bool SYM = false>
__global__ void __launch_bounds__(kBlock2, (KH > 5 && SEP && sizeof(T) == 1 ? 2 : FIRGPU_T2_MINBLOCKS))   // 7 rows x 16 horizontal sums
fir2d_tiled_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TiledParams p)
{
    using G = Geo2<T, KH, KW>;
    using G1 = typename G::G1;
    constexpr int ES = G::ES;
    constexpr int OS = sizeof(OutT);
    constexpr int NOUT = G1::SC * OS / 4;         // output words per thread per row: 16 bytes, or 8 for int16 -> uint8

    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + kStages2 * G::STAGE;
    const int tid = threadIdx.x;

    int n_my = 0;
    if ((int)blockIdx.x < p.total_items) n_my = (p.total_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
    const int spi = p.stages_per_item;

    const uint32_t empty_base = bar_base + 8 * kStages2;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages2; ++s) {
            mbar_init(bar_base + 8 * s, 1);                   // "full": TMA transaction barrier
            mbar_init(empty_base + 8 * s, kThreads2 / 32);    // "empty": one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // (No programmatic dependent launch here: measured on row-band shards of the 8- and 4-GPU plans it changes nothing for
    // the 5-row kernels and costs 15-20 % for 3x3 int16, whose resident CTAs per SM are capped below the occupancy limit —
    // the next grid's CTAs then land unevenly in the free slots.  profiles/r2_shard2d_stages_per_item_{pdl,nopdl}.jsonl)

    // item -> (image, band, strip); consecutive items are neighbouring strips of
    // one band, so bands that share halo rows are processed at the same time
    // and the shared rows are L2 hits.
    auto decode = [&](int ordinal, int &img, int &row0, int &col0) {
        const int item = (int)blockIdx.x + ordinal * (int)gridDim.x;
        const int strip = item % p.n_strips;
        const int t = item / p.n_strips;
        const int band = t % p.n_bands;
        img = t / p.n_bands;
        row0 = band * p.band_rows;
        col0 = strip * G::OUTW;
    };

    // ---- producer cursor (thread 0) ------------------------------------------------
    int p_item = 0, p_j = 0, p_stage = 0;
    auto issue = [&]() {
        int img, row0, col0;
        decode(p_item, img, row0, col0);
        const uint32_t bar = bar_base + 8 * p_stage;
        const uint32_t dst = smem_base + p_stage * G::STAGE;
        mbar_expect_tx(bar, G::STAGE);
        const int c1 = row0 - G::UP + p_j * G::BH + p.halo_top;
#pragma unroll
        for (int bx = 0; bx < G::NBX; ++bx)
            tma_load_3d(dst + bx * (G::BH * G::BOXB), &tmap, (col0 - G::PADS) * ES / kMapElemBytes + bx * kBoxElems, c1, img, bar);
        if (++p_j == spi) {
            p_j = 0;
            ++p_item;
        }
        p_stage = p_stage + 1 == kStages2 ? 0 : p_stage + 1;
    };
    const long long total_stages = (long long)n_my * spi;
#if FIRGPU_T2_WARPSPEC
    if (tid >= kThreads2) {
        // Producer warp: one lane walks the whole stage sequence, refilling a
        // stage as soon as every consumer warp has released it.
        if (tid == kThreads2) {
            uint32_t empty_phase = 0;
            for (long long q = 0; q < total_stages; ++q) {
                if (q >= kStages2) {
                    mbar_wait_patient(empty_base + 8 * p_stage, (empty_phase >> p_stage) & 1u);
                    empty_phase ^= 1u << p_stage;
                }
                issue();
            }
        }
        return;
    }
#else
    if (tid == 0) {
        for (int i = 0; i < kStages2 && i < total_stages; ++i) issue();
    }
#endif

    // ---- per-thread constants ---------------------------------------------------------
    const bool active = tid < G::ACTIVE;
    const uint32_t b_main = 16 + tid * 16;
    const uint32_t off_m = stage_offset<G>(b_main);
    const uint32_t off_l = stage_offset<G>(b_main - G1::HLB);
    const uint32_t off_r = stage_offset<G>(b_main + 16);

    uint32_t cpe[SEP ? 1 : KH][kMaxPairs2], cpo[SEP ? 1 : KH][kMaxPairs2];
    if constexpr (SEP) {
#pragma unroll
        for (int k = 0; k < kMaxPairs2; ++k) {
            cpe[0][k] = p.rp_even[k];
            cpo[0][k] = p.rp_odd[k];
        }
    } else {
#pragma unroll
        for (int m = 0; m < KH; ++m)
#pragma unroll
            for (int k = 0; k < kMaxPairs2; ++k) {
                cpe[m][k] = p.cp_even[m][k];
                cpo[m][k] = p.cp_odd[m][k];
            }
    }

    constexpr bool col_sym = SEP && SYM;          // separable with a symmetric column vector (host-checked, p.col_sym)
    // rolling window of the last KH input rows
    uint32_t Bw[SEP ? 1 : KH][G1::NB];            // general: byte-packed samples
    int Rw[SEP ? KH : 1][G1::S];                  // separable: horizontal sums
    constexpr bool kCross = !SEP && KH == 5 && KW == 5;      // general 5x5: left-over taps paired across kernel rows
    uint32_t Vw[kCross ? KH : 1][2 * G1::NB];     // ... row rr interleaved bytewise with the row before it
    static_assert(FR == 0 || (kCross && FR <= 2), "FFMA rows exist for the general 5x5 kernel only");
    constexpr int NC = kCross ? (KH - FR) / 2 : 0;           // cross-paired kernel-row pairs (FR + 2j, FR + 2j + 1)
    constexpr int NF = G1::S + KW - 1;                       // float samples per row: offsets -lo .. S - 1 + hi
    [[maybe_unused]] float Fprev[FR > 1 ? NF : 1];           // FR = 2: the previous row's float samples

    int stage = 0;
    uint32_t phase_bits = 0;
    for (int k = 0; k < n_my; ++k) {
        int img, row0, col0;
        decode(k, img, row0, col0);
        const int col = col0 + tid * G1::SC;
        int rows_here = p.height - row0;
        if (rows_here > p.band_rows) rows_here = p.band_rows;
        OutT *ycol = reinterpret_cast<OutT *>(p.y) + (long long)img * p.img_stride_y + (long long)row0 * p.ld_y + col;
        const bool vec = p.y_vec_ok && col + G1::SC <= p.width;
        const int room = p.width - col;            // columns this thread may write
        for (int j = 0; j < spi; ++j) {
            mbar_wait(bar_base + 8 * stage, (phase_bits >> stage) & 1u);
            phase_bits ^= 1u << stage;
            if (active && room > 0) {
                const uint32_t st = smem_base + stage * G::STAGE;
#pragma unroll 1
                for (int g = 0; g < G::RPS; ++g) {
#pragma unroll
                    for (int rr = 0; rr < KH; ++rr) {
                        const int r = g * KH + rr;                    // row inside the stage
                        const int o = j * G::BH + r - (KH - 1);       // output row inside the band
                        uint32_t raw[G1::NR];
                        const uint32_t ro = st + r * G::BOXB;
                        lds_words<G1::HLB / 4>(ro + off_l, &raw[0]);
                        lds_words<4>(ro + off_m, &raw[G1::HLB / 4]);
                        lds_words<G1::HRB / 4>(ro + off_r, &raw[G1::HLB / 4 + 4]);
                        if constexpr (ES == 2) {
#pragma unroll
                            for (int q = 0; q < G1::NR; ++q) raw[q] = __vimin_s16x2_relu(raw[q], 0x00ff00ffu);
                        }
                        uint32_t Bnew[G1::NB];
#pragma unroll
                        for (int q = 0; q < G1::NB; ++q) {
                            if constexpr (ES == 1) Bnew[q] = raw[q];
                            else Bnew[q] = __byte_perm(raw[2 * q], raw[2 * q + 1], 0x6420u);
                        }
                        if constexpr (SEP) {
#pragma unroll
                            for (int i = 0; i < G1::S; ++i) Rw[rr][i] = 0;
                            mac_row<G1, KW>(cpe[0], cpo[0], Bnew, Rw[rr]);
                        } else {
#pragma unroll
                            for (int q = 0; q < G1::NB; ++q) Bw[rr][q] = Bnew[q];
                            if constexpr (kCross) {
                                const int prev = (rr + KH - 1) % KH;
#pragma unroll
                                for (int q = 0; q < G1::NB; ++q) {
                                    Vw[rr][2 * q] = __byte_perm(Bnew[q], Bw[prev][q], 0x5140u);
                                    Vw[rr][2 * q + 1] = __byte_perm(Bnew[q], Bw[prev][q], 0x7362u);
                                }
                            }
                        }
                        [[maybe_unused]] float Fnew[FR > 0 ? NF : 1];      // this row's samples as floats: 2^23 + byte, minus 2^23
                        if constexpr (FR > 0) {
#pragma unroll
                            for (int e = 0; e < NF; ++e) {
                                const int sidx = G1::HL4 - G1::lo + e;  // byte of the packed row
                                const uint32_t bits = __byte_perm(Bnew[sidx >> 2], 0x4b000000u, 0x7440u | (uint32_t)(sidx & 3));
                                Fnew[e] = __uint_as_float(bits) - 8388608.0f;
                            }
                        }
                        if (o >= 0 && o < rows_here) {
                            int acc[G1::S];
#pragma unroll
                            for (int i = 0; i < G1::S; ++i) acc[i] = FR > 0 ? (int)p.acc0 : p.bias;
                            [[maybe_unused]] float accf[FR > 0 ? G1::S : 1];
                            if constexpr (FR > 0) {
#pragma unroll
                                for (int i = 0; i < G1::S; ++i) accf[i] = p.fbase;
#pragma unroll
                                for (int m = 0; m < FR; ++m)
#pragma unroll
                                    for (int t = 0; t < KW; ++t)
#pragma unroll
                                        for (int i = 0; i < G1::S; ++i)
                                            accf[i] = __fmaf_rn(p.ftap[m][t], m == 0 ? Fnew[i + t] : Fprev[i + t], accf[i]);
                            }
#pragma unroll
                            for (int m = FR; m < KH; ++m) {
                                const int slot = (rr - m + KH) % KH;   // input row (newest - m) <-> kernel row m
                                if constexpr (SEP) {
                                    if constexpr (col_sym) {
                                        // symmetric column vector (every low-pass / binomial / Gaussian kernel): the two rows
                                        // that share a tap are added first — an add on the ALU pipe replaces a multiply-add
                                        // on the heavy FMA pipe, the busiest unit of this kernel
                                        if (m < KH / 2) {
                                            const int mirror = (rr - (KH - 1 - m) + KH) % KH;
#pragma unroll
                                            for (int i = 0; i < G1::S; ++i) acc[i] += p.col_taps[m] * (Rw[slot][i] + Rw[mirror][i]);
                                        } else if (m == KH / 2) {
#pragma unroll
                                            for (int i = 0; i < G1::S; ++i) acc[i] += p.col_taps[m] * Rw[slot][i];
                                        }
                                    } else {
#pragma unroll
                                        for (int i = 0; i < G1::S; ++i) acc[i] += p.col_taps[m] * Rw[slot][i];
                                    }
                                } else if constexpr (kCross) {
                                    if (m < FR + 2 * NC) mac_row_pairs<G1, KW>(cpe[m], cpo[m], Bw[slot], acc);
                                    else mac_row<G1, KW>(cpe[m], cpo[m], Bw[slot], acc);
                                } else {
                                    mac_row<G1, KW>(cpe[m], cpo[m], Bw[slot], acc);
                                }
                            }
                            if constexpr (kCross) {       // kernel rows (FR+2j, FR+2j+1) <-> input rows (rr-FR-2j, rr-FR-2j-1)
#pragma unroll
                                for (int j = 0; j < NC; ++j)
                                    mac_cross<G1>(p.xp_even[j], p.xp_odd[j], Vw[(rr + 2 * KH - FR - 2 * j) % KH], acc);
                            }
                            if constexpr (FR > 0) {
#pragma unroll
                                for (int i = 0; i < G1::S; ++i)
                                    acc[i] = (int)((uint32_t)acc[i] + (__float_as_uint(accf[i]) << p.fsh));
                            }
                            uint32_t V[G1::S / 2];
#pragma unroll
                            for (int q = 0; q < G1::S / 2; ++q) V[q] = post_pair_mode<POST>(acc[2 * q], acc[2 * q + 1], p);
                            uint32_t O[NOUT];
                            if constexpr (OS == 2) {
#pragma unroll
                                for (int q = 0; q < NOUT; ++q) O[q] = V[q];
                            } else {
#pragma unroll
                                for (int q = 0; q < NOUT; ++q) O[q] = __byte_perm(V[2 * q], V[2 * q + 1], 0x6420u);
                            }
                            OutT *dst = ycol + (long long)o * p.ld_y;
                            if (vec) {
                                if constexpr (NOUT == 4) store_v4(dst, O[0], O[1], O[2], O[3]);
                                else *reinterpret_cast<uint2 *>(dst) = make_uint2(O[0], O[1]);
                            } else {
#pragma unroll
                                for (int e = 0; e < G1::SC; ++e) {
                                    if (e < room) {
                                        if constexpr (OS == 2) dst[e] = (OutT)((O[e / 2] >> (16 * (e & 1))) & 0xffffu);
                                        else dst[e] = (OutT)((O[e / 4] >> (8 * (e & 3))) & 0xffu);
                                    }
                                }
                            }
                        }
                        if constexpr (FR > 1) {
#pragma unroll
                            for (int e = 0; e < NF; ++e) Fprev[e] = Fnew[e];
                        }
                    }
                }
            }
#if FIRGPU_T2_WARPSPEC
            __syncwarp();                                  // this warp is done with the stage
            if ((tid & 31) == 0) staged::mbar_arrive(empty_base + 8 * stage);
#else
            __syncthreads();                               // every thread is done with this stage
            if (tid == 0 && (long long)k * spi + j + kStages2 < total_stages) issue();
#endif
            stage = stage + 1 == kStages2 ? 0 : stage + 1;
        }
    }
}

// ---- launch helpers -------------------------------------------------------------
template <typename T, int KH, int KW, int POST, bool SEP, typename OutT = T, int FR = 0, bool SYM = false>
int launch_tiled_one(const CUtensorMap &map, const TiledParams &p, cudaStream_t stream)
{
    using G = Geo2<T, KH, KW>;
    auto kern = fir2d_tiled_kernel<T, KH, KW, POST, SEP, OutT, FR, SYM>;
    static std::mutex mu;
    static int ctas_per_sm[64] = {0};
    int dev = 0;
    FIRGPU_CUDA(cudaGetDevice(&dev));
    int per_sm, sms;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev < 64 && ctas_per_sm[dev] == 0) {
            FIRGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
            int n = 0;
            FIRGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kBlock2, G::SMEM));
            ctas_per_sm[dev] = n < 1 ? 1 : n;
        }
        per_sm = dev < 64 ? ctas_per_sm[dev] : 1;
    }
    // Measured on B200 (profiles/r1_sweep_tunables.md): two stages in flight; 3-row
    // int16 kernels are fastest with 3 resident CTAs per SM; 5-row and all uint8 kernels
    // (twice the arithmetic per byte) with all that fit.
    if (KH <= 3 && sizeof(T) == 2 && per_sm > 3) per_sm = 3;
    if (const char *e = getenv("FIRGPU_CTAS_PER_SM")) {       // (tools/sweep_tiled.cu changes it between launches: not cached)
        const int forced = atoi(e);
        if (forced >= 1 && forced < per_sm) per_sm = forced;
    }
    sms = staged::sm_count(dev);
    long long grid = (long long)sms * per_sm;
    if (grid > p.total_items) grid = p.total_items;
    // a few items per CTA (a row-band shard of an 8-GPU plan, a small image): hide the launch latency behind the
    // previous kernel's tail; long launches gain nothing from it
    // Band height.  Items (band x strip x image) are dealt to the persistent CTAs in rounds; with only a few rounds —
    // a row-band shard of an 8-GPU plan — a part-filled last round costs up to 20 %.  Choose the stages per item
    // (band = stages * BH - (KH - 1) output rows) that loses least to a part-filled round, to the KH - 1 rows every item
    // loads without producing output, and to a short last band; plenty of rounds keep the default (about 64 rows).
    // Measured on the [4, 2048, 16384] int16 shard (config 4 at N = 8): 3x3 0.856 (default 8 stages) -> 0.917, rank-1 5x5
    // 0.741 -> 0.774 of the HBM peak; the model is coarse (the best forced value reaches 0.93 / 0.82).
    TiledParams q = p;
    int spi_lo = 2, spi_hi = 16;
    if (const char *e = getenv("FIRGPU_T2_SPI")) spi_lo = spi_hi = atoi(e);      // experiments: force the stages per item
    if (p.total_items < 16 * (long long)sms * per_sm || spi_lo == spi_hi) {
        double best = -1.0;
        for (int spi = spi_lo; spi <= spi_hi; ++spi) {
            const int band = spi * G::BH - (KH - 1);
            if (band < 1) continue;
            const long long bands = (p.height + band - 1) / band;
            const long long items = bands * p.n_strips * (long long)p.n_img;
            const long long g = std::min<long long>((long long)sms * per_sm, items);
            const double rounds = (double)((items + g - 1) / g);
            // cost model: every item walks spi stages whatever it holds; useful work = output rows
            const double useful = (double)p.height * p.n_strips * p.n_img / (double)G::BH;      // stages of output rows
            const double spent = rounds * g * spi;
            const double eff = useful / spent;
            if (eff > best) {
                best = eff;
                q.stages_per_item = spi;
                q.band_rows = band;
                q.n_bands = (int)bands;
                q.total_items = (int)items;
            }
        }
        grid = (long long)sms * per_sm;
        if (grid > q.total_items) grid = q.total_items;
    }
    FIRGPU_CUDA(staged::launch_pdl(kern, (unsigned)grid, kBlock2, G::SMEM, stream, false, map, q));
    note_launch(FIRGPU_PATH_STAGED);
    return 1;
}

template <typename T, int KH, int KW, bool SEP, typename OutT, bool SYM = false>
int launch_tiled_post(const CUtensorMap &map, const TiledParams &p, int post, cudaStream_t s)
{
#if FIRGPU_T2_FFMA
    if constexpr (!SEP && KH == 5 && KW == 5) {     // general 5x5: FFMA rows where the host proved them exact
        if (p.ffma_rows == 1) {
            switch (post) {
            case POST_SHIFT: return launch_tiled_one<T, KH, KW, POST_SHIFT, SEP, OutT, 1>(map, p, s);
            case POST_HI16: return launch_tiled_one<T, KH, KW, POST_HI16, SEP, OutT, 1>(map, p, s);
            case POST_WRAP: return launch_tiled_one<T, KH, KW, POST_WRAP, SEP, OutT, 1>(map, p, s);
            default: return 0;
            }
        }
        if (p.ffma_rows == 2) {
            switch (post) {
            case POST_SHIFT: return launch_tiled_one<T, KH, KW, POST_SHIFT, SEP, OutT, 2>(map, p, s);
            case POST_HI16: return launch_tiled_one<T, KH, KW, POST_HI16, SEP, OutT, 2>(map, p, s);
            case POST_WRAP: return launch_tiled_one<T, KH, KW, POST_WRAP, SEP, OutT, 2>(map, p, s);
            default: return 0;
            }
        }
    }
#endif
    switch (post) {
    case POST_SHIFT: return launch_tiled_one<T, KH, KW, POST_SHIFT, SEP, OutT, 0, SYM>(map, p, s);
    case POST_HI16: return launch_tiled_one<T, KH, KW, POST_HI16, SEP, OutT, 0, SYM>(map, p, s);
    case POST_WRAP: return launch_tiled_one<T, KH, KW, POST_WRAP, SEP, OutT, 0, SYM>(map, p, s);
    default: return 0;
    }
}

template <typename T, int KH, int KW, typename OutT = T>
int launch_tiled_modes(const CUtensorMap &map, const TiledParams &p, int post, bool sep, cudaStream_t s)
{
    if constexpr (KH > 1 && KW >= 3 && !(KH == 3 && KW == 3)) {
        if (sep) return p.col_sym ? launch_tiled_post<T, KH, KW, true, OutT, true>(map, p, post, s)
                                  : launch_tiled_post<T, KH, KW, true, OutT, false>(map, p, post, s);
    }
    return launch_tiled_post<T, KH, KW, false, OutT>(map, p, post, s);
}

template <typename T, typename OutT = T>
int launch_tiled_shape(const CUtensorMap &map, const TiledParams &p, int kh, int kw, int post, bool sep, cudaStream_t s)
{
    if (kh == 1 && kw == 3) return launch_tiled_modes<T, 1, 3, OutT>(map, p, post, sep, s);   // short-row batches of the 1-D filter
    if (kh == 1 && kw == 5) return launch_tiled_modes<T, 1, 5, OutT>(map, p, post, sep, s);
    if (kh == 3 && kw == 1) return launch_tiled_modes<T, 3, 1, OutT>(map, p, post, sep, s);
    if (kh == 3 && kw == 3) return launch_tiled_modes<T, 3, 3, OutT>(map, p, post, sep, s);
    if (kh == 3 && kw == 5) return launch_tiled_modes<T, 3, 5, OutT>(map, p, post, sep, s);
    if (kh == 5 && kw == 1) return launch_tiled_modes<T, 5, 1, OutT>(map, p, post, sep, s);
    if (kh == 5 && kw == 3) return launch_tiled_modes<T, 5, 3, OutT>(map, p, post, sep, s);
    if (kh == 5 && kw == 5) return launch_tiled_modes<T, 5, 5, OutT>(map, p, post, sep, s);
#if !FIRGPU_QUICK
    if (kh == 1 && kw == 7) return launch_tiled_modes<T, 1, 7, OutT>(map, p, post, sep, s);
    if (kh == 7 && kw == 1) return launch_tiled_modes<T, 7, 1, OutT>(map, p, post, sep, s);
    if (kh == 3 && kw == 7) return launch_tiled_modes<T, 3, 7, OutT>(map, p, post, sep, s);
    if (kh == 7 && kw == 3) return launch_tiled_modes<T, 7, 3, OutT>(map, p, post, sep, s);
    if (kh == 5 && kw == 7) return launch_tiled_modes<T, 5, 7, OutT>(map, p, post, sep, s);
    if (kh == 7 && kw == 5) return launch_tiled_modes<T, 7, 5, OutT>(map, p, post, sep, s);
    if (kh == 7 && kw == 7) return launch_tiled_modes<T, 7, 7, OutT>(map, p, post, sep, s);
#endif
    return 0;
}

// Does the (kh, kw) shape have a separable instantiation?
inline bool shape_has_separable(int kh, int kw) { return kh > 1 && kw >= 3 && !(kh == 3 && kw == 3); }
#if FIRGPU_QUICK
constexpr int kMaxWindow = 5;
#else
constexpr int kMaxWindow = 7;
#endif
// Instantiated windows: odd heights and widths up to 7 (the QUICK build: 5), except 1x1.  Even sizes ride on the next
// odd one (fir2d_tiled.cu appends a zero row / column, which keeps the reference's centre L // 2).
inline bool shape_supported(int kh, int kw)
{
    return (kh & 1) && (kw & 1) && kh >= 1 && kw >= 1 && kh <= kMaxWindow && kw <= kMaxWindow && !(kh == 1 && kw == 1);
}
int launch_tiled_u8(const CUtensorMap &map, const TiledParams &p, int kh, int kw, int post, bool sep, cudaStream_t s);
int launch_tiled_i16(const CUtensorMap &map, const TiledParams &p, int kh, int kw, int post, bool sep, cudaStream_t s);
int launch_tiled_i16u8(const CUtensorMap &map, const TiledParams &p, int kh, int kw, int post, bool sep, cudaStream_t s);   // int16 pixels in, uint8 out

}  // namespace tiled
}  // namespace firgpu
