// Staged 1-D fixed-point FIR for sm_100a (device code and launch templates): the fast path behind
// firgpu_fir1d_fixed (reference arithmetic:
// fir_1d/model/python/fir_1d_fixed_ref.py:95-128, batched as
// fir_1d/sim/vector/gen_fixed_output.py:44-59).
//
// Design (see DESIGN.md "fir1d_staged"):
//   * persistent CTAs, each walking a strided list of 16 KiB input tiles;
//   * every tile is staged into shared memory together with a 16-byte halo on
//     both sides by ONE 1-D bulk copy (cp.async.bulk -> UBLKCP, the TMA
//     engine) that completes on an mbarrier; kStages tiles are in flight per
//     CTA, so HBM latency is covered without holding registers;
//   * only the bytes a bulk copy cannot express (row ends that need zero
//     padding, rows whose base is not 16-byte aligned) are patched by the
//     threads with guarded loads;
//   * each thread owns 16 contiguous input bytes per sub-step: one LDS.128 plus
//     two small halo loads, one coalesced 128-bit store;
//   * arithmetic is packed: int16 containers are clamped two at a time
//     (vimin_s16x2_relu), samples travel as four bytes per register, and the
//     multiply-accumulate is dp2a (two signed 16-bit taps x two unsigned 8-bit
//     samples per instruction) into an int32 accumulator that starts at the
//     rounding bias.  When the taps can be pre-scaled so that the shift is 16
//     bits, the shift + pack of two outputs is a single byte-permute of the
//     accumulators' upper halves; saturation to [0,255] is again packed.
//   * the host proves, from the actual taps, that neither the reference's
//     acc_bits wrap nor int32 can overflow; otherwise the call is served by
//     the generic int64 kernel (fir1d_generic.cu).
//   * dilation 2 = interleaved I/Q: the byte-permutes de-interleave into an I
//     stream and a Q stream, the same MAC runs on each, and the final permute
//     re-interleaves for free.
// Tensor cores are deliberately not used: 3/5 taps is not a contraction.
#pragma once
#include "common.cuh"

#include <cstdlib>
#include <mutex>
#include <type_traits>

namespace firgpu {
namespace staged {

constexpr int kChunk = 16;                                 // input bytes per thread per sub-step
constexpr int kPad = 16;                                   // halo bytes staged on each side
#ifndef FIRGPU_STORE_CS
#define FIRGPU_STORE_CS 0
#endif

// Tunables of one kernel build.  In-flight bytes per SM (= CTAs/SM x stages x
// tile) is what the HBM system is sensitive to: about 64 KiB per SM is the
// measured optimum (profiles/r1_sweep_tunables.md); more outstanding bulk reads
// lower the achieved DRAM bandwidth of a 50/50 read/write stream.
template <int THREADS, int SUB, int STAGES, int MINBLOCKS, int CTAS_PER_SM> struct Cfg {
    static constexpr int kThreads = THREADS;               // threads per CTA
    static constexpr int kSub = SUB;                       // 16-byte sub-steps per thread per tile
    static constexpr int kStages = STAGES;                 // tiles in flight per CTA
    static constexpr int kMinBlocks = MINBLOCKS;           // __launch_bounds__ register cap
    static constexpr int kCtasPerSm = CTAS_PER_SM;         // persistent CTAs per SM (<= occupancy)
    static constexpr int kTileBytes = THREADS * kChunk * SUB;
    // a stage = tile + a halo pad of PADB bytes on both sides (16 for <= 9 taps, more for long filters)
    __host__ __device__ static constexpr int stage_bytes(int padb) { return ((kTileBytes + 2 * padb + 127) / 128) * 128; }
    __host__ __device__ static constexpr int smem_bytes(int padb) { return STAGES * stage_bytes(padb) + STAGES * 8; }
    static constexpr int kSmemBytes = STAGES * (((kTileBytes + 2 * kPad + 127) / 128) * 128) + STAGES * 8;   // 16-byte pads
};

// Per-container choice for long rows (HBM sweep on B200, 2^28-sample stream, 5 taps;
// profiles/r1_sweep_tunables.md): GB/s of the best configuration in the comment.
template <typename InT, typename OutT> struct TuneFor { using type = Cfg<256, 4, 2, 2, 2>; };
#ifdef FIRGPU_K_THREADS        // tools/sweep_staged.cu builds one explicit configuration
template <> struct TuneFor<int16_t, int16_t> { using type = Cfg<FIRGPU_K_THREADS, FIRGPU_K_SUB, FIRGPU_K_STAGES, FIRGPU_K_MINBLOCKS, 0>; };
template <> struct TuneFor<int16_t, uint8_t> { using type = Cfg<FIRGPU_K_THREADS, FIRGPU_K_SUB, FIRGPU_K_STAGES, FIRGPU_K_MINBLOCKS, 0>; };
template <> struct TuneFor<uint8_t, uint8_t> { using type = Cfg<FIRGPU_K_THREADS, FIRGPU_K_SUB, FIRGPU_K_STAGES, FIRGPU_K_MINBLOCKS, 0>; };
#else
template <> struct TuneFor<int16_t, int16_t> { using type = Cfg<256, 8, 2, 1, 1>; };   // 6592 GB/s
template <> struct TuneFor<int16_t, uint8_t> { using type = Cfg<256, 6, 2, 1, 2>; };   // 6463 GB/s
template <> struct TuneFor<uint8_t, uint8_t> { using type = Cfg<256, 4, 2, 2, 2>; };   // 6310 GB/s
// uint8 in, int16 out writes two bytes per byte read: half the bulk reads in flight (8 KiB tiles, 32 KiB per SM) suit that
// mix — 0.95 of the HBM peak against 0.84 with the 16 KiB default; 4 KiB tiles 0.72, 12 KiB 0.89, 32 KiB x 1 CTA 0.89,
// three stages 0.88, three CTAs 0.86 (profiles/r2_sweep_u8_in_i16_out_tile_configs.txt)
#ifndef FIRGPU_U8I16_SUB
template <> struct TuneFor<uint8_t, int16_t> { using type = Cfg<256, 2, 2, 2, 2>; };
#endif
#ifdef FIRGPU_U8I16_SUB        // tools/build_u8i16_variants.sh: one explicit configuration of the uint8 -> int16 kernel
template <> struct TuneFor<uint8_t, int16_t> { using type = Cfg<256, FIRGPU_U8I16_SUB, FIRGPU_U8I16_STAGES, FIRGPU_U8I16_MINBLOCKS, FIRGPU_U8I16_CTAS>; };
#endif
#endif
// Rows shorter than two long-row tiles: 8 KiB tiles, three CTAs per SM.
using SmallRowCfg = Cfg<256, 2, 3, 3, 3>;
// Flattened batches (row ends inside the kernel, RE = true).
#ifndef FIRGPU_FLAT_SUB
#define FIRGPU_FLAT_SUB 6             // 16-byte sub-steps per thread per tile (tile = 4 KiB x this); measured 4 / 6 / 8: 0.77 / 0.81 / 0.80
#endif
#ifndef FIRGPU_FLAT_CTAS
#define FIRGPU_FLAT_CTAS 2
#endif
using FlatCfg = Cfg<256, FIRGPU_FLAT_SUB, 2, FIRGPU_FLAT_CTAS, FIRGPU_FLAT_CTAS>;
// WRAP post mode on long rows: the post stage costs about four more instructions per sample, and with the eight
// warps per SM of the int16 long-row configuration the kernel becomes latency-bound (ncu: issue 41 %, DRAM 62 %);
// the same 64 KiB in flight per SM as two CTAs of 16 KiB tiles.
#ifndef FIRGPU_WRAP_CFG
#define FIRGPU_WRAP_CFG 256, 4, 2, 2, 2
#endif
using WrapCfg = Cfg<FIRGPU_WRAP_CFG>;

constexpr int kMaxFastTaps = 129;               // dp2a datapath (taps > 9 run in zero-padded buckets, see launch_long)
constexpr int kMaxWideTaps = 9;                 // IMAD / wide datapaths
constexpr int kMaxPairs = 65;
constexpr int kMaxFilters = 4;                 // tap sets applied to one staged input (multi-filter mode)

struct StagedParams {
    const unsigned char *x;
    unsigned char *y;
    long long rows, cols, ld_x, ld_y;          // cols / ld in scalars of the container type
    long long halo_left, halo_right;           // readable neighbour scalars (shard seams)
    long long tiles_per_row, total_tiles;
    int shift;                                 // SHIFT post mode: frac_bits
    int bias;                                  // accumulator start value
    int x_bulk_ok;                             // base and row pitch are 16-byte aligned
    int y_vec_ok;
    int n_filters;                             // 1, or kMaxFilters in multi-filter mode
    long long filter_stride_y;                 // elements between the outputs of consecutive tap sets
    uint32_t cp_even[kMaxFilters][kMaxPairs];  // dp2a tap pairs, window starting on an even sample
    uint32_t cp_odd[kMaxFilters][kMaxPairs];   // ... on an odd sample
    int32_t hq[kMaxWideTaps];                  // IMAD / wide mode taps
    uint32_t in_lo2, in_hi2, out_lo2, out_hi2; // wide mode: clamp bounds, each duplicated into both 16-bit lanes
    // WRAP post mode (live acc_bits wrap, fir_1d_fixed_ref.py:110-118): sign-extend the low acc_bits bits of the
    // int32 sum (wrap_sh = 32 - acc_bits), add the rounding bias WITHOUT re-wrapping, floor-shift, saturate
    int wrap_sh;
    uint32_t post_bias, frac_mask;             // 2^(frac-1), 2^frac - 1
    // flattened batch: the one long row is really rows of `period` scalars, each zero-padded on its own
    // (gen_fixed_output.py:44-59); 0 = rows are what rows / cols say
    long long period;
    uint32_t rem_tile_step, rem_sub_step;      // (grid * tile scalars) % period, (threads * chunk scalars) % period (launch_cfg)
    int wrap_split;                            // WRAP: use the general split-shift form (frac_bits > acc_bits, or frac + 32 - acc_bits < 2)
    int wrap_mul;                              // WRAP: 2^(acc_bits - frac_bits) (multiply-add form)
    int pdl;                                   // launch with programmatic stream serialisation
    // streaming (rows == 1): the halo_left scalars in front of column 0 are not in memory before x but in `hist`
    // (hist[j] is column j - halo_left); the CTA of the last tile copies the row's last halo_left scalars to hist_out
    const unsigned char *hist;
    unsigned char *hist_out;
};

enum { POST_SHIFT = 0, POST_HI16 = 1, POST_WRAP = 2 };

// ---- PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FIRGPU_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FIRGPU_DONE_%=;\n"
        "bra FIRGPU_WAIT_%=;\n"
        "FIRGPU_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine; size and both
// addresses are multiples of 16 bytes.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ int dp2a_lo(uint32_t taps, uint32_t bytes, int acc)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(taps), "r"(bytes), "r"(acc));
    return d;
}
__device__ __forceinline__ int dp2a_hi(uint32_t taps, uint32_t bytes, int acc)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(taps), "r"(bytes), "r"(acc));
    return d;
}
template <int N> __device__ __forceinline__ void lds_words(uint32_t addr, uint32_t *dst);
template <> __device__ __forceinline__ void lds_words<0>(uint32_t, uint32_t *) {}
template <> __device__ __forceinline__ void lds_words<1>(uint32_t addr, uint32_t *dst)
{
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(dst[0]) : "r"(addr));
}
template <> __device__ __forceinline__ void lds_words<2>(uint32_t addr, uint32_t *dst)
{
    asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(dst[0]), "=r"(dst[1]) : "r"(addr));
}
template <> __device__ __forceinline__ void lds_words<4>(uint32_t addr, uint32_t *dst)
{
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(dst[0]), "=r"(dst[1]), "=r"(dst[2]), "=r"(dst[3]) : "r"(addr));
}

__device__ __forceinline__ void store_v4(void *ptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
#if FIRGPU_STORE_CS
    asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(ptr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
#else
    *reinterpret_cast<uint4 *>(ptr) = make_uint4(a, b, c, d);
#endif
}
// 32 bytes per lane in one instruction (STG.E.ENL2.256): a warp writes 1 KiB of whole sectors.  Two 16-byte stores at a
// 32-byte lane stride half-fill every sector twice (uint8 in, int16 out: 0.83 of the HBM peak with them).
__device__ __forceinline__ void store_v8(void *ptr, const uint32_t *o)
{
#if FIRGPU_STORE_CS
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                 "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
#else
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                 "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
#endif
}
__device__ __forceinline__ void store_v2(void *ptr, uint32_t a, uint32_t b)
{
#if FIRGPU_STORE_CS
    asm volatile("st.global.cs.v2.b32 [%0], {%1,%2};" ::"l"(ptr), "r"(a), "r"(b) : "memory");
#else
    *reinterpret_cast<uint2 *>(ptr) = make_uint2(a, b);
#endif
}

// NW 32-bit words ending at a 16-byte boundary (left halo: the unaligned remainder comes first).
template <int NW> __device__ __forceinline__ void lds_before(uint32_t end_addr, uint32_t *dst)
{
    constexpr int R = NW % 4, Q = NW / 4;
    const uint32_t a = end_addr - NW * 4;
    if constexpr (R == 1) lds_words<1>(a, dst);
    if constexpr (R == 2) lds_words<2>(a, dst);
    if constexpr (R == 3) {
        lds_words<1>(a, dst);
        lds_words<2>(a + 4, dst + 1);
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) lds_words<4>(a + R * 4 + q * 16, dst + R + q * 4);
}
// NW 32-bit words starting at a 16-byte boundary (right halo).
template <int NW> __device__ __forceinline__ void lds_after(uint32_t addr, uint32_t *dst)
{
    constexpr int R = NW % 4, Q = NW / 4;
#pragma unroll
    for (int q = 0; q < Q; ++q) lds_words<4>(addr + q * 16, dst + q * 4);
    if constexpr (R == 1) lds_words<1>(addr + Q * 16, dst + Q * 4);
    if constexpr (R == 2) lds_words<2>(addr + Q * 16, dst + Q * 4);
    if constexpr (R == 3) {
        lds_words<2>(addr + Q * 16, dst + Q * 4);
        lds_words<1>(addr + Q * 16 + 8, dst + Q * 4 + 2);
    }
}

__host__ __device__ constexpr long long floor_div(long long a, long long b)
{
    return a >= 0 ? a / b : -((-a + b - 1) / b);
}

// ---- compile-time geometry of one (taps, dilation, container) combination ----
template <typename InT, int L, int DIL> struct Geo {
    static constexpr int ES = sizeof(InT);
    static constexpr int hi = L / 2;              // taps reaching towards higher indices
    static constexpr int lo = L - 1 - hi;         // ... lower indices
    static constexpr int HL4 = (lo + 3) / 4 * 4;  // stream halo, whole byte-quads
    static constexpr int HR4 = (hi + 3) / 4 * 4;
    static constexpr int SC = kChunk / ES;        // scalars per chunk
    static constexpr int S = SC / DIL;            // samples per stream per chunk
    static constexpr int HLB = HL4 * DIL * ES;    // halo bytes actually loaded
    static constexpr int HRB = HR4 * DIL * ES;
    static constexpr int NR = (HLB + kChunk + HRB) / 4;   // raw 32-bit words per chunk
    static constexpr int NB = (HL4 + S + HR4) / 4;        // byte-quads per stream
    static constexpr int NPE = (L + 1) / 2;
    static constexpr int NPO = L / 2 + 1;
    static constexpr int PADB = ((HLB > HRB ? HLB : HRB) + 15) / 16 * 16 < kPad ? kPad : ((HLB > HRB ? HLB : HRB) + 15) / 16 * 16;
    static constexpr int PADS = PADB / ES;        // scalars per staged pad
    static_assert(L >= 1 && L <= kMaxFastTaps, "tap count outside the staged kernel's range");
};

struct TilePlan {
    long long row, col0;            // tile origin (scalars)
    long long bulk_lo, bulk_hi;     // scalar range moved by the bulk copy (global columns)
    long long need_hi;              // staged data is consumed up to this column
};

template <typename G, int TS>
__device__ __forceinline__ TilePlan plan_tile(const StagedParams &p, long long row, unsigned tcol)
{
    TilePlan t;
    t.row = row;
    t.col0 = (long long)tcol * TS;
    const long long g0 = t.col0 - G::PADS, g1 = t.col0 + TS + G::PADS;
    const long long vlo = -p.halo_left, vhi = p.cols + p.halo_right;
    long long in_tile = p.cols - t.col0;
    if (in_tile > TS) in_tile = TS;
    t.need_hi = t.col0 + (in_tile + G::SC - 1) / G::SC * G::SC + G::PADS;
    if (t.need_hi > g1) t.need_hi = g1;
    constexpr int Q = 16 / G::ES;                 // scalars per 16 bytes
    long long lo = floor_div((p.hist ? 0 : vlo) + Q - 1, Q) * Q, hi = floor_div(vhi, Q) * Q;   // (a history buffer is not in front of x)
    if (lo < g0) lo = g0;
    if (hi > t.need_hi) hi = t.need_hi;
    if (!p.x_bulk_ok || hi <= lo) lo = hi = g0;
    t.bulk_lo = lo;
    t.bulk_hi = hi;
    return t;
}

// Walk (row, tile-in-row) forward by `step` tiles; tiles_per_row < 2^31.
__device__ __forceinline__ void advance_tile(long long &row, unsigned &tcol, unsigned step, unsigned tiles_per_row)
{
    tcol += step;
    if (tcol >= tiles_per_row) {
        const unsigned q = tcol / tiles_per_row;
        row += q;
        tcol -= q * tiles_per_row;
    }
}

// One stream of byte-packed samples -> S accumulators.
template <typename G, int L, int MAC>
__device__ __forceinline__ void mac_stream(const StagedParams &p, const int f, const uint32_t (&B)[G::NB], int (&acc)[G::S])
{
#pragma unroll
    for (int i = 0; i < G::S; ++i) {
        int a = p.bias;
        if constexpr (MAC == 0) {
            const int s0 = G::HL4 + i - G::lo;
            if ((s0 & 1) == 0) {
#pragma unroll
                for (int k = 0; k < G::NPE; ++k) {
                    const int q = s0 / 2 + k;
                    if ((q >> 1) < G::NB)
                        a = (q & 1) ? dp2a_hi(p.cp_even[f][k], B[q >> 1], a) : dp2a_lo(p.cp_even[f][k], B[q >> 1], a);
                }
            } else {
#pragma unroll
                for (int k = 0; k < G::NPO; ++k) {
                    const int q = (s0 - 1) / 2 + k;
                    if ((q >> 1) < G::NB)
                        a = (q & 1) ? dp2a_hi(p.cp_odd[f][k], B[q >> 1], a) : dp2a_lo(p.cp_odd[f][k], B[q >> 1], a);
                }
            }
        } else {
#pragma unroll
            for (int d = -G::lo; d <= G::hi; ++d) {
                const int s = G::HL4 + i + d;
                const uint32_t v = __byte_perm(B[s >> 2], 0u, 0x4440u | (unsigned)(s & 3));
                a = (int)((uint32_t)a + (uint32_t)p.hq[G::hi - d] * v);      // modulo 2^32: exact under the acc_bits wrap
            }
        }
        acc[i] = a;
    }
}

// Two accumulators -> two saturated outputs in 16-bit lanes.
template <bool HI16>
__device__ __forceinline__ uint32_t post_pair(int a0, int a1, int shift)
{
    uint32_t v;
    if constexpr (HI16) {
        v = __byte_perm((uint32_t)a0, (uint32_t)a1, 0x7632u);
    } else {
        v = __byte_perm((uint32_t)(a0 >> shift), (uint32_t)(a1 >> shift), 0x5410u);
    }
    return __vimin_s16x2_relu(v, 0x00ff00ffu);
}

// Same for the kernel's POST mode.  WRAP: the int32 sum is exact modulo 2^32 (dp2a and IMAD wrap), so its low
// acc_bits bits are those of the reference's unbounded sum; (w + bias) >> frac is evaluated as
// (w >> frac) + (((w mod 2^frac) + bias) >> frac), which cannot overflow for any acc_bits <= 32, frac <= 31.
template <int POST, typename Params>
__device__ __forceinline__ uint32_t post_pair_mode(int a0, int a1, const Params &p)
{
    if constexpr (POST == POST_WRAP) {
        int v0, v1;
        if (!p.wrap_split) {
            // t = a << (32 - acc_bits) = w * 2^(32-acc_bits) exactly (w = the wrapped accumulator); one 32 x 32 -> 64
            // multiply-add then gives (w + 2^(frac-1)) >> frac in its upper word:
            //   t * 2^(acc_bits-frac) + 2^31 = (w + 2^(frac-1)) * 2^(32-frac)       (needs 2 <= frac + 32 - acc_bits <= 32)
            const long long r0 = (long long)(a0 << p.wrap_sh) * (long long)p.wrap_mul + 0x80000000LL;
            const long long r1 = (long long)(a1 << p.wrap_sh) * (long long)p.wrap_mul + 0x80000000LL;
            v0 = (int)(r0 >> 32);
            v1 = (int)(r1 >> 32);
        } else {
            // general form: sign-extend, then (w + bias) >> frac as (w >> frac) + (((w mod 2^frac) + bias) >> frac)
            const int w0 = (a0 << p.wrap_sh) >> p.wrap_sh, w1 = (a1 << p.wrap_sh) >> p.wrap_sh;
            v0 = (w0 >> p.shift) + (int)((((uint32_t)w0 & p.frac_mask) + p.post_bias) >> p.shift);
            v1 = (w1 >> p.shift) + (int)((((uint32_t)w1 & p.frac_mask) + p.post_bias) >> p.shift);
        }
        v0 = __vimin_s32_relu(v0, 255);
        v1 = __vimin_s32_relu(v1, 255);
        return __byte_perm((uint32_t)v0, (uint32_t)v1, 0x5410u);
    } else {
        return post_pair<POST == POST_HI16>(a0, a1, p.shift);
    }
}

// One thread's chunk, step 1: raw staged words (left halo | 16 bytes | right
// halo) -> byte-packed sample streams (one per I/Q component).
template <typename InT, int L, int DIL>
__device__ __forceinline__ void pack_chunk(uint32_t (&raw)[Geo<InT, L, DIL>::NR], uint32_t (&B)[DIL][Geo<InT, L, DIL>::NB])
{
    using G = Geo<InT, L, DIL>;
    constexpr int ES = G::ES;
    if constexpr (ES == 2) {
#pragma unroll
        for (int j = 0; j < G::NR; ++j) raw[j] = __vimin_s16x2_relu(raw[j], 0x00ff00ffu);
    }
    if constexpr (ES == 1 && DIL == 1) {
#pragma unroll
        for (int j = 0; j < G::NB; ++j) B[0][j] = raw[j];
    } else if constexpr (ES == 2 && DIL == 1) {
#pragma unroll
        for (int j = 0; j < G::NB; ++j) B[0][j] = __byte_perm(raw[2 * j], raw[2 * j + 1], 0x6420u);
    } else if constexpr (ES == 1 && DIL == 2) {
#pragma unroll
        for (int j = 0; j < G::NB; ++j) {
            B[0][j] = __byte_perm(raw[2 * j], raw[2 * j + 1], 0x6420u);
            B[1][j] = __byte_perm(raw[2 * j], raw[2 * j + 1], 0x7531u);
        }
    } else {
#pragma unroll
        for (int j = 0; j < G::NB; ++j) {
            const uint32_t ta = __byte_perm(raw[4 * j], raw[4 * j + 1], 0x6240u);
            const uint32_t tb = __byte_perm(raw[4 * j + 2], raw[4 * j + 3], 0x6240u);
            B[0][j] = __byte_perm(ta, tb, 0x5410u);
            B[1][j] = __byte_perm(ta, tb, 0x7632u);
        }
    }
}

// Step 2: byte-packed streams x tap set f -> output words in memory order.
// ENDS (flattened batches only): the chunk's window straddles a row boundary; `cut_s` samples of each stream's
// window belong to the row on the left.  Every row is zero-padded on its own, so the sums are formed twice — once
// with the right row's samples zeroed, once with the left row's — and each output takes the sum of its own row.
template <typename InT, typename OutT, int L, int DIL, int POST, int MAC, bool ENDS>
__device__ __forceinline__ void mac_post_chunk(const StagedParams &p, const int f, const uint32_t (&B)[DIL][Geo<InT, L, DIL>::NB],
                                               const int cut_s, uint32_t (&O)[Geo<InT, L, DIL>::SC * sizeof(OutT) / 4])
{
    using G = Geo<InT, L, DIL>;
    constexpr int OS = sizeof(OutT);
    constexpr int NOUT = G::SC * OS / 4;
    int acc[DIL][G::S];
    if constexpr (!ENDS) {
#pragma unroll
        for (int d = 0; d < DIL; ++d) mac_stream<G, L, MAC>(p, f, B[d], acc[d]);
    } else {
#pragma unroll
        for (int d = 0; d < DIL; ++d) {
            uint32_t BL[G::NB], BR[G::NB];
#pragma unroll
            for (int j = 0; j < G::NB; ++j) {
                // bits of this quad that lie in the RIGHT row: all ones shifted up by 8 x (bytes of the quad in the left
                // row, clamped to 0..4) — shl.b32 yields 0 for a shift of 32, which plain C++ does not promise
                const int mb = __vimin_s32_relu(8 * cut_s - 32 * j, 32);
                uint32_t mr;
                asm("shl.b32 %0, %1, %2;" : "=r"(mr) : "r"(0xffffffffu), "r"(mb));
                BL[j] = B[d][j] & ~mr;
                BR[j] = B[d][j] & mr;
            }
            int aL[G::S], aR[G::S];
            mac_stream<G, L, MAC>(p, f, BL, aL);
            mac_stream<G, L, MAC>(p, f, BR, aR);
#pragma unroll
            for (int i = 0; i < G::S; ++i) acc[d][i] = (G::HL4 + i < cut_s) ? aL[i] : aR[i];
        }
    }

    // accumulators -> saturated 16-bit lanes in memory order
    uint32_t V[G::SC / 2];
    if constexpr (DIL == 1) {
#pragma unroll
        for (int j = 0; j < G::S / 2; ++j) V[j] = post_pair_mode<POST>(acc[0][2 * j], acc[0][2 * j + 1], p);
    } else {
#pragma unroll
        for (int j = 0; j < G::S; ++j) V[j] = post_pair_mode<POST>(acc[0][j], acc[1][j], p);
    }
    if constexpr (OS == 2) {
#pragma unroll
        for (int j = 0; j < NOUT; ++j) O[j] = V[j];
    } else {
#pragma unroll
        for (int j = 0; j < NOUT; ++j) O[j] = __byte_perm(V[2 * j], V[2 * j + 1], 0x6420u);
    }
}

// Flattened batches: how many scalars of the window of a chunk whose first scalar sits at position r of its row
// belong to the row on the LEFT of a row boundary (0 = the window lies inside one row).  The host guarantees
// period >= window, so a window meets at most one boundary.
template <typename G, int DIL>
__device__ __forceinline__ int row_cut(uint32_t r, uint32_t period)
{
    constexpr uint32_t HLS = G::HL4 * DIL, HRS = G::HR4 * DIL;
    if (r < HLS) return (int)(HLS - r);                    // the row starts inside the left halo
    if (r + G::SC + HRS > period) return (int)(period - r + HLS);   // the row ends inside the chunk or its right halo
    return 0;
}

// Wide datapath (MAC == 2): int16 samples used at full width — the signed int16
// path of the extended entry points.  Each staged pair is clamped to the
// caller's sample range (packed, skipped for the full int16 range), unpacked
// to int32, accumulated with IMAD (taps are constant-bank operands), shifted,
// and packed back with saturation (cvt.pack.sat.s16.s32 = two outputs per
// instruction), then clamped to the output range.  dilation needs no
// de-interleave here: every scalar has its own register.
template <int L, int DIL, bool CLAMP>
__device__ __forceinline__ void wide_chunk(const StagedParams &p, const uint32_t (&raw)[Geo<int16_t, L, DIL>::NR], uint32_t (&O)[4])
{
    using G = Geo<int16_t, L, DIL>;
    constexpr int HLS = G::HLB / 2;                // halo scalars in front of the chunk
    int X[G::NR * 2];
#pragma unroll
    for (int j = 0; j < G::NR; ++j) {
        uint32_t w = raw[j];
        if constexpr (CLAMP) w = __vmaxs2(__vmins2(w, p.in_hi2), p.in_lo2);
        // sign-extend the low half: prmt selector 9 = byte 1 with sign replication (the
        // __byte_perm intrinsic masks that bit off, so the instruction is written out)
        asm("prmt.b32 %0, %1, %1, 0x9910;" : "=r"(X[2 * j]) : "r"(w));
        X[2 * j + 1] = (int)w >> 16;
    }
    int v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        int a = p.bias;
#pragma unroll
        for (int k = 0; k < L; ++k) a += p.hq[k] * X[HLS + e + (G::hi - k) * DIL];
        v[e] = a >> p.shift;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t o;
        asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(o) : "r"(v[2 * j + 1]), "r"(v[2 * j]));
        if constexpr (CLAMP) o = __vmaxs2(__vmins2(o, p.out_hi2), p.out_lo2);
        O[j] = o;
    }
}

// Write one chunk's output words: a vector store, or guarded scalars at ragged
// row ends / unaligned outputs (`room` = scalars left in the row).
template <typename OutT, int SC>
__device__ __forceinline__ void store_chunk(OutT *dst, const uint32_t (&O)[SC * sizeof(OutT) / 4], bool vec, long long room)
{
    constexpr int OS = sizeof(OutT);
    constexpr int NOUT = SC * OS / 4;
    if (vec) {
        if constexpr (NOUT == 2) {
            store_v2(dst, O[0], O[1]);
        } else if constexpr (NOUT == 8) {
            if ((reinterpret_cast<uintptr_t>(dst) & 31u) == 0) store_v8(dst, O);
            else {
                store_v4(dst, O[0], O[1], O[2], O[3]);
                store_v4(dst + 16 / OS, O[4], O[5], O[6], O[7]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < NOUT; j += 4) store_v4(dst + j * 4 / OS, O[j], O[j + 1], O[j + 2], O[j + 3]);
        }
    } else {
#pragma unroll
        for (int e = 0; e < SC; ++e) {
            if (e < room) {
                if constexpr (OS == 2) dst[e] = (OutT)((O[e / 2] >> (16 * (e & 1))) & 0xffffu);
                else dst[e] = (OutT)((O[e / 4] >> (8 * (e & 3))) & 0xffu);
            }
        }
    }
}

// RE ("row ends"): this instantiation takes flattened batches (p.period = row length) and zero-pads every row
// on its own; the others carry no code for it.
template <typename C, typename InT, typename OutT, int L, int DIL, int POST, int MAC, int NF, bool RE = false>
__global__ void __launch_bounds__(C::kThreads, C::kMinBlocks)
fir1d_staged_kernel(const __grid_constant__ StagedParams p)
{
    using G = Geo<InT, L, DIL>;
    constexpr int kThreads = C::kThreads, kSub = C::kSub, kStages = C::kStages, kStageBytes = C::stage_bytes(G::PADB);
    constexpr int TS = C::kTileBytes / G::ES;      // scalars per tile
    constexpr int ES = G::ES;
    constexpr int OS = sizeof(OutT);
    constexpr int NOUT = G::SC * OS / 4;          // output words per chunk

    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + kStages * kStageBytes;
    const int tid = threadIdx.x;

    long long n_my = 0;
    if ((long long)blockIdx.x < p.total_tiles)
        n_my = (p.total_tiles - 1 - blockIdx.x) / gridDim.x + 1;
    const unsigned tpr = (unsigned)p.tiles_per_row;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(bar_base + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Programmatic dependent launch: everything above overlapped the previous kernel's tail; its results (this
    // kernel's input, e.g. the uint8 copy of a float container) are visible from here on.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // Our dependents (the next kernel in the stream) may become resident from now on: every CTA of this grid is
    // resident at this point (one wave), so theirs spread evenly over the SMs; their prologue overlaps our run and
    // their own griddepcontrol.wait still waits for this whole grid.  (Triggering at a CTA's last tile instead was
    // measured 22 % slower back to back: SMs whose CTA had already exited took two CTAs of the next grid.)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // Producer cursor (thread 0 only): the tile whose bulk copy is issued next.
    long long p_row = 0;
    unsigned p_tcol = 0;
    int p_stage = 0;
    auto issue = [&]() {
        const TilePlan t = plan_tile<G, TS>(p, p_row, p_tcol);
        if (t.bulk_hi > t.bulk_lo) {
            const uint32_t bytes = (uint32_t)(t.bulk_hi - t.bulk_lo) * ES;
            const uint32_t dst = smem_base + p_stage * kStageBytes + (uint32_t)(t.bulk_lo - (t.col0 - G::PADS)) * ES;
            const unsigned char *src = p.x + (t.row * p.ld_x + t.bulk_lo) * ES;
            fence_proxy_async();
            mbar_expect_tx(bar_base + 8 * p_stage, bytes);
            bulk_g2s(dst, src, bytes, bar_base + 8 * p_stage);
        }
        advance_tile(p_row, p_tcol, gridDim.x, tpr);
        p_stage = p_stage + 1 == kStages ? 0 : p_stage + 1;
    };
    if (tid == 0) {
        advance_tile(p_row, p_tcol, blockIdx.x, tpr);
        for (int i = 0; i < kStages && i < n_my; ++i) issue();
    }

    long long c_row = 0;
    unsigned c_tcol = 0;
    advance_tile(c_row, c_tcol, blockIdx.x, tpr);
    int stage = 0;
    uint32_t phase_bits = 0;
    const uint32_t lane_off = G::PADB + tid * kChunk;
    // flattened batch: position of this thread's first chunk inside its row, kept up to date with adds only
    constexpr bool kRowEnds = RE;
    static_assert(!RE || (MAC != 2 && L <= kMaxWideTaps), "row ends: short filters on the packed datapaths only");
    const uint32_t period = kRowEnds ? (uint32_t)p.period : 0u;
    uint32_t tile_rem = 0, lane_rem = 0;
    if constexpr (kRowEnds) {
        if (period) {
            tile_rem = (uint32_t)(((unsigned long long)blockIdx.x * TS) % period);
            lane_rem = (uint32_t)(tid * G::SC) % period;
        }
    }
    // Interior tiles — the bulk copy covers the whole stage and the tile lies wholly inside the row — need no plan: the
    // two bounds below (first / last column a whole-stage bulk copy may touch) are loop invariants.  Everything else
    // (row ends, shard seams, unaligned rows) takes the general plan.  Used where the per-tile plan is a visible share
    // of the instruction stream (smaller tiles, heavier post stage): A/B on one box, fraction of the HBM peak with /
    // without — WRAP post 0.986 / 0.936, IMAD 0.885 / 0.817, row ends 0.77 / 0.70, I/Q 0.978 / 0.961; the plain long-row
    // kernels schedule worse with it (int16 -> int16 0.990 / 1.000, int16 -> uint8 0.938 / 0.995) and keep the plan.
    constexpr bool kInteriorFast = RE || POST == POST_WRAP || MAC == 1 || DIL == 2;
    constexpr int Q16 = 16 / G::ES;
    const long long bulk_first = floor_div((p.hist ? 0 : -p.halo_left) + Q16 - 1, Q16) * Q16;
    const long long bulk_last = floor_div(p.cols + p.halo_right, Q16) * Q16;
    for (long long i = 0; i < n_my; ++i) {
        const long long col0 = (long long)c_tcol * TS;
        const long long g0 = col0 - G::PADS;
        const uint32_t st_base = smem_base + stage * kStageBytes;
        const bool interior = kInteriorFast && p.x_bulk_ok && g0 >= bulk_first && col0 + TS + G::PADS <= bulk_last && col0 + TS <= p.cols;
        bool has_bulk = true;
        if (!interior) {
        const TilePlan t = plan_tile<G, TS>(p, c_row, c_tcol);

        // Patch what the bulk copy does not cover: zero padding outside the
        // readable range, and everything when the row is not 16-byte aligned.
        has_bulk = t.bulk_hi > t.bulk_lo;
        const long long left_end = has_bulk ? t.bulk_lo : t.need_hi;
        const long long right_begin = has_bulk ? t.bulk_hi : t.need_hi;
        if (left_end > g0 || t.need_hi > right_begin) {
            const long long vlo = -p.halo_left, vhi = p.cols + p.halo_right;
            const InT *xrow = reinterpret_cast<const InT *>(p.x) + c_row * p.ld_x;
            const InT *hist = reinterpret_cast<const InT *>(p.hist);
            InT *sts = reinterpret_cast<InT *>(smem + stage * kStageBytes);
            auto sample = [&](long long g) -> InT {
                if (g < vlo || g >= vhi) return (InT)0;
                return (g < 0 && hist) ? hist[g - vlo] : xrow[g];
            };
            for (long long g = g0 + tid; g < left_end; g += kThreads) sts[g - g0] = sample(g);
            for (long long g = right_begin + tid; g < t.need_hi; g += kThreads) sts[g - g0] = sample(g);
            __syncthreads();
        }
        }
        if (has_bulk) {
            mbar_wait(bar_base + 8 * stage, (phase_bits >> stage) & 1u);
            phase_bits ^= 1u << stage;
        }

        // Stage -> registers for all sub-steps, then hand the stage back.
        uint32_t raw[kSub][G::NR];
#pragma unroll
        for (int u = 0; u < kSub; ++u) {
            const uint32_t a = st_base + lane_off + u * (kThreads * kChunk);
            lds_before<G::HLB / 4>(a, &raw[u][0]);
            lds_words<4>(a, &raw[u][G::HLB / 4]);
            lds_after<G::HRB / 4>(a + kChunk, &raw[u][G::HLB / 4 + 4]);
        }
        __syncthreads();
        if (tid == 0 && i + kStages < n_my) issue();

        OutT *ythread = reinterpret_cast<OutT *>(p.y) + c_row * p.ld_y + col0 + tid * G::SC;
        uint32_t r_chunk = 0;                              // row position of chunk u = 0; advanced per chunk below
        if constexpr (kRowEnds) {
            r_chunk = tile_rem + lane_rem;
            if (r_chunk >= period) r_chunk -= period;
        }
        // one chunk: pack, then per tap set accumulate + post-process + store
        auto do_chunk = [&](const int u, const bool vec, const long long room) {
            if constexpr (MAC == 2) {
                uint32_t O[NOUT];
                wide_chunk<L, DIL, POST != 0>(p, raw[u], O);           // POST slot = "clamp ranges are not the full int16 range"
                store_chunk<OutT, G::SC>(ythread + u * (kThreads * G::SC), O, vec, room);
            } else {
                uint32_t B[DIL][G::NB];
                pack_chunk<InT, L, DIL>(raw[u], B);
                int cut = 0;
                bool ends = false;
                if constexpr (kRowEnds) {
                    if (period) {
                        cut = row_cut<G, DIL>(r_chunk, period);
                        ends = __any_sync(0xffffffffu, cut != 0);                 // warp-uniform: no divergent double work
                        r_chunk += p.rem_sub_step;
                        if (r_chunk >= period) r_chunk -= period;
                    }
                }
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    uint32_t O[NOUT];
                    if constexpr (kRowEnds) {
                        if (ends) mac_post_chunk<InT, OutT, L, DIL, POST, MAC, true>(p, f, B, cut / DIL, O);
                        else mac_post_chunk<InT, OutT, L, DIL, POST, MAC, false>(p, f, B, 0, O);
                    } else {
                        mac_post_chunk<InT, OutT, L, DIL, POST, MAC, false>(p, f, B, 0, O);
                    }
                    if (room > 0) store_chunk<OutT, G::SC>(ythread + f * p.filter_stride_y + u * (kThreads * G::SC), O, vec, room);
                }
            }
        };
        if (p.y_vec_ok && col0 + TS <= p.cols) {
            // whole tile inside the row: no per-chunk bounds logic at all
#pragma unroll
            for (int u = 0; u < kSub; ++u) do_chunk(u, true, G::SC);
        } else {
            const long long left = p.cols - col0 - tid * G::SC;      // scalars from this thread's first chunk to the row end
#pragma unroll
            for (int u = 0; u < kSub; ++u) {
                const long long room = left - u * (kThreads * G::SC);
                if constexpr (MAC == 2) {
                    if (room <= 0) continue;
                }
                do_chunk(u, p.y_vec_ok && room >= G::SC, room);         // (lanes past the row end compute and store nothing)
            }
        }
        advance_tile(c_row, c_tcol, gridDim.x, tpr);
        stage = stage + 1 == kStages ? 0 : stage + 1;
        if constexpr (kRowEnds) {
            tile_rem += p.rem_tile_step;
            if (tile_rem >= period) tile_rem -= period;
        }
    }
    // streaming: the next block's history = this block's tail (written to the OTHER history buffer: the CTA of the
    // first tile may still be reading the current one)
    if (p.hist_out && (long long)blockIdx.x == (p.total_tiles - 1) % gridDim.x) {
        const InT *xrow = reinterpret_cast<const InT *>(p.x);
        InT *out = reinterpret_cast<InT *>(p.hist_out);
        for (long long j = tid; j < p.halo_left; j += kThreads) out[j] = xrow[p.cols - p.halo_left + j];
    }
}

// Per-process launch knobs, read once (getenv walks the whole environment; a launch should not).
struct LaunchEnv {
    int ctas_per_sm = 0;       // FIRGPU_CTAS_PER_SM: cap on resident CTAs per SM (profiling runs)
    bool pdl = true;           // FIRGPU_NO_PDL=1 launches without programmatic stream serialisation (A/B runs)
    int pdl_max_tiles = 16;    // FIRGPU_PDL_MAX_TILES: largest launch (tiles per CTA) that carries the attribute (measured: 4 -> 16 tiles
                               // lifts back-to-back 2^24 / 2^25-sample passes from 0.91 / 0.86 to 0.96 / 0.89 of the HBM peak; 110 tiles lose 1 %)
    LaunchEnv()
    {
        if (const char *e = getenv("FIRGPU_CTAS_PER_SM")) ctas_per_sm = atoi(e);
        if (const char *e = getenv("FIRGPU_NO_PDL")) pdl = !(e[0] == '1');
        if (const char *e = getenv("FIRGPU_PDL_MAX_TILES")) pdl_max_tiles = atoi(e);
    }
};
inline const LaunchEnv &launch_env()
{
    static const LaunchEnv env;
    return env;
}
// SM count per device, cached (cudaDeviceGetAttribute is a driver round trip).
inline int sm_count(int dev)
{
    static int sms[64] = {0};
    if (dev >= 0 && dev < 64 && sms[dev]) return sms[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 1;
    if (dev >= 0 && dev < 64) sms[dev] = n;
    return n;
}
// Launch with (optionally) the programmatic-stream-serialisation attribute: the kernel may become resident while
// its predecessor in the stream drains; it executes griddepcontrol.wait before it touches global memory.
template <typename Kern, typename... Args>
inline cudaError_t launch_pdl(Kern kern, unsigned grid, unsigned block, size_t smem, cudaStream_t stream, bool pdl, const Args &...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

template <typename C, typename InT, typename OutT, int L, int DIL, int POST, int MAC, int NF = 1, bool RE = false>
int launch_cfg(const StagedParams &p_in, cudaStream_t stream)
{
    auto kern = fir1d_staged_kernel<C, InT, OutT, L, DIL, POST, MAC, NF, RE>;
    using G = Geo<InT, L, DIL>;
    constexpr int kSmem = C::smem_bytes(G::PADB);
    // a flattened batch needs every chunk window to meet at most one row boundary
    if (p_in.period && (!RE || p_in.period < (long long)(G::HL4 * DIL + G::SC + G::HR4 * DIL) || p_in.period >= (1LL << 31) || p_in.rows != 1))
        return 0;
    static std::mutex mu;
    static int ctas_per_sm[64] = {0};
    int dev = 0;
    FIRGPU_CUDA(cudaGetDevice(&dev));
    int per_sm;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev < 64 && ctas_per_sm[dev] == 0) {
            FIRGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
            int n = 0;
            FIRGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, C::kThreads, kSmem));
            if (C::kCtasPerSm > 0 && n > C::kCtasPerSm) n = C::kCtasPerSm;
            ctas_per_sm[dev] = n < 1 ? 1 : n;
        }
        per_sm = dev < 64 ? ctas_per_sm[dev] : 1;
    }
    const LaunchEnv &env = launch_env();
    if (env.ctas_per_sm >= 1 && env.ctas_per_sm < per_sm) per_sm = env.ctas_per_sm;
    StagedParams p = p_in;
    const long long ts = C::kTileBytes / (int)sizeof(InT);
    p.tiles_per_row = (p.cols + ts - 1) / ts;
    if (p.tiles_per_row >= (1LL << 31)) return 0;
    p.total_tiles = p.tiles_per_row * p.rows;
    long long grid = (long long)sm_count(dev) * per_sm;
    if (grid > p.total_tiles) grid = p.total_tiles;
    if (p.period) {
        p.rem_tile_step = (uint32_t)((unsigned long long)(grid * ts) % (unsigned long long)p.period);
        p.rem_sub_step = (uint32_t)((unsigned long long)(C::kThreads * G::SC) % (unsigned long long)p.period);
    }
    // Programmatic stream serialisation only for launches of at most pdl_max_tiles tiles per CTA (about 30 microseconds), where
    // the launch latency it hides is a visible share; long kernels measured 1 % slower with it.
    const bool pdl = env.pdl && p.pdl && p.total_tiles <= (long long)env.pdl_max_tiles * grid;
    FIRGPU_CUDA(launch_pdl(kern, (unsigned)grid, C::kThreads, kSmem, stream, pdl, p));
    note_launch(FIRGPU_PATH_STAGED);
    return 1;
}

template <typename InT, typename OutT, int L, int DIL, int POST, int MAC>
int launch_one(const StagedParams &p, cudaStream_t stream)
{
    // four tap sets over one staged input (the generators' coefficient loop):
    // uint8 images, 3 / 5 taps only
    if constexpr (sizeof(InT) == 1 && sizeof(OutT) == 1 && (L == 3 || L == 5) && DIL == 1 && MAC == 0) {
        if (p.n_filters == kMaxFilters)
            return p.period ? launch_cfg<Cfg<256, 2, 2, 2, 2>, InT, OutT, L, DIL, POST, MAC, kMaxFilters, true>(p, stream)
                            : launch_cfg<Cfg<256, 2, 2, 2, 2>, InT, OutT, L, DIL, POST, MAC, kMaxFilters>(p, stream);
    }
    if (p.n_filters != 1) return 0;
    // flattened batches (dense rows filtered as one long row, row ends inside the kernel): one configuration,
    // 8 KiB tiles and three CTAs per SM — the row-end path is latency-bound and wants the warps
    if constexpr (MAC != 2 && L <= kMaxWideTaps) {
        if (p.period) return launch_cfg<FlatCfg, InT, OutT, L, DIL, POST, MAC, 1, true>(p, stream);
    }
    if (p.period) return 0;
    using Big = typename TuneFor<InT, OutT>::type;
    using Long = typename std::conditional<POST == POST_WRAP, WrapCfg, Big>::type;
#ifndef FIRGPU_K_THREADS
    if (p.cols * (long long)sizeof(InT) < 2LL * Big::kTileBytes)
        return launch_cfg<SmallRowCfg, InT, OutT, L, DIL, POST, MAC>(p, stream);
#endif
    return launch_cfg<Long, InT, OutT, L, DIL, POST, MAC>(p, stream);
}

// post: POST_SHIFT / POST_HI16 / POST_WRAP (wide TU: 0 / 1 = clamp flag).  mac: 0 = dp2a (signed 16-bit taps),
// 1 = IMAD with full int32 taps — built for the WRAP post mode only, which is exact for any taps.
template <typename InT, typename OutT, int L, int DIL>
int launch_modes(const StagedParams &p, int post, int mac, cudaStream_t s)
{
#ifdef FIRGPU_STAGED_WIDE_TU             // fir1d_staged_wide.cu: the wide (signed int16) datapath only
    static_assert(sizeof(InT) == 2 && sizeof(OutT) == 2, "wide datapath is int16 -> int16");
    (void)mac;
    return post ? launch_one<InT, OutT, L, DIL, 1, 2>(p, s) : launch_one<InT, OutT, L, DIL, 0, 2>(p, s);
#else
    if (mac == 1) return post == POST_WRAP ? launch_one<InT, OutT, L, DIL, POST_WRAP, 1>(p, s) : 0;
    switch (post) {
    case POST_SHIFT: return launch_one<InT, OutT, L, DIL, POST_SHIFT, 0>(p, s);
    case POST_HI16: return launch_one<InT, OutT, L, DIL, POST_HI16, 0>(p, s);
    case POST_WRAP: return launch_one<InT, OutT, L, DIL, POST_WRAP, 0>(p, s);
    default: return 0;
    }
#endif
}

template <typename InT, typename OutT, int DIL>
int launch_taps(const StagedParams &p, int n_taps, int post, int mac, cudaStream_t s)
{
    switch (n_taps) {
#ifndef FIRGPU_QUICK                     // (make QUICK=1: experiment builds carry 3 and 5 taps only)
    case 1: return launch_modes<InT, OutT, 1, DIL>(p, post, mac, s);
    case 2: return launch_modes<InT, OutT, 2, DIL>(p, post, mac, s);
    case 4: return launch_modes<InT, OutT, 4, DIL>(p, post, mac, s);
    case 6: return launch_modes<InT, OutT, 6, DIL>(p, post, mac, s);
    case 7: return launch_modes<InT, OutT, 7, DIL>(p, post, mac, s);
    case 8: return launch_modes<InT, OutT, 8, DIL>(p, post, mac, s);
    case 9: return launch_modes<InT, OutT, 9, DIL>(p, post, mac, s);
#endif
    case 3: return launch_modes<InT, OutT, 3, DIL>(p, post, mac, s);
    case 5: return launch_modes<InT, OutT, 5, DIL>(p, post, mac, s);
    default: return 0;
    }
}

// Long filters (10..129 taps): the same kernel, instantiated for a few tap-count
// buckets; the host zero-pads the taps to the bucket (keeping the centre) —
// zero taps cost dp2a slots but no accuracy.  8 KiB tiles, two CTAs per SM.
using LongTapCfg = Cfg<256, 2, 2, 1, 2>;
constexpr int kLongBuckets[] = {13, 17, 25, 33, 49, 65, 97, 129};
constexpr int kMaxLongTapsU8 = 65;              // uint8 samples: buckets up to here (see launch_long)
template <typename T, typename OutT, int LB>
int launch_long_bucket(const StagedParams &p, int post, cudaStream_t s)
{
    switch (post) {
    case POST_SHIFT: return launch_cfg<LongTapCfg, T, OutT, LB, 1, POST_SHIFT, 0>(p, s);
    case POST_HI16: return launch_cfg<LongTapCfg, T, OutT, LB, 1, POST_HI16, 0>(p, s);
    case POST_WRAP: return launch_cfg<LongTapCfg, T, OutT, LB, 1, POST_WRAP, 0>(p, s);
    default: return 0;
    }
}
template <typename T, typename OutT = T>
int launch_long(const StagedParams &p, int bucket, int post, cudaStream_t s)
{
    switch (bucket) {
    case 13: return launch_long_bucket<T, OutT, 13>(p, post, s);
#ifndef FIRGPU_QUICK
    case 17: return launch_long_bucket<T, OutT, 17>(p, post, s);
    case 25: return launch_long_bucket<T, OutT, 25>(p, post, s);
    case 33: return launch_long_bucket<T, OutT, 33>(p, post, s);
    case 49: return launch_long_bucket<T, OutT, 49>(p, post, s);
    case 65: return launch_long_bucket<T, OutT, 65>(p, post, s);
    // 97 / 129 taps: int16 samples only — with uint8 samples (16 per chunk) these kernels unroll to > 1000 dp2a per
    // sub-step and ptxas needs over 40 minutes per translation unit; such calls run on the generic kernel
    case 97:
        if constexpr (sizeof(T) == 2) return launch_long_bucket<T, OutT, 97>(p, post, s);
        else return 0;
    case 129:
        if constexpr (sizeof(T) == 2) return launch_long_bucket<T, OutT, 129>(p, post, s);
        else return 0;
#endif
    default: return 0;
    }
}
int launch_staged_long_u8(const StagedParams &p, int bucket, int post, cudaStream_t s);
int launch_staged_long_i16(const StagedParams &p, int bucket, int post, cudaStream_t s);
int launch_staged_long_i16u8(const StagedParams &p, int bucket, int post, cudaStream_t s);   // int16 samples in, uint8 out
int launch_staged_long_u8i16(const StagedParams &p, int bucket, int post, cudaStream_t s);   // uint8 samples in, int16 container out

// One translation unit per (input, output, dilation) combination instantiates these
// (fir1d_staged_inst.cu compiled eight times), which keeps the build parallel.
#define FIRGPU_DECLARE_STAGED(name) int name(const StagedParams &p, int n_taps, int post, int mac, cudaStream_t s)
FIRGPU_DECLARE_STAGED(launch_staged_u8_u8_d1);
FIRGPU_DECLARE_STAGED(launch_staged_u8_u8_d2);
FIRGPU_DECLARE_STAGED(launch_staged_u8_i16_d1);
FIRGPU_DECLARE_STAGED(launch_staged_u8_i16_d2);
FIRGPU_DECLARE_STAGED(launch_staged_i16_u8_d1);
FIRGPU_DECLARE_STAGED(launch_staged_i16_u8_d2);
FIRGPU_DECLARE_STAGED(launch_staged_i16_i16_d1);
FIRGPU_DECLARE_STAGED(launch_staged_i16_i16_d2);
int launch_staged_wide(const StagedParams &p, int n_taps, int dilation, bool clamp, cudaStream_t s);

}  // namespace staged
}  // namespace firgpu
