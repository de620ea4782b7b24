// Host side of the tiled 2-D fixed-point FIR (kernels: fir2d_tiled.cuh):
// eligibility, exactness proof, rank-1 factorisation of the integer kernel,
// CUtensorMap construction, work decomposition.
#include "fir2d_tiled.cuh"
#include "fastplan.h"

#include <cstdlib>
#include <cstring>

namespace firgpu {
namespace {
using namespace tiled;

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static const EncodeTiledFn fn = [] {            // resolved once (thread-safe static initialisation)
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return (EncodeTiledFn)ptr;
        return (EncodeTiledFn) nullptr;
    }();
    return fn;
}

// Encoded tensor maps, per host thread: a map depends only on (base, dims, strides, box rows), and callers filter the
// same buffers again and again (a row-band shard per step, the same image with several kernels); encoding costs a
// driver call per launch otherwise.  Eight entries, round robin.
struct MapKey {
    const void *base;
    cuuint64_t dims[3], strides[2];
    cuuint32_t box_rows;
    bool operator==(const MapKey &o) const
    {
        return base == o.base && dims[0] == o.dims[0] && dims[1] == o.dims[1] && dims[2] == o.dims[2] && strides[0] == o.strides[0] &&
               strides[1] == o.strides[1] && box_rows == o.box_rows;
    }
};
struct MapCache {
    static constexpr int kEntries = 8;
    MapKey key[kEntries] = {};
    CUtensorMap map[kEntries];
    bool used[kEntries] = {};
    int next = 0;
};
thread_local MapCache tl_maps;

long long gcd_ll(long long a, long long b)
{
    a = a < 0 ? -a : a;
    b = b < 0 ? -b : b;
    while (b) {
        const long long t = a % b;
        a = b;
        b = t;
    }
    return a;
}

// hq[kh][kw] == a[m] * b[n] exactly?  (integer rank-1 factorisation)
bool factor_rank1(const int32_t *hq, int kh, int kw, long long *a, long long *b)
{
    int r = -1;
    for (int m = 0; m < kh && r < 0; ++m)
        for (int n = 0; n < kw; ++n)
            if (hq[m * kw + n] != 0) {
                r = m;
                break;
            }
    if (r < 0) return false;                       // all-zero kernel: nothing to gain
    long long g = 0;
    int c = -1;
    for (int n = 0; n < kw; ++n) {
        g = gcd_ll(g, hq[r * kw + n]);
        if (c < 0 && hq[r * kw + n] != 0) c = n;
    }
    for (int n = 0; n < kw; ++n) b[n] = hq[r * kw + n] / g;
    for (int m = 0; m < kh; ++m) {
        if ((long long)hq[m * kw + c] % b[c] != 0) return false;
        a[m] = (long long)hq[m * kw + c] / b[c];
        for (int n = 0; n < kw; ++n)
            if (a[m] * b[n] != (long long)hq[m * kw + n]) return false;
    }
    return true;
}

}  // namespace

bool fir2d_tiled_window_ok(int kh, int kw)
{
    return kh >= 1 && kw >= 1 && kh <= kMaxWindow && kw <= kMaxWindow && shape_supported(kh | 1, kw | 1);
}

int try_launch_fir2d_fixed_tiled(const Fir2dArgs &a_in, const int32_t *hq_in)
{
    // Even window sizes: a zero row / column appended at the END keeps the centre (L // 2 == (L + 1) // 2 for even L,
    // fir_1d_fixed_ref.py:98-107 convention in both directions) and adds nothing to any sum or bound.
    if (a_in.kh < 1 || a_in.kw < 1 || a_in.kh > kMaxKH || a_in.kw > kMaxKH) return 0;
    Fir2dArgs a = a_in;
    int32_t padded[kMaxKH * kMaxKH];
    const int32_t *hq = hq_in;
    if (!(a.kh & 1) || !(a.kw & 1)) {
        a.kh = a_in.kh | 1;
        a.kw = a_in.kw | 1;
        if (a.kh > kMaxKH || a.kw > kMaxKH) return 0;
        for (int m = 0; m < a.kh; ++m)
            for (int n = 0; n < a.kw; ++n) padded[m * a.kw + n] = (m < a_in.kh && n < a_in.kw) ? hq_in[m * a_in.kw + n] : 0;
        hq = padded;
    }
    const bool narrowing = a.in_dtype == FIRGPU_I16 && a.out_dtype == FIRGPU_U8;       // int16 pixels in, uint8 out
    if ((a.in_dtype != a.out_dtype && !narrowing) || (a.in_dtype != FIRGPU_U8 && a.in_dtype != FIRGPU_I16)) return 0;
    if (!shape_supported(a.kh, a.kw)) return 0;
    if (force_generic()) return 0;
    const int es = dtype_size(a.in_dtype);
    // TMA needs a 16-byte aligned base and 16-byte multiples for every pitch.
    if ((uintptr_t)a.x % 16 != 0 || (a.ld_x * es) % 16 != 0) return 0;
    if (a.n_img > 1 && (a.img_stride_x * es) % 16 != 0) return 0;
    // uint8 pixels travel as 16-bit map elements: whole pairs only, unless the rows carry a zero pad byte
    const int64_t map_w = a.width + ((es == 1 && (a.width & 1) && a.x_row_padded) ? 1 : 0);
    if (es == 1 && (map_w & 1)) return 0;
    if (a.width >= (1LL << 31) - 4096 || a.height + a.halo_top + a.halo_bottom >= (1LL << 31) - 64 || a.n_img >= (1LL << 31))
        return 0;
    EncodeTiledFn encode = encode_fn();
    if (!encode) return 0;

    const int n_taps = a.kh * a.kw;
    TiledParams p{};
    bool sep = false;
    FastPlan f;
    long long fa[kMaxKH], fb[8];
    // Post stage as in the 1-D kernel (fir1d_staged.cu, select_plan): shift-by-16 with pre-scaled taps, plain shift, or —
    // live acc_bits wrap, a shifted result beyond 16 bits — the WRAP form (exact modulo 2^32 for acc_bits <= 32).
    auto plan = [&](int tap_limit) {
        FastPlan g = plan_fast(hq, n_taps, a.frac_bits, a.acc_bits, tap_limit);
        if (!g.ok) g = plan_wrap(hq, n_taps, a.frac_bits, a.acc_bits, tap_limit);
        return g;
    };
    if (shape_has_separable(a.kh, a.kw) && !(getenv("FIRGPU_NO_SEPARABLE")) && factor_rank1(hq, a.kh, a.kw, fa, fb)) {
        bool b_fits = true;
        for (int n = 0; n < a.kw; ++n)
            if (fb[n] > 32767 || fb[n] < -32768) b_fits = false;
        for (int m = 0; m < a.kh; ++m)
            if (fa[m] > 0x7fffffffLL || fa[m] < -0x80000000LL) b_fits = false;
        if (b_fits) {
            f = plan(0x7fffffff);           // row pass: dp2a with vector b (int16); column pass: IMAD with vector a (int32)
            sep = f.ok;
        }
    }
    if (!sep) {
        f = plan(32767);
        if (!f.ok) return 0;
    }
    const int post = f.wrap ? POST_WRAP : (f.hi16 ? POST_HI16 : POST_SHIFT);
    if (sep) {
        const long long scale = f.hi16 ? (1LL << (16 - a.frac_bits)) : 1;
        int32_t b32[8];
        for (int n = 0; n < a.kw; ++n) b32[n] = (int32_t)fb[n];
        build_pairs(b32, a.kw, kMaxPairs2, p.rp_even, p.rp_odd);
        for (int m = 0; m < a.kh; ++m) p.col_taps[m] = (int32_t)(fa[m] * scale);
        p.col_sym = getenv("FIRGPU_NO_COL_SYM") ? 0 : 1;     // symmetric column vector: mirrored rows are added before the multiply
        for (int m = 0; m < a.kh; ++m)
            if (fa[m] != fa[a.kh - 1 - m]) p.col_sym = 0;
    } else {
        for (int m = 0; m < a.kh; ++m) build_pairs(f.taps + m * a.kw, a.kw, kMaxPairs2, p.cp_even[m], p.cp_odd[m]);
        if (a.kh == 5 && a.kw == 5) {
            // Kernel rows 0 .. fr-1 (the newest input rows) as exact FFMA sums on the other FMA pipe (fir2d_tiled.cuh,
            // fir2d_tiled_kernel): admitted when every partial sum of their products stays inside a 2^23-wide window.
            // The float taps are the UNSCALED hq (the dp2a taps may carry the shift-by-16 pre-scale, applied to the
            // float sum as a left shift when it joins the int32 accumulator).
            int fr = 0;                                      // measured slower than dp2a alone (profiles/README.md R2.5c)
#if FIRGPU_T2_FFMA
            if (const char *e = getenv("FIRGPU_T2_FFMA_ROWS")) fr = atoi(e);
#endif
            if (fr < 0) fr = 0;
            if (fr > 2) fr = 2;
            const int fsh = f.hi16 ? 16 - a.frac_bits : 0;
            for (; fr > 0; --fr) {
                long long sum_abs = 0;
                for (int k = 0; k < fr * a.kw; ++k) sum_abs += hq[k] < 0 ? -(long long)hq[k] : (long long)hq[k];
                if (255LL * sum_abs < (1LL << 23)) break;
            }
            p.ffma_rows = fr;
            if (fr > 0) {
                long long neg = 0;
                for (int m = 0; m < fr; ++m)
                    for (int t = 0; t < 5; ++t) {
                        const int32_t h = tap_at(hq + m * a.kw, 5, t - 2);      // multiplies the sample at offset t - 2
                        p.ftap[m][t] = (float)h;                                // |h| < 2^23 / 255: exact
                        if (h < 0) neg += -(long long)h;
                    }
                const uint32_t base_bits = 0x4b000000u + (uint32_t)(255LL * neg);   // bits of 2^23 + 255 * neg (< 2^24)
                memcpy(&p.fbase, &base_bits, sizeof(float));
                p.fsh = fsh;
                p.acc0 = (uint32_t)f.bias - (base_bits << fsh);
            }
            // the tap of a 5-wide kernel row that has no partner in its row: offset +2 (taps[0]) for windows that start on
            // an even sample, offset -2 (taps[4]) for odd ones; the dp2a rows fr, fr+1 (and fr+2, fr+3) share a dp2a
            // (fir2d_tiled.cuh, mac_cross)
            for (int j = 0; j < (5 - fr) / 2; ++j) {
                const int32_t *r0 = f.taps + (fr + 2 * j) * a.kw, *r1 = f.taps + (fr + 2 * j + 1) * a.kw;
                p.xp_even[j] = pack_pair(tap_at(r0, 5, 2), tap_at(r1, 5, 2));
                p.xp_odd[j] = pack_pair(tap_at(r0, 5, -2), tap_at(r1, 5, -2));
            }
        }
    }
    p.wrap_sh = f.wrap_sh;
    p.wrap_split = f.wrap_split ? 1 : 0;
    p.wrap_mul = f.wrap_mul;
    p.post_bias = f.post_bias;
    p.frac_mask = f.frac_mask;

    const int bh = rows_per_stage(a.kh);                     // Geo2::BH
    const int stages_per_item = (kBandTarget + a.kh - 1 + bh - 1) / bh;
    p.band_rows = stages_per_item * bh - (a.kh - 1);
    p.stages_per_item = stages_per_item;
    const int outw = kStripBytes / es - 2 * (16 / es);       // Geo2::OUTW
    p.y = (unsigned char *)a.y;
    p.ld_y = a.ld_y;
    p.img_stride_y = a.img_stride_y;
    p.width = (int)a.width;
    p.height = (int)a.height;
    p.n_img = (int)a.n_img;
    p.n_strips = (int)((a.width + outw - 1) / outw);
    p.n_bands = (int)((a.height + p.band_rows - 1) / p.band_rows);
    const long long items = (long long)p.n_img * p.n_bands * p.n_strips;
    if (items >= (1LL << 31) - (1 << 20)) return 0;
    p.total_items = (int)items;
    p.halo_top = (int)a.halo_top;
    p.shift = f.shift;
    p.bias = f.bias;
    const int os = dtype_size(a.out_dtype);
    const int va = 16 * os / es;                             // bytes per thread store: 16, or 8 when narrowing
    p.y_vec_ok = ((uintptr_t)a.y % va == 0) && (a.ld_y * os) % va == 0 && (a.n_img == 1 || (a.img_stride_y * os) % va == 0);

    const long long map_rows = a.halo_top + a.height + a.halo_bottom;
    cuuint64_t dims[3] = {(cuuint64_t)(map_w * es / kMapElemBytes), (cuuint64_t)map_rows, (cuuint64_t)a.n_img};
    cuuint64_t strides[2] = {(cuuint64_t)(a.ld_x * es),
                             (cuuint64_t)((a.n_img > 1 ? a.img_stride_x : map_rows * a.ld_x) * es)};
    cuuint32_t box[3] = {(cuuint32_t)kBoxElems, (cuuint32_t)bh, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    void *base = (void *)((const unsigned char *)a.x - a.halo_top * a.ld_x * es);
    const MapKey key{base, {dims[0], dims[1], dims[2]}, {strides[0], strides[1]}, box[1]};
    MapCache &cache = tl_maps;
    int slot = -1;
    for (int i = 0; i < MapCache::kEntries; ++i)
        if (cache.used[i] && cache.key[i] == key) slot = i;
    if (slot < 0) {
        slot = cache.next;
        const CUresult r = encode(&cache.map[slot], CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, base, dims,
                                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {                              // shape the TMA cannot describe: generic kernel
            cache.used[slot] = false;
            return 0;
        }
        cache.key[slot] = key;
        cache.used[slot] = true;
        cache.next = (slot + 1) % MapCache::kEntries;
    }
    const CUtensorMap &map = cache.map[slot];

    if (narrowing) return launch_tiled_i16u8(map, p, a.kh, a.kw, post, sep, a.stream);
    return es == 1 ? launch_tiled_u8(map, p, a.kh, a.kw, post, sep, a.stream)
                   : launch_tiled_i16(map, p, a.kh, a.kw, post, sep, a.stream);
}

}  // namespace firgpu
