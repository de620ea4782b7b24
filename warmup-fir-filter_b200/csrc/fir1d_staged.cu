// Host side of the staged 1-D fixed-point FIR (kernels: fir1d_staged.cuh):
// proves from the actual taps that the packed int32 datapath is exact, builds
// the dp2a tap pairs, and dispatches to the instantiation for the containers.
#include "fir1d_staged.cuh"
#include "fastplan.h"

#include <algorithm>
#include <cstdlib>

namespace firgpu {
namespace {
using namespace staged;

int g_mac_mode = -1;   // -1 = read FIRGPU_MAC once; 0 dp2a, 1 imad
int mac_mode()
{
    if (g_mac_mode < 0) {
        const char *e = getenv("FIRGPU_MAC");
        g_mac_mode = (e && e[0] == '1') ? 1 : 0;
    }
    return g_mac_mode;
}

// Every tap set must be provably exact in int32, and all of them share one post mode: the 16-bit-shift form when every
// set can be pre-scaled, else the plain shift, else — wrap live, taps beyond int16, a shifted result beyond 16 bits — the
// WRAP form (dp2a while the taps fit int16, IMAD with full int32 taps otherwise; IMAD exists for single tap sets of at most
// nine taps).
bool select_plan(const int32_t *taps_in, int n_taps, int n_filters, int frac_bits, int acc_bits, bool long_bucket, bool force_imad,
                 FastPlan *plans, int *post_out, int *mac_out)
{
    int post = POST_HI16, mac = 0;
    auto plan_all = [&](auto &&make) {
        for (int f = 0; f < n_filters; ++f) {
            plans[f] = make(taps_in + (size_t)f * n_taps);
            if (!plans[f].ok) return false;
        }
        return true;
    };
    bool ok = !force_imad && plan_all([&](const int32_t *t) { return plan_fast(t, n_taps, frac_bits, acc_bits, 32767); });
    if (ok) {
        for (int f = 0; f < n_filters; ++f)
            if (!plans[f].hi16) post = POST_SHIFT;
        if (post == POST_SHIFT)
            ok = plan_all([&](const int32_t *t) { return plan_fast(t, n_taps, frac_bits, acc_bits, 32767, false); });
    }
    if (!ok) {
        post = POST_WRAP;
        ok = !force_imad && plan_all([&](const int32_t *t) { return plan_wrap(t, n_taps, frac_bits, acc_bits, 32767); });
        if (!ok && !long_bucket && n_filters == 1) {
            mac = 1;
            ok = plan_all([&](const int32_t *t) { return plan_wrap(t, n_taps, frac_bits, acc_bits, 0x7fffffff); });
        }
    }
    *post_out = post;
    *mac_out = mac;
    return ok;
}

}  // namespace

int try_launch_fir1d_fixed_staged_multi(const Fir1dArgs &a, const int32_t *hq_host, int n_filters, int64_t filter_stride_y)
{
    if (n_filters != 1 && n_filters != kMaxFilters) return 0;
    if (a.dilation != 1 && a.dilation != 2) return 0;
    if (a.in_dtype != FIRGPU_U8 && a.in_dtype != FIRGPU_I16) return 0;
    if (a.out_dtype != FIRGPU_U8 && a.out_dtype != FIRGPU_I16) return 0;
    if (a.dilation == 2 && (a.cols & 1)) return 0;           // I/Q pairs must be whole
    if (a.dilation == 2 && ((a.halo_left | a.halo_right) & 1)) return 0;
    if (force_generic()) return 0;
    const bool force_imad = mac_mode() == 1 && n_filters == 1 && a.n_taps <= kMaxWideTaps;   // FIRGPU_MAC=1: A/B runs of the IMAD datapath
    if (a.n_taps > kMaxFastTaps) return 0;
    // 10..129 taps: zero-pad to the next bucket, keeping the centre tap where it is
    // (output n reads x[n - (k - L/2)], so the padded filter must keep h[L/2] at LB/2).
    int n_taps = a.n_taps, bucket = 0;
    int32_t padded[kMaxFastTaps];
    const int32_t *taps_in = hq_host;
    if (a.n_taps > kMaxWideTaps) {
        if (n_filters != 1 || a.dilation != 1) return 0;
        if (a.in_dtype == FIRGPU_U8 && a.n_taps > kMaxLongTapsU8) return 0;
        for (int b : kLongBuckets)
            if (!bucket && b >= a.n_taps) bucket = b;
        const int off = bucket / 2 - a.n_taps / 2;
        for (int k = 0; k < bucket; ++k) padded[k] = (k >= off && k < off + a.n_taps) ? hq_host[k - off] : 0;
        taps_in = padded;
        n_taps = bucket;
    }
    FastPlan plans[kMaxFilters];
    int post = POST_HI16, mac = 0;
    const bool ok = select_plan(taps_in, n_taps, n_filters, a.frac_bits, a.acc_bits, bucket != 0, force_imad, plans, &post, &mac);
    if (!ok) return 0;
    const FastPlan &f0 = plans[0];

    const int es = dtype_size(a.in_dtype), os = dtype_size(a.out_dtype);
    StagedParams p{};
    p.x = (const unsigned char *)a.x;
    p.y = (unsigned char *)a.y;
    p.rows = a.rows;
    p.cols = a.cols;
    p.ld_x = a.ld_x;
    p.ld_y = a.ld_y;
    p.halo_left = a.halo_left;
    p.halo_right = a.halo_right;
    p.n_filters = n_filters;
    p.filter_stride_y = filter_stride_y;
    p.shift = f0.shift;
    p.bias = f0.bias;
    p.wrap_sh = f0.wrap_sh;
    p.post_bias = f0.post_bias;
    p.frac_mask = f0.frac_mask;
    p.wrap_split = f0.wrap_split ? 1 : 0;
    p.wrap_mul = f0.wrap_mul;
    p.period = a.period;
    p.pdl = 1;
    if (a.hist) {                              // streaming block: single row, single tap set, history of halo_left scalars
        if (a.rows != 1 || n_filters != 1 || a.period || a.halo_left < 1 || a.cols < a.halo_left) return 0;
        p.hist = (const unsigned char *)a.hist;
        p.hist_out = (unsigned char *)a.hist_out;
    }
    p.x_bulk_ok = ((uintptr_t)a.x % 16 == 0) && (a.rows == 1 || (a.ld_x * es) % 16 == 0);
    const int out_vec = (int)((kChunk / es) * os);           // bytes per vector store: 8, 16 or 32
    const int out_align = out_vec > 16 ? 16 : out_vec;
    p.y_vec_ok = ((uintptr_t)a.y % out_align == 0) && (a.rows == 1 || (a.ld_y * os) % out_align == 0) &&
                 (n_filters == 1 || (filter_stride_y * os) % out_align == 0);
    for (int f = 0; f < n_filters; ++f) build_pairs(plans[f].taps, n_taps, kMaxPairs, p.cp_even[f], p.cp_odd[f]);
    for (int k = 0; k < n_taps && k < kMaxWideTaps; ++k) p.hq[k] = f0.taps[k];
    if (bucket) {
        if (a.in_dtype == FIRGPU_U8)
            return a.out_dtype == FIRGPU_U8 ? launch_staged_long_u8(p, bucket, post, a.stream)
                                            : launch_staged_long_u8i16(p, bucket, post, a.stream);
        return a.out_dtype == FIRGPU_U8 ? launch_staged_long_i16u8(p, bucket, post, a.stream)
                                        : launch_staged_long_i16(p, bucket, post, a.stream);
    }
    const bool d2 = a.dilation == 2;
    if (a.in_dtype == FIRGPU_U8) {
        if (a.out_dtype == FIRGPU_U8)
            return d2 ? launch_staged_u8_u8_d2(p, a.n_taps, post, mac, a.stream) : launch_staged_u8_u8_d1(p, a.n_taps, post, mac, a.stream);
        return d2 ? launch_staged_u8_i16_d2(p, a.n_taps, post, mac, a.stream) : launch_staged_u8_i16_d1(p, a.n_taps, post, mac, a.stream);
    }
    if (a.out_dtype == FIRGPU_U8)
        return d2 ? launch_staged_i16_u8_d2(p, a.n_taps, post, mac, a.stream) : launch_staged_i16_u8_d1(p, a.n_taps, post, mac, a.stream);
    return d2 ? launch_staged_i16_i16_d2(p, a.n_taps, post, mac, a.stream) : launch_staged_i16_i16_d1(p, a.n_taps, post, mac, a.stream);
}

// Host-only query behind firgpu_fir1d_fixed_plan: which datapath would serve these taps on uint8 / int16 containers.
int plan_fir1d_fixed(const int32_t *hq_host, int n_taps_in, int frac_bits, int acc_bits, int *path, int *post, int *mac)
{
    *path = FIRGPU_PATH_GENERIC;
    *post = *mac = 0;
    if (n_taps_in < 1 || n_taps_in > kMaxFastTaps || force_generic()) return 0;
    int n_taps = n_taps_in, bucket = 0;
    int32_t padded[kMaxFastTaps];
    const int32_t *taps_in = hq_host;
    if (n_taps_in > kMaxWideTaps) {
        for (int b : kLongBuckets)
            if (!bucket && b >= n_taps_in) bucket = b;
        const int off = bucket / 2 - n_taps_in / 2;
        for (int k = 0; k < bucket; ++k) padded[k] = (k >= off && k < off + n_taps_in) ? hq_host[k - off] : 0;
        taps_in = padded;
        n_taps = bucket;
    }
    FastPlan plans[kMaxFilters];
    if (select_plan(taps_in, n_taps, 1, frac_bits, acc_bits, bucket != 0, false, plans, post, mac)) *path = FIRGPU_PATH_STAGED;
    return 0;
}

int try_launch_fir1d_fixed_staged(const Fir1dArgs &a, const int32_t *hq_host)
{
    return try_launch_fir1d_fixed_staged_multi(a, hq_host, 1, 0);
}

// Wide datapath: int16 samples used at full width with caller-given clamp
// ranges (extended entry point).  Exactness proof: |sample| <= M and
// M * sum|hq| + bias < 2^31 and below the acc_bits wrap point.
int try_launch_fir1d_fixed_staged_wide(const Fir1dArgs &a, const int32_t *hq_host)
{
    if (a.in_dtype != FIRGPU_I16 || a.out_dtype != FIRGPU_I16) return 0;
    if (a.dilation != 1 && a.dilation != 2) return 0;
    if (a.dilation == 2 && ((a.cols | a.halo_left | a.halo_right) & 1)) return 0;
    if (a.n_taps > kMaxWideTaps || (a.centre >= 0 && a.centre != a.n_taps / 2)) return 0;
    if (a.frac_bits < 1 || a.frac_bits > 30) return 0;
    if (a.in_lo < -32768 || a.in_hi > 32767 || a.out_lo < -32768 || a.out_hi > 32767) return 0;
    if (force_generic()) return 0;
    long long sum_abs = 0;
    for (int k = 0; k < a.n_taps; ++k) sum_abs += hq_host[k] < 0 ? -(long long)hq_host[k] : (long long)hq_host[k];
    const long long m = std::max<long long>(a.in_lo < 0 ? -(long long)a.in_lo : a.in_lo, a.in_hi < 0 ? -(long long)a.in_hi : a.in_hi);
    const long long bias = 1LL << (a.frac_bits - 1);
    if (sum_abs > (1LL << 40)) return 0;
    const long long max_acc = m * sum_abs;
    if (max_acc + bias > 0x7fffffffLL) return 0;
    if (a.acc_bits < 63 && max_acc > (1LL << (a.acc_bits - 1)) - 1) return 0;

    StagedParams p{};
    p.x = (const unsigned char *)a.x;
    p.y = (unsigned char *)a.y;
    p.rows = a.rows;
    p.cols = a.cols;
    p.ld_x = a.ld_x;
    p.ld_y = a.ld_y;
    p.halo_left = a.halo_left;
    p.halo_right = a.halo_right;
    p.n_filters = 1;
    p.pdl = 1;
    if (a.hist) {
        if (a.rows != 1 || a.halo_left < 1 || a.cols < a.halo_left) return 0;
        p.hist = (const unsigned char *)a.hist;
        p.hist_out = (unsigned char *)a.hist_out;
    }
    p.shift = a.frac_bits;
    p.bias = (int)bias;
    p.x_bulk_ok = ((uintptr_t)a.x % 16 == 0) && (a.rows == 1 || (a.ld_x * 2) % 16 == 0);
    p.y_vec_ok = ((uintptr_t)a.y % 16 == 0) && (a.rows == 1 || (a.ld_y * 2) % 16 == 0);
    for (int k = 0; k < a.n_taps; ++k) p.hq[k] = hq_host[k];
    auto dup = [](int v) { return ((uint32_t)v & 0xffffu) | ((uint32_t)v << 16); };
    p.in_lo2 = dup(a.in_lo);
    p.in_hi2 = dup(a.in_hi);
    p.out_lo2 = dup(a.out_lo);
    p.out_hi2 = dup(a.out_hi);
    const bool clamp = !(a.in_lo == -32768 && a.in_hi == 32767 && a.out_lo == -32768 && a.out_hi == 32767);
    return launch_staged_wide(p, a.n_taps, a.dilation, clamp, a.stream);
}

}  // namespace firgpu
