// Host-side proof that the packed int32 datapath (dp2a / IMAD into an int32
// accumulator, packed 16-bit saturate) reproduces the reference's arithmetic
// exactly for a given set of integer taps, shared by the 1-D and 2-D fast
// kernels.  Reference semantics being preserved:
//   sum -> wrap to acc_bits -> + 2^(frac-1) -> >> frac -> clamp [0,255]
//   (fir_1d/model/python/fir_1d_fixed_ref.py:95-126).
#pragma once
#include <stdint.h>

namespace firgpu {

constexpr int kFastPlanMaxTaps = 129;

struct FastPlan {
    bool ok = false;
    bool hi16 = false;                       // taps pre-scaled so that the shift is exactly 16 bits
    bool wrap = false;                       // live acc_bits wrap evaluated in the post stage (plan_wrap)
    int shift = 0, bias = 0;
    int wrap_sh = 0;                         // wrap: 32 - acc_bits
    bool wrap_split = false;                 // wrap: general split-shift post form (the multiply-add form does not apply)
    int wrap_mul = 0;                        // wrap: 2^(acc_bits - frac_bits)
    uint32_t post_bias = 0, frac_mask = 0;   // wrap: 2^(frac-1), 2^frac - 1
    int32_t taps[kFastPlanMaxTaps] = {0};    // possibly pre-scaled
};

// tap_limit: 32767 when taps travel as signed 16-bit dp2a operands, INT32_MAX for IMAD.
inline FastPlan plan_fast(const int32_t *hq, int n_taps, int frac_bits, int acc_bits, int tap_limit, bool allow_hi16 = true)
{
    FastPlan f;
    if (n_taps < 1 || n_taps > kFastPlanMaxTaps || frac_bits < 1 || frac_bits > 30) return f;
    long long sum_abs = 0;
    for (int k = 0; k < n_taps; ++k) sum_abs += hq[k] < 0 ? -(long long)hq[k] : (long long)hq[k];
    const long long max_acc = 255LL * sum_abs;                  // |sum| can never exceed this
    // The reference wraps to acc_bits once per output; that is the identity
    // when the sum provably fits.
    if (acc_bits < 63 && max_acc > (1LL << (acc_bits - 1)) - 1) return f;
    if (allow_hi16 && frac_bits <= 16) {
        const int up = 16 - frac_bits;
        bool fits = (max_acc << up) + 32768 <= 0x7fffffffLL;
        for (int k = 0; fits && k < n_taps; ++k) {
            const long long sv = (long long)hq[k] * (1LL << up);
            if (sv > tap_limit || sv < -(long long)tap_limit - 1) fits = false;
        }
        if (fits) {
            f.ok = true;
            f.hi16 = true;
            f.shift = 16;
            f.bias = 32768;
            for (int k = 0; k < n_taps; ++k) f.taps[k] = (int32_t)((long long)hq[k] * (1LL << up));
            return f;
        }
    }
    const long long bias = 1LL << (frac_bits - 1);
    if (max_acc + bias > 0x7fffffffLL) return f;
    if (((max_acc + bias) >> frac_bits) > 32767) return f;     // packed 16-bit saturate stage
    for (int k = 0; k < n_taps; ++k)
        if (hq[k] > tap_limit || hq[k] < -(long long)tap_limit - 1) return f;
    f.ok = true;
    f.hi16 = false;
    f.shift = frac_bits;
    f.bias = (int)bias;
    for (int k = 0; k < n_taps; ++k) f.taps[k] = hq[k];
    return f;
}

// Wrap plan: int32 arithmetic is exact modulo 2^32, so for acc_bits <= 32 the low acc_bits bits of the int32 sum ARE
// the reference's wrapped accumulator whatever the taps and however large the true sum (fir_1d_fixed_ref.py:110-115);
// the post stage sign-extends them, adds the bias without re-wrapping (:118) and floor-shifts (:121).  The sum starts
// at 0, not at the bias.  tap_limit as in plan_fast (dp2a operands must fit int16; IMAD takes any int32).
inline FastPlan plan_wrap(const int32_t *hq, int n_taps, int frac_bits, int acc_bits, int tap_limit)
{
    FastPlan f;
    if (n_taps < 1 || n_taps > kFastPlanMaxTaps || frac_bits < 1 || frac_bits > 31 || acc_bits < 1) return f;
    long long sum_abs = 0;
    for (int k = 0; k < n_taps; ++k) {
        if (hq[k] > tap_limit || hq[k] < -(long long)tap_limit - 1) return f;
        sum_abs += hq[k] < 0 ? -(long long)hq[k] : (long long)hq[k];
    }
    if (acc_bits > 32) {                     // a wider accumulator never wraps a sum that fits int32: same as acc_bits = 32
        if (255LL * sum_abs > 0x7fffffffLL) return f;
        acc_bits = 32;
    }
    f.ok = true;
    f.wrap = true;
    f.shift = frac_bits;
    f.bias = 0;
    f.wrap_sh = 32 - acc_bits;
    // multiply-add post form: (a << wrap_sh) * 2^(acc_bits - frac_bits) + 2^31, upper word (see post_pair_mode)
    f.wrap_split = !(frac_bits <= acc_bits && acc_bits - frac_bits <= 30);
    f.wrap_mul = f.wrap_split ? 0 : (int)(1u << (acc_bits - frac_bits));
    f.post_bias = 1u << (frac_bits - 1);
    f.frac_mask = (1u << frac_bits) - 1u;
    for (int k = 0; k < n_taps; ++k) f.taps[k] = hq[k];
    return f;
}

// Coefficient for the sample at offset d from the output position:
// y[n] += h[hi - d] * x[n + d]   (fir_1d_fixed_ref.py:98-107 with idx = n - k + L//2).
inline int32_t tap_at(const int32_t *taps, int n_taps, int d)
{
    const int hi = n_taps / 2, lo = n_taps - 1 - hi;
    return (d < -lo || d > hi) ? 0 : taps[hi - d];
}
inline uint32_t pack_pair(int32_t first, int32_t second)
{
    return ((uint32_t)first & 0xffffu) | ((uint32_t)second << 16);
}
// dp2a tap pairs for windows that start on an even / odd sample (see
// fir1d_staged.cuh, mac_stream).
inline void build_pairs(const int32_t *taps, int n_taps, int n_pairs, uint32_t *even, uint32_t *odd)
{
    const int hi = n_taps / 2, lo = n_taps - 1 - hi;
    (void)hi;
    for (int k = 0; k < n_pairs; ++k) {
        even[k] = pack_pair(tap_at(taps, n_taps, -lo + 2 * k), tap_at(taps, n_taps, -lo + 2 * k + 1));
        odd[k] = pack_pair(tap_at(taps, n_taps, -lo - 1 + 2 * k), tap_at(taps, n_taps, -lo + 2 * k));
    }
}

}  // namespace firgpu
