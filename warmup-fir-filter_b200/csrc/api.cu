// C-ABI of libfirgpu (include/firgpu.h): argument checks, kernel-family
// dispatch, and the host-buffer pipelines.
//
// Nothing here computes a filter on the CPU.  Every compute entry point ends
// in a CUDA launch or fails with the cudaError_t it got; there is no fallback.
#include "common.cuh"
#include "fastplan.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

namespace firgpu {

namespace {
thread_local char tl_error[512] = "";
thread_local int tl_path = FIRGPU_PATH_NONE;
thread_local int64_t tl_launches = 0;
}  // namespace

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return 0;
    snprintf(tl_error, sizeof(tl_error), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

void note_launch(int path)
{
    tl_path = path;
    ++tl_launches;
}

namespace {

bool fixed_in_ok(int dt) { return dt >= FIRGPU_U8 && dt <= FIRGPU_F64; }

int check_1d(const char *fn, const void *x, const void *y, int in_dtype, int64_t rows, int64_t cols,
             int64_t ld_x, int64_t ld_y, const void *taps, int n_taps, int dilation,
             int64_t halo_left, int64_t halo_right)
{
    if (!fixed_in_ok(in_dtype)) return set_error(FIRGPU_E_DTYPE, "%s: unsupported in_dtype %d", fn, in_dtype);
    if (rows < 0 || cols < 0) return set_error(FIRGPU_E_SHAPE, "%s: negative shape rows=%lld cols=%lld", fn, (long long)rows, (long long)cols);
    if (rows > 1 && (ld_x < cols || ld_y < cols))
        return set_error(FIRGPU_E_SHAPE, "%s: leading dimension smaller than cols (ld_x=%lld ld_y=%lld cols=%lld)", fn,
                         (long long)ld_x, (long long)ld_y, (long long)cols);
    if (n_taps < 1 || n_taps > FIRGPU_MAX_TAPS) return set_error(FIRGPU_E_TAPS, "%s: n_taps=%d outside [1,%d]", fn, n_taps, FIRGPU_MAX_TAPS);
    if (dilation < 1) return set_error(FIRGPU_E_DILATION, "%s: dilation=%d must be >= 1", fn, dilation);
    if (halo_left < 0 || halo_right < 0) return set_error(FIRGPU_E_HALO, "%s: negative halo", fn);
    if (!taps) return set_error(FIRGPU_E_NULL, "%s: taps pointer is NULL", fn);
    if (rows > 0 && cols > 0 && (!x || !y)) return set_error(FIRGPU_E_NULL, "%s: NULL data pointer with non-empty shape", fn);
    return 0;
}

int check_2d(const char *fn, const void *x, const void *y, int in_dtype, int64_t n_img, int64_t height, int64_t width,
             int64_t ld_x, int64_t ld_y, const void *taps, int kh, int kw, int64_t halo_top, int64_t halo_bottom)
{
    if (!fixed_in_ok(in_dtype)) return set_error(FIRGPU_E_DTYPE, "%s: unsupported in_dtype %d", fn, in_dtype);
    if (n_img < 0 || height < 0 || width < 0) return set_error(FIRGPU_E_SHAPE, "%s: negative shape", fn);
    if (height > 1 && (ld_x < width || ld_y < width)) return set_error(FIRGPU_E_SHAPE, "%s: leading dimension smaller than width", fn);
    if (kh < 1 || kw < 1 || (int64_t)kh * kw > FIRGPU_MAX_TAPS)
        return set_error(FIRGPU_E_TAPS, "%s: kernel %dx%d outside [1,%d] taps", fn, kh, kw, FIRGPU_MAX_TAPS);
    if (halo_top < 0 || halo_bottom < 0) return set_error(FIRGPU_E_HALO, "%s: negative halo", fn);
    if (!taps) return set_error(FIRGPU_E_NULL, "%s: taps pointer is NULL", fn);
    if (n_img > 0 && height > 0 && width > 0 && (!x || !y)) return set_error(FIRGPU_E_NULL, "%s: NULL data pointer with non-empty shape", fn);
    return 0;
}

int check_bits(const char *fn, int frac_bits, int acc_bits)
{
    if (frac_bits < 1 || acc_bits < 1) return set_error(FIRGPU_E_BITS, "%s: frac_bits=%d and acc_bits=%d must be >= 1", fn, frac_bits, acc_bits);
    return 0;
}

// ---------------------------------------------------------------------------
// Host-buffer pipeline: chunks travel H2D -> kernel -> D2H on kSlots streams,
// so the two copy engines and the SMs overlap.  Pageable host memory is
// bounced through pinned slot buffers; pinned / registered memory is DMA'd
// directly.
// ---------------------------------------------------------------------------
constexpr int kSlots = 4;
constexpr size_t kChunkDefault = 32u << 20;  // input bytes per chunk

// FIRGPU_CHUNK_MB overrides the chunk size (tuning / tests of the segment path).
size_t chunk_target()
{
    static size_t v = 0;
    if (!v) {
        const char *e = getenv("FIRGPU_CHUNK_MB");
        const long mb = e ? atol(e) : 0;
        v = mb >= 1 && mb <= 1024 ? (size_t)mb << 20 : kChunkDefault;
    }
    return v;
}

struct HostCtx {
    std::mutex mu;
    cudaStream_t stream[kSlots] = {};
    void *d_in[kSlots] = {}, *d_out[kSlots] = {};
    void *p_in[kSlots] = {}, *p_out[kSlots] = {};
    size_t d_in_cap = 0, d_out_cap = 0, p_in_cap = 0, p_out_cap = 0;
    bool ready = false;
};

std::mutex g_ctx_mu;
std::map<int, HostCtx *> g_ctx;

HostCtx *ctx_for(int device)
{
    std::lock_guard<std::mutex> lock(g_ctx_mu);
    auto it = g_ctx.find(device);
    if (it != g_ctx.end()) return it->second;
    HostCtx *c = new HostCtx();
    g_ctx[device] = c;
    return c;
}

int ctx_reserve(HostCtx *c, size_t in_bytes, size_t out_bytes, bool pin_in, bool pin_out)
{
    if (!c->ready) {
        for (int s = 0; s < kSlots; ++s) FIRGPU_CUDA(cudaStreamCreateWithFlags(&c->stream[s], cudaStreamNonBlocking));
        c->ready = true;
    }
    auto grow = [](void **bufs, size_t *cap, size_t want, bool pinned) -> int {
        if (want <= *cap) return 0;
        for (int s = 0; s < kSlots; ++s) {
            if (bufs[s]) FIRGPU_CUDA(pinned ? cudaFreeHost(bufs[s]) : cudaFree(bufs[s]));
            bufs[s] = nullptr;
        }
        *cap = 0;
        for (int s = 0; s < kSlots; ++s)
            FIRGPU_CUDA(pinned ? cudaHostAlloc(&bufs[s], want, cudaHostAllocDefault) : cudaMalloc(&bufs[s], want));
        *cap = want;
        return 0;
    };
    int rc;
    if ((rc = grow(c->d_in, &c->d_in_cap, in_bytes, false))) return rc;
    if ((rc = grow(c->d_out, &c->d_out_cap, out_bytes, false))) return rc;
    if (pin_in && (rc = grow(c->p_in, &c->p_in_cap, in_bytes, true))) return rc;
    if (pin_out && (rc = grow(c->p_out, &c->p_out_cap, out_bytes, true))) return rc;
    return 0;
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

struct DeviceGuard {
    int prev = -1;
    bool active = false;
    int enter(int device)
    {
        if (device < 0) return 0;
        FIRGPU_CUDA(cudaGetDevice(&prev));
        if (prev != device) {
            FIRGPU_CUDA(cudaSetDevice(device));
            active = true;
        }
        return 0;
    }
    ~DeviceGuard()
    {
        if (active) cudaSetDevice(prev);
    }
};

// Leaves a host pipeline only after every slot stream is idle.  The success path has already waited for each slot;
// on an error exit, copies queued on the other slots may still be reading or writing the CALLER's (pinned) buffers
// and the slot buffers the next call will reuse.
struct SlotDrain {
    HostCtx *c;
    ~SlotDrain()
    {
        if (!c->ready) return;
        for (int s = 0; s < kSlots; ++s) cudaStreamSynchronize(c->stream[s]);
        cudaGetLastError();
    }
};

// A rectangular piece of a [rows x cols] problem plus the neighbour columns /
// rows that travel with it.
struct Piece {
    int64_t r0, nr;          // rows (1-D) or image rows incl. nothing of the halo (2-D)
    int64_t c0, nc;          // columns
    int64_t before, after;   // halo elements (1-D: scalars; 2-D: rows)
};

// Copy `nr` rows of `row_bytes` bytes between pitched buffers.
int copy_rows_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, size_t nr,
                    cudaMemcpyKind kind, cudaStream_t s)
{
    if (nr == 0 || row_bytes == 0) return 0;
    if (nr == 1 || (dpitch == row_bytes && spitch == row_bytes)) {
        FIRGPU_CUDA(cudaMemcpyAsync(dst, src, row_bytes * nr, kind, s));
    } else {
        FIRGPU_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, nr, kind, s));
    }
    return 0;
}

// Staging copies between pageable user memory and the pinned slot buffers run on
// several host threads: one core moves about 10 GB/s, PCIe wants 50.
int staging_threads()
{
    static int n = 0;
    if (!n) {
        const char *e = getenv("FIRGPU_STAGING_THREADS");
        const int forced = e ? atoi(e) : 0;
        const int hw = (int)std::thread::hardware_concurrency();
        n = forced >= 1 ? forced : std::max(1, std::min(8, hw / 2));
    }
    return n;
}

void copy_rows_serial(char *dst, size_t dpitch, const char *src, size_t spitch, size_t row_bytes, size_t r0, size_t r1)
{
    for (size_t r = r0; r < r1; ++r) memcpy(dst + r * dpitch, src + r * spitch, row_bytes);
}

void copy_rows_host(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, size_t nr)
{
    char *d = (char *)dst;
    const char *s = (const char *)src;
    if (nr == 1 || (dpitch == row_bytes && spitch == row_bytes)) {   // one contiguous block: treat it as rows of 1 MiB
        const size_t total = row_bytes * nr;
        row_bytes = dpitch = spitch = (size_t)1 << 20;
        nr = total / row_bytes;
        const size_t tail = total - nr * row_bytes;
        if (tail) memcpy(d + nr * row_bytes, s + nr * row_bytes, tail);
    }
    const size_t total = row_bytes * nr;
    int threads = staging_threads();
    if (total < ((size_t)4 << 20) || nr < 2) threads = 1;
    if ((size_t)threads > nr) threads = (int)nr;
    if (threads <= 1) {
        copy_rows_serial(d, dpitch, s, spitch, row_bytes, 0, nr);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve(threads - 1);
    for (int t = 1; t < threads; ++t)
        pool.emplace_back(copy_rows_serial, d, dpitch, s, spitch, row_bytes, nr * t / threads, nr * (t + 1) / threads);
    copy_rows_serial(d, dpitch, s, spitch, row_bytes, 0, nr / threads);
    for (auto &th : pool) th.join();
}

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// 1-D driver (n_out outputs per input, out_stride host elements apart; a piece's
// device outputs are ldd_y * nr elements apart).  `run(d_x, d_y, nr, nc, ldd_x, ldd_y, before, after, stream)`
// launches the device entry point on one piece.
template <typename Run>
int pipeline_1d(const void *x, void *y, int es, int os, int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                int64_t reach_before, int64_t reach_after, int device, Run run, int n_out = 1, int64_t out_stride = 0,
                int64_t avail_before = 0, int64_t avail_after = 0)
{
    // avail_before / avail_after (rows == 1 only): real samples of a longer host signal that lie in front of x / behind
    // x + cols — this call filters a segment of it (one device's share of a multi-device call) and reads them as halo.
    if (rows == 0 || cols == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    DeviceGuard guard;
    int rc = guard.enter(device);
    if (rc) return rc;
    int dev = 0;
    FIRGPU_CUDA(cudaGetDevice(&dev));
    HostCtx *c = ctx_for(dev);
    std::lock_guard<std::mutex> lock(c->mu);

    // Pieces: whole rows while a row fits the chunk target, else column
    // segments of one row with the taps' reach attached on both sides.
    std::vector<Piece> pieces;
    const int64_t row_bytes = cols * es;
    const int64_t q_in = 16 / es;                      // scalars per 16 bytes
    if ((size_t)row_bytes <= chunk_target() && avail_before == 0 && avail_after == 0) {
        int64_t per = (int64_t)(chunk_target() / (size_t)row_bytes);
        if (per < 1) per = 1;
        // keep at least kSlots pieces in flight when there is enough work
        if (rows >= 2 * kSlots && per > (rows + 2 * kSlots - 1) / (2 * kSlots) && (size_t)(row_bytes * rows) > (4u << 20))
            per = (rows + 2 * kSlots - 1) / (2 * kSlots);
        for (int64_t r = 0; r < rows; r += per) pieces.push_back({r, std::min(per, rows - r), 0, cols, 0, 0});
    } else {
        // (ramping the segment length up at the start of the job and down at its end, to shorten the only two
        // transfers that run with the other PCIe direction idle, was measured: no gain at 16-64 MiB chunks.)
        const int64_t seg = round_up((int64_t)(chunk_target() / es), 16384);
        for (int64_t r = 0; r < rows; ++r)
            for (int64_t c0 = 0; c0 < cols; c0 += seg) {
                const int64_t nc = std::min(seg, cols - c0);
                pieces.push_back({r, 1, c0, nc, std::min(reach_before, c0 + avail_before), std::min(reach_after, cols - c0 - nc + avail_after)});
            }
    }
    // Device layout of a piece: rows at a 16-byte-aligned pitch; a column
    // segment sits 16-byte aligned with its leading halo right before it.
    int64_t max_in = 0, max_out = 0;
    for (const Piece &p : pieces) {
        const int64_t ldd = round_up(p.nc + (p.before ? round_up(p.before, 256 / es) : 0) + p.after, q_in);
        max_in = std::max(max_in, ldd * p.nr * es + 256);
        max_out = std::max(max_out, round_up(p.nc, 16) * p.nr * os * n_out + 256);
    }
    const bool x_pinned = is_pinned(x), y_pinned = is_pinned(y);
    if ((rc = ctx_reserve(c, (size_t)max_in, (size_t)max_out, !x_pinned, !y_pinned))) return rc;
    SlotDrain drain{c};                                // on every exit, errors included: no copy of ours outlives the call

    const bool trace = getenv("FIRGPU_TRACE") != nullptr;
    double t_sync = 0, t_out = 0, t_in = 0, t_enq = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    struct Pending {
        bool busy = false;
        Piece p;
        int64_t ldd_y = 0;
    } pending[kSlots];
    auto finalize = [&](int s) -> int {
        if (!pending[s].busy) return 0;
        const double t0 = now();
        FIRGPU_CUDA(cudaStreamSynchronize(c->stream[s]));
        t_sync += now() - t0;
        const double t1 = now();
        if (!y_pinned) {
            const Piece &p = pending[s].p;
            for (int f = 0; f < n_out; ++f)
                copy_rows_host((char *)y + (f * out_stride + p.r0 * ld_y + p.c0) * os, (size_t)ld_y * os,
                               (char *)c->p_out[s] + (size_t)f * pending[s].ldd_y * p.nr * os,
                               (size_t)pending[s].ldd_y * os, (size_t)p.nc * os, (size_t)p.nr);
        }
        t_out += now() - t1;
        pending[s].busy = false;
        return 0;
    };

    for (size_t k = 0; k < pieces.size(); ++k) {
        const int s = (int)(k % kSlots);
        if ((rc = finalize(s))) return rc;
        const double t2 = now();
        const Piece &p = pieces[k];
        // A column segment is copied from `pre` scalars before its first one, with pre >= its halo and a multiple of
        // 256 bytes: source and destination of the DMA then start 256-byte aligned (segments start at multiples of
        // 16384 scalars), which measured 8 % faster than starting at the 2-scalar halo.  The extra scalars are ignored.
        const int64_t lead = p.before ? round_up(p.before, 256 / es) : 0;   // aligned offset of the segment in the slot
        const int64_t pre = lead <= p.c0 + avail_before ? lead : p.before;  // (a segment at the very start of the host signal has only its halo in front)
        const int64_t span = pre + p.nc + p.after;                     // scalars copied per row
        const int64_t ldd_x = round_up(lead + p.nc + p.after, q_in);
        const int64_t ldd_y = round_up(p.nc, 16);
        const char *src = (const char *)x + (p.r0 * ld_x + p.c0 - pre) * es;
        char *d_first = (char *)c->d_in[s] + (lead - pre) * es;        // where the first copied scalar lands
        if (x_pinned) {
            if ((rc = copy_rows_async(d_first, (size_t)ldd_x * es, src, (size_t)ld_x * es, (size_t)span * es,
                                      (size_t)p.nr, cudaMemcpyHostToDevice, c->stream[s])))
                return rc;
        } else {
            copy_rows_host((char *)c->p_in[s], (size_t)span * es, src, (size_t)ld_x * es, (size_t)span * es, (size_t)p.nr);
            t_in += now() - t2;
            if ((rc = copy_rows_async(d_first, (size_t)ldd_x * es, c->p_in[s], (size_t)span * es, (size_t)span * es,
                                      (size_t)p.nr, cudaMemcpyHostToDevice, c->stream[s])))
                return rc;
        }
        const double t3 = now();
        rc = run((const char *)c->d_in[s] + lead * es, c->d_out[s], p.nr, p.nc, ldd_x, ldd_y, p.before, p.after,
                 c->stream[s]);
        if (rc) return rc;
        for (int f = 0; f < n_out; ++f) {
            const size_t d_off = (size_t)f * ldd_y * p.nr * os;
            char *dst = y_pinned ? (char *)y + (f * out_stride + p.r0 * ld_y + p.c0) * os : (char *)c->p_out[s] + d_off;
            const size_t dpitch = y_pinned ? (size_t)ld_y * os : (size_t)ldd_y * os;
            if ((rc = copy_rows_async(dst, dpitch, (char *)c->d_out[s] + d_off, (size_t)ldd_y * os, (size_t)p.nc * os,
                                      (size_t)p.nr, cudaMemcpyDeviceToHost, c->stream[s])))
                return rc;
        }
        pending[s].busy = true;
        pending[s].p = p;
        pending[s].ldd_y = ldd_y;
        t_enq += now() - t3;
    }
    for (int s = 0; s < kSlots; ++s)
        if ((rc = finalize(s))) return rc;
    if (trace)
        fprintf(stderr, "[firgpu] pipeline_1d: %zu pieces, pinned in/out %d/%d, staging threads %d, sync %.2f ms, copy-out %.2f ms, copy-in %.2f ms, enqueue %.2f ms\n",
                pieces.size(), (int)x_pinned, (int)y_pinned, staging_threads(), t_sync * 1e3, t_out * 1e3, t_in * 1e3, t_enq * 1e3);
    return 0;
}

// 2-D driver: one image row band per piece, with the kernel's vertical reach
// attached above and below.
template <typename Run>
int pipeline_2d(const void *x, void *y, int es, int os, int64_t n_img, int64_t height, int64_t width,
                int64_t reach_above, int64_t reach_below, int device, Run run)
{
    if (n_img == 0 || height == 0 || width == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    DeviceGuard guard;
    int rc = guard.enter(device);
    if (rc) return rc;
    int dev = 0;
    FIRGPU_CUDA(cudaGetDevice(&dev));
    HostCtx *c = ctx_for(dev);
    std::lock_guard<std::mutex> lock(c->mu);

    const int64_t row_bytes = width * es;
    int64_t band = (int64_t)(chunk_target() / (size_t)std::max<int64_t>(row_bytes, 1));
    if (band < 8) band = 8;
    if (band > height) band = height;
    std::vector<Piece> pieces;     // r0 = image * height + first row
    for (int64_t b = 0; b < n_img; ++b)
        for (int64_t r = 0; r < height; r += band) {
            const int64_t nr = std::min(band, height - r);
            pieces.push_back({b * height + r, nr, 0, width, std::min(reach_above, r), std::min(reach_below, height - r - nr)});
        }
    const int64_t ldd_x = round_up(width, 16 / es), ldd_y = round_up(width, 16);
    const size_t max_in = (size_t)(ldd_x * (band + reach_above + reach_below) * es + 256);
    const size_t max_out = (size_t)(ldd_y * band * os + 256);
    const bool x_pinned = is_pinned(x), y_pinned = is_pinned(y);
    if ((rc = ctx_reserve(c, max_in, max_out, !x_pinned, !y_pinned))) return rc;
    SlotDrain drain{c};

    struct Pending {
        bool busy = false;
        Piece p;
    } pending[kSlots];
    auto finalize = [&](int s) -> int {
        if (!pending[s].busy) return 0;
        FIRGPU_CUDA(cudaStreamSynchronize(c->stream[s]));
        if (!y_pinned) {
            const Piece &p = pending[s].p;
            copy_rows_host((char *)y + p.r0 * width * os, (size_t)width * os, c->p_out[s], (size_t)ldd_y * os,
                           (size_t)width * os, (size_t)p.nr);
        }
        pending[s].busy = false;
        return 0;
    };
    for (size_t k = 0; k < pieces.size(); ++k) {
        const int s = (int)(k % kSlots);
        if ((rc = finalize(s))) return rc;
        const Piece &p = pieces[k];
        const int64_t nrows = p.before + p.nr + p.after;
        const char *src = (const char *)x + (p.r0 - p.before) * width * es;
        if (x_pinned) {
            if ((rc = copy_rows_async(c->d_in[s], (size_t)ldd_x * es, src, (size_t)width * es, (size_t)width * es,
                                      (size_t)nrows, cudaMemcpyHostToDevice, c->stream[s])))
                return rc;
        } else {
            copy_rows_host(c->p_in[s], (size_t)width * es, src, (size_t)width * es, (size_t)width * es, (size_t)nrows);
            if ((rc = copy_rows_async(c->d_in[s], (size_t)ldd_x * es, c->p_in[s], (size_t)width * es, (size_t)width * es,
                                      (size_t)nrows, cudaMemcpyHostToDevice, c->stream[s])))
                return rc;
        }
        rc = run((const char *)c->d_in[s] + p.before * ldd_x * es, c->d_out[s], p.nr, width, ldd_x, ldd_y, p.before, p.after,
                 c->stream[s]);
        if (rc) return rc;
        char *dst = y_pinned ? (char *)y + p.r0 * width * os : (char *)c->p_out[s];
        const size_t dpitch = y_pinned ? (size_t)width * os : (size_t)ldd_y * os;
        if ((rc = copy_rows_async(dst, dpitch, c->d_out[s], (size_t)ldd_y * os, (size_t)width * os, (size_t)p.nr,
                                  cudaMemcpyDeviceToHost, c->stream[s])))
            return rc;
        pending[s].busy = true;
        pending[s].p = p;
    }
    for (int s = 0; s < kSlots; ++s)
        if ((rc = finalize(s))) return rc;
    return 0;
}

bool wide_container(int in_dtype) { return in_dtype == FIRGPU_I32 || in_dtype == FIRGPU_F32 || in_dtype == FIRGPU_F64; }

// int32 / float containers of a useful size are converted to the reference's uint8 pixels first (csrc/prep_u8.cu).
bool convert_first(int in_dtype, int64_t samples)
{
    return wide_container(in_dtype) && samples >= 4096 && !force_generic();
}

// uint8 copy of rows x (halo_left + cols + halo_right) samples in scratch memory: *body points at column 0 of row 0
// (16-byte aligned), *ld_t is the pitch (a multiple of 16).  The caller frees *tmp with cudaFreeAsync on `stream`.
int uint8_copy_1d(const void *x, int in_dtype, int64_t rows, int64_t cols, int64_t ld_x, int64_t halo_left, int64_t halo_right,
                  cudaStream_t stream, uint8_t **tmp, uint8_t **body, int64_t *ld_t)
{
    const int64_t lead = (halo_left + 127) / 128 * 128;                  // body and pitch on 128-byte (L2 line) boundaries
    *ld_t = (lead + cols + halo_right + 127) / 128 * 128;
    int rc = scratch_alloc((void **)tmp, (size_t)(*ld_t * rows), stream);
    if (rc) return rc;
    *body = *tmp + lead;
    rc = launch_prep_u8(x, in_dtype, *tmp, rows, halo_left + cols + halo_right, ld_x, *ld_t, halo_left, lead - halo_left, stream);
    if (rc) cudaFreeAsync(*tmp, stream);
    return rc;
}

int out_size_fixed(int out_dtype) { return out_dtype == FIRGPU_U8 ? 1 : out_dtype == FIRGPU_I16 ? 2 : 0; }

// "Flattened pass": a dense batch of rows (pitch == cols) is filtered by the staged kernel as ONE row of rows * cols
// samples — base pointers are 16-byte aligned, so every tile is a bulk copy whatever the row pitch (odd image widths,
// the reference's 2999 x 4499 image) and no tile is left mostly empty by a short row.  The kernel is told the row
// length (`period`) and forms the sums of the few chunks whose window straddles a row boundary once per side, so each
// row keeps its own zero padding (fir_1d/sim/vector/gen_fixed_output.py:44-59).  One launch.
// Rows wider than kFlattenMaxBytes with a 16-byte aligned pitch stay on the per-row paths (row-streaming tensor-map
// kernel up to 8 KiB, per-row tile walk beyond), which carry no per-chunk boundary test.  Measured, int16, 2^27 samples,
// fraction of the HBM peak, in-kernel row ends / per-row paths: 128 B rows 0.81 / 0.15, 256 B 0.81 / 0.30, 512 B 0.81 / 0.55,
// 1 KiB 0.81 / 0.81, 2 KiB 0.85 / 0.68, 4 KiB 0.88 / 0.85, 8 KiB 0.83 / 0.96, 2 MiB 0.98 / 1.02 (profiles/README.md R2.4).
constexpr long kFlattenMaxBytes = 4096;

bool flatten_ok(const void *x, const void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                int n_taps, int dilation)
{
    if (rows < 2 || ld_x != cols || ld_y != cols) return false;
    if (in_dtype != FIRGPU_U8 && in_dtype != FIRGPU_I16) return false;
    const int es = in_dtype == FIRGPU_U8 ? 1 : 2, os = out_size_fixed(out_dtype);
    static const long max_bytes = [] {
        const char *e = getenv("FIRGPU_FLATTEN_MAX_BYTES");
        return e ? atol(e) : kFlattenMaxBytes;
    }();
    static const bool disabled = getenv("FIRGPU_NO_FLATTEN") != nullptr;
    if (disabled || force_generic()) return false;
    if ((cols * es) % 16 == 0 && (cols * os) % 16 == 0 && cols * es > max_bytes) return false;
    if (((uintptr_t)x | (uintptr_t)y) % 16) return false;
    if (dilation == 2 && (cols & 1)) return false;                             // I/Q pairs must not straddle rows
    (void)n_taps;                                                              // (the launcher checks cols against the chunk window)
    return true;
}
int out_size_ideal(int out_dtype) { return out_dtype == FIRGPU_F64 ? 8 : out_dtype == FIRGPU_F32 ? 4 : 0; }

}  // namespace

// Stream-ordered scratch memory (the uint8 copy of int32 / float containers).  A pool of the library's own, one per
// device, that keeps what it has allocated between calls (the device's default pool hands its memory back at
// every synchronisation, which would put a fresh cuMemCreate on each call); firgpu_release_cache() trims it.
namespace {
std::mutex g_pool_mu;
std::map<int, cudaMemPool_t> g_pools;
}  // namespace

int scratch_alloc(void **ptr, size_t bytes, cudaStream_t stream)
{
    int dev = 0;
    FIRGPU_CUDA(cudaGetDevice(&dev));
    cudaMemPool_t pool = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        auto it = g_pools.find(dev);
        if (it == g_pools.end()) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            FIRGPU_CUDA(cudaMemPoolCreate(&pool, &props));
            uint64_t keep = ~0ULL;
            FIRGPU_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
            g_pools[dev] = pool;
        } else {
            pool = it->second;
        }
    }
    FIRGPU_CUDA(cudaMallocFromPoolAsync(ptr, bytes ? bytes : 1, pool, stream));
    return 0;
}

// Extended fixed-point call on device memory (arguments already checked).
int fir1d_fixed_ex_device(const Fir1dArgs &a, const int32_t *hq, bool fast_only)
{
    // Any alignment is the centred filter with zero taps in front of / behind the real ones (a zero tap adds nothing
    // to the integer sum), so causal and other off-centre filters run on the centred fast kernels too.
    Fir1dArgs fast = a;
    int32_t padded[kFastPlanMaxTaps];
    const int32_t *fast_taps = hq;
    const int n_taps = a.n_taps, c = a.centre < 0 ? n_taps / 2 : a.centre;
    if (c != n_taps / 2) {
        const int reach = std::max(c, n_taps - 1 - c), wide_taps = 2 * reach + 1, off = reach - c;
        if (wide_taps <= kFastPlanMaxTaps) {
            for (int k = 0; k < wide_taps; ++k) padded[k] = (k >= off && k < off + n_taps) ? hq[k - off] : 0;
            fast.n_taps = wide_taps;
            fast.centre = -1;
            fast_taps = padded;
        }
    }
    int rc;
    if (fast.is_reference_model()) {        // the reference's own parameters: the fast kernels apply
        rc = try_launch_fir1d_fixed_staged(fast, fast_taps);
    } else {                                // other clamp ranges on int16 containers: wide datapath
        rc = try_launch_fir1d_fixed_staged_wide(fast, fast_taps);
    }
    if (rc != 0 || fast_only) return rc;    // served (1), error, or "no fast kernel applies" for a caller that has its own fallback
    rc = launch_fir1d_fixed_generic(a, hq);
    return rc ? rc : 1;
}

}  // namespace firgpu

using namespace firgpu;

extern "C" {

int firgpu_version(void) { return FIRGPU_VERSION; }
const char *firgpu_last_error(void) { return tl_error; }
int firgpu_last_path(void) { return tl_path; }
int64_t firgpu_launch_count(void) { return tl_launches; }

int firgpu_device_count(int *count)
{
    if (!count) return set_error(FIRGPU_E_NULL, "firgpu_device_count: NULL");
    FIRGPU_CUDA(cudaGetDeviceCount(count));
    return 0;
}

int firgpu_fir1d_fixed(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols, int64_t ld_x,
                       int64_t ld_y, const int32_t *hq, int n_taps, int dilation, int frac_bits, int acc_bits,
                       int64_t halo_left_valid, int64_t halo_right_valid, void *stream)
{
    int rc = check_1d("firgpu_fir1d_fixed", x, y, in_dtype, rows, cols, ld_x, ld_y, hq, n_taps, dilation,
                      halo_left_valid, halo_right_valid);
    if (rc) return rc;
    if ((rc = check_bits("firgpu_fir1d_fixed", frac_bits, acc_bits))) return rc;
    if (!out_size_fixed(out_dtype)) return set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_fixed: out_dtype must be U8 or I16, got %d", out_dtype);
    if (rows == 0 || cols == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    Fir1dArgs a{x, y, in_dtype, out_dtype, rows, cols, ld_x, ld_y, n_taps, dilation, frac_bits, acc_bits,
                halo_left_valid, halo_right_valid, (cudaStream_t)stream};
    // int32 / float containers: convert to the reference's uint8 pixels in one streaming
    // pass (into a stream-ordered temporary), then filter on the fast kernels.
    if (convert_first(in_dtype, rows * cols) && n_taps <= 65 && (dilation == 1 || dilation == 2)) {
        uint8_t *tmp = nullptr, *body = nullptr;
        int64_t ld_t = 0;
        if ((rc = uint8_copy_1d(x, in_dtype, rows, cols, ld_x, halo_left_valid, halo_right_valid, (cudaStream_t)stream, &tmp, &body, &ld_t)))
            return rc;
        // the uint8 copy then takes the ordinary routing below (short-row kernel, staged kernel, generic kernel)
        rc = firgpu_fir1d_fixed(body, y, FIRGPU_U8, out_dtype, rows, cols, ld_t, ld_y, hq, n_taps, dilation, frac_bits, acc_bits,
                                halo_left_valid, halo_right_valid, stream);
        cudaFreeAsync(tmp, (cudaStream_t)stream);
        return rc;
    }
    if (flatten_ok(x, y, in_dtype, out_dtype, rows, cols, ld_x, ld_y, n_taps, dilation) && halo_left_valid == 0 && halo_right_valid == 0) {
        Fir1dArgs flat = a;
        flat.rows = 1;
        flat.cols = flat.ld_x = flat.ld_y = rows * cols;
        flat.period = cols;
        rc = try_launch_fir1d_fixed_staged(flat, hq);
        if (rc == 1) return 0;
        if (rc != 0) return rc;
    }
    // Many short rows (images, short channels): the row-streaming tensor-map
    // kernel keeps whole CTAs busy where the 1-D tile kernel would stage one
    // mostly empty tile per row.  Its TMA out-of-bounds zero fill is the
    // per-row zero padding.
    if (rows >= 16 && cols * dtype_size(in_dtype) <= 8192 && dilation == 1 && halo_left_valid == 0 && halo_right_valid == 0 &&
        in_dtype == out_dtype) {
        Fir2dArgs b{x, y, in_dtype, out_dtype, 1, rows, cols, ld_x, ld_y, 0, 0, 1, n_taps, frac_bits, acc_bits, 0, 0,
                    (cudaStream_t)stream};
        rc = try_launch_fir2d_fixed_tiled(b, hq);
        if (rc == 1) return 0;
        if (rc != 0) return rc;
    }
    rc = try_launch_fir1d_fixed_staged(a, hq);
    if (rc == 1) return 0;
    if (rc != 0) return rc;
    return launch_fir1d_fixed_generic(a, hq);
}

int firgpu_fir1d_fixed_multi(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols,
                             int64_t ld_x, int64_t ld_y, int64_t filter_stride_y, const int32_t *hq, int n_filters,
                             int n_taps, int dilation, int frac_bits, int acc_bits, void *stream)
{
    if (n_filters < 0) return set_error(FIRGPU_E_SHAPE, "firgpu_fir1d_fixed_multi: n_filters=%d", n_filters);
    const int os = out_size_fixed(out_dtype);
    if (!os) return set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_fixed_multi: out_dtype must be U8 or I16, got %d", out_dtype);
    int rc = check_1d("firgpu_fir1d_fixed_multi", x, y, in_dtype, rows, cols, ld_x, ld_y, hq, n_taps, dilation, 0, 0);
    if (rc) return rc;
    if ((rc = check_bits("firgpu_fir1d_fixed_multi", frac_bits, acc_bits))) return rc;
    if (rows == 0 || cols == 0 || n_filters == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    if (convert_first(in_dtype, rows * cols)) {          // one conversion for all tap sets
        uint8_t *tmp = nullptr, *body = nullptr;
        int64_t ld_t = 0;
        if ((rc = uint8_copy_1d(x, in_dtype, rows, cols, ld_x, 0, 0, (cudaStream_t)stream, &tmp, &body, &ld_t))) return rc;
        rc = firgpu_fir1d_fixed_multi(body, y, FIRGPU_U8, out_dtype, rows, cols, ld_t, ld_y, filter_stride_y, hq, n_filters,
                                      n_taps, dilation, frac_bits, acc_bits, stream);
        cudaFreeAsync(tmp, (cudaStream_t)stream);
        return rc;
    }
    // Groups of four tap sets are served by ONE pass of the staged kernel (the
    // input is read from HBM once); the rest one tap set at a time.
    const bool flatten = flatten_ok(x, y, in_dtype, out_dtype, rows, cols, ld_x, ld_y, n_taps, dilation) &&
                         ((size_t)filter_stride_y * os) % 16 == 0;
    int f = 0;
    for (; f + 4 <= n_filters; f += 4) {
        Fir1dArgs a{x, (char *)y + (size_t)f * (size_t)filter_stride_y * os, in_dtype, out_dtype, rows, cols, ld_x, ld_y,
                    n_taps, dilation, frac_bits, acc_bits, 0, 0, (cudaStream_t)stream};
        if (flatten) {                                // dense rows: one long row with in-kernel row ends
            Fir1dArgs flat = a;
            flat.rows = 1;
            flat.cols = flat.ld_x = flat.ld_y = rows * cols;
            flat.period = cols;
            rc = try_launch_fir1d_fixed_staged_multi(flat, hq + (size_t)f * n_taps, 4, filter_stride_y);
            if (rc == 1) continue;
            if (rc != 0) return rc;
        }
        rc = try_launch_fir1d_fixed_staged_multi(a, hq + (size_t)f * n_taps, 4, filter_stride_y);
        if (rc == 1) continue;
        if (rc != 0) return rc;
        break;                                        // not eligible: fall through to single passes
    }
    for (; f < n_filters; ++f) {
        rc = firgpu_fir1d_fixed(x, (char *)y + (size_t)f * (size_t)filter_stride_y * os, in_dtype, out_dtype, rows,
                                cols, ld_x, ld_y, hq + (size_t)f * n_taps, n_taps, dilation, frac_bits, acc_bits, 0, 0,
                                stream);
        if (rc) return rc;
    }
    return 0;
}

int firgpu_fir1d_ideal(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols, int64_t ld_x,
                       int64_t ld_y, const double *h, int n_taps, int dilation, int64_t halo_left_valid,
                       int64_t halo_right_valid, void *stream)
{
    int rc = check_1d("firgpu_fir1d_ideal", x, y, in_dtype, rows, cols, ld_x, ld_y, h, n_taps, dilation, halo_left_valid,
                      halo_right_valid);
    if (rc) return rc;
    if (!out_size_ideal(out_dtype)) return set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_ideal: out_dtype must be F64 or F32, got %d", out_dtype);
    if (rows == 0 || cols == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    if (convert_first(in_dtype, rows * cols) && n_taps <= 9 && (dilation == 1 || dilation == 2)) {   // the floating-point model also starts from the rounded, clamped pixel
        uint8_t *tmp = nullptr, *body = nullptr;
        int64_t ld_t = 0;
        if ((rc = uint8_copy_1d(x, in_dtype, rows, cols, ld_x, halo_left_valid, halo_right_valid, (cudaStream_t)stream, &tmp, &body, &ld_t)))
            return rc;
        Fir1dArgs t{body, y, FIRGPU_U8, out_dtype, rows, cols, ld_t, ld_y, n_taps, dilation, 0, 0,
                    halo_left_valid, halo_right_valid, (cudaStream_t)stream};
        rc = launch_fir1d_ideal(t, h);
        cudaFreeAsync(tmp, (cudaStream_t)stream);
        return rc;
    }
    Fir1dArgs a{x, y, in_dtype, out_dtype, rows, cols, ld_x, ld_y, n_taps, dilation, 0, 0,
                halo_left_valid, halo_right_valid, (cudaStream_t)stream};
    return launch_fir1d_ideal(a, h);
}

int firgpu_fir2d_fixed(const void *x, void *y, int in_dtype, int out_dtype, int64_t n_img, int64_t height, int64_t width,
                       int64_t ld_x, int64_t ld_y, int64_t img_stride_x, int64_t img_stride_y, const int32_t *hq, int kh,
                       int kw, int frac_bits, int acc_bits, int64_t halo_top_valid, int64_t halo_bottom_valid, void *stream)
{
    int rc = check_2d("firgpu_fir2d_fixed", x, y, in_dtype, n_img, height, width, ld_x, ld_y, hq, kh, kw, halo_top_valid,
                      halo_bottom_valid);
    if (rc) return rc;
    if ((rc = check_bits("firgpu_fir2d_fixed", frac_bits, acc_bits))) return rc;
    if (!out_size_fixed(out_dtype)) return set_error(FIRGPU_E_DTYPE, "firgpu_fir2d_fixed: out_dtype must be U8 or I16, got %d", out_dtype);
    if (n_img == 0 || height == 0 || width == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    // Images the tensor-map kernel cannot read in place are copied once into an aligned temporary and filtered there:
    // int32 / float containers (rounded and clamped to uint8 pixels on the way), and uint8 / int16 images whose pitch
    // or base is not 16-byte aligned or whose uint8 width is odd (every row of the copy ends in zero bytes).
    const bool windows_ok = fir2d_tiled_window_ok(kh, kw);
    const int es_in = dtype_size(in_dtype);
    const bool convert = out_dtype == FIRGPU_U8 && convert_first(in_dtype, n_img * height * width);
    const bool repitch = !convert && (in_dtype == FIRGPU_U8 || in_dtype == FIRGPU_I16) &&
                         (in_dtype == out_dtype || (in_dtype == FIRGPU_I16 && out_dtype == FIRGPU_U8)) &&
                         n_img * height * width >= 4096 && !force_generic() &&
                         ((uintptr_t)x % 16 != 0 || (ld_x * es_in) % 16 != 0 || (n_img > 1 && (img_stride_x * es_in) % 16 != 0) ||
                          (in_dtype == FIRGPU_U8 && (width & 1)));
    if (windows_ok && (convert || repitch)) {
        const int t_dtype = convert ? FIRGPU_U8 : in_dtype;
        const int es_t = dtype_size(t_dtype);
        const int64_t band = halo_top_valid + height + halo_bottom_valid;
        // pitch: a multiple of 128 bytes (whole L2 lines per row), with at least one zero element after an odd uint8 width
        const int64_t ld_t = ((width + (t_dtype == FIRGPU_U8 ? (width & 1) : 0)) * es_t + 127) / 128 * 128 / es_t;
        uint8_t *tmp = nullptr;
        if ((rc = scratch_alloc((void **)&tmp, (size_t)(ld_t * band * n_img * es_t), (cudaStream_t)stream))) return rc;
        for (int64_t b = 0; b < n_img && !rc; ++b) {
            const char *src = (const char *)x + (b * img_stride_x - halo_top_valid * ld_x) * es_in;
            uint8_t *dst = tmp + b * band * ld_t * es_t;
            rc = convert ? launch_prep_u8(src, in_dtype, dst, band, width, ld_x, ld_t, 0, 0, (cudaStream_t)stream)
                         : launch_repitch(src, in_dtype, dst, band, width, ld_x, ld_t, (cudaStream_t)stream);
        }
        if (!rc) {
            Fir2dArgs t{tmp + halo_top_valid * ld_t * es_t, y, t_dtype, out_dtype, n_img, height, width, ld_t, ld_y, band * ld_t,
                        img_stride_y, kh, kw, frac_bits, acc_bits, halo_top_valid, halo_bottom_valid, (cudaStream_t)stream};
            t.x_row_padded = 1;
            rc = try_launch_fir2d_fixed_tiled(t, hq);
            if (rc == 1) rc = 0;
            else if (rc == 0) rc = launch_fir2d_fixed_generic(t, hq);
        }
        cudaFreeAsync(tmp, (cudaStream_t)stream);
        return rc;
    }
    Fir2dArgs a{x, y, in_dtype, out_dtype, n_img, height, width, ld_x, ld_y, img_stride_x, img_stride_y, kh, kw,
                frac_bits, acc_bits, halo_top_valid, halo_bottom_valid, (cudaStream_t)stream};
    rc = try_launch_fir2d_fixed_tiled(a, hq);
    if (rc == 1) return 0;
    if (rc != 0) return rc;
    return launch_fir2d_fixed_generic(a, hq);
}

int firgpu_fir2d_ideal(const void *x, void *y, int in_dtype, int out_dtype, int64_t n_img, int64_t height, int64_t width,
                       int64_t ld_x, int64_t ld_y, int64_t img_stride_x, int64_t img_stride_y, const double *h, int kh,
                       int kw, int64_t halo_top_valid, int64_t halo_bottom_valid, void *stream)
{
    int rc = check_2d("firgpu_fir2d_ideal", x, y, in_dtype, n_img, height, width, ld_x, ld_y, h, kh, kw, halo_top_valid,
                      halo_bottom_valid);
    if (rc) return rc;
    if (!out_size_ideal(out_dtype)) return set_error(FIRGPU_E_DTYPE, "firgpu_fir2d_ideal: out_dtype must be F64 or F32, got %d", out_dtype);
    if (n_img == 0 || height == 0 || width == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    // int32 / float images: rounded and clamped to uint8 pixels first (what the model does to every sample anyway), so
    // that the register kernel applies
    if (convert_first(in_dtype, n_img * height * width) && (kh == 3 || kh == 5) && (kw == 3 || kw == 5)) {
        const int es_in = dtype_size(in_dtype);
        const int64_t band = halo_top_valid + height + halo_bottom_valid;
        const int64_t ld_t = (width + 15) / 16 * 16;
        uint8_t *tmp = nullptr;
        if ((rc = scratch_alloc((void **)&tmp, (size_t)(ld_t * band * n_img), (cudaStream_t)stream))) return rc;
        for (int64_t b = 0; b < n_img && !rc; ++b)
            rc = launch_prep_u8((const char *)x + (b * img_stride_x - halo_top_valid * ld_x) * es_in, in_dtype, tmp + b * band * ld_t,
                                band, width, ld_x, ld_t, 0, 0, (cudaStream_t)stream);
        if (!rc) {
            Fir2dArgs t{tmp + halo_top_valid * ld_t, y, FIRGPU_U8, out_dtype, n_img, height, width, ld_t, ld_y, band * ld_t,
                        img_stride_y, kh, kw, 0, 0, halo_top_valid, halo_bottom_valid, (cudaStream_t)stream};
            rc = launch_fir2d_ideal(t, h);
        }
        cudaFreeAsync(tmp, (cudaStream_t)stream);
        return rc;
    }
    Fir2dArgs a{x, y, in_dtype, out_dtype, n_img, height, width, ld_x, ld_y, img_stride_x, img_stride_y, kh, kw, 0, 0,
                halo_top_valid, halo_bottom_valid, (cudaStream_t)stream};
    return launch_fir2d_ideal(a, h);
}

// ---- host-buffer entry points ---------------------------------------------
int firgpu_fir1d_fixed_host(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols, int64_t ld_x,
                            int64_t ld_y, const int32_t *hq, int n_taps, int dilation, int frac_bits, int acc_bits,
                            int device)
{
    int rc = check_1d("firgpu_fir1d_fixed_host", x, y, in_dtype, rows, cols, ld_x, ld_y, hq, n_taps, dilation, 0, 0);
    if (rc) return rc;
    if ((rc = check_bits("firgpu_fir1d_fixed_host", frac_bits, acc_bits))) return rc;
    const int os = out_size_fixed(out_dtype);
    if (!os) return set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_fixed_host: out_dtype must be U8 or I16, got %d", out_dtype);
    const int64_t centre = n_taps / 2;
    return pipeline_1d(x, y, dtype_size(in_dtype), os, rows, cols, ld_x, ld_y, (n_taps - 1 - centre) * (int64_t)dilation,
                       centre * (int64_t)dilation, device,
                       [&](const void *dx, void *dy, int64_t nr, int64_t nc, int64_t lx, int64_t ly, int64_t before,
                           int64_t after, cudaStream_t s) {
                           return firgpu_fir1d_fixed(dx, dy, in_dtype, out_dtype, nr, nc, lx, ly, hq, n_taps, dilation,
                                                     frac_bits, acc_bits, before, after, (void *)s);
                       });
}

int firgpu_fir1d_fixed_multi_host(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols,
                                  int64_t ld_x, int64_t ld_y, int64_t filter_stride_y, const int32_t *hq, int n_filters,
                                  int n_taps, int dilation, int frac_bits, int acc_bits, int device)
{
    int rc = check_1d("firgpu_fir1d_fixed_multi_host", x, y, in_dtype, rows, cols, ld_x, ld_y, hq, n_taps, dilation, 0, 0);
    if (rc) return rc;
    if ((rc = check_bits("firgpu_fir1d_fixed_multi_host", frac_bits, acc_bits))) return rc;
    const int os = out_size_fixed(out_dtype);
    if (!os) return set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_fixed_multi_host: out_dtype must be U8 or I16, got %d", out_dtype);
    if (n_filters < 1 || n_filters > 64) return set_error(FIRGPU_E_SHAPE, "firgpu_fir1d_fixed_multi_host: n_filters=%d outside [1,64]", n_filters);
    const int64_t centre = n_taps / 2;
    return pipeline_1d(x, y, dtype_size(in_dtype), os, rows, cols, ld_x, ld_y, (n_taps - 1 - centre) * (int64_t)dilation,
                       centre * (int64_t)dilation, device,
                       [&](const void *dx, void *dy, int64_t nr, int64_t nc, int64_t lx, int64_t ly, int64_t before,
                           int64_t after, cudaStream_t s) {
                           if (before == 0 && after == 0)
                               return firgpu_fir1d_fixed_multi(dx, dy, in_dtype, out_dtype, nr, nc, lx, ly, ly * nr, hq,
                                                               n_filters, n_taps, dilation, frac_bits, acc_bits, (void *)s);
                           // column segments of one very long row carry halos: one tap set at a time
                           for (int f = 0; f < n_filters; ++f) {
                               int r = firgpu_fir1d_fixed(dx, (char *)dy + (size_t)f * ly * nr * os, in_dtype, out_dtype, nr, nc, lx,
                                                          ly, hq + (size_t)f * n_taps, n_taps, dilation, frac_bits, acc_bits,
                                                          before, after, (void *)s);
                               if (r) return r;
                           }
                           return 0;
                       },
                       n_filters, filter_stride_y);
}

int firgpu_fir1d_ideal_host(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols, int64_t ld_x,
                            int64_t ld_y, const double *h, int n_taps, int dilation, int device)
{
    int rc = check_1d("firgpu_fir1d_ideal_host", x, y, in_dtype, rows, cols, ld_x, ld_y, h, n_taps, dilation, 0, 0);
    if (rc) return rc;
    const int os = out_size_ideal(out_dtype);
    if (!os) return set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_ideal_host: out_dtype must be F64 or F32, got %d", out_dtype);
    const int64_t centre = n_taps / 2;
    return pipeline_1d(x, y, dtype_size(in_dtype), os, rows, cols, ld_x, ld_y, (n_taps - 1 - centre) * (int64_t)dilation,
                       centre * (int64_t)dilation, device,
                       [&](const void *dx, void *dy, int64_t nr, int64_t nc, int64_t lx, int64_t ly, int64_t before,
                           int64_t after, cudaStream_t s) {
                           return firgpu_fir1d_ideal(dx, dy, in_dtype, out_dtype, nr, nc, lx, ly, h, n_taps, dilation,
                                                     before, after, (void *)s);
                       });
}

int firgpu_fir2d_fixed_host(const void *x, void *y, int in_dtype, int out_dtype, int64_t n_img, int64_t height,
                            int64_t width, const int32_t *hq, int kh, int kw, int frac_bits, int acc_bits, int device)
{
    int rc = check_2d("firgpu_fir2d_fixed_host", x, y, in_dtype, n_img, height, width, width, width, hq, kh, kw, 0, 0);
    if (rc) return rc;
    if ((rc = check_bits("firgpu_fir2d_fixed_host", frac_bits, acc_bits))) return rc;
    const int os = out_size_fixed(out_dtype);
    if (!os) return set_error(FIRGPU_E_DTYPE, "firgpu_fir2d_fixed_host: out_dtype must be U8 or I16, got %d", out_dtype);
    const int64_t ch = kh / 2;
    // rows above come from i - m + ch with m up to kh-1; rows below from m = 0
    return pipeline_2d(x, y, dtype_size(in_dtype), os, n_img, height, width, kh - 1 - ch, ch, device,
                       [&](const void *dx, void *dy, int64_t nr, int64_t w, int64_t lx, int64_t ly, int64_t above,
                           int64_t below, cudaStream_t s) {
                           return firgpu_fir2d_fixed(dx, dy, in_dtype, out_dtype, 1, nr, w, lx, ly, 0, 0, hq, kh, kw,
                                                     frac_bits, acc_bits, above, below, (void *)s);
                       });
}

int firgpu_fir2d_ideal_host(const void *x, void *y, int in_dtype, int out_dtype, int64_t n_img, int64_t height,
                            int64_t width, const double *h, int kh, int kw, int device)
{
    int rc = check_2d("firgpu_fir2d_ideal_host", x, y, in_dtype, n_img, height, width, width, width, h, kh, kw, 0, 0);
    if (rc) return rc;
    const int os = out_size_ideal(out_dtype);
    if (!os) return set_error(FIRGPU_E_DTYPE, "firgpu_fir2d_ideal_host: out_dtype must be F64 or F32, got %d", out_dtype);
    const int64_t ch = kh / 2;
    return pipeline_2d(x, y, dtype_size(in_dtype), os, n_img, height, width, kh - 1 - ch, ch, device,
                       [&](const void *dx, void *dy, int64_t nr, int64_t w, int64_t lx, int64_t ly, int64_t above,
                           int64_t below, cudaStream_t s) {
                           return firgpu_fir2d_ideal(dx, dy, in_dtype, out_dtype, 1, nr, w, lx, ly, 0, 0, h, kh, kw,
                                                     above, below, (void *)s);
                       });
}

int firgpu_fir1d_compare(const void *x, int in_dtype, int64_t rows, int64_t cols, int64_t ld_x, const int32_t *hq,
                         const double *h, int n_taps, int frac_bits, int acc_bits, double *metrics, void *stream)
{
    int rc = check_1d("firgpu_fir1d_compare", x, x, in_dtype, rows, cols, ld_x, ld_x, hq, n_taps, 1, 0, 0);
    if (rc) return rc;
    if ((rc = check_bits("firgpu_fir1d_compare", frac_bits, acc_bits))) return rc;
    if (!h || !metrics) return set_error(FIRGPU_E_NULL, "firgpu_fir1d_compare: NULL taps / metrics pointer");
    for (int i = 0; i < 7; ++i) metrics[i] = 0.0;
    if (rows == 0 || cols == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    if (convert_first(in_dtype, rows * cols) && n_taps <= 9) {         // int32 / float containers: uint8 pixels first
        uint8_t *tmp = nullptr, *body = nullptr;
        int64_t ld_t = 0;
        if ((rc = uint8_copy_1d(x, in_dtype, rows, cols, ld_x, 0, 0, (cudaStream_t)stream, &tmp, &body, &ld_t))) return rc;
        rc = launch_fir1d_compare(body, FIRGPU_U8, rows, cols, ld_t, hq, h, n_taps, frac_bits, acc_bits, metrics, (cudaStream_t)stream);
        cudaFreeAsync(tmp, (cudaStream_t)stream);
        return rc;
    }
    return launch_fir1d_compare(x, in_dtype, rows, cols, ld_x, hq, h, n_taps, frac_bits, acc_bits, metrics, (cudaStream_t)stream);
}

int firgpu_fir1d_compare_host(const void *x, int in_dtype, int64_t rows, int64_t cols, int64_t ld_x, const int32_t *hq,
                              const double *h, int n_taps, int frac_bits, int acc_bits, double *metrics, int device)
{
    int rc = check_1d("firgpu_fir1d_compare_host", x, x, in_dtype, rows, cols, ld_x, ld_x, hq, n_taps, 1, 0, 0);
    if (rc) return rc;
    if ((rc = check_bits("firgpu_fir1d_compare_host", frac_bits, acc_bits))) return rc;
    if (!h || !metrics) return set_error(FIRGPU_E_NULL, "firgpu_fir1d_compare_host: NULL taps / metrics pointer");
    for (int i = 0; i < 7; ++i) metrics[i] = 0.0;
    if (rows == 0 || cols == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    // Row pieces travel H2D on the slot streams while earlier pieces are being reduced; every piece adds into one
    // device accumulator (the kernels publish with atomics), read back once at the end.
    DeviceGuard guard;
    if ((rc = guard.enter(device))) return rc;
    int dev = 0;
    FIRGPU_CUDA(cudaGetDevice(&dev));
    HostCtx *c = ctx_for(dev);
    std::lock_guard<std::mutex> lock(c->mu);
    const int es = dtype_size(in_dtype);
    const int64_t pitch = round_up(cols * es, 16);
    int64_t per = std::max<int64_t>(1, (int64_t)(chunk_target() / (size_t)pitch));
    if (rows >= 2 * kSlots && per > (rows + 2 * kSlots - 1) / (2 * kSlots) && pitch * rows > (4 << 20))
        per = (rows + 2 * kSlots - 1) / (2 * kSlots);
    per = std::min(per, rows);
    const bool x_pinned = is_pinned(x);
    if ((rc = ctx_reserve(c, (size_t)(pitch * per + 256), 16, !x_pinned, false))) return rc;
    SlotDrain drain{c};
    double *d_sums = nullptr;
    if ((rc = scratch_alloc((void **)&d_sums, 8 * sizeof(double), c->stream[0]))) return rc;
    FIRGPU_CUDA(cudaMemsetAsync(d_sums, 0, 8 * sizeof(double), c->stream[0]));
    FIRGPU_CUDA(cudaStreamSynchronize(c->stream[0]));
    bool busy[kSlots] = {};
    int64_t k = 0;
    for (int64_t r0 = 0; r0 < rows && !rc; r0 += per, ++k) {
        const int s = (int)(k % kSlots);
        const int64_t nr = std::min(per, rows - r0);
        if (busy[s]) FIRGPU_CUDA(cudaStreamSynchronize(c->stream[s]));
        const char *src = (const char *)x + r0 * ld_x * es;
        size_t spitch = (size_t)ld_x * es;
        if (!x_pinned) {
            copy_rows_host(c->p_in[s], (size_t)cols * es, src, spitch, (size_t)cols * es, (size_t)nr);
            src = (const char *)c->p_in[s];
            spitch = (size_t)cols * es;
        }
        rc = copy_rows_async(c->d_in[s], (size_t)pitch, src, spitch, (size_t)cols * es, (size_t)nr, cudaMemcpyHostToDevice, c->stream[s]);
        if (!rc)
            rc = launch_fir1d_compare_into(c->d_in[s], in_dtype, nr, cols, pitch / es, hq, h, n_taps, frac_bits, acc_bits, d_sums, c->stream[s]);
        busy[s] = true;
    }
    double host[8] = {};
    if (!rc) {
        for (int s = 0; s < kSlots; ++s)
            if (busy[s]) FIRGPU_CUDA(cudaStreamSynchronize(c->stream[s]));
        FIRGPU_CUDA(cudaMemcpyAsync(host, d_sums, sizeof(host), cudaMemcpyDeviceToHost, c->stream[0]));
    }
    cudaFreeAsync(d_sums, c->stream[0]);
    FIRGPU_CUDA(cudaStreamSynchronize(c->stream[0]));
    if (rc) return rc;
    compare_unpack(host, metrics);
    return 0;
}

// ---- recorded launch sequences (CUDA graphs) ----------------------------------------------------------------------
// Between firgpu_record_begin and firgpu_record_end the device entry points called on `stream` from this thread are
// captured instead of executed; the resulting graph replays the whole sequence with one submission.  This is the
// batch entry point: the reference pipeline's 7 images x 8 tap sets, or the steps of a streaming loop, become one
// host call (a few microseconds) instead of one per launch.
struct firgpu_graph_s {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int device = 0;
    size_t nodes = 0;
};

int firgpu_record_begin(void *stream)
{
    if (!stream) return set_error(FIRGPU_E_NULL, "firgpu_record_begin: a recorded sequence needs an explicit (non-default) stream");
    FIRGPU_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
    return 0;
}

int firgpu_record_end(void *stream, firgpu_graph_t *out)
{
    if (!stream || !out) return set_error(FIRGPU_E_NULL, "firgpu_record_end: NULL");
    cudaGraph_t graph = nullptr;
    FIRGPU_CUDA(cudaStreamEndCapture((cudaStream_t)stream, &graph));
    firgpu_graph_s *g = new firgpu_graph_s();
    g->graph = graph;
    cudaGetDevice(&g->device);
    size_t n_nodes = 0;
    if (cudaGraphGetNodes(graph, nullptr, &n_nodes) == cudaSuccess && n_nodes) {     // kernels only: that is what a replay launches
        std::vector<cudaGraphNode_t> nodes(n_nodes);
        cudaGraphGetNodes(graph, nodes.data(), &n_nodes);
        for (size_t i = 0; i < n_nodes; ++i) {
            cudaGraphNodeType t;
            if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++g->nodes;
        }
    }
    cudaError_t e = cudaGraphInstantiate(&g->exec, graph, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(graph);
        delete g;
        return check_cuda(e, "cudaGraphInstantiate");
    }
    *out = g;
    return 0;
}

int firgpu_record_abort(void *stream)
{
    if (!stream) return 0;
    cudaGraph_t graph = nullptr;
    cudaStreamEndCapture((cudaStream_t)stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return 0;
}

int firgpu_graph_launch(firgpu_graph_t g, void *stream)
{
    if (!g || !g->exec) return set_error(FIRGPU_E_NULL, "firgpu_graph_launch: NULL graph");
    FIRGPU_CUDA(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
    tl_launches += (int64_t)g->nodes;
    return 0;
}

int firgpu_graph_nodes(firgpu_graph_t g, int64_t *nodes)
{
    if (!g || !nodes) return set_error(FIRGPU_E_NULL, "firgpu_graph_nodes: NULL");
    *nodes = (int64_t)g->nodes;
    return 0;
}

int firgpu_graph_destroy(firgpu_graph_t g)
{
    if (!g) return 0;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return 0;
}

// ---- extended entry points: alignment, clamp / saturation bounds, raw samples ----
namespace {
int check_ex(const char *fn, int n_taps, int centre, int in_lo, int in_hi, int out_lo, int out_hi)
{
    if (centre < 0 || centre >= n_taps) return set_error(FIRGPU_E_TAPS, "%s: centre=%d outside [0,%d)", fn, centre, n_taps);
    if (in_lo > in_hi || out_lo > out_hi) return set_error(FIRGPU_E_BITS, "%s: empty clamp range", fn);
    if (in_lo < -32768 || in_hi > 32767) return set_error(FIRGPU_E_BITS, "%s: sample range must fit int16", fn);
    return 0;
}
}  // namespace

int firgpu_fir1d_fixed_ex(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols, int64_t ld_x,
                          int64_t ld_y, const int32_t *hq, int n_taps, int dilation, int frac_bits, int acc_bits,
                          int64_t halo_left_valid, int64_t halo_right_valid, int centre, int in_lo, int in_hi, int out_lo,
                          int out_hi, void *stream)
{
    int rc = check_1d("firgpu_fir1d_fixed_ex", x, y, in_dtype, rows, cols, ld_x, ld_y, hq, n_taps, dilation,
                      halo_left_valid, halo_right_valid);
    if (rc) return rc;
    if ((rc = check_bits("firgpu_fir1d_fixed_ex", frac_bits, acc_bits))) return rc;
    if ((rc = check_ex("firgpu_fir1d_fixed_ex", n_taps, centre, in_lo, in_hi, out_lo, out_hi))) return rc;
    const int os = out_size_fixed(out_dtype);
    if (!os) return set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_fixed_ex: out_dtype must be U8 or I16, got %d", out_dtype);
    if ((out_dtype == FIRGPU_U8 && (out_lo < 0 || out_hi > 255)) || out_lo < -32768 || out_hi > 32767)
        return set_error(FIRGPU_E_BITS, "firgpu_fir1d_fixed_ex: output range [%d,%d] does not fit the output container", out_lo, out_hi);
    if (rows == 0 || cols == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    Fir1dArgs a{x, y, in_dtype, out_dtype, rows, cols, ld_x, ld_y, n_taps, dilation, frac_bits, acc_bits,
                halo_left_valid, halo_right_valid, (cudaStream_t)stream};
    a.centre = centre;
    a.in_lo = in_lo;
    a.in_hi = in_hi;
    a.out_lo = out_lo;
    a.out_hi = out_hi;
    rc = fir1d_fixed_ex_device(a, hq, false);
    return rc == 1 ? 0 : rc;
}

int firgpu_fir1d_ideal_ex(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols, int64_t ld_x,
                          int64_t ld_y, const double *h, int n_taps, int dilation, int64_t halo_left_valid,
                          int64_t halo_right_valid, int centre, int in_lo, int in_hi, int raw, void *stream)
{
    int rc = check_1d("firgpu_fir1d_ideal_ex", x, y, in_dtype, rows, cols, ld_x, ld_y, h, n_taps, dilation, halo_left_valid,
                      halo_right_valid);
    if (rc) return rc;
    if ((rc = check_ex("firgpu_fir1d_ideal_ex", n_taps, centre, in_lo, in_hi, 0, 0))) return rc;
    if (!out_size_ideal(out_dtype)) return set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_ideal_ex: out_dtype must be F64 or F32, got %d", out_dtype);
    if (rows == 0 || cols == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    Fir1dArgs a{x, y, in_dtype, out_dtype, rows, cols, ld_x, ld_y, n_taps, dilation, 0, 0,
                halo_left_valid, halo_right_valid, (cudaStream_t)stream};
    a.centre = centre;
    a.in_lo = in_lo;
    a.in_hi = in_hi;
    a.raw = raw ? 1 : 0;
    return launch_fir1d_ideal(a, h);
}

int firgpu_fir1d_fixed_ex_host(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols,
                               int64_t ld_x, int64_t ld_y, const int32_t *hq, int n_taps, int dilation, int frac_bits,
                               int acc_bits, int centre, int in_lo, int in_hi, int out_lo, int out_hi, int device)
{
    int rc = check_1d("firgpu_fir1d_fixed_ex_host", x, y, in_dtype, rows, cols, ld_x, ld_y, hq, n_taps, dilation, 0, 0);
    if (rc) return rc;
    if ((rc = check_ex("firgpu_fir1d_fixed_ex_host", n_taps, centre, in_lo, in_hi, out_lo, out_hi))) return rc;
    const int os = out_size_fixed(out_dtype);
    if (!os) return set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_fixed_ex_host: out_dtype must be U8 or I16, got %d", out_dtype);
    return pipeline_1d(x, y, dtype_size(in_dtype), os, rows, cols, ld_x, ld_y, (n_taps - 1 - centre) * (int64_t)dilation,
                       centre * (int64_t)dilation, device,
                       [&](const void *dx, void *dy, int64_t nr, int64_t nc, int64_t lx, int64_t ly, int64_t before,
                           int64_t after, cudaStream_t s) {
                           return firgpu_fir1d_fixed_ex(dx, dy, in_dtype, out_dtype, nr, nc, lx, ly, hq, n_taps, dilation,
                                                        frac_bits, acc_bits, before, after, centre, in_lo, in_hi, out_lo,
                                                        out_hi, (void *)s);
                       });
}

int firgpu_fir1d_ideal_ex_host(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols,
                               int64_t ld_x, int64_t ld_y, const double *h, int n_taps, int dilation, int centre,
                               int in_lo, int in_hi, int raw, int device)
{
    int rc = check_1d("firgpu_fir1d_ideal_ex_host", x, y, in_dtype, rows, cols, ld_x, ld_y, h, n_taps, dilation, 0, 0);
    if (rc) return rc;
    if ((rc = check_ex("firgpu_fir1d_ideal_ex_host", n_taps, centre, in_lo, in_hi, 0, 0))) return rc;
    const int os = out_size_ideal(out_dtype);
    if (!os) return set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_ideal_ex_host: out_dtype must be F64 or F32, got %d", out_dtype);
    return pipeline_1d(x, y, dtype_size(in_dtype), os, rows, cols, ld_x, ld_y, (n_taps - 1 - centre) * (int64_t)dilation,
                       centre * (int64_t)dilation, device,
                       [&](const void *dx, void *dy, int64_t nr, int64_t nc, int64_t lx, int64_t ly, int64_t before,
                           int64_t after, cudaStream_t s) {
                           return firgpu_fir1d_ideal_ex(dx, dy, in_dtype, out_dtype, nr, nc, lx, ly, h, n_taps, dilation,
                                                        before, after, centre, in_lo, in_hi, raw, (void *)s);
                       });
}

// ---- one host array, several devices --------------------------------------------------------------------------
// A single ctypes caller with one big NumPy array gets the whole box: the problem is partitioned like firgpu/shard.py
// does it for one process per GPU (rows split evenly; a single long row cut into contiguous segments whose taps' reach
// is read from the neighbouring host samples, cuts on 16384-scalar boundaries so I/Q pairs stay whole) and every
// device runs its own H2D -> kernel -> D2H pipeline on a host thread of its own.  No device talks to another.
int firgpu_fir1d_fixed_host_multi(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols, int64_t ld_x,
                                  int64_t ld_y, const int32_t *hq, int n_taps, int dilation, int frac_bits, int acc_bits,
                                  const int *devices, int n_devices)
{
    int rc = check_1d("firgpu_fir1d_fixed_host_multi", x, y, in_dtype, rows, cols, ld_x, ld_y, hq, n_taps, dilation, 0, 0);
    if (rc) return rc;
    if ((rc = check_bits("firgpu_fir1d_fixed_host_multi", frac_bits, acc_bits))) return rc;
    const int os = out_size_fixed(out_dtype), es = dtype_size(in_dtype);
    if (!os) return set_error(FIRGPU_E_DTYPE, "firgpu_fir1d_fixed_host_multi: out_dtype must be U8 or I16, got %d", out_dtype);
    if (!devices || n_devices < 1 || n_devices > 64) return set_error(FIRGPU_E_SHAPE, "firgpu_fir1d_fixed_host_multi: n_devices=%d", n_devices);
    if (rows == 0 || cols == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    struct Share {
        int device;
        int64_t r0, nr, c0, nc;
        int rc = 0, path = FIRGPU_PATH_NONE;
        int64_t launches = 0;
        char error[512] = "";
    };
    std::vector<Share> shares;
    const int64_t centre = n_taps / 2, reach_before = (n_taps - 1 - centre) * (int64_t)dilation, reach_after = centre * (int64_t)dilation;
    if (rows >= 2) {
        const int parts = (int)std::min<int64_t>(n_devices, rows);
        for (int d = 0; d < parts; ++d) {
            const int64_t r0 = rows * d / parts, r1 = rows * (d + 1) / parts;
            shares.push_back({devices[d], r0, r1 - r0, 0, cols});
        }
    } else {
        const int64_t unit = 16384, units = (cols + unit - 1) / unit;
        const int parts = (int)std::min<int64_t>(n_devices, units);
        for (int d = 0; d < parts; ++d) {
            const int64_t c0 = std::min(cols, units * d / parts * unit), c1 = std::min(cols, units * (d + 1) / parts * unit);
            if (c1 > c0) shares.push_back({devices[d], 0, 1, c0, c1 - c0});
        }
    }
    auto work = [&](Share &sh) {
        const char *xs = (const char *)x + (sh.r0 * ld_x + sh.c0) * es;
        char *ys = (char *)y + (sh.r0 * ld_y + sh.c0) * os;
        const int64_t before0 = tl_launches;
        sh.rc = pipeline_1d(xs, ys, es, os, sh.nr, sh.nc, ld_x, ld_y, reach_before, reach_after, sh.device,
                            [&](const void *dx, void *dy, int64_t nr, int64_t nc, int64_t lx, int64_t ly, int64_t before, int64_t after,
                                cudaStream_t s) {
                                return firgpu_fir1d_fixed(dx, dy, in_dtype, out_dtype, nr, nc, lx, ly, hq, n_taps, dilation, frac_bits,
                                                          acc_bits, before, after, (void *)s);
                            },
                            1, 0, rows == 1 ? sh.c0 : 0, rows == 1 ? cols - sh.c0 - sh.nc : 0);
        sh.path = tl_path;
        sh.launches = tl_launches - before0;
        if (sh.rc) snprintf(sh.error, sizeof(sh.error), "%s", tl_error);
    };
    std::vector<std::thread> pool;
    for (size_t i = 1; i < shares.size(); ++i) pool.emplace_back(work, std::ref(shares[i]));
    work(shares[0]);
    for (auto &t : pool) t.join();
    for (size_t i = 0; i < shares.size(); ++i) {
        if (i) tl_launches += shares[i].launches;           // (share 0 ran on this thread and counted itself)
        if (shares[i].path == FIRGPU_PATH_GENERIC || tl_path == FIRGPU_PATH_NONE) tl_path = shares[i].path;
        if (shares[i].rc && !rc) {
            rc = shares[i].rc;
            snprintf(tl_error, sizeof(tl_error), "device %d: %s", shares[i].device, shares[i].error);
        }
    }
    return rc;
}

// ---- complex taps on interleaved I/Q (fir1d_complex.cu explains the decomposition) -----------------------------
namespace {
// Real 2L-tap filters A (Re, centre 2c+1) and B (Im, centre 2c) of the complex taps (hq_re, hq_im).
void complex_tap_filters(const int32_t *hq_re, const int32_t *hq_im, int n_taps, std::vector<int32_t> &ga, std::vector<int32_t> &gb)
{
    ga.assign(2 * (size_t)n_taps, 0);
    gb.assign(2 * (size_t)n_taps, 0);
    for (int k = 0; k < n_taps; ++k) {
        ga[2 * k] = -hq_im[k];
        ga[2 * k + 1] = hq_re[k];
        gb[2 * k] = hq_re[k];
        gb[2 * k + 1] = hq_im[k];
    }
}

int complex_device(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                   const std::vector<int32_t> &ga, const std::vector<int32_t> &gb, int n_taps, int frac_bits, int acc_bits,
                   int64_t halo_left, int64_t halo_right, int in_lo, int in_hi, int out_lo, int out_hi, cudaStream_t stream)
{
    const int os = out_size_fixed(out_dtype), c = n_taps / 2;
    // scratch rows for the two real filters, 16-byte aligned pitch
    const int64_t ld_t = round_up(cols, 16);
    unsigned char *tmp = nullptr;
    int rc = scratch_alloc((void **)&tmp, (size_t)(2 * ld_t * rows * os), stream);
    if (rc) return rc;
    unsigned char *ta = tmp, *tb = tmp + (size_t)ld_t * rows * os;
    Fir1dArgs a{x, ta, in_dtype, out_dtype, rows, cols, ld_x, ld_t, 2 * n_taps, 1, frac_bits, acc_bits, halo_left, halo_right, stream};
    a.in_lo = in_lo;
    a.in_hi = in_hi;
    a.out_lo = out_lo;
    a.out_hi = out_hi;
    a.centre = 2 * c + 1;
    rc = fir1d_fixed_ex_device(a, ga.data(), false);
    if (rc == 1) {
        a.y = tb;
        a.centre = 2 * c;
        rc = fir1d_fixed_ex_device(a, gb.data(), false);
    }
    if (rc == 1) {
        rc = 0;
        if (rows == 1 || (ld_t == cols && ld_y == cols)) {           // one dense block (cols is even, so the parity carries over rows)
            rc = launch_select_even_odd(ta, tb, y, rows * cols, os, stream);
        } else {
            for (int64_t r = 0; r < rows && !rc; ++r)
                rc = launch_select_even_odd(ta + (size_t)r * ld_t * os, tb + (size_t)r * ld_t * os, (char *)y + (size_t)r * ld_y * os, cols, os, stream);
        }
    }
    cudaFreeAsync(tmp, stream);
    return rc;
}

int check_complex(const char *fn, const void *x, const void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols, int64_t ld_x,
                  int64_t ld_y, const int32_t *hq_re, const int32_t *hq_im, int n_taps, int frac_bits, int acc_bits, int64_t halo_left,
                  int64_t halo_right, int in_lo, int in_hi, int out_lo, int out_hi)
{
    int rc = check_1d(fn, x, y, in_dtype, rows, cols, ld_x, ld_y, hq_re, n_taps, 1, halo_left, halo_right);
    if (rc) return rc;
    if (!hq_im) return set_error(FIRGPU_E_NULL, "%s: imaginary taps pointer is NULL", fn);
    if ((rc = check_bits(fn, frac_bits, acc_bits))) return rc;
    if ((rc = check_ex(fn, n_taps, n_taps / 2, in_lo, in_hi, out_lo, out_hi))) return rc;
    if (2 * n_taps > FIRGPU_MAX_TAPS) return set_error(FIRGPU_E_TAPS, "%s: n_taps=%d exceeds %d complex taps", fn, n_taps, FIRGPU_MAX_TAPS / 2);
    if (!out_size_fixed(out_dtype)) return set_error(FIRGPU_E_DTYPE, "%s: out_dtype must be U8 or I16, got %d", fn, out_dtype);
    if ((out_dtype == FIRGPU_U8 && (out_lo < 0 || out_hi > 255)) || out_lo < -32768 || out_hi > 32767)
        return set_error(FIRGPU_E_BITS, "%s: output range [%d,%d] does not fit the output container", fn, out_lo, out_hi);
    if ((cols | halo_left | halo_right) & 1) return set_error(FIRGPU_E_SHAPE, "%s: cols and halos count interleaved scalars and must be even", fn);
    for (int k = 0; k < n_taps; ++k)
        if (hq_im[k] == INT32_MIN) return set_error(FIRGPU_E_TAPS, "%s: imaginary tap %d cannot be negated in int32", fn, k);
    return 0;
}
}  // namespace

int firgpu_fir1d_fixed_complex(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols, int64_t ld_x,
                               int64_t ld_y, const int32_t *hq_re, const int32_t *hq_im, int n_taps, int frac_bits, int acc_bits,
                               int64_t halo_left_valid, int64_t halo_right_valid, int in_lo, int in_hi, int out_lo, int out_hi,
                               void *stream)
{
    int rc = check_complex("firgpu_fir1d_fixed_complex", x, y, in_dtype, out_dtype, rows, cols, ld_x, ld_y, hq_re, hq_im, n_taps, frac_bits,
                           acc_bits, halo_left_valid, halo_right_valid, in_lo, in_hi, out_lo, out_hi);
    if (rc) return rc;
    if (rows == 0 || cols == 0) {
        tl_path = FIRGPU_PATH_EMPTY;
        return 0;
    }
    std::vector<int32_t> ga, gb;
    complex_tap_filters(hq_re, hq_im, n_taps, ga, gb);
    return complex_device(x, y, in_dtype, out_dtype, rows, cols, ld_x, ld_y, ga, gb, n_taps, frac_bits, acc_bits, halo_left_valid,
                          halo_right_valid, in_lo, in_hi, out_lo, out_hi, (cudaStream_t)stream);
}

int firgpu_fir1d_fixed_complex_host(const void *x, void *y, int in_dtype, int out_dtype, int64_t rows, int64_t cols, int64_t ld_x,
                                    int64_t ld_y, const int32_t *hq_re, const int32_t *hq_im, int n_taps, int frac_bits,
                                    int acc_bits, int in_lo, int in_hi, int out_lo, int out_hi, int device)
{
    int rc = check_complex("firgpu_fir1d_fixed_complex_host", x, y, in_dtype, out_dtype, rows, cols, ld_x, ld_y, hq_re, hq_im, n_taps,
                           frac_bits, acc_bits, 0, 0, in_lo, in_hi, out_lo, out_hi);
    if (rc) return rc;
    std::vector<int32_t> ga, gb;
    complex_tap_filters(hq_re, hq_im, n_taps, ga, gb);
    const int64_t c = n_taps / 2;
    // reach in interleaved scalars, rounded up to whole complex samples: filter B reads 2L-1-2c scalars back, A 2c+1 ahead
    return pipeline_1d(x, y, dtype_size(in_dtype), out_size_fixed(out_dtype), rows, cols, ld_x, ld_y, 2 * (n_taps - c), 2 * c + 2, device,
                       [&](const void *dx, void *dy, int64_t nr, int64_t nc, int64_t lx, int64_t ly, int64_t before, int64_t after,
                           cudaStream_t s) {
                           return complex_device(dx, dy, in_dtype, out_dtype, nr, nc, lx, ly, ga, gb, n_taps, frac_bits, acc_bits, before,
                                                 after, in_lo, in_hi, out_lo, out_hi, s);
                       });
}

int firgpu_fir1d_fixed_plan(const int32_t *hq, int n_taps, int frac_bits, int acc_bits, int *path, int *post, int *mac)
{
    if (!hq || !path || !post || !mac) return set_error(FIRGPU_E_NULL, "firgpu_fir1d_fixed_plan: NULL");
    if (n_taps < 1 || n_taps > FIRGPU_MAX_TAPS) return set_error(FIRGPU_E_TAPS, "firgpu_fir1d_fixed_plan: n_taps=%d outside [1,%d]", n_taps, FIRGPU_MAX_TAPS);
    int rc = check_bits("firgpu_fir1d_fixed_plan", frac_bits, acc_bits);
    if (rc) return rc;
    return plan_fir1d_fixed(hq, n_taps, frac_bits, acc_bits, path, post, mac);
}

int firgpu_get_device(int *device)
{
    if (!device) return set_error(FIRGPU_E_NULL, "firgpu_get_device: NULL");
    FIRGPU_CUDA(cudaGetDevice(device));
    return 0;
}

int firgpu_set_device(int device)
{
    FIRGPU_CUDA(cudaSetDevice(device));
    return 0;
}

int firgpu_pinned_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return set_error(FIRGPU_E_NULL, "firgpu_pinned_alloc: NULL");
    FIRGPU_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}

int firgpu_pinned_free(void *ptr)
{
    if (ptr) FIRGPU_CUDA(cudaFreeHost(ptr));
    return 0;
}

int firgpu_release_cache(void)
{
    std::lock_guard<std::mutex> lock(g_ctx_mu);
    int prev = 0;
    cudaGetDevice(&prev);
    for (auto &kv : g_ctx) {
        HostCtx *c = kv.second;
        std::lock_guard<std::mutex> l2(c->mu);
        cudaSetDevice(kv.first);
        for (int s = 0; s < kSlots; ++s) {
            if (c->d_in[s]) cudaFree(c->d_in[s]);
            if (c->d_out[s]) cudaFree(c->d_out[s]);
            if (c->p_in[s]) cudaFreeHost(c->p_in[s]);
            if (c->p_out[s]) cudaFreeHost(c->p_out[s]);
            c->d_in[s] = c->d_out[s] = c->p_in[s] = c->p_out[s] = nullptr;
            if (c->ready) cudaStreamDestroy(c->stream[s]);
        }
        c->d_in_cap = c->d_out_cap = c->p_in_cap = c->p_out_cap = 0;
        c->ready = false;
    }
    {
        std::lock_guard<std::mutex> l3(g_pool_mu);
        for (auto &kv : g_pools) cudaMemPoolTrimTo(kv.second, 0);
    }
    cudaSetDevice(prev);
    cudaGetLastError();
    return 0;
}

}  // extern "C"
