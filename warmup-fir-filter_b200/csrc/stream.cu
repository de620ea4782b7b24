// Device-resident streaming filters (firgpu_stream_*): the causal models with the (taps-1)-sample history carried
// between blocks in HBM — what reg_ is in the reference's C++ class (fir_1d/model/cpp/fir_1d_ref.cpp:16-33,
// Fir1D::process_sample shifts reg_ and sums h[k] * reg_[k], i.e. y[n] = sum_k h[k] x[n-k]).
//
// A block is filtered where it lies (device memory of the caller).  Fast kernels (staged fixed-point kernel, register
// kernel of the raw float64 model) read the history themselves: the scalars in front of column 0 come from the history
// buffer, and the CTA of the last tile writes the block's tail to the OTHER of two history buffers (the CTA of the
// first tile may still be reading the current one) — ONE launch per block.  Everything else: one pass of the ordinary
// causal kernels with a zero history computes every output, then one small kernel (a) rewrites the first
// (taps-1)*dilation outputs, whose windows reach into the previous block, from the saved history and (b) replaces the
// history by the block's tail.  Nothing is concatenated or copied; blocks no longer than the history are served by
// the small kernel alone.
// Host blocks travel through a chunked H2D -> filter -> D2H pipeline whose filter steps are chained by events
// (the history makes consecutive chunks dependent; their copies still overlap).
#include "common.cuh"

#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

namespace firgpu {
namespace {

constexpr int kHeadThreads = 256;
constexpr int kHostSlots = 3;
constexpr int64_t kHostChunk = 8LL << 20;          // samples per host pipeline chunk

struct StreamState {
    bool fixed = true;
    int device = 0, in_dtype = 0, out_dtype = 0, n_taps = 0, dilation = 1;
    int frac_bits = 0, acc_bits = 0, in_lo = 0, in_hi = 255, out_lo = 0, out_hi = 255, raw = 0;
    std::vector<int32_t> hq;
    std::vector<double> h;
    int64_t hist = 0;                              // history scalars: (n_taps - 1) * dilation
    void *d_state = nullptr, *d_taps = nullptr;    // d_state: two history buffers of `hist` scalars; `cur` is the live one
    int cur = 0;
    size_t state_stride = 0;                       // bytes between the two buffers
    void *state(int which) const { return (char *)d_state + (size_t)which * state_stride; }
    std::mutex mu;
    // host pipeline (created on first use)
    cudaStream_t hs[kHostSlots] = {};
    cudaEvent_t filtered[kHostSlots] = {};
    void *d_x[kHostSlots] = {}, *d_y[kHostSlots] = {}, *p_x[kHostSlots] = {}, *p_y[kHostSlots] = {};
    int64_t cap = 0;
    bool host_ready = false;
};

template <typename InT> __device__ __forceinline__ double ideal_sample(InT v, int lo, int hi, int raw)
{
    return raw ? (double)v : (double)prep_sample_range(v, lo, hi);
}

// Outputs 0 .. min(n, hist)-1 of the block from the saved history + the block's first samples (same arithmetic as
// the generic kernels: int64 sum + post_accumulate, or unfused float64 / float32 fmaf in ascending taps), then the
// new history = the last `hist` scalars of history ++ block.  One CTA; the old history is read into shared memory
// first, so the update is safe for blocks shorter than the history.
template <typename InT, typename OutT, bool FIXED>
__global__ void __launch_bounds__(kHeadThreads)
stream_head_kernel(const InT *__restrict__ x, OutT *__restrict__ y, InT *__restrict__ state, long long n, int hist,
                   const void *__restrict__ taps, int n_taps, int dilation, int frac_bits, int acc_bits, int in_lo, int in_hi,
                   int out_lo, int out_hi, int raw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    InT *old = reinterpret_cast<InT *>(smem_raw);
    for (int j = threadIdx.x; j < hist; j += kHeadThreads) old[j] = state[j];
    __syncthreads();
    const long long m = n < hist ? n : hist;
    for (long long i = threadIdx.x; i < m; i += kHeadThreads) {
        if constexpr (FIXED) {
            const int32_t *hq = static_cast<const int32_t *>(taps);
            long long acc = 0;
            for (int k = 0; k < n_taps; ++k) {
                const long long src = i - (long long)k * dilation;
                const InT v = src >= 0 ? x[src] : old[hist + src];
                acc += (long long)prep_sample_range(v, in_lo, in_hi) * (long long)hq[k];
            }
            y[i] = (OutT)post_accumulate(acc, frac_bits, acc_bits, out_lo, out_hi);
        } else {
            const double *h = static_cast<const double *>(taps);
            OutT acc = (OutT)0;
            for (int k = 0; k < n_taps; ++k) {
                const long long src = i - (long long)k * dilation;
                const double s = ideal_sample(src >= 0 ? x[src] : old[hist + src], in_lo, in_hi, raw);
                if constexpr (sizeof(OutT) == 8) acc = __dadd_rn(acc, __dmul_rn(h[k], s));
                else acc = fmaf((float)h[k], (float)s, acc);
            }
            y[i] = acc;
        }
    }
    for (int j = threadIdx.x; j < hist; j += kHeadThreads) {
        const long long src = n - hist + j;
        state[j] = src >= 0 ? x[src] : old[j + n];
    }
}

template <typename InT, typename OutT, bool FIXED>
int launch_head(StreamState *s, const void *x, void *y, int64_t n, cudaStream_t stream)
{
    if (s->hist == 0) return 0;
    stream_head_kernel<InT, OutT, FIXED><<<1, kHeadThreads, (size_t)s->hist * sizeof(InT), stream>>>(
        (const InT *)x, (OutT *)y, (InT *)s->state(s->cur), (long long)n, (int)s->hist, s->d_taps, s->n_taps, s->dilation, s->frac_bits,
        s->acc_bits, s->in_lo, s->in_hi, s->out_lo, s->out_hi, s->raw);
    FIRGPU_CUDA(cudaGetLastError());
    note_launch(FIRGPU_PATH_GENERIC);
    return 0;
}

template <typename InT> int launch_head_in(StreamState *s, const void *x, void *y, int64_t n, cudaStream_t stream)
{
    if (s->fixed)
        return s->out_dtype == FIRGPU_U8 ? launch_head<InT, uint8_t, true>(s, x, y, n, stream) : launch_head<InT, int16_t, true>(s, x, y, n, stream);
    return s->out_dtype == FIRGPU_F64 ? launch_head<InT, double, false>(s, x, y, n, stream) : launch_head<InT, float, false>(s, x, y, n, stream);
}

int launch_head_any(StreamState *s, const void *x, void *y, int64_t n, cudaStream_t stream)
{
    switch (s->in_dtype) {
    case FIRGPU_U8: return launch_head_in<uint8_t>(s, x, y, n, stream);
    case FIRGPU_I16: return launch_head_in<int16_t>(s, x, y, n, stream);
    case FIRGPU_I32: return launch_head_in<int32_t>(s, x, y, n, stream);
    case FIRGPU_F32: return launch_head_in<float>(s, x, y, n, stream);
    case FIRGPU_F64: return launch_head_in<double>(s, x, y, n, stream);
    default: return set_error(FIRGPU_E_DTYPE, "firgpu_stream_process: unsupported in_dtype %d", s->in_dtype);
    }
}

int out_size(const StreamState *s) { return dtype_size(s->out_dtype); }

struct DeviceScope {
    int prev = -1;
    bool switched = false;
    int enter(int device)
    {
        FIRGPU_CUDA(cudaGetDevice(&prev));
        if (prev != device) {
            FIRGPU_CUDA(cudaSetDevice(device));
            switched = true;
        }
        return 0;
    }
    ~DeviceScope()
    {
        if (switched) cudaSetDevice(prev);
    }
};

// One block that already lies in device memory (stream-ordered, asynchronous).
int process_device(StreamState *s, const void *x, void *y, int64_t n, cudaStream_t stream)
{
    int rc = 0;
    if (n > s->hist) {
        // one launch: a fast kernel that reads the history buffer and leaves the next one behind.  Not while the stream
        // is being recorded: which of the two buffers is live is host state, and a replayed sequence must find the
        // history where it left it — the two-launch form below keeps it in one buffer.
        cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &capturing) != cudaSuccess) {
            cudaGetLastError();
            capturing = cudaStreamCaptureStatusNone;
        }
        if (s->hist > 0 && capturing == cudaStreamCaptureStatusNone) {
            Fir1dArgs a{x, y, s->in_dtype, s->out_dtype, 1, n, n, n, s->n_taps, s->dilation, s->frac_bits, s->acc_bits,
                        s->hist, 0, stream};
            a.centre = 0;
            a.in_lo = s->in_lo;
            a.in_hi = s->in_hi;
            a.out_lo = s->out_lo;
            a.out_hi = s->out_hi;
            a.raw = s->raw;
            a.hist = s->state(s->cur);
            a.hist_out = s->state(s->cur ^ 1);
            rc = s->fixed ? fir1d_fixed_ex_device(a, s->hq.data(), true) : try_launch_fir1d_ideal_fast(a, s->h.data());
            if (rc == 1) {
                s->cur ^= 1;
                return 0;
            }
            if (rc != 0) return rc;
        }
        // no fast kernel for these parameters: every output with a zero history; the first `hist` of them are rewritten below
        if (s->fixed)
            rc = firgpu_fir1d_fixed_ex(x, y, s->in_dtype, s->out_dtype, 1, n, n, n, s->hq.data(), s->n_taps, s->dilation, s->frac_bits,
                                       s->acc_bits, 0, 0, 0, s->in_lo, s->in_hi, s->out_lo, s->out_hi, (void *)stream);
        else
            rc = firgpu_fir1d_ideal_ex(x, y, s->in_dtype, s->out_dtype, 1, n, n, n, s->h.data(), s->n_taps, s->dilation, 0, 0, 0,
                                       s->in_lo, s->in_hi, s->raw, (void *)stream);
        if (rc) return rc;
    }
    return launch_head_any(s, x, y, n, stream);
}

int host_reserve(StreamState *s, int64_t samples, bool pin_x, bool pin_y)
{
    if (!s->host_ready) {
        for (int i = 0; i < kHostSlots; ++i) {
            FIRGPU_CUDA(cudaStreamCreateWithFlags(&s->hs[i], cudaStreamNonBlocking));
            FIRGPU_CUDA(cudaEventCreateWithFlags(&s->filtered[i], cudaEventDisableTiming));
        }
        s->host_ready = true;
    }
    const int es = dtype_size(s->in_dtype), os = out_size(s);
    if (samples > s->cap) {
        for (int i = 0; i < kHostSlots; ++i) {
            if (s->d_x[i]) cudaFree(s->d_x[i]);
            if (s->d_y[i]) cudaFree(s->d_y[i]);
            if (s->p_x[i]) cudaFreeHost(s->p_x[i]);
            if (s->p_y[i]) cudaFreeHost(s->p_y[i]);
            s->d_x[i] = s->d_y[i] = s->p_x[i] = s->p_y[i] = nullptr;
        }
        s->cap = 0;
        for (int i = 0; i < kHostSlots; ++i) {
            FIRGPU_CUDA(cudaMalloc(&s->d_x[i], (size_t)samples * es));
            FIRGPU_CUDA(cudaMalloc(&s->d_y[i], (size_t)samples * os));
        }
        s->cap = samples;
    }
    for (int i = 0; i < kHostSlots; ++i) {
        if (pin_x && !s->p_x[i]) FIRGPU_CUDA(cudaHostAlloc(&s->p_x[i], (size_t)s->cap * es, cudaHostAllocDefault));
        if (pin_y && !s->p_y[i]) FIRGPU_CUDA(cudaHostAlloc(&s->p_y[i], (size_t)s->cap * os, cudaHostAllocDefault));
    }
    return 0;
}

bool host_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

void drain(StreamState *s)
{
    for (int i = 0; i < kHostSlots; ++i)
        if (s->hs[i]) cudaStreamSynchronize(s->hs[i]);
}

int create_common(firgpu_stream_t *out, StreamState *s, const void *taps, size_t tap_bytes)
{
    DeviceScope scope;
    int rc = scope.enter(s->device);
    if (rc) {
        delete s;
        return rc;
    }
    s->state_stride = ((size_t)std::max<int64_t>(s->hist, 1) * dtype_size(s->in_dtype) + 255) / 256 * 256;
    const size_t state_bytes = 2 * s->state_stride;
    cudaError_t e = cudaMalloc(&s->d_state, state_bytes);
    if (e == cudaSuccess) e = cudaMemset(s->d_state, 0, state_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_taps, tap_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(s->d_taps, taps, tap_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (s->d_state) cudaFree(s->d_state);
        if (s->d_taps) cudaFree(s->d_taps);
        delete s;
        return check_cuda(e, "firgpu_stream_create");
    }
    *out = reinterpret_cast<firgpu_stream_t>(s);
    return 0;
}

int check_create(const char *fn, firgpu_stream_t *out, int in_dtype, const void *taps, int n_taps, int dilation, int in_lo, int in_hi)
{
    if (!out || !taps) return set_error(FIRGPU_E_NULL, "%s: NULL pointer", fn);
    if (in_dtype < FIRGPU_U8 || in_dtype > FIRGPU_F64) return set_error(FIRGPU_E_DTYPE, "%s: unsupported in_dtype %d", fn, in_dtype);
    if (n_taps < 1 || n_taps > FIRGPU_MAX_TAPS) return set_error(FIRGPU_E_TAPS, "%s: n_taps=%d outside [1,%d]", fn, n_taps, FIRGPU_MAX_TAPS);
    if (dilation < 1) return set_error(FIRGPU_E_DILATION, "%s: dilation=%d must be >= 1", fn, dilation);
    if ((int64_t)(n_taps - 1) * dilation * dtype_size(in_dtype) > 48 * 1024)
        return set_error(FIRGPU_E_TAPS, "%s: a history of (n_taps-1)*dilation = %lld samples does not fit the head kernel's 48 KiB", fn,
                         (long long)(n_taps - 1) * dilation);
    if (in_lo > in_hi || in_lo < -32768 || in_hi > 32767) return set_error(FIRGPU_E_BITS, "%s: sample range [%d,%d]", fn, in_lo, in_hi);
    return 0;
}

}  // namespace
}  // namespace firgpu

using namespace firgpu;

extern "C" {

int firgpu_stream_create_fixed(firgpu_stream_t *out, int in_dtype, int out_dtype, const int32_t *hq, int n_taps, int dilation,
                               int frac_bits, int acc_bits, int in_lo, int in_hi, int out_lo, int out_hi, int device)
{
    int rc = check_create("firgpu_stream_create_fixed", out, in_dtype, hq, n_taps, dilation, in_lo, in_hi);
    if (rc) return rc;
    if (out_dtype != FIRGPU_U8 && out_dtype != FIRGPU_I16)
        return set_error(FIRGPU_E_DTYPE, "firgpu_stream_create_fixed: out_dtype must be U8 or I16, got %d", out_dtype);
    if (frac_bits < 1 || acc_bits < 1) return set_error(FIRGPU_E_BITS, "firgpu_stream_create_fixed: frac_bits=%d and acc_bits=%d must be >= 1", frac_bits, acc_bits);
    if (out_lo > out_hi || (out_dtype == FIRGPU_U8 && (out_lo < 0 || out_hi > 255)) || out_lo < -32768 || out_hi > 32767)
        return set_error(FIRGPU_E_BITS, "firgpu_stream_create_fixed: output range [%d,%d] does not fit the output container", out_lo, out_hi);
    if (device < 0) FIRGPU_CUDA(cudaGetDevice(&device));
    StreamState *s = new (std::nothrow) StreamState();
    if (!s) return set_error(FIRGPU_E_NULL, "firgpu_stream_create_fixed: out of host memory");
    s->fixed = true;
    s->device = device;
    s->in_dtype = in_dtype;
    s->out_dtype = out_dtype;
    s->n_taps = n_taps;
    s->dilation = dilation;
    s->frac_bits = frac_bits;
    s->acc_bits = acc_bits;
    s->in_lo = in_lo;
    s->in_hi = in_hi;
    s->out_lo = out_lo;
    s->out_hi = out_hi;
    s->hq.assign(hq, hq + n_taps);
    s->hist = (int64_t)(n_taps - 1) * dilation;
    return create_common(out, s, hq, sizeof(int32_t) * (size_t)n_taps);
}

int firgpu_stream_create_ideal(firgpu_stream_t *out, int in_dtype, int out_dtype, const double *h, int n_taps, int dilation,
                               int in_lo, int in_hi, int raw, int device)
{
    int rc = check_create("firgpu_stream_create_ideal", out, in_dtype, h, n_taps, dilation, in_lo, in_hi);
    if (rc) return rc;
    if (out_dtype != FIRGPU_F64 && out_dtype != FIRGPU_F32)
        return set_error(FIRGPU_E_DTYPE, "firgpu_stream_create_ideal: out_dtype must be F64 or F32, got %d", out_dtype);
    if (device < 0) FIRGPU_CUDA(cudaGetDevice(&device));
    StreamState *s = new (std::nothrow) StreamState();
    if (!s) return set_error(FIRGPU_E_NULL, "firgpu_stream_create_ideal: out of host memory");
    s->fixed = false;
    s->device = device;
    s->in_dtype = in_dtype;
    s->out_dtype = out_dtype;
    s->n_taps = n_taps;
    s->dilation = dilation;
    s->in_lo = in_lo;
    s->in_hi = in_hi;
    s->raw = raw ? 1 : 0;
    s->h.assign(h, h + n_taps);
    s->hist = (int64_t)(n_taps - 1) * dilation;
    return create_common(out, s, h, sizeof(double) * (size_t)n_taps);
}

int firgpu_stream_process(firgpu_stream_t handle, const void *x, void *y, int64_t n, void *stream)
{
    StreamState *s = reinterpret_cast<StreamState *>(handle);
    if (!s) return set_error(FIRGPU_E_NULL, "firgpu_stream_process: NULL handle");
    if (n < 0) return set_error(FIRGPU_E_SHAPE, "firgpu_stream_process: n=%lld", (long long)n);
    if (n == 0) return 0;
    if (!x || !y) return set_error(FIRGPU_E_NULL, "firgpu_stream_process: NULL data pointer");
    std::lock_guard<std::mutex> lock(s->mu);
    DeviceScope scope;
    int rc = scope.enter(s->device);
    if (rc) return rc;
    return process_device(s, x, y, n, (cudaStream_t)stream);
}

int firgpu_stream_process_host(firgpu_stream_t handle, const void *x, void *y, int64_t n)
{
    StreamState *s = reinterpret_cast<StreamState *>(handle);
    if (!s) return set_error(FIRGPU_E_NULL, "firgpu_stream_process_host: NULL handle");
    if (n < 0) return set_error(FIRGPU_E_SHAPE, "firgpu_stream_process_host: n=%lld", (long long)n);
    if (n == 0) return 0;
    if (!x || !y) return set_error(FIRGPU_E_NULL, "firgpu_stream_process_host: NULL data pointer");
    std::lock_guard<std::mutex> lock(s->mu);
    DeviceScope scope;
    int rc = scope.enter(s->device);
    if (rc) return rc;
    const int es = dtype_size(s->in_dtype), os = out_size(s);
    const bool x_pinned = host_pinned(x), y_pinned = host_pinned(y);
    const int64_t chunk = std::min<int64_t>(n, kHostChunk);
    if ((rc = host_reserve(s, chunk, !x_pinned, !y_pinned))) return rc;

    struct Pending {
        bool busy = false;
        int64_t c0 = 0, nc = 0;
    } pending[kHostSlots];
    auto finalize = [&](int i) -> int {
        if (!pending[i].busy) return 0;
        pending[i].busy = false;
        FIRGPU_CUDA(cudaStreamSynchronize(s->hs[i]));
        if (!y_pinned) memcpy((char *)y + pending[i].c0 * os, s->p_y[i], (size_t)pending[i].nc * os);
        return 0;
    };
    int64_t k = 0;
    int prev = -1;
    for (int64_t c0 = 0; c0 < n && !rc; c0 += chunk, ++k) {
        const int i = (int)(k % kHostSlots);
        const int64_t nc = std::min(chunk, n - c0);
        if ((rc = finalize(i))) break;
        const void *src = (const char *)x + c0 * es;
        if (!x_pinned) {
            memcpy(s->p_x[i], src, (size_t)nc * es);
            src = s->p_x[i];
        }
        cudaError_t e = cudaMemcpyAsync(s->d_x[i], src, (size_t)nc * es, cudaMemcpyHostToDevice, s->hs[i]);
        // the filter step of this chunk needs the history left by the previous one
        if (e == cudaSuccess && prev >= 0) e = cudaStreamWaitEvent(s->hs[i], s->filtered[prev], 0);
        if (e != cudaSuccess) {
            rc = check_cuda(e, "firgpu_stream_process_host: copy in");
            break;
        }
        if ((rc = process_device(s, s->d_x[i], s->d_y[i], nc, s->hs[i]))) break;
        e = cudaEventRecord(s->filtered[i], s->hs[i]);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(y_pinned ? (void *)((char *)y + c0 * os) : s->p_y[i], s->d_y[i], (size_t)nc * os, cudaMemcpyDeviceToHost, s->hs[i]);
        if (e != cudaSuccess) {
            rc = check_cuda(e, "firgpu_stream_process_host: copy out");
            break;
        }
        pending[i].busy = true;
        pending[i].c0 = c0;
        pending[i].nc = nc;
        prev = i;
    }
    if (rc) {
        drain(s);            // nothing of ours may still read or write the caller's buffers after we return
        return rc;
    }
    for (int j = 0; j < kHostSlots; ++j) {
        const int i = (int)((k + j) % kHostSlots);      // oldest first
        if ((rc = finalize(i))) {
            drain(s);
            return rc;
        }
    }
    return 0;
}

int firgpu_stream_reset(firgpu_stream_t handle, void *stream)
{
    StreamState *s = reinterpret_cast<StreamState *>(handle);
    if (!s) return set_error(FIRGPU_E_NULL, "firgpu_stream_reset: NULL handle");
    std::lock_guard<std::mutex> lock(s->mu);
    DeviceScope scope;
    int rc = scope.enter(s->device);
    if (rc) return rc;
    if (s->hist) FIRGPU_CUDA(cudaMemsetAsync(s->state(s->cur), 0, (size_t)s->hist * dtype_size(s->in_dtype), (cudaStream_t)stream));
    return 0;
}

int firgpu_stream_history(firgpu_stream_t handle, int64_t *scalars)
{
    StreamState *s = reinterpret_cast<StreamState *>(handle);
    if (!s || !scalars) return set_error(FIRGPU_E_NULL, "firgpu_stream_history: NULL");
    *scalars = s->hist;
    return 0;
}

int firgpu_stream_destroy(firgpu_stream_t handle)
{
    StreamState *s = reinterpret_cast<StreamState *>(handle);
    if (!s) return 0;
    {
        std::lock_guard<std::mutex> lock(s->mu);
        DeviceScope scope;
        if (scope.enter(s->device) == 0) {
            drain(s);
            for (int i = 0; i < kHostSlots; ++i) {
                if (s->d_x[i]) cudaFree(s->d_x[i]);
                if (s->d_y[i]) cudaFree(s->d_y[i]);
                if (s->p_x[i]) cudaFreeHost(s->p_x[i]);
                if (s->p_y[i]) cudaFreeHost(s->p_y[i]);
                if (s->hs[i]) cudaStreamDestroy(s->hs[i]);
                if (s->filtered[i]) cudaEventDestroy(s->filtered[i]);
            }
            if (s->d_state) cudaFree(s->d_state);
            if (s->d_taps) cudaFree(s->d_taps);
        }
        cudaGetLastError();
    }
    delete s;
    return 0;
}

}  // extern "C"
