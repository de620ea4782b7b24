/*
 * firgpu — C ABI of the B200-native fixed-point FIR path.
 *
 * This header is the drop-in boundary.  The reference (young-0320/
 * warmup-fir-filter) has no FFI layer of its own: its hot path is a set of
 * Python functions.  Each entry point below names the reference interface it
 * stands behind (paths relative to the reference root); INTEGRATION.md shows
 * the ctypes stub a reference maintainer would add at those lines.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in signatures
 *     (streams travel as void* = cudaStream_t, NULL = the legacy default
 *     stream);
 *   - every function returns 0 on success, a negative FIRGPU_E_* code for a
 *     rejected argument, or a positive cudaError_t; firgpu_last_error() gives
 *     the thread-local message;
 *   - "device" entry points take DEVICE pointers and are stream-ordered and
 *     asynchronous; "_host" entry points take HOST pointers, run the
 *     host->device copy, the kernels and the device->host copy as a chunked
 *     pipeline, and return when the result is in y;
 *   - integer taps: the caller quantises real coefficients on the host with
 *     the reference's own NumPy calls (np.rint, np.clip, astype —
 *     fir_1d/model/python/fir_1d_fixed_ref.py:79-81) and passes int32 values;
 *     that float64 step is part of the bit-exact contract and is deliberately
 *     not re-implemented in C;
 *   - leading dimensions (ld_*, img_stride_*) are in ELEMENTS of the array's
 *     own dtype;
 *   - there is no CPU fallback anywhere behind this ABI: without a CUDA
 *     device every compute entry point fails with a cudaError_t.
 *
 * Sample containers (in_dtype): FIRGPU_U8 is the reference's native pixel
 * type; FIRGPU_I16/I32 are integer containers whose values the kernel clamps
 * to [0,255] exactly as the reference clamps any integer input
 * (fir_1d/model/python/fir_1d_ref.py:40-41); FIRGPU_F32/F64 are rounded with
 * floor(x + 0.5) in float64 first (fir_1d_ref.py:35-38).  Fixed-point outputs
 * (out_dtype) are FIRGPU_U8 or the same values in an FIRGPU_I16 container.
 */
#ifndef FIRGPU_H
#define FIRGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FIRGPU_VERSION 200          /* 0.2.0 */
#define FIRGPU_MAX_TAPS 1024        /* per filter; 2-D: kh*kw */

enum firgpu_dtype { FIRGPU_U8 = 0, FIRGPU_I16 = 1, FIRGPU_I32 = 2, FIRGPU_F32 = 3, FIRGPU_F64 = 4 };

enum firgpu_status {
    FIRGPU_OK = 0,
    FIRGPU_E_DTYPE = -1,      /* unsupported in/out container */
    FIRGPU_E_SHAPE = -2,      /* negative size, ld < cols, ... */
    FIRGPU_E_TAPS = -3,       /* n_taps outside [1, FIRGPU_MAX_TAPS] */
    FIRGPU_E_BITS = -4,       /* frac_bits < 1 or acc_bits < 1 */
    FIRGPU_E_DILATION = -5,   /* dilation < 1 */
    FIRGPU_E_NULL = -6,       /* NULL pointer with non-empty shape */
    FIRGPU_E_HALO = -7        /* negative halo_*_valid */
};

/* Which kernel family served the last call on this thread (diagnostics and
 * tests: proves the fast path ran, or why it did not). */
enum firgpu_path {
    FIRGPU_PATH_NONE = 0,
    FIRGPU_PATH_GENERIC = 1,      /* int64 accumulate, any shape / alignment / acc_bits */
    FIRGPU_PATH_STAGED = 2,       /* a fast kernel: bulk-copy (TMA) staged / tensor-map tiled packed int32 datapath,
                                     or the register kernels of the floating-point model */
    FIRGPU_PATH_EMPTY = 3         /* zero-sized input, nothing launched */
};

int firgpu_version(void);
const char *firgpu_last_error(void);
int firgpu_last_path(void);
/* Number of kernels this library has launched from the calling thread. */
int64_t firgpu_launch_count(void);
int firgpu_device_count(int *count);
/* Current CUDA device of the calling thread (device entry points run on it). */
int firgpu_get_device(int *device);
int firgpu_set_device(int device);

/* ------------------------------------------------------------------------
 * 1-D fixed-point FIR, row-batched.
 *
 * Stands behind
 *   fir_1d_fixed_golden(x, h, frac_bits=12, acc_bits=32, coeff_bits=16)
 *       fir_1d/model/python/fir_1d_fixed_ref.py:12-130      (rows == 1)
 *   _run_fixed_rowwise(x_u8, h, ...)
 *       fir_1d/sim/vector/gen_fixed_output.py:34-60          (rows == height)
 *
 * For every row r and n in [0, cols):
 *   acc  = sum_k hq[k] * xs[r][n - (k - n_taps/2) * dilation]
 *   y    = sat_0^255( (wrap_to_acc_bits(acc) + 2^(frac_bits-1)) >> frac_bits )
 * xs = input clamped to [0,255]; samples outside [0, cols) are 0, EXCEPT that
 * up to halo_left_valid samples before and halo_right_valid samples after
 * each row are real neighbour data the kernel reads in place (they must be
 * readable memory).  That is how a shard of a longer signal is filtered
 * without a seam; both are 0 for a whole signal.
 * dilation == 2 filters interleaved I/Q (Re and Im independently, the only
 * FIR expressible with lib/mycomplex.cpp:12-22).
 * ------------------------------------------------------------------------ */
int firgpu_fir1d_fixed(const void *x, void *y, int in_dtype, int out_dtype,
                       int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                       const int32_t *hq, int n_taps, int dilation,
                       int frac_bits, int acc_bits,
                       int64_t halo_left_valid, int64_t halo_right_valid,
                       void *stream);

/* Which datapath firgpu_fir1d_fixed would use for these integer taps on uint8 /
 * int16 containers.  Host-only (no device needed): the exactness proof of the
 * packed int32 kernels is made from the taps alone.
 *   *path  FIRGPU_PATH_STAGED (a fast kernel) or FIRGPU_PATH_GENERIC (int64 kernel)
 *   *post  0 = shift, 1 = taps pre-scaled to a 16-bit shift, 2 = live acc_bits
 *          wrap evaluated in the post stage (fir_1d_fixed_ref.py:110-118)
 *   *mac   0 = dp2a (taps fit int16), 1 = IMAD with 32-bit taps
 * Filters of 66 .. 129 taps: the answer given is the one for int16 samples
 * (uint8 samples beyond 65 taps run on the generic kernel). */
int firgpu_fir1d_fixed_plan(const int32_t *hq, int n_taps, int frac_bits, int acc_bits,
                            int *path, int *post, int *mac);

/* Same filter applied with n_filters tap sets to one input, reading x once:
 * hq is [n_filters][n_taps] row-major (host), y is [n_filters] outputs
 * separated by filter_stride_y elements.  Stands behind the coefficient loop
 * of _generate_fixed_outputs_for_tap_map (gen_fixed_output.py:91-104). */
int firgpu_fir1d_fixed_multi(const void *x, void *y, int in_dtype, int out_dtype,
                             int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                             int64_t filter_stride_y,
                             const int32_t *hq, int n_filters, int n_taps, int dilation,
                             int frac_bits, int acc_bits, void *stream);

/* ------------------------------------------------------------------------
 * 1-D ideal (floating-point) FIR, row-batched.
 *
 * Stands behind
 *   fir_1d_ideal(x, h)           fir_1d/model/python/fir_1d_ref.py:43-65
 *   _run_ideal_rowwise(x_u8, h)  fir_1d/sim/vector/gen_ideal_output.py:37-50
 *
 * out_dtype FIRGPU_F64: bit-identical to the reference (float64, taps in
 * ascending k, separate round-to-nearest multiply and add, no FMA).
 * out_dtype FIRGPU_F32: float32 arithmetic, |err| <= 1e-6 * sum_k|h[k]|*255.
 * ------------------------------------------------------------------------ */
int firgpu_fir1d_ideal(const void *x, void *y, int in_dtype, int out_dtype,
                       int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                       const double *h, int n_taps, int dilation,
                       int64_t halo_left_valid, int64_t halo_right_valid,
                       void *stream);

/* ------------------------------------------------------------------------
 * Extended 1-D entry points: the reference model with its fixed choices made
 * parameters.  (centre = n_taps/2, in [0,255], out [0,255], raw = 0) IS the
 * reference model and is served by the same kernels as the plain calls.  Other
 * alignments of the fixed-point model also run on them (the taps are zero-padded
 * to the centred form), other ranges on int16 containers on the wide datapath,
 * the causal raw float64 model on the register kernel; the remaining settings
 * run on the generic int64 / fp64 kernels.
 *
 *   centre   taps ahead of the output position: y[n] = sum_k h[k] x[n-(k-centre)*dilation].
 *            centre = 0 is the causal form of Fir1D::process_sample / fir_1d_idel
 *            (fir_1d/model/cpp/fir_1d_ref.cpp:16-63); with halo_left_valid =
 *            (n_taps-1)*dilation carried samples in front of x it is chunked
 *            streaming with state.
 *   in_lo/in_hi, out_lo/out_hi   sample clamp and output saturation ranges:
 *            (-32768, 32767) twice with out_dtype I16 is the signed int16
 *            output path the reference lists as future work
 *            (fir_1d/docs/fir_1d_3tap_compare_analysis_v1.md:185).
 *   raw      ideal model only: filter the samples as given, no round / clamp
 *            (the C++ Fir1D model takes raw doubles).
 * ------------------------------------------------------------------------ */
int firgpu_fir1d_fixed_ex(const void *x, void *y, int in_dtype, int out_dtype,
                          int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                          const int32_t *hq, int n_taps, int dilation,
                          int frac_bits, int acc_bits,
                          int64_t halo_left_valid, int64_t halo_right_valid,
                          int centre, int in_lo, int in_hi, int out_lo, int out_hi,
                          void *stream);
int firgpu_fir1d_ideal_ex(const void *x, void *y, int in_dtype, int out_dtype,
                          int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                          const double *h, int n_taps, int dilation,
                          int64_t halo_left_valid, int64_t halo_right_valid,
                          int centre, int in_lo, int in_hi, int raw, void *stream);
int firgpu_fir1d_fixed_ex_host(const void *x, void *y, int in_dtype, int out_dtype,
                               int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                               const int32_t *hq, int n_taps, int dilation,
                               int frac_bits, int acc_bits,
                               int centre, int in_lo, int in_hi, int out_lo, int out_hi,
                               int device);
int firgpu_fir1d_ideal_ex_host(const void *x, void *y, int in_dtype, int out_dtype,
                               int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                               const double *h, int n_taps, int dilation,
                               int centre, int in_lo, int in_hi, int raw, int device);

/* ------------------------------------------------------------------------
 * Complex taps on interleaved I/Q  [I0,Q0,I1,Q1,...]  (cols, ld_*, halo_* count
 * scalars and must be even).
 *
 * The reference stops at real taps: lib/mycomplex.h:10-41 defines complex +
 * complex and complex * double only (that is the dilation == 2 mode above).
 * This is the step SURVEY 8(f)4 lists after streaming; the specification is
 * ours, parity unpinned except for hq_im == 0, where it equals
 * the real-tap I/Q filter:
 *   acc_re[n] = sum_k hq_re[k] * xs_re[m] - hq_im[k] * xs_im[m]    m = n - (k - n_taps/2)
 *   acc_im[n] = sum_k hq_re[k] * xs_im[m] + hq_im[k] * xs_re[m]
 *   each through the 1-D post stage (wrap to acc_bits once, + 2^(frac-1),
 *   >> frac, saturate to [out_lo, out_hi]); samples clamped to [in_lo, in_hi]
 *   ((0,255) twice = the reference's ranges, (-32768,32767) = signed baseband).
 * ------------------------------------------------------------------------ */
int firgpu_fir1d_fixed_complex(const void *x, void *y, int in_dtype, int out_dtype,
                               int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                               const int32_t *hq_re, const int32_t *hq_im, int n_taps,
                               int frac_bits, int acc_bits,
                               int64_t halo_left_valid, int64_t halo_right_valid,
                               int in_lo, int in_hi, int out_lo, int out_hi, void *stream);
int firgpu_fir1d_fixed_complex_host(const void *x, void *y, int in_dtype, int out_dtype,
                                    int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                                    const int32_t *hq_re, const int32_t *hq_im, int n_taps,
                                    int frac_bits, int acc_bits,
                                    int in_lo, int in_hi, int out_lo, int out_hi, int device);

/* ------------------------------------------------------------------------
 * Fused fixed-vs-ideal comparison (row-batched, no outputs materialised).
 *
 * Stands behind _compute_metrics
 *   fir_1d/sim/vector/gen_3tap_compare_report.py:67-112 (and the 5-tap twin),
 * which the reference applies to a saved ideal/fixed file pair.  Here both
 * models are evaluated per sample in one pass over x and reduced to
 *   metrics[0] = sum |fixed - ideal|     metrics[4] = #(fixed == 0)
 *   metrics[1] = sum (fixed - ideal)^2   metrics[5] = #(fixed == 255)
 *   metrics[2] = sum (fixed - ideal)     metrics[6] = #(ideal < 0 or > 255)
 *   metrics[3] = max |fixed - ideal|
 * (float64; divide by rows*cols for the report's means).  metrics is a HOST
 * array of 7 doubles; the call synchronises the stream.
 * ------------------------------------------------------------------------ */
int firgpu_fir1d_compare(const void *x, int in_dtype, int64_t rows, int64_t cols, int64_t ld_x,
                         const int32_t *hq, const double *h, int n_taps,
                         int frac_bits, int acc_bits, double *metrics, void *stream);
int firgpu_fir1d_compare_host(const void *x, int in_dtype, int64_t rows, int64_t cols, int64_t ld_x,
                              const int32_t *hq, const double *h, int n_taps,
                              int frac_bits, int acc_bits, double *metrics, int device);

/* ------------------------------------------------------------------------
 * 2-D fixed-point / ideal FIR over a batch of images.
 *
 * The reference has NO fir_2d implementation (fir_2d/model/cpp/CMakeLists.txt
 * is empty); this is the counterpart BASELINE.json asks for, specified in
 * same pre/post-processing as the 1-D model, one accumulator
 * per pixel over the kh x kw window (hq row-major [kh][kw]),
 *   acc = sum_m sum_n hq[m][n] * xs[i - m + kh/2][j - n + kw/2],
 * zero outside the image, except that halo_top_valid rows above and
 * halo_bottom_valid rows below each image are real neighbour rows (row-band
 * shards).  Would stand behind fir_2d.model.python.fir_2d_fixed_ref /
 * fir_2d_ref.
 * ------------------------------------------------------------------------ */
int firgpu_fir2d_fixed(const void *x, void *y, int in_dtype, int out_dtype,
                       int64_t n_img, int64_t height, int64_t width,
                       int64_t ld_x, int64_t ld_y, int64_t img_stride_x, int64_t img_stride_y,
                       const int32_t *hq, int kh, int kw,
                       int frac_bits, int acc_bits,
                       int64_t halo_top_valid, int64_t halo_bottom_valid,
                       void *stream);

int firgpu_fir2d_ideal(const void *x, void *y, int in_dtype, int out_dtype,
                       int64_t n_img, int64_t height, int64_t width,
                       int64_t ld_x, int64_t ld_y, int64_t img_stride_x, int64_t img_stride_y,
                       const double *h, int kh, int kw,
                       int64_t halo_top_valid, int64_t halo_bottom_valid,
                       void *stream);

/* ------------------------------------------------------------------------
 * Host-buffer entry points: what a ctypes caller with NumPy arrays uses.
 * x and y are HOST pointers (pinned memory is copied directly and overlaps
 * fully; pageable memory goes through the library's pinned staging ring).
 * The call is a pipeline over chunks: H2D copy, kernel, D2H copy on three
 * streams, and returns after the last D2H completes.  device < 0 = current.
 * ------------------------------------------------------------------------ */
int firgpu_fir1d_fixed_host(const void *x, void *y, int in_dtype, int out_dtype,
                            int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                            const int32_t *hq, int n_taps, int dilation,
                            int frac_bits, int acc_bits, int device);

/* The same call spread over several devices of one box (devices[n_devices],
 * CUDA ordinals): rows are split evenly, a single long row is cut into
 * contiguous segments that read the taps' reach from the neighbouring host
 * samples (the partitioning of SURVEY 8(e): no collective, no device-to-device
 * traffic).  One host thread and one copy/compute pipeline per device; returns
 * when y is complete. */
int firgpu_fir1d_fixed_host_multi(const void *x, void *y, int in_dtype, int out_dtype,
                                  int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                                  const int32_t *hq, int n_taps, int dilation,
                                  int frac_bits, int acc_bits,
                                  const int *devices, int n_devices);

/* n_filters tap sets over one host input: y holds n_filters outputs,
 * filter_stride_y elements apart (gen_fixed_output.py:91-104 in one call). */
int firgpu_fir1d_fixed_multi_host(const void *x, void *y, int in_dtype, int out_dtype,
                                  int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                                  int64_t filter_stride_y,
                                  const int32_t *hq, int n_filters, int n_taps, int dilation,
                                  int frac_bits, int acc_bits, int device);

int firgpu_fir1d_ideal_host(const void *x, void *y, int in_dtype, int out_dtype,
                            int64_t rows, int64_t cols, int64_t ld_x, int64_t ld_y,
                            const double *h, int n_taps, int dilation, int device);

int firgpu_fir2d_fixed_host(const void *x, void *y, int in_dtype, int out_dtype,
                            int64_t n_img, int64_t height, int64_t width,
                            const int32_t *hq, int kh, int kw,
                            int frac_bits, int acc_bits, int device);

int firgpu_fir2d_ideal_host(const void *x, void *y, int in_dtype, int out_dtype,
                            int64_t n_img, int64_t height, int64_t width,
                            const double *h, int kh, int kw, int device);

/* ------------------------------------------------------------------------
 * Streaming filters with the history resident on the device.
 *
 * Stand behind the reference's C++ class
 *   Fir1D::process_sample / reset     fir_1d/model/cpp/fir_1d_ref.cpp:16-35
 * whose reg_ member carries the last n_taps samples between calls.  A stream
 * object is the causal filter y[n] = sum_k h[k] * x[n - k*dilation] over an
 * unbounded signal delivered in blocks; the (n_taps-1)*dilation most recent
 * scalars live in HBM between calls.  Concatenating the outputs of any
 * blocking of a signal is bit-identical to one firgpu_fir1d_*_ex call with
 * centre = 0 over the whole signal (and, for the raw float64 ideal model, to
 * Fir1D::process_sample sample by sample).
 *
 *   create_fixed  the fixed-point model (parameters as firgpu_fir1d_fixed_ex;
 *                 (0,255,0,255) = the reference's clamp / saturation ranges);
 *   create_ideal  the floating-point model (parameters as firgpu_fir1d_ideal_ex;
 *                 raw = 1 with F64 samples is the C++ Fir1D);
 *   process       x, y: DEVICE pointers to n scalars; stream-ordered and
 *                 asynchronous; blocks of one object must be submitted in
 *                 signal order on streams that respect that order;
 *   process_host  x, y: HOST pointers (pinned memory is DMA'd directly);
 *                 returns when y is complete;
 *   reset         zero the history (Fir1D::reset).
 * device < 0 = the current device.  Objects are thread-safe (one call at a
 * time per object).
 * ------------------------------------------------------------------------ */
typedef struct firgpu_stream_s *firgpu_stream_t;
int firgpu_stream_create_fixed(firgpu_stream_t *out, int in_dtype, int out_dtype,
                               const int32_t *hq, int n_taps, int dilation,
                               int frac_bits, int acc_bits,
                               int in_lo, int in_hi, int out_lo, int out_hi, int device);
int firgpu_stream_create_ideal(firgpu_stream_t *out, int in_dtype, int out_dtype,
                               const double *h, int n_taps, int dilation,
                               int in_lo, int in_hi, int raw, int device);
int firgpu_stream_process(firgpu_stream_t s, const void *x, void *y, int64_t n, void *stream);
int firgpu_stream_process_host(firgpu_stream_t s, const void *x, void *y, int64_t n);
int firgpu_stream_reset(firgpu_stream_t s, void *stream);
int firgpu_stream_history(firgpu_stream_t s, int64_t *scalars);   /* (n_taps-1)*dilation */
int firgpu_stream_destroy(firgpu_stream_t s);

/* ------------------------------------------------------------------------
 * Recorded launch sequences (the batch entry point).
 *
 * The reference pipeline filters 7 images with 8 coefficient sets, one
 * Python call per image row (fir_1d/sim/vector/gen_fixed_output.py:91-104,
 * 44-59); here a whole such sequence of DEVICE entry-point calls is recorded
 * once and replayed with one submission:
 *
 *     firgpu_record_begin(stream);
 *     ... any firgpu_* device entry points on `stream`, this thread ...
 *     firgpu_record_end(stream, &g);          nothing has executed yet
 *     firgpu_graph_launch(g, stream);         as often as wanted
 *
 * Pointers, shapes and taps are frozen into the recording; the buffers'
 * CONTENTS are read at replay time.  Each distinct call pattern must have run
 * once eagerly before it is recorded (first use of a kernel configures it).
 * Entry points that synchronise (firgpu_fir1d_compare, every *_host form)
 * cannot be recorded.  `stream` must not be the NULL stream.
 * ------------------------------------------------------------------------ */
typedef struct firgpu_graph_s *firgpu_graph_t;
int firgpu_record_begin(void *stream);
int firgpu_record_end(void *stream, firgpu_graph_t *out);
int firgpu_record_abort(void *stream);                 /* leave recording mode after an error, discarding it */
int firgpu_graph_launch(firgpu_graph_t g, void *stream);
int firgpu_graph_nodes(firgpu_graph_t g, int64_t *kernels);   /* kernel launches one replay performs */
int firgpu_graph_destroy(firgpu_graph_t g);

/* Page-locked host memory for callers without their own allocator (NumPy
 * arrays wrapped around it are DMA'd directly by the _host entry points). */
int firgpu_pinned_alloc(void **ptr, size_t bytes);
int firgpu_pinned_free(void *ptr);

/* Release cached device / pinned buffers and streams held for the _host
 * pipeline on every device. */
int firgpu_release_cache(void);

#ifdef __cplusplus
}
#endif
#endif /* FIRGPU_H */
