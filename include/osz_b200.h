/*
 * osz_b200.h -- C ABI of the B200-native openseize hot path.
 *
 * The reference (mscaudill/openseize, pure Python) has no native/FFI boundary
 * of its own: its operators call numpy/scipy routines chunk by chunk from the
 * generator functions of src/openseize/core/numerical.py.  This header is the
 * boundary a maintainer would bind (ctypes) to replace those per-chunk calls;
 * each entry point names the reference call site it replaces.  See
 * INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - every function returns 0 (OSZ_OK) or a negative osz_status; the message
 *     of the last failure on the calling thread is osz_last_error().
 *   - data pointers named *_dev are DEVICE pointers, *_host are host pointers.
 *   - signals are time-contiguous rows: sample t of row r is p[r*ld + t]
 *     (ld in elements).  Other layouts are packed with osz_pack_rows_f64.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *     All exec calls are asynchronous on that stream.
 *   - plans are immutable after creation and may be shared between streams.
 *   - there is no CPU path: every exec call launches sm_100a kernels and fails
 *     with OSZ_ERR_CUDA when no device is present.
 */
#ifndef OSZ_B200_H
#define OSZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum osz_status {
    OSZ_OK = 0,
    OSZ_ERR_ARG = -1,      /* invalid argument                              */
    OSZ_ERR_CUDA = -2,     /* CUDA runtime / launch failure                 */
    OSZ_ERR_UNSUPPORTED = -3, /* valid request this build has no kernel for */
    OSZ_ERR_ALLOC = -4
} osz_status;

/* ---- library ---------------------------------------------------------- */
int osz_version(void);
const char *osz_last_error(void);
/* sm_count, compute capability and opt-in shared memory of device `dev`. */
int osz_device_info(int dev, int *sm_count, int *cc_major, int *cc_minor,
                    int64_t *smem_optin_bytes, int64_t *global_mem_bytes);
/* number of kernels this library has launched in this process (all threads);
 * bench.py reports the difference across its timed region as gpu_launches. */
int64_t osz_launch_count(void);

/* ---- runtime plumbing (for hosts that do not bring their own allocator;
 *      replaces nothing in the reference -- it is the pinned-host streaming
 *      substrate under core/producer.py's chunk iterator) ----------------- */
int osz_dev_malloc(void **p_dev, size_t bytes);
int osz_dev_free(void *p_dev);
int osz_host_alloc(void **p_host, size_t bytes);          /* pinned */
int osz_host_free(void *p_host);
int osz_stream_create(void **stream);
int osz_stream_destroy(void *stream);
int osz_stream_sync(void *stream);
int osz_memcpy_h2d_async(void *dst_dev, const void *src_host, size_t bytes, void *stream);
int osz_memcpy_d2h_async(void *dst_host, const void *src_dev, size_t bytes, void *stream);
int osz_memcpy_d2d_async(void *dst_dev, const void *src_dev, size_t bytes, void *stream);
/* strided rows (ArrayProducer yields views: core/producer.py:289-295) */
int osz_memcpy2d_h2d_async(void *dst_dev, size_t dst_pitch_bytes, const void *src_host,
                           size_t src_pitch_bytes, size_t width_bytes, size_t height,
                           void *stream);
int osz_memcpy2d_d2h_async(void *dst_host, size_t dst_pitch_bytes, const void *src_dev,
                           size_t src_pitch_bytes, size_t width_bytes, size_t height,
                           void *stream);
int osz_memset_async(void *dst_dev, int value, size_t bytes, void *stream);

/* (outer, n, inner) <-> (outer*inner rows, n) time-contiguous.
 * src element (o, t, i) is src[(o*n + t)*inner + i]; row = o*inner + i.
 * Replaces the axis-generic slicing of core/arraytools.py:43-82 for data
 * whose sample axis is not the last one. */
int osz_pack_rows_f64(const double *src_dev, int64_t outer, int64_t n, int64_t inner,
                      double *dst_dev, int64_t ld_dst, void *stream);
int osz_unpack_rows_f64(const double *src_dev, int64_t ld_src, int64_t outer, int64_t n,
                        int64_t inner, double *dst_dev, void *stream);
/* complex128 variant (STFT output) */
int osz_unpack_rows_c128(const double *src_dev, int64_t ld_src, int64_t outer, int64_t n,
                         int64_t inner, double *dst_dev, void *stream);
/* z[r][t] = re[r][t] + i im[r][t]: (rows, n) real and imaginary rows -> (rows, n)
 * complex128 (interleaved).  The analytic signal x + i H(x) of
 * experimental/coupling/transforms.py:185-192 (protools.multiply / protools.add of two
 * producers) assembled on the device. */
int osz_zip_complex_f64(const double *re_dev, int64_t ldre, const double *im_dev, int64_t ldim,
                        int64_t rows, int64_t n, double *z_dev, void *stream);
/* widen float32 / int16 chunks to float64 on the device (the reference
 * returns float64 for every input dtype). */
int osz_widen_f32_f64(const float *src_dev, double *dst_dev, int64_t count, void *stream);
int osz_widen_i16_f64(const int16_t *src_dev, double *dst_dev, int64_t count, void *stream);
/* row-strided variants: chunk rows land inside a consumer's halo'd buffer */
int osz_widen_rows_f32_f64(const float *src_dev, int64_t ld_src, double *dst_dev, int64_t ld_dst,
                           int64_t rows, int64_t n, void *stream);
int osz_widen_rows_i16_f64(const int16_t *src_dev, int64_t ld_src, double *dst_dev,
                           int64_t ld_dst, int64_t rows, int64_t n, void *stream);
/* EDF ingest fused into the upload (the reference's Reader._records / _read_array
 * / _decipher, file_io/edf.py:382-419,452-556).  rec_dev holds whole data records as
 * they sit in the file, per_record int16 each, every selected channel with spr
 * samples per record at chan_off[c] inside a record; row c of dst receives samples
 * skip .. skip+n-1 (counted from the first record in rec_dev):
 *   dst[c][i] = rec[(skip+i)/spr][chan_off[c] + (skip+i)%spr] * slope[c] + offset[c]
 * multiply and add rounded separately like numpy's `arr * slopes; result += offsets`.
 * The records cross PCIe as int16: a quarter of the float64 width, and no host-side
 * de-interleaving. */
int osz_decode_edf_records_f64(const int16_t *rec_dev, int64_t per_record, int64_t spr,
                               const int *chan_off_dev, const double *slope_dev,
                               const double *offset_dev, int64_t skip, double *dst_dev,
                               int64_t ld_dst, int64_t rows, int64_t n, void *stream);

/* Arithmetic of a plan (osz_*_plan_set_compute): the reference's float64, or the
 * opt-in float32 mode -- float64 samples in and out, the kernel's arithmetic in
 * float32, within north_star's float32 tolerance (1e-5 of the output peak). */
enum { OSZ_COMPUTE_F64 = 0, OSZ_COMPUTE_F32 = 1 };

/* ---- FIR: replaces _cconvolve + overlap-add of nm.oaconvolve
 *      (core/numerical.py:229-269) --------------------------------------- */
typedef struct osz_fir_plan osz_fir_plan;
enum {
    OSZ_FIR_AUTO = 0,
    OSZ_FIR_DIRECT = 1,
    OSZ_FIR_FFT = 2,
    /* overlap-save FFT evaluated in float32, float64 samples in and out: opt-in
     * (the reference is float64); differs by ~1e-6 of the output peak */
    OSZ_FIR_FFT_F32 = 3
};
/* taps: the window the reference passes to oaconvolve (numerical.py:158). */
int osz_fir_plan_create(osz_fir_plan **plan, const double *taps_host, int ntaps, int algo);
int osz_fir_plan_destroy(osz_fir_plan *plan);
int osz_fir_plan_algo(const osz_fir_plan *plan);   /* resolved algorithm */
/* Valid linear convolution of one halo'd span per row:
 *   y[r][i] = sum_{k<ntaps} taps[k] * x[r][i + ntaps-1-k],  0 <= i < n_out
 * x_dev points at the FIRST halo sample; each row holds n_out + ntaps-1
 * readable samples.  The host keeps the ntaps-1 halo between chunks (the
 * reference's `overlap` carry, numerical.py:220-226,268-269). */
int osz_fir_exec_f64(const osz_fir_plan *plan, const double *x_dev, int64_t ldx,
                     int64_t rows, int64_t n_out, double *y_dev, int64_t ldy, void *stream);

/* ---- IIR biquad cascade: replaces scipy.signal.sosfilt as called by
 *      nm.sosfilt / nm.sosfiltfilt (core/numerical.py:334,399,402,410) and
 *      scipy.signal.lfilter for len(a)=len(b)=3 (:445,508,511,519) -------- */
typedef struct osz_sos_plan osz_sos_plan;
/* sos: (nsec, 6) rows [b0 b1 b2 a0 a1 a2], a0 normalised like scipy. */
int osz_sos_plan_create(osz_sos_plan **plan, const double *sos_host, int nsec);
int osz_sos_plan_destroy(osz_sos_plan *plan);
/* Filter n samples per row through the cascade (DF2T), as a time-parallel
 * scan.  state_dev is (rows, nsec, 2) delay registers, read as the initial
 * condition and overwritten with the final one (the reference's `z`,
 * numerical.py:329-335).  reverse != 0 runs from sample n-1 down to 0 and
 * writes y in place of the same index (flip -> sosfilt -> flip,
 * numerical.py:401-403).  y_dev may be NULL: only the final state is wanted
 * (the look-ahead pass, numerical.py:397-399). */
int osz_sos_exec_f64(const osz_sos_plan *plan, const double *x_dev, int64_t ldx,
                     int64_t rows, int64_t n, int reverse, double *state_dev,
                     double *y_dev, int64_t ldy, void *stream);
/* state[r][s][j] = zi[s][j] * x[r][sample]  (numerical.py:385,399,410):
 * zi_host is scipy.signal.sosfilt_zi(sos), shape (nsec, 2). */
int osz_sos_state_from_sample_f64(const osz_sos_plan *plan, const double *zi_host,
                                  const double *x_dev, int64_t ldx, int64_t rows,
                                  int64_t sample, double *state_dev, void *stream);

/* The look-ahead pass of nm.sosfiltfilt / nm.filtfilt in one call (core/numerical.py:397-399,
 * :508-509: `z = zi * chunk[last]; _, z = sosfilt(sos, flip(chunk), zi=z)`): the state left by
 * filtering x_dev (rows, n) -- reverse: last sample first -- starting from
 * zi_host * (the first sample processed).  state_dev (rows, nsec, 2) is written only. */
int osz_sos_lookahead_f64(const osz_sos_plan *plan, const double *zi_host, const double *x_dev,
                          int64_t ldx, int64_t rows, int64_t n, int reverse, double *state_dev,
                          void *stream);

/* State the cascade holds after filtering the n samples of x_dev (rows, n) FROM REST
 * (reverse: last sample first), without running the recurrence: a weighted sum of the
 * last `settle` samples processed (osz_sos_plan_settle), whose weights decay like the
 * impulse response -- exact to ~1e-18 of the state scale.  This is the look-ahead pass of
 * nm.sosfiltfilt / nm.filtfilt (core/numerical.py:397-399, :508-509), whose start state
 * zi * x[last] has been forgotten after `settle` samples.  One or two sections only
 * (osz_sos_plan_has_weights) and n >= settle; OSZ_ERR_UNSUPPORTED otherwise. */
int osz_sos_tail_state_f64(const osz_sos_plan *plan, const double *x_dev, int64_t ldx,
                           int64_t rows, int64_t n, int reverse, double *state_dev, void *stream);
/* Samples after which the cascade has forgotten its start state (||T^n|| < 1e-18), -1 for
 * a pole on or outside the unit circle. */
int64_t osz_sos_plan_settle(const osz_sos_plan *plan);
int osz_sos_plan_has_weights(const osz_sos_plan *plan);

/* ---- transfer-function IIR of any order: replaces scipy.signal.lfilter as
 *      called by nm.lfilter / nm.filtfilt (core/numerical.py:445,508,511,519)
 *      when max(len a, len b) > 3 (second order rides the biquad scan) ------ */
typedef struct osz_tf_plan osz_tf_plan;
/* b, a as handed to scipy.signal.lfilter (normalised by a[0] like scipy). */
int osz_tf_plan_create(osz_tf_plan **plan, const double *b_host, int nb, const double *a_host,
                       int na);
int osz_tf_plan_destroy(osz_tf_plan *plan);
int osz_tf_plan_states(const osz_tf_plan *plan);   /* max(nb, na) - 1 */
/* Transposed direct form II, evaluated in scipy's order.  state_dev is
 * (rows, states) delay registers: initial condition in, final condition out
 * (the reference's `z`, numerical.py:439-446).  reverse / y_dev == NULL as for
 * osz_sos_exec_f64. */
int osz_tf_exec_f64(const osz_tf_plan *plan, const double *x_dev, int64_t ldx, int64_t rows,
                    int64_t n, int reverse, double *state_dev, double *y_dev, int64_t ldy,
                    void *stream);
/* state[r][j] = zi[j] * x[r][sample]: zi_host is scipy.signal.lfilter_zi(b, a)
 * (numerical.py:487-496,508,519). */
int osz_tf_state_from_sample_f64(const osz_tf_plan *plan, const double *zi_host,
                                 const double *x_dev, int64_t ldx, int64_t rows, int64_t sample,
                                 double *state_dev, void *stream);

/* ---- polyphase resampling: replaces scipy.signal.resample_poly/upfirdn as
 *      called by nm.polyphase_resample (core/numerical.py:610,631) -------- */
typedef struct osz_upfirdn_plan osz_upfirdn_plan;
/* h: anti-alias taps exactly as handed to resample_poly(window=h) (the plan
 * applies scipy's h *= up).  up/down already reduced by their gcd. */
int osz_upfirdn_plan_create(osz_upfirdn_plan **plan, const double *h_host, int ntaps,
                            int up, int down);
int osz_upfirdn_plan_destroy(osz_upfirdn_plan *plan);
/* OSZ_COMPUTE_F32: the decimating kernel (up == 1) narrows the samples when it
 * stages a tile and runs taps, windows and sums in float32; other plans keep
 * float64 (osz_upfirdn_plan_compute tells). */
int osz_upfirdn_plan_set_compute(osz_upfirdn_plan *plan, int compute);
int osz_upfirdn_plan_compute(const osz_upfirdn_plan *plan);
/* Which float64 decimating kernel (up == 1) the plan runs: OSZ_UFD_AUTO (the
 * tensor-core one when its tile geometry fits), OSZ_UFD_POLYPHASE (CUDA-core
 * polyphase filter with register sliding windows) or OSZ_UFD_MMA (banded
 * Toeplitz product on the FP64 tensor cores, mma.sync m8n8k4).  Both evaluate
 * the same sum; they differ in summation order only.  osz_upfirdn_plan_kernel
 * returns the one that will run (OSZ_UFD_GENERAL for up > 1). */
enum { OSZ_UFD_AUTO = 0, OSZ_UFD_POLYPHASE = 1, OSZ_UFD_MMA = 2, OSZ_UFD_GENERAL = 3 };
int osz_upfirdn_plan_set_kernel(osz_upfirdn_plan *plan, int kernel);
int osz_upfirdn_plan_kernel(const osz_upfirdn_plan *plan);
/* Global output sample j of the resampled recording is
 *   y[j] = sum_k h'[j*down + half - k*up] * x[k],  half = (ntaps-1)/2
 * (scipy's resample_poly after its pre-pad/pre-remove bookkeeping).  This call
 * computes j in [out_first, out_first + n_out) for each row.  x_dev[r*ldx + m]
 * holds global input sample (x_first + m); samples outside [x_first,
 * x_first + x_len) that the sum touches are taken as zero (recording edges). */
int osz_upfirdn_exec_f64(const osz_upfirdn_plan *plan, const double *x_dev, int64_t ldx,
                         int64_t rows, int64_t x_first, int64_t x_len, int64_t out_first,
                         int64_t n_out, double *y_dev, int64_t ldy, void *stream);

/* ---- fused IIR pass + decimating FIR: the LAST pass of nm.sosfiltfilt / nm.filtfilt
 *      (the backward scipy.signal.sosfilt / lfilter call, core/numerical.py:402,410,
 *      511,519) -- or a forward nm.sosfilt / nm.lfilter pass (:334,445) -- followed
 *      by the oaconvolve (mode 'same', :229-269) and resample_poly (:610,631) calls of
 *      the two operators a Pipeline composes after it (tools/pipeline.py:109-124).
 *      The pass output stays in shared memory; the decimating filter (the upfirdn
 *      plan built from the convolved taps) runs on it on the tensor cores. -------- */
/* Time spans per row the fused kernel would use (>= 1); 0 = this pair of plans / this
 * chunk length cannot run fused (run the two kernels separately). */
int osz_sosdec_spans(const osz_sos_plan *sos, const osz_upfirdn_plan *ufd, int64_t rows,
                     int64_t n);
/* x_dev (rows, n): input of the pass; state_dev (rows, nsec, 2): state entering the pass,
 * replaced by the state leaving it; reverse: the pass runs from the chunk's last sample to
 * its first; A: global index of the chunk's first sample.  Column c of y_dev is global
 * output out_first + c of  y[j] = sum_k h'[j*down + half - k] * p[k]  (p = the pass output,
 * as in osz_upfirdn_exec_f64); written are the outputs whose whole window of ntaps samples
 * lies inside one span of the chunk.  edges_dev (rows, nspan, 2, ntaps-1) receives the first
 * and last ntaps-1 pass outputs of every span (real-time order) for
 * osz_sosdec_boundary_f64. */
int osz_sosdec_exec_f64(const osz_sos_plan *sos, const osz_upfirdn_plan *ufd, const double *x_dev,
                        int64_t ldx, int64_t rows, int64_t n, int reverse, double *state_dev,
                        int nspan, int64_t A, double *y_dev, int64_t ldy, int64_t out_first,
                        int64_t n_out, double *edges_dev, void *stream);
/* The outputs osz_sosdec_exec_f64 left out: windows that straddle the chunk's start
 * (prev_tail_dev: (rows, ntaps-1) last pass outputs of the previous chunk, row pitch
 * prev_ld; NULL = recording start, zeros before it), a boundary between two spans, or --
 * with has_end -- the chunk's end (recording end: zeros after it).  Only outputs
 * j_min <= j <= j_max are written. */
int osz_sosdec_boundary_f64(const osz_upfirdn_plan *ufd, const double *edges_dev, int nspan,
                            int reverse, const double *prev_tail_dev, int64_t prev_ld,
                            int has_end, int64_t rows, int64_t n, int64_t A, double *y_dev,
                            int64_t ldy, int64_t out_first, int64_t n_out, int64_t j_min,
                            int64_t j_max, void *stream);

/* ---- windowed DFT: replaces detrend + window + rfft + scale of
 *      nm.modified_dft / nm.periodogram (core/numerical.py:691-716,781-794)
 *      applied to the sliding windows of nm._spectra_estimatives (:817-849) */
typedef struct osz_spec_plan osz_spec_plan;
enum { OSZ_DETREND_NONE = 0, OSZ_DETREND_CONSTANT = 1, OSZ_DETREND_LINEAR = 2 };
/* window_host: nfft window coefficients (scipy.signal.get_window, :694);
 * norm: 1/(fs*sum(w^2)) or 1/sum(w)^2 (:703-708). */
int osz_spec_plan_create(osz_spec_plan **plan, int nfft, int stride,
                         const double *window_host, int detrend, double norm);
int osz_spec_plan_destroy(osz_spec_plan *plan);
int osz_spec_plan_path(const osz_spec_plan *plan); /* 1 = shared-memory pow2, 2 = generic */
/* Arithmetic of the plan's transforms.  OSZ_COMPUTE_F64 (default) is the
 * reference's; OSZ_COMPUTE_F32 is the opt-in float32 mode (float64 samples in,
 * float64 sums out; samples are centred in float64, then window product and
 * FFT run in float32 -- within north_star's float32 tolerance, 1e-5 of the
 * largest bin).  It exists for osz_welch_accum_f64, osz_periodogram_f64 and
 * osz_stft_f64 at nfft = 512 .. 4096; other plans keep float64
 * (osz_spec_plan_compute tells).
 * window_host: the same coefficients given to osz_spec_plan_create. */
int osz_spec_plan_set_compute(osz_spec_plan *plan, int compute, const double *window_host);
int osz_spec_plan_compute(const osz_spec_plan *plan);
/* Fused Welch accumulate: for each row adds the one-sided periodograms of
 * segments s = 0..nseg-1 (segment s = x[r][s*stride .. s*stride+nfft)) into
 * psd_sum_dev[r*ldp + k], k <= nfft/2.  The caller divides by the segment
 * count (the reference's running mean, spectra/estimators.py:150-152). */
int osz_welch_accum_f64(const osz_spec_plan *plan, const double *x_dev, int64_t ldx,
                        int64_t rows, int64_t nseg, double *psd_sum_dev, int64_t ldp,
                        void *stream);
/* Per-segment outputs.  out is [seg][row][nfft/2+1]; periodogram: float64,
 * modified DFT (STFT): interleaved complex128. */
int osz_periodogram_f64(const osz_spec_plan *plan, const double *x_dev, int64_t ldx,
                        int64_t rows, int64_t nseg, double *out_dev, void *stream);
int osz_stft_f64(const osz_spec_plan *plan, const double *x_dev, int64_t ldx,
                 int64_t rows, int64_t nseg, double *out_dev, void *stream);
/* One zero-padded segment per row, for periodogram / modified_dft calls with
 * nfft > n (numerical.py:688-699): out[r][i] = (x[r][i] - trend_r(i)) * w[i]
 * for i < n and 0 for n <= i < nfft; trend per `detrend` over the n samples,
 * w = get_window(window, n) already on the device.  The transform then runs
 * on `out` with a unit window and no detrending. */
int osz_spec_prepare_f64(const double *x_dev, int64_t ldx, int64_t rows, int64_t n,
                         int64_t nfft, const double *window_dev, int detrend,
                         double *out_dev, int64_t ldo, void *stream);

/* ---- producer tools on device-resident chunks ---------------
 * Mask compaction of MaskedProducer (core/producer.py:427-444: np.take of the
 * kept samples along the sample axis): y[r][j] = x[r][idx[j]], idx_dev the
 * ascending int64 positions np.flatnonzero(mask chunk) on the device. */
int osz_take_cols_f64(const double *x_dev, int64_t ldx, int64_t rows, const int64_t *idx_dev,
                      int64_t nkeep, double *y_dev, int64_t ldy, void *stream);
/* protools.mean / protools.std along the production axis (core/protools.py:
 * 529-536, 580-590), one call per chunk:  acc[r] += (n * mean_chunk[r],
 * n * mean(chunk^2)[r], n), mean over the values that are not NaN when
 * ignore_nan (np.nanmean) -- the reference's chunk-weighted combination.
 * acc_dev: rows x 3 (zeroed by the caller before the first chunk);
 * scratch_dev: rows x osz_row_moments_slots() x 3 doubles. */
int osz_row_moments_slots(void);
int osz_row_moments_f64(const double *x_dev, int64_t ldx, int64_t rows, int64_t n,
                        int ignore_nan, double *acc_dev, double *scratch_dev, void *stream);
/* protools.standardize along the production axis (:659-662):
 * y[r][i] = (x[r][i] - mean[r]) / std[r]. */
int osz_row_standardize_f64(const double *x_dev, int64_t ldx, int64_t rows, int64_t n,
                            const double *mean_dev, const double *std_dev, double *y_dev,
                            int64_t ldy, void *stream);
/* The same three for an axis that is NOT the production axis (:538-543, 592-595,
 * 663-668): per sample i the np.(nan)mean / np.(nan)std over the rows of the
 * chunk; any of mean_out (n), std_out (n), y (rows x n standardized) may be
 * null. */
int osz_col_moments_f64(const double *x_dev, int64_t ldx, int64_t rows, int64_t n,
                        int ignore_nan, double *mean_out_dev, double *std_out_dev,
                        double *y_dev, int64_t ldy, void *stream);

/* ---- float32 I/O mode (opt-in; the reference returns float64 for every input dtype,
 *      core/numerical.py:699): float samples in and out for the operators
 *      whose arithmetic already runs in float32 -- 8 / 4.2 / 4 bytes per sample of HBM
 *      traffic for FIR / resampling / Welch instead of 16 / 8.4 / 8 -- and for the biquad
 *      scan, whose recurrence, scan and carried state stay float64 (a float32 recurrence
 *      misses north_star's 1e-5 tolerance).  Results within 1e-5 of the
 *      output peak.  Python: openseize_b200.set_io("float32"). --------------------- */
/* plan created with OSZ_FIR_FFT_F32 (at most 1025 taps) */
int osz_fir_exec_f32(const osz_fir_plan *plan, const float *x_dev, int64_t ldx, int64_t rows,
                     int64_t n_out, float *y_dev, int64_t ldy, void *stream);
int osz_sos_exec_f32(const osz_sos_plan *plan, const float *x_dev, int64_t ldx, int64_t rows,
                     int64_t n, int reverse, double *state_dev, float *y_dev, int64_t ldy,
                     void *stream);
int osz_sos_state_from_sample_f32(const osz_sos_plan *plan, const double *zi_host,
                                  const float *x_dev, int64_t ldx, int64_t rows, int64_t sample,
                                  double *state_dev, void *stream);
int osz_sos_lookahead_f32(const osz_sos_plan *plan, const double *zi_host, const float *x_dev,
                          int64_t ldx, int64_t rows, int64_t n, int reverse, double *state_dev,
                          void *stream);
/* decimating plan (up == 1) set to OSZ_COMPUTE_F32 */
int osz_upfirdn_exec_f32(const osz_upfirdn_plan *plan, const float *x_dev, int64_t ldx,
                         int64_t rows, int64_t x_first, int64_t x_len, int64_t out_first,
                         int64_t n_out, float *y_dev, int64_t ldy, void *stream);
/* plan set to OSZ_COMPUTE_F32 (power-of-two nfft 512 ... 4096); sums stay float64 */
int osz_welch_accum_f32(const osz_spec_plan *plan, const float *x_dev, int64_t ldx, int64_t rows,
                        int64_t nseg, double *psd_sum_dev, int64_t ldp, void *stream);
/* rows between the sample types: operators without a float32 kernel of their own run in
 * float64 between a widen (osz_widen_rows_f32_f64) and a narrow */
int osz_narrow_rows_f64_f32(const double *src_dev, int64_t ld_src, float *dst_dev, int64_t ld_dst,
                            int64_t rows, int64_t n, void *stream);
int osz_widen_rows_i16_f32(const int16_t *src_dev, int64_t ld_src, float *dst_dev, int64_t ld_dst,
                           int64_t rows, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OSZ_B200_H */
