/*
 * resampler_b200.h -- C ABI of the B200-native ResamplerFir path.
 *
 * Drop-in boundary for hasenbanck/resampler v0.5.1's `ResamplerFir`
 * (reference: src/resampler_fir.rs, src/fir/, src/window.rs).  The reference
 * has no FFI of its own; these entry points are what a `resampler-cuda` crate
 * binds so that `ResamplerFir::new / new_from_hz /
 * buffer_size_output / resample / delay / reset` keep their meaning, plus the
 * batched multi-stream entry points that are new.
 *
 * Plain C: opaque handle, plain pointers and sizes, int status codes.  All
 * lengths called `*_len` are in f32 VALUES (all channels), exactly like the
 * reference's slices; `*_frames` are per-channel frames.
 *
 * Threading: a handle must not be used from two threads at once (the reference
 * takes `&mut self`, resampler_fir.rs:509); different handles are independent.
 * There is no CPU fallback: every compute entry point fails with
 * RSB_ERR_NO_DEVICE / RSB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef RESAMPLER_B200_H
#define RESAMPLER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rsb_fir rsb_fir; /* N independent streams sharing one configuration, on one GPU */

enum rsb_status {
    RSB_OK = 0,
    /* ResampleError::InvalidInputBufferSize  (src/error.rs:5, checked first, resampler_fir.rs:514) */
    RSB_ERR_INVALID_INPUT_BUFFER_SIZE = 1,
    /* ResampleError::InvalidOutputBufferSize (src/error.rs:7, checked second, resampler_fir.rs:517) */
    RSB_ERR_INVALID_OUTPUT_BUFFER_SIZE = 2,
    /* the reference panics with "input sample rate must be greater than zero" (resampler_fir.rs:302-305) */
    RSB_ERR_ZERO_INPUT_RATE = 10,
    /* the reference panics with "output sample rate must be greater than zero" (resampler_fir.rs:306-309) */
    RSB_ERR_ZERO_OUTPUT_RATE = 11,
    RSB_ERR_INVALID_ARGUMENT = 12,
    /* batched calls only: an `out` buffer was too small for what its calls produced */
    RSB_ERR_OUTPUT_CAPACITY = 13,
    RSB_ERR_NO_DEVICE = 100,
    RSB_ERR_CUDA = 101,
    RSB_ERR_OUT_OF_MEMORY = 102,
    RSB_ERR_PLAN_OVERFLOW = 103
};

/* Latency (resampler_fir.rs:138-162): taps = 16 / 32 / 64 / 128 */
enum rsb_latency { RSB_LATENCY_SAMPLE8 = 0, RSB_LATENCY_SAMPLE16 = 1, RSB_LATENCY_SAMPLE32 = 2,
                   RSB_LATENCY_SAMPLE64 = 3 /* reference default */ };
/* Attenuation (resampler_fir.rs:101-123): Kaiser beta 7 / 10 / 13 */
enum rsb_attenuation { RSB_ATTENUATION_DB60 = 0, RSB_ATTENUATION_DB90 = 1,
                       RSB_ATTENUATION_DB120 = 2 /* reference default */ };

enum rsb_memspace { RSB_MEM_DEVICE = 0, /* pointers are device pointers on the handle's GPU */
                    RSB_MEM_HOST = 1    /* host pointers (pinned is fastest); the call copies */ };

/* Which convolution kernel serves a handle.
 *   EXACT: two-phase dot product in the reference's AVX-512 order (fir/avx512.rs:5-50);
 *          output samples are bit-identical to that path.
 *   FAST : interpolates the two phase rows once per output frame and reuses the row for
 *          every stream/channel of the tile; samples within 1e-6 absolute of EXACT.
 *   TENSOR: the same product on the tcgen05 tensor cores with every operand split into two
 *          fp16 parts after a power-of-two prescale (three fp16 products per tap, fp32
 *          accumulation in tensor memory); samples within 1e-6 absolute of EXACT (measured
 *          <= 6e-7 on full-scale noise).  Serves batches whose streams share one plan and whose
 *          inputs and outputs lie at one constant stride (1, 2, 4 or 8 channels); other batches
 *          fall back to FAST / EXACT.  LAYOUT ADVICE: give every stream's input and output buffer
 *          a 16-byte aligned start and lay them out at one stride that is a multiple of 16 bytes
 *          (e.g. round buffer_size_output() up to a multiple of 4 values): the current kernel
 *          moves rows with TMA; output rows that are not 16-byte aligned are served by its
 *          predecessor (plain stores, ~0.7x the throughput), device input pointers that are not
 *          by FAST / EXACT.
 *
 * Non-finite and out-of-range samples.  EXACT follows IEEE arithmetic sample by sample like the
 * reference: the same outputs become Inf / NaN, all others are bit-identical.  TENSOR and FAST
 * multiply whole tiles (64 / 32 output frames x their input window) with zero-padded filter
 * rows: an Inf / NaN input sample -- and for TENSOR any sample with |x| >= 4094, the fp16 range
 * after its 2^4 prescale -- makes every output of the tiles whose input range contains it
 * unspecified (typically NaN) for that stream AND channel; other channels, other streams and
 * outputs further than one tile plus one filter length away are not affected.  Select
 * RSB_KERNEL_EXACT for material that may contain such samples
 * (tests/test_gpu_tensor_margin.py::test_non_finite_and_out_of_range_samples_contract).
 *   AUTO : TENSOR where it applies (one plan, constant stride, >= 64 stream-channels), else
 *          FAST where it applies (stream count, ratio), else EXACT. */
enum rsb_kernel { RSB_KERNEL_AUTO = 0, RSB_KERNEL_EXACT = 1, RSB_KERNEL_FAST = 2,
                  RSB_KERNEL_TENSOR = 3 };

enum rsb_flags {
    RSB_FLAG_NONE = 0,
    RSB_FLAG_ASYNC = 1,        /* device memspace: return after enqueueing (without it a call
                                  returns when its outputs are complete).  The count
                                  arrays are written by rsb_fir_sync() or by the fourth
                                  submit after this one (four submits may be in flight), so
                                  they must stay valid until then.  Host memspace: honoured by
                                  rsb_fir_process_batch calls that run as a pipeline of time
                                  slices (pinned buffers; counts are written at once, the output
                                  buffers are complete after rsb_fir_sync(); back-to-back calls
                                  then overlap one call's last copies with the next one's first);
                                  ignored (synchronous) otherwise */
    RSB_FLAG_RECORD_CALLS = 2, /* keep per-call (consumed, produced) for rsb_fir_last_call_counts */
    RSB_FLAG_KEEP_PLAN = 4     /* keep the device plan for rsb_fir_last_plan */
};

const char *rsb_version(void);
const char *rsb_status_string(int status);
const char *rsb_last_error(void); /* thread-local detail of the last failure */
int rsb_device_count(void);       /* usable CUDA devices; 0 when none / no driver */

/* ---- construction: ResamplerFir::new_from_hz (resampler_fir.rs:295-404) ----
 * One handle = n_streams independent resamplers with identical parameters on GPU
 * `device`.  ResamplerFir::new (resampler_fir.rs:252-266) maps SampleRate -> Hz
 * (lib.rs:219-234) and calls this. */
int rsb_fir_create(rsb_fir **out, int device, uint32_t n_streams, uint32_t channels,
                   uint32_t input_rate_hz, uint32_t output_rate_hz, int latency, int attenuation);
void rsb_fir_destroy(rsb_fir *h);
int rsb_fir_set_kernel(rsb_fir *h, int kernel);

uint32_t rsb_fir_channels(const rsb_fir *h);
/* Kernel (rsb_kernel, never AUTO) the most recent batch actually ran on; AUTO before the first. */
int rsb_fir_last_kernel(const rsb_fir *h);
uint32_t rsb_fir_n_streams(const rsb_fir *h);
uint32_t rsb_fir_taps(const rsb_fir *h);
double rsb_fir_ratio(const rsb_fir *h);
int rsb_fir_device(const rsb_fir *h);
/* ResamplerFir::buffer_size_output (resampler_fir.rs:456-465), in values */
size_t rsb_fir_buffer_size_output(const rsb_fir *h);
/* ResamplerFir::delay (resampler_fir.rs:630-632) */
size_t rsb_fir_delay(const rsb_fir *h);
/* ResamplerFir::reset (resampler_fir.rs:638-642); stream < 0 resets every stream */
int rsb_fir_reset(rsb_fir *h, int64_t stream);
/* copies the [1024][taps] coefficient table (resampler_fir.rs:406-421) to `out` */
int rsb_fir_coeffs(const rsb_fir *h, float *out, size_t out_len);

/* ---- ResamplerFir::resample (resampler_fir.rs:509-621) for one stream, host slices.
 * Synchronous.  Returns RSB_ERR_INVALID_INPUT_BUFFER_SIZE / _OUTPUT_ (in that order)
 * when a length is not a multiple of `channels`. */
int rsb_fir_resample(rsb_fir *h, uint32_t stream, const float *input, size_t input_len,
                     float *output, size_t output_len, size_t *consumed, size_t *produced);

/* ---- batched submit: n independent resample() calls, one per listed stream ----
 * Semantically identical to calling rsb_fir_resample once for every i in [0, n)
 * with (streams[i], in[i], in_lens[i], out[i], out_lens[i]); streams == NULL means
 * 0..n-1; a stream may appear at most once.  consumed/produced receive values. */
int rsb_fir_submit_batch(rsb_fir *h, uint32_t n, const uint32_t *streams,
                         const float *const *in, const size_t *in_lens, float *const *out,
                         const size_t *out_lens, size_t *consumed, size_t *produced,
                         int memspace, uint32_t flags);

/* ---- batched multi-call: the canonical caller loop per stream, in one launch ----
 * For every listed stream runs, as if sequentially on that stream,
 *     while off < total_len { c,p = resample(in[off .. off+min(call_len, total_len-off)],
 *                                           buf[out_cap_len]); append buf[..p] to out;
 *                             off += c; if c == 0 break }
 * (resample/src/main.rs:226-254, resampler_fir.rs:719-735).  out_cap_len == 0 means
 * buffer_size_output().  Appended output goes to out[i] (capacity out_capacities[i]
 * values; too small => RSB_ERR_OUTPUT_CAPACITY, nothing is written past the end).
 * consumed_totals / produced_totals / n_calls (each may be NULL) receive per-stream
 * totals in values / calls. */
int rsb_fir_process_batch(rsb_fir *h, uint32_t n, const uint32_t *streams,
                          const float *const *in, const size_t *total_lens, size_t call_len,
                          size_t out_cap_len, float *const *out, const size_t *out_capacities,
                          size_t *consumed_totals, size_t *produced_totals, uint32_t *n_calls,
                          int memspace, uint32_t flags);

/* ---- CLI batch path: the format step in front of the canonical loop ----
 * (resample/src/main.rs:128-156, then resample_batch_fir :226-254)
 * Sample formats as hound 3.5 hands them to the CLI (`samples::<i32>()` / `samples::<f32>()`):
 * integer samples become `s as f32 / 2^(bits-1)` (main.rs:131-136); 8-bit WAV data is unsigned
 * on disk and 128 is subtracted first; 24-bit samples are 3 packed little-endian bytes.
 * Reference quirk kept for parity: for 32-bit samples `(1 << 31) as f32` is an i32 shift,
 * i.e. -2^31, so RSB_PCM_S32 yields -s / 2^31 (inverted polarity) exactly like the CLI. */
enum rsb_pcm_format { RSB_PCM_U8 = 0, RSB_PCM_S16 = 1, RSB_PCM_S24 = 2, RSB_PCM_S32 = 3,
                      RSB_PCM_F32 = 4 };
/* Like rsb_fir_process_batch, but in[i] points at in_frames[i] frames of raw samples with
 * src_channels interleaved channels.  src_channels == 1: every sample is duplicated into all
 * of the handle's channels (main.rs:139-146, mono -> stereo); src_channels == channels: passes
 * through (:148-150); anything else: RSB_ERR_INVALID_ARGUMENT ("Unsupported channel count",
 * :151-154).  The conversion runs on the GPU (one pass, pcm_ingest.cu) into a staging buffer
 * owned by the handle; host memspace moves the RAW samples over PCIe (half the bytes of f32
 * for 16-bit sources).  call_len, out_cap_len, consumed_totals and produced_totals are in f32
 * values of the converted signal (the CLI uses call_len = 512).  RSB_FLAG_ASYNC as for
 * rsb_fir_process_batch (device memspace; the raw buffers must stay valid until the sync). */
int rsb_fir_process_pcm_batch(rsb_fir *h, uint32_t n, const uint32_t *streams,
                              const void *const *in, const size_t *in_frames, int format,
                              uint32_t src_channels, size_t call_len, size_t out_cap_len,
                              float *const *out, const size_t *out_capacities,
                              size_t *consumed_totals, size_t *produced_totals, uint32_t *n_calls,
                              int memspace, uint32_t flags);
/* ---- opt-in tail handling (NOT in the reference) ----
 * The reference never flushes: the last taps - 1 input frames stay in the history and the
 * output lags by delay() = taps / 2 input frames (resampler_fir.rs:630-632).  This feeds
 * delay() frames of silence to every listed stream -- exactly
 * resample(&[0.0; delay * channels], out) per stream -- and appends what that produces
 * (values, in produced[]) to out[i].  The streams' state carries on as after any other call.
 * Synchronous. */
int rsb_fir_flush_batch(rsb_fir *h, uint32_t n, const uint32_t *streams, float *const *out,
                        const size_t *out_capacities, size_t *produced, int memspace,
                        uint32_t flags);
/* 1 when the most recent PCM batch ran with the format step INSIDE the tensor kernel (s16 or
 * packed s24 sources of mono or stereo streams, equally long, rows 16-byte aligned at one constant stride,
 * tensor kernel selected): the kernel's TMA producer streams the raw frames and its splitter converts them;
 * only each stream's last 4096 frames go through the separate format-step kernel (they feed
 * the history).  Results are bit-identical to the unfused path. */
int rsb_fir_last_pcm_fused(const rsb_fir *h);
/* device time (ms, CUDA events) of the format-step kernel of the most recent PCM batch */
int rsb_fir_last_ingest_ms(rsb_fir *h, float *ms);

/* completes RSB_FLAG_ASYNC work (writes the deferred counts) and waits for the GPU.  A failure of an
 * asynchronous submit (plan workspace overflow, output capacity) that surfaced while a LATER submit
 * was being started does not abort that later submit: it is returned here, once. */
int rsb_fir_sync(rsb_fir *h);

/* per-call (consumed, produced) values of job `job` of the last synchronous batch
 * submitted with RSB_FLAG_RECORD_CALLS */
int rsb_fir_last_call_counts(rsb_fir *h, uint32_t job, uint32_t *consumed, uint32_t *produced,
                             size_t max_calls, size_t *n_calls);
/* per-output-frame plan of job `job` of the last synchronous batch submitted with
 * RSB_FLAG_KEEP_PLAN, expanded ON THE DEVICE by the same code the kernels use:
 * input_offset (relative to the call's read position), phase1, phase2, bits of frac
 * (resampler_fir.rs:544, 558-565).  Arrays may be NULL. */
int rsb_fir_last_plan(rsb_fir *h, uint32_t job, uint32_t *input_offset, uint32_t *phase1,
                      uint32_t *phase2, uint32_t *frac_bits, size_t max_frames, size_t *n_frames);

/* ---- timing on the handle's CUDA stream (for benchmarks) ---- */
int rsb_fir_timer_start(rsb_fir *h);
int rsb_fir_timer_stop(rsb_fir *h, float *elapsed_ms); /* waits for the stop event */
/* device time (ms, CUDA events on the handle's stream) of the convolution kernel of the most
 * recent batches, newest first, at most 64; waits for the stream */
int rsb_fir_conv_times(rsb_fir *h, float *ms, size_t max, size_t *n);
/* debug: per-phase SM cycle totals of the fast kernel (thread 0 of every CTA): [0] tile setup,
 * [1] TMA issue + zero fill, [2] banded-row build, [3] TMA wait, [4] de-interleave,
 * [5] register-tiled product, [6] store.  Returns and clears the counters; sets the enable flag. */
int rsb_debug_phase_cycles(rsb_fir *h, int enable, uint64_t *out8);
/* debug: per-role SM cycle totals of the tensor kernel (CTA 0; the index list is in
 * fir_tc2.cu).  Copies min(count, 24) counters into out (may be NULL), clears them and sets
 * the enable flag. */
int rsb_debug_tc_cycles(rsb_fir *h, int enable, uint64_t *out, uint32_t count);
/* debug: watchdog record of a tensor-kernel launch that failed on a stuck barrier wait:
 * {wait tag, block, warp, parity | barrier address << 8}, then for warp w of that block at [4 + w]:
 * tag | parity << 8 | barrier address << 12; zeros when none.  Clears the record. */
int rsb_debug_tc_hang(uint32_t out32[32]);
/* kernels launched on this handle since creation (your own count for gpu_launches) */
uint64_t rsb_fir_launch_count(const rsb_fir *h);
/* the handle's cudaStream_t, as an opaque pointer */
void *rsb_fir_cuda_stream(const rsb_fir *h);
/* Host-memspace rsb_fir_process_batch calls that ran as a pipeline of time slices (host->device
 * copy of slice k+1, kernels of slice k and device->host copy of slice k-1 overlapped; results
 * identical to the unsliced loop), and the slices they were cut into. */
int rsb_fir_host_pipeline_stats(const rsb_fir *h, uint64_t *batches, uint64_t *slices);
/* Pinned-memory copy rates of the device's PCIe link in GB/s: mode 0 host->device alone (out[0]),
 * 1 device->host alone (out[1]), 2 both at once, 3 both at once as 2-D copies of 1024 rows (the
 * shape of a batch slice: one row per stream).  The ceiling host-memspace calls are held against. */
int rsb_pcie_probe(int device, size_t bytes, int iters, int mode, double out[2]);

/* ---- memory helpers ---- */
void *rsb_alloc_pinned(size_t bytes);
void rsb_free_pinned(void *p);
void *rsb_alloc_device(int device, size_t bytes);
void rsb_free_device(int device, void *p);
/* kind: 0 host->device, 1 device->host, 2 device->device; synchronous */
int rsb_memcpy(int device, void *dst, const void *src, size_t bytes, int kind);
/* fills a device buffer of n f32 values with the synthetic test signal of SURVEY.md 8(d):
 * 0.5*sin(2*pi*f*frame/rate) + 0.25*u, f = 110*(1 + (stream % 64)) + 7*channel,
 * u = splitmix64 counter hash in [-1, 1).  Layout [stream][frame][channel]. */
int rsb_fill_synthetic(int device, float *dst, uint32_t first_stream, uint32_t n_streams,
                       uint64_t frames, uint32_t channels, uint32_t rate_hz, uint64_t seed);

/* device micro-benchmarks backing the kernel design (tools/microbench.py): id 0 scalar FFMA
 * peak, id 1 packed FFMA2 peak, id >= 10 shared-memory banded-GEMM inner-loop candidates.
 * result = TFLOP/s, aux = variant specific (resident CTAs per SM). */
int rsb_microbench(int device, int id, int arg, double *result, double *aux);

/* ---- ResamplerFft (src/resampler_fft.rs): the reference's overlap-add FFT resampler, batched ----
 * Fixed-size chunks: every call consumes chunk_size_input() values and produces chunk_size_output()
 * values per stream (both include the channels); rates are the SampleRate enum's ten values. */
typedef struct rsb_fft rsb_fft;
int rsb_fft_create(rsb_fft **out, int device, uint32_t n_streams, uint32_t channels,
                   uint32_t input_rate_hz, uint32_t output_rate_hz);      /* ResamplerFft::new :75-128 */
void rsb_fft_destroy(rsb_fft *h);
size_t rsb_fft_chunk_size_input(const rsb_fft *h);                         /* :135-137 */
size_t rsb_fft_chunk_size_output(const rsb_fft *h);                        /* :143-145 */
size_t rsb_fft_delay(const rsb_fft *h);                                    /* :151-153 */
int rsb_fft_reset(rsb_fft *h, int64_t stream);                             /* clears the overlap (-1: all streams) */
/* one chunk of one stream, host slices: ResamplerFft::resample (:182-246), status 1 / 2 for a short
 * input / output buffer in the reference's order */
int rsb_fft_resample(rsb_fft *h, uint32_t stream, const float *input, size_t input_len, float *output,
                     size_t output_len);
/* batched: per job min(in_len / chunk_size_input, out_len / chunk_size_output) consecutive chunks
 * (written to chunks_done), one launch; RSB_FLAG_ASYNC for device memspace */
int rsb_fft_process_batch(rsb_fft *h, uint32_t n, const uint32_t *streams, const float *const *in,
                          const size_t *in_lens, float *const *out, const size_t *out_lens,
                          size_t *chunks_done, int memspace, uint32_t flags);
int rsb_fft_sync(rsb_fft *h);
uint64_t rsb_fft_launch_count(const rsb_fft *h);
const char *rsb_fft_last_error(void);

/* Device-side filter design (window.rs:17-131 on the GPU, bit-identical to the host design; worth it
 * only for callers that create resamplers for many distinct rate pairs: a table is designed once
 * per (cutoff bits, taps, attenuation) and cached either way, like the reference's FIR_CACHE).
 * rsb_set_device_filter_design: tables missing from the cache are designed on the handle's GPU from
 * now on (returns the previous setting).  rsb_device_design_table: uncached, for tests / timing;
 * out receives the [1024][taps] table, elapsed_ms (optional) the device time of its kernels. */
int rsb_set_device_filter_design(int enable);
int rsb_device_design_table(int device, uint32_t input_rate_hz, uint32_t output_rate_hz, int latency,
                            int attenuation, float *out, size_t out_len, float *elapsed_ms);

/* ---- host-only entry points (no GPU needed): filter design and the phase planner ---- */
/* host twin of the device-side design's sine (a restatement of glibc's sinf): equals libm's sinf bit
 * for bit on the hosts this library is built for; the CPU tests check exactly that */
float rsb_host_sinf_restated(float x);
double rsb_host_bessel_i0(double x);                              /* window.rs:96-112 */
double rsb_host_cutoff_kaiser(uint32_t taps, double beta);        /* window.rs:114-131 */
void rsb_host_kaiser_window(uint32_t n, double beta, int symmetric, float *out); /* :66-94 */
void rsb_host_make_sincs(uint32_t sample_count, uint32_t factor, float cutoff, double beta,
                         int symmetric, float *out);              /* window.rs:17-55 */
/* table exactly as rsb_fir_create builds it; out_len >= 1024*taps */
int rsb_host_design_table(uint32_t input_rate_hz, uint32_t output_rate_hz, int latency,
                          int attenuation, float *out, size_t out_len, uint32_t *cutoff_bits);
/* Runs the planner (the same code the device executes) on the host for one stream:
 * starting from (position_bits, available) performs the canonical loop (single_call != 0:
 * exactly one call) and returns per-call counts in frames and the expanded per-frame plan.
 * Any output array may be NULL.  Returns the number of calls. */
size_t rsb_host_plan(uint32_t input_rate_hz, uint32_t output_rate_hz, int latency,
                     uint64_t *position_bits, uint32_t *available, uint64_t total_frames,
                     uint32_t call_frames, uint32_t cap_frames, int single_call,
                     uint32_t *calls_consumed, uint32_t *calls_produced, size_t max_calls,
                     uint32_t *input_offset, uint32_t *phase1, uint32_t *phase2,
                     uint32_t *frac_bits, size_t max_frames, size_t *n_frames,
                     size_t *n_segments);

#ifdef __cplusplus
}
#endif
#endif /* RESAMPLER_B200_H */
