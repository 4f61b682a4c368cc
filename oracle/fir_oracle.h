/*
 * fir_oracle.h -- CPU oracle for the ResamplerFir hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  The product
 * (resampler_b200/) never links or imports this code.
 *
 * What it is: a line-by-line C restatement of hasenbanck/resampler v0.5.1
 *   src/window.rs:17-131         (Kaiser-sinc polyphase filter design)
 *   src/resampler_fir.rs:295-404 (constructor: ratio, taps, cutoff)
 *   src/resampler_fir.rs:456-465 (buffer_size_output)
 *   src/resampler_fir.rs:509-621 (resample(): the streaming state machine)
 *   src/fir/avx512.rs:5-50       (dual-phase dot product, AVX-512 order)
 *   src/fir/mod.rs:47-62         (scalar dot product)
 *   resample/src/main.rs:128-156 (CLI format step: integer PCM -> f32, mono -> stereo)
 * The Rust reference cannot be compiled in this image (no rustc/cargo), so
 * there is no oracle/_ref build.
 *
 * Pinning status:
 *   - filter design: PINNED against the reference's own known-answer tests
 *     (src/window.rs:152-410), replayed in tests/test_oracle_golden.py.
 *   - convolution formula: PINNED to the reference's SIMD-vs-scalar test
 *     (src/fir/mod.rs:137-192, 1e-5) and the >= 90 dB stop-band test
 *     (src/resampler_fir.rs:741-815).
 *   - CLI format step (orc_pcm_to_f32): PARITY UNPINNED by the reference (the CLI has no
 *     tests); the arithmetic is one exact division by a power of two per sample and is
 *     checked against hand-computed values in tests/test_oracle_golden.py.  The i32 values
 *     come from hound 3.5.1 (Cargo.lock; not vendored): 8-bit WAV samples are unsigned bytes
 *     minus 128, 16/24/32-bit ones are little-endian two's complement.
 *   - (consumed, produced) sequences, phase indices and output sample values:
 *     PARITY UNPINNED by the reference -- no reference test or fixture holds
 *     any of them.  They rest on this restatement, cross-checked by a second,
 *     independent numpy restatement (oracle/plan_numpy.py) and by the
 *     characterisation vectors of SURVEY.md section 8(a).
 */
#ifndef FIR_ORACLE_H
#define FIR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_PHASES 1024         /* resampler_fir.rs:17 */
#define ORC_INPUT_CAPACITY 4096 /* resampler_fir.rs:18 */
#define ORC_BUFFER_SIZE 8192    /* resampler_fir.rs:19 */

enum { ORC_OK = 0, ORC_ERR_INPUT_SIZE = 1, ORC_ERR_OUTPUT_SIZE = 2 };
enum { ORC_CONV_AVX512 = 0, ORC_CONV_SCALAR = 1, ORC_CONV_AVX512_INTRIN = 2 };

/* ---- filter design (window.rs) ---- */
double orc_bessel_i0(double x);                                   /* window.rs:96-112 */
void orc_kaiser_window(size_t n, double beta, int symmetric, float *out); /* window.rs:66-94 */
double orc_cutoff_kaiser(size_t sample_count, double beta);       /* window.rs:114-131 */
/* out is [factor][sample_count] row-major.  window.rs:17-55 */
void orc_make_sincs(size_t sample_count, size_t factor, float f_cutoff, double beta,
                    int symmetric, float *out, float *sum_out);
double orc_attenuation_beta(int attenuation);                     /* resampler_fir.rs:117-123 */
int orc_latency_taps(int latency);                                /* resampler_fir.rs:153-161 */
/* cutoff as the constructor derives it (resampler_fir.rs:311-326). */
float orc_cutoff_for(uint32_t in_hz, uint32_t out_hz, int taps, double beta);

/* ---- dot products ---- */
float orc_convolve_avx512_order(const float *x, const float *c1, const float *c2, float frac,
                                size_t taps);                     /* fir/avx512.rs:5-50 */
float orc_convolve_scalar(const float *x, const float *c1, const float *c2, float frac,
                          size_t taps);                           /* fir/mod.rs:47-62 */
/* the same routine with real AVX-512 intrinsics (cpu_bench.c, built with -mavx512f);
 * only callable when orc_cpu_has_avx512f() */
float orc_convolve_avx512_intrin(const float *x, const float *c1, const float *c2, float frac,
                                 size_t taps);

/* ---- the streaming resampler (resampler_fir.rs:179-201) ---- */
typedef struct orc_fir {
    size_t channels;
    float *coeffs;          /* [1024][taps], 64-byte aligned */
    float *input_buffers;   /* [channels][8192] planar */
    size_t read_position;
    size_t available_frames;
    double position;
    double ratio;
    size_t taps;
    int conv_kind;
} orc_fir;

/* Per-output-frame trace, filled when non-NULL pointers are handed to
 * orc_fir_resample_traced (test instrumentation; not in the reference). */
typedef struct orc_trace {
    uint32_t *input_offset;
    uint32_t *phase1;
    uint32_t *phase2;
    uint32_t *frac_bits;
    size_t capacity;   /* entries available in each array */
    size_t count;      /* entries written by the last call */
} orc_trace;

/* returns NULL when a rate is zero (the reference panics, :302-309) */
orc_fir *orc_fir_new(size_t channels, uint32_t in_hz, uint32_t out_hz, int latency,
                     int attenuation);
void orc_fir_free(orc_fir *r);
void orc_fir_set_conv(orc_fir *r, int conv_kind);
size_t orc_fir_buffer_size_output(const orc_fir *r);              /* :456-465 */
size_t orc_fir_delay(const orc_fir *r);                           /* :630-632 */
void orc_fir_reset(orc_fir *r);                                   /* :638-642 */
int orc_fir_resample(orc_fir *r, const float *input, size_t in_len, float *output,
                     size_t out_len, size_t *consumed, size_t *produced);  /* :509-621 */
int orc_fir_resample_traced(orc_fir *r, const float *input, size_t in_len, float *output,
                            size_t out_len, size_t *consumed, size_t *produced,
                            orc_trace *trace);

/* The canonical caller loop (resample/src/main.rs:226-254 and
 * resampler_fir.rs:719-735): feeds `total_len` values in chunks of `call_len`
 * values, output buffer of `out_cap_len` values per call (0 = buffer_size_output),
 * appends to `out` (capacity out_capacity values).  Per-call counts are stored
 * when consumed_calls/produced_calls are non-NULL (capacity max_calls).
 * Returns the number of calls made; *out_total = values produced. */
size_t orc_fir_process(orc_fir *r, const float *in, size_t total_len, size_t call_len,
                       size_t out_cap_len, float *out, size_t out_capacity, size_t *out_total,
                       size_t *in_total, uint32_t *consumed_calls, uint32_t *produced_calls,
                       size_t max_calls, orc_trace *trace);

/* ---- CLI format step (resample/src/main.rs:128-156) ----
 * format: 0 = unsigned 8-bit, 1 = s16le, 2 = packed s24le, 3 = s32le, 4 = f32.
 * Converts n_src source values; each is written `dup` times in a row (dup = 2 for the CLI's
 * mono -> stereo duplication, :139-146; 1 for stereo, :148-150).  dst holds n_src * dup. */
void orc_pcm_to_f32(const void *src, int format, size_t n_src, size_t dup, float *dst);

/* ---- multi-threaded CPU baseline (cpu_bench.c) ---- */
/* Runs `n_streams` independent resamplers, one stream per thread at a time
 * (streams handed out round-robin to `n_threads` threads), each through the
 * canonical caller loop with `call_frames`-frame calls.  in: [n_streams][frames*channels]
 * (stream stride in_stride values); out: [n_streams][out_stride].  use_avx512
 * selects real AVX-512 intrinsics (the reference's runtime pick on such hosts,
 * resampler_fir.rs:332) when the CPU has them; otherwise the scalar-coded
 * 16-lane restatement.  Returns wall seconds of the threaded region;
 * produced[n_streams] receives values produced per stream. */
double orc_cpu_bench(size_t n_streams, size_t channels, uint32_t in_hz, uint32_t out_hz,
                     int latency, int attenuation, const float *in, size_t in_stride,
                     size_t frames, size_t call_frames, float *out, size_t out_stride,
                     uint64_t *produced, int n_threads, int use_avx512);
int orc_cpu_has_avx512f(void);

#ifdef __cplusplus
}
#endif
#endif
