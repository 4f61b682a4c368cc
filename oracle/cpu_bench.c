/*
 * cpu_bench.c -- multi-threaded CPU baseline (TEST/BENCH INFRASTRUCTURE ONLY).
 *
 * The reference's best-SIMD path: fir/avx512.rs:5-50 with real AVX-512
 * intrinsics, driven by the restated resample() loop (fir_oracle.c), one
 * stream per thread (the reference spawns no threads itself; SURVEY.md 8(d)
 * defines the multi-threaded baseline as one stream per thread).
 * Build: gcc -O3 -mavx512f -mfma -ffp-contract=off.
 */
#define _GNU_SOURCE
#include "fir_oracle.h"

#include <immintrin.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

int orc_cpu_has_avx512f(void) { return __builtin_cpu_supports("avx512f") ? 1 : 0; }

/* fir/avx512.rs:5-50, intrinsic for intrinsic. */
__attribute__((target("avx512f,fma")))
float orc_convolve_avx512_intrin(const float *x, const float *c1, const float *c2, float frac,
                                 size_t taps) {
    size_t iters = taps / 16;
    __m512 acc1 = _mm512_setzero_ps();
    __m512 acc2 = _mm512_setzero_ps();
    for (size_t i = 0; i < iters; ++i) {
        size_t off = i * 16;
        __m512 xv = _mm512_loadu_ps(x + off);
        __m512 k1 = _mm512_load_ps(c1 + off);
        __m512 k2 = _mm512_load_ps(c2 + off);
        acc1 = _mm512_fmadd_ps(k1, xv, acc1);
        acc2 = _mm512_fmadd_ps(k2, xv, acc2);
    }
    __m512 fv = _mm512_set1_ps(frac);
    __m512 omf = _mm512_set1_ps(1.0f - frac);
    __m512 w1 = _mm512_mul_ps(acc1, omf);
    __m512 w2 = _mm512_mul_ps(acc2, fv);
    __m512 v = _mm512_add_ps(w1, w2);
    /* _mm512_reduce_add_ps as Rust's stdarch spells it: 512 -> 256 -> 128 ->
     * (a + movehl) -> lane0 + lane1 */
    __m256 lo = _mm512_castps512_ps256(v);
    __m256 hi = _mm256_castpd_ps(_mm512_extractf64x4_pd(_mm512_castps_pd(v), 1));
    __m256 s8 = _mm256_add_ps(lo, hi);
    __m128 s4 = _mm_add_ps(_mm256_castps256_ps128(s8), _mm256_extractf128_ps(s8, 1));
    __m128 s2 = _mm_add_ps(s4, _mm_movehl_ps(s4, s4));
    float a = _mm_cvtss_f32(s2);
    float b = _mm_cvtss_f32(_mm_shuffle_ps(s2, s2, 1));
    return a + b;
}

typedef struct {
    size_t n_streams, channels, in_stride, frames, call_frames, out_stride;
    uint32_t in_hz, out_hz;
    int latency, attenuation, conv_kind;
    const float *in;
    float *out;
    uint64_t *produced;
    int tid, n_threads;
} bench_job;

static void *bench_worker(void *arg) {
    bench_job *j = (bench_job *)arg;
    for (size_t s = (size_t)j->tid; s < j->n_streams; s += (size_t)j->n_threads) {
        orc_fir *r = orc_fir_new(j->channels, j->in_hz, j->out_hz, j->latency, j->attenuation);
        orc_fir_set_conv(r, j->conv_kind);
        size_t total = 0;
        orc_fir_process(r, j->in + s * j->in_stride, j->frames * j->channels,
                        j->call_frames * j->channels, 0, j->out ? j->out + s * j->out_stride : NULL,
                        j->out ? j->out_stride : 0, &total, NULL, NULL, NULL, 0, NULL);
        j->produced[s] = total;
        orc_fir_free(r);
    }
    return NULL;
}

double orc_cpu_bench(size_t n_streams, size_t channels, uint32_t in_hz, uint32_t out_hz,
                     int latency, int attenuation, const float *in, size_t in_stride,
                     size_t frames, size_t call_frames, float *out, size_t out_stride,
                     uint64_t *produced, int n_threads, int use_avx512) {
    if (n_threads < 1) n_threads = 1;
    /* build (and cache) the coefficient table outside the timed region, as the
     * reference's constructor would before the stream starts */
    orc_fir *warm = orc_fir_new(channels, in_hz, out_hz, latency, attenuation);
    if (!warm) return -1.0;
    orc_fir_free(warm);
    int kind = (use_avx512 && orc_cpu_has_avx512f()) ? ORC_CONV_AVX512_INTRIN : ORC_CONV_AVX512;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    bench_job *jobs = (bench_job *)malloc(sizeof(bench_job) * (size_t)n_threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < n_threads; ++t) {
        bench_job *j = &jobs[t];
        j->n_streams = n_streams; j->channels = channels; j->in_stride = in_stride;
        j->frames = frames; j->call_frames = call_frames; j->out_stride = out_stride;
        j->in_hz = in_hz; j->out_hz = out_hz; j->latency = latency; j->attenuation = attenuation;
        j->conv_kind = kind; j->in = in; j->out = out; j->produced = produced;
        j->tid = t; j->n_threads = n_threads;
        pthread_create(&th[t], NULL, bench_worker, j);
    }
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
