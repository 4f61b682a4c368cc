/*
 * fir_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see fir_oracle.h).
 *
 * Restates hasenbanck/resampler v0.5.1; every function cites the reference
 * lines it follows.  Compile with -ffp-contract=off: the reference's blend
 * `acc1*(1-f) + acc2*f` is three separately rounded operations and its f32
 * filter design must stay in f32 without fused contractions.
 */
#include "fir_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* window.rs                                                           */
/* ------------------------------------------------------------------ */

/* window.rs:96-112.  Power series of I0; term_k = term_{k-1} * (x^2/4) / k^2,
 * stop as soon as the sum stops changing, at most 1499 terms. */
double orc_bessel_i0(double x) {
    double base = x * x / 4.0;
    double term = 1.0;
    double result = 1.0;
    for (int idx = 1; idx < 1500; ++idx) {
        term = term * base / (double)(idx * idx);
        double previous = result;
        result += term;
        if (result == previous) break;
    }
    return result;
}

/* window.rs:66-94.  Periodic: x/(N/2)-1; Symmetric: 2x/(N-1)-1; f64 until the
 * final `as f32`.  powi(2) is x*x. */
void orc_kaiser_window(size_t n, double beta, int symmetric, float *out) {
    double bessel_beta = orc_bessel_i0(beta);
    for (size_t index = 0; index < n; ++index) {
        double x = (double)index;
        double nx;
        if (symmetric)
            nx = 2.0 * x / (double)(n - 1) - 1.0;
        else
            nx = x / ((double)n / 2.0) - 1.0;
        double value = orc_bessel_i0(beta * sqrt(1.0 - nx * nx)) / bessel_beta;
        out[index] = (float)value;
    }
}

/* window.rs:114-131 */
double orc_cutoff_kaiser(size_t sample_count, double beta) {
    double n = (double)sample_count;
    double a_db = beta / 0.1102 + 8.7;
    double delta_f_nyquist = (a_db - 7.95) / (14.36 * n);
    const double SAFETY_MARGIN = 1.005;
    double cutoff = 1.0 - (delta_f_nyquist * SAFETY_MARGIN);
    if (cutoff < 0.7) cutoff = 0.7;
    if (cutoff > 1.0) cutoff = 1.0;
    return cutoff;
}

/* window.rs:29-37: f32 sinc with f32 PI and the platform sinf. */
static float orc_sinc(float value) {
    const float PI_F32 = 3.14159274101257324219f; /* core::f32::consts::PI */
    if (value == 0.0f) return 1.0f;
    float vp = value * PI_F32;
    return sinf(vp) / vp;
}

/* window.rs:17-55.  y[x] = w[x]*sinc((x - N/2) as f32 * cutoff / factor as f32)
 * in f32; `sum` is a sequential f32 accumulation; sum /= factor;
 * sincs[factor-n-1][p] = y[factor*p+n] / sum. */
void orc_make_sincs(size_t sample_count, size_t factor, float f_cutoff, double beta,
                    int symmetric, float *out, float *sum_out) {
    size_t totpoints = sample_count * factor;
    float *y = (float *)malloc(totpoints * sizeof(float));
    float *window = (float *)malloc(totpoints * sizeof(float));
    orc_kaiser_window(totpoints, beta, symmetric, window);
    volatile float sum = 0.0f; /* volatile: forbid any reassociation/vectorisation */
    for (size_t x = 0; x < totpoints; ++x) {
        float arg = (float)((int32_t)x - (int32_t)(totpoints / 2)) * f_cutoff / (float)factor;
        float val = window[x] * orc_sinc(arg);
        sum = sum + val;
        y[x] = val;
    }
    float s = sum;
    s /= (float)factor;
    for (size_t p = 0; p < sample_count; ++p)
        for (size_t n = 0; n < factor; ++n)
            out[(factor - n - 1) * sample_count + p] = y[factor * p + n] / s;
    if (sum_out) *sum_out = s;
    free(window);
    free(y);
}

/* resampler_fir.rs:117-123 */
double orc_attenuation_beta(int attenuation) {
    switch (attenuation) {
        case 0: return 7.0;   /* Db60 */
        case 1: return 10.0;  /* Db90 */
        default: return 13.0; /* Db120 */
    }
}

/* resampler_fir.rs:153-161 */
int orc_latency_taps(int latency) {
    switch (latency) {
        case 0: return 16;   /* Sample8 */
        case 1: return 32;   /* Sample16 */
        case 2: return 64;   /* Sample32 */
        default: return 128; /* Sample64 */
    }
}

/* resampler_fir.rs:311-326: f64 ratio/cutoff, `as f32` only at the end. */
float orc_cutoff_for(uint32_t in_hz, uint32_t out_hz, int taps, double beta) {
    double in_r = (double)in_hz, out_r = (double)out_hz;
    double base_cutoff = orc_cutoff_kaiser((size_t)taps, beta);
    double cutoff = (in_r <= out_r) ? base_cutoff : base_cutoff * (out_r / in_r);
    return (float)cutoff;
}

/* ------------------------------------------------------------------ */
/* fir/avx512.rs and fir/mod.rs                                        */
/* ------------------------------------------------------------------ */

/* fir/avx512.rs:5-50 written lane by lane: 16 lanes, FMA chains over
 * taps/16 steps for both phases (:36-37), unfused blend with `1.0 - frac`
 * computed in f32 (:41-45), then _mm512_reduce_add_ps (:48) = halving tree
 * (l,l+8) -> (l,l+4) -> (l,l+2) -> (0,1) as Rust's stdarch defines it. */
float orc_convolve_avx512_order(const float *x, const float *c1, const float *c2, float frac,
                                size_t taps) {
    float acc1[16], acc2[16], v[16];
    for (int l = 0; l < 16; ++l) acc1[l] = acc2[l] = 0.0f;
    size_t iters = taps / 16;
    for (size_t i = 0; i < iters; ++i) {
        for (int l = 0; l < 16; ++l) {
            float xv = x[i * 16 + l];
            acc1[l] = fmaf(c1[i * 16 + l], xv, acc1[l]);
            acc2[l] = fmaf(c2[i * 16 + l], xv, acc2[l]);
        }
    }
    float omf = 1.0f - frac;
    for (int l = 0; l < 16; ++l) {
        float w1 = acc1[l] * omf;
        float w2 = acc2[l] * frac;
        v[l] = w1 + w2;
    }
    for (int l = 0; l < 8; ++l) v[l] = v[l] + v[l + 8];
    for (int l = 0; l < 4; ++l) v[l] = v[l] + v[l + 4];
    for (int l = 0; l < 2; ++l) v[l] = v[l] + v[l + 2];
    return v[0] + v[1];
}

/* fir/mod.rs:47-62: unfused mul/add, single chains. */
float orc_convolve_scalar(const float *x, const float *c1, const float *c2, float frac,
                          size_t taps) {
    float sum1 = 0.0f, sum2 = 0.0f;
    for (size_t i = 0; i < taps; ++i) {
        float xv = x[i];
        float p1 = c1[i] * xv;
        sum1 = sum1 + p1;
        float p2 = c2[i] * xv;
        sum2 = sum2 + p2;
    }
    float a = sum1 * (1.0f - frac);
    float b = sum2 * frac;
    return a + b;
}

/* ------------------------------------------------------------------ */
/* resampler_fir.rs                                                    */
/* ------------------------------------------------------------------ */

/* resampler_fir.rs:425-443: process-wide table cache keyed by
 * (cutoff.to_bits(), taps, attenuation) (:91-95).  Entries live forever, like
 * the reference's LazyLock<Mutex<HashMap>>. */
typedef struct { uint32_t cutoff_bits; size_t taps; int attenuation; float *coeffs; } orc_cache_entry;
static orc_cache_entry g_cache[64];
static int g_cache_n = 0;
static pthread_mutex_t g_cache_mu = PTHREAD_MUTEX_INITIALIZER;

static float *orc_get_or_create_coeffs(float cutoff, size_t taps, int attenuation) {
    uint32_t bits;
    memcpy(&bits, &cutoff, 4);
    pthread_mutex_lock(&g_cache_mu);
    for (int i = 0; i < g_cache_n; ++i)
        if (g_cache[i].cutoff_bits == bits && g_cache[i].taps == taps &&
            g_cache[i].attenuation == attenuation) {
            float *c = g_cache[i].coeffs;
            pthread_mutex_unlock(&g_cache_mu);
            return c;
        }
    float *coeffs = (float *)aligned_alloc(64, ORC_PHASES * taps * sizeof(float));
    orc_make_sincs(taps, ORC_PHASES, cutoff, orc_attenuation_beta(attenuation), 1, coeffs,
                   NULL);                                         /* :406-421 */
    if (g_cache_n < 64) {
        g_cache[g_cache_n].cutoff_bits = bits;
        g_cache[g_cache_n].taps = taps;
        g_cache[g_cache_n].attenuation = attenuation;
        g_cache[g_cache_n].coeffs = coeffs;
        g_cache_n++;
    }
    pthread_mutex_unlock(&g_cache_mu);
    return coeffs;
}

/* resampler_fir.rs:295-404 */
orc_fir *orc_fir_new(size_t channels, uint32_t in_hz, uint32_t out_hz, int latency,
                     int attenuation) {
    if (in_hz == 0 || out_hz == 0) return NULL; /* reference panics, :302-309 */
    orc_fir *r = (orc_fir *)calloc(1, sizeof(orc_fir));
    r->channels = channels;
    r->ratio = (double)in_hz / (double)out_hz;                    /* :311-313 */
    r->taps = (size_t)orc_latency_taps(latency);                  /* :315 */
    double beta = orc_attenuation_beta(attenuation);              /* :316 */
    float cutoff = orc_cutoff_for(in_hz, out_hz, (int)r->taps, beta); /* :317-326 */
    r->coeffs = orc_get_or_create_coeffs(cutoff, r->taps, attenuation < 0 ? 0 : (attenuation > 2 ? 2 : attenuation));
    r->input_buffers = (float *)calloc(ORC_BUFFER_SIZE * channels, sizeof(float)); /* :329 */
    r->read_position = 0;
    r->available_frames = 0;
    r->position = 0.0;
    r->conv_kind = ORC_CONV_AVX512; /* runtime pick on an avx512f host, :332 */
    return r;
}

void orc_fir_free(orc_fir *r) {
    if (!r) return;
    free(r->input_buffers); /* coeffs belong to the cache */
    free(r);
}

void orc_fir_set_conv(orc_fir *r, int conv_kind) { r->conv_kind = conv_kind; }

/* resampler_fir.rs:456-465 */
size_t orc_fir_buffer_size_output(const orc_fir *r) {
    double max_usable_frames = (double)(ORC_INPUT_CAPACITY - r->taps);
    size_t max_output_frames = (size_t)ceil(max_usable_frames / r->ratio) + 2;
    return max_output_frames * r->channels;
}

size_t orc_fir_delay(const orc_fir *r) { return r->taps / 2; }    /* :630-632 */

void orc_fir_reset(orc_fir *r) {                                  /* :638-642 */
    r->read_position = 0;
    r->available_frames = 0;
    r->position = 0.0;
}

static size_t min_sz(size_t a, size_t b) { return a < b ? a : b; }

/* resampler_fir.rs:509-621 */
int orc_fir_resample_traced(orc_fir *r, const float *input, size_t in_len, float *output,
                            size_t out_len, size_t *consumed, size_t *produced,
                            orc_trace *trace) {
    const size_t ch = r->channels, taps = r->taps;
    if (in_len % ch != 0) return ORC_ERR_INPUT_SIZE;              /* :514-516 */
    if (out_len % ch != 0) return ORC_ERR_OUTPUT_SIZE;            /* :517-519 */

    size_t input_frames = in_len / ch;
    size_t output_capacity = out_len / ch;

    size_t write_position = r->read_position + r->available_frames;           /* :524 */
    size_t remaining_capacity =
        ORC_BUFFER_SIZE > write_position ? ORC_BUFFER_SIZE - write_position : 0; /* :525 */
    size_t frames_to_copy = min_sz(min_sz(input_frames, remaining_capacity),
                                   ORC_INPUT_CAPACITY - r->available_frames);  /* :526-528 */

    for (size_t f = 0; f < frames_to_copy; ++f)                                /* :531-537 */
        for (size_t c = 0; c < ch; ++c)
            r->input_buffers[ORC_BUFFER_SIZE * c + write_position + f] = input[f * ch + c];
    r->available_frames += frames_to_copy;                                     /* :538 */

    size_t output_frame_count = 0;
    if (trace) trace->count = 0;
    for (;;) {
        size_t input_offset = (size_t)floor(r->position);                      /* :544 */
        if (input_offset + taps > r->available_frames) break;                  /* :549 */
        if (output_frame_count >= output_capacity) break;                      /* :553 */

        double position_fract = r->position - trunc(r->position);  /* f64::fract, :557 */
        double phase_f = position_fract * (double)ORC_PHASES;                  /* :562 */
        if (phase_f > (double)(ORC_PHASES - 1)) phase_f = (double)(ORC_PHASES - 1);
        size_t phase1 = (size_t)phase_f;                                       /* :563 */
        size_t phase2 = min_sz(phase1 + 1, ORC_PHASES - 1);                    /* :564 */
        float frac = (float)(phase_f - (double)phase1);                        /* :565 */

        if (trace && trace->count < trace->capacity) {
            size_t t = trace->count++;
            uint32_t fb;
            memcpy(&fb, &frac, 4);
            if (trace->input_offset) trace->input_offset[t] = (uint32_t)input_offset;
            if (trace->phase1) trace->phase1[t] = (uint32_t)phase1;
            if (trace->phase2) trace->phase2[t] = (uint32_t)phase2;
            if (trace->frac_bits) trace->frac_bits[t] = fb;
        }

        for (size_t c = 0; c < ch; ++c) {                                      /* :567-586 */
            size_t actual_pos = r->read_position + input_offset;
            const float *x = r->input_buffers + ORC_BUFFER_SIZE * c + actual_pos;
            const float *c1 = r->coeffs + phase1 * taps;
            const float *c2 = r->coeffs + phase2 * taps;
            float s;
            if (r->conv_kind == ORC_CONV_AVX512_INTRIN)
                s = orc_convolve_avx512_intrin(x, c1, c2, frac, taps);
            else if (r->conv_kind == ORC_CONV_SCALAR)
                s = orc_convolve_scalar(x, c1, c2, frac, taps);
            else
                s = orc_convolve_avx512_order(x, c1, c2, frac, taps);
            output[output_frame_count * ch + c] = s;
        }
        output_frame_count += 1;                                               /* :588 */
        r->position += r->ratio;                                               /* :589 */
    }

    size_t consumed_frames = min_sz((size_t)floor(r->position), r->available_frames); /* :596 */
    r->read_position += consumed_frames;                                       /* :600 */
    r->available_frames -= consumed_frames;                                    /* :601 */
    r->position -= (double)consumed_frames;                                    /* :602 */

    if (r->read_position > ORC_INPUT_CAPACITY) {                               /* :605-615 */
        for (size_t c = 0; c < ch; ++c) {
            float *buf = r->input_buffers + ORC_BUFFER_SIZE * c;
            memmove(buf, buf + r->read_position, r->available_frames * sizeof(float));
        }
        r->read_position = 0;
    }

    *consumed = frames_to_copy * ch;                                           /* :617-620 */
    *produced = output_frame_count * ch;
    return ORC_OK;
}

int orc_fir_resample(orc_fir *r, const float *input, size_t in_len, float *output,
                     size_t out_len, size_t *consumed, size_t *produced) {
    return orc_fir_resample_traced(r, input, in_len, output, out_len, consumed, produced, NULL);
}

/* The canonical caller loop: resample/src/main.rs:226-254, and the reference's
 * test helper resampler_fir.rs:719-735. */
size_t orc_fir_process(orc_fir *r, const float *in, size_t total_len, size_t call_len,
                       size_t out_cap_len, float *out, size_t out_capacity, size_t *out_total,
                       size_t *in_total, uint32_t *consumed_calls, uint32_t *produced_calls,
                       size_t max_calls, orc_trace *trace) {
    size_t cap = out_cap_len ? out_cap_len : orc_fir_buffer_size_output(r);
    float *buf = (float *)malloc((cap ? cap : 1) * sizeof(float));
    size_t offset = 0, total = 0, calls = 0;
    orc_trace sub;
    size_t tr_used = 0;
    while (offset < total_len) {
        size_t remaining = total_len - offset;
        size_t chunk = remaining < call_len ? remaining : call_len;
        size_t c = 0, p = 0;
        orc_trace *tp = NULL;
        if (trace) {
            sub.input_offset = trace->input_offset ? trace->input_offset + tr_used : NULL;
            sub.phase1 = trace->phase1 ? trace->phase1 + tr_used : NULL;
            sub.phase2 = trace->phase2 ? trace->phase2 + tr_used : NULL;
            sub.frac_bits = trace->frac_bits ? trace->frac_bits + tr_used : NULL;
            sub.capacity = trace->capacity - tr_used;
            sub.count = 0;
            tp = &sub;
        }
        int err = orc_fir_resample_traced(r, in + offset, chunk, buf, cap, &c, &p, tp);
        if (err) break;
        if (tp) tr_used += sub.count;
        size_t room = out_capacity - total;
        size_t ncopy = p < room ? p : room;
        if (out && ncopy) memcpy(out + total, buf, ncopy * sizeof(float));
        total += p;
        if (calls < max_calls) {
            if (consumed_calls) consumed_calls[calls] = (uint32_t)c;
            if (produced_calls) produced_calls[calls] = (uint32_t)p;
        }
        calls += 1;
        offset += c;
        if (c == 0) break;
    }
    if (trace) trace->count = tr_used;
    free(buf);
    if (out_total) *out_total = total;
    if (in_total) *in_total = offset;
    return calls;
}

/* ---- CLI format step: resample/src/main.rs:128-156 ---- */
void orc_pcm_to_f32(const void *src, int format, size_t n_src, size_t dup, float *dst) {
    const uint8_t *b = (const uint8_t *)src;
    for (size_t i = 0; i < n_src; ++i) {
        float x;
        if (format == 4) {                    /* SampleFormat::Float, main.rs:129 */
            uint32_t u = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) |
                         ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
            memcpy(&x, &u, 4);
        } else {                              /* SampleFormat::Int, main.rs:130-136 */
            int32_t s;
            int bits;
            if (format == 0) {                /* hound: u8 on disk, minus 128 */
                s = (int32_t)b[i] - 128;
                bits = 8;
            } else if (format == 1) {
                s = (int16_t)((uint16_t)b[2 * i] | ((uint16_t)b[2 * i + 1] << 8));
                bits = 16;
            } else if (format == 2) {
                uint32_t u = (uint32_t)b[3 * i] | ((uint32_t)b[3 * i + 1] << 8) |
                             ((uint32_t)b[3 * i + 2] << 16);
                s = (u & 0x800000u) ? (int32_t)(u | 0xff000000u) : (int32_t)u;
                bits = 24;
            } else {
                uint32_t u = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) |
                             ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
                s = (int32_t)u;
                bits = 32;
            }
            /* `(1 << (bits - 1)) as f32`: the literal is an i32, and for 32-bit samples
             * `1i32 << 31` is i32::MIN (Rust's shift only panics for amounts >= 32), so
             * max_value = -2^31 and the CLI yields -s / 2^31: 32-bit integer files come out
             * with inverted polarity.  The restatement keeps that. */
            float max_value = bits == 32 ? -2147483648.0f : (float)(1 << (bits - 1));
            x = (float)s / max_value;         /* `s as f32 / max_value`, main.rs:135 */
        }
        for (size_t d = 0; d < dup; ++d) dst[i * dup + d] = x;   /* main.rs:141-146 */
    }
}
