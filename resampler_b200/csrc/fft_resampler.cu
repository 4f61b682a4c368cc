// fft_resampler.cu -- `ResamplerFft` on the GPU: the reference's second
// resampler, overlap-add through a pair of FFTs (src/resampler_fft.rs; the approach follows Rubato).
//
// Per chunk of N_in frames and per channel (`FftResampler::resample`, :388-424): zero-pad to 2 N_in,
// forward real FFT, multiply the first `new_length` bins by the filter's spectrum, place them into a
// spectrum of N_out + 1 bins, inverse real FFT of 2 N_out points (unnormalised), add the previous
// chunk's second half, keep this chunk's second half.  N_in / N_out come from the reference's
// conversion table scaled to at least 512 input frames (src/fft/planner.rs:33-233): 1176 -> 1280 for
// 44.1 -> 48 kHz, 512 -> 1536 for 16 -> 48 kHz, ...
//
// Kernel: one CTA per (stream, chunk) -- the overlap-add is a sum of two terms, accumulated with
// atomics -- the channels in turn.  The real transforms run as half-size complex Stockham auto-sort
// FFTs in shared memory (mixed radix 8 / 4 / 2 / 3 / 5 / 7: every size of the table factors into
// these) with the usual pack / untangle steps, twiddles from tables computed in f64 on the host; the
// input chunk and the output chunk are staged so that global memory sees whole interleaved rows.
// The filter spectrum is the f64 DFT of the reference's f32 filter (:349-386), rounded to f32 once.
//
// This path is NOT near its roofline.  It exists to cover the reference's second
// public type with the same API and parity discipline; the FIR path is the product's hot path.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/resampler_b200.h"
#include "filter_design.h"

namespace {

thread_local std::string g_fft_error;
int fft_fail(int code, const std::string &msg) {
    g_fft_error = msg;
    return code;
}
#define FFT_CUDA(x)                                                                              \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess) return fft_fail(RSB_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
    } while (0)

constexpr int kMaxPasses = 8;
struct FftPlan {
    uint32_t n;
    uint32_t n_pass;
    uint32_t radix[kMaxPasses];
};

struct FftParams {
    FftPlan fwd, inv;
    const float2 *tw_fwd;      // exp(-2 pi i m / n_in), m < n_in: the forward transform's roots
    const float2 *tw_inv;      // exp(+2 pi i m / n_out), m < n_out: the inverse transform's roots
    const float2 *tw_post;     // exp(-2 pi i k / (2 n_in)), k <= n_in: real-input untangling
    const float2 *tw_pre;      // exp(+2 pi i k / (2 n_out)), k <= n_out: real-output tangling
    const float2 *filter;      // [n_in + 1]
    float *overlap;            // [streams][channels][n_out]
    uint32_t n_in, n_out, new_length, channels, buf_len;
};

struct FftJob {
    const float *in;
    float *out;
    const float *ov_read;      // the stream's overlap before this call  [channels][n_out]
    float *ov_write;           // ... after it (the other buffer)
    uint32_t stream;
    uint32_t chunks;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiplication by -i (forward) / +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 rot90(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

// R-point DFT of u in place.  2, 4 and 8 are the usual add / rotate networks; 3, 5 and 7 multiply by
// the roots W_R^m held in registers (wr).
template <int R, bool INV>
__device__ __forceinline__ void dft_small(float2 (&u)[R], const float2 (&wr)[R]) {
    if constexpr (R == 2) {
        const float2 a = u[0], b = u[1];
        u[0] = cadd(a, b);
        u[1] = csub(a, b);
    } else if constexpr (R == 4) {
        const float2 a0 = cadd(u[0], u[2]), a1 = csub(u[0], u[2]), a2 = cadd(u[1], u[3]), a3 = rot90<INV>(csub(u[1], u[3]));
        u[0] = cadd(a0, a2);
        u[1] = cadd(a1, a3);
        u[2] = csub(a0, a2);
        u[3] = csub(a1, a3);
    } else if constexpr (R == 8) {
        float2 e[4] = {u[0], u[2], u[4], u[6]}, o[4] = {u[1], u[3], u[5], u[7]};
        float2 dummy[4];
        dft_small<4, INV>(e, dummy);
        dft_small<4, INV>(o, dummy);
        const float h = 0.70710678118654752440f;
        // o[q] *= W_8^q
        const float2 o1 = INV ? make_float2(h * (o[1].x - o[1].y), h * (o[1].x + o[1].y))
                              : make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
        const float2 o2 = rot90<INV>(o[2]);
        const float2 o3 = INV ? make_float2(-h * (o[3].x + o[3].y), h * (o[3].x - o[3].y))
                              : make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
        u[0] = cadd(e[0], o[0]); u[4] = csub(e[0], o[0]);
        u[1] = cadd(e[1], o1);   u[5] = csub(e[1], o1);
        u[2] = cadd(e[2], o2);   u[6] = csub(e[2], o2);
        u[3] = cadd(e[3], o3);   u[7] = csub(e[3], o3);
    } else {
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            float2 acc = u[0];
#pragma unroll
            for (int r = 1; r < R; ++r) {
                const float2 w = wr[(q * r) % R];
                acc.x = fmaf(u[r].x, w.x, fmaf(-u[r].y, w.y, acc.x));
                acc.y = fmaf(u[r].x, w.y, fmaf(u[r].y, w.x, acc.y));
            }
            v[q] = acc;
        }
#pragma unroll
        for (int q = 0; q < R; ++q) u[q] = v[q];
    }
}

// One Stockham pass of radix R: x -> y, p = product of the radices already done.  `tw` is the
// transform's table W_n^m (global memory, L1 resident); the R-point DFT's own roots W_R^m sit in registers.
template <int R, bool INV>
__device__ void stockham_pass(const float2 *x, float2 *y, uint32_t n, uint32_t p, const float2 *tw) {
    const uint32_t t = n / R;
    const uint32_t tw_step = n / (p * R);       // W_{pR}^{k r} = W_n^{k r tw_step}
    float2 wr[R];
#pragma unroll
    for (int m = 0; m < R; ++m) wr[m] = (R == 3 || R == 5 || R == 7) ? __ldg(&tw[m * (n / R)]) : make_float2(0.f, 0.f);
    for (uint32_t i = threadIdx.x; i < t; i += blockDim.x) {
        const uint32_t k = i % p;
        const uint32_t j = (i / p) * p * R + k;
        const uint32_t kw = k * tw_step;
        float2 u[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float2 v = x[i + r * t];
            u[r] = (r == 0 || p == 1) ? v : cmul(v, __ldg(&tw[kw * r]));    // kw * r < n
        }
        dft_small<R, INV>(u, wr);
#pragma unroll
        for (int q = 0; q < R; ++q) y[j + q * p] = u[q];
    }
}

// Complex FFT over two shared buffers; returns the buffer holding the result.
template <bool INV>
__device__ float2 *fft_smem(float2 *a, float2 *b, const FftPlan &P, const float2 *tw) {
    uint32_t p = 1;
    for (uint32_t s = 0; s < P.n_pass; ++s) {
        const uint32_t R = P.radix[s];
        switch (R) {
            case 8: stockham_pass<8, INV>(a, b, P.n, p, tw); break;
            case 7: stockham_pass<7, INV>(a, b, P.n, p, tw); break;
            case 5: stockham_pass<5, INV>(a, b, P.n, p, tw); break;
            case 4: stockham_pass<4, INV>(a, b, P.n, p, tw); break;
            case 3: stockham_pass<3, INV>(a, b, P.n, p, tw); break;
            default: stockham_pass<2, INV>(a, b, P.n, p, tw); break;
        }
        __syncthreads();
        p *= R;
        float2 *tmp = a;
        a = b;
        b = tmp;
    }
    return a;
}

// One CTA per (job, chunk): the chunks of a stream are independent up to the overlap-add, which is a
// sum of exactly two terms per output value -- this chunk's first half and the previous chunk's
// second half (:419-423) -- so both are ADDED into a zeroed output with atomics (a + b == b + a: the
// result does not depend on which CTA comes first).  Chunk 0 also adds the overlap carried in the
// handle; the last chunk stores its second half as the new overlap (the other buffer of a pair, so a
// concurrent chunk 0 never reads what the last chunk writes).
__global__ void __launch_bounds__(256) fft_resample_kernel(const FftJob *jobs, const FftParams P) {
    extern __shared__ __align__(16) uint8_t sm_fft[];
    float2 *buf_a = reinterpret_cast<float2 *>(sm_fft);
    float2 *buf_b = buf_a + P.buf_len;
    float *s_in = reinterpret_cast<float *>(buf_b + P.buf_len);          // [n_in * channels]
    float *s_out = s_in + (size_t)P.n_in * P.channels;                    // [2 n_out * channels]: both halves
    const FftJob job = jobs[blockIdx.y];
    const uint32_t chunk = blockIdx.x;
    if (chunk >= job.chunks) return;
    const uint32_t ch = P.channels, n_in = P.n_in, n_out = P.n_out;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;

    const float *src = job.in + (size_t)chunk * n_in * ch;
    for (uint32_t i = tid; i < n_in * ch; i += nt) s_in[i] = src[i];              // :195-202, coalesced
    __syncthreads();
    for (uint32_t c = 0; c < ch; ++c) {
        // Real transforms through half-size complex ones.  Forward (:390-397): the zero-padded chunk
        // x[0 .. 2 n_in) is packed as z[n] = x[2n] + i x[2n+1], n < n_in (zero beyond n_in / 2).
        for (uint32_t n = tid; n < n_in; n += nt)
            buf_a[n] = 2 * n < n_in ? make_float2(s_in[(2 * n) * ch + c], s_in[(2 * n + 1) * ch + c])
                                    : make_float2(0.0f, 0.0f);
        __syncthreads();
        float2 *Z = fft_smem<false>(buf_a, buf_b, P.fwd, P.tw_fwd);
        float2 *Yb = Z == buf_a ? buf_b : buf_a;
        // untangle: X[k] = (Z[k] + conj(Z[M-k])) / 2 - (i / 2) W_2M^k (Z[k] - conj(Z[M-k])), k <= M;
        // times the filter, truncated / zero-extended to n_out + 1 bins (:399-411)
        for (uint32_t k = tid; k <= n_out; k += nt) {
            float2 v = make_float2(0.0f, 0.0f);
            if (k < P.new_length) {
                const float2 a = Z[k == n_in ? 0 : k], bq = Z[k == 0 ? 0 : n_in - k];
                const float2 b = make_float2(bq.x, -bq.y);
                const float2 sum = cadd(a, b), dif = csub(a, b);
                const float2 t = cmul(__ldg(&P.tw_post[k]), dif);            // W (Z[k] - conj(Z[M-k]))
                const float2 X = make_float2(0.5f * (sum.x + t.y), 0.5f * (sum.y - t.x));   // sum/2 - (i/2) t
                v = cmul(X, __ldg(&P.filter[k]));
            }
            if (k == 0 || k == n_out) v.y = 0.0f;
            Yb[k] = v;
        }
        __syncthreads();
        // tangle for the inverse (:413-417): Z'[k] = (Y[k] + conj(Y[M'-k])) + i W_2M'^{-k}... with the
        // inverse roots e^{+i pi k / M'}: Z'[k] = (Y[k] + conj(Y[M'-k])) + i e^{+i pi k / M'} (Y[k] - conj(Y[M'-k]))
        for (uint32_t k = tid; k < n_out; k += nt) {
            const float2 a = Yb[k], bq = Yb[n_out - k];
            const float2 b = make_float2(bq.x, -bq.y);
            const float2 sum = cadd(a, b), dif = csub(a, b);
            const float2 t = cmul(__ldg(&P.tw_pre[k]), dif);
            Z[k] = make_float2(sum.x - t.y, sum.y + t.x);                     // sum + i t
        }
        __syncthreads();
        float2 *z = fft_smem<true>(Z, Yb, P.inv, P.tw_inv);                      // z[n] = y[2n] + i y[2n+1]
        for (uint32_t n = tid; n < n_out; n += nt) {
            s_out[(2 * n) * ch + c] = z[n].x;
            s_out[(2 * n + 1) * ch + c] = z[n].y;
        }
        __syncthreads();
    }
    // overlap-add (:419-423) as two contributions
    float *dst = job.out + (size_t)chunk * n_out * ch;
    const bool first = chunk == 0, last = chunk + 1 == job.chunks;
    for (uint32_t i = tid; i < n_out * ch; i += nt) {
        float v = s_out[i];
        if (first) {                       // overlap layout in the handle: [channel][n_out]
            const uint32_t n = i / ch, c = i - n * ch;
            v += job.ov_read[(size_t)c * n_out + n];
        }
        atomicAdd(&dst[i], v);
        const float tail = s_out[(size_t)n_out * ch + i];
        if (last) {
            const uint32_t n = i / ch, c = i - n * ch;
            job.ov_write[(size_t)c * n_out + n] = tail;
        } else {
            atomicAdd(&dst[(size_t)n_out * ch + i], tail);
        }
    }
}

bool factorize(uint32_t n, FftPlan &P) {
    P.n = n;
    P.n_pass = 0;
    const uint32_t order[6] = {8, 4, 2, 3, 5, 7};
    for (uint32_t r : order)
        while (n % r == 0 && n > 1) {
            if (P.n_pass == kMaxPasses) return false;
            P.radix[P.n_pass++] = r;
            n /= r;
        }
    return n == 1;
}

// src/lib.rs:191-233 and src/fft/planner.rs:33-233
bool conversion_sizes(uint32_t in_hz, uint32_t out_hz, uint32_t &n_in, uint32_t &n_out) {
    auto family = [](uint32_t hz) -> uint32_t {
        switch (hz) {
            case 22050: case 44100: case 88200: case 176400: return 22050;
            case 16000: case 32000: return 16000;
            case 48000: case 96000: case 192000: case 384000: return 48000;
            default: return 0;
        }
    };
    const uint32_t fi = family(in_hz), fo = family(out_hz);
    if (!fi || !fo) return false;
    uint32_t bi = 2, bo = 2;
    if (fi == 22050 && fo == 48000) { bi = 588; bo = 1280; }
    else if (fi == 48000 && fo == 22050) { bi = 1280; bo = 588; }
    else if (fi == 16000 && fo == 48000) { bi = 64; bo = 192; }
    else if (fi == 48000 && fo == 16000) { bi = 192; bo = 64; }
    else if (fi == 16000 && fo == 22050) { bi = 640; bo = 882; }
    else if (fi == 22050 && fo == 16000) { bi = 882; bo = 640; }
    bi *= in_hz / fi;
    bo *= out_hz / fo;
    uint32_t mult = (uint32_t)std::fmax(std::ceil(512.0f / (float)bi), 1.0f);      // planner.rs:211-220
    uint32_t p2 = 1;
    while (p2 < mult) p2 <<= 1;                                                    // next_power_of_two
    n_in = bi * p2;
    n_out = bo * p2;
    return true;
}

}  // namespace

struct rsb_fft {
    int device = 0;
    uint32_t n_streams = 0, channels = 0, in_hz = 0, out_hz = 0, n_in = 0, n_out = 0, new_length = 0;
    FftParams P{};
    float2 *d_tw_fwd = nullptr, *d_tw_inv = nullptr, *d_filter = nullptr;
    float *d_overlap = nullptr;           // two buffers of [streams][channels][n_out]
    std::vector<uint8_t> ov_sel;          // which one is live, per stream
    cudaStream_t stream = nullptr;
    size_t smem = 0;
    // staging of host-memspace calls and the job table
    void *d_stage_in = nullptr, *d_stage_out = nullptr, *d_jobs = nullptr;
    size_t cap_in = 0, cap_out = 0, cap_jobs = 0;
    uint64_t launches = 0;
};

extern "C" {

const char *rsb_fft_last_error(void) { return g_fft_error.c_str(); }

int rsb_fft_create(rsb_fft **out, int device, uint32_t n_streams, uint32_t channels, uint32_t input_rate_hz,
                   uint32_t output_rate_hz) {
    if (!out) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (n_streams == 0 || channels == 0) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "n_streams and channels must be positive");
    uint32_t n_in = 0, n_out = 0;
    // ResamplerFft::new takes the SampleRate enum (resampler_fft.rs:75-79): the ten fixed rates only
    if (!conversion_sizes(input_rate_hz, output_rate_hz, n_in, n_out))
        return fft_fail(RSB_ERR_INVALID_ARGUMENT, "ResamplerFft supports the SampleRate enum's rates only");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fft_fail(RSB_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    }
    if (device < 0 || device >= count) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "bad device index");
    FFT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FFT_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fft_fail(RSB_ERR_NO_DEVICE, "an sm_100 device is required");
    rsb_fft *h = new rsb_fft;
    h->device = device;
    h->n_streams = n_streams;
    h->channels = channels;
    h->in_hz = input_rate_hz;
    h->out_hz = output_rate_hz;
    h->n_in = n_in;
    h->n_out = n_out;
    h->new_length = n_in < n_out ? n_in + 1 : n_out;                 // :399-402
    FftParams &P = h->P;
    if ((n_in & 1u) || (n_out & 1u) || !factorize(n_in, P.fwd) || !factorize(n_out, P.inv)) {
        delete h;
        return fft_fail(RSB_ERR_INVALID_ARGUMENT, "FFT size does not factor into 2, 3, 5, 7");
    }
    // ---- filter (:349-386): cutoff, periodic Kaiser sinc, scaled by 1 / (2 n_in), its spectrum ----
    const double kBeta = 10.0;
    const double cutoff = n_in > n_out ? rsb::kaiser_cutoff(n_out, kBeta) * ((double)n_out / (double)n_in)
                                       : rsb::kaiser_cutoff(n_in, kBeta);
    std::vector<float> sincs(n_in);
    rsb::make_sincs_for_kaiser(n_in, 1, (float)cutoff, kBeta, false, sincs.data());
    const uint32_t N1 = 2 * n_in, N2 = 2 * n_out;
    std::vector<float> ft(N1, 0.0f);
    for (uint32_t i = 0; i < n_in; ++i) ft[i] = sincs[i] / (float)(2 * n_in);
    const double kPi = 3.14159265358979323846;
    std::vector<double> cs(N1), sn(N1);
    for (uint32_t m = 0; m < N1; ++m) { cs[m] = std::cos(2.0 * kPi * m / N1); sn[m] = std::sin(2.0 * kPi * m / N1); }
    std::vector<float2> filt(n_in + 1), tw1(N1), tw2(N2);
    for (uint32_t k = 0; k <= n_in; ++k) {
        double re = 0.0, im = 0.0;
        for (uint32_t n = 0; n < n_in; ++n) {
            const uint32_t m = (uint32_t)(((uint64_t)k * n) % N1);
            re += (double)ft[n] * cs[m];
            im -= (double)ft[n] * sn[m];
        }
        filt[k] = make_float2((float)re, (float)im);
    }
    // tw1 / tw2: [roots of the half-size transform | untangling roots of the full size]
    std::vector<float2> twf(n_in), twi(n_out);
    tw1.resize(n_in + 1);
    tw2.resize(n_out + 1);
    for (uint32_t m = 0; m < n_in; ++m)
        twf[m] = make_float2((float)std::cos(2.0 * kPi * m / n_in), (float)-std::sin(2.0 * kPi * m / n_in));
    for (uint32_t m = 0; m < n_out; ++m)
        twi[m] = make_float2((float)std::cos(2.0 * kPi * m / n_out), (float)std::sin(2.0 * kPi * m / n_out));
    for (uint32_t k = 0; k <= n_in; ++k) tw1[k] = make_float2((float)cs[k], (float)-sn[k]);
    for (uint32_t k = 0; k <= n_out; ++k)
        tw2[k] = make_float2((float)std::cos(2.0 * kPi * k / N2), (float)std::sin(2.0 * kPi * k / N2));
    auto bail = [&](const char *what) {
        rsb_fft_destroy(h);
        return fft_fail(RSB_ERR_CUDA, what);
    };
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("stream");
    if (cudaMalloc(&h->d_tw_fwd, sizeof(float2) * (2 * n_in + 1)) != cudaSuccess ||
        cudaMalloc(&h->d_tw_inv, sizeof(float2) * (2 * n_out + 1)) != cudaSuccess ||
        cudaMalloc(&h->d_filter, sizeof(float2) * (n_in + 1)) != cudaSuccess ||
        cudaMalloc(&h->d_overlap, 2 * sizeof(float) * (size_t)n_streams * channels * n_out) != cudaSuccess)
        return bail("out of device memory");
    cudaMemcpy(h->d_tw_fwd, twf.data(), sizeof(float2) * n_in, cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_tw_fwd + n_in, tw1.data(), sizeof(float2) * (n_in + 1), cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_tw_inv, twi.data(), sizeof(float2) * n_out, cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_tw_inv + n_out, tw2.data(), sizeof(float2) * (n_out + 1), cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_filter, filt.data(), sizeof(float2) * (n_in + 1), cudaMemcpyHostToDevice);
    cudaMemset(h->d_overlap, 0, 2 * sizeof(float) * (size_t)n_streams * channels * n_out);
    h->ov_sel.assign(n_streams, 0);
    P.tw_fwd = h->d_tw_fwd;
    P.tw_inv = h->d_tw_inv;
    P.tw_post = h->d_tw_fwd + n_in;
    P.tw_pre = h->d_tw_inv + n_out;
    P.filter = h->d_filter;
    P.overlap = h->d_overlap;
    P.n_in = n_in;
    P.n_out = n_out;
    P.new_length = h->new_length;
    P.channels = channels;
    P.buf_len = std::max(n_in, n_out + 1);
    h->smem = sizeof(float2) * 2 * P.buf_len + sizeof(float) * (size_t)channels * (n_in + 2 * n_out);
    if (h->smem > 227 * 1024) return bail("channel count too large for this FFT size");
    if (cudaFuncSetAttribute(fft_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem) != cudaSuccess)
        return bail("shared memory");
    if (cudaGetLastError() != cudaSuccess) return bail("create");
    *out = h;
    return RSB_OK;
}

void rsb_fft_destroy(rsb_fft *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    cudaFree(h->d_tw_fwd);
    cudaFree(h->d_tw_inv);
    cudaFree(h->d_filter);
    cudaFree(h->d_overlap);
    cudaFree(h->d_stage_in);
    cudaFree(h->d_stage_out);
    cudaFree(h->d_jobs);
    delete h;
}

size_t rsb_fft_chunk_size_input(const rsb_fft *h) { return h ? (size_t)h->n_in * h->channels : 0; }      // :135-137
size_t rsb_fft_chunk_size_output(const rsb_fft *h) { return h ? (size_t)h->n_out * h->channels : 0; }    // :143-145
size_t rsb_fft_delay(const rsb_fft *h) { return h ? h->n_in / 2 : 0; }                                    // :151-153
uint64_t rsb_fft_launch_count(const rsb_fft *h) { return h ? h->launches : 0; }

int rsb_fft_reset(rsb_fft *h, int64_t stream) {
    if (!h) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "null handle");
    FFT_CUDA(cudaSetDevice(h->device));
    const size_t per = (size_t)h->channels * h->n_out;
    const size_t half = per * h->n_streams;
    if (stream < 0) FFT_CUDA(cudaMemsetAsync(h->d_overlap, 0, 2 * sizeof(float) * half, h->stream));
    else if ((uint64_t)stream < h->n_streams) {
        FFT_CUDA(cudaMemsetAsync(h->d_overlap + per * (size_t)stream, 0, sizeof(float) * per, h->stream));
        FFT_CUDA(cudaMemsetAsync(h->d_overlap + half + per * (size_t)stream, 0, sizeof(float) * per, h->stream));
    } else return fft_fail(RSB_ERR_INVALID_ARGUMENT, "bad stream index");
    return RSB_OK;
}

int rsb_fft_sync(rsb_fft *h) {
    if (!h) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "null handle");
    FFT_CUDA(cudaSetDevice(h->device));
    FFT_CUDA(cudaStreamSynchronize(h->stream));
    return RSB_OK;
}

// Whole chunks per job: chunks_done[i] = min(in_lens[i] / chunk_size_input, out_lens[i] / chunk_size_output)
// consecutive `resample()` calls of the reference (:182-246), state (the overlap) carried in the handle.
int rsb_fft_process_batch(rsb_fft *h, uint32_t n, const uint32_t *streams, const float *const *in, const size_t *in_lens,
                          float *const *out, const size_t *out_lens, size_t *chunks_done, int memspace, uint32_t flags) {
    if (!h) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "null handle");
    if (n == 0) return RSB_OK;
    if (!in || !in_lens || !out || !out_lens) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "null array argument");
    if (memspace != RSB_MEM_DEVICE && memspace != RSB_MEM_HOST) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "bad memspace");
    FFT_CUDA(cudaSetDevice(h->device));
    const size_t csi = rsb_fft_chunk_size_input(h), cso = rsb_fft_chunk_size_output(h);
    std::vector<FftJob> jobs(n);
    std::vector<uint8_t> seen(h->n_streams, 0);
    std::vector<size_t> off_in(n), off_out(n);
    size_t tot_in = 0, tot_out = 0;
    uint32_t max_chunks = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t s = streams ? streams[i] : i;
        if (s >= h->n_streams) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "bad stream index");
        if (seen[s]) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "stream listed twice in one batch");
        seen[s] = 1;
        const size_t c = std::min(in_lens[i] / csi, out_lens[i] / cso);
        if (c > 0xffffffffull) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "too many chunks");
        if (c && (!in[i] || !out[i])) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "null buffer");
        const size_t per = (size_t)h->channels * h->n_out, half = per * h->n_streams;
        const uint32_t sel = h->ov_sel[s];
        jobs[i] = FftJob{in[i], out[i], h->d_overlap + sel * half + per * s, h->d_overlap + (sel ^ 1u) * half + per * s,
                         s, (uint32_t)c};
        max_chunks = std::max<uint32_t>(max_chunks, (uint32_t)c);
        if (chunks_done) chunks_done[i] = c;
        off_in[i] = tot_in;
        off_out[i] = tot_out;
        tot_in += c * csi;
        tot_out += c * cso;
    }
    auto reserve = [](void *&p, size_t &cap, size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const cudaError_t e = cudaMalloc(&p, bytes + bytes / 4 + 256);
        if (e == cudaSuccess) cap = bytes + bytes / 4 + 256;
        return e;
    };
    if (memspace == RSB_MEM_HOST) {
        FFT_CUDA(reserve(h->d_stage_in, h->cap_in, tot_in * sizeof(float)));
        FFT_CUDA(reserve(h->d_stage_out, h->cap_out, tot_out * sizeof(float)));
        for (uint32_t i = 0; i < n; ++i) {
            if (!jobs[i].chunks) continue;
            FFT_CUDA(cudaMemcpyAsync(static_cast<float *>(h->d_stage_in) + off_in[i], in[i], jobs[i].chunks * csi * sizeof(float),
                                     cudaMemcpyHostToDevice, h->stream));
            jobs[i].in = static_cast<float *>(h->d_stage_in) + off_in[i];
            jobs[i].out = static_cast<float *>(h->d_stage_out) + off_out[i];
        }
    }
    FFT_CUDA(reserve(h->d_jobs, h->cap_jobs, sizeof(FftJob) * n));
    // the job table is read by the kernel after this call may have returned (RSB_FLAG_ASYNC).  `jobs` is
    // pageable memory: cudaMemcpyAsync returns once such a source has been staged for the DMA, so
    // the vector may die with this call; the copy itself is ordered on the stream before the kernel
    FFT_CUDA(cudaMemcpyAsync(h->d_jobs, jobs.data(), sizeof(FftJob) * n, cudaMemcpyHostToDevice, h->stream));
    if (max_chunks) {
        // the chunks' contributions are accumulated into the output: it starts from zero
        for (uint32_t i = 0; i < n; ++i)
            if (jobs[i].chunks)
                FFT_CUDA(cudaMemsetAsync(jobs[i].out, 0, jobs[i].chunks * cso * sizeof(float), h->stream));
        if (max_chunks > 0x7fffffffu || n > 65535u) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "batch too large; split it");
        fft_resample_kernel<<<dim3(max_chunks, n), (getenv("RSB_FFT_THREADS") ? atoi(getenv("RSB_FFT_THREADS")) : 256), h->smem, h->stream>>>(static_cast<const FftJob *>(h->d_jobs), h->P);
        FFT_CUDA(cudaGetLastError());
        h->launches += 1;
        for (uint32_t i = 0; i < n; ++i)
            if (jobs[i].chunks) h->ov_sel[jobs[i].stream] ^= 1u;
    }
    if (memspace == RSB_MEM_HOST) {
        for (uint32_t i = 0; i < n; ++i)
            if (jobs[i].chunks)
                FFT_CUDA(cudaMemcpyAsync(out[i], static_cast<float *>(h->d_stage_out) + off_out[i],
                                         jobs[i].chunks * cso * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        FFT_CUDA(cudaStreamSynchronize(h->stream));
    } else if (!(flags & RSB_FLAG_ASYNC)) {
        FFT_CUDA(cudaStreamSynchronize(h->stream));
    }
    return RSB_OK;
}

// One chunk of one stream, host slices: `ResamplerFft::resample` (:182-246) incl. its error precedence
int rsb_fft_resample(rsb_fft *h, uint32_t stream, const float *input, size_t input_len, float *output, size_t output_len) {
    if (!h) return fft_fail(RSB_ERR_INVALID_ARGUMENT, "null handle");
    if (input_len < rsb_fft_chunk_size_input(h)) return fft_fail(RSB_ERR_INVALID_INPUT_BUFFER_SIZE, "input shorter than chunk_size_input()");
    if (output_len < rsb_fft_chunk_size_output(h)) return fft_fail(RSB_ERR_INVALID_OUTPUT_BUFFER_SIZE, "output shorter than chunk_size_output()");
    const size_t il = rsb_fft_chunk_size_input(h), ol = rsb_fft_chunk_size_output(h);
    return rsb_fft_process_batch(h, 1, &stream, &input, &il, &output, &ol, nullptr, RSB_MEM_HOST, 0);
}

}  // extern "C"
