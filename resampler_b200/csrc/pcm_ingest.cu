// pcm_ingest.cu -- the format step in front of the FIR path, on the device (sm_100a).
//
// Reference: the CLI's batch path, resample/src/main.rs:128-156 --
//   * integer PCM -> f32:  `s as f32 / (1 << (bits_per_sample - 1)) as f32`  (:131-136; the
//     samples arrive as i32 from hound 3.5's `samples::<i32>()`, which turns 8-bit WAV bytes
//     into signed values by subtracting 128),
//   * mono -> stereo: every sample pushed twice (:139-146), stereo passes through (:148-150).
// Here both are one pass over HBM: raw samples are read once with vector loads, converted and
// written once as the interleaved f32 frames the convolution kernels consume.  A division by a
// power of two is exact in binary floating point, so `v * 2^-(bits-1)` equals the reference's
// division bit for bit; `i32 as f32` rounds to nearest even like `__int2float_rn`.
//
// HBM-bound byte work: every thread produces one 16-byte store per step (a warp writes 512
// contiguous bytes), reads the 4 / dup source values behind it with one vector load, keeps eight
// independent steps in flight, and the grid is a multiple of the SM count walking 32 KB output
// chunks of all jobs in order.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/resampler_b200.h"
#include "pcm_ingest.h"

namespace rsb {

namespace {

constexpr uint32_t kChunkVec = 2048;   // float4 stores per chunk (32 KB of output)
constexpr uint32_t kThreads = 256;

template <int FMT> struct Src;
template <> struct Src<RSB_PCM_U8> {
    using T = uint8_t;
    static __device__ __forceinline__ float cv(T v) { return (float)((int)v - 128) * (1.0f / 128.0f); }
};
template <> struct Src<RSB_PCM_S16> {
    using T = int16_t;
    static __device__ __forceinline__ float cv(T v) { return (float)(int)v * (1.0f / 32768.0f); }
};
template <> struct Src<RSB_PCM_S32> {
    using T = int32_t;
    static __device__ __forceinline__ float cv(T v) {
        // main.rs:131: `(1 << 31) as f32` is an i32 shift = i32::MIN = -2^31, so the reference
        // divides 32-bit samples by a NEGATIVE full scale (inverted polarity); kept as is
        return __int2float_rn(v) * (-1.0f / 2147483648.0f);
    }
};
template <> struct Src<RSB_PCM_F32> {
    using T = float;
    static __device__ __forceinline__ float cv(T v) { return v; }
};

// packed little-endian 24-bit samples
__device__ __forceinline__ float cv_s24(const uint8_t *p) {
    const int v = (int)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16));
    return (float)((v << 8) >> 8) * (1.0f / 8388608.0f);
}

// ---- one 16-byte output piece = output values [4q, 4q + 4); source index = output index / dup.
// A piece is fetched RAW (the 4 / dup source samples packed in at most four 32-bit registers)
// and converted only when it is stored, so the loads a thread keeps in flight cost as few
// registers as the format allows (s16 stereo: 2 per piece instead of 4).
template <int FMT, int DUP>
struct Piece {
    using T = typename Src<FMT>::T;
    static constexpr int kCnt = 4 / DUP;                      // source samples per piece
    static constexpr int kBytes = kCnt * (int)sizeof(T);       // 2, 4, 8 or 16
    static constexpr int kWords = (kBytes + 3) / 4;
    struct Raw { uint32_t r[kWords]; };

    static __device__ __forceinline__ Raw ld(const void *src, uint64_t q, bool aligned, uint32_t) {
        const T *s = static_cast<const T *>(src) + (uint64_t)kCnt * q;
        Raw v;
        if (aligned) {
            if constexpr (kBytes == 2) v.r[0] = *reinterpret_cast<const uint16_t *>(s);
            else if constexpr (kBytes == 4) v.r[0] = *reinterpret_cast<const uint32_t *>(s);
            else if constexpr (kBytes == 8) {
                const uint2 w = *reinterpret_cast<const uint2 *>(s);
                v.r[0] = w.x; v.r[1] = w.y;
            } else {
                const uint4 w = *reinterpret_cast<const uint4 *>(s);
                v.r[0] = w.x; v.r[1] = w.y; v.r[2] = w.z; v.r[3] = w.w;
            }
        } else {
#pragma unroll
            for (int w = 0; w < kWords; ++w) v.r[w] = 0;
#pragma unroll
            for (int k = 0; k < kCnt; ++k) {
                uint32_t e;
                if constexpr (sizeof(T) == 4) e = reinterpret_cast<const uint32_t *>(s)[k];
                else if constexpr (sizeof(T) == 2) e = reinterpret_cast<const uint16_t *>(s)[k];
                else e = reinterpret_cast<const uint8_t *>(s)[k];
                v.r[k * (int)sizeof(T) / 4] |= e << (8 * ((k * (int)sizeof(T)) % 4));
            }
        }
        return v;
    }
    static __device__ __forceinline__ float elem(const Raw &v, int k) {
        const uint32_t w = v.r[k * (int)sizeof(T) / 4] >> (8 * ((k * (int)sizeof(T)) % 4));
        if constexpr (FMT == RSB_PCM_F32) return __uint_as_float(w);
        else if constexpr (sizeof(T) == 4) return Src<FMT>::cv((T)w);
        else if constexpr (sizeof(T) == 2) return Src<FMT>::cv((T)(uint16_t)w);
        else return Src<FMT>::cv((T)(uint8_t)w);
    }
    static __device__ __forceinline__ float4 cv(const Raw &v) {
        if constexpr (DUP == 1) return make_float4(elem(v, 0), elem(v, 1), elem(v, 2), elem(v, 3));
        else {
            const float a = elem(v, 0), b = elem(v, 1);
            return make_float4(a, a, b, b);
        }
    }
};

// generic duplication factor (mono into 4 / 8 channels ...): converted on load
template <int FMT>
struct Piece<FMT, 0> {
    using T = typename Src<FMT>::T;
    struct Raw { float4 f; };
    static __device__ __forceinline__ Raw ld(const void *src, uint64_t q, bool, uint32_t dup) {
        const T *s = static_cast<const T *>(src);
        const uint64_t o = 4 * q;
        return Raw{make_float4(Src<FMT>::cv(s[o / dup]), Src<FMT>::cv(s[(o + 1) / dup]),
                               Src<FMT>::cv(s[(o + 2) / dup]), Src<FMT>::cv(s[(o + 3) / dup]))};
    }
    static __device__ __forceinline__ float4 cv(const Raw &v) { return v.f; }
};

__device__ __forceinline__ float cv_s24_bits(uint32_t v) {
    return (float)((int)(v << 8) >> 8) * (1.0f / 8388608.0f);
}

// packed 24-bit: an aligned stereo piece is 12 bytes = three words; everything else goes
// through byte loads and is converted on load
template <>
struct Piece<RSB_PCM_S24, 1> {
    struct Raw { uint32_t r[3]; };
    static __device__ __forceinline__ Raw ld(const void *src, uint64_t q, bool aligned, uint32_t) {
        const uint8_t *s = static_cast<const uint8_t *>(src) + 12 * q;
        Raw v;
        if (aligned) {
            const uint32_t *w = reinterpret_cast<const uint32_t *>(s);
            v.r[0] = w[0]; v.r[1] = w[1]; v.r[2] = w[2];
        } else {
#pragma unroll
            for (int w = 0; w < 3; ++w)
                v.r[w] = (uint32_t)s[4 * w] | ((uint32_t)s[4 * w + 1] << 8) |
                         ((uint32_t)s[4 * w + 2] << 16) | ((uint32_t)s[4 * w + 3] << 24);
        }
        return v;
    }
    static __device__ __forceinline__ float4 cv(const Raw &v) {
        return make_float4(cv_s24_bits(v.r[0]), cv_s24_bits((v.r[0] >> 24) | (v.r[1] << 8)),
                           cv_s24_bits((v.r[1] >> 16) | (v.r[2] << 16)), cv_s24_bits(v.r[2] >> 8));
    }
};
template <>
struct Piece<RSB_PCM_S24, 2> {
    struct Raw { uint32_t a, b; };   // the two samples' 24 bits
    static __device__ __forceinline__ Raw ld(const void *src, uint64_t q, bool aligned, uint32_t) {
        const uint8_t *s = static_cast<const uint8_t *>(src) + 6 * q;
        if (aligned) {
            // 6 bytes at an even offset: three halfwords
            const uint16_t *h = reinterpret_cast<const uint16_t *>(s);
            const uint32_t h0 = h[0], h1 = h[1], h2 = h[2];
            return Raw{h0 | ((h1 & 0xffu) << 16), (h1 >> 8) | (h2 << 8)};
        }
        return Raw{(uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16),
                   (uint32_t)s[3] | ((uint32_t)s[4] << 8) | ((uint32_t)s[5] << 16)};
    }
    static __device__ __forceinline__ float4 cv(const Raw &v) {
        const float a = cv_s24_bits(v.a), b = cv_s24_bits(v.b);
        return make_float4(a, a, b, b);
    }
};
template <>
struct Piece<RSB_PCM_S24, 0> {
    struct Raw { float4 f; };
    static __device__ __forceinline__ Raw ld(const void *src, uint64_t q, bool, uint32_t dup) {
        const uint8_t *s = static_cast<const uint8_t *>(src);
        const uint64_t o = 4 * q;
        return Raw{make_float4(cv_s24(s + 3 * (o / dup)), cv_s24(s + 3 * ((o + 1) / dup)),
                               cv_s24(s + 3 * ((o + 2) / dup)), cv_s24(s + 3 * ((o + 3) / dup)))};
    }
    static __device__ __forceinline__ float4 cv(const Raw &v) { return v.f; }
};

template <int FMT>
__device__ __forceinline__ float load1(const void *src, uint64_t i) {
    if constexpr (FMT == RSB_PCM_S24) return cv_s24(static_cast<const uint8_t *>(src) + 3 * i);
    else return Src<FMT>::cv(static_cast<const typename Src<FMT>::T *>(src)[i]);
}

// One CTA walks chunks c = blockIdx.x, + gridDim.x, ...; chunk c belongs to the job j with
// chunk_first[j] <= c < chunk_first[j + 1]: a division when every job has the same number of
// chunks (cpj != 0), else a binary search (uniform per CTA).  A thread keeps kPer independent
// loads in flight before its kPer 16-byte stores.
// resident CTAs per SM the register allocation is asked to allow: five when a piece in flight is
// at most two registers (40 warps keep 8 loads each in flight; six would spill), free otherwise
template <int FMT, int DUP>
constexpr int min_ctas() { return sizeof(typename Piece<FMT, DUP>::Raw) <= 8 ? 5 : 1; }

template <int FMT, int DUP>
__global__ void __launch_bounds__(kThreads, min_ctas<FMT, DUP>())
pcm_ingest_kernel(const PcmJob *__restrict__ jobs, const uint64_t *__restrict__ chunk_first,
                  uint32_t n_jobs, uint64_t n_chunks, uint32_t dup, uint64_t cpj) {
    constexpr uint32_t kPer = kChunkVec / kThreads;
    using P = Piece<FMT, DUP>;
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        uint32_t lo;
        uint64_t first;
        if (cpj) {
            lo = (uint32_t)(c / cpj);
            first = (uint64_t)lo * cpj;
        } else {
            uint32_t hi = n_jobs;   // invariant: chunk_first[lo] <= c < chunk_first[hi]
            lo = 0;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (chunk_first[mid] <= c) lo = mid; else hi = mid;
            }
            first = chunk_first[lo];
        }
        const PcmJob J = jobs[lo];
        const uint64_t q0 = (c - first) * kChunkVec + threadIdx.x;
        const uint64_t n_vec = J.n_out >> 2;   // whole float4 pieces of the job
        const bool aligned = J.src_aligned != 0;
        float4 *dst = reinterpret_cast<float4 *>(J.dst);
        typename P::Raw v[kPer];
        if (q0 - threadIdx.x + kChunkVec <= n_vec) {   // interior chunk: no guards
#pragma unroll
            for (uint32_t k = 0; k < kPer; ++k) v[k] = P::ld(J.src, q0 + k * kThreads, aligned, dup);
#pragma unroll
            for (uint32_t k = 0; k < kPer; ++k) __stcs(dst + q0 + k * kThreads, P::cv(v[k]));
            continue;
        }
#pragma unroll
        for (uint32_t k = 0; k < kPer; ++k) {
            const uint64_t q = q0 + k * kThreads;
            if (q < n_vec) v[k] = P::ld(J.src, q, aligned, dup);
        }
#pragma unroll
        for (uint32_t k = 0; k < kPer; ++k) {
            const uint64_t q = q0 + k * kThreads;
            if (q < n_vec) dst[q] = P::cv(v[k]);
        }
        // the job's last 1-3 values (this is the job's last chunk)
        const uint64_t tail0 = n_vec << 2;
        if (threadIdx.x < J.n_out - tail0) {
            const uint64_t o = tail0 + threadIdx.x;
            J.dst[o] = load1<FMT>(J.src, o / dup);
        }
    }
}

template <int FMT, int DUP>
void launch_one(const PcmJob *jobs, const uint64_t *chunk_first, uint32_t n_jobs, uint64_t n_chunks,
                uint32_t dup, uint64_t cpj, int sm_count, cudaStream_t stream) {
    // persistent grid: exactly the CTAs that are resident at once, a multiple of the SM count
    static int per_sm = 0;
    if (per_sm == 0 &&
        (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcm_ingest_kernel<FMT, DUP>, kThreads, 0) !=
             cudaSuccess || per_sm <= 0))
        per_sm = 4;
    uint64_t grid = (uint64_t)sm_count * (uint64_t)per_sm;
    if (n_chunks < grid) grid = n_chunks;
    pcm_ingest_kernel<FMT, DUP><<<(uint32_t)grid, kThreads, 0, stream>>>(jobs, chunk_first, n_jobs,
                                                                         n_chunks, dup, cpj);
}

template <int FMT>
void launch_fmt(const PcmJob *jobs, const uint64_t *chunk_first, uint32_t n_jobs, uint64_t n_chunks,
                uint32_t dup, uint64_t cpj, int sm_count, cudaStream_t stream) {
    if (dup == 1) launch_one<FMT, 1>(jobs, chunk_first, n_jobs, n_chunks, dup, cpj, sm_count, stream);
    else if (dup == 2) launch_one<FMT, 2>(jobs, chunk_first, n_jobs, n_chunks, dup, cpj, sm_count, stream);
    else launch_one<FMT, 0>(jobs, chunk_first, n_jobs, n_chunks, dup, cpj, sm_count, stream);
}

}  // namespace

uint32_t pcm_bytes_per_sample(int format) {
    switch (format) {
        case RSB_PCM_U8: return 1;
        case RSB_PCM_S16: return 2;
        case RSB_PCM_S24: return 3;
        case RSB_PCM_S32: return 4;
        case RSB_PCM_F32: return 4;
        default: return 0;
    }
}

uint64_t pcm_chunks(uint64_t n_out) {
    const uint64_t vec = (n_out + 3) >> 2;   // the tail piece counts as one
    return (vec + kChunkVec - 1) / kChunkVec;
}

bool launch_pcm_ingest(const PcmJob *jobs, const uint64_t *chunk_first, uint32_t n_jobs,
                       uint64_t n_chunks, uint64_t chunks_per_job, int format, uint32_t dup,
                       int sm_count, cudaStream_t stream) {
    if (n_jobs == 0 || n_chunks == 0) return true;
    switch (format) {
        case RSB_PCM_U8: launch_fmt<RSB_PCM_U8>(jobs, chunk_first, n_jobs, n_chunks, dup, chunks_per_job, sm_count, stream); break;
        case RSB_PCM_S16: launch_fmt<RSB_PCM_S16>(jobs, chunk_first, n_jobs, n_chunks, dup, chunks_per_job, sm_count, stream); break;
        case RSB_PCM_S24: launch_fmt<RSB_PCM_S24>(jobs, chunk_first, n_jobs, n_chunks, dup, chunks_per_job, sm_count, stream); break;
        case RSB_PCM_S32: launch_fmt<RSB_PCM_S32>(jobs, chunk_first, n_jobs, n_chunks, dup, chunks_per_job, sm_count, stream); break;
        case RSB_PCM_F32: launch_fmt<RSB_PCM_F32>(jobs, chunk_first, n_jobs, n_chunks, dup, chunks_per_job, sm_count, stream); break;
        default: return false;
    }
    return true;
}

}  // namespace rsb
