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

// the four source values behind output values [4q, 4q + 4): source index = output index / dup
template <int FMT, int DUP>
__device__ __forceinline__ float4 load4(const void *src, uint64_t q, bool aligned, uint32_t dup) {
    using S = Src<FMT>;
    using T = typename S::T;
    const T *s = static_cast<const T *>(src);
    float4 r;
    if constexpr (DUP == 1) {
        if (aligned) {
            // 4 source values in one load of 4 * sizeof(T) bytes
            struct alignas(4 * sizeof(T)) V { T v[4]; };
            const V v = *reinterpret_cast<const V *>(s + 4 * q);
            r = make_float4(S::cv(v.v[0]), S::cv(v.v[1]), S::cv(v.v[2]), S::cv(v.v[3]));
        } else {
            r = make_float4(S::cv(s[4 * q]), S::cv(s[4 * q + 1]), S::cv(s[4 * q + 2]), S::cv(s[4 * q + 3]));
        }
    } else if constexpr (DUP == 2) {
        float a, b;
        if (aligned) {
            struct alignas(2 * sizeof(T)) V { T v[2]; };
            const V v = *reinterpret_cast<const V *>(s + 2 * q);
            a = S::cv(v.v[0]);
            b = S::cv(v.v[1]);
        } else {
            a = S::cv(s[2 * q]);
            b = S::cv(s[2 * q + 1]);
        }
        r = make_float4(a, a, b, b);
    } else {
        const uint64_t o = 4 * q;
        r = make_float4(S::cv(s[o / dup]), S::cv(s[(o + 1) / dup]), S::cv(s[(o + 2) / dup]),
                        S::cv(s[(o + 3) / dup]));
    }
    return r;
}

template <int FMT>
__device__ __forceinline__ float load1(const void *src, uint64_t i) {
    return Src<FMT>::cv(static_cast<const typename Src<FMT>::T *>(src)[i]);
}

__device__ __forceinline__ float cv_s24_bits(uint32_t v) {
    return (float)((int)(v << 8) >> 8) * (1.0f / 8388608.0f);
}

template <int DUP>
__device__ __forceinline__ float4 load4_s24(const void *src, uint64_t q, bool aligned, uint32_t dup) {
    const uint8_t *s = static_cast<const uint8_t *>(src);
    const uint64_t o = 4 * q;
    if constexpr (DUP == 1) {
        if (aligned) {
            // 4 samples = 12 bytes = three aligned words
            const uint32_t *w = reinterpret_cast<const uint32_t *>(s + 12 * q);
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            return make_float4(cv_s24_bits(w0), cv_s24_bits((w0 >> 24) | (w1 << 8)),
                               cv_s24_bits((w1 >> 16) | (w2 << 16)), cv_s24_bits(w2 >> 8));
        }
    }
    if constexpr (DUP == 1)
        return make_float4(cv_s24(s + 3 * o), cv_s24(s + 3 * o + 3), cv_s24(s + 3 * o + 6),
                           cv_s24(s + 3 * o + 9));
    if constexpr (DUP == 2) {
        const float a = cv_s24(s + 3 * (2 * q)), b = cv_s24(s + 3 * (2 * q + 1));
        return make_float4(a, a, b, b);
    }
    return make_float4(cv_s24(s + 3 * (o / dup)), cv_s24(s + 3 * ((o + 1) / dup)),
                       cv_s24(s + 3 * ((o + 2) / dup)), cv_s24(s + 3 * ((o + 3) / dup)));
}

// One CTA walks chunks c = blockIdx.x, + gridDim.x, ...; chunk c belongs to the job j with
// chunk_first[j] <= c < chunk_first[j + 1]: a division when every job has the same number of
// chunks (cpj != 0), else a binary search (uniform per CTA).  A thread keeps kPer independent
// loads in flight before its kPer 16-byte stores.
template <int FMT, int DUP>
__device__ __forceinline__ float4 load_piece(const PcmJob &J, uint64_t q, bool aligned, uint32_t dup) {
    if constexpr (FMT == RSB_PCM_S24) return load4_s24<DUP>(J.src, q, aligned, dup);
    else return load4<FMT, DUP>(J.src, q, aligned, dup);
}

template <int FMT, int DUP>
__global__ void __launch_bounds__(kThreads)
pcm_ingest_kernel(const PcmJob *__restrict__ jobs, const uint64_t *__restrict__ chunk_first,
                  uint32_t n_jobs, uint64_t n_chunks, uint32_t dup, uint64_t cpj) {
    constexpr uint32_t kPer = kChunkVec / kThreads;
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
        float4 v[kPer];
        if (q0 - threadIdx.x + kChunkVec <= n_vec) {   // interior chunk: no guards
#pragma unroll
            for (uint32_t k = 0; k < kPer; ++k) v[k] = load_piece<FMT, DUP>(J, q0 + k * kThreads, aligned, dup);
#pragma unroll
            for (uint32_t k = 0; k < kPer; ++k) __stcs(dst + q0 + k * kThreads, v[k]);
            continue;
        }
#pragma unroll
        for (uint32_t k = 0; k < kPer; ++k) {
            const uint64_t q = q0 + k * kThreads;
            if (q < n_vec) v[k] = load_piece<FMT, DUP>(J, q, aligned, dup);
        }
#pragma unroll
        for (uint32_t k = 0; k < kPer; ++k) {
            const uint64_t q = q0 + k * kThreads;
            if (q < n_vec) dst[q] = v[k];
        }
        // the job's last 1-3 values (this is the job's last chunk)
        const uint64_t tail0 = n_vec << 2;
        if (threadIdx.x < J.n_out - tail0) {
            const uint64_t o = tail0 + threadIdx.x;
            float x;
            if constexpr (FMT == RSB_PCM_S24) x = cv_s24(static_cast<const uint8_t *>(J.src) + 3 * (o / dup));
            else x = load1<FMT>(J.src, o / dup);
            J.dst[o] = x;
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
