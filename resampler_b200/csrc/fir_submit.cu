// fir_submit.cu -- the small-call streaming path: ONE launch per batched submit.
//
// rsb_fir_submit_batch is "N independent resample() calls, one per listed stream"
// (resampler_fir.rs:509-621).  For small calls the five-kernel pipeline of a big batch (plan,
// tile index, filter tiles, convolution, state update) is all launch latency, and streams whose
// call sizes differ share no plan, so the tensor / FFMA2 kernels (one filter tile for many
// streams) do not apply at all.  Here one CTA serves one stream's call from start to finish:
//
//   1. thread 0 runs the exact phase planner (planner.h, the same code as everywhere else) on the
//      stream's own state -> a handful of plan segments in shared memory, (copied, produced);
//   2. the CTA expands the per-frame plan (offset, phase1, frac: `phase_point`, :544, :558-565)
//      block by block into shared memory;
//   3. half-warps compute the output samples in the reference's AVX-512 order (16 lanes, FMA
//      chains, unfused blend, halving tree: fir/avx512.rs:22-48) -> samples are BIT-IDENTICAL to
//      the oracle, divergent streams included;
//   4. the CTA writes the new history tail into the stream's other history buffer and the
//      scalars (position, buffered frames) back to the device state and to the result record.
#include "fir_kernels.h"

namespace rsb {

namespace {

constexpr uint32_t kSubmitSegs = 128;       // plan segments of ONE call (bound: segs_per_call_bound)
constexpr uint32_t kSubmitBlock = 1024;     // output frames expanded per pass

struct SmemSink {
    PlanSeg *segs;
    uint32_t n;
    uint32_t out0;
    __device__ __forceinline__ void seg(int64_t base_bits, int64_t step_bits, uint32_t cnt) {
        if (n < kSubmitSegs) {
            PlanSeg s;
            s.base_bits = base_bits;
            s.step_bits = step_bits;
            s.n = cnt;
            s.out0 = out0;
            s.vbase = 0;
            segs[n] = s;
        }
        n += 1;
        out0 += cnt;
    }
};

template <int TAPS>
__global__ void __launch_bounds__(256) submit_fused_kernel(const SubmitJob *jobs, SubmitResult *results,
                                                           StreamStateDev st, const float *coeffs,
                                                           double ratio, uint32_t ch) {
    __shared__ PlanSeg s_segs[kSubmitSegs];
    __shared__ int32_t s_v[kSubmitBlock];
    __shared__ uint32_t s_p1[kSubmitBlock];
    __shared__ float s_frac[kSubmitBlock];
    __shared__ uint32_t s_nseg, s_produced, s_copied, s_h0, s_h1, s_status;

    const SubmitJob job = jobs[blockIdx.x];
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        PlanState s;
        s.position = st.position[job.stream];
        s.available = st.hist_len[job.stream];
        s_h0 = s.available;
        SmemSink sink{s_segs, 0u, 0u};
        DivCache dc;
        div_cache_reset(dc);
        const CallResult r = plan_call(s, ratio, (uint32_t)TAPS, job.in_frames, job.cap_frames, sink, dc);
        s_nseg = sink.n < kSubmitSegs ? sink.n : kSubmitSegs;
        s_status = sink.n > kSubmitSegs ? 1u : 0u;
        s_produced = s_status ? 0u : r.produced;
        s_copied = r.copied;
        s_h1 = s.available;
        st.position[job.stream] = s.position;        // resampler_fir.rs:602
        st.hist_len[job.stream] = s.available;       // :601
        SubmitResult res;
        res.position = s.position;
        res.copied = r.copied;
        res.produced = s_produced;
        res.available = s.available;
        res.status = s_status;
        results[blockIdx.x] = res;
    }
    __syncthreads();
    const uint32_t produced = s_produced, n_seg = s_nseg;
    const int64_t H = (int64_t)s_h0;
    const float *hist = job.hist;
    const float *in = job.in;
    const uint32_t lane16 = tid & 15u;
    const uint32_t hw = tid >> 4, n_hw = blockDim.x >> 4;

    for (uint32_t k0 = 0; k0 < produced; k0 += kSubmitBlock) {
        const uint32_t nk = min(kSubmitBlock, produced - k0);
        // ---- per-frame plan of this block ----
        for (uint32_t i = tid; i < nk; i += blockDim.x) {
            const uint32_t o = k0 + i;
            uint32_t lo = 0, hi = n_seg;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_segs[mid].out0 <= o) lo = mid; else hi = mid;
            }
            const PlanSeg sg = s_segs[lo];
            const double pos = bits2d(sg.base_bits + (int64_t)(o - sg.out0) * sg.step_bits);
            const PhasePoint pp = phase_point(pos);
            s_v[i] = (int32_t)pp.off;
            s_p1[i] = pp.phase1;
            s_frac[i] = pp.frac;
        }
        __syncthreads();
        // ---- samples, two per warp: the 16 lanes of a half-warp are the zmm lanes ----
        const uint32_t total = nk * ch;
        for (uint32_t base = 0; base < total; base += n_hw) {
            const uint32_t idx_raw = base + hw;
            const bool valid = idx_raw < total;
            const uint32_t idx = valid ? idx_raw : total - 1;
            const uint32_t k = idx / ch, c = idx - k * ch;
            const int64_t v = s_v[k];
            const uint32_t p1 = s_p1[k];
            const uint32_t p2 = p1 + 1 < kPhases - 1 ? p1 + 1 : kPhases - 1;
            const float frac = s_frac[k];
            const float *c1 = coeffs + (size_t)p1 * TAPS;
            const float *c2 = coeffs + (size_t)p2 * TAPS;
            float acc1 = 0.0f, acc2 = 0.0f;
#pragma unroll
            for (int i = 0; i < TAPS / 16; ++i) {
                const int64_t vv = v + i * 16 + lane16;
                const float x = vv < H ? __ldg(hist + ((int64_t)kHistFrames - H + vv) * ch + c)
                                       : __ldg(in + (vv - H) * ch + c);
                acc1 = __fmaf_rn(__ldg(c1 + i * 16 + lane16), x, acc1);
                acc2 = __fmaf_rn(__ldg(c2 + i * 16 + lane16), x, acc2);
            }
            const float omf = __fsub_rn(1.0f, frac);
            float s = __fadd_rn(__fmul_rn(acc1, omf), __fmul_rn(acc2, frac));
            s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 8));
            s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));
            s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
            s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
            if (valid && lane16 == 0) job.out[(size_t)(k0 + k) * ch + c] = s;
        }
        __syncthreads();
    }

    // ---- new history: the last H1 frames of [history | copied input] -> the other buffer ----
    const int64_t H1 = (int64_t)s_h1;
    const int64_t v_end = H + (int64_t)s_copied;
    const int64_t v0 = v_end - H1;
    const int64_t n_vals = H1 * ch;
    float *new_hist = job.hist_next;
    for (int64_t i = tid; i < n_vals; i += blockDim.x) {
        const int64_t f = i / ch, c = i - f * ch;
        const int64_t vv = v0 + f;
        const float x = vv < H ? hist[((int64_t)kHistFrames - H + vv) * ch + c] : in[(vv - H) * ch + c];
        new_hist[((int64_t)kHistFrames - H1 + f) * ch + c] = x;
    }
}

}  // namespace

void launch_submit_fused(const SubmitJob *jobs, SubmitResult *results, uint32_t n_jobs, StreamStateDev st,
                         const float *coeffs, double ratio, uint32_t taps, uint32_t channels,
                         cudaStream_t stream) {
    if (n_jobs == 0) return;
    switch (taps) {
        case 16: submit_fused_kernel<16><<<n_jobs, 256, 0, stream>>>(jobs, results, st, coeffs, ratio, channels); break;
        case 32: submit_fused_kernel<32><<<n_jobs, 256, 0, stream>>>(jobs, results, st, coeffs, ratio, channels); break;
        case 64: submit_fused_kernel<64><<<n_jobs, 256, 0, stream>>>(jobs, results, st, coeffs, ratio, channels); break;
        default: submit_fused_kernel<128><<<n_jobs, 256, 0, stream>>>(jobs, results, st, coeffs, ratio, channels); break;
    }
}

}  // namespace rsb
