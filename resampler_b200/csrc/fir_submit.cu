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
#include <cstdlib>

#include "fir_submit.h"

namespace rsb {

namespace {

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
        res.hist0 = s_h0;
        res.n_seg = s_nseg;
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

// ---- thread-per-output variant (the default) ----
// The half-warp kernel above spends ~60 warp instructions per output sample: every lane owns
// only taps/16 FMAs per phase, the rest is addressing, shuffles and the blend.  Here ONE thread
// computes an output sample and keeps all 16 "zmm lanes" as private accumulators, so the
// reference's arithmetic (per-lane FMA chains over taps l, l + 16, ..., unfused blend per lane,
// halving tree 8 / 4 / 2 / 1: fir/avx512.rs:22-48) runs entirely in registers, in the same order:
// the samples stay BIT-IDENTICAL to the oracle.  Neighbouring lanes are advanced two at a time with
// packed fp32 instructions (fma / mul / add .f32x2: each half is the IEEE scalar operation).
//   * The call's input window [history | new input] does not depend on the plan: warps 1-3 copy it
//     into shared memory WHILE thread 0 runs the serial planner (a window that does not fit is read
//     from global memory instead).
//   * The per-frame plan is expanded for up to 1024 frames at a time; after that the warps work
//     independently on blocks of 32 outputs (no CTA-wide barrier per block).
//   * The two coefficient rows of an output differ from thread to thread (every output has its own
//     phase), so a warp first copies the rows of its block's frames into shared memory with
//     16-byte asynchronous copies (row stride TB + 4 floats, the phase-2 rows in a region of their
//     own: the later per-thread 16-byte reads of eight consecutive slots touch eight different bank
//     groups), TB = 16 taps at a time.
constexpr uint32_t kTpThreads = 128;
constexpr uint32_t kSuper = 1024;          // frames expanded per pass
constexpr uint32_t kTpRowsMax = 33;        // staged phases per warp at most (32 mono outputs + 1): offset of the phase-2 rows
constexpr int kTpTapBlock = 16;            // taps staged per round (16: half the staging memory of 32, more CTAs per SM)

__device__ __forceinline__ float2 fmul2(const float2 a, const float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<const uint64_t &>(a)), "l"(reinterpret_cast<const uint64_t &>(b)));
    return d;
}
__device__ __forceinline__ float2 fadd2(const float2 a, const float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<const uint64_t &>(a)), "l"(reinterpret_cast<const uint64_t &>(b)));
    return d;
}
// d = a * b + d on both halves (fma.rn.f32x2, SASS FFMA2): two IEEE fp32 FMAs in one instruction
__device__ __forceinline__ void ffma2(float2 &d, const float2 a, const float2 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<const uint64_t &>(a)), "l"(reinterpret_cast<const uint64_t &>(b)));
}

// value e of the virtual buffer [history | new input] (interleaved), from shared or global memory
template <bool XS>
__device__ __forceinline__ float win_at(const float *s_x, const float *hist, const float *in, int64_t HC, uint32_t e) {
    if (XS) return s_x[e];
    return (int64_t)e < HC ? __ldg(hist + e) : __ldg(in + ((int64_t)e - HC));
}

// One block of 32 consecutive output VALUES (frame-major, channel-minor) of the call, by one warp.
template <int TAPS, bool XS>
__device__ __forceinline__ void tp_block(uint32_t idx0, uint32_t n_vals, uint32_t k0, uint32_t ch, const int32_t *s_v,
                                         const uint16_t *s_p1, const float *s_frac, const float *s_x, const float *hist,
                                         const float *in, int64_t HC, const float *coeffs, float *cw, uint16_t *slot_p1,
                                         float *out, uint32_t lane) {
    constexpr int TB = TAPS < kTpTapBlock ? TAPS : kTpTapBlock;
    constexpr int CS = TB + 4;
    constexpr uint32_t CPR = TB / 4, kStageRows = 32u / (2u * CPR);
    // staging role of this lane: 16-byte piece st_q of phase st_ph, frames st_r, st_r + kStageRows, ...
    const uint32_t st_q = lane % CPR, st_ph = (lane / CPR) & 1u, st_r = lane / (2u * CPR);
    // phase-1 rows from cw, phase-2 rows kTpRowsMax rows behind them: consecutive slots are CS floats apart
    const uint32_t st_dst = (uint32_t)__cvta_generic_to_shared(cw) + (st_ph * kTpRowsMax * CS + st_q * 4u) * 4u;
    const uint32_t idx = idx0 + lane;
    const bool valid = idx < n_vals;
    const uint32_t idc = valid ? idx : n_vals - 1;
    const uint32_t fg = idc / ch, c = idc - fg * ch;                 // frame of the call, channel
    const uint32_t f_lo = idx0 / ch;
    const uint32_t f_hi = (min(idx0 + 31u, n_vals - 1)) / ch;
    const uint32_t nr = f_hi - f_lo + 1u;
    const uint32_t fr = fg - f_lo;
    const uint32_t xb = (uint32_t)s_v[fg - k0] * ch + c;
    // Frames of the block that share a phase share one staged pair of rows (a 1 : 3 ratio cycles
    // through three phases: 3 row pairs instead of 32; 44.1 -> 48 kHz has 160 phases: no sharing).
    // Lane r speaks for frame f_lo + r (nr <= 32).
    const uint32_t my_p1 = s_p1[f_lo + min(lane, nr - 1u) - k0];
    const uint32_t same = __match_any_sync(0xffffffffu, lane < nr ? my_p1 : 0x10000u + lane);
    const uint32_t leader = (uint32_t)__ffs((int)same) - 1u;                    // first frame with this phase
    const uint32_t lead_mask = __ballot_sync(0xffffffffu, lane < nr && leader == lane);
    const uint32_t n_lead = (uint32_t)__popc(lead_mask);
    const uint32_t my_slot = (uint32_t)__popc(lead_mask & ((1u << leader) - 1u));   // staged slot of frame `lane`
    const uint32_t fr_slot = __shfl_sync(0xffffffffu, my_slot, fr);                  // ... of this thread's frame
    if (lane < nr && leader == lane) slot_p1[my_slot] = (uint16_t)my_p1;             // phase of staged slot j, j < n_lead
    __syncwarp();
    float2 acc1[8], acc2[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) acc1[l] = acc2[l] = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int tb = 0; tb < TAPS / TB; ++tb) {
        {
            const float *src_l = coeffs + tb * TB + st_q * 4u;
            for (uint32_t j0 = 0; j0 < n_lead; j0 += 4u * kStageRows) {
                // four row pairs per trip: the phase look-ups first, then the copies
                uint32_t p[4];
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u)
                    p[u] = min((uint32_t)slot_p1[min(j0 + u * kStageRows + st_r, 31u)] + st_ph, kPhases - 1);
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    const uint32_t j = j0 + u * kStageRows + st_r;
                    if (j < n_lead)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(st_dst + j * (CS * 4u)),
                                     "l"(src_l + (size_t)p[u] * TAPS)
                                     : "memory");
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncwarp();
        const float4 *r1 = reinterpret_cast<const float4 *>(cw + (size_t)fr_slot * CS);
        const float4 *r2 = reinterpret_cast<const float4 *>(cw + (size_t)fr_slot * CS + kTpRowsMax * CS);
        const uint32_t xq = xb + (uint32_t)tb * TB * ch;
#pragma unroll
        for (int q = 0; q < TB / 4; ++q) {
            const float4 a = r1[q], b = r2[q];
            const float2 x01 = make_float2(win_at<XS>(s_x, hist, in, HC, xq + (4 * q) * ch),
                                           win_at<XS>(s_x, hist, in, HC, xq + (4 * q + 1) * ch));
            const float2 x23 = make_float2(win_at<XS>(s_x, hist, in, HC, xq + (4 * q + 2) * ch),
                                           win_at<XS>(s_x, hist, in, HC, xq + (4 * q + 3) * ch));
            // tap t = tb * TB + 4 q + e belongs to zmm lane t % 16 (TB is a multiple of 16)
            ffma2(acc1[(2 * q) & 7], make_float2(a.x, a.y), x01);
            ffma2(acc2[(2 * q) & 7], make_float2(b.x, b.y), x01);
            ffma2(acc1[(2 * q + 1) & 7], make_float2(a.z, a.w), x23);
            ffma2(acc2[(2 * q + 1) & 7], make_float2(b.z, b.w), x23);
        }
        __syncwarp();
    }
    // blend per lane (unfused mul, mul, add), then the halving tree of _mm512_reduce_add_ps
    const float frac = s_frac[fg - k0];
    const float omf = __fsub_rn(1.0f, frac);
    const float2 omf2 = make_float2(omf, omf), frac2 = make_float2(frac, frac);
    float2 t[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        // products packed, sums scalar: ptxas contracts mul.f32x2 + add.f32x2 into a packed FMA,
        // the reference's blend is unfused (avx512.rs:41-45)
        const float2 m1 = fmul2(acc1[l], omf2), m2 = fmul2(acc2[l], frac2);
        t[l] = make_float2(__fadd_rn(m1.x, m2.x), __fadd_rn(m1.y, m2.y));
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) t[l] = fadd2(t[l], t[l + 4]);      // lanes l += l + 8
#pragma unroll
    for (int l = 0; l < 2; ++l) t[l] = fadd2(t[l], t[l + 2]);      // l += l + 4
    t[0] = fadd2(t[0], t[1]);                                       // l += l + 2
    const float y = __fadd_rn(t[0].x, t[0].y);                      // l += l + 1
    if (valid) out[idx] = y;
}

// Integer-ratio runs.  When out_hz / gcd = P is small the phases cycle with period P: output frames
// k, k + P, k + 2 P, ... share their two coefficient rows and their windows slide by D = in_hz / gcd
// input frames.  A lane = (group, residue r of the cycle, channel) then computes M such outputs with
// ONE pair of 16-byte coefficient loads per tap quad (straight from the table: the P row pairs in use
// stay in L1) and, for D == 1, a sliding window of M + 3 input values -- a third of the
// L1 / shared-memory traffic of tp_block(), which is that kernel's limiter.  The arithmetic is the
// same per-lane FMA chains, unfused blend and halving tree, evaluated lane group by lane group
// (zmm lanes 0-3 and 8-11, then 4-7 and 12-15) so that only 8 accumulators per output are live:
// bit-identical.  The block covers (32 / (P ch)) P M consecutive frames starting at fb; the plan is
// CHECKED (same phase, offsets D apart) and the block handed to tp_block() when it does not hold.
constexpr int kRunM = 3;      // outputs per lane (4 for 16 / 32 taps with sliding windows, see the kernel)
template <int TAPS, int M, bool SLIDE>
__device__ __forceinline__ bool tp_run_block(uint32_t fb, uint32_t k0, uint32_t ch, uint32_t P, uint32_t D, const int32_t *s_v,
                                             const uint16_t *s_p1, const float *s_frac, const float *s_x,
                                             const float *coeffs, float *out, uint32_t lane) {
    const uint32_t L = P * ch, G = 32u / L;
    const uint32_t g_raw = lane / L, w = lane - g_raw * L;
    const bool on = g_raw < G;
    const uint32_t g = on ? g_raw : 0u;              // spare lanes shadow group 0 (no store)
    const uint32_t r = w / ch, c = w - r * ch;
    const uint32_t f_first = fb + g * (P * M) + r;   // this lane's first frame of the call
    const uint32_t f0 = f_first - k0;                // ... in the pass tables
    const int32_t v0 = s_v[f0];
    const uint32_t p1 = s_p1[f0];
    bool ok = true;
    float omf[M], fr[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const uint32_t f = f0 + (uint32_t)m * P;
        ok = ok && s_v[f] == v0 + (int32_t)((uint32_t)m * D) && s_p1[f] == p1;
        fr[m] = s_frac[f];
        omf[m] = __fsub_rn(1.0f, fr[m]);
    }
    if (!__all_sync(0xffffffffu, ok)) return false;
    const float4 *c1 = reinterpret_cast<const float4 *>(coeffs + (size_t)p1 * TAPS);
    const float4 *c2 = reinterpret_cast<const float4 *>(coeffs + (size_t)min(p1 + 1u, kPhases - 1) * TAPS);
    const float *xp = s_x + (size_t)v0 * ch + c;     // tap t of output m: xp[(t + m D) ch]
    float s8a[M][4], y[M];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float s8[M][4];
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
            const int q = half + 2 * hi;             // zmm lanes 4 q .. 4 q + 3
            float a1[M][4], a2[M][4];
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
                for (int e = 0; e < 4; ++e) a1[m][e] = a2[m][e] = 0.0f;
#pragma unroll(TAPS / 16 > 2 ? 2 : TAPS / 16)
            for (int j = 0; j < TAPS / 16; ++j) {
                const float4 k1 = __ldg(c1 + 4 * j + q), k2 = __ldg(c2 + 4 * j + q);
                const float k1v[4] = {k1.x, k1.y, k1.z, k1.w}, k2v[4] = {k2.x, k2.y, k2.z, k2.w};
                const float *xq = xp + (size_t)(16 * j + 4 * q) * ch;
                if (SLIDE) {
                    float xv[M + 3];
#pragma unroll
                    for (int i = 0; i < M + 3; ++i) xv[i] = xq[(uint32_t)i * ch];
#pragma unroll
                    for (int m = 0; m < M; ++m)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            a1[m][e] = __fmaf_rn(k1v[e], xv[m + e], a1[m][e]);
                            a2[m][e] = __fmaf_rn(k2v[e], xv[m + e], a2[m][e]);
                        }
                } else {
#pragma unroll
                    for (int m = 0; m < M; ++m)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float x = xq[((uint32_t)m * D + (uint32_t)e) * ch];
                            a1[m][e] = __fmaf_rn(k1v[e], x, a1[m][e]);
                            a2[m][e] = __fmaf_rn(k2v[e], x, a2[m][e]);
                        }
                }
            }
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float t = __fadd_rn(__fmul_rn(a1[m][e], omf[m]), __fmul_rn(a2[m][e], fr[m]));   // blend, unfused
                    s8[m][e] = hi == 0 ? t : __fadd_rn(s8[m][e], t);                                     // lanes l += l + 8
                }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) {
            if (half == 0) {
#pragma unroll
                for (int e = 0; e < 4; ++e) s8a[m][e] = s8[m][e];
            } else {
                float s4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) s4[e] = __fadd_rn(s8a[m][e], s8[m][e]);                      // l += l + 4
                y[m] = __fadd_rn(__fadd_rn(s4[0], s4[2]), __fadd_rn(s4[1], s4[3]));                     // l += l + 2, l += l + 1
            }
        }
    }
    if (on) {
#pragma unroll
        for (int m = 0; m < M; ++m) out[(size_t)(f_first + (uint32_t)m * P) * ch + c] = y[m];
    }
    return true;
}

// m outputs per lane (a whole block: 3 or 4; the ragged end of a pass: whatever whole units remain)
template <int TAPS>
__device__ __forceinline__ bool tp_run_dispatch(uint32_t m, bool slide, uint32_t fb, uint32_t k0, uint32_t ch, uint32_t P,
                                                uint32_t D, const int32_t *s_v, const uint16_t *s_p1, const float *s_frac,
                                                const float *s_x, const float *coeffs, float *out, uint32_t lane) {
#define RSB_RUN(M_) \
    (slide ? tp_run_block<TAPS, M_, true>(fb, k0, ch, P, D, s_v, s_p1, s_frac, s_x, coeffs, out, lane) \
           : tp_run_block<TAPS, M_, false>(fb, k0, ch, P, D, s_v, s_p1, s_frac, s_x, coeffs, out, lane))
    switch (m) {
        case 4:
            if constexpr (TAPS <= 32) return tp_run_block<TAPS, 4, true>(fb, k0, ch, P, D, s_v, s_p1, s_frac, s_x, coeffs, out, lane);
            else return false;
        case 3: return RSB_RUN(3);
        case 2: return RSB_RUN(2);
        case 1: return RSB_RUN(1);
        default: return false;
    }
#undef RSB_RUN
}

// The serial part, one call per WARP (lane 0 works: calls of different sizes take different paths
// through the planner, 32 of them in one warp would run one after the other): exact phase plan of
// the call (planner.h) from the stream's state -> result record, plan segments, new scalar state.
// All calls of a submit are planned at the same time instead of one after the other at the head
// of their convolution CTAs.
template <int TAPS>
__global__ void __launch_bounds__(256) submit_plan_kernel(const SubmitJob *jobs, SubmitResult *results, uint32_t n_jobs,
                                                          StreamStateDev st, double ratio, PlanSeg *seg_store) {
    const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n_jobs || (threadIdx.x & 31u) != 0) return;
    const SubmitJob job = jobs[j];
    PlanState s;
    s.position = st.position[job.stream];
    s.available = st.hist_len[job.stream];
    const uint32_t h0 = s.available;
    SmemSink sink{seg_store + (size_t)j * kSubmitSegs, 0u, 0u};
    DivCache dc;
    div_cache_reset(dc);
    const CallResult r = plan_call(s, ratio, (uint32_t)TAPS, job.in_frames, job.cap_frames, sink, dc);
    const uint32_t status = sink.n > kSubmitSegs ? 1u : 0u;
    st.position[job.stream] = s.position;        // resampler_fir.rs:602
    st.hist_len[job.stream] = s.available;       // :601
    SubmitResult res;
    res.position = s.position;
    res.copied = r.copied;
    res.produced = status ? 0u : r.produced;
    res.available = s.available;
    res.status = status;
    res.hist0 = h0;
    res.n_seg = sink.n < kSubmitSegs ? sink.n : kSubmitSegs;
    results[j] = res;
}

template <int TAPS>
__global__ void __launch_bounds__(kTpThreads, 5) submit_fused_tp_kernel(const SubmitJob *jobs, const SubmitResult *results,
                                                                     const float *coeffs,
                                                                     const PlanSeg *seg_store, uint32_t ch,
                                                                     uint32_t rows_per_warp, uint32_t xw_vals,
                                                                     uint32_t run_p, uint32_t run_d) {
    constexpr int TB = TAPS < kTpTapBlock ? TAPS : kTpTapBlock;
    constexpr int CS = TB + 4;
    extern __shared__ __align__(16) float sm_tp[];
    float *s_c = sm_tp;                                            // [4 warps][kTpRowsMax + rows_per_warp][CS]
    float *s_x = sm_tp + 4u * (kTpRowsMax + rows_per_warp) * CS;   // [xw_vals]
    __shared__ PlanSeg s_segs[kSubmitSegs];
    __shared__ int32_t s_v[kSuper];
    __shared__ float s_frac[kSuper];
    __shared__ uint16_t s_p1[kSuper];
    __shared__ uint16_t s_slot[kTpThreads / 32u][32];      // per warp: phase of each staged row pair

    const SubmitJob job = jobs[blockIdx.x];
    const SubmitResult res = results[blockIdx.x];        // written by submit_plan_kernel
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t H = res.hist0;
    const int64_t HC = (int64_t)H * ch;                           // values of live history
    const float *hist = job.hist + ((int64_t)kHistFrames * ch - HC);   // hist[e] = value e of [history | input]
    const float *in = job.in;
    const uint32_t n_win = (H + job.in_frames) * ch;
    const bool xs = n_win <= xw_vals;
    const uint32_t produced = res.status ? 0u : res.produced, n_seg = res.n_seg;
    if (tid < n_seg) s_segs[tid] = seg_store[(size_t)blockIdx.x * kSubmitSegs + tid];
    if (xs) {
        // the call's input window, four independent loads in flight per thread
        for (uint32_t i0 = tid; i0 < n_win; i0 += 4u * kTpThreads) {
            float xv[4];
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) {
                const uint32_t i = i0 + u * kTpThreads;
                xv[u] = i < n_win ? ((int64_t)i < HC ? __ldg(hist + i) : __ldg(in + ((int64_t)i - HC))) : 0.0f;
            }
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) {
                const uint32_t i = i0 + u * kTpThreads;
                if (i < n_win) s_x[i] = xv[u];
            }
        }
    }
    __syncthreads();
    float *cw = s_c + (size_t)warp * (kTpRowsMax + rows_per_warp) * CS;

    // integer-ratio runs (tp_run_block): outputs per lane 4 with sliding windows at 16 / 32 taps, else 3
    // (registers; with strided windows 4 also puts the groups of an 8-channel batch on the same banks:
    // 35 -> 50 us); a pass is a whole number of blocks
    const bool runs = run_p != 0 && xs;
    const bool wide = TAPS <= 32 && run_d == 1u;
    const uint32_t unit = runs ? (32u / (run_p * ch)) * run_p : 1u;         // frames per output-per-lane
    const uint32_t bf = unit * (wide ? 4u : (uint32_t)kRunM);
    const uint32_t pass = runs ? (kSuper / bf) * bf : kSuper;
    for (uint32_t k0 = 0; k0 < produced; k0 += pass) {
        const uint32_t nf = min(pass, produced - k0);
        // ---- per-frame plan of this pass ----
        for (uint32_t i = tid; i < nf; i += kTpThreads) {
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
            s_p1[i] = (uint16_t)pp.phase1;
            s_frac[i] = pp.frac;
        }
        __syncthreads();
        // ---- blocks of 32 output values, warps independent ----
        const uint32_t v0 = k0 * ch, n_vals = (k0 + nf) * ch;
        if (runs) {
            // integer-ratio runs: whole blocks of m_full outputs per lane, the ragged end of the pass with
            // as many whole units as remain; what is left (less than one unit, or a block whose plan is not
            // periodic) takes the general routine
            for (uint32_t fb = k0 + warp * bf; fb < k0 + nf; fb += (kTpThreads / 32u) * bf) {
                const uint32_t fe = min(fb + bf, k0 + nf);
                const uint32_t m = (fe - fb) / unit;
                uint32_t f_gen = fb;
                if (m != 0 && tp_run_dispatch<TAPS>(m, run_d == 1u, fb, k0, ch, run_p, run_d, s_v, s_p1, s_frac, s_x, coeffs, job.out, lane))
                    f_gen = fb + m * unit;
                for (uint32_t idx0 = f_gen * ch; idx0 < fe * ch; idx0 += 32u)
                    tp_block<TAPS, true>(idx0, fe * ch, k0, ch, s_v, s_p1, s_frac, s_x, hist, in, HC, coeffs, cw, s_slot[warp], job.out, lane);
            }
        } else {
            for (uint32_t idx0 = v0 + warp * 32u; idx0 < n_vals; idx0 += kTpThreads) {
                if (xs) tp_block<TAPS, true>(idx0, n_vals, k0, ch, s_v, s_p1, s_frac, s_x, hist, in, HC, coeffs, cw, s_slot[warp], job.out, lane);
                else tp_block<TAPS, false>(idx0, n_vals, k0, ch, s_v, s_p1, s_frac, s_x, hist, in, HC, coeffs, cw, s_slot[warp], job.out, lane);
            }
        }
        __syncthreads();
    }

    // ---- new history: the last H1 frames of [history | copied input] -> the other buffer ----
    const int64_t n_keep = (int64_t)res.available * ch;
    const int64_t e_first = HC + (int64_t)res.copied * ch - n_keep;
    float *new_hist = job.hist_next + ((int64_t)kHistFrames * ch - n_keep);
    for (int64_t i = tid; i < n_keep; i += kTpThreads) {
        const int64_t e = e_first + i;
        new_hist[i] = e < HC ? hist[e] : in[e - HC];
    }
}

}  // namespace

// shared memory of the thread-per-output kernel for this handle geometry and the largest call of the
// submit; 0 = does not apply
static size_t tp_smem_bytes(uint32_t taps, uint32_t ch, uint32_t max_in_frames, uint32_t *rows_per_warp,
                            uint32_t *xw_vals) {
    if (ch == 0 || ch > 32) return 0;
    const uint32_t tb = taps < (uint32_t)kTpTapBlock ? taps : (uint32_t)kTpTapBlock, cs = tb + 4;
    *rows_per_warp = 31u / ch + 2u;
    const size_t c_bytes = (size_t)4 * (kTpRowsMax + *rows_per_warp) * cs * sizeof(float);
    // window: the call's new input plus the history a stream carries in steady state (taps + a few
    // frames of look-ahead); a stream holding more than that reads its window from global memory
    size_t want = ((size_t)max_in_frames + taps + 96u) * ch;
    const size_t cap = (48u * 1024u) / sizeof(float);
    if (want > cap) want = 0;           // large calls: no staging, every CTA takes the global path
    *xw_vals = (uint32_t)want;
    return c_bytes + want * sizeof(float);
}

// Two-stage form (planner kernel, then the thread-per-output convolution) available for this geometry?
bool submit_two_stage(uint32_t taps, uint32_t channels, uint32_t max_in_frames) {
    uint32_t rpw = 0, xw = 0;
    return !getenv("RSB_SUBMIT_HALFWARP") && tp_smem_bytes(taps, channels, max_in_frames, &rpw, &xw) != 0;
}

// Stage 1 (may run on another stream than stage 2, e.g. underneath the previous submit's convolution:
// it touches the job table, the result records, the plan segments and the streams' scalar state only)
void launch_submit_plan(const SubmitJob *jobs, SubmitResult *results, uint32_t n_jobs, StreamStateDev st, double ratio,
                        uint32_t taps, PlanSeg *seg_store, cudaStream_t stream) {
    if (n_jobs == 0) return;
    const uint32_t grid = (n_jobs + 7) / 8;
    switch (taps) {
        case 16: submit_plan_kernel<16><<<grid, 256, 0, stream>>>(jobs, results, n_jobs, st, ratio, seg_store); break;
        case 32: submit_plan_kernel<32><<<grid, 256, 0, stream>>>(jobs, results, n_jobs, st, ratio, seg_store); break;
        case 64: submit_plan_kernel<64><<<grid, 256, 0, stream>>>(jobs, results, n_jobs, st, ratio, seg_store); break;
        default: submit_plan_kernel<128><<<grid, 256, 0, stream>>>(jobs, results, n_jobs, st, ratio, seg_store); break;
    }
}

// Stage 2: samples and the new history
void launch_submit_conv(const SubmitJob *jobs, const SubmitResult *results, uint32_t n_jobs, const float *coeffs,
                        uint32_t taps, uint32_t channels, uint32_t max_in_frames, const PlanSeg *seg_store,
                        uint32_t in_hz, uint32_t out_hz, cudaStream_t stream) {
    if (n_jobs == 0) return;
    uint32_t rpw = 0, xw = 0;
    const size_t smem = tp_smem_bytes(taps, channels, max_in_frames, &rpw, &xw);
    // integer-ratio runs (tp_run_block): phase cycle P = out_hz / gcd short enough for two or more
    // (cycle x channels) groups per warp
    uint32_t a = in_hz, g = out_hz;
    while (a) { const uint32_t t = g % a; g = a; a = t; }
    uint32_t run_p = g ? out_hz / g : 0, run_d = g ? in_hz / g : 0;
    if (run_p == 0 || run_p * channels > 16u || run_d > 64u || getenv("RSB_SUBMIT_NO_RUNS")) run_p = 0;
    auto go = [&](auto kern, int which) {
        // the opt-in is per kernel and device: repeated only when a submit needs more than any before
        // it on this device (a few microseconds of host time per call otherwise)
        static size_t granted[4][64];
        int dev = 0;
        cudaGetDevice(&dev);
        size_t &have = granted[which][dev & 63];
        if (smem > have) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            have = smem;
        }
        kern<<<n_jobs, kTpThreads, smem, stream>>>(jobs, results, coeffs, seg_store, channels, rpw, xw, run_p, run_d);
    };
    switch (taps) {
        case 16: go(submit_fused_tp_kernel<16>, 0); break;
        case 32: go(submit_fused_tp_kernel<32>, 1); break;
        case 64: go(submit_fused_tp_kernel<64>, 2); break;
        default: go(submit_fused_tp_kernel<128>, 3); break;
    }
}

// Single-launch fallback (half-warp kernel: plan, samples and state in one kernel)
void launch_submit_fused(const SubmitJob *jobs, SubmitResult *results, uint32_t n_jobs, StreamStateDev st,
                         const float *coeffs, double ratio, uint32_t taps, uint32_t channels,
                         uint32_t /*max_in_frames*/, PlanSeg * /*seg_store*/, cudaStream_t stream) {
    if (n_jobs == 0) return;
    switch (taps) {
        case 16: submit_fused_kernel<16><<<n_jobs, 256, 0, stream>>>(jobs, results, st, coeffs, ratio, channels); break;
        case 32: submit_fused_kernel<32><<<n_jobs, 256, 0, stream>>>(jobs, results, st, coeffs, ratio, channels); break;
        case 64: submit_fused_kernel<64><<<n_jobs, 256, 0, stream>>>(jobs, results, st, coeffs, ratio, channels); break;
        default: submit_fused_kernel<128><<<n_jobs, 256, 0, stream>>>(jobs, results, st, coeffs, ratio, channels); break;
    }
}

}  // namespace rsb
