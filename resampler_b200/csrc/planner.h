// planner.h -- exact emulation of ResamplerFir's f64 phase accumulator, shared
// by host and device code (__host__ __device__).
//
// Reference behaviour reproduced (hasenbanck/resampler v0.5.1):
//   src/resampler_fir.rs:521-528  frames_to_copy = min(in, 8192-write_pos, 4096-available)
//   src/resampler_fir.rs:542-590  output loop: off = floor(pos); stop when off+taps > available
//                                 or produced >= capacity; pos += ratio once per output frame
//   src/resampler_fir.rs:558-565  phase_f = min(fract(pos)*1024, 1023); phase1 = trunc;
//                                 phase2 = min(phase1+1, 1023); frac = (phase_f-phase1) as f32
//   src/resampler_fir.rs:592-602  consumed = min(floor(pos), available); pos -= consumed
//
// The accumulator is a *rounded* f64 recurrence p <- fl(p + ratio); its rounding
// pattern is observable, so positions must be
// bit-identical.  Instead of one dependent DADD per output frame the walk uses
// a closed form per binade: while p stays inside one binade [2^e, 2^(e+1)) every
// value is a multiple of u = ulp(2^e) and fl(p + ratio) = p + D*u with a
// constant integer D (after at most one step that resolves a round-half-even
// tie).  So positions inside a binade form an exact arithmetic progression in
// the *bit pattern*: bits(p_j) = bits(p_1) + j*D.  Only the steps that enter a
// new binade are executed as real f64 additions.  A "segment" is one such
// progression; a call of 512 frames needs ~20 segments instead of ~557 DADDs.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RSB_HD __host__ __device__ __forceinline__
#else
#define RSB_HD inline
#endif

namespace rsb {

constexpr uint32_t kPhases = 1024;          // resampler_fir.rs:17
constexpr uint32_t kInputCapacity = 4096;   // resampler_fir.rs:18

// One arithmetic progression of accumulator values (all inside one binade).
struct PlanSeg {
    int64_t base_bits;   // bit pattern of the first position
    int64_t step_bits;   // bits(p_{j+1}) - bits(p_j)  (0 when n == 1)
    uint32_t n;          // number of output frames in the segment
    uint32_t out0;       // index of its first output frame within the unit's flat output
    int64_t vbase;       // virtual input frame that position 0.0 refers to (see fir_common.h)
};

RSB_HD int64_t d2bits(double d) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(d);
#else
    int64_t b; memcpy(&b, &d, 8); return b;
#endif
}
RSB_HD double bits2d(int64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double d; memcpy(&d, &b, 8); return d;
#endif
}
// rounded f64 add that the compiler may not contract or reassociate
RSB_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}

// Per-output-frame quantities derived from a position (resampler_fir.rs:544, 558-565).
struct PhasePoint {
    uint32_t off;      // floor(position)
    uint32_t phase1;
    uint32_t phase2;
    float frac;
};
RSB_HD PhasePoint phase_point(double position) {
    PhasePoint r;
#if defined(__CUDA_ARCH__)
    // explicit _rn intrinsics: nvcc may not contract or reorder these
    const double fl = floor(position);
    const double fract = __dsub_rn(position, fl);       // position >= 0: fract() == p - floor(p)
    double phase_f = __dmul_rn(fract, (double)kPhases);
    if (phase_f > (double)(kPhases - 1)) phase_f = (double)(kPhases - 1);
    r.off = (uint32_t)fl;
    r.phase1 = (uint32_t)phase_f;                       // truncating cast (:563)
    r.frac = __double2float_rn(__dsub_rn(phase_f, (double)r.phase1));
#else
    const double fl = __builtin_floor(position);
    const double fract = position - fl;
    double phase_f = fract * (double)kPhases;
    if (phase_f > (double)(kPhases - 1)) phase_f = (double)(kPhases - 1);
    r.off = (uint32_t)fl;
    r.phase1 = (uint32_t)phase_f;
    r.frac = (float)(phase_f - (double)r.phase1);       // round-to-nearest-even cast
#endif
    r.phase2 = r.phase1 + 1 < kPhases - 1 ? r.phase1 + 1 : kPhases - 1;
    return r;
}

// ceil(num / den) for 0 < num, den < 2^53 with a small quotient (a call's output count).  A
// 64-bit integer division costs ~100 instructions on the GPU and an f64 division ~250 cycles of
// latency, and both sit on the planner's serial critical path (about 20 per 512-frame call), so
// the quotient is estimated with ONE multiply by a cached reciprocal and corrected exactly with
// integer arithmetic.  The step `den` is a constant of the binade (see the header comment), so
// the reciprocal is computed once per binade and reused by every later call.
struct DivCache {
    int64_t den[16];
    double inv[16];
};
RSB_HD void div_cache_reset(DivCache &c) {
    for (int i = 0; i < 16; ++i) { c.den[i] = 0; c.inv[i] = 0.0; }
}
RSB_HD int64_t ceil_div_53(int64_t num, int64_t den, DivCache &c, int slot) {
    slot &= 15;
    if (c.den[slot] != den) {
        c.den[slot] = den;
#if defined(__CUDA_ARCH__)
        c.inv[slot] = __ddiv_rn(1.0, (double)den);
#else
        c.inv[slot] = 1.0 / (double)den;
#endif
    }
#if defined(__CUDA_ARCH__)
    int64_t q = (int64_t)__dmul_rn((double)num, c.inv[slot]);   // floor estimate, off by <= 1
#else
    volatile double est = (double)num * c.inv[slot];
    int64_t q = (int64_t)est;
#endif
    int64_t rem = num - q * den;
    while (rem < 0) { q -= 1; rem += den; }
    while (rem >= den) { q += 1; rem -= den; }
    return rem > 0 ? q + 1 : q;
}

RSB_HD bool same_binade(int64_t a, int64_t b) {
    // equal biased exponents, and not zero/subnormal
    return ((a ^ b) >> 52) == 0 && ((a >> 52) & 0x7ff) != 0;
}

// Emits the segments of ONE call's output loop.  `pos` is the accumulator on
// entry (updated to its value on exit, i.e. last output position + ratio);
// available = frames in the buffer after the append; returns frames produced.
// Sink must provide: void seg(int64_t base_bits, int64_t step_bits, uint32_t n).
template <class Sink>
RSB_HD uint32_t plan_call_outputs(double &pos, double ratio, uint32_t available, uint32_t taps,
                                  uint32_t cap, Sink &sink, DivCache &dc) {
    if (available < taps || cap == 0) return 0;
    // off + taps > available  <=>  floor(p) >= available - taps + 1  <=>  p >= L
    const double L = (double)(available - taps + 1);
    uint32_t count = 0;
    double p = pos;
    while (p < L && count < cap) {
        const double p1 = dadd(p, ratio);
        const int64_t b0 = d2bits(p), b1 = d2bits(p1);
        bool done = false;
        if (count + 1 < cap && p1 < L && same_binade(b0, b1)) {
            const double p2 = dadd(p1, ratio);
            const int64_t b2 = d2bits(p2);
            const int64_t D = b2 - b1;
            if (same_binade(b1, b2) && D > 0) {
                // progression starts at p1: p1 was produced by an in-binade step, so a
                // round-half-even tie (if ratio sits exactly between two grid points) has
                // already settled on the even mantissa and D is the constant step from here.
                const int64_t mant_lim = (int64_t)1 << 53;
                const int64_t B = (b1 & (((int64_t)1 << 52) - 1)) | ((int64_t)1 << 52);
                // elements j with B + j*D < 2^53 stay below 2^(e+1)
                const int e = (int)((b1 >> 52) & 0x7ff) - 1023;
                int64_t n = ceil_div_53(mant_lim - B, D, dc, e);
                // elements with position < L
                if (L < bits2d((int64_t)(e + 1 + 1023) << 52)) {
                    // L / u is an exact integer < 2^53
                    const double scale = bits2d((int64_t)(52 - e + 1023) << 52);
                    const int64_t LQ = (int64_t)(L * scale);
                    const int64_t nL = ceil_div_53(LQ - B, D, dc, e);   // LQ > B because p1 < L
                    if (nL < n) n = nL;
                }
                const int64_t room = (int64_t)(cap - count - 1);
                if (room < n) n = room;
                // merge the leading element when it continues the same progression
                if (b1 - b0 == D) {
                    sink.seg(b0, D, (uint32_t)(n + 1));
                } else {
                    sink.seg(b0, 0, 1u);
                    sink.seg(b1, D, (uint32_t)n);
                }
                count += (uint32_t)(n + 1);
                const double last = bits2d(b1 + (n - 1) * D);
                p = dadd(last, ratio);
                done = true;
            }
        }
        if (!done) {
            sink.seg(b0, 0, 1u);
            count += 1;
            p = p1;
        }
    }
    pos = p;
    return count;
}

// Streaming state of one resampler that is not sample data.
struct PlanState {
    double position;      // resampler_fir.rs:193
    uint32_t available;   // resampler_fir.rs:191 (frames buffered, = history length between calls)
};

struct CallResult {
    uint32_t copied;      // frames_to_copy  (consumed value / channels)
    uint32_t produced;    // output frames
    uint32_t advanced;    // consumed_frames of :596 (how far read_position moved)
};

// One resample() call on the state machine.  `remaining_capacity` of :525-527 never
// binds (read_position <= 4096 at every entry because of the compaction at :605-615),
// so read_position itself is unobservable and is not modelled.
template <class Sink>
RSB_HD CallResult plan_call(PlanState &st, double ratio, uint32_t taps, uint32_t in_frames,
                            uint32_t cap_frames, Sink &sink, DivCache &dc) {
    CallResult r;
    uint32_t room = kInputCapacity - st.available;
    r.copied = in_frames < room ? in_frames : room;               // :526-528
    st.available += r.copied;                                     // :538
    r.produced = plan_call_outputs(st.position, ratio, st.available, taps, cap_frames, sink, dc);
#if defined(__CUDA_ARCH__)
    double fl = floor(st.position);
#else
    double fl = __builtin_floor(st.position);
#endif
    // :596  (position.floor() as usize).min(available_frames); `as usize` saturates
    uint32_t adv = fl >= (double)st.available ? st.available : (uint32_t)fl;
    r.advanced = adv;
    st.available -= adv;                                          // :601
    st.position = dadd(st.position, -(double)adv);                // :602 (exact)
    return r;
}

}  // namespace rsb
