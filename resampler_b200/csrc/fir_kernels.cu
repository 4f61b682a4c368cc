// fir_kernels.cu -- plan / tile / exact-convolution / state kernels for sm_100a.
//
// Pipeline of one submit (all on the handle's stream):
//   plan_units_kernel   one thread per plan unit: exact f64 phase walk in closed form
//                       (planner.h), emits segments + per-call counts + tile slices
//   tile_index_kernel   one warp per tile: first segment + the tile's 32 plan entries
//   conv_*_kernel       persistent CTAs over (tile, stream group)
//   update_state_kernel one CTA per job: history tail -> other history buffer, scalars
#include "fir_kernels.h"

namespace rsb {

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
struct SegSink {
    PlanSeg *segs;
    uint32_t cap;
    uint32_t n;
    uint32_t out0;
    int64_t vbase;
    __device__ __forceinline__ void seg(int64_t base_bits, int64_t step_bits, uint32_t cnt) {
        if (n < cap) {
            PlanSeg s;
            s.base_bits = base_bits;
            s.step_bits = step_bits;
            s.n = cnt;
            s.out0 = out0;
            s.vbase = vbase;
            segs[n] = s;
        }
        n += 1;
        out0 += cnt;
    }
};

__global__ void plan_units_kernel(UnitDev *units, uint32_t n_units, StreamStateDev st,
                                  double ratio, uint32_t taps, PlanSeg *segs, CallCounts *calls,
                                  uint32_t tile_out, uint32_t *tile_total) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    UnitDev U = units[u];
    PlanState s;
    s.position = st.position[U.rep_stream];
    s.available = st.hist_len[U.rep_stream];
    const uint32_t hist0 = s.available;
    SegSink sink{segs + U.seg_off, U.seg_cap, 0u, 0u, 0};
    uint64_t offset = 0;     // frames handed over so far (the caller loop's input_offset)
    int64_t advanced = 0;    // frames the read position has moved (virtual frame of position 0)
    uint32_t n_calls = 0;
    uint32_t status = 0;
    DivCache dc;
    div_cache_reset(dc);
    for (;;) {
        if (!U.single_call && offset >= U.total_frames) break;
        if (n_calls >= U.max_calls) { status = 2; break; }
        const uint64_t remaining = U.total_frames - offset;
        const uint32_t chunk =
            U.single_call ? (uint32_t)(remaining > 0xffffffffull ? 0xffffffffull : remaining)
                          : (uint32_t)(remaining < U.call_frames ? remaining : U.call_frames);
        sink.vbase = advanced;
        const CallResult r = plan_call(s, ratio, taps, chunk, U.cap_frames, sink, dc);
        if (n_calls < U.call_cap) {
            CallCounts cc;
            cc.copied = r.copied;
            cc.produced = r.produced;
            calls[U.call_off + n_calls] = cc;
        }
        n_calls += 1;
        offset += r.copied;
        advanced += r.advanced;
        if (U.single_call || r.copied == 0) break;
    }
    if (sink.n > U.seg_cap) status = 1;
    const uint32_t n_tiles = (uint32_t)((sink.out0 + tile_out - 1) / tile_out);
    uint32_t tile_off = 0;
    if (status != 1 && n_tiles) tile_off = atomicAdd(tile_total, n_tiles);
    UnitDev *o = units + u;
    o->total_out = sink.out0;
    o->total_copied = offset;
    o->n_calls = n_calls;
    o->n_segs = sink.n < U.seg_cap ? sink.n : U.seg_cap;
    o->n_tiles = status == 1 ? 0u : n_tiles;
    o->tile_off = tile_off;
    o->status = status;
    o->hist_len0 = hist0;
    o->final_available = s.available;
    o->final_position = s.position;
}

void launch_plan(UnitDev *units, uint32_t n_units, StreamStateDev st, double ratio, uint32_t taps,
                 PlanSeg *segs, CallCounts *calls, uint32_t tile_out, uint32_t *tile_total,
                 cudaStream_t stream) {
    const uint32_t threads = 32;
    plan_units_kernel<<<(n_units + threads - 1) / threads, threads, 0, stream>>>(
        units, n_units, st, ratio, taps, segs, calls, tile_out, tile_total);
}

// ---------------------------------------------------------------------------
// tile records
// ---------------------------------------------------------------------------
// One warp per tile: lane 0 finds the tile's first segment by binary search, every lane
// expands one output frame of the tile (exact position -> offset, phase, frac), and -- for the
// fast kernel -- the warp builds the tile's banded filter matrix G[kTileOut][gs] in global
// memory:  G[k][(v_k - v_base) + t] = c[phase1_k][t]*(1-frac_k) + c[phase2_k][t]*frac_k, zero
// elsewhere.  Every stream group that processes the tile fetches it with one TMA bulk copy.
template <int TAPS>
__global__ void tile_index_kernel(const UnitDev *units, const PlanSeg *segs, TileRec *tiles,
                                  PlanEntry *entries, const float *coeffs, float *gtiles,
                                  uint32_t gs) {
    const uint32_t u = blockIdx.x;
    const UnitDev &U = units[u];
    const uint32_t n_tiles = U.n_tiles;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t t = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= n_tiles) return;
    const uint32_t o0 = t * kTileOut;
    const PlanSeg *sg = segs + U.seg_off;
    // last segment with out0 <= o0
    uint32_t lo = 0, hi = U.n_segs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sg[mid].out0 <= o0) lo = mid; else hi = mid;
    }
    const uint64_t left = U.total_out - o0;
    const uint32_t n_out = (uint32_t)(left < kTileOut ? left : kTileOut);
    PlanEntry e;
    e.v = 0; e.phase1 = 0; e.frac = 0.f; e.off = 0;
    if (lane < n_out) {
        uint32_t s = lo;
        PlanSeg cur = sg[s];
        const uint32_t o = o0 + lane;
        while (o >= cur.out0 + cur.n) cur = sg[++s];
        const double pos = bits2d(cur.base_bits + (int64_t)(o - cur.out0) * cur.step_bits);
        const PhasePoint pp = phase_point(pos);
        e.v = (int32_t)(cur.vbase + (int64_t)pp.off);
        e.phase1 = pp.phase1;
        e.frac = pp.frac;
        e.off = pp.off;
    }
    const size_t tile = (size_t)U.tile_off + t;
    entries[tile * kTileOut + lane] = e;
    // window: starts at the first frame needed, aligned so that (v_base - H) % 4 == 0
    const int32_t H = (int32_t)U.hist_len0;
    const int32_t v_first = __shfl_sync(0xffffffffu, e.v, 0);
    const int32_t v_last = __shfl_sync(0xffffffffu, e.v, (int)n_out - 1);
    const int32_t v_base = v_first - ((((v_first - H) % 4) + 4) % 4);
    const uint32_t winp = (uint32_t)((v_last - v_base) + TAPS + 3) & ~3u;
    if (lane == 0) {
        TileRec r;
        r.unit = u;
        r.o_start = o0;
        r.seg = U.seg_off + lo;
        r.n_out = n_out;
        r.v_base = v_base;
        r.winp = winp;
        r.n_members = U.n_members;
        r.member_off = U.member_off;
        r.hist_len0 = U.hist_len0;
        r.pad0 = 0;
        r.total_frames = U.total_frames;
        tiles[tile] = r;
    }
    if (gtiles == nullptr) return;
    // ---- banded rows ----
    constexpr int kQ = TAPS / 4;                       // float4 per coefficient row
    float4 *grow = reinterpret_cast<float4 *>(gtiles + tile * kTileOut * gs);
    const uint32_t gq = gs >> 2;                       // float4 per G row
    for (uint32_t k = 0; k < kTileOut; ++k, grow += gq) {
        const uint32_t p1 = __shfl_sync(0xffffffffu, e.phase1, (int)k);
        const float fr = __shfl_sync(0xffffffffu, e.frac, (int)k);
        const int32_t vk = __shfl_sync(0xffffffffu, e.v, (int)k);
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < n_out && lane < kQ) {
            const uint32_t p2 = p1 + 1 < kPhases - 1 ? p1 + 1 : kPhases - 1;
            const float4 a = __ldg(reinterpret_cast<const float4 *>(coeffs + (size_t)p1 * TAPS) + lane);
            const float4 b = __ldg(reinterpret_cast<const float4 *>(coeffs + (size_t)p2 * TAPS) + lane);
            const float omf = __fsub_rn(1.0f, fr);
            h.x = __fmaf_rn(b.x, fr, __fmul_rn(a.x, omf));
            h.y = __fmaf_rn(b.y, fr, __fmul_rn(a.y, omf));
            h.z = __fmaf_rn(b.z, fr, __fmul_rn(a.z, omf));
            h.w = __fmaf_rn(b.w, fr, __fmul_rn(a.w, omf));
        }
        const int d = k < n_out ? (int)(vk - v_base) : 0;
        const int r = d & 3, q0 = d >> 2;
        // output float4 q holds taps 4*(q - q0) - r .. +3: parts of h[q - q0 - 1] and h[q - q0]
        for (uint32_t qb = 0; qb < gq; qb += 32) {
            const int q = (int)(qb + lane);
            const int src = q - q0;                    // coefficient float4 index of the upper part
            const float4 hi4 = make_float4(__shfl_sync(0xffffffffu, h.x, src & 31),
                                           __shfl_sync(0xffffffffu, h.y, src & 31),
                                           __shfl_sync(0xffffffffu, h.z, src & 31),
                                           __shfl_sync(0xffffffffu, h.w, src & 31));
            const float4 lo4 = make_float4(__shfl_sync(0xffffffffu, h.x, (src - 1) & 31),
                                           __shfl_sync(0xffffffffu, h.y, (src - 1) & 31),
                                           __shfl_sync(0xffffffffu, h.z, (src - 1) & 31),
                                           __shfl_sync(0xffffffffu, h.w, (src - 1) & 31));
            const bool hi_ok = src >= 0 && src < kQ, lo_ok = src - 1 >= 0 && src - 1 < kQ;
            const float hv[4] = {hi_ok ? hi4.x : 0.f, hi_ok ? hi4.y : 0.f, hi_ok ? hi4.z : 0.f,
                                 hi_ok ? hi4.w : 0.f};
            const float lv[4] = {lo_ok ? lo4.x : 0.f, lo_ok ? lo4.y : 0.f, lo_ok ? lo4.z : 0.f,
                                 lo_ok ? lo4.w : 0.f};
            float o4[4];
#pragma unroll
            for (int el = 0; el < 4; ++el) {
                // tap = 4*src + el - r
                const int idx = el - r;                // index into hv (>= 0) or lv (< 0)
                o4[el] = idx >= 0 ? hv[idx & 3] : lv[(idx + 4) & 3];
            }
            if ((uint32_t)q < gq) grow[q] = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
    }
}

void launch_tiles(const UnitDev *units, uint32_t n_units, const PlanSeg *segs, TileRec *tiles,
                  PlanEntry *entries, uint32_t max_tiles_per_unit, uint32_t taps,
                  const float *coeffs, float *gtiles, uint32_t gs, cudaStream_t stream) {
    if (n_units == 0 || max_tiles_per_unit == 0) return;
    const uint32_t threads = 128, per_cta = threads / 32;
    // gridDim.y is limited to 65535: units in x, tile chunks in y
    uint32_t chunks = (max_tiles_per_unit + per_cta - 1) / per_cta;
    if (chunks > 65535u) chunks = 65535u;   // callers bound tiles per unit below 65535*4
    dim3 grid(n_units, chunks);
    switch (taps) {
        case 16: tile_index_kernel<16><<<grid, threads, 0, stream>>>(units, segs, tiles, entries, coeffs, gtiles, gs); break;
        case 32: tile_index_kernel<32><<<grid, threads, 0, stream>>>(units, segs, tiles, entries, coeffs, gtiles, gs); break;
        case 64: tile_index_kernel<64><<<grid, threads, 0, stream>>>(units, segs, tiles, entries, coeffs, gtiles, gs); break;
        default: tile_index_kernel<128><<<grid, threads, 0, stream>>>(units, segs, tiles, entries, coeffs, gtiles, gs); break;
    }
}

// ---------------------------------------------------------------------------
// exact convolution: fir/avx512.rs:5-50 mapped lane for lane onto half-warps
// ---------------------------------------------------------------------------
// 16 lanes of a half-warp are the 16 lanes of the zmm registers: lane l owns
// acc1[l], acc2[l] (FMA chains over taps/16 steps, :36-37), blends them with
// separately rounded mul, mul, add (:41-45), and the half-warp reduces with the
// halving tree of _mm512_reduce_add_ps (:48) via xor-shuffles 8, 4, 2, 1.
template <int TAPS>
__global__ void __launch_bounds__(256) conv_exact_kernel(ConvParams P) {
    __shared__ int64_t s_v[kTileOut];
    __shared__ uint32_t s_p1[kTileOut];
    __shared__ float s_frac[kTileOut];

    const uint32_t ch = P.channels;
    const uint32_t n_items = *P.tile_total * P.groups;
    const uint32_t lane16 = threadIdx.x & 15u;
    const uint32_t half = (threadIdx.x >> 4) & 1u;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t n_warps = blockDim.x >> 5;

    for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        const uint32_t t = w / P.groups, g = w - t * P.groups;
        const TileRec rec = P.tiles[t];
        const UnitDev &U = P.units[rec.unit];
        const uint32_t m0 = g * P.streams_per_group;
        if (m0 >= U.n_members) continue;
        const uint32_t nm = min(P.streams_per_group, U.n_members - m0);
        const int64_t H = (int64_t)U.hist_len0;

        __syncthreads();
        if (threadIdx.x < rec.n_out) {
            const PlanEntry e = P.entries[(size_t)t * kTileOut + threadIdx.x];
            s_v[threadIdx.x] = e.v;
            s_p1[threadIdx.x] = e.phase1;
            s_frac[threadIdx.x] = e.frac;
        }
        __syncthreads();

        const uint32_t per_stream = rec.n_out * ch;
        const uint32_t total = nm * per_stream;
        for (uint32_t base = warp * 2; base < total; base += n_warps * 2) {
            const uint32_t idx_raw = base + half;
            const bool valid = idx_raw < total;
            const uint32_t idx = valid ? idx_raw : total - 1;
            const uint32_t m = idx / per_stream;
            const uint32_t rem = idx - m * per_stream;
            const uint32_t k = rem / ch;
            const uint32_t c = rem - k * ch;
            const JobDev job = P.jobs[U.member_off + m0 + m];
            const float *hist = job.hist;
            const int64_t v = s_v[k];
            const uint32_t p1 = s_p1[k];
            const uint32_t p2 = p1 + 1 < kPhases - 1 ? p1 + 1 : kPhases - 1;
            const float frac = s_frac[k];
            const float *c1 = P.coeffs + (size_t)p1 * TAPS;
            const float *c2 = P.coeffs + (size_t)p2 * TAPS;
            float acc1 = 0.0f, acc2 = 0.0f;
#pragma unroll
            for (int i = 0; i < TAPS / 16; ++i) {
                const int64_t vv = v + i * 16 + lane16;
                const float x = vv < H ? hist[((int64_t)kHistFrames - H + vv) * ch + c]
                                       : job.in[(vv - H) * ch + c];
                acc1 = __fmaf_rn(__ldg(c1 + i * 16 + lane16), x, acc1);
                acc2 = __fmaf_rn(__ldg(c2 + i * 16 + lane16), x, acc2);
            }
            const float omf = __fsub_rn(1.0f, frac);
            float s = __fadd_rn(__fmul_rn(acc1, omf), __fmul_rn(acc2, frac));
            s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 8));
            s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));
            s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
            s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
            const uint64_t o = (uint64_t)rec.o_start + k;
            if (valid && lane16 == 0 && o < job.out_capacity) job.out[o * ch + c] = s;
        }
    }
}

void launch_conv_exact(const ConvParams &p, uint32_t max_items, int sm_count,
                       cudaStream_t stream) {
    if (max_items == 0) return;
    uint32_t grid = (uint32_t)sm_count * 8u;
    if (grid > max_items) grid = max_items;
    switch (p.taps) {
        case 16: conv_exact_kernel<16><<<grid, 256, 0, stream>>>(p); break;
        case 32: conv_exact_kernel<32><<<grid, 256, 0, stream>>>(p); break;
        case 64: conv_exact_kernel<64><<<grid, 256, 0, stream>>>(p); break;
        default: conv_exact_kernel<128><<<grid, 256, 0, stream>>>(p); break;
    }
}

// ---------------------------------------------------------------------------
// state write-back
// ---------------------------------------------------------------------------
__global__ void update_state_kernel(const UnitDev *units, const JobDev *jobs, StreamStateDev st,
                                    uint32_t ch) {
    const JobDev job = jobs[blockIdx.x];
    const UnitDev &U = units[job.unit];
    const float *old_hist = job.hist;
    float *new_hist = job.hist_next;
    const int64_t H0 = U.hist_len0, H1 = U.final_available;
    const int64_t v_end = H0 + (int64_t)U.total_copied;
    const int64_t v0 = v_end - H1;   // first virtual frame that stays buffered
    const int64_t n_vals = H1 * ch;
    for (int64_t i = threadIdx.x; i < n_vals; i += blockDim.x) {
        const int64_t f = i / ch, c = i - f * ch;
        const int64_t vv = v0 + f;
        const float x = vv < H0 ? old_hist[((int64_t)kHistFrames - H0 + vv) * ch + c]
                                : job.in[(vv - H0) * ch + c];
        new_hist[((int64_t)kHistFrames - H1 + f) * ch + c] = x;
    }
}

// Scalar state of every stream of the submit (runs on the plan stream right after the plan).
__global__ void state_scalars_kernel(const UnitDev *units, const JobDev *jobs, uint32_t n_jobs,
                                     StreamStateDev st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_jobs) return;
    const JobDev &job = jobs[i];
    const UnitDev &U = units[job.unit];
    st.position[job.stream] = U.final_position;      // resampler_fir.rs:602
    st.hist_len[job.stream] = U.final_available;     // resampler_fir.rs:601
}

void launch_state_scalars(const UnitDev *units, const JobDev *jobs, uint32_t n_jobs,
                          StreamStateDev st, cudaStream_t stream) {
    if (n_jobs == 0) return;
    state_scalars_kernel<<<(n_jobs + 255) / 256, 256, 0, stream>>>(units, jobs, n_jobs, st);
}

void launch_update(const UnitDev *units, const JobDev *jobs, uint32_t n_jobs, StreamStateDev st,
                   uint32_t channels, cudaStream_t stream) {
    if (n_jobs == 0) return;
    update_state_kernel<<<n_jobs, 128, 0, stream>>>(units, jobs, st, channels);
}

__global__ void reset_state_kernel(StreamStateDev st, uint32_t first, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    st.position[first + i] = 0.0;     // resampler_fir.rs:638-642
    st.hist_len[first + i] = 0u;
}

void launch_reset(StreamStateDev st, uint32_t first, uint32_t count, cudaStream_t stream) {
    if (count == 0) return;
    reset_state_kernel<<<(count + 255) / 256, 256, 0, stream>>>(st, first, count);
}

// ---------------------------------------------------------------------------
// plan expansion (debug / parity of phase indices)
// ---------------------------------------------------------------------------
__global__ void expand_plan_kernel(const UnitDev *units, uint32_t unit, const PlanEntry *entries,
                                   uint32_t n_frames, uint32_t *off, uint32_t *p1, uint32_t *p2,
                                   uint32_t *fb) {
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_frames) return;
    const UnitDev &U = units[unit];
    // the very entries the convolution kernels consume
    const PlanEntry e = entries[(size_t)(U.tile_off + o / kTileOut) * kTileOut + (o % kTileOut)];
    if (off) off[o] = e.off;
    if (p1) p1[o] = e.phase1;
    if (p2) p2[o] = e.phase1 + 1 < kPhases - 1 ? e.phase1 + 1 : kPhases - 1;
    if (fb) fb[o] = __float_as_uint(e.frac);
}

void launch_expand_plan(const UnitDev *units, uint32_t unit, const PlanEntry *entries,
                        uint32_t n_frames, uint32_t *off, uint32_t *p1, uint32_t *p2, uint32_t *fb,
                        cudaStream_t stream) {
    if (n_frames == 0) return;
    expand_plan_kernel<<<(n_frames + 255) / 256, 256, 0, stream>>>(units, unit, entries, n_frames,
                                                                    off, p1, p2, fb);
}

// ---------------------------------------------------------------------------
// synthetic input
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void fill_synthetic_kernel(float *dst, uint32_t first_stream, uint32_t n_streams,
                                      uint64_t frames, uint32_t ch, uint32_t rate, uint64_t seed) {
    const uint64_t per_stream = frames * ch;
    const uint64_t total = per_stream * n_streams;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s_local = i / per_stream;
        const uint64_t r = i - s_local * per_stream;
        const uint64_t n = r / ch;
        const uint32_t c = (uint32_t)(r - n * ch);
        const uint64_t s = first_stream + s_local;
        const uint64_t f = 110ull * (1ull + (s % 64ull)) + 7ull * c;   // Hz
        const uint64_t ph = (f * n) % rate;                            // exact phase numerator
        const float sine = sinpif(2.0f * (float)ph / (float)rate);
        const uint64_t h = splitmix64(seed ^ (s << 32) ^ ((uint64_t)c << 62) ^ n);
        const float u = (float)(h >> 40) * (1.0f / 8388608.0f) - 1.0f;   // [-1, 1)
        dst[i] = 0.5f * sine + 0.25f * u;
    }
}

void launch_fill_synthetic(float *dst, uint32_t first_stream, uint32_t n_streams, uint64_t frames,
                           uint32_t channels, uint32_t rate_hz, uint64_t seed,
                           cudaStream_t stream) {
    const uint64_t total = frames * channels * n_streams;
    if (total == 0) return;
    uint64_t blocks = (total + 255) / 256;
    if (blocks > 148ull * 32ull) blocks = 148ull * 32ull;
    fill_synthetic_kernel<<<(uint32_t)blocks, 256, 0, stream>>>(dst, first_stream, n_streams, frames,
                                                                channels, rate_hz, seed);
}

}  // namespace rsb
