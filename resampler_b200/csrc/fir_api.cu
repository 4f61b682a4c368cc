// fir_api.cu -- host library behind include/resampler_b200.h.
//
// Mirrors the reference's ResamplerFir (src/resampler_fir.rs) for N streams on
// one GPU.  All streaming state (history frames, f64 position, buffered frame
// count) lives on the device; the host keeps only a "cohort" id per stream so
// that streams whose state and call sequence are identical share one plan.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

// planner.h templates are __host__ __device__; the host-only sink below is fine on the host
#pragma nv_diag_suppress 20011
#pragma nv_diag_suppress 20014

#include "../../include/resampler_b200.h"
#include "filter_design.h"
#include "filter_design_device.h"
#include "fir_kernels.h"
#include "fir_submit.h"
#include "pcm_ingest.h"

namespace {

thread_local std::string g_last_error;
std::atomic<bool> g_device_design{false};   // rsb_set_device_filter_design

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

#define RSB_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            return fail(e__ == cudaErrorMemoryAllocation ? RSB_ERR_OUT_OF_MEMORY         \
                                                         : RSB_ERR_CUDA,                 \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));            \
        }                                                                                \
    } while (0)

// grow-only device / pinned buffers
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

struct Pending {
    bool active = false;
    uint32_t n = 0;
    uint32_t n_units = 0;
    size_t *consumed = nullptr;
    size_t *produced = nullptr;
    uint32_t *n_calls = nullptr;
    bool check_capacity = false;
};

// Everything one submit needs on the device and in pinned memory.  Two of them alternate so
// that the (serial) plan of submit i+1 can run on the plan stream while the convolution of
// submit i is still running on the main stream.
struct Workspace {
    DevBuf d_units, d_jobs, d_segs, d_calls, d_tiles, d_entries, d_gtiles, d_tct, d_counter;
    PinBuf h_units, h_jobs, h_units_back, h_calls_back, h_segs, h_counter;
    DevBuf d_pcm_jobs;   // job table of the PCM format step (rsb_fir_process_pcm_batch)
    PinBuf h_pcm_jobs;
    std::vector<uint32_t> job_unit;          // unit of each job of the batch
    std::vector<uint32_t> job_stream;        // stream of each job of the batch
    bool host_planned = false;               // the plan ran on the host: results already mirrored
    uint64_t seq = 0;                        // submit number that last used this workspace
    std::vector<uint64_t> job_out_capacity;  // frames
    uint32_t n_units = 0;
    bool has_calls = false, has_plan = false;
    Pending pending;
    cudaEvent_t ev_plan = nullptr;   // plan + tile kernels and result read-back done
    cudaEvent_t ev_done = nullptr;   // convolution + state update done
};

}  // namespace

struct rsb_fir {
    int device = 0;
    int sm_count = 148;
    size_t mem_pitch = 0;    // largest pitch cudaMemcpy2D accepts (cudaDeviceProp::memPitch)
    cudaStream_t stream = nullptr;
    uint32_t n_streams = 0, channels = 0, in_hz = 0, out_hz = 0, taps = 0;
    int latency = 0, attenuation = 0;
    double ratio = 0.0;
    int kernel_mode = RSB_KERNEL_AUTO;
    int last_kernel = RSB_KERNEL_AUTO;   // kernel the most recent batch ran on
    std::shared_ptr<const rsb::FirTable> table;
    float *d_coeffs = nullptr;
    rsb::StreamStateDev st{};
    std::vector<uint64_t> cohort;   // host-side plan cohort of each stream (0 = fresh state)
    std::vector<uint8_t> hist_sel;  // which history buffer of a stream is live (host mirror)
    // Host mirror of the scalar stream state (resampler_fir.rs:191-193).  With it a batch of few
    // plan units is planned on a host core (one 60 s unit: 1.2 ms, the same code takes 15 ms
    // on one GPU thread); batches of many units are planned on the GPU, one thread per unit,
    // and the mirror is refreshed from their read-back.
    std::vector<double> m_pos;
    std::vector<uint32_t> m_avail;
    std::vector<uint8_t> m_ok;
    std::vector<uint64_t> m_pending;   // submit number whose read-back will refresh the mirror
    uint64_t next_cohort = 1;
    uint64_t launches = 0;
    int deferred_rc = 0;               // failure of an asynchronous submit found while starting a later one
    std::string deferred_msg;
    cudaStream_t plan_stream = nullptr;   // plan / tile kernels and the state scalars
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_sync = nullptr;
    // ring of event pairs bracketing the convolution kernel of the most recent batches
    static constexpr int kConvRing = 64;
    cudaEvent_t ev_conv[kConvRing][2] = {};
    uint64_t conv_batches = 0;

    // Tensor-kernel plan cache: the tile records and the G store depend only on (start state of
    // the cohort, call signature), not on sample data; a batch whose plan key equals the cached
    // one (e.g. the same batch shape after reset()) reuses them and skips the builders.
    struct Tc2PlanKey {
        uint64_t pos_bits = 0, total_frames = 0;
        uint32_t avail = 0, call_frames = 0, cap_frames = 0, single = 0;
        bool valid = false;
        bool operator==(const Tc2PlanKey &o) const {
            return valid && o.valid && pos_bits == o.pos_bits && total_frames == o.total_frames &&
                   avail == o.avail && call_frames == o.call_frames && cap_frames == o.cap_frames &&
                   single == o.single;
        }
    } tc2_key;
    DevBuf d_tct2, d_gmat2;          // 64-output tile records, fp16 G store (+ MMA lists) of tc2_key
    uint64_t tc2_cache_hits = 0;
    // fused submits (fir_submit.cu): kFusedSlots of them in flight, so that the host runs ahead of the
    // GPU and the next submits' job tables and planner kernels are queued underneath the running
    // convolution (with two slots the host could only start on submit i once i - 2 had finished, and
    // its enqueue + the plan stage then sat on the critical path: 69 us per 4096-call submit against
    // a 45 us convolution); results are delivered when a slot is reused or at sync
    static constexpr uint32_t kFusedSlots = 4;
    struct FusedSlot {
        PinBuf h_jobs, h_res;
        DevBuf d_jobs, d_res, d_segs;
        cudaEvent_t ev_done = nullptr, ev_plan = nullptr;
        bool active = false;
        uint32_t n = 0;
        uint64_t seq = 0;
        std::vector<uint32_t> streams;
        size_t *consumed = nullptr, *produced = nullptr;
    } fused[kFusedSlots];
    uint64_t fused_count = 0;
    std::vector<uint32_t> seen_epoch;   // duplicate-stream check without clearing an array per submit
    uint32_t epoch = 0;
    std::vector<uint64_t> m_fused_seq;  // fused submit whose result will refresh a stream's mirror
    // host-memspace pipeline of rsb_fir_process_batch: time slices, H2D / kernels / D2H overlapped
    struct HostPipe {
        cudaStream_t s_in = nullptr, s_out = nullptr;
        cudaEvent_t ev_h2d[2] = {}, ev_conv[2] = {}, ev_d2h[2] = {};
        DevBuf d_in[2], d_out[2];
        std::vector<rsb::CallCounts> calls;     // dry-run plan of the whole batch
        std::vector<rsb::PlanSeg> seg_scratch;
        uint64_t batches = 0, slices = 0;
    } pipe;
    Workspace ws[2];
    uint64_t submits = 0;             // ws[submits & 1] is the next one to use
    DevBuf d_stage_in, d_stage_out, d_dbg;   // host-memspace staging (those calls are synchronous)
    // PCM format step (rsb_fir_process_pcm_batch): raw samples of host-memspace calls
    DevBuf d_pcm_raw, d_zero;
    cudaEvent_t ev_pcm[2] = {};
    bool pcm_timed = false;
    bool pcm_fused_last = false;   // the last PCM batch ran with the format step inside the tensor kernel
    Workspace &last_ws() { return ws[(submits + 1) & 1]; }
};

namespace {

using rsb::JobDev;
using rsb::UnitDev;

struct UnitKey {
    uint64_t cohort, total_frames;
    uint32_t call_frames, cap_frames, single;
    bool operator<(const UnitKey &o) const {
        return std::tie(cohort, total_frames, call_frames, cap_frames, single) <
               std::tie(o.cohort, o.total_frames, o.call_frames, o.cap_frames, o.single);
    }
};

struct JobHost {
    uint32_t stream;
    const float *in;
    float *out;
    uint64_t total_frames;     // offered frames
    uint64_t out_capacity;     // frames
    uint32_t call_frames;      // frames per call (ignored when single)
    uint32_t cap_frames;       // output capacity per call, frames
};

// Raw-sample input of a batch (rsb_fir_process_pcm_batch): JobHost::in then points at raw
// samples, which the format step converts into the f32 staging buffer the kernels read.
struct PcmSpec {
    int format;
    uint32_t dup;   // f32 values written per source value (mono source: the handle's channels)
    uint32_t bps;   // bytes per source sample
};

uint32_t segs_per_call_bound(double ratio) {
    int e = 0;
    std::frexp(ratio, &e);              // ratio = m * 2^e, m in [0.5, 1)
    int lowest = std::min(0, e - 1);    // lowest binade the accumulator visits after one step
    return (uint32_t)(2 * (14 - lowest) + 6);
}

// Host twin of plan_units_kernel (fir_kernels.cu): the same planner.h code on a host core.
struct BoundedSink {
    rsb::PlanSeg *segs;
    uint32_t cap;
    uint32_t n;
    uint32_t out0;
    int64_t vbase;
    void seg(int64_t base_bits, int64_t step_bits, uint32_t cnt) {
        if (n < cap) {
            rsb::PlanSeg sg;
            sg.base_bits = base_bits;
            sg.step_bits = step_bits;
            sg.n = cnt;
            sg.out0 = out0;
            sg.vbase = vbase;
            segs[n] = sg;
        }
        n += 1;
        out0 += cnt;
    }
};

void plan_unit_host(UnitDev &U, rsb::PlanState s, double ratio, uint32_t taps, rsb::PlanSeg *segs,
                    rsb::CallCounts *calls, uint32_t tile_out, uint64_t &tile_total) {
    const uint32_t hist0 = s.available;
    BoundedSink sink{segs + U.seg_off, U.seg_cap, 0u, 0u, 0};
    uint64_t offset = 0;
    int64_t advanced = 0;
    uint32_t n_calls = 0, status = 0;
    rsb::DivCache dc;
    rsb::div_cache_reset(dc);
    for (;;) {
        if (!U.single_call && offset >= U.total_frames) break;
        if (n_calls >= U.max_calls) { status = 2; break; }
        const uint64_t remaining = U.total_frames - offset;
        const uint32_t chunk =
            U.single_call ? (uint32_t)std::min<uint64_t>(remaining, 0xffffffffull)
                          : (uint32_t)std::min<uint64_t>(remaining, U.call_frames);
        sink.vbase = advanced;
        const rsb::CallResult r = rsb::plan_call(s, ratio, taps, chunk, U.cap_frames, sink, dc);
        if (calls && n_calls < U.call_cap) {
            rsb::CallCounts cc;
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
    const uint32_t n_tiles = (uint32_t)(((uint64_t)sink.out0 + tile_out - 1) / tile_out);
    uint32_t tile_off = 0;
    if (status != 1 && n_tiles) {
        tile_off = (uint32_t)tile_total;
        tile_total += n_tiles;
    }
    U.total_out = sink.out0;
    U.total_copied = offset;
    U.n_calls = n_calls;
    U.n_segs = sink.n < U.seg_cap ? sink.n : U.seg_cap;
    U.n_tiles = status == 1 ? 0u : n_tiles;
    U.tile_off = tile_off;
    U.status = status;
    U.hist_len0 = hist0;
    U.final_available = s.available;
    U.final_position = s.position;
}

// batches with at most this many plan units are planned on the host
constexpr uint32_t kHostPlanMaxUnits = 8;

int finalize_fused_all(rsb_fir *h);

int finalize_pending(rsb_fir *h, Workspace &W) {
    if (!W.pending.active) return RSB_OK;
    Pending p = W.pending;
    W.pending.active = false;
    RSB_CUDA(cudaEventSynchronize(W.ev_plan));
    const UnitDev *ub = W.h_units_back.as<UnitDev>();
    int rc = RSB_OK;
    for (uint32_t u = 0; u < p.n_units; ++u) {
        if (ub[u].status != 0)
            rc = fail(RSB_ERR_PLAN_OVERFLOW, "plan workspace bound exceeded (status " +
                                                 std::to_string(ub[u].status) + ")");
    }
    const uint32_t ch = h->channels;
    for (uint32_t i = 0; i < p.n; ++i) {
        const UnitDev &U = ub[W.job_unit[i]];
        if (!W.host_planned && W.job_stream.size() == p.n && !h->m_ok[W.job_stream[i]] &&
            h->m_pending[W.job_stream[i]] == W.seq) {
            // the newest GPU-planned batch of this stream: its read-back is the stream's state
            h->m_pos[W.job_stream[i]] = U.final_position;
            h->m_avail[W.job_stream[i]] = U.final_available;
            h->m_ok[W.job_stream[i]] = 1;
        }
        if (p.consumed) p.consumed[i] = (size_t)U.total_copied * ch;
        if (p.produced) p.produced[i] = (size_t)U.total_out * ch;
        if (p.n_calls) p.n_calls[i] = U.n_calls;
        if (p.check_capacity && U.total_out > W.job_out_capacity[i] && rc == RSB_OK)
            rc = fail(RSB_ERR_OUTPUT_CAPACITY, "output buffer of job " + std::to_string(i) +
                                                   " too small: produced " +
                                                   std::to_string(U.total_out) + " frames");
    }
    return rc;
}

int finalize_all(rsb_fir *h) {
    // older submit first
    int rc = finalize_pending(h, h->ws[h->submits & 1]);
    int rc2 = finalize_pending(h, h->ws[(h->submits + 1) & 1]);
    return rc != RSB_OK ? rc : rc2;
}

// Core: runs `jobs` (already validated) on the device.
int run_batch(rsb_fir *h, const std::vector<JobHost> &jobs, bool single, int memspace,
              uint32_t flags, size_t *consumed, size_t *produced, uint32_t *n_calls_out,
              const PcmSpec *pcm = nullptr) {
    RSB_CUDA(cudaSetDevice(h->device));
    {   // fused submits in flight: their read-back makes the host mirror valid again
        const int rf = finalize_fused_all(h);
        if (rf != RSB_OK) return rf;
    }
    Workspace &W = h->ws[h->submits & 1];
    // this workspace was last used two submits ago: collect its counts, wait for its kernels.  A
    // failure of that older (asynchronous) submit is not this batch's: it is kept for
    // rsb_fir_sync() and the current batch still runs.
    int rc = finalize_pending(h, W);
    if (rc != RSB_OK && h->deferred_rc == RSB_OK) {
        h->deferred_rc = rc;
        h->deferred_msg = g_last_error;
    }
    rc = RSB_OK;
    RSB_CUDA(cudaEventSynchronize(W.ev_done));
    const uint32_t n = (uint32_t)jobs.size();
    const uint32_t ch = h->channels;
    if (n == 0) return RSB_OK;

    // ---- group jobs into plan units ----
    std::map<UnitKey, uint32_t> key_to_unit;
    std::vector<UnitKey> unit_keys;
    std::vector<uint32_t> unit_count;
    W.job_unit.assign(n, 0);
    W.job_out_capacity.assign(n, 0);
    for (uint32_t i = 0; i < n; ++i) {
        UnitKey k{h->cohort[jobs[i].stream], jobs[i].total_frames,
                  single ? 0u : jobs[i].call_frames, jobs[i].cap_frames, single ? 1u : 0u};
        auto it = key_to_unit.find(k);
        uint32_t u;
        if (it == key_to_unit.end()) {
            u = (uint32_t)unit_keys.size();
            key_to_unit.emplace(k, u);
            unit_keys.push_back(k);
            unit_count.push_back(0);
        } else {
            u = it->second;
        }
        W.job_unit[i] = u;
        W.job_out_capacity[i] = jobs[i].out_capacity;
        unit_count[u] += 1;
    }
    const uint32_t n_units = (uint32_t)unit_keys.size();

    // ---- choose kernel + tile geometry ----
    bool use_fast = false;
    if (h->kernel_mode != RSB_KERNEL_EXACT && rsb::fast_supported(ch, h->taps, h->ratio)) {
        if (h->kernel_mode == RSB_KERNEL_FAST) {
            use_fast = true;
        } else {
            // AUTO: the fast kernel wants enough streams per plan to fill its column tile
            uint32_t biggest = *std::max_element(unit_count.begin(), unit_count.end());
            use_fast = (uint64_t)biggest * ch >= 32;
        }
    }
    if (use_fast && memspace == RSB_MEM_DEVICE && !pcm) {
        // the fast kernel stages input with 16-byte vector loads
        for (uint32_t i = 0; i < n; ++i)
            if ((reinterpret_cast<uintptr_t>(jobs[i].in) & 15u) != 0) {
                if (h->kernel_mode == RSB_KERNEL_FAST)
                    return fail(RSB_ERR_INVALID_ARGUMENT,
                                "fast kernel needs 16-byte aligned device input pointers");
                use_fast = false;
                break;
            }
    }
    const uint32_t tile_out = rsb::kTileOut;
    const uint32_t spg = use_fast ? rsb::fast_streams_per_group(ch, h->taps, h->ratio)
                                  : rsb::kExactStreamsPerGroup;

    // ---- size the workspace ----
    const uint32_t spc = segs_per_call_bound(h->ratio);
    const bool rec_calls = (flags & RSB_FLAG_RECORD_CALLS) != 0;
    RSB_CUDA(W.h_units.reserve(sizeof(UnitDev) * n_units));
    RSB_CUDA(W.h_units_back.reserve(sizeof(UnitDev) * n_units));
    RSB_CUDA(W.h_jobs.reserve(sizeof(JobDev) * n));
    UnitDev *hu = W.h_units.as<UnitDev>();
    uint64_t seg_total = 0, call_total = 0, tile_total = 0, max_tiles_unit = 0;
    uint32_t member_off = 0, max_groups = 1;
    for (uint32_t u = 0; u < n_units; ++u) {
        const UnitKey &k = unit_keys[u];
        UnitDev U;
        std::memset(&U, 0, sizeof(U));
        U.total_frames = k.total_frames;
        U.call_frames = k.call_frames;
        U.cap_frames = k.cap_frames;
        U.single_call = k.single;
        double out_bound_d =
            std::ceil(((double)rsb::kInputCapacity + (double)k.total_frames) / h->ratio) + 2.0;
        // a single call produces at most its output capacity (strong up-sampling: the frames the
        // buffered input could yield may be billions while the caller's buffer is small)
        if (k.single) out_bound_d = std::min(out_bound_d, (double)k.cap_frames + 2.0);
        if (out_bound_d > 4.0e9 || k.total_frames > 0x7ff00000ull)
            return fail(RSB_ERR_INVALID_ARGUMENT, "too many frames per stream in one batch; split it");
        const uint64_t out_bound = (uint64_t)out_bound_d;
        uint64_t max_calls = 1;
        if (!k.single) {
            const uint64_t full = k.call_frames ? (k.total_frames + k.call_frames - 1) / k.call_frames
                                                : 0;
            const uint64_t by_cap = k.cap_frames ? out_bound / k.cap_frames : 0;
            max_calls = 2 * full + by_cap + 4;
        }
        if (max_calls > 0x7fffffffull)
            return fail(RSB_ERR_INVALID_ARGUMENT, "too many calls in one batch");
        U.max_calls = (uint32_t)max_calls;
        // a call cannot produce more segments than output frames
        const uint64_t per_call = std::min<uint64_t>(
            spc, k.cap_frames ? std::min<uint64_t>(k.cap_frames, out_bound) : 0);
        const uint64_t seg_cap = std::min<uint64_t>(max_calls * per_call, out_bound) + 1;
        if (seg_total + seg_cap > 0xffffffffull)
            return fail(RSB_ERR_INVALID_ARGUMENT, "plan too large for one batch; split it");
        U.seg_off = (uint32_t)seg_total;
        U.seg_cap = (uint32_t)seg_cap;
        seg_total += seg_cap;
        if (rec_calls) {
            U.call_off = (uint32_t)call_total;
            U.call_cap = U.max_calls;
            call_total += U.max_calls;
        }
        const uint64_t tiles = (out_bound + tile_out - 1) / tile_out;
        if (tiles > 65535ull * 4ull)
            return fail(RSB_ERR_INVALID_ARGUMENT, "too many output frames per stream in one batch");
        U.tile_cap = (uint32_t)tiles;
        tile_total += tiles;
        max_tiles_unit = std::max(max_tiles_unit, tiles);
        U.member_off = member_off;
        U.n_members = 0;   // filled below
        member_off += unit_count[u];
        max_groups = std::max(max_groups, (unit_count[u] + spg - 1) / spg);
        hu[u] = U;
    }
    if (tile_total * max_groups > 0xffffffffull)
        return fail(RSB_ERR_INVALID_ARGUMENT, "too many work items in one batch; split it");

    // ---- host-memory staging ----
    const bool host_mem = memspace == RSB_MEM_HOST;
    const bool stage_in = host_mem || pcm != nullptr;   // the kernels read d_stage_in
    std::vector<size_t> in_off, out_off, in_vals, raw_off;
    size_t in_total = 0, out_total = 0, raw_total = 0;
    if (stage_in) {
        in_off.resize(n);
        out_off.resize(n);
        in_vals.resize(n);
        raw_off.resize(n);
        for (uint32_t i = 0; i < n; ++i) {
            uint64_t frames = jobs[i].total_frames;
            if (single) frames = std::min<uint64_t>(frames, rsb::kInputCapacity);   // :526-528
            in_vals[i] = (size_t)frames * ch;
            in_off[i] = in_total;
            in_total += (in_vals[i] + 3) & ~(size_t)3;    // keep 16-byte alignment
            out_off[i] = out_total;
            out_total += ((size_t)jobs[i].out_capacity * ch + 3) & ~(size_t)3;
            if (pcm) {
                raw_off[i] = raw_total;
                raw_total += (in_vals[i] / pcm->dup * pcm->bps + 15) & ~(size_t)15;
            }
        }
        RSB_CUDA(h->d_stage_in.reserve(in_total * sizeof(float) + 16));
        if (host_mem) RSB_CUDA(h->d_stage_out.reserve(out_total * sizeof(float) + 16));
        if (host_mem && pcm) RSB_CUDA(h->d_pcm_raw.reserve(raw_total + 16));
    }

    // jobs are stored grouped by unit: members of U are hj[U.member_off .. +U.n_members)
    JobDev *hj = W.h_jobs.as<JobDev>();
    const size_t hist_stride = (size_t)rsb::kHistFrames * ch;
    for (uint32_t i = 0; i < n; ++i) {
        JobDev J;
        J.stream = jobs[i].stream;
        J.unit = W.job_unit[i];
        J.out_capacity = jobs[i].out_capacity;
        J.in = stage_in ? h->d_stage_in.as<float>() + in_off[i] : jobs[i].in;
        J.out = host_mem ? h->d_stage_out.as<float>() + out_off[i] : jobs[i].out;
        const uint32_t sel = h->hist_sel[J.stream];
        J.hist = h->st.hist[sel] + hist_stride * J.stream;
        J.hist_next = h->st.hist[sel ^ 1u] + hist_stride * J.stream;
        UnitDev &U = hu[J.unit];
        hj[U.member_off + U.n_members] = J;
        if (U.n_members == 0) U.rep_stream = J.stream;
        U.n_members += 1;
    }

    // member inputs at one constant stride (and one plan unit): the batch's input is a 2-D tensor
    uint64_t in_stride = 0;
    bool in_uniform = false;
    if (n_units == 1 && (ch == 1 || ch == 2 || ch == 4 || ch == 8) && unit_keys[0].total_frames > 0) {
        const uintptr_t base = reinterpret_cast<uintptr_t>(hj[0].in);
        in_stride = n > 1 ? (uint64_t)(reinterpret_cast<uintptr_t>(hj[1].in) - base) : 0;
        in_uniform = n == 1 || reinterpret_cast<uintptr_t>(hj[1].in) > base;
        for (uint32_t i = 1; in_uniform && i < n; ++i)
            in_uniform = reinterpret_cast<uintptr_t>(hj[i].in) == base + (uint64_t)i * in_stride;
        if (n == 1) in_stride = ((uint64_t)unit_keys[0].total_frames * ch * 4 + 15) & ~15ull;
    }
    // tensor-core kernel: needs that tensor view (TMA streams the input chunk by chunk)
    bool use_tc = false;
    CUtensorMap tc_tmap;
    // AUTO: when at least half of a 128-row group is filled (below that the MMA's M is mostly idle)
    const bool want_tc = h->kernel_mode == RSB_KERNEL_TENSOR ||
                         (h->kernel_mode == RSB_KERNEL_AUTO && (uint64_t)n * ch >= 64);
    if (want_tc && in_uniform &&
        (rsb::tc2_supported(ch, h->taps, h->ratio) || rsb::tc_supported(ch, h->taps, h->ratio)))
        use_tc = rsb::tc_make_input_tensor_map(&tc_tmap, hj[0].in,
                                               getenv("RSB_DEBUG_FAKE_IN_STRIDE") ? (uint64_t)atoll(getenv("RSB_DEBUG_FAKE_IN_STRIDE")) : in_stride,
                                               unit_keys[0].total_frames, n, ch);

    RSB_CUDA(W.d_units.reserve(sizeof(UnitDev) * n_units));
    RSB_CUDA(W.d_jobs.reserve(sizeof(JobDev) * n));
    RSB_CUDA(W.d_segs.reserve(sizeof(rsb::PlanSeg) * seg_total));
    RSB_CUDA(W.d_tiles.reserve(sizeof(rsb::TileRec) * tile_total));
    RSB_CUDA(W.d_entries.reserve(sizeof(rsb::PlanEntry) * rsb::kTileOut * tile_total));
    const uint32_t gs = use_fast ? rsb::fast_row_stride(ch, h->taps, h->ratio) : 0;
    if (!use_tc && use_fast) RSB_CUDA(W.d_gtiles.reserve(sizeof(float) * rsb::kTileOut * gs * tile_total));
    RSB_CUDA(W.d_counter.reserve(sizeof(uint32_t) * 4));
    if (rec_calls) {
        RSB_CUDA(W.d_calls.reserve(sizeof(rsb::CallCounts) * std::max<uint64_t>(call_total, 1)));
        RSB_CUDA(W.h_calls_back.reserve(sizeof(rsb::CallCounts) * std::max<uint64_t>(call_total, 1)));
    }

    cudaStream_t s = h->stream;        // convolution, history update, data copies
    cudaStream_t sp = h->plan_stream;  // plan, tiles, state scalars, result read-back

    // ---- plan on a host core when the batch has few plan units whose state is mirrored ----
    bool host_plan = n_units <= kHostPlanMaxUnits;
    for (uint32_t u = 0; host_plan && u < n_units; ++u) host_plan = h->m_ok[hu[u].rep_stream] != 0;
    W.host_planned = host_plan;
    W.seq = h->submits;
    W.job_stream.resize(n);
    for (uint32_t i = 0; i < n; ++i) W.job_stream[i] = jobs[i].stream;
    uint64_t host_tiles = 0;
    rsb_fir::Tc2PlanKey plan_key;
    if (host_plan) {
        RSB_CUDA(W.h_segs.reserve(sizeof(rsb::PlanSeg) * seg_total));
        RSB_CUDA(W.h_counter.reserve(sizeof(uint32_t) * 4));
        rsb::PlanSeg *hs = W.h_segs.as<rsb::PlanSeg>();
        rsb::CallCounts *hc = rec_calls ? W.h_calls_back.as<rsb::CallCounts>() : nullptr;
        if (n_units == 1) {
            const uint32_t rep = hu[0].rep_stream;
            std::memcpy(&plan_key.pos_bits, &h->m_pos[rep], sizeof(uint64_t));
            plan_key.avail = h->m_avail[rep];
            plan_key.total_frames = hu[0].total_frames;
            plan_key.call_frames = hu[0].call_frames;
            plan_key.cap_frames = hu[0].cap_frames;
            plan_key.single = hu[0].single_call;
            plan_key.valid = true;
        }
        for (uint32_t u = 0; u < n_units; ++u) {
            const uint32_t rep = hu[u].rep_stream;
            plan_unit_host(hu[u], rsb::PlanState{h->m_pos[rep], h->m_avail[rep]}, h->ratio, h->taps,
                           hs, hc, tile_out, host_tiles);
        }
        std::memcpy(W.h_units_back.p, hu, sizeof(UnitDev) * n_units);
        for (uint32_t i = 0; i < n; ++i) {
            const UnitDev &U = hu[W.job_unit[i]];
            h->m_pos[jobs[i].stream] = U.final_position;
            h->m_avail[jobs[i].stream] = U.final_available;
        }
    } else {
        for (uint32_t i = 0; i < n; ++i) {
            h->m_ok[jobs[i].stream] = 0;
            h->m_pending[jobs[i].stream] = h->submits;
        }
    }
    // Tensor kernel generation 2 (fir_tc2.cu): additionally needs the plan's totals on the host
    // (its TMA stores clip to what the plan produces) and equally strided, 16-byte aligned outputs.
    bool use_tc2 = false;
    CUtensorMap tc2_out_map;
    uint64_t tc2_tiles = 0;
    if (use_tc && host_plan && hu[0].total_out > 0 && rsb::tc2_supported(ch, h->taps, h->ratio) &&
        rsb::tc2_g_stages(ch, h->taps, h->ratio) >= 2 && !getenv("RSB_TC_V1")) {
        const uintptr_t obase = reinterpret_cast<uintptr_t>(hj[0].out);
        uint64_t ostride = n > 1 ? (uint64_t)(reinterpret_cast<uintptr_t>(hj[1].out) - obase) : 0;
        bool ok = n == 1 || reinterpret_cast<uintptr_t>(hj[1].out) > obase;
        uint64_t min_cap = ~0ull;
        for (uint32_t i = 0; i < n; ++i) {
            if (ok && i > 0) ok = reinterpret_cast<uintptr_t>(hj[i].out) == obase + (uint64_t)i * ostride;
            min_cap = std::min<uint64_t>(min_cap, hj[i].out_capacity);
        }
        const uint64_t valid = std::min<uint64_t>(hu[0].total_out, min_cap);
        if (n == 1) ostride = (valid * ch * 4 + 15) & ~15ull;
        if (ok && valid > 0)
            use_tc2 = rsb::tc2_make_output_tensor_map(&tc2_out_map, hj[0].out,
                                                      getenv("RSB_DEBUG_FAKE_OUT_STRIDE") ? (uint64_t)atoll(getenv("RSB_DEBUG_FAKE_OUT_STRIDE")) : ostride,
                                                      valid, n, ch);
        tc2_tiles = (hu[0].total_out + rsb::kTc2TileOut - 1) / rsb::kTc2TileOut;
    }
    bool tc2_hit = false;
    if (use_tc2) {
        tc2_hit = plan_key == h->tc2_key && !getenv("RSB_TC_NO_PLAN_CACHE");
        if (!tc2_hit) {
            h->tc2_key.valid = false;      // the buffers are about to be rebuilt (or freed by a failed reserve)
            RSB_CUDA(h->d_tct2.reserve(sizeof(rsb::Tc2Tile) * tc2_tiles));
            RSB_CUDA(h->d_gmat2.reserve(rsb::tc2_gmat_bytes_per_tile(h->taps, h->ratio) * tc2_tiles));
        } else {
            h->tc2_cache_hits += 1;
        }
    } else if (use_tc && !rsb::tc_supported(ch, h->taps, h->ratio)) {
        use_tc = false;      // generation 1 cannot take this ratio: FFMA2 / exact kernel
        if (use_fast) RSB_CUDA(W.d_gtiles.reserve(sizeof(float) * rsb::kTileOut * gs * tile_total));
    } else if (use_tc) {
        RSB_CUDA(W.d_gtiles.reserve(sizeof(float) * rsb::tc_gmat_floats_per_tile(h->taps, h->ratio) *
                                    tile_total));
        RSB_CUDA(W.d_tct.reserve(sizeof(rsb::TcTile) * tile_total));
    }
    RSB_CUDA(cudaMemcpyAsync(W.d_units.p, hu, sizeof(UnitDev) * n_units, cudaMemcpyHostToDevice, sp));
    RSB_CUDA(cudaMemcpyAsync(W.d_jobs.p, hj, sizeof(JobDev) * n, cudaMemcpyHostToDevice, sp));
    RSB_CUDA(cudaMemsetAsync(W.d_counter.p, 0, sizeof(uint32_t) * 4, sp));
    // host buffers that are equally sized and equally strided move with ONE 2-D copy
    bool h2d_uniform = host_mem && !pcm && n > 1, d2h_uniform = host_mem && n > 1 && n_units == 1;
    ptrdiff_t in_pitch = 0, out_pitch = 0;
    if (host_mem && n > 1) {
        in_pitch = reinterpret_cast<const char *>(jobs[1].in) - reinterpret_cast<const char *>(jobs[0].in);
        out_pitch = reinterpret_cast<char *>(jobs[1].out) - reinterpret_cast<char *>(jobs[0].out);
        for (uint32_t i = 1; i < n; ++i) {
            if (in_vals[i] != in_vals[0] || in_off[i] - in_off[i - 1] != in_off[1] - in_off[0] ||
                reinterpret_cast<const char *>(jobs[i].in) - reinterpret_cast<const char *>(jobs[0].in) !=
                    (ptrdiff_t)i * in_pitch)
                h2d_uniform = false;
            if (out_off[i] - out_off[i - 1] != out_off[1] - out_off[0] ||
                reinterpret_cast<char *>(jobs[i].out) - reinterpret_cast<char *>(jobs[0].out) !=
                    (ptrdiff_t)i * out_pitch)
                d2h_uniform = false;
        }
        // separately allocated host arrays can be terabytes apart: one 2-D copy only within the
        // pitch the runtime accepts, else one copy per stream
        if (in_pitch < (ptrdiff_t)(in_vals[0] * sizeof(float)) || (size_t)in_pitch > h->mem_pitch)
            h2d_uniform = false;
        if (out_pitch <= 0 || (size_t)out_pitch > h->mem_pitch) d2h_uniform = false;
    }
    bool pcm_fused = false;   // the tensor kernel reads the raw s16 frames itself
    uint32_t pcm_raw_mode = 0;   // TcParams::raw16
    if (pcm) {
        // ---- format step: raw samples -> interleaved f32 frames in d_stage_in ----
        uint64_t n_chunks = 0, chunks_per_job = 0;
        RSB_CUDA(W.h_pcm_jobs.reserve(sizeof(rsb::PcmJob) * n + sizeof(uint64_t) * (n + 1)));
        RSB_CUDA(W.d_pcm_jobs.reserve(sizeof(rsb::PcmJob) * n + sizeof(uint64_t) * (n + 1)));
        rsb::PcmJob *pj = W.h_pcm_jobs.as<rsb::PcmJob>();
        uint64_t *cf = reinterpret_cast<uint64_t *>(pj + n);
        // equally sized host buffers at one pitch move with ONE 2-D copy
        bool raw_uniform = host_mem && n > 1 && in_vals[0] != 0;
        ptrdiff_t raw_pitch = 0;
        if (raw_uniform) {
            raw_pitch = reinterpret_cast<const char *>(jobs[1].in) - reinterpret_cast<const char *>(jobs[0].in);
            for (uint32_t i = 1; raw_uniform && i < n; ++i)
                raw_uniform = in_vals[i] == in_vals[0] &&
                              reinterpret_cast<const char *>(jobs[i].in) -
                                      reinterpret_cast<const char *>(jobs[0].in) == (ptrdiff_t)i * raw_pitch;
            if (raw_pitch < (ptrdiff_t)(in_vals[0] / pcm->dup * pcm->bps) || (size_t)raw_pitch > h->mem_pitch)
                raw_uniform = false;
        }
        if (raw_uniform)
            RSB_CUDA(cudaMemcpy2DAsync(h->d_pcm_raw.p, raw_off[1] - raw_off[0], jobs[0].in,
                                       (size_t)raw_pitch, in_vals[0] / pcm->dup * pcm->bps, n,
                                       cudaMemcpyHostToDevice, s));
        for (uint32_t i = 0; i < n; ++i) {
            const size_t raw_bytes = in_vals[i] / pcm->dup * pcm->bps;
            const void *src = jobs[i].in;
            if (host_mem) {
                void *d = static_cast<char *>(h->d_pcm_raw.p) + raw_off[i];
                if (raw_bytes && !raw_uniform)
                    RSB_CUDA(cudaMemcpyAsync(d, jobs[i].in, raw_bytes, cudaMemcpyHostToDevice, s));
                src = d;
            }
            pj[i].src = src;
            pj[i].dst = h->d_stage_in.as<float>() + in_off[i];
            pj[i].n_out = in_vals[i];
            pj[i].src_aligned = (reinterpret_cast<uintptr_t>(src) & 15u) == 0 ? 1u : 0u;
            pj[i].pad_ = 0;
        }
        // Stereo s16 batches that run on the tensor kernel skip the format pass: the kernel's TMA
        // producer streams the RAW frames (tensor map over the raw rows) and its splitter
        // converts them.  Only the last 4096 frames of every stream are still converted here,
        // for the history update (update_state_kernel reads them as f32).
        pcm_fused = false;
        const uint32_t src_ch = ch / pcm->dup;   // 1 (mono source) or ch
        const uint32_t sb = pcm->bps;            // bytes per raw sample
        if (use_tc && host_plan && (pcm->format == RSB_PCM_S16 || pcm->format == RSB_PCM_S24) && ch <= 2 &&
            !getenv("RSB_PCM_UNFUSED")) {
            const uintptr_t base = reinterpret_cast<uintptr_t>(pj[0].src);
            const uint64_t stride = n > 1 ? (uint64_t)(reinterpret_cast<uintptr_t>(pj[1].src) - base)
                                          : (((uint64_t)in_vals[0] / pcm->dup * sb + 15) & ~15ull);
            bool ok = n == 1 || reinterpret_cast<uintptr_t>(pj[1].src) > base;
            for (uint32_t i = 1; ok && i < n; ++i)
                ok = reinterpret_cast<uintptr_t>(pj[i].src) == base + (uint64_t)i * stride;
            // jobs are grouped by unit in hj; one unit => hj order == job order
            CUtensorMap raw_map;
            if (ok && rsb::tc_make_raw16_tensor_map(&raw_map, pj[0].src, stride,
                                                    unit_keys[0].total_frames, n, src_ch, ch, sb)) {
                tc_tmap = raw_map;
                pcm_fused = true;
                pcm_raw_mode = src_ch == ch ? 1u : 2u;
            }
        }
        if (pcm_fused) {
            const uint64_t copied = hu[0].total_copied;
            const uint64_t start = (copied > rsb::kHistFrames ? copied - rsb::kHistFrames : 0) & ~7ull;
            for (uint32_t i = 0; i < n; ++i) {
                pj[i].src = static_cast<const char *>(pj[i].src) + start * sb * src_ch;   // raw frames
                pj[i].dst += start * ch;
                pj[i].n_out -= start * ch;
                pj[i].src_aligned = (reinterpret_cast<uintptr_t>(pj[i].src) & 15u) == 0 ? 1u : 0u;
            }
        }
        chunks_per_job = rsb::pcm_chunks(pj[0].n_out);
        for (uint32_t i = 0; i < n; ++i) {
            cf[i] = n_chunks;
            n_chunks += rsb::pcm_chunks(pj[i].n_out);
            if (rsb::pcm_chunks(pj[i].n_out) != chunks_per_job) chunks_per_job = 0;
        }
        cf[n] = n_chunks;
        RSB_CUDA(cudaMemcpyAsync(W.d_pcm_jobs.p, pj, sizeof(rsb::PcmJob) * n + sizeof(uint64_t) * (n + 1),
                                 cudaMemcpyHostToDevice, s));
        RSB_CUDA(cudaEventRecord(h->ev_pcm[0], s));
        rsb::launch_pcm_ingest(W.d_pcm_jobs.as<rsb::PcmJob>(),
                               reinterpret_cast<const uint64_t *>(W.d_pcm_jobs.as<rsb::PcmJob>() + n), n,
                               n_chunks, chunks_per_job, pcm->format, pcm->dup, h->sm_count, s);
        RSB_CUDA(cudaEventRecord(h->ev_pcm[1], s));
        h->pcm_timed = n_chunks != 0;
        h->pcm_fused_last = pcm_fused;
        if (n_chunks) h->launches += 1;
    } else if (host_mem) {
        if (h2d_uniform && in_vals[0]) {
            RSB_CUDA(cudaMemcpy2DAsync(h->d_stage_in.p, (in_off[1] - in_off[0]) * sizeof(float),
                                       jobs[0].in, (size_t)in_pitch, in_vals[0] * sizeof(float), n,
                                       cudaMemcpyHostToDevice, s));
        } else {
            for (uint32_t i = 0; i < n; ++i)
                if (in_vals[i])
                    RSB_CUDA(cudaMemcpyAsync(h->d_stage_in.as<float>() + in_off[i], jobs[i].in,
                                             in_vals[i] * sizeof(float), cudaMemcpyHostToDevice, s));
        }
    }

    // ---- launch ----
    if (host_plan) {
        // the plan already exists: upload the segments in use and the tile total
        for (uint32_t u = 0; u < n_units; ++u)
            if (hu[u].n_segs)
                RSB_CUDA(cudaMemcpyAsync(W.d_segs.as<rsb::PlanSeg>() + hu[u].seg_off,
                                         W.h_segs.as<rsb::PlanSeg>() + hu[u].seg_off,
                                         sizeof(rsb::PlanSeg) * hu[u].n_segs, cudaMemcpyHostToDevice, sp));
        uint32_t *hcnt = W.h_counter.as<uint32_t>();
        hcnt[0] = (uint32_t)host_tiles;
        RSB_CUDA(cudaMemcpyAsync(W.d_counter.p, hcnt, sizeof(uint32_t), cudaMemcpyHostToDevice, sp));
    } else {
        rsb::launch_plan(W.d_units.as<UnitDev>(), n_units, h->st, h->ratio, h->taps,
                         W.d_segs.as<rsb::PlanSeg>(), W.d_calls.as<rsb::CallCounts>(), tile_out,
                         W.d_counter.as<uint32_t>(), sp);
    }
    // the streams' scalar state (position, buffered frames) moves on the plan stream, so the
    // next submit can be planned while this one is still convolving
    rsb::launch_state_scalars(W.d_units.as<UnitDev>(), W.d_jobs.as<JobDev>(), n, h->st, sp);
    if (!host_plan) {
        RSB_CUDA(cudaMemcpyAsync(W.h_units_back.p, W.d_units.p, sizeof(UnitDev) * n_units,
                                 cudaMemcpyDeviceToHost, sp));
        if (rec_calls && call_total)
            RSB_CUDA(cudaMemcpyAsync(W.h_calls_back.p, W.d_calls.p,
                                     sizeof(rsb::CallCounts) * call_total, cudaMemcpyDeviceToHost, sp));
    }
    RSB_CUDA(cudaEventRecord(W.ev_plan, sp));
    RSB_CUDA(cudaStreamWaitEvent(s, W.ev_plan, 0));
    // Tile records, per-frame plan entries and the banded filter tiles: wide, short kernels on
    // the main stream in front of the convolution.  (Running them on the plan stream underneath
    // the previous submit's convolution was measured: they take 1.1 ms alone but slow the
    // persistent convolution kernel by 1.5 ms.)
    // (a cached tensor plan needs neither tile records nor per-frame entries, unless the caller
    // wants to read the plan back)
    const bool skip_tiles = tc2_hit && !(flags & RSB_FLAG_KEEP_PLAN);
    if (!skip_tiles)
        rsb::launch_tiles(W.d_units.as<UnitDev>(), n_units, W.d_segs.as<rsb::PlanSeg>(),
                          W.d_tiles.as<rsb::TileRec>(), W.d_entries.as<rsb::PlanEntry>(),
                          (uint32_t)max_tiles_unit, h->taps, h->d_coeffs,
                          use_fast && !use_tc ? W.d_gtiles.as<float>() : nullptr, gs, s);
    if (use_tc2 && !tc2_hit) {
        rsb::launch_tc2_tiles(W.d_units.as<UnitDev>(), W.d_entries.as<rsb::PlanEntry>(),
                              h->d_tct2.as<rsb::Tc2Tile>(), h->taps, h->ratio, (uint32_t)tc2_tiles, s);
        rsb::launch_tc2_gmat(W.d_units.as<UnitDev>(), W.d_entries.as<rsb::PlanEntry>(), h->d_coeffs,
                             h->d_tct2.as<rsb::Tc2Tile>(), nullptr, nullptr, h->d_gmat2.as<uint8_t>(), h->taps,
                             h->ratio, (uint32_t)tc2_tiles,
                             getenv("RSB_TC_COMP") ? atof(getenv("RSB_TC_COMP")) : rsb::kTc2TruncationComp, s);
        h->tc2_key = plan_key;
    } else if (use_tc2) {
        // cached
    } else if (use_tc)
        rsb::launch_tc_gmat(W.d_units.as<UnitDev>(), W.d_tiles.as<rsb::TileRec>(),
                            W.d_entries.as<rsb::PlanEntry>(), h->d_coeffs, W.d_gtiles.as<float>(),
                            W.d_tct.as<rsb::TcTile>(), h->taps, h->ratio, (uint32_t)tile_total, s);
    rsb::ConvParams P;
    P.units = W.d_units.as<UnitDev>();
    P.jobs = W.d_jobs.as<JobDev>();
    P.segs = W.d_segs.as<rsb::PlanSeg>();
    P.tiles = W.d_tiles.as<rsb::TileRec>();
    P.entries = W.d_entries.as<rsb::PlanEntry>();
    P.gtiles = W.d_gtiles.as<float>();
    P.tile_total = W.d_counter.as<uint32_t>();
    P.work_counter = W.d_counter.as<uint32_t>() + 1;
    P.coeffs = h->d_coeffs;
    P.st = h->st;
    P.channels = ch;
    P.taps = h->taps;
    P.groups = max_groups;
    P.streams_per_group = spg;
    P.group_block = std::min<uint32_t>(max_groups, 2u);
    // single plan unit with equally strided member inputs: one TMA tensor map for the batch
    CUtensorMap tmap;
    P.tmap_valid = 0;
    if (use_fast && !use_tc && in_uniform &&
        rsb::fast_make_input_tensor_map(&tmap, hj[0].in, in_stride, unit_keys[0].total_frames, n, ch,
                                        h->taps, h->ratio))
        P.tmap_valid = 1;
    const uint32_t max_items = (uint32_t)(tile_total * max_groups);
    const int ring = (int)(h->conv_batches % rsb_fir::kConvRing);
    RSB_CUDA(cudaEventRecord(h->ev_conv[ring][0], s));
    if (use_tc2) {
        rsb::Tc2Params T;
        T.units = P.units;
        T.jobs = P.jobs;
        T.tct = h->d_tct2.as<rsb::Tc2Tile>();
        T.gmat = h->d_gmat2.as<uint8_t>();
        T.work_counter = P.work_counter;
        T.channels = ch;
        const uint32_t mpg = rsb::tc2_rows_per_group() / ch;
        T.groups = (n + mpg - 1) / mpg;
        // A run of consecutive tiles is one work item: long enough to amortise filling the input
        // ring (~4 tiles; measured best at 160-190 tiles), and cut so that the items divide evenly
        // among the CTAs: k items per CTA (at least 8: with fewer the rounding of k * SMs / groups to
        // whole runs leaves CTAs idle in small batches), n_runs = floor(k * SMs / groups) runs.
        {
            const uint64_t total = tc2_tiles * T.groups, sms = (uint64_t)h->sm_count;
            const uint64_t k = std::max<uint64_t>(8, (total + sms * 85) / (sms * 170));
            const uint64_t n_runs = std::max<uint64_t>(1, k * sms / T.groups);
            const uint64_t rt = (tc2_tiles + n_runs - 1) / n_runs;
            T.run_tiles = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(rt, 16), 256);
        }
        if (getenv("RSB_TC_RUN_TILES")) T.run_tiles = (uint32_t)atoi(getenv("RSB_TC_RUN_TILES"));
        T.kt_max = rsb::tc2_kt_extent(h->taps, h->ratio);
        T.issuers = rsb::tc2_issuers(h->taps, h->ratio);
        T.variant = getenv("RSB_TC_VARIANT") ? (uint32_t)atoi(getenv("RSB_TC_VARIANT")) : 0u;
        T.g_stages = rsb::tc2_g_stages(ch, h->taps, h->ratio, T.variant);
        T.prefetch_chunks = getenv("RSB_TC_PREFETCH") ? (uint32_t)atoi(getenv("RSB_TC_PREFETCH")) : 0u;
        T.ablate = getenv("RSB_TC_ABLATE") ? (uint32_t)atoi(getenv("RSB_TC_ABLATE")) : 0u;
        T.out_scale = rsb::tc2_out_scale();
        // both epilogue teams drain every tile, one column half each: the accumulator is back with
        // the issuers one tcgen05.ld after the tile completes
        T.epi_split = getenv("RSB_TC_EPI_SPLIT") ? (uint32_t)atoi(getenv("RSB_TC_EPI_SPLIT")) : 1u;
        T.hint_crit = getenv("RSB_TC_HINT_CRIT") ? (uint32_t)atoi(getenv("RSB_TC_HINT_CRIT")) : 0x989680u;
        T.hint_other = getenv("RSB_TC_HINT_OTHER") ? (uint32_t)atoi(getenv("RSB_TC_HINT_OTHER")) : 0x989680u;
        if (getenv("RSB_TC_GSTAGES")) T.g_stages = std::min<uint32_t>(T.g_stages, (uint32_t)atoi(getenv("RSB_TC_GSTAGES")));
        T.raw16 = pcm_fused ? pcm_raw_mode : 0u;
        T.raw_bytes = pcm && pcm_fused ? pcm->bps : 2u;
        if (getenv("RSB_TC_ISSUERS")) T.issuers = atoi(getenv("RSB_TC_ISSUERS")) == 1 ? 1u : T.issuers;
        rsb::launch_conv_tc2(T, tc_tmap, tc2_out_map, h->sm_count, false, s);
    } else if (use_tc) {
        rsb::TcParams T;
        T.units = P.units;
        T.jobs = P.jobs;
        T.tct = W.d_tct.as<rsb::TcTile>();
        T.gmat = W.d_gtiles.as<float>();
        T.work_counter = P.work_counter;
        T.channels = ch;
        const uint32_t mpg = rsb::tc_rows_per_group() / ch;
        T.groups = (n + mpg - 1) / mpg;
        // a run of consecutive tiles is one work item: long enough to amortise filling the input
        // ring (~6 tiles), short enough that every CTA gets several items and the tail is short
        // (measured on the headline batch: 192 tiles 15.8 ms, 512 tiles 16.0 ms, 1024 tiles 16.1 ms)
        const uint64_t want = (tile_total * T.groups) / ((uint64_t)h->sm_count * 8u) + 1;
        T.run_tiles = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 32), 192);
        if (getenv("RSB_TC_RUN_TILES")) T.run_tiles = (uint32_t)atoi(getenv("RSB_TC_RUN_TILES"));
        T.kt_max = rsb::tc_kt_extent(h->taps, h->ratio);
        T.issuers = rsb::tc_issuers(h->taps, h->ratio);
        T.raw16 = pcm_fused ? pcm_raw_mode : 0u;
        T.raw_bytes = pcm && pcm_fused ? pcm->bps : 2u;
        if (getenv("RSB_TC_ISSUERS")) T.issuers = atoi(getenv("RSB_TC_ISSUERS")) == 1 ? 1u : T.issuers;
        rsb::launch_conv_tc(T, tc_tmap, h->sm_count, !host_plan, s);
    } else if (use_fast) {
        rsb::launch_conv_fast(P, P.tmap_valid ? &tmap : nullptr, h->ratio, max_items, h->sm_count, s);
    } else {
        rsb::launch_conv_exact(P, max_items, h->sm_count, s);
    }
    h->last_kernel = use_tc ? RSB_KERNEL_TENSOR : use_fast ? RSB_KERNEL_FAST : RSB_KERNEL_EXACT;
    RSB_CUDA(cudaEventRecord(h->ev_conv[ring][1], s));
    h->conv_batches += 1;
    rsb::launch_update(W.d_units.as<UnitDev>(), W.d_jobs.as<JobDev>(), n, h->st, ch, s);
    h->launches += (use_tc2 ? (tc2_hit ? (skip_tiles ? 4 : 5) : 7) : use_tc ? 6 : 5) - (host_plan ? 1 : 0);
    RSB_CUDA(cudaGetLastError());
    RSB_CUDA(cudaEventRecord(W.ev_done, s));
    h->submits += 1;

    // cohorts: every unit becomes a fresh cohort (same state + same calls => same new state)
    {
        std::vector<uint64_t> new_id(n_units);
        for (uint32_t u = 0; u < n_units; ++u) new_id[u] = h->next_cohort++;
        for (uint32_t i = 0; i < n; ++i) {
            h->cohort[jobs[i].stream] = new_id[W.job_unit[i]];
            h->hist_sel[jobs[i].stream] ^= 1u;   // the update kernel wrote the other buffer
        }
    }
    W.n_units = n_units;
    W.has_calls = rec_calls;
    W.has_plan = (flags & RSB_FLAG_KEEP_PLAN) != 0;

    W.pending.active = true;
    W.pending.n = n;
    W.pending.n_units = n_units;
    W.pending.consumed = consumed;
    W.pending.produced = produced;
    W.pending.n_calls = n_calls_out;
    W.pending.check_capacity = !single;
    if ((flags & RSB_FLAG_ASYNC) && !host_mem) return RSB_OK;

    rc = finalize_pending(h, W);
    // a synchronous call returns with its outputs complete (the handle's streams are
    // non-blocking: nothing else would order a caller's later copy behind the convolution)
    if (!host_mem) RSB_CUDA(cudaEventSynchronize(W.ev_done));
    if (host_mem) {
        const UnitDev *ub = W.h_units_back.as<UnitDev>();
        uint64_t min_cap = ~0ull;
        for (uint32_t i = 0; i < n; ++i) min_cap = std::min<uint64_t>(min_cap, jobs[i].out_capacity);
        if (d2h_uniform && ub[0].total_out <= min_cap) {
            // one plan unit => every stream produced the same number of frames
            const size_t width = (size_t)ub[0].total_out * ch * sizeof(float);
            if (width)
                RSB_CUDA(cudaMemcpy2DAsync(jobs[0].out, (size_t)out_pitch, h->d_stage_out.p,
                                           (out_off[1] - out_off[0]) * sizeof(float), width, n,
                                           cudaMemcpyDeviceToHost, s));
        } else {
            for (uint32_t i = 0; i < n; ++i) {
                const UnitDev &U = ub[W.job_unit[i]];
                const uint64_t frames = std::min<uint64_t>(U.total_out, jobs[i].out_capacity);
                if (frames)
                    RSB_CUDA(cudaMemcpyAsync(jobs[i].out, h->d_stage_out.as<float>() + out_off[i],
                                             (size_t)frames * ch * sizeof(float),
                                             cudaMemcpyDeviceToHost, s));
            }
        }
        RSB_CUDA(cudaStreamSynchronize(s));
    }
    return rc;
}

// Delivers the results of a fused submit: counts to the caller's arrays, scalar state to the host
// mirror (unless a later submit has taken the stream on).
int finalize_fused(rsb_fir *h, rsb_fir::FusedSlot &F) {
    if (!F.active) return RSB_OK;
    F.active = false;
    RSB_CUDA(cudaEventSynchronize(F.ev_done));
    const rsb::SubmitResult *res = F.h_res.as<rsb::SubmitResult>();
    const uint32_t ch = h->channels;
    int rc = RSB_OK;
    for (uint32_t i = 0; i < F.n; ++i) {
        const uint32_t s = F.streams[i];
        if (res[i].status != 0) rc = fail(RSB_ERR_PLAN_OVERFLOW, "plan workspace bound exceeded (fused submit)");
        if (F.consumed) F.consumed[i] = (size_t)res[i].copied * ch;
        if (F.produced) F.produced[i] = (size_t)res[i].produced * ch;
        if (h->m_fused_seq[s] == F.seq && !h->m_ok[s]) {
            h->m_pos[s] = res[i].position;
            h->m_avail[s] = res[i].available;
            h->m_ok[s] = 1;
        }
    }
    return rc;
}

int finalize_fused_all(rsb_fir *h) {
    int rc = RSB_OK;
    for (uint32_t k = 0; k < rsb_fir::kFusedSlots; ++k) {           // oldest first
        const int r = finalize_fused(h, h->fused[(h->fused_count + k) % rsb_fir::kFusedSlots]);
        if (rc == RSB_OK) rc = r;
    }
    return rc;
}

// One launch for the whole submit (device memspace): see fir_submit.cu.
// Returns RSB_OK / an error, or kNotTaken when an AUTO submit is large enough for the tile kernels
// (nothing has been changed then).
constexpr int kNotTaken = -1000;
int run_submit_fused(rsb_fir *h, uint32_t n, const uint32_t *streams, const float *const *in,
                     const size_t *in_lens, float *const *out, const size_t *out_lens, size_t *consumed,
                     size_t *produced, uint32_t flags) {
    RSB_CUDA(cudaSetDevice(h->device));
    rsb_fir::FusedSlot &F = h->fused[h->fused_count % rsb_fir::kFusedSlots];
    int rc = finalize_fused(h, F);
    if (rc != RSB_OK) return rc;
    // GPU-planned batches of the general path may still owe these streams a mirror refresh; their
    // kernels run on the same streams in order, nothing to wait for here
    const uint32_t ch = h->channels;
    if (!F.ev_done) RSB_CUDA(cudaEventCreateWithFlags(&F.ev_done, cudaEventDisableTiming));
    RSB_CUDA(F.h_jobs.reserve(sizeof(rsb::SubmitJob) * n));
    RSB_CUDA(F.h_res.reserve(sizeof(rsb::SubmitResult) * n));
    RSB_CUDA(F.d_jobs.reserve(sizeof(rsb::SubmitJob) * n));
    RSB_CUDA(F.d_res.reserve(sizeof(rsb::SubmitResult) * n));
    RSB_CUDA(F.d_segs.reserve(sizeof(rsb::PlanSeg) * rsb::kSubmitSegs * n));
    rsb::SubmitJob *hj = F.h_jobs.as<rsb::SubmitJob>();
    F.streams.resize(n);
    uint32_t *fs = F.streams.data();
    const size_t hist_stride = (size_t)rsb::kHistFrames * ch;
    // ---- pass 1: validation (error precedence of resampler_fir.rs:514-519 per job) and the job
    // table; touches scratch only.  The host cost of a submit is this loop: no 64-bit divisions
    // (channel counts are powers of two in practice), one pass over the caller's arrays. ----
    const bool pow2 = (ch & (ch - 1)) == 0;
    const uint32_t sh = pow2 ? (uint32_t)__builtin_ctz(ch) : 0u;
    const size_t mask = pow2 ? (size_t)ch - 1 : 0;
    h->epoch += 1;
    if (h->epoch == 0) { std::fill(h->seen_epoch.begin(), h->seen_epoch.end(), 0u); h->epoch = 1; }
    const uint32_t epoch = h->epoch, n_streams = h->n_streams;
    uint32_t *seen = h->seen_epoch.data();
    const uint8_t *hsel = h->hist_sel.data();
    const uint64_t *coh = h->cohort.data();
    float *const hist0 = h->st.hist[0], *const hist1 = h->st.hist[1];
    bool same = true;
    uint32_t max_in = 0;
    uint64_t in_total = 0;
    const uint64_t c0 = coh[streams ? (streams[0] < n_streams ? streams[0] : 0u) : 0u];
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t s = streams ? streams[i] : i;
        const size_t il = in_lens[i], ol = out_lens[i];
        if (pow2 ? (il & mask) != 0 : il % ch != 0)
            return fail(RSB_ERR_INVALID_INPUT_BUFFER_SIZE, "input length of job " + std::to_string(i) +
                                                               " is not a multiple of channels");
        if (pow2 ? (ol & mask) != 0 : ol % ch != 0)
            return fail(RSB_ERR_INVALID_OUTPUT_BUFFER_SIZE, "output length of job " + std::to_string(i) +
                                                                " is not a multiple of channels");
        if (s >= n_streams) return fail(RSB_ERR_INVALID_ARGUMENT, "bad stream index");
        if (streams) {      // the identity mapping cannot list a stream twice
            if (seen[s] == epoch) return fail(RSB_ERR_INVALID_ARGUMENT, "stream listed twice in one batch");
            seen[s] = epoch;
        }
        if (il && !in[i]) return fail(RSB_ERR_INVALID_ARGUMENT, "null input pointer");
        if (ol && !out[i]) return fail(RSB_ERR_INVALID_ARGUMENT, "null output pointer");
        const uint64_t fin = pow2 ? il >> sh : il / ch, fout = pow2 ? ol >> sh : ol / ch;
        rsb::SubmitJob J;
        J.in = in[i];
        J.out = out[i];
        const uint32_t sel = hsel[s];
        J.hist = (sel ? hist1 : hist0) + hist_stride * s;
        J.hist_next = (sel ? hist0 : hist1) + hist_stride * s;
        J.in_frames = (uint32_t)std::min<uint64_t>(fin, rsb::kInputCapacity);   // :526-528
        J.cap_frames = (uint32_t)std::min<uint64_t>(fout, 0xffffffffull);
        J.stream = s;
        J.pad = 0;
        hj[i] = J;
        fs[i] = s;
        max_in = std::max(max_in, J.in_frames);
        in_total += J.in_frames;
        same = same && coh[s] == c0 && J.in_frames == hj[0].in_frames && J.cap_frames == hj[0].cap_frames;
    }
    // AUTO: the tile kernels win once a submit carries tens of millions of samples ...
    if (h->kernel_mode != RSB_KERNEL_EXACT &&
        ((double)in_total + (double)n * (rsb::kInputCapacity / 8)) / h->ratio * ch > (double)(32ull << 20))
        return kNotTaken;
    // ... and much earlier when every stream makes the SAME call from the same state: then one plan
    // and one set of filter tiles serve all streams (the tensor kernel), while this path re-reads two
    // coefficient rows per output.  Measured crossover ~1e8 sample-taps (1024 stereo x 512 frames at
    // 128 taps: 70 us against 113 us here; 2048 frames: 87 against 446; 4096 mono x 160 frames at 32
    // taps stays here: 79 against 150).
    if (h->kernel_mode == RSB_KERNEL_AUTO && same && (uint64_t)n * ch >= 64 &&
        (double)in_total / h->ratio * ch * (double)h->taps > 1.0e8)
        return kNotTaken;
    // ---- pass 2: bookkeeping ----
    const uint64_t seq = ++h->fused_count;         // slot index of THIS submit was taken above
    // all jobs of one cohort with one call signature keep sharing a cohort afterwards
    const uint64_t shared = h->next_cohort;
    if (same) h->next_cohort += 1;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t s = fs[i];
        h->cohort[s] = same ? shared : h->next_cohort++;
        h->hist_sel[s] ^= 1u;
        h->m_ok[s] = 0;
        h->m_fused_seq[s] = seq;
        h->m_pending[s] = ~0ull;      // no general-path read-back may refresh this stream any more
    }
    cudaStream_t s = h->stream, sp = h->plan_stream;
    if (rsb::submit_two_stage(h->taps, ch, max_in)) {
        // Stage 1 on the plan stream (job table up, planner kernel, result records down): it touches
        // the scalar stream state only, so it runs underneath the previous submit's convolution.
        // Stage 2 (samples, history) on the main stream behind it.
        if (!F.ev_plan) RSB_CUDA(cudaEventCreateWithFlags(&F.ev_plan, cudaEventDisableTiming));
        RSB_CUDA(cudaMemcpyAsync(F.d_jobs.p, hj, sizeof(rsb::SubmitJob) * n, cudaMemcpyHostToDevice, sp));
        rsb::launch_submit_plan(F.d_jobs.as<rsb::SubmitJob>(), F.d_res.as<rsb::SubmitResult>(), n, h->st, h->ratio,
                                h->taps, F.d_segs.as<rsb::PlanSeg>(), sp);
        RSB_CUDA(cudaMemcpyAsync(F.h_res.p, F.d_res.p, sizeof(rsb::SubmitResult) * n, cudaMemcpyDeviceToHost, sp));
        RSB_CUDA(cudaEventRecord(F.ev_plan, sp));
        RSB_CUDA(cudaStreamWaitEvent(s, F.ev_plan, 0));
        rsb::launch_submit_conv(F.d_jobs.as<rsb::SubmitJob>(), F.d_res.as<rsb::SubmitResult>(), n, h->d_coeffs, h->taps,
                                ch, max_in, F.d_segs.as<rsb::PlanSeg>(), h->in_hz, h->out_hz, s);
        RSB_CUDA(cudaGetLastError());
        RSB_CUDA(cudaEventRecord(F.ev_done, s));
    } else {
        // single-launch fallback: the kernel touches the scalar state on the main stream: order the two
        RSB_CUDA(cudaEventRecord(h->ev_sync, sp));
        RSB_CUDA(cudaStreamWaitEvent(s, h->ev_sync, 0));
        RSB_CUDA(cudaMemcpyAsync(F.d_jobs.p, hj, sizeof(rsb::SubmitJob) * n, cudaMemcpyHostToDevice, s));
        rsb::launch_submit_fused(F.d_jobs.as<rsb::SubmitJob>(), F.d_res.as<rsb::SubmitResult>(), n, h->st,
                                 h->d_coeffs, h->ratio, h->taps, ch, max_in, F.d_segs.as<rsb::PlanSeg>(), s);
        RSB_CUDA(cudaGetLastError());
        RSB_CUDA(cudaMemcpyAsync(F.h_res.p, F.d_res.p, sizeof(rsb::SubmitResult) * n, cudaMemcpyDeviceToHost, s));
        RSB_CUDA(cudaEventRecord(F.ev_done, s));
        RSB_CUDA(cudaStreamWaitEvent(sp, F.ev_done, 0));
    }
    h->launches += 2;      // plan + convolution (one launch on the half-warp fallback)
    h->last_kernel = RSB_KERNEL_EXACT;
    F.active = true;
    F.n = n;
    F.seq = seq;
    F.consumed = consumed;
    F.produced = produced;
    if (flags & RSB_FLAG_ASYNC) return RSB_OK;
    return finalize_fused(h, F);
}

// Host-memspace batches of one plan unit, long enough to be worth slicing: the batch is cut at call
// boundaries into time slices, and slice k+1's host->device copy, slice k's kernels and slice
// k-1's device->host copy run concurrently (three streams, two device buffers each way).  A dry
// run of the (sample-independent) plan on a host core fixes the slice boundaries at cumulative
// `copied` offsets, so every slice offers exactly the frames the unsliced loop would have consumed
// by then: the calls, their counts and the samples are those of the unsliced batch
// (resampler_fir.rs:509-621 called in a loop; state carries from slice to slice in the handle).
// Returns false when the batch does not qualify (the caller then takes the unsliced path).
bool process_host_pipelined(rsb_fir *h, const std::vector<JobHost> &jobs, size_t *consumed, size_t *produced,
                            uint32_t *n_calls_out, bool async, int *rc_out) {
    const uint32_t n = (uint32_t)jobs.size();
    const uint32_t ch = h->channels;
    const uint64_t T = jobs[0].total_frames;
    const uint32_t cf = jobs[0].call_frames;
    if (cf == 0 || T < 16ull * cf || (uint64_t)n * T * ch * sizeof(float) < (64ull << 20)) return false;
    for (uint32_t i = 1; i < n; ++i)
        if (jobs[i].total_frames != T || h->cohort[jobs[i].stream] != h->cohort[jobs[0].stream]) return false;
    if (cudaSetDevice(h->device) != cudaSuccess) return false;
    if (finalize_fused_all(h) != RSB_OK || finalize_all(h) != RSB_OK) return false;
    const uint32_t rep = jobs[0].stream;
    if (!h->m_ok[rep]) return false;

    // ---- dry run of the whole plan: per-call counts ----
    UnitDev U;
    std::memset(&U, 0, sizeof(U));
    U.total_frames = T;
    U.call_frames = cf;
    U.cap_frames = jobs[0].cap_frames;
    const double out_bound_d = std::ceil(((double)rsb::kInputCapacity + (double)T) / h->ratio) + 2.0;
    if (out_bound_d > 4.0e9 || T > 0x7ff00000ull) return false;
    const uint64_t max_calls = 2 * ((T + cf - 1) / cf) + (U.cap_frames ? (uint64_t)out_bound_d / U.cap_frames : 0) + 4;
    if (max_calls > 0x7fffffffull) return false;
    U.max_calls = (uint32_t)max_calls;
    U.call_cap = U.max_calls;
    U.seg_cap = 0;      // segments are not kept (status 1 = "segment bound" is expected)
    h->pipe.calls.resize(max_calls);
    uint64_t tiles_dummy = 0;
    plan_unit_host(U, rsb::PlanState{h->m_pos[rep], h->m_avail[rep]}, h->ratio, h->taps, nullptr,
                   h->pipe.calls.data(), rsb::kTileOut, tiles_dummy);
    if (U.n_calls >= U.max_calls || U.total_copied != T || U.n_calls < 16) return false;
    for (uint32_t i = 0; i < n; ++i)
        if (jobs[i].out_capacity < U.total_out) return false;     // the unsliced path reports it
    const uint32_t n_calls = U.n_calls;
    const rsb::CallCounts *cc = h->pipe.calls.data();

    // ---- slice boundaries (calls per slice) ----
    const uint64_t row_bytes_per_call = (uint64_t)n * cf * ch * sizeof(float);
    const uint64_t want_slices = getenv("RSB_HOST_PIPE_SLICES") ? std::max(3, atoi(getenv("RSB_HOST_PIPE_SLICES"))) : 48;
    uint64_t cps = (n_calls + want_slices - 1) / want_slices;
    cps = std::max<uint64_t>(cps, 8);
    cps = std::min<uint64_t>(cps, std::max<uint64_t>(1, (512ull << 20) / std::max<uint64_t>(row_bytes_per_call, 1)));
    const uint32_t n_slices = (uint32_t)((n_calls + cps - 1) / cps);
    if (n_slices < 3) return false;
    struct Slice { uint64_t in_off, in_frames, out_off, out_frames; };
    std::vector<Slice> sl(n_slices);
    uint64_t in_acc = 0, out_acc = 0, max_in = 0, max_out = 0;
    for (uint32_t k = 0, c = 0; k < n_slices; ++k) {
        Slice S{in_acc, 0, out_acc, 0};
        for (uint64_t j = 0; j < cps && c < n_calls; ++j, ++c) {
            S.in_frames += cc[c].copied;
            S.out_frames += cc[c].produced;
        }
        in_acc += S.in_frames;
        out_acc += S.out_frames;
        max_in = std::max(max_in, S.in_frames);
        max_out = std::max(max_out, S.out_frames);
        sl[k] = S;
    }

    auto &Pp = h->pipe;
    auto cuda_ok = [&](cudaError_t e) {
        if (e == cudaSuccess) return true;
        *rc_out = fail(RSB_ERR_CUDA, cudaGetErrorString(e));
        return false;
    };
    if (!Pp.s_in) {
        if (!cuda_ok(cudaStreamCreateWithFlags(&Pp.s_in, cudaStreamNonBlocking))) return true;
        if (!cuda_ok(cudaStreamCreateWithFlags(&Pp.s_out, cudaStreamNonBlocking))) return true;
        for (int b = 0; b < 2; ++b) {
            if (!cuda_ok(cudaEventCreateWithFlags(&Pp.ev_h2d[b], cudaEventDisableTiming))) return true;
            if (!cuda_ok(cudaEventCreateWithFlags(&Pp.ev_conv[b], cudaEventDisableTiming))) return true;
            if (!cuda_ok(cudaEventCreateWithFlags(&Pp.ev_d2h[b], cudaEventDisableTiming))) return true;
        }
    }
    const size_t in_pitch = (((size_t)max_in * ch * sizeof(float)) + 15) & ~(size_t)15;
    const size_t out_pitch = (((size_t)max_out * ch * sizeof(float)) + 15) & ~(size_t)15;
    for (int b = 0; b < 2; ++b) {
        if (!cuda_ok(Pp.d_in[b].reserve(in_pitch * n + 16))) return true;
        if (!cuda_ok(Pp.d_out[b].reserve(out_pitch * n + 16))) return true;
    }
    // host rows at one pitch move with one 2-D copy per slice and direction
    bool in_uniform = n > 1 && !getenv("RSB_HOST_PIPE_1D"), out_uniform = in_uniform;
    ptrdiff_t h_in_pitch = 0, h_out_pitch = 0;
    if (n > 1) {
        h_in_pitch = reinterpret_cast<const char *>(jobs[1].in) - reinterpret_cast<const char *>(jobs[0].in);
        h_out_pitch = reinterpret_cast<char *>(jobs[1].out) - reinterpret_cast<char *>(jobs[0].out);
        for (uint32_t i = 1; i < n; ++i) {
            if (reinterpret_cast<const char *>(jobs[i].in) - reinterpret_cast<const char *>(jobs[0].in) != (ptrdiff_t)i * h_in_pitch)
                in_uniform = false;
            if (reinterpret_cast<char *>(jobs[i].out) - reinterpret_cast<char *>(jobs[0].out) != (ptrdiff_t)i * h_out_pitch)
                out_uniform = false;
        }
        if (h_in_pitch < (ptrdiff_t)(T * ch * sizeof(float)) || (size_t)h_in_pitch > h->mem_pitch) in_uniform = false;
        if (h_out_pitch < (ptrdiff_t)(U.total_out * ch * sizeof(float)) || (size_t)h_out_pitch > h->mem_pitch) out_uniform = false;
    }

    std::vector<JobHost> sj(n);
    int rc = RSB_OK;
    for (uint32_t k = 0; k < n_slices && rc == RSB_OK; ++k) {
        const int b = (int)(Pp.slices & 1u);
        const Slice &S = sl[k];
        const size_t in_w = (size_t)S.in_frames * ch * sizeof(float), out_w = (size_t)S.out_frames * ch * sizeof(float);
        // the slice two before this one (possibly of the previous, asynchronous call) has been
        // convolved out of this input buffer
        if (Pp.slices >= 2 && !cuda_ok(cudaStreamWaitEvent(Pp.s_in, Pp.ev_conv[b], 0))) return true;
        if (in_w) {
            if (in_uniform) {
                if (!cuda_ok(cudaMemcpy2DAsync(Pp.d_in[b].p, in_pitch, jobs[0].in + S.in_off * ch, (size_t)h_in_pitch,
                                               in_w, n, cudaMemcpyHostToDevice, Pp.s_in))) return true;
            } else {
                for (uint32_t i = 0; i < n; ++i)
                    if (!cuda_ok(cudaMemcpyAsync(static_cast<char *>(Pp.d_in[b].p) + i * in_pitch, jobs[i].in + S.in_off * ch,
                                                 in_w, cudaMemcpyHostToDevice, Pp.s_in))) return true;
            }
        }
        if (!cuda_ok(cudaEventRecord(Pp.ev_h2d[b], Pp.s_in))) return true;
        if (!cuda_ok(cudaStreamWaitEvent(h->stream, Pp.ev_h2d[b], 0))) return true;
        // slice k-2's results have left this output buffer
        if (Pp.slices >= 2 && !cuda_ok(cudaStreamWaitEvent(h->stream, Pp.ev_d2h[b], 0))) return true;
        for (uint32_t i = 0; i < n; ++i) {
            sj[i] = jobs[i];
            sj[i].in = reinterpret_cast<const float *>(static_cast<const char *>(Pp.d_in[b].p) + i * in_pitch);
            sj[i].out = reinterpret_cast<float *>(static_cast<char *>(Pp.d_out[b].p) + i * out_pitch);
            sj[i].total_frames = S.in_frames;
            sj[i].out_capacity = S.out_frames;
        }
        rc = run_batch(h, sj, false, RSB_MEM_DEVICE, RSB_FLAG_ASYNC, nullptr, nullptr, nullptr);
        if (rc != RSB_OK) break;
        if (!cuda_ok(cudaEventRecord(Pp.ev_conv[b], h->stream))) return true;
        if (!cuda_ok(cudaStreamWaitEvent(Pp.s_out, Pp.ev_conv[b], 0))) return true;
        if (out_w) {
            if (out_uniform) {
                if (!cuda_ok(cudaMemcpy2DAsync(jobs[0].out + S.out_off * ch, (size_t)h_out_pitch, Pp.d_out[b].p, out_pitch,
                                               out_w, n, cudaMemcpyDeviceToHost, Pp.s_out))) return true;
            } else {
                for (uint32_t i = 0; i < n; ++i)
                    if (!cuda_ok(cudaMemcpyAsync(jobs[i].out + S.out_off * ch, static_cast<char *>(Pp.d_out[b].p) + i * out_pitch,
                                                 out_w, cudaMemcpyDeviceToHost, Pp.s_out))) return true;
            }
        }
        if (!cuda_ok(cudaEventRecord(Pp.ev_d2h[b], Pp.s_out))) return true;
        Pp.slices += 1;
    }
    if (!async || rc != RSB_OK) {
        const int rf = finalize_all(h);
        if (rc == RSB_OK) rc = rf;
        if (!cuda_ok(cudaStreamSynchronize(Pp.s_out))) return true;
        if (!cuda_ok(cudaStreamSynchronize(h->stream))) return true;
    }
    Pp.batches += 1;
    if (rc == RSB_OK) {
        for (uint32_t i = 0; i < n; ++i) {
            if (consumed) consumed[i] = (size_t)T * ch;
            if (produced) produced[i] = (size_t)U.total_out * ch;
            if (n_calls_out) n_calls_out[i] = n_calls;
        }
    }
    *rc_out = rc;
    return true;
}

int check_handle(const rsb_fir *h) {
    if (!h) return fail(RSB_ERR_INVALID_ARGUMENT, "null handle");
    return RSB_OK;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char *rsb_version(void) { return "resampler_b200 0.1.0 (sm_100a)"; }

const char *rsb_status_string(int status) {
    switch (status) {
        case RSB_OK: return "ok";
        case RSB_ERR_INVALID_INPUT_BUFFER_SIZE: return "Input buffer size is invalid";    // error.rs:13
        case RSB_ERR_INVALID_OUTPUT_BUFFER_SIZE: return "Output buffer size is invalid";  // error.rs:14
        case RSB_ERR_ZERO_INPUT_RATE: return "input sample rate must be greater than zero";
        case RSB_ERR_ZERO_OUTPUT_RATE: return "output sample rate must be greater than zero";
        case RSB_ERR_INVALID_ARGUMENT: return "invalid argument";
        case RSB_ERR_OUTPUT_CAPACITY: return "output buffer too small for the batch";
        case RSB_ERR_NO_DEVICE: return "no usable CUDA device";
        case RSB_ERR_CUDA: return "CUDA error";
        case RSB_ERR_OUT_OF_MEMORY: return "out of device memory";
        case RSB_ERR_PLAN_OVERFLOW: return "plan workspace overflow";
        default: return "unknown status";
    }
}

const char *rsb_last_error(void) { return g_last_error.c_str(); }

int rsb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int rsb_fir_create(rsb_fir **out, int device, uint32_t n_streams, uint32_t channels,
                   uint32_t input_rate_hz, uint32_t output_rate_hz, int latency, int attenuation) {
    if (!out) return fail(RSB_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    // same order as the reference's asserts (resampler_fir.rs:302-309)
    if (input_rate_hz == 0) return fail(RSB_ERR_ZERO_INPUT_RATE, rsb_status_string(RSB_ERR_ZERO_INPUT_RATE));
    if (output_rate_hz == 0) return fail(RSB_ERR_ZERO_OUTPUT_RATE, rsb_status_string(RSB_ERR_ZERO_OUTPUT_RATE));
    const int taps = rsb::latency_to_taps(latency);
    const double beta = rsb::attenuation_to_beta(attenuation);
    if (taps < 0 || beta < 0.0) return fail(RSB_ERR_INVALID_ARGUMENT, "bad latency/attenuation");
    if (channels == 0 || n_streams == 0)
        return fail(RSB_ERR_INVALID_ARGUMENT, "channels and n_streams must be positive");
    int n_dev = rsb_device_count();
    if (n_dev <= 0) return fail(RSB_ERR_NO_DEVICE, "no CUDA device visible; there is no CPU fallback");
    if (device < 0 || device >= n_dev) return fail(RSB_ERR_INVALID_ARGUMENT, "bad device index");
    RSB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RSB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(RSB_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major) +
                                           std::to_string(prop.minor) +
                                           "; this library contains sm_100a code only");

    std::unique_ptr<rsb_fir> h(new rsb_fir());
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->mem_pitch = prop.memPitch;
    h->n_streams = n_streams;
    h->channels = channels;
    h->in_hz = input_rate_hz;
    h->out_hz = output_rate_hz;
    h->latency = latency;
    h->attenuation = attenuation;
    h->taps = (uint32_t)taps;
    h->ratio = (double)input_rate_hz / (double)output_rate_hz;   // :311-313
    const float cutoff = rsb::design_cutoff(input_rate_hz, output_rate_hz, h->taps, beta);
    // cache miss: the host design, or (rsb_set_device_filter_design) the same table designed on the GPU
    h->table = rsb::get_or_create_table(
        cutoff, h->taps, attenuation,
        g_device_design ? +[](float c, uint32_t t, double b, float *out, void *ctx) {
            return rsb::design_table_on_device(*static_cast<int *>(ctx), c, t, b, out, nullptr);
        } : nullptr,
        &device);
    h->cohort.assign(n_streams, 0);
    h->hist_sel.assign(n_streams, 0);
    h->m_pos.assign(n_streams, 0.0);
    h->m_avail.assign(n_streams, 0u);
    h->m_ok.assign(n_streams, 1);
    h->m_pending.assign(n_streams, 0);
    h->m_fused_seq.assign(n_streams, 0);
    h->seen_epoch.assign(n_streams, 0);

    RSB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    RSB_CUDA(cudaStreamCreateWithFlags(&h->plan_stream, cudaStreamNonBlocking));
    RSB_CUDA(cudaEventCreate(&h->ev_t0));
    RSB_CUDA(cudaEventCreate(&h->ev_t1));
    RSB_CUDA(cudaEventCreateWithFlags(&h->ev_sync, cudaEventDisableTiming));
    for (cudaEvent_t &e : h->ev_pcm) RSB_CUDA(cudaEventCreate(&e));
    for (Workspace &W : h->ws) {
        RSB_CUDA(cudaEventCreateWithFlags(&W.ev_plan, cudaEventDisableTiming));
        RSB_CUDA(cudaEventCreateWithFlags(&W.ev_done, cudaEventDisableTiming));
    }
    for (int i = 0; i < rsb_fir::kConvRing; ++i) {
        RSB_CUDA(cudaEventCreate(&h->ev_conv[i][0]));
        RSB_CUDA(cudaEventCreate(&h->ev_conv[i][1]));
    }
    const size_t tab_bytes = h->table->coeffs.size() * sizeof(float);
    RSB_CUDA(cudaMalloc(&h->d_coeffs, tab_bytes));
    RSB_CUDA(cudaMemcpyAsync(h->d_coeffs, h->table->coeffs.data(), tab_bytes,
                             cudaMemcpyHostToDevice, h->stream));
    const size_t hist_bytes = (size_t)n_streams * rsb::kHistFrames * channels * sizeof(float);
    RSB_CUDA(cudaMalloc(&h->st.position, sizeof(double) * n_streams));
    RSB_CUDA(cudaMalloc(&h->st.hist_len, sizeof(uint32_t) * n_streams));
    RSB_CUDA(cudaMalloc(&h->st.hist[0], hist_bytes));
    RSB_CUDA(cudaMalloc(&h->st.hist[1], hist_bytes));
    RSB_CUDA(cudaMemsetAsync(h->st.position, 0, sizeof(double) * n_streams, h->stream));
    RSB_CUDA(cudaMemsetAsync(h->st.hist_len, 0, sizeof(uint32_t) * n_streams, h->stream));
    RSB_CUDA(cudaMemsetAsync(h->st.hist[0], 0, hist_bytes, h->stream));   // :329 zero-filled
    RSB_CUDA(cudaMemsetAsync(h->st.hist[1], 0, hist_bytes, h->stream));
    RSB_CUDA(cudaStreamSynchronize(h->stream));
    *out = h.release();
    return RSB_OK;
}

void rsb_fir_destroy(rsb_fir *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->plan_stream) cudaStreamSynchronize(h->plan_stream);
    cudaFree(h->d_coeffs);
    cudaFree(h->st.position);
    cudaFree(h->st.hist_len);
    cudaFree(h->st.hist[0]);
    cudaFree(h->st.hist[1]);
    for (Workspace &W : h->ws) {
        for (DevBuf *b : {&W.d_units, &W.d_jobs, &W.d_segs, &W.d_calls, &W.d_tiles, &W.d_entries,
                          &W.d_gtiles, &W.d_tct, &W.d_counter, &W.d_pcm_jobs})
            b->release();
        for (PinBuf *b : {&W.h_units, &W.h_jobs, &W.h_units_back, &W.h_calls_back, &W.h_segs,
                          &W.h_counter, &W.h_pcm_jobs})
            b->release();
        if (W.ev_plan) cudaEventDestroy(W.ev_plan);
        if (W.ev_done) cudaEventDestroy(W.ev_done);
    }
    for (DevBuf *b : {&h->d_stage_in, &h->d_stage_out, &h->d_dbg, &h->d_pcm_raw, &h->d_zero, &h->d_tct2, &h->d_gmat2})
        b->release();
    for (auto &F : h->fused) {
        F.h_jobs.release(); F.h_res.release(); F.d_jobs.release(); F.d_res.release(); F.d_segs.release();
        if (F.ev_done) cudaEventDestroy(F.ev_done);
        if (F.ev_plan) cudaEventDestroy(F.ev_plan);
    }
    for (int b = 0; b < 2; ++b) {
        h->pipe.d_in[b].release();
        h->pipe.d_out[b].release();
        for (cudaEvent_t e : {h->pipe.ev_h2d[b], h->pipe.ev_conv[b], h->pipe.ev_d2h[b]}) if (e) cudaEventDestroy(e);
    }
    if (h->pipe.s_in) cudaStreamDestroy(h->pipe.s_in);
    if (h->pipe.s_out) cudaStreamDestroy(h->pipe.s_out);
    for (cudaEvent_t e : h->ev_pcm) if (e) cudaEventDestroy(e);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    if (h->ev_sync) cudaEventDestroy(h->ev_sync);
    for (int i = 0; i < rsb_fir::kConvRing; ++i)
        for (int j = 0; j < 2; ++j)
            if (h->ev_conv[i][j]) cudaEventDestroy(h->ev_conv[i][j]);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->plan_stream) cudaStreamDestroy(h->plan_stream);
    delete h;
}

int rsb_fir_set_kernel(rsb_fir *h, int kernel) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    if (kernel < RSB_KERNEL_AUTO || kernel > RSB_KERNEL_TENSOR)
        return fail(RSB_ERR_INVALID_ARGUMENT, "bad kernel id");
    if (kernel == RSB_KERNEL_TENSOR && !rsb::tc2_supported(h->channels, h->taps, h->ratio) &&
        !rsb::tc_supported(h->channels, h->taps, h->ratio))
        return fail(RSB_ERR_INVALID_ARGUMENT, "tensor kernel does not support this configuration");
    if (kernel == RSB_KERNEL_FAST && !rsb::fast_supported(h->channels, h->taps, h->ratio))
        return fail(RSB_ERR_INVALID_ARGUMENT, "fast kernel does not support this configuration");
    h->kernel_mode = kernel;
    return RSB_OK;
}

uint32_t rsb_fir_channels(const rsb_fir *h) { return h ? h->channels : 0; }
int rsb_fir_last_kernel(const rsb_fir *h) { return h ? h->last_kernel : RSB_KERNEL_AUTO; }
uint32_t rsb_fir_n_streams(const rsb_fir *h) { return h ? h->n_streams : 0; }
uint32_t rsb_fir_taps(const rsb_fir *h) { return h ? h->taps : 0; }
double rsb_fir_ratio(const rsb_fir *h) { return h ? h->ratio : 0.0; }
int rsb_fir_device(const rsb_fir *h) { return h ? h->device : -1; }

size_t rsb_fir_buffer_size_output(const rsb_fir *h) {
    if (!h) return 0;
    // resampler_fir.rs:456-465
    const double usable = (double)(rsb::kInputCapacity - h->taps);
    const size_t frames = (size_t)std::ceil(usable / h->ratio) + 2;
    return frames * h->channels;
}

size_t rsb_fir_delay(const rsb_fir *h) { return h ? h->taps / 2 : 0; }   // :630-632

int rsb_fir_reset(rsb_fir *h, int64_t stream) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    RSB_CUDA(cudaSetDevice(h->device));
    if (stream >= (int64_t)h->n_streams) return fail(RSB_ERR_INVALID_ARGUMENT, "bad stream index");
    // the scalar state lives on the plan stream (ordered with every plan kernel)
    if (stream < 0) {
        rsb::launch_reset(h->st, 0, h->n_streams, h->plan_stream);
        std::fill(h->cohort.begin(), h->cohort.end(), 0);
        std::fill(h->m_pos.begin(), h->m_pos.end(), 0.0);
        std::fill(h->m_avail.begin(), h->m_avail.end(), 0u);
        std::fill(h->m_ok.begin(), h->m_ok.end(), 1);
    } else {
        rsb::launch_reset(h->st, (uint32_t)stream, 1, h->plan_stream);
        h->cohort[(size_t)stream] = 0;
        h->m_pos[(size_t)stream] = 0.0;
        h->m_avail[(size_t)stream] = 0u;
        h->m_ok[(size_t)stream] = 1;
    }
    h->launches += 1;
    RSB_CUDA(cudaGetLastError());
    return RSB_OK;
}

int rsb_fir_coeffs(const rsb_fir *h, float *out, size_t out_len) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    if (!out || out_len < h->table->coeffs.size())
        return fail(RSB_ERR_INVALID_ARGUMENT, "coefficient buffer too small");
    // read back from the device copy: this is what the kernels use
    RSB_CUDA(cudaSetDevice(h->device));
    RSB_CUDA(cudaMemcpy(out, h->d_coeffs, h->table->coeffs.size() * sizeof(float),
                        cudaMemcpyDeviceToHost));
    return RSB_OK;
}

int rsb_fir_resample(rsb_fir *h, uint32_t stream, const float *input, size_t input_len,
                     float *output, size_t output_len, size_t *consumed, size_t *produced) {
    const float *in_ptrs[1] = {input};
    float *out_ptrs[1] = {output};
    return rsb_fir_submit_batch(h, 1, &stream, in_ptrs, &input_len, out_ptrs, &output_len, consumed,
                                produced, RSB_MEM_HOST, 0);
}

int rsb_fir_submit_batch(rsb_fir *h, uint32_t n, const uint32_t *streams, const float *const *in,
                         const size_t *in_lens, float *const *out, const size_t *out_lens,
                         size_t *consumed, size_t *produced, int memspace, uint32_t flags) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    if (n == 0) return RSB_OK;
    if (!in || !in_lens || !out || !out_lens)
        return fail(RSB_ERR_INVALID_ARGUMENT, "null array argument");
    if (memspace != RSB_MEM_DEVICE && memspace != RSB_MEM_HOST)
        return fail(RSB_ERR_INVALID_ARGUMENT, "bad memspace");
    const uint32_t ch = h->channels;
    // Small device-resident submits take the fused path (fir_submit.cu): plan, samples in the
    // reference's order and state update in two launches, per-stream call sizes welcome.  Large
    // ones (tens of millions of samples) are better served by the tile kernels below.
    static const bool no_fused = getenv("RSB_NO_FUSED_SUBMIT") != nullptr;
    if (memspace == RSB_MEM_DEVICE && !(flags & (RSB_FLAG_RECORD_CALLS | RSB_FLAG_KEEP_PLAN)) &&
        (h->kernel_mode == RSB_KERNEL_AUTO || h->kernel_mode == RSB_KERNEL_EXACT) && !no_fused) {
        const int rf = run_submit_fused(h, n, streams, in, in_lens, out, out_lens, consumed, produced, flags);
        if (rf != kNotTaken) return rf;
    }
    // error precedence of resampler_fir.rs:514-519, per job, before any state changes
    for (uint32_t i = 0; i < n; ++i) {
        if (in_lens[i] % ch != 0)
            return fail(RSB_ERR_INVALID_INPUT_BUFFER_SIZE, "input length of job " + std::to_string(i) +
                                                               " is not a multiple of channels");
        if (out_lens[i] % ch != 0)
            return fail(RSB_ERR_INVALID_OUTPUT_BUFFER_SIZE, "output length of job " +
                                                                std::to_string(i) +
                                                                " is not a multiple of channels");
    }
    std::vector<JobHost> jobs(n);
    std::vector<uint8_t> seen(h->n_streams, 0);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t s = streams ? streams[i] : i;
        if (s >= h->n_streams) return fail(RSB_ERR_INVALID_ARGUMENT, "bad stream index");
        if (seen[s]) return fail(RSB_ERR_INVALID_ARGUMENT, "stream listed twice in one batch");
        seen[s] = 1;
        if (in_lens[i] && !in[i]) return fail(RSB_ERR_INVALID_ARGUMENT, "null input pointer");
        if (out_lens[i] && !out[i]) return fail(RSB_ERR_INVALID_ARGUMENT, "null output pointer");
        const uint64_t cap = out_lens[i] / ch;
        // a call never takes more than INPUT_CAPACITY frames (:526-528): clamping what is
        // offered changes nothing and bounds what the kernels may touch
        const uint64_t offered = std::min<uint64_t>(in_lens[i] / ch, rsb::kInputCapacity);
        jobs[i] = JobHost{s, in[i], out[i], offered, cap, 0u,
                          (uint32_t)std::min<uint64_t>(cap, 0xffffffffull)};
    }
    return run_batch(h, jobs, true, memspace, flags, consumed, produced, nullptr);
}

int rsb_fir_process_batch(rsb_fir *h, uint32_t n, const uint32_t *streams, const float *const *in,
                          const size_t *total_lens, size_t call_len, size_t out_cap_len,
                          float *const *out, const size_t *out_capacities, size_t *consumed_totals,
                          size_t *produced_totals, uint32_t *n_calls, int memspace, uint32_t flags) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    if (n == 0) return RSB_OK;
    if (!in || !total_lens || !out || !out_capacities)
        return fail(RSB_ERR_INVALID_ARGUMENT, "null array argument");
    if (memspace != RSB_MEM_DEVICE && memspace != RSB_MEM_HOST)
        return fail(RSB_ERR_INVALID_ARGUMENT, "bad memspace");
    const uint32_t ch = h->channels;
    if (out_cap_len == 0) out_cap_len = rsb_fir_buffer_size_output(h);
    // every call of the loop sees slices of call_len (or the remainder) and out_cap_len values
    if (call_len % ch != 0)
        return fail(RSB_ERR_INVALID_INPUT_BUFFER_SIZE, "call_len is not a multiple of channels");
    if (out_cap_len % ch != 0)
        return fail(RSB_ERR_INVALID_OUTPUT_BUFFER_SIZE, "out_cap_len is not a multiple of channels");
    if (call_len / ch > 0xffffffffull || out_cap_len / ch > 0xffffffffull)
        return fail(RSB_ERR_INVALID_ARGUMENT, "call too large");
    std::vector<JobHost> jobs(n);
    std::vector<uint8_t> seen(h->n_streams, 0);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t s = streams ? streams[i] : i;
        if (s >= h->n_streams) return fail(RSB_ERR_INVALID_ARGUMENT, "bad stream index");
        if (seen[s]) return fail(RSB_ERR_INVALID_ARGUMENT, "stream listed twice in one batch");
        seen[s] = 1;
        if (total_lens[i] % ch != 0)
            return fail(RSB_ERR_INVALID_INPUT_BUFFER_SIZE, "input length of job " + std::to_string(i) +
                                                               " is not a multiple of channels");
        if (total_lens[i] && call_len == 0)
            return fail(RSB_ERR_INVALID_ARGUMENT, "call_len must be positive");
        if (total_lens[i] && !in[i]) return fail(RSB_ERR_INVALID_ARGUMENT, "null input pointer");
        if (out_capacities[i] && !out[i]) return fail(RSB_ERR_INVALID_ARGUMENT, "null output pointer");
        jobs[i] = JobHost{s, in[i], out[i], total_lens[i] / ch, out_capacities[i] / ch,
                          (uint32_t)(call_len / ch), (uint32_t)(out_cap_len / ch)};
    }
    if (memspace == RSB_MEM_HOST && (flags & ~(uint32_t)RSB_FLAG_ASYNC) == 0 && !getenv("RSB_NO_HOST_PIPELINE")) {
        int rc = RSB_OK;
        if (process_host_pipelined(h, jobs, consumed_totals, produced_totals, n_calls,
                                   (flags & RSB_FLAG_ASYNC) != 0, &rc)) return rc;
    }
    return run_batch(h, jobs, false, memspace, flags, consumed_totals, produced_totals, n_calls);
}

// CLI batch path: format step (resample/src/main.rs:128-156) + canonical loop (:226-254)
int rsb_fir_process_pcm_batch(rsb_fir *h, uint32_t n, const uint32_t *streams,
                              const void *const *in, const size_t *in_frames, int format,
                              uint32_t src_channels, size_t call_len, size_t out_cap_len,
                              float *const *out, const size_t *out_capacities,
                              size_t *consumed_totals, size_t *produced_totals, uint32_t *n_calls,
                              int memspace, uint32_t flags) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    if (n == 0) return RSB_OK;
    if (!in || !in_frames || !out || !out_capacities)
        return fail(RSB_ERR_INVALID_ARGUMENT, "null array argument");
    if (memspace != RSB_MEM_DEVICE && memspace != RSB_MEM_HOST)
        return fail(RSB_ERR_INVALID_ARGUMENT, "bad memspace");
    const uint32_t ch = h->channels;
    PcmSpec spec;
    spec.format = format;
    spec.bps = rsb::pcm_bytes_per_sample(format);
    if (spec.bps == 0) return fail(RSB_ERR_INVALID_ARGUMENT, "unknown PCM format");
    // main.rs:139-156: a mono source is duplicated into every channel, a source with the
    // resampler's channel count passes through, anything else is "Unsupported channel count"
    if (src_channels != 1 && src_channels != ch)
        return fail(RSB_ERR_INVALID_ARGUMENT, "Unsupported channel count: " + std::to_string(src_channels));
    spec.dup = src_channels == ch ? 1u : ch;
    if (out_cap_len == 0) out_cap_len = rsb_fir_buffer_size_output(h);
    if (call_len % ch != 0)
        return fail(RSB_ERR_INVALID_INPUT_BUFFER_SIZE, "call_len is not a multiple of channels");
    if (out_cap_len % ch != 0)
        return fail(RSB_ERR_INVALID_OUTPUT_BUFFER_SIZE, "out_cap_len is not a multiple of channels");
    if (call_len / ch > 0xffffffffull || out_cap_len / ch > 0xffffffffull)
        return fail(RSB_ERR_INVALID_ARGUMENT, "call too large");
    std::vector<JobHost> jobs(n);
    std::vector<uint8_t> seen(h->n_streams, 0);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t s = streams ? streams[i] : i;
        if (s >= h->n_streams) return fail(RSB_ERR_INVALID_ARGUMENT, "bad stream index");
        if (seen[s]) return fail(RSB_ERR_INVALID_ARGUMENT, "stream listed twice in one batch");
        seen[s] = 1;
        if (in_frames[i] && call_len == 0)
            return fail(RSB_ERR_INVALID_ARGUMENT, "call_len must be positive");
        if (in_frames[i] && !in[i]) return fail(RSB_ERR_INVALID_ARGUMENT, "null input pointer");
        if (out_capacities[i] && !out[i]) return fail(RSB_ERR_INVALID_ARGUMENT, "null output pointer");
        jobs[i] = JobHost{s, static_cast<const float *>(in[i]), out[i], in_frames[i],
                          out_capacities[i] / ch, (uint32_t)(call_len / ch), (uint32_t)(out_cap_len / ch)};
    }
    return run_batch(h, jobs, false, memspace, flags, consumed_totals, produced_totals, n_calls, &spec);
}

// Opt-in tail handling.  The reference never flushes: its last
// taps - 1 input frames stay in the history and the output lags by delay() = taps / 2 input
// frames (resampler_fir.rs:630-632).  Feeding delay() frames of silence emits the output whose
// filter centre lies on the last real input frames; it is exactly
// resample(&[0.0; delay * channels], ..) per stream, state carried on as after any call.
int rsb_fir_flush_batch(rsb_fir *h, uint32_t n, const uint32_t *streams, float *const *out,
                        const size_t *out_capacities, size_t *produced, int memspace,
                        uint32_t flags) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    if (n == 0) return RSB_OK;
    if (memspace != RSB_MEM_DEVICE && memspace != RSB_MEM_HOST)
        return fail(RSB_ERR_INVALID_ARGUMENT, "bad memspace");
    const size_t len = (size_t)(h->taps / 2) * h->channels;
    std::vector<float> host_zero;
    const float *zero = nullptr;
    if (memspace == RSB_MEM_HOST) {
        host_zero.assign(len, 0.0f);
        zero = host_zero.data();
    } else {
        RSB_CUDA(cudaSetDevice(h->device));
        if (!h->d_zero.p) {
            RSB_CUDA(h->d_zero.reserve(64 * 64 * sizeof(float)));   // taps / 2 <= 64 frames, <= 64 ch
            RSB_CUDA(cudaMemset(h->d_zero.p, 0, h->d_zero.cap));
        }
        if (len * sizeof(float) > h->d_zero.cap) {
            RSB_CUDA(h->d_zero.reserve(len * sizeof(float)));
            RSB_CUDA(cudaMemset(h->d_zero.p, 0, h->d_zero.cap));
        }
        zero = h->d_zero.as<float>();
    }
    std::vector<const float *> in(n, zero);
    std::vector<size_t> lens(n, len);
    return rsb_fir_process_batch(h, n, streams, in.data(), lens.data(), len, 0, out, out_capacities,
                                 nullptr, produced, nullptr, memspace, flags & ~(uint32_t)RSB_FLAG_ASYNC);
}

int rsb_fir_last_pcm_fused(const rsb_fir *h) { return h && h->pcm_fused_last ? 1 : 0; }

int rsb_fir_last_ingest_ms(rsb_fir *h, float *ms) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    if (!ms) return fail(RSB_ERR_INVALID_ARGUMENT, "null argument");
    if (!h->pcm_timed) return fail(RSB_ERR_INVALID_ARGUMENT, "no PCM batch has run on this handle");
    RSB_CUDA(cudaSetDevice(h->device));
    RSB_CUDA(cudaEventSynchronize(h->ev_pcm[1]));
    RSB_CUDA(cudaEventElapsedTime(ms, h->ev_pcm[0], h->ev_pcm[1]));
    return RSB_OK;
}

int rsb_fir_sync(rsb_fir *h) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    RSB_CUDA(cudaSetDevice(h->device));
    int rc = finalize_all(h);
    const int rf = finalize_fused_all(h);
    if (rc == RSB_OK) rc = rf;
    if (h->deferred_rc != RSB_OK) {      // an older asynchronous submit failed: reported here, once
        if (rc == RSB_OK) {
            rc = h->deferred_rc;
            g_last_error = h->deferred_msg;
        }
        h->deferred_rc = RSB_OK;
    }
    RSB_CUDA(cudaStreamSynchronize(h->plan_stream));
    RSB_CUDA(cudaStreamSynchronize(h->stream));
    if (h->pipe.s_out) RSB_CUDA(cudaStreamSynchronize(h->pipe.s_out));
    return rc;
}

int rsb_fir_last_call_counts(rsb_fir *h, uint32_t job, uint32_t *consumed, uint32_t *produced,
                             size_t max_calls, size_t *n_calls) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    int rc = finalize_all(h);
    if (rc != RSB_OK && rc != RSB_ERR_OUTPUT_CAPACITY) return rc;
    Workspace &W = h->last_ws();
    if (!W.has_calls) return fail(RSB_ERR_INVALID_ARGUMENT, "last batch did not record calls");
    if (job >= W.job_unit.size()) return fail(RSB_ERR_INVALID_ARGUMENT, "bad job index");
    RSB_CUDA(cudaEventSynchronize(W.ev_plan));
    const UnitDev &U = W.h_units_back.as<UnitDev>()[W.job_unit[job]];
    const rsb::CallCounts *cc = W.h_calls_back.as<rsb::CallCounts>() + U.call_off;
    const size_t nc = std::min<size_t>(U.n_calls, U.call_cap);
    if (n_calls) *n_calls = nc;
    for (size_t i = 0; i < nc && i < max_calls; ++i) {
        if (consumed) consumed[i] = cc[i].copied * h->channels;
        if (produced) produced[i] = cc[i].produced * h->channels;
    }
    return RSB_OK;
}

int rsb_fir_last_plan(rsb_fir *h, uint32_t job, uint32_t *input_offset, uint32_t *phase1,
                      uint32_t *phase2, uint32_t *frac_bits, size_t max_frames, size_t *n_frames) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    int rc = finalize_all(h);
    if (rc != RSB_OK && rc != RSB_ERR_OUTPUT_CAPACITY) return rc;
    Workspace &W = h->last_ws();
    if (!W.has_plan) return fail(RSB_ERR_INVALID_ARGUMENT, "last batch did not keep its plan");
    if (job >= W.job_unit.size()) return fail(RSB_ERR_INVALID_ARGUMENT, "bad job index");
    RSB_CUDA(cudaSetDevice(h->device));
    RSB_CUDA(cudaEventSynchronize(W.ev_plan));
    const uint32_t unit = W.job_unit[job];
    const UnitDev &U = W.h_units_back.as<UnitDev>()[unit];
    const size_t total = (size_t)U.total_out;
    if (n_frames) *n_frames = total;
    const size_t cnt = std::min(total, max_frames);
    if (cnt == 0) return RSB_OK;
    RSB_CUDA(h->d_dbg.reserve(cnt * 4 * sizeof(uint32_t)));
    uint32_t *d = h->d_dbg.as<uint32_t>();
    rsb::launch_expand_plan(W.d_units.as<UnitDev>(), unit, W.d_entries.as<rsb::PlanEntry>(),
                            (uint32_t)cnt, d, d + cnt, d + 2 * cnt, d + 3 * cnt, h->stream);
    h->launches += 1;
    RSB_CUDA(cudaGetLastError());
    RSB_CUDA(cudaStreamSynchronize(h->stream));
    uint32_t *dst[4] = {input_offset, phase1, phase2, frac_bits};
    for (int k = 0; k < 4; ++k)
        if (dst[k])
            RSB_CUDA(cudaMemcpy(dst[k], d + (size_t)k * cnt, cnt * sizeof(uint32_t),
                                cudaMemcpyDeviceToHost));
    return RSB_OK;
}

int rsb_fir_timer_start(rsb_fir *h) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    RSB_CUDA(cudaSetDevice(h->device));
    // work enqueued after this call must not start before the start event on either stream
    RSB_CUDA(cudaEventRecord(h->ev_sync, h->plan_stream));
    RSB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_sync, 0));
    RSB_CUDA(cudaEventRecord(h->ev_t0, h->stream));
    RSB_CUDA(cudaStreamWaitEvent(h->plan_stream, h->ev_t0, 0));
    return RSB_OK;
}

int rsb_fir_timer_stop(rsb_fir *h, float *elapsed_ms) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    RSB_CUDA(cudaSetDevice(h->device));
    RSB_CUDA(cudaEventRecord(h->ev_sync, h->plan_stream));
    RSB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_sync, 0));
    RSB_CUDA(cudaEventRecord(h->ev_t1, h->stream));
    RSB_CUDA(cudaEventSynchronize(h->ev_t1));
    float ms = 0.0f;
    RSB_CUDA(cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1));
    if (elapsed_ms) *elapsed_ms = ms;
    return RSB_OK;
}

int rsb_fir_conv_times(rsb_fir *h, float *ms, size_t max, size_t *n) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    RSB_CUDA(cudaSetDevice(h->device));
    RSB_CUDA(cudaStreamSynchronize(h->stream));
    const uint64_t have = std::min<uint64_t>(h->conv_batches, rsb_fir::kConvRing);
    const size_t cnt = (size_t)std::min<uint64_t>(have, max);
    if (n) *n = cnt;
    for (size_t i = 0; i < cnt; ++i) {
        // most recent first
        const int slot = (int)((h->conv_batches - 1 - i) % rsb_fir::kConvRing);
        float t = 0.f;
        RSB_CUDA(cudaEventElapsedTime(&t, h->ev_conv[slot][0], h->ev_conv[slot][1]));
        if (ms) ms[i] = t;
    }
    return RSB_OK;
}

int rsb_debug_phase_cycles(rsb_fir *h, int enable, uint64_t *out8) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    RSB_CUDA(cudaSetDevice(h->device));
    RSB_CUDA(cudaStreamSynchronize(h->stream));
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "u64");
    rsb::fast_phase_profile(enable, reinterpret_cast<unsigned long long *>(out8));
    RSB_CUDA(cudaGetLastError());
    return RSB_OK;
}

int rsb_debug_tc_hang(uint32_t out32[32]) {
    if (!out32) return RSB_ERR_INVALID_ARGUMENT;
    rsb::tc2_hang_record(out32);
    return RSB_OK;
}

int rsb_debug_tc_cycles(rsb_fir *h, int enable, uint64_t *out, uint32_t count) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    RSB_CUDA(cudaSetDevice(h->device));
    RSB_CUDA(cudaDeviceSynchronize());
    if (getenv("RSB_TC_V1")) {
        unsigned long long tmp[24] = {0};
        rsb::tc_phase_profile(enable, tmp);
        if (out) std::memcpy(out, tmp, sizeof(uint64_t) * std::min<uint32_t>(count, 24u));
    } else {
        rsb::tc2_phase_profile(enable, reinterpret_cast<unsigned long long *>(out), count);
    }
    RSB_CUDA(cudaGetLastError());
    return RSB_OK;
}

uint64_t rsb_fir_launch_count(const rsb_fir *h) { return h ? h->launches : 0; }

int rsb_fir_host_pipeline_stats(const rsb_fir *h, uint64_t *batches, uint64_t *slices) {
    if (check_handle(h)) return RSB_ERR_INVALID_ARGUMENT;
    if (batches) *batches = h->pipe.batches;
    if (slices) *slices = h->pipe.slices;
    return RSB_OK;
}

// Pinned-memory copy rates of this GPU's PCIe link (GB/s): mode 0 host->device alone, 1 device->host
// alone, 2 both directions at once (out[0] = host->device, out[1] = device->host).  Plain
// cudaMemcpyAsync on two streams: the ceiling the host-memspace calls are measured against.
int rsb_pcie_probe(int device, size_t bytes, int iters, int mode, double out[2]) {
    if (!out || bytes == 0 || iters <= 0 || mode < 0 || mode > 3)
        return fail(RSB_ERR_INVALID_ARGUMENT, "bad probe arguments");
    if (cudaSetDevice(device) != cudaSuccess) return fail(RSB_ERR_NO_DEVICE, "no such device");
    out[0] = out[1] = 0.0;
    void *h_up = nullptr, *h_dn = nullptr, *d_up = nullptr, *d_dn = nullptr;
    cudaStream_t s_up = nullptr, s_dn = nullptr;
    cudaEvent_t e[4] = {};
    int rc = RSB_OK;
    auto ok = [&](cudaError_t err) {
        if (err != cudaSuccess && rc == RSB_OK) rc = fail(RSB_ERR_CUDA, cudaGetErrorString(err));
        return err == cudaSuccess;
    };
    // mode 3: both directions as 2-D copies of 1024 rows (the shape of a host-memspace batch slice:
    // one row per stream, host rows a whole stream apart = twice the row here)
    const size_t rows = 1024, row_bytes = ((bytes / rows) + 15) & ~(size_t)15;
    const size_t host_bytes = mode == 3 ? 2 * row_bytes * rows : bytes;
    if (mode == 3) bytes = row_bytes * rows;
    ok(cudaMallocHost(&h_up, host_bytes)) && ok(cudaMallocHost(&h_dn, host_bytes)) && ok(cudaMalloc(&d_up, bytes)) &&
        ok(cudaMalloc(&d_dn, bytes)) && ok(cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking)) &&
        ok(cudaStreamCreateWithFlags(&s_dn, cudaStreamNonBlocking));
    for (int i = 0; i < 4 && rc == RSB_OK; ++i) ok(cudaEventCreate(&e[i]));
    if (rc == RSB_OK) {
        std::memset(h_up, 1, host_bytes);
        std::memset(h_dn, 0, host_bytes);
        const bool up = mode != 1, dn = mode != 0;
        for (int pass = 0; pass < 2 && rc == RSB_OK; ++pass) {      // pass 0 warms up
            const int reps = pass == 0 ? 1 : iters;
            if (up) ok(cudaEventRecord(e[0], s_up));
            if (dn) ok(cudaEventRecord(e[2], s_dn));
            for (int i = 0; i < reps; ++i) {
                if (mode == 3) {
                    ok(cudaMemcpy2DAsync(d_up, row_bytes, h_up, 2 * row_bytes, row_bytes, rows, cudaMemcpyHostToDevice, s_up));
                    ok(cudaMemcpy2DAsync(h_dn, 2 * row_bytes, d_dn, row_bytes, row_bytes, rows, cudaMemcpyDeviceToHost, s_dn));
                } else {
                    if (up) ok(cudaMemcpyAsync(d_up, h_up, bytes, cudaMemcpyHostToDevice, s_up));
                    if (dn) ok(cudaMemcpyAsync(h_dn, d_dn, bytes, cudaMemcpyDeviceToHost, s_dn));
                }
            }
            if (up) ok(cudaEventRecord(e[1], s_up));
            if (dn) ok(cudaEventRecord(e[3], s_dn));
            if (up) ok(cudaEventSynchronize(e[1]));
            if (dn) ok(cudaEventSynchronize(e[3]));
            if (pass == 1 && rc == RSB_OK) {
                float ms = 0.f;
                if (up && ok(cudaEventElapsedTime(&ms, e[0], e[1]))) out[0] = (double)bytes * reps / (ms * 1e-3) / 1e9;
                if (dn && ok(cudaEventElapsedTime(&ms, e[2], e[3]))) out[1] = (double)bytes * reps / (ms * 1e-3) / 1e9;
            }
        }
    }
    for (cudaEvent_t ev : e) if (ev) cudaEventDestroy(ev);
    if (s_up) cudaStreamDestroy(s_up);
    if (s_dn) cudaStreamDestroy(s_dn);
    if (d_up) cudaFree(d_up);
    if (d_dn) cudaFree(d_dn);
    if (h_up) cudaFreeHost(h_up);
    if (h_dn) cudaFreeHost(h_dn);
    return rc;
}
void *rsb_fir_cuda_stream(const rsb_fir *h) { return h ? (void *)h->stream : nullptr; }

void *rsb_alloc_pinned(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        g_last_error = "cudaMallocHost failed";
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void rsb_free_pinned(void *p) { if (p) cudaFreeHost(p); }

void *rsb_alloc_device(int device, size_t bytes) {
    void *p = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        g_last_error = "cudaMalloc failed";
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void rsb_free_device(int device, void *p) {
    if (!p) return;
    cudaSetDevice(device);
    cudaFree(p);
}

int rsb_memcpy(int device, void *dst, const void *src, size_t bytes, int kind) {
    RSB_CUDA(cudaSetDevice(device));
    cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice
                     : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    RSB_CUDA(cudaMemcpy(dst, src, bytes, k));
    return RSB_OK;
}

int rsb_fill_synthetic(int device, float *dst, uint32_t first_stream, uint32_t n_streams,
                       uint64_t frames, uint32_t channels, uint32_t rate_hz, uint64_t seed) {
    if (!dst || rate_hz == 0) return fail(RSB_ERR_INVALID_ARGUMENT, "bad argument");
    RSB_CUDA(cudaSetDevice(device));
    rsb::launch_fill_synthetic(dst, first_stream, n_streams, frames, channels, rate_hz, seed, 0);
    RSB_CUDA(cudaGetLastError());
    RSB_CUDA(cudaDeviceSynchronize());
    return RSB_OK;
}

// ---- host-only entry points ----
double rsb_host_bessel_i0(double x) { return rsb::bessel_i0(x); }
double rsb_host_cutoff_kaiser(uint32_t taps, double beta) { return rsb::kaiser_cutoff(taps, beta); }
void rsb_host_kaiser_window(uint32_t n, double beta, int symmetric, float *out) {
    rsb::make_kaiser_window(n, beta, symmetric != 0, out);
}
void rsb_host_make_sincs(uint32_t sample_count, uint32_t factor, float cutoff, double beta,
                         int symmetric, float *out) {
    rsb::make_sincs_for_kaiser(sample_count, factor, cutoff, beta, symmetric != 0, out);
}

float rsb_host_sinf_restated(float x) { return rsb::sinf_restated(x); }

int rsb_set_device_filter_design(int enable) {
    const int before = g_device_design ? 1 : 0;
    g_device_design = enable != 0;
    return before;
}

int rsb_device_design_table(int device, uint32_t input_rate_hz, uint32_t output_rate_hz, int latency,
                            int attenuation, float *out, size_t out_len, float *elapsed_ms) {
    if (input_rate_hz == 0) return fail(RSB_ERR_ZERO_INPUT_RATE, rsb_status_string(RSB_ERR_ZERO_INPUT_RATE));
    if (output_rate_hz == 0) return fail(RSB_ERR_ZERO_OUTPUT_RATE, rsb_status_string(RSB_ERR_ZERO_OUTPUT_RATE));
    const int taps = rsb::latency_to_taps(latency);
    const double beta = rsb::attenuation_to_beta(attenuation);
    if (taps < 0 || beta < 0.0 || !out) return fail(RSB_ERR_INVALID_ARGUMENT, "bad latency/attenuation/buffer");
    if (out_len < (size_t)1024 * taps) return fail(RSB_ERR_INVALID_ARGUMENT, "table buffer too small");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return fail(RSB_ERR_NO_DEVICE, "no such CUDA device");
    }
    const float cutoff = rsb::design_cutoff(input_rate_hz, output_rate_hz, (uint32_t)taps, beta);
    if (!rsb::design_table_on_device(device, cutoff, (uint32_t)taps, beta, out, elapsed_ms))
        return fail(RSB_ERR_CUDA, "device-side filter design failed");
    return RSB_OK;
}

int rsb_host_design_table(uint32_t input_rate_hz, uint32_t output_rate_hz, int latency,
                          int attenuation, float *out, size_t out_len, uint32_t *cutoff_bits) {
    if (input_rate_hz == 0) return fail(RSB_ERR_ZERO_INPUT_RATE, rsb_status_string(RSB_ERR_ZERO_INPUT_RATE));
    if (output_rate_hz == 0) return fail(RSB_ERR_ZERO_OUTPUT_RATE, rsb_status_string(RSB_ERR_ZERO_OUTPUT_RATE));
    const int taps = rsb::latency_to_taps(latency);
    const double beta = rsb::attenuation_to_beta(attenuation);
    if (taps < 0 || beta < 0.0) return fail(RSB_ERR_INVALID_ARGUMENT, "bad latency/attenuation");
    const float cutoff = rsb::design_cutoff(input_rate_hz, output_rate_hz, (uint32_t)taps, beta);
    auto t = rsb::get_or_create_table(cutoff, (uint32_t)taps, attenuation);
    if (cutoff_bits) *cutoff_bits = t->cutoff_bits;
    if (out) {
        if (out_len < t->coeffs.size()) return fail(RSB_ERR_INVALID_ARGUMENT, "table buffer too small");
        std::memcpy(out, t->coeffs.data(), t->coeffs.size() * sizeof(float));
    }
    return RSB_OK;
}

namespace {
struct HostSink {
    std::vector<rsb::PlanSeg> *segs;
    uint32_t out0 = 0;
    int64_t vbase = 0;
    void seg(int64_t base_bits, int64_t step_bits, uint32_t n) {
        segs->push_back(rsb::PlanSeg{base_bits, step_bits, n, out0, vbase});
        out0 += n;
    }
};
}  // namespace

size_t rsb_host_plan(uint32_t input_rate_hz, uint32_t output_rate_hz, int latency,
                     uint64_t *position_bits, uint32_t *available, uint64_t total_frames,
                     uint32_t call_frames, uint32_t cap_frames, int single_call,
                     uint32_t *calls_consumed, uint32_t *calls_produced, size_t max_calls,
                     uint32_t *input_offset, uint32_t *phase1, uint32_t *phase2,
                     uint32_t *frac_bits, size_t max_frames, size_t *n_frames,
                     size_t *n_segments) {
    const int taps = rsb::latency_to_taps(latency);
    if (taps < 0 || input_rate_hz == 0 || output_rate_hz == 0 || !position_bits || !available) {
        if (n_frames) *n_frames = 0;
        return 0;
    }
    const double ratio = (double)input_rate_hz / (double)output_rate_hz;
    rsb::PlanState st;
    st.position = rsb::bits2d((int64_t)*position_bits);
    st.available = *available;
    std::vector<rsb::PlanSeg> segs;
    HostSink sink;
    sink.segs = &segs;
    uint64_t offset = 0;
    int64_t advanced = 0;
    size_t n_calls = 0;
    rsb::DivCache dc;
    rsb::div_cache_reset(dc);
    // mirrors plan_units_kernel (fir_kernels.cu)
    for (;;) {
        if (!single_call && offset >= total_frames) break;
        const uint64_t remaining = total_frames - offset;
        const uint32_t chunk =
            single_call ? (uint32_t)std::min<uint64_t>(remaining, 0xffffffffull)
                        : (uint32_t)std::min<uint64_t>(remaining, call_frames);
        sink.vbase = advanced;
        const rsb::CallResult r = rsb::plan_call(st, ratio, (uint32_t)taps, chunk, cap_frames, sink, dc);
        if (n_calls < max_calls) {
            if (calls_consumed) calls_consumed[n_calls] = r.copied;
            if (calls_produced) calls_produced[n_calls] = r.produced;
        }
        n_calls += 1;
        offset += r.copied;
        advanced += r.advanced;
        if (single_call || r.copied == 0) break;
    }
    size_t k = 0;
    for (const rsb::PlanSeg &sg : segs) {
        for (uint32_t j = 0; j < sg.n; ++j, ++k) {
            if (k >= max_frames) continue;
            const rsb::PhasePoint pp =
                rsb::phase_point(rsb::bits2d(sg.base_bits + (int64_t)j * sg.step_bits));
            if (input_offset) input_offset[k] = pp.off;
            if (phase1) phase1[k] = pp.phase1;
            if (phase2) phase2[k] = pp.phase2;
            if (frac_bits) {
                uint32_t b;
                std::memcpy(&b, &pp.frac, 4);
                frac_bits[k] = b;
            }
        }
    }
    if (n_frames) *n_frames = k;
    if (n_segments) *n_segments = segs.size();
    *position_bits = (uint64_t)rsb::d2bits(st.position);
    *available = st.available;
    return n_calls;
}

}  // extern "C"
