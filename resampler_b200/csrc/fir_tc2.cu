// fir_tc2.cu -- the tensor-core convolution kernel for sm_100a (tcgen05 + TMEM + TMA), generation 2.
//
// Formulation.  For a tile of 64 consecutive output frames,  OUT[row][k] = sum_j X[row][j] * G[k][j]
// with row = one channel of one member stream and G the tile's interpolated, banded filter rows
// (the reference blends two dot products per output, fir/avx512.rs:41-45; blending the rows first
// is the same linear map).  The product runs on the 5th-generation tensor cores as
// D[128 rows x 64 outputs] += A[128 x 16] * B[64 x 16]^T  per K step of 16 input frames
// (tcgen05.mma kind::f16, M = 128, N = 64, fp32 accumulator in tensor memory).
//
// Precision: split fp16 with power-of-two prescale.  X = x * 2^4 and G = g * 2^13 are each split
// into an fp16 "hi" part (11 significant bits) and an fp16 "lo" part (the next 11 bits; thanks to
// the prescale the lo parts stay in fp16's normal range for |x| >= 2^-6, below that their absolute
// error is 2^-29 in units of x).  Three products are accumulated in ONE fp32 accumulator:
// X_lo*G_hi and X_hi*G_lo first, then X_hi*G_hi with its K steps outside-in (both ends of the band
// towards its centre) so that the large centre taps are added last; the epilogue multiplies by
// 2^-17 (exact).  Every fp16 product is exact in fp32 (11 x 11 bits), so the representation error
// is ~2^-22 relative, as for 3xTF32, at half the tensor work: a K step covers 16 frames instead
// of 8.  The parity bar against the oracle is 1e-6 absolute.  Inputs must satisfy |x| < 4094
// (fp16 range after the prescale); see resampler_b200.h for the non-finite contract.
//
// Data flow of one CTA (persistent, one per SM, 25 warps; every hand-off is an mbarrier):
//   * work item = (run of consecutive tiles) x (group of 128 rows = 128 / CH member streams).
//     A run walks forward in time, so every input frame is fetched from HBM/L2 ONCE per row and
//     kept in a TMEM ring while the ~4 tiles whose windows cover it are computed.
//   * warp 22 (one thread): work scheduler (atomic counter) + TMA producer of the input: one 2-D
//     tensor copy per 16-frame chunk (box = 16 frames x 128/CH members), anchored at the first
//     new input frame (a TMA box must start 16-byte aligned in global memory).
//   * warps 4-15 ("splitter", thread == row; three teams of four warps take every third chunk:
//     one chunk is a long serial latency chain, three in flight hide it):
//     shared memory -> de-interleave -> scale -> fp16 hi / lo pairs -> tcgen05.st into the two
//     TMEM rings (ring column == 2 input frames, TMEM lane == row).  The rings ARE the A operands.
//   * warp 23 (one thread): bulk copies of the tile's [G_hi, G_lo] matrices (prebuilt by
//     tc2_gmat_kernel in the canonical no-swizzle K-major core-matrix layout) into a stage ring.
//   * warps 20 and 21: issue the tcgen05.mma (39+ per 128-tap tile) for alternate tiles, one
//     accumulator each; ONE tcgen05.commit per tile.
//   * warp 24 ("janitor", one thread): follows the tile-completion barriers in order and hands the
//     ring slots no later tile reads back to the splitter (the issuers commit nothing but t_done).
//   * warps 0-3 and 16-19 (epilogue, two teams; team h drains column half h of EVERY tile, so an
//     accumulator is back with the issuers one tcgen05.ld after its tile completes):
//     tcgen05.ld of the warp's 32 accumulator lanes x 32 columns, scale, then two rounds of 16
//     frames each through ONE 128B-swizzled staging buffer per warp and TMA tensor STORES (one
//     warp stores its own members' boxes; no CTA-wide barrier anywhere in the steady state).  The
//     quarter-tile rounds halve the staging memory, which buys the third G stage for 128 taps.
//     The store map clips partial tiles and capacity.
//
// TMEM map (512 columns x 128 lanes): [0,192) X hi ring (384 frames), [192,384) X lo ring,
// [384,448) and [448,512) the two accumulators.
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "fir_kernels.h"
#include "sm100_ptx.cuh"

namespace rsb {

// Optional per-role cycle accounting (one thread per role of CTA 0; enabled by a debug call):
// issuer 0: [0] wait x_full [1] wait g_full [2] wait d_empty [3] issue [4] commit + meta;
// splitter wg 0: [5] wait x_empty [6] wait xs_full [7] loads + split [11] wait::st + arrive
// [12] tcgen05.st [15] loop overhead; epilogue warp 0: [8] wait t_done [9] tcgen05.ld + staging
// [10] TMA store issue + waits; [13] kernel cycles of CTA 0; [14] tiles of CTA 0;
// janitor: [16] wait t_done [17] releases; G producer: [18] wait t_done [19] issue;
// input producer: [23] wait item slot [24] item set-up [25] wait xs_empty [26] TMA issue.
__device__ unsigned long long g_tc2_cycles[32];
__device__ int g_tc2_prof = 0;
// Watchdog of the kernel's mbarrier waits: a wait that has not completed after ~2 s records
// {tag, block, warp, parity | barrier address << 8} in a host-mapped buffer and traps (the launch
// fails with an error instead of hanging the GPU).  Tags: see the kW* constants.
__device__ unsigned int *g_tc2_hang = nullptr;
__device__ int g_tc2_watchdog = 0;

namespace {

using namespace ptx;

enum : uint32_t { kWItemFull = 1, kWItemEmpty, kWXsEmpty, kWXsFull, kWGDone, kWJanDone, kWDEmpty, kWGFull,
                  kWXFull, kWXEmpty, kWEpiDone };
// Suspend-time hint of try_wait: without it a failed attempt returns almost at once and the ~15
// waiting warps of a CTA burn a third of the SM's issue slots re-polling (ncu: 7500 warp
// instructions per tile, 0.5 IPC per scheduler = the whole tile time).  The warp still wakes up as
// soon as the phase completes.
constexpr uint32_t kSuspendHint = 0x989680u;
__constant__ uint32_t g_tc2_hint[2] = {kSuspendHint, kSuspendHint};   // [critical roles, others]
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity, uint32_t hint = kSuspendHint) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint)
        : "memory");
    return ok != 0;
}
// Cold path of a wait (kept out of line: the hot loops stay small).  With the watchdog enabled
// (debug, RSB_TC_WATCHDOG=1) a wait that has not completed after ~2 s records itself and traps.
__device__ __forceinline__ uint32_t wait_hint(uint32_t tag) {
    const bool crit = tag == kWDEmpty || tag == kWGFull || tag == kWXFull || tag == kWEpiDone;
    return g_tc2_hint[crit ? 0 : 1];
}
__device__ __noinline__ void mbar_wait_slow(uint64_t *bar, uint32_t parity, uint32_t tag) {
    unsigned int *rec = g_tc2_hang;
    if (!g_tc2_watchdog || rec == nullptr) {
        const uint32_t hint = wait_hint(tag);
        while (!mbar_try(bar, parity, hint)) { }
        return;
    }
    const long long t0 = clock64();
    for (;;) {
        if (mbar_try(bar, parity)) return;
        if (clock64() - t0 > 4000000000ll) {
            // first CTA to time out claims the record; every stuck warp of that CTA adds its own
            // wait (slot 4 + warp), then the kernel traps half a second later
            const unsigned int prev = atomicCAS(&rec[0], 0u, tag);
            if (prev == 0u) {
                rec[1] = blockIdx.x;
                rec[2] = threadIdx.x >> 5;
                rec[3] = parity | (smem_u32(bar) << 8);
            }
            if (prev == 0u || rec[1] == blockIdx.x)
                rec[4 + (threadIdx.x >> 5)] = tag | (parity << 8) | (smem_u32(bar) << 12);
            __threadfence_system();
            while (clock64() - t0 < 5000000000ll) { }
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait_wd(uint64_t *bar, uint32_t parity, uint32_t tag) {
    if (mbar_try(bar, parity, wait_hint(tag))) return;   // try_wait's already-complete path is the cheapest test
    mbar_wait_slow(bar, parity, tag);
}

struct RoleClock {
    bool on;
    long long t;
    __device__ __forceinline__ void start(bool enable) {
        on = enable;
        t = on ? clock64() : 0;
    }
    __device__ __forceinline__ void lap(int i) {
        if (on) {
            const long long n = clock64();
            atomicAdd(&g_tc2_cycles[i], (unsigned long long)(n - t));
            t = n;
        }
    }
    __device__ __forceinline__ void count(int i, unsigned long long n) {
        if (on) atomicAdd(&g_tc2_cycles[i], n);
    }
};

constexpr uint32_t kRows = 128;                    // MMA M
constexpr uint32_t kN = kTc2TileOut;               // MMA N (64 output frames)
constexpr uint32_t kChunk = 16;                    // frames per input chunk / ring slot / K step
constexpr uint32_t kRing = 384;                    // frames in a TMEM ring
constexpr uint32_t kSlots = kRing / kChunk;        // 24
constexpr uint32_t kSlotCols = kChunk / 2;         // 8 TMEM columns per slot (two fp16 per column)
constexpr uint32_t kColHi = 0, kColLo = kRing / 2, kColD = kRing;
// Role layout (template parameters ST / ET / XS of the kernel): ST splitter teams and ET epilogue
// teams of four warps each, XS landing buffers for input chunks.  XS MUST be a multiple of ST: a
// team then meets every use of "its" stages in turn; otherwise it revisits a stage only every few
// uses and its parity wait can alias when tensor copies land out of order (found by the
// watchdog).  Warps: 0-3 epilogue team 0, then 4 x ST splitter warps, then epilogue team 1 (ET == 2),
// then issuer 0, issuer 1, input TMA producer + scheduler, G producer, janitor.
constexpr uint32_t kMaxXStages = 8;
constexpr uint32_t kXStageBytes = kRows * kChunk * 4;   // 8192 for every channel count
constexpr uint32_t kMaxGStages = 3;
constexpr uint32_t kDone = 8;                      // tile-completion barriers (ring)
__host__ __device__ constexpr uint32_t role_threads(uint32_t st, uint32_t et) { return (4u * (st + et) + 5u) * 32u; }
constexpr uint32_t kItemSlots = 2;
constexpr uint32_t kListEntries = 32;               // main-pass MMA list per tile (128 bytes, after [G_hi, G_lo])
constexpr uint32_t kKsTabBytes = 16;                // per K step: first active output quarter | quarters << 4
constexpr uint32_t kListBytes = kListEntries * 4 + kKsTabBytes;
constexpr uint32_t kListEnd = 0xffffffffu;
constexpr float kScaleX = 16.0f;                   // 2^4
constexpr float kScaleG = 8192.0f;                 // 2^13
constexpr float kScaleOut = 1.0f / (16.0f * 8192.0f);

// tcgen05 instruction descriptor: D = f32, A = B = f16, both K-major, N = 64, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((kN >> 3) << 17) | ((kRows >> 4) << 24);

__host__ __device__ inline uint32_t kt_max_of(uint32_t taps, double ratio) {
    // first and last output of a tile start at most floor(63*ratio)+1 frames apart; up to 15
    // frames of alignment in front; rounded up to the MMA's K step
    const uint32_t span = (uint32_t)((double)(kN - 1) * ratio) + 1u;
    return (15u + span + taps + 15u) & ~15u;
}

struct Item {
    uint32_t t0, t1;       // tiles [t0, t1) of the unit
    uint32_t group;        // member group
    uint32_t n_chunks;     // input chunks of the run
    int32_t vb;            // virtual frame of chunk 0 (H + multiple of kChunk)
    uint32_t valid;
};

// Shared-memory descriptor of a B operand: K-major, no swizzle.  A core matrix is 8 rows x 16
// bytes (8 fp16 along K); the two core matrices of a K step are 64 rows x 16 B = 1024 bytes apart
// (LBO), 8-row groups 128 bytes apart (SBO); descriptor version 1.  The low word holds the start
// address (>> 4) and the LBO, the high word SBO and version.
__device__ __forceinline__ uint32_t b_desc_lo(uint32_t smem_addr) {
    return ((smem_addr & 0x3ffffu) >> 4) | ((1024u >> 4) << 16);
}
constexpr uint32_t kBDescHi = (128u >> 4) | (1u << 14);
constexpr uint32_t kBDescKStep = 2048u >> 4;   // descriptor increment per K step of 16 frames

struct Smem {
    uint64_t xs_full[kMaxXStages], xs_empty[kMaxXStages];
    uint64_t g_full[kMaxGStages];
    uint64_t t_done[kDone];     // tile d's MMAs have completed: barrier d % 8 (tcgen05.commit)
    uint64_t x_full[kSlots], x_empty[kSlots];
    uint64_t d_empty[2];
    uint64_t item_full[kItemSlots], item_empty[kItemSlots];
    Item item[kItemSlots];
    uint32_t tmem_base;
};

// ---- PTX not in sm100_ptx.cuh ----
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo,
                                              uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        ".reg .b64 bd;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_f16_ts_x2(uint32_t d_tmem, uint32_t a0, uint32_t b0_lo,
                                                 uint32_t a1, uint32_t b1_lo, uint32_t b_hi,
                                                 uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t, e;\n\t"
        ".reg .b64 bd0, bd1;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 bd0, {%2, %5};\n\t"
        "mov.b64 bd1, {%4, %5};\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd0, %6, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%3], bd1, %6, t;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a0), "r"(b0_lo), "r"(a1), "r"(b1_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Four MMAs into the same accumulator with one election: the first honours `accumulate`, the
// others always accumulate.
__device__ __forceinline__ void tc_mma_f16_ts_x4(uint32_t d_tmem, uint32_t a0, uint32_t a1,
                                                 uint32_t a2, uint32_t a3, uint32_t b0_lo,
                                                 uint32_t b1_lo, uint32_t b2_lo, uint32_t b3_lo,
                                                 uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t, e;\n\t"
        ".reg .b64 d0, d1, d2, d3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 d0, {%5, %9};\n\t"
        "mov.b64 d1, {%6, %9};\n\t"
        "mov.b64 d2, {%7, %9};\n\t"
        "mov.b64 d3, {%8, %9};\n\t"
        "setp.ne.b32 p, %11, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], d0, %10, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], d1, %10, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%3], d2, %10, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%4], d3, %10, t;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0_lo), "r"(b1_lo), "r"(b2_lo), "r"(b3_lo), "r"(b_hi),
        "r"(idesc), "r"(accumulate)
        : "memory");
}
// Up to four MMAs of the tile's list with one election: each has its own accumulator column range,
// A columns, B descriptor and N (instruction descriptor); MMAs k >= n are predicated off.
__device__ __forceinline__ void tc_mma_f16_ts_list4(uint32_t n, uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3,
                                                    uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                    uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3,
                                                    uint32_t b_hi, uint32_t i0, uint32_t i1, uint32_t i2,
                                                    uint32_t i3) {
    asm volatile(
        "{\n\t"
        ".reg .pred t, e, e1, e2, e3;\n\t"
        ".reg .b64 x0, x1, x2, x3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.gt.u32 e1, %0, 1;\n\t"
        "setp.gt.u32 e2, %0, 2;\n\t"
        "setp.gt.u32 e3, %0, 3;\n\t"
        "and.pred e1, e1, e;\n\t"
        "and.pred e2, e2, e;\n\t"
        "and.pred e3, e3, e;\n\t"
        "mov.b64 x0, {%9, %13};\n\t"
        "mov.b64 x1, {%10, %13};\n\t"
        "mov.b64 x2, {%11, %13};\n\t"
        "mov.b64 x3, {%12, %13};\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], [%5], x0, %14, t;\n\t"
        "@e1 tcgen05.mma.cta_group::1.kind::f16 [%2], [%6], x1, %15, t;\n\t"
        "@e2 tcgen05.mma.cta_group::1.kind::f16 [%3], [%7], x2, %16, t;\n\t"
        "@e3 tcgen05.mma.cta_group::1.kind::f16 [%4], [%8], x3, %17, t;\n\t"
        "}" ::"r"(n),
        "r"(d0), "r"(d1), "r"(d2), "r"(d3), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(b2), "r"(b3),
        "r"(b_hi), "r"(i0), "r"(i1), "r"(i2), "r"(i3)
        : "memory");
}
// As above, but the FIRST MMA honours `acc0` (a tile's very first MMA overwrites its accumulator).
__device__ __forceinline__ void tc_mma_f16_ts_list4a(uint32_t n, uint32_t acc0, uint32_t d0, uint32_t d1, uint32_t d2,
                                                     uint32_t d3, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                     uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t b_hi,
                                                     uint32_t i0, uint32_t i1, uint32_t i2, uint32_t i3) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t, e, e1, e2, e3;\n\t"
        ".reg .b64 x0, x1, x2, x3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.gt.u32 e1, %0, 1;\n\t"
        "setp.gt.u32 e2, %0, 2;\n\t"
        "setp.gt.u32 e3, %0, 3;\n\t"
        "and.pred e1, e1, e;\n\t"
        "and.pred e2, e2, e;\n\t"
        "and.pred e3, e3, e;\n\t"
        "mov.b64 x0, {%9, %13};\n\t"
        "mov.b64 x1, {%10, %13};\n\t"
        "mov.b64 x2, {%11, %13};\n\t"
        "mov.b64 x3, {%12, %13};\n\t"
        "setp.ne.b32 p, %18, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], [%5], x0, %14, p;\n\t"
        "@e1 tcgen05.mma.cta_group::1.kind::f16 [%2], [%6], x1, %15, t;\n\t"
        "@e2 tcgen05.mma.cta_group::1.kind::f16 [%3], [%7], x2, %16, t;\n\t"
        "@e3 tcgen05.mma.cta_group::1.kind::f16 [%4], [%8], x3, %17, t;\n\t"
        "}" ::"r"(n),
        "r"(d0), "r"(d1), "r"(d2), "r"(d3), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(b2), "r"(b3),
        "r"(b_hi), "r"(i0), "r"(i1), "r"(i2), "r"(i3), "r"(acc0)
        : "memory");
}
// TMA tensor store shared -> global of one 2-D box (SASS: UTMASTG), bulk-group completion
__device__ __forceinline__ void tensor_s2g_2d(const CUtensorMap *tm, int c0, int c1, uint32_t smem_addr) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm),
                 "r"(c0), "r"(c1), "r"(smem_addr)
                 : "memory");
}
// L2 prefetch of one 2-D box (no shared memory involved): hides the HBM latency of the input
// stream behind a handful of landing buffers
__device__ __forceinline__ void tensor_prefetch_2d(const CUtensorMap *tm, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// (a, b) = two consecutive frames of one row -> packed fp16 hi pair and lo pair; the even frame
// sits in the low half word (element 2c of TMEM column c)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t &hi, uint32_t &lo) {
    const float A = a * kScaleX, B = b * kScaleX;
    const __half2 h = __floats2half2_rn(A, B);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(__fsub_rn(A, hf.x), __fsub_rn(B, hf.y));
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// One input chunk (16 frames) of this thread's row out of the TMA landing buffer.
// RAW: 0 = f32 input; 1 = raw s16 / s24 frames of CH channels (CH = 1 or 2); 2 = raw MONO frames
// duplicated into both channels of a stereo stream (the CLI's mono -> stereo, main.rs:139-146).
// SB: bytes per raw sample, 2 (s16) or 3 (packed little-endian s24).
template <int CH, int RAW, int SB>
__device__ __forceinline__ void load_chunk(uint32_t base, const float *stage_ptr, uint32_t ml, uint32_t c,
                                           float (&x)[kChunk]) {
    constexpr bool kRawStereo = RAW == 1 && CH == 2;
    constexpr bool kRawMono = (RAW == 1 && CH == 1) || RAW == 2;
    if constexpr (SB == 3 && kRawStereo) {
        // packed s24 stereo: unswizzled 96-byte rows; sample k of channel c sits at byte 6k + 3c
        const uint32_t rb = base + ml * 96u;
        uint32_t w[25];
#pragma unroll
        for (uint32_t u = 0; u < 6; ++u) {
            const float4 q4 = lds128(rb + 16u * u);
            w[4 * u] = __float_as_uint(q4.x); w[4 * u + 1] = __float_as_uint(q4.y);
            w[4 * u + 2] = __float_as_uint(q4.z); w[4 * u + 3] = __float_as_uint(q4.w);
        }
        w[24] = 0;
#pragma unroll
        for (uint32_t k = 0; k < kChunk; ++k) {
            const uint32_t o0 = 6 * k, o1 = 6 * k + 3;
            const uint32_t v0 = __funnelshift_r(w[o0 >> 2], w[(o0 >> 2) + 1], (o0 & 3u) * 8u);
            const uint32_t v1 = __funnelshift_r(w[o1 >> 2], w[(o1 >> 2) + 1], (o1 & 3u) * 8u);
            const uint32_t vv = c ? v1 : v0;
            x[k] = (float)((int)(vv << 8) >> 8) * (1.0f / 8388608.0f);
        }
    } else if constexpr (SB == 3 && kRawMono) {
        // packed s24 mono: unswizzled 48-byte rows
        const uint32_t rb = base + ml * 48u;
        uint32_t w[13];
#pragma unroll
        for (uint32_t u = 0; u < 3; ++u) {
            const float4 q4 = lds128(rb + 16u * u);
            w[4 * u] = __float_as_uint(q4.x); w[4 * u + 1] = __float_as_uint(q4.y);
            w[4 * u + 2] = __float_as_uint(q4.z); w[4 * u + 3] = __float_as_uint(q4.w);
        }
        w[12] = 0;
#pragma unroll
        for (uint32_t k = 0; k < kChunk; ++k) {
            const uint32_t o = 3 * k;
            const uint32_t vv = __funnelshift_r(w[o >> 2], w[(o >> 2) + 1], (o & 3u) * 8u);
            x[k] = (float)((int)(vv << 8) >> 8) * (1.0f / 8388608.0f);
        }
    } else if constexpr (kRawMono) {
        // raw s16 mono: unswizzled 32-byte rows (16 frames), one per member; both channel rows of
        // a stereo member read the same samples (RAW == 2)
        const uint32_t rb = base + ml * 32u;
#pragma unroll
        for (uint32_t u = 0; u < 2; ++u) {
            const float4 q4 = lds128(rb + 16u * u);
            const uint32_t w[4] = {__float_as_uint(q4.x), __float_as_uint(q4.y), __float_as_uint(q4.z),
                                   __float_as_uint(q4.w)};
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                x[8 * u + 2 * k] = (float)(int)(short)(w[k] & 0xffffu) * (1.0f / 32768.0f);
                x[8 * u + 2 * k + 1] = (float)((int)w[k] >> 16) * (1.0f / 32768.0f);
            }
        }
    } else if constexpr (kRawStereo) {
        // raw s16 stereo: 64-byte rows (16 frames of 2 x s16), 64B swizzle: unit u at
        // u ^ ((row >> 1) & 3); a word is one frame, the channel picks its half;
        // s / 2^15 is the reference's `s as f32 / 32768.0` (main.rs:131-136)
        const uint32_t rb = base + ml * 64u;
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) {
            const float4 q4 = lds128(rb + ((u ^ ((ml >> 1) & 3u)) << 4));
            const uint32_t w[4] = {__float_as_uint(q4.x), __float_as_uint(q4.y), __float_as_uint(q4.z),
                                   __float_as_uint(q4.w)};
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                const int sv = c ? (int)w[k] >> 16 : (int)(short)(w[k] & 0xffffu);
                x[4 * u + k] = (float)sv * (1.0f / 32768.0f);
            }
        }
    } else if constexpr (CH == 2) {
        // 128-byte rows (16 stereo frames), 128B swizzle: unit u at u ^ (row & 7); the two channel
        // rows of a member read the same addresses (a broadcast inside the warp)
        const uint32_t rb = base + ml * 128u;
#pragma unroll
        for (uint32_t u = 0; u < 8; ++u) {
            const float4 q4 = lds128(rb + ((u ^ (ml & 7u)) << 4));
            x[2 * u] = c ? q4.y : q4.x;
            x[2 * u + 1] = c ? q4.w : q4.z;
        }
    } else if constexpr (CH == 1) {
        // 64-byte rows (16 mono frames), 64B swizzle: unit u at u ^ ((row >> 1) & 3)
        const uint32_t rb = base + ml * 64u;
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) {
            const float4 q4 = lds128(rb + ((u ^ ((ml >> 1) & 3u)) << 4));
            x[4 * u] = q4.x; x[4 * u + 1] = q4.y; x[4 * u + 2] = q4.z; x[4 * u + 3] = q4.w;
        }
    } else {
        // 4 / 8 channels: rows of 16 frames x CH floats, no swizzle; one 4-byte load per frame
        const float *rowp = stage_ptr + ml * (kChunk * CH) + c;
#pragma unroll
        for (uint32_t f = 0; f < kChunk; ++f) x[f] = rowp[f * CH];
    }
}

template <int CH, int RAW = 0, int SB = 2, int ST = 3, int ET = 2, int XS = 6, int QS = 1>
__global__ void __launch_bounds__(role_threads(ST, ET), 1)
conv_tc2_kernel(const __grid_constant__ Tc2Params P, const __grid_constant__ CUtensorMap tmap_in,
                const __grid_constant__ CUtensorMap tmap_out) {
    static_assert(CH == 1 || CH == 2 || CH == 4 || CH == 8, "tensor kernel: 1, 2, 4 or 8 channels");
    static_assert(RAW == 0 || (RAW == 1 && CH <= 2) || (RAW == 2 && CH == 2), "raw input: mono / stereo");
    static_assert(SB == 2 || (SB == 3 && RAW != 0), "raw samples: 16 or packed 24 bits");
    constexpr uint32_t kSplitTeams = ST, kEpiTeams = ET, kXStages = XS;
    static_assert(kXStages % kSplitTeams == 0 && kSlots % kSplitTeams == 0 && kXStages <= kMaxXStages,
                  "stage / slot ownership must be fixed per team");
    static_assert(ET == 1 || ET == 2, "one or two epilogue teams");
    constexpr uint32_t kWarpEpi1 = 4 + 4 * kSplitTeams;                 // first warp of epilogue team 1
    constexpr uint32_t kWarpIssuer0 = 4 + 4 * (kSplitTeams + kEpiTeams - 1), kWarpIssuer1 = kWarpIssuer0 + 1,
                       kWarpTmaIn = kWarpIssuer0 + 2, kWarpG = kWarpIssuer0 + 3, kWarpJanitor = kWarpIssuer0 + 4;
    constexpr uint32_t kItemConsumers = 4 * (kSplitTeams + kEpiTeams) + 4;   // every warp but the scheduler
    constexpr uint32_t kStageBufs = kEpiTeams == 1 ? 2u : 1u;               // staging buffers per epilogue warp
    constexpr bool kRawStereo = RAW == 1 && CH == 2;
    constexpr uint32_t kRawFrameBytes = RAW == 0 ? 0u : (kRawStereo ? 2u : 1u) * SB;
    // bytes one input chunk lands in shared memory: f32 rows, or one row of 16 raw frames per member
    constexpr uint32_t kXLandBytes = RAW != 0 ? (kRows / CH) * kChunk * kRawFrameBytes : kXStageBytes;
    constexpr uint32_t kMpg = kRows / CH;               // members per group
    constexpr uint32_t kMpw = 32 / CH;                  // members per epilogue warp
    // epilogue staging: a half tile (32 frames) of one warp's members = CH boxes of kMpw rows x 128
    // bytes, 128B-swizzled, every box 1024-byte aligned
    constexpr uint32_t kBoxBytes = kMpw * 128u < 1024u ? 1024u : kMpw * 128u;
    constexpr uint32_t kHalfBytes = CH * kBoxBytes;
    // QS: the epilogue stages a QUARTER tile (16 frames) per round in one buffer per warp: half the
    // staging memory, which buys a third G stage for the 128-tap configurations
    constexpr bool kQuarter = QS != 0 && CH >= 2 && ET == 2;
    constexpr uint32_t kWarpStageBytes = kQuarter ? (CH / 2) * kBoxBytes : kHalfBytes;   // x kStageBufs (ET == 1)
    extern __shared__ __align__(1024) uint8_t smem_tc2[];
    __shared__ Smem S;
    const uint32_t g_bytes = P.kt_max * 128u;           // one G half: kt_max/8 K groups x 1024 B
    uint8_t *xst = smem_tc2;                             // [kXStages][kXStageBytes]
    uint8_t *ost = smem_tc2 + kXStages * kXStageBytes;   // [kEpiTeams][4 quadrants][kStageBufs][kHalfBytes]
    uint8_t *gst = ost + 2 * 4 * kWarpStageBytes;        // [g_stages]{[2][g_bytes], MMA list}
    const uint32_t g_stage_bytes = 2u * g_bytes + kListBytes;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const bool prof = g_tc2_prof != 0 && blockIdx.x == 0;
    const long long t_kernel = prof ? clock64() : 0;
    RoleClock rc;
    rc.start(false);
    if (tid == 0) {
        for (uint32_t i = 0; i < kXStages; ++i) { mbar_init(&S.xs_full[i], 1); mbar_init(&S.xs_empty[i], 4); }
        for (uint32_t i = 0; i < kMaxGStages; ++i) mbar_init(&S.g_full[i], 1);
        for (uint32_t i = 0; i < kDone; ++i) mbar_init(&S.t_done[i], 1);
        for (uint32_t i = 0; i < kSlots; ++i) { mbar_init(&S.x_full[i], 4); mbar_init(&S.x_empty[i], 1); }
        // every epilogue warp that drains the accumulator + janitor
        for (uint32_t i = 0; i < 2; ++i) mbar_init(&S.d_empty[i], kEpiTeams == 2 && P.epi_split ? 9 : 5);
        for (uint32_t i = 0; i < kItemSlots; ++i) {
            mbar_init(&S.item_full[i], 1);
            mbar_init(&S.item_empty[i], kItemConsumers);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpIssuer0) tmem_alloc(&S.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    const UnitDev &U = P.units[0];
    const int32_t H = (int32_t)U.hist_len0;
    const uint32_t n_tiles = (uint32_t)((U.total_out + kN - 1) / kN);
    const Tc2Tile *tct = P.tct;
    const uint32_t n_runs = (n_tiles + P.run_tiles - 1) / P.run_tiles;
    const uint32_t n_items = n_runs * P.groups;
    const uint32_t n_gst = P.g_stages;

    // every consumer warp fetches the next item from the scheduler's two-slot queue (one arrival
    // per warp)
    auto get_item = [&](uint32_t it) {
        const uint32_t slot = it % kItemSlots;
        mbar_wait_wd(&S.item_full[slot], (it / kItemSlots) & 1u, kWItemFull);
        const Item I = S.item[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.item_empty[slot]);
        return I;
    };

    if (warp == kWarpTmaIn) {
        // ===== scheduler + TMA producer of the input chunks =====
        if (lane == 0) {
            uint32_t xs_seq = 0;
            rc.start(prof);
            for (uint32_t it = 0;; ++it) {
                const uint32_t slot = it % kItemSlots;
                mbar_wait_wd(&S.item_empty[slot], ((it / kItemSlots) & 1u) ^ 1u, kWItemEmpty);
                rc.lap(23);
                const uint32_t idx = atomicAdd(P.work_counter, 1u);
                Item I;
                I.valid = idx < n_items ? 1u : 0u;
                I.t0 = I.t1 = I.group = I.n_chunks = 0;
                I.vb = 0;
                if (I.valid) {
                    const uint32_t r = idx / P.groups;
                    I.group = idx - r * P.groups;
                    I.t0 = r * P.run_tiles;
                    I.t1 = min(I.t0 + P.run_tiles, n_tiles);
                    const Tc2Tile first = tct[I.t0], last = tct[I.t1 - 1];
                    // tile K ranges start on the chunk grid anchored at the first frame of the new
                    // input (virtual frame H): a TMA box must start 16-byte aligned in global
                    // memory, and no chunk straddles the history / input seam
                    I.vb = first.k0;
                    I.n_chunks = (uint32_t)(last.k0 + (int32_t)last.kt - I.vb) / kChunk;
                }
                S.item[slot] = I;
                mbar_arrive(&S.item_full[slot]);
                if (!I.valid) break;
                const int32_t m0 = (int32_t)(I.group * kMpg);
                // inner coordinate in tensor-map elements: frames (mono f32, stereo 8-byte frames,
                // raw frames), bytes (packed s24) or floats (4 / 8 channels)
                constexpr int32_t kCoordMul = (int32_t)(SB == 3 ? kRawFrameBytes : CH >= 4 ? CH : 1);
                const uint32_t pf = P.prefetch_chunks;      // 0: no L2 prefetch
                for (uint32_t j = 0; j < min(pf, I.n_chunks); ++j) {
                    const int32_t v = I.vb + (int32_t)(j * kChunk);
                    if (v >= H) tensor_prefetch_2d(&tmap_in, (v - H) * kCoordMul, m0);
                }
                rc.lap(24);
                for (uint32_t j = 0; j < I.n_chunks; ++j) {
                    const int32_t v = I.vb + (int32_t)(j * kChunk);
                    if (pf != 0 && j + pf < I.n_chunks) {
                        const int32_t vp = v + (int32_t)(pf * kChunk);
                        if (vp >= H) tensor_prefetch_2d(&tmap_in, (vp - H) * kCoordMul, m0);
                    }
                    if (v < H) continue;       // touches the history: the splitter loads it itself
                    const uint32_t s = xs_seq % kXStages;
                    mbar_wait_wd(&S.xs_empty[s], ((xs_seq / kXStages) & 1u) ^ 1u, kWXsEmpty);
                    rc.lap(25);
                    if (P.ablate & 2u) {
                        mbar_arrive(&S.xs_full[s]);
                    } else {
                        mbar_arrive_expect_tx(&S.xs_full[s], kXLandBytes);
                        tensor_g2s_2d(xst + s * kXStageBytes, &tmap_in, (v - H) * kCoordMul, m0, &S.xs_full[s]);
                    }
                    rc.lap(26);
                    ++xs_seq;
                }
            }
        }
        __syncwarp();
    } else if (warp == kWarpG) {
        // ===== TMA producer of the G matrices =====
        uint32_t g_seq = 0;
        rc.start(prof && lane == 0);
        for (uint32_t it = 0;; ++it) {
            const Item I = get_item(it);
            if (!I.valid) break;
            if (lane == 0) {
                for (uint32_t t = I.t0; t < I.t1; ++t) {
                    const Tc2Tile m = tct[t];
                    const uint32_t s = g_seq % n_gst;
                    if (g_seq >= n_gst) {     // the stage's previous tile has been multiplied
                        const uint32_t prev = g_seq - n_gst;
                        mbar_wait_wd(&S.t_done[prev % kDone], (prev / kDone) & 1u, kWGDone);
                    }
                    rc.lap(18);
                    const uint32_t bytes = m.kt * 128u;
                    const uint8_t *src = P.gmat + (size_t)m.g_idx * g_stage_bytes;
                    uint8_t *dst = gst + (size_t)s * g_stage_bytes;
                    if ((P.ablate & 1u) && g_seq >= n_gst) {
                        mbar_arrive(&S.g_full[s]);
                    } else {
                        mbar_arrive_expect_tx(&S.g_full[s], 2 * bytes + kListBytes);
                        bulk_g2s(dst, src, bytes, &S.g_full[s]);
                        bulk_g2s(dst + g_bytes, src + g_bytes, bytes, &S.g_full[s]);
                        bulk_g2s(dst + 2 * g_bytes, src + 2 * g_bytes, kListBytes, &S.g_full[s]);
                    }
                    ++g_seq;
                    rc.lap(19);
                }
            }
            g_seq = __shfl_sync(0xffffffffu, g_seq, 0);
        }
    } else if (warp == kWarpJanitor) {
        // ===== janitor: hands ring slots back to the splitter once no later tile reads them =====
        uint32_t d_seq = 0;          // tiles of this CTA so far
        uint32_t r_slot = 0;         // ring slot of the next chunk to release
        rc.start(prof && lane == 0);
        for (uint32_t it = 0;; ++it) {
            const Item I = get_item(it);
            if (!I.valid) break;
            if (lane == 0) {
                uint32_t rel = 0;    // chunks of this run released so far
                Tc2Tile mn = tct[I.t0];
                for (uint32_t t = I.t0; t < I.t1; ++t, ++d_seq) {
                    const bool last = t + 1 >= I.t1;
                    if (!last) mn = tct[t + 1];
                    mbar_wait_wd(&S.t_done[d_seq % kDone], (d_seq / kDone) & 1u, kWJanDone);
                    rc.lap(16);
                    // chunks that end at or before the next tile's first frame are free again
                    const uint32_t upto = last ? I.n_chunks : (uint32_t)(mn.k0 - I.vb) / kChunk;
                    for (; rel < upto; ++rel) {
                        mbar_arrive(&S.x_empty[r_slot]);
                        if (++r_slot == kSlots) r_slot = 0;
                    }
                    // keeps the issuers at most two tiles ahead of this warp (the t_done ring)
                    mbar_arrive(&S.d_empty[d_seq & 1u]);
                    rc.lap(17);
                }
            }
            d_seq = __shfl_sync(0xffffffffu, d_seq, 0);
            r_slot = __shfl_sync(0xffffffffu, r_slot, 0);
        }
    } else if (warp == kWarpIssuer0 || warp == kWarpIssuer1) {
        // ===== MMA issuers.  The whole warp runs the loop (warp-uniform values), one elected lane
        // issues.  Two warps issue alternate tiles into the two accumulators whenever two
        // consecutive tiles' K ranges fit the ring together (P.issuers == 2); otherwise the first
        // issues every tile. =====
        const uint32_t mine = warp == kWarpIssuer0 ? 0u : 1u;
        const uint32_t step = P.issuers;
        uint32_t base_slot = 0, base_par = 0;   // ring slot / phase parity of the run's chunk 0
        uint32_t d_seq = 0;                     // tiles of this CTA so far; tile -> accumulator d_seq & 1
        rc.start(prof && lane == 0 && mine == 0);
        for (uint32_t it = 0;; ++it) {
            const Item I = get_item(it);
            if (!I.valid) break;
            if (mine < step) {
                rc.count(14, I.t1 - I.t0);
                uint32_t waited = 0;                      // chunks of this run already seen full
                const uint32_t d_run = d_seq;
                uint32_t t = I.t0 + (step == 2 ? ((d_seq ^ mine) & 1u) : 0u);
                d_seq += t - I.t0;
                Tc2Tile m = {0, 0, 0, 0, 0, 0, 0, 0};
                if (t < I.t1) m = tct[t];
                for (; t < I.t1; t += step, d_seq += step) {
                    Tc2Tile mn = m;
                    if (t + step < I.t1) mn = tct[t + step];      // prefetch the next own tile
                    const uint32_t j0 = (uint32_t)(m.k0 - I.vb) / kChunk;
                    const uint32_t n_ks = m.kt / kChunk;
                    rc.lap(4);
                    // The accumulator wait comes FIRST: it passes only after the epilogue has
                    // drained tile d_seq - 2, i.e. after every earlier user of this tile's G
                    // stage has completed, so the parity wait on g_full below cannot be looking
                    // at the stage's previous phase.
                    const uint32_t b = d_seq & 1u;
                    mbar_wait_wd(&S.d_empty[b], ((d_seq >> 1) & 1u) ^ 1u, kWDEmpty);
                    rc.lap(2);
                    const uint32_t gs = d_seq % n_gst;
                    mbar_wait_wd(&S.g_full[gs], (d_seq / n_gst) & 1u, kWGFull);
                    rc.lap(1);
                    // Only the tile's OWN chunks [j0, j0 + n_ks) are waited for: they cannot be
                    // recycled before this tile completes.  (Earlier chunks may already have been
                    // released by the janitor and refilled for the ring's next revolution while
                    // this warp was held up elsewhere; a parity wait on them would alias and never
                    // complete -- found by the watchdog with a slow epilogue.)
                    // A splitter team fills its chunks in order, so the LAST chunk of each team
                    // inside the range vouches for the team's earlier ones: at most kSplitTeams
                    // waits per tile instead of one per chunk (an already-complete try_wait still
                    // costs ~90 cycles of this warp's critical path).
                    {
                        const uint32_t lo_c = waited > j0 ? waited : j0, hi_c = j0 + n_ks;
                        uint32_t cw = hi_c - lo_c > kSplitTeams ? hi_c - kSplitTeams : lo_c;
                        const uint32_t lin = base_slot + cw;
                        uint32_t w_slot = lin % kSlots, w_par = base_par ^ ((lin / kSlots) & 1u);
                        for (; cw < hi_c; ++cw) {
                            mbar_wait_wd(&S.x_full[w_slot], w_par, kWXFull);
                            if (++w_slot == kSlots) { w_slot = 0; w_par ^= 1u; }
                        }
                        waited = hi_c;
                    }
                    rc.lap(0);
                    tc_fence_after();

                    const uint32_t d_tmem = tmem + kColD + b * kN;
                    const uint32_t s0 = (base_slot + j0) % kSlots;
                    const uint32_t ghi = b_desc_lo(smem_u32(gst + (size_t)gs * g_stage_bytes));
                    const uint32_t glo = ghi + (g_bytes >> 4);
                    const uint32_t ahi = tmem + kColHi, alo = tmem + kColLo;
                    constexpr uint32_t kWrapCols = kSlots * kSlotCols;
                    auto wrap = [](uint32_t c) { return c >= kWrapCols ? c - kWrapCols : c; };
                    // small terms first: X_lo * G_hi and X_hi * G_lo, two K steps per issue block.  A K
                    // step's band covers only some of the tile's four output quarters (the tile's
                    // K-step table: first active quarter | quarters << 4): the MMA's N, accumulator
                    // columns and G rows are trimmed to them -- the skipped blocks of G are zero, and
                    // the tensor pipe is the kernel's largest energy consumer.  The tile's very
                    // first pair runs at full width: it initialises the accumulator.
                    const uint32_t col0 = s0 * kSlotCols;
                    uint32_t col = col0, dk = 0, ks = 0;
                    const uint32_t n_ks_issue = (P.ablate & 32u) ? 0u : n_ks;
                    uint32_t tw[4];
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(tw[0]), "=r"(tw[1]), "=r"(tw[2]), "=r"(tw[3])
                                 : "r"(smem_u32(gst + (size_t)gs * g_stage_bytes + 2u * g_bytes) + kListEntries * 4u));
                    const uint64_t tab_lo = (uint64_t)tw[0] | ((uint64_t)tw[1] << 32);
                    const uint64_t tab_hi = (uint64_t)tw[2] | ((uint64_t)tw[3] << 32);
                    auto ks_entry = [&](uint32_t k) -> uint32_t {
                        const uint32_t b = (uint32_t)((k < 8u ? tab_lo >> (8u * k) : tab_hi >> (8u * (k - 8u))) & 0xffu);
                        return (k == 0u || (P.ablate & 64u)) ? 0x40u : b;        // K step 0: full width
                    };
                    constexpr uint32_t kIdescNoN = kIdesc & ~(0x3fu << 17);
#pragma unroll 2
                    for (; ks < n_ks_issue; ks += 2) {
                        const bool two = ks + 1 < n_ks_issue;
                        const uint32_t c1 = wrap(col + kSlotCols);
                        const uint32_t e0 = ks_entry(ks), e1 = two ? ks_entry(ks + 1) : 0x40u;
                        const uint32_t q0a = (e0 & 0xfu) * 16u, ia = kIdescNoN | (((e0 >> 4) * 2u) << 17);
                        const uint32_t q0b = (e1 & 0xfu) * 16u, ib = kIdescNoN | (((e1 >> 4) * 2u) << 17);
                        tc_mma_f16_ts_list4a(two ? 4u : 2u, ks != 0, d_tmem + q0a, d_tmem + q0a, d_tmem + q0b, d_tmem + q0b,
                                             alo + col, ahi + col, alo + c1, ahi + c1, ghi + dk + q0a, glo + dk + q0a,
                                             ghi + dk + kBDescKStep + q0b, glo + dk + kBDescKStep + q0b, kBDescHi, ia, ia,
                                             ib, ib);
                        dk += 2 * kBDescKStep;
                        col = wrap(c1 + kSlotCols);
                    }
                    // X_hi * G_hi: the tile's MMA list (built with the G matrices, same stage).  Every
                    // K step appears once per output quarter: "early" (outside-in) for the quarters
                    // whose main lobe lies elsewhere, "late" for those it carries, as MMAs over
                    // column sub-ranges (entry = K step | first quarter << 8 | quarters << 12).  The
                    // tensor core's fp32 accumulation truncates; adding an output's main lobe last
                    // keeps the number of truncations at full magnitude to one or two.
                    if (P.ablate & 32u) {
                        tc_mma_f16_ts(d_tmem, ahi + col0, ghi, kBDescHi, kIdesc, 1u);
                    } else {
                        const uint32_t lst = smem_u32(gst + (size_t)gs * g_stage_bytes + 2u * g_bytes);
#pragma unroll 1
                        for (uint32_t e = 0; e < kListEntries; e += 4) {
                            uint32_t w[4];
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                                         : "r"(lst + e * 4u));
                            if (w[0] == kListEnd) break;
                            uint32_t dd[4], aa[4], bb[4], ii[4], nv = 0;
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) {
                                const uint32_t ksx = w[k] & 0xffu, q0 = (w[k] >> 8) & 0xfu, nq = (w[k] >> 12) & 0xfu;
                                uint32_t sl = s0 + ksx;
                                sl = sl >= kSlots ? sl - kSlots : sl;
                                dd[k] = d_tmem + q0 * 16u;
                                aa[k] = ahi + sl * kSlotCols;
                                bb[k] = ghi + ksx * kBDescKStep + q0 * 16u;
                                ii[k] = (kIdesc & ~(0x3fu << 17)) | ((nq * 2u) << 17);   // N = 16 * nq
                                nv += w[k] != kListEnd ? 1u : 0u;
                            }
                            tc_mma_f16_ts_list4(nv, dd[0], dd[1], dd[2], dd[3], aa[0], aa[1], aa[2], aa[3], bb[0], bb[1],
                                                bb[2], bb[3], kBDescHi, ii[0], ii[1], ii[2], ii[3]);
                            if (nv < 4u) break;
                        }
                    }
                    rc.lap(3);
                    // the tile's only commit: epilogue (accumulator ready), G producer (stage
                    // free) and janitor (ring slots free) all follow this barrier
                    tc_commit_elect(&S.t_done[d_seq % kDone]);
                    m = mn;
                }
                d_seq = d_run + (I.t1 - I.t0);
            } else {
                d_seq += I.t1 - I.t0;
            }
            // ring position of the next run's chunk 0
            base_slot += I.n_chunks % kSlots;
            base_par ^= (I.n_chunks / kSlots) & 1u;
            if (base_slot >= kSlots) { base_slot -= kSlots; base_par ^= 1u; }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 4 + 4 * kSplitTeams) {
        // ===== splitter: shared memory (TMA landing buffer) -> fp16 hi / lo rings in TMEM.  Three
        // teams of four warps (one per TMEM lane quadrant) take every third chunk. =====
        const uint32_t wg = (warp - 4u) >> 2;
        const uint32_t row = tid & 127u;                    // TMEM lane
        const uint32_t ml = row / CH, c = row % CH;         // member inside the group, channel
        const uint32_t lane_base = ((warp & 3u) * 32u) << 16;
        uint32_t q_seq = 0;             // chunks of this CTA so far (both warpgroups count all)
        uint32_t xs_seq = 0;            // TMA-landed chunks so far (both warpgroups count all)
        uint32_t rs = 0, par = 1;       // ring slot of chunk q_seq and the parity its x_empty wait uses
        rc.start(prof && row == 0 && wg == 0);
        for (uint32_t it = 0;; ++it) {
            const Item I = get_item(it);
            if (!I.valid) break;
            const uint32_t mg = I.group * kMpg + ml;
            const bool member_ok = mg < U.n_members;
            const float *hist = nullptr, *in = nullptr;
            if (I.vb < H && member_ok) {
                const JobDev *job = P.jobs + U.member_off + mg;
                hist = job->hist;
                in = job->in;
            }
            for (uint32_t j = 0; j < I.n_chunks; ++j, ++q_seq) {
                const int32_t v = I.vb + (int32_t)(j * kChunk);
                const bool landed = v >= H;
                if ((q_seq % kSplitTeams) == wg) {
                    rc.lap(15);
                    // Load and split BEFORE waiting for the ring slot: with the values already
                    // split in registers only the TMEM stores remain between "slot free" and
                    // "chunk readable".
                    float x[kChunk];
                    int fast_slot = -1;
                    if (P.ablate & 8u) {
#pragma unroll
                        for (uint32_t f = 0; f < kChunk; ++f) x[f] = 0.f;
                        if (landed) {
                            const uint32_t s = xs_seq % kXStages;
                            mbar_wait_wd(&S.xs_full[s], (xs_seq / kXStages) & 1u, kWXsFull);
                            __syncwarp();
                            fast_slot = (int)s;
                        }
                    } else if (landed) {
                        const uint32_t s = xs_seq % kXStages;
                        mbar_wait_wd(&S.xs_full[s], (xs_seq / kXStages) & 1u, kWXsFull);
                        __syncwarp();
                        rc.lap(6);
                        load_chunk<CH, RAW, SB>(smem_u32(xst + s * kXStageBytes),
                                                reinterpret_cast<const float *>(xst + s * kXStageBytes), ml, c, x);
                        fast_slot = (int)s;
                    } else {
                        // the chunk lies in the history (only at the very start of a batch)
#pragma unroll
                        for (uint32_t f = 0; f < kChunk; ++f) {
                            const int64_t vv = (int64_t)v + f;
                            float xv = 0.f;
                            if (member_ok && vv >= 0) {
                                if (vv < H) xv = hist[((int64_t)kHistFrames - H + vv) * CH + c];
                                else if ((uint64_t)(vv - H) < U.total_frames) xv = in[(vv - H) * CH + c];
                            }
                            x[f] = xv;
                        }
                    }
                    uint32_t hi[kSlotCols], lo[kSlotCols];
#pragma unroll
                    for (uint32_t f = 0; f < kSlotCols; ++f) split_pair(x[2 * f], x[2 * f + 1], hi[f], lo[f]);
                    rc.lap(7);
                    // The landing buffer is handed back only after the loaded values have been
                    // consumed: the empty asm below takes every split value as an input, so the
                    // loads and the arithmetic are complete before the arrive.
                    asm volatile("" ::"r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]),
                                 "r"(hi[6]), "r"(hi[7]), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]),
                                 "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7])
                                 : "memory");
                    if (fast_slot >= 0) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&S.xs_empty[fast_slot]);
                    }
                    mbar_wait_wd(&S.x_empty[rs], par, kWXEmpty);
                    __syncwarp();
                    rc.lap(5);
                    tc_fence_after();
                    const uint32_t colw = rs * kSlotCols;
                    if (!(P.ablate & 8u)) {
                        tmem_st8(tmem + lane_base + kColHi + colw, hi);
                        tmem_st8(tmem + lane_base + kColLo + colw, lo);
                    }
                    rc.lap(12);
                    // the stores must have completed before hi[] / lo[] are overwritten by the
                    // next chunk (tcgen05.st reads its source registers asynchronously) and
                    // before the issuers are told
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.x_full[rs]);
                    rc.lap(11);
                }
                if (landed) ++xs_seq;
                if (++rs == kSlots) { rs = 0; par ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: accumulator (TMEM) -> scale -> swizzled staging -> TMA tensor stores.
        // A warp owns the 32 accumulator lanes of its quadrant = kMpw members and stores their
        // boxes itself (no barrier wider than a warp). =====
        // Two teams: team h drains column half h of EVERY tile, so an accumulator is back with the
        // issuers one tcgen05.ld after its tile completes (alternating whole tiles between the teams
        // held it through the first half's staging and stores: the accumulator turnaround was the
        // kernel's critical loop).  One team: both halves in turn.
        const uint32_t team = warp >= kWarpEpi1 ? 1u : 0u;
        const bool epi_split = kEpiTeams == 2 && P.epi_split != 0;
        const uint32_t quad = warp & 3u;
        const uint32_t ml = lane / CH, c = lane % CH;       // member inside the warp, channel
        const uint32_t lane_base = (quad * 32u) << 16;
        const uint32_t sb0 = smem_u32(ost + (team * 4u + quad) * (kQuarter ? kWarpStageBytes : kStageBufs * kHalfBytes));
        uint32_t h_seq = 0;
        uint32_t d_seq = 0;
        const float out_scale = P.out_scale;
        rc.start(prof && tid == 0);
        for (uint32_t it = 0;; ++it) {
            const Item I = get_item(it);
            if (!I.valid) break;
            const int32_t m_first = (int32_t)(I.group * kMpg + quad * kMpw);
            for (uint32_t t = I.t0; t < I.t1; ++t, ++d_seq) {
                if (kEpiTeams == 2 && !epi_split && (d_seq & 1u) != team) continue;
                const uint32_t o_start = t * kN;
                const uint32_t b = d_seq & 1u;
                rc.lap(10);
                mbar_wait_wd(&S.t_done[d_seq % kDone], (d_seq / kDone) & 1u, kWEpiDone);
                __syncwarp();     // lanes leave the polling loop at different times
                rc.lap(8);
                tc_fence_after();
                if (P.ablate & 16u) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.d_empty[b]);
                    continue;
                }
#pragma unroll 1
                for (uint32_t hf = (epi_split ? team : 0u); hf < (epi_split ? team + 1u : 2u); ++hf) {
                    uint32_t acc[32];
                    tmem_ld32(tmem + lane_base + kColD + b * kN + hf * 32u, acc);
                    tmem_wait_ld();
                    if (epi_split || hf == 1) {   // this warp's part is drained: hand it back to the issuers
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&S.d_empty[b]);
                    }
                    rc.lap(9);
                    if constexpr (kQuarter) {
#pragma unroll
                        for (uint32_t qr = 0; qr < 2; ++qr) {
                            // the staging buffer's previous store has read it
                            if (lane == 0) bulk_wait_read<0>();
                            __syncwarp();
                            rc.lap(20);
                            if constexpr (CH == 2) {
                                // lane pair (member, L) / (member, R): the L lane writes frames 0-7 of
                                // the quarter (both channels: units 0-3 of the member's box row), the
                                // R lane frames 8-15 (units 4-7), after exchanging 8 values
                                float own[8], got[8];
#pragma unroll
                                for (uint32_t i = 0; i < 8; ++i) {
                                    const float keep = __uint_as_float(c ? acc[16 * qr + 8 + i] : acc[16 * qr + i]) * out_scale;
                                    const float send = __uint_as_float(c ? acc[16 * qr + i] : acc[16 * qr + 8 + i]) * out_scale;
                                    own[i] = keep;
                                    got[i] = __shfl_xor_sync(0xffffffffu, send, 1);
                                }
                                const uint32_t rb = sb0 + ml * 128u;
#pragma unroll
                                for (uint32_t u = 0; u < 4; ++u) {
                                    const float l0 = c ? got[2 * u] : own[2 * u], r0 = c ? own[2 * u] : got[2 * u];
                                    const float l1 = c ? got[2 * u + 1] : own[2 * u + 1], r1 = c ? own[2 * u + 1] : got[2 * u + 1];
                                    sts128(rb + (((c * 4u + u) ^ (ml & 7u)) << 4), l0, r0, l1, r1);
                                }
                            } else {
#pragma unroll
                                for (uint32_t o = 0; o < 16; ++o) {
                                    const uint32_t fi = o * CH + c;
                                    const uint32_t box = fi >> 5, unit = (fi >> 2) & 7u;
                                    sts32(sb0 + box * kBoxBytes + ml * 128u + ((unit ^ (ml & 7u)) << 4) + (fi & 3u) * 4u,
                                          __uint_as_float(acc[16 * qr + o]) * out_scale);
                                }
                            }
                            rc.lap(21);
                            fence_proxy_async();
                            __syncwarp();
                            rc.lap(22);
                            if (lane == 0 && !(P.ablate & 4u)) {
                                const int32_t f0 = (int32_t)((o_start + hf * 32u + qr * 16u) * CH);   // float coordinate
#pragma unroll
                                for (uint32_t bx = 0; bx < (uint32_t)CH / 2u; ++bx)
                                    tensor_s2g_2d(&tmap_out, f0 + (int32_t)(bx * 32u), m_first, sb0 + bx * kBoxBytes);
                                bulk_commit();
                            }
                        }
                        continue;
                    }
                    // the staging buffer's previous stores have read it
                    if (lane == 0) bulk_wait_read<kStageBufs - 1>();
                    __syncwarp();
                    const uint32_t sb = sb0 + (kStageBufs == 2 ? (h_seq & 1u) * kHalfBytes : 0u);
                    ++h_seq;
                    rc.lap(20);
                    if constexpr (CH == 1) {
                        // the thread's 32 frames are one 128-byte box row: eight 16-byte units
                        const uint32_t rb = sb + ml * 128u;
#pragma unroll
                        for (uint32_t u = 0; u < 8; ++u)
                            sts128(rb + ((u ^ (ml & 7u)) << 4), __uint_as_float(acc[4 * u]) * out_scale,
                                   __uint_as_float(acc[4 * u + 1]) * out_scale,
                                   __uint_as_float(acc[4 * u + 2]) * out_scale,
                                   __uint_as_float(acc[4 * u + 3]) * out_scale);
                    } else if constexpr (CH == 2) {
                        // Lane pair (member, L) / (member, R): after exchanging 16 values the L lane
                        // owns frames 0-15 of both channels = box 0's whole 128-byte row of the
                        // member, the R lane frames 16-31 = box 1's row: eight conflict-free 16-byte
                        // stores each instead of 32 scalar ones.
                        float own[16], got[16];
#pragma unroll
                        for (uint32_t i = 0; i < 16; ++i) {
                            const float keep = __uint_as_float(c ? acc[16 + i] : acc[i]) * out_scale;
                            const float send = __uint_as_float(c ? acc[i] : acc[16 + i]) * out_scale;
                            own[i] = keep;
                            got[i] = __shfl_xor_sync(0xffffffffu, send, 1);
                        }
                        const uint32_t rb = sb + c * kBoxBytes + ml * 128u;
#pragma unroll
                        for (uint32_t u = 0; u < 8; ++u) {
                            // frames 2u, 2u + 1 of this lane's 16: (L, R, L, R)
                            const float l0 = c ? got[2 * u] : own[2 * u], r0 = c ? own[2 * u] : got[2 * u];
                            const float l1 = c ? got[2 * u + 1] : own[2 * u + 1], r1 = c ? own[2 * u + 1] : got[2 * u + 1];
                            sts128(rb + ((u ^ (ml & 7u)) << 4), l0, r0, l1, r1);
                        }
                    } else {
                        // float (frame o, channel c) of the member's half-tile row: index o*CH + c;
                        // box = index / 32, 16-byte unit inside the box row XOR-swizzled by the row
#pragma unroll
                        for (uint32_t o = 0; o < 32; ++o) {
                            const uint32_t fi = o * CH + c;
                            const uint32_t box = fi >> 5, unit = (fi >> 2) & 7u;
                            sts32(sb + box * kBoxBytes + ml * 128u + ((unit ^ (ml & 7u)) << 4) + (fi & 3u) * 4u,
                                  __uint_as_float(acc[o]) * out_scale);
                        }
                    }
                    rc.lap(21);
                    fence_proxy_async();
                    __syncwarp();
                    rc.lap(22);
                    if (lane == 0 && !(P.ablate & 4u)) {
                        const int32_t f0 = (int32_t)((o_start + hf * 32u) * CH);   // float coordinate
#pragma unroll
                        for (uint32_t bx = 0; bx < (uint32_t)CH; ++bx)
                            tensor_s2g_2d(&tmap_out, f0 + (int32_t)(bx * 32u), m_first, sb + bx * kBoxBytes);
                        bulk_commit();
                    }
                }
            }
        }
        if (lane == 0) bulk_wait_all();
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kWarpIssuer0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
    if (prof && tid == 0) atomicAdd(&g_tc2_cycles[13], (unsigned long long)(clock64() - t_kernel));
}

// One CTA per stored G tile: builds the tile's banded filter matrix, scaled by 2^13 and split into
// fp16 hi and lo parts, in the canonical K-major core-matrix layout the MMA reads
// ([K group of 8][64 rows][8 halves]), K counted from k0 = the first needed virtual frame rounded
// down to the 16-frame chunk grid.  A thread owns one output row and walks along its two
// coefficient rows, one 16-byte piece (8 K values) of the hi and of the lo matrix per step;
// consecutive lanes write consecutive pieces.  `src_tile[g]` is the tile a stored matrix is built
// from (identity when the tiles are not de-duplicated).
template <int TAPS>
__global__ void __launch_bounds__(256)
tc2_gmat_kernel(const UnitDev *units, const PlanEntry *entries, const float *coeffs, const Tc2Tile *tct,
                const uint32_t *src_tile, const uint32_t *n_stored, uint8_t *gmat, uint32_t kt_max, float comp) {
    const UnitDev &U = units[0];
    const uint32_t g = blockIdx.x;
    if (n_stored ? g >= *n_stored : false) return;
    const uint32_t t = src_tile ? src_tile[g] : g;
    const uint32_t n_tiles = (uint32_t)((U.total_out + kN - 1) / kN);
    if (t >= n_tiles) return;
    const Tc2Tile m = tct[t];
    const PlanEntry *ent = entries + ((size_t)U.tile_off * kTileOut + (size_t)t * kN);
    const uint32_t tid = threadIdx.x;
    const uint32_t o = tid & 63u;                        // output row of this thread
    const bool row_ok = o < m.n_out;
    PlanEntry e = ent[0];
    if (row_ok) e = ent[o];
    const uint32_t p1 = e.phase1;
    const uint32_t p2 = p1 + 1 < kPhases - 1 ? p1 + 1 : kPhases - 1;
    const float fr = e.frac, omf = __fsub_rn(1.0f, fr);
    const int d = e.v - m.k0;                            // K index of the row's tap 0
    const float *ca = coeffs + (size_t)p1 * TAPS;
    const float *cb = coeffs + (size_t)p2 * TAPS;
    const size_t g_bytes = (size_t)kt_max * 128u;
    const size_t tile_bytes = 2 * g_bytes + kListBytes;
    uint4 *dst_hi = reinterpret_cast<uint4 *>(gmat + (size_t)g * tile_bytes);
    uint4 *dst_lo = reinterpret_cast<uint4 *>(gmat + (size_t)g * tile_bytes + g_bytes);
    // ---- the tile's main-pass MMA list ----
    __shared__ int s_v[kN];
    if (tid < kN) s_v[tid] = ent[tid < m.n_out ? tid : 0].v;
    __syncthreads();
    if (tid == 0) {
        uint32_t *list = reinterpret_cast<uint32_t *>(gmat + (size_t)g * tile_bytes + 2 * g_bytes);
        const int n_ks = (int)(m.kt / kChunk);
        // K steps holding the main lobe (taps TAPS/2 - 1 and TAPS/2) of each quarter's outputs
        int lo[4], hi[4];
        for (int q = 0; q < 4; ++q) {
            lo[q] = 1;
            hi[q] = 0;
            if (16u * q < m.n_out) {
                const uint32_t last = min(16u * q + 15u, m.n_out - 1u);
                lo[q] = (s_v[16 * q] - m.k0 + TAPS / 2 - 1) / (int)kChunk;
                hi[q] = min((s_v[last] - m.k0 + TAPS / 2) / (int)kChunk, n_ks - 1);
            }
        }
        // output quarters whose band overlaps K step ks (the other blocks of G are zero)
        auto active = [&](int ks) {
            uint32_t mk = 0;
            for (uint32_t q = 0; q < 4; ++q) {
                if (16u * q >= m.n_out) break;
                const uint32_t last = min(16u * q + 15u, m.n_out - 1u);
                const int d_first = s_v[16 * q] - m.k0, d_last = s_v[last] - m.k0;
                if (16 * ks + 16 > d_first && 16 * ks < d_last + TAPS) mk |= 1u << q;
            }
            return mk;
        };
        uint32_t n = 0;
        auto emit_ranges = [&](int ks, uint32_t mask) {
            mask &= active(ks);
            for (uint32_t q = 0; q < 4;) {
                if (!((mask >> q) & 1u)) { ++q; continue; }
                uint32_t q1 = q;
                while (q1 < 4 && ((mask >> q1) & 1u)) ++q1;
                if (n < kListEntries - 1) list[n++] = (uint32_t)ks | (q << 8) | ((q1 - q) << 12);
                q = q1;
            }
        };
        auto central = [&](int ks) {
            uint32_t mk = 0;
            for (int q = 0; q < 4; ++q)
                if (ks >= lo[q] && ks <= hi[q]) mk |= 1u << q;
            return mk;
        };
        // early: outside-in over all K steps, for the quarters the step is not central for
        for (int f = 0, b = n_ks - 1; f <= b; ++f, --b) {
            emit_ranges(f, ~central(f) & 0xfu);
            if (b != f) emit_ranges(b, ~central(b) & 0xfu);
        }
        // late: ascending, for the quarters the step is central for
        for (int ks = 0; ks < n_ks; ++ks) emit_ranges(ks, central(ks));
        // K-step table of the lo passes: first active quarter | quarters << 4 (contiguous: the band
        // moves monotonically through the tile); a step without any active quarter keeps full width
        uint8_t *tab = reinterpret_cast<uint8_t *>(list + kListEntries);
        for (int ks = 0; ks < 16; ++ks) {
            uint32_t e = 0x40u;
            if (ks < n_ks) {
                const uint32_t mk = active(ks);
                if (mk) {
                    const uint32_t q0 = (uint32_t)__ffs((int)mk) - 1u, q1 = 31u - (uint32_t)__clz((int)mk);
                    e = q0 | ((q1 - q0 + 1u) << 4);
                }
            }
            tab[ks] = (uint8_t)e;
        }
        for (; n < kListEntries; ++n) list[n] = kListEnd;
    }
    for (uint32_t kg = tid >> 6; kg < m.kt / 8; kg += 4) {
        uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
        const int j0 = (int)(8 * kg) - d;
        if (row_ok && j0 > -8 && j0 < TAPS) {
            float gv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int j = j0 + q;
                gv[q] = 0.f;
                if (j >= 0 && j < TAPS)
                    gv[q] = __fmaf_rn(__ldg(cb + j), fr, __fmul_rn(__ldg(ca + j), omf)) * kScaleG;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const __half2 h = __floats2half2_rn(gv[2 * q], gv[2 * q + 1]);
                const float2 hf = __half22float2(h);
                // lo = residual + g * comp: the compensation of the accumulator's truncation rides
                // in the lo matrix (the residual is ~2^-12 g, fp32 resolves the 2^-24 g term easily;
                // (1 + 0.85 * 2^-24) itself is not representable in fp32, so the epilogue cannot apply it)
                const __half2 l = __floats2half2_rn(__fmaf_rn(gv[2 * q], comp, __fsub_rn(gv[2 * q], hf.x)),
                                                    __fmaf_rn(gv[2 * q + 1], comp, __fsub_rn(gv[2 * q + 1], hf.y)));
                hi[q] = *reinterpret_cast<const uint32_t *>(&h);
                lo[q] = *reinterpret_cast<const uint32_t *>(&l);
            }
        }
        dst_hi[kg * kN + o] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        dst_lo[kg * kN + o] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// Tile records of the tensor kernel, one thread per 64-output tile (= two 32-output plan tiles,
// whose entries are adjacent): K range on the chunk grid anchored at virtual frame H.
__global__ void tc2_tiles_kernel(const UnitDev *units, const PlanEntry *entries, Tc2Tile *tct, uint32_t taps,
                                 uint32_t kt_max) {
    const UnitDev &U = units[0];
    const uint32_t n_tiles = (uint32_t)((U.total_out + kN - 1) / kN);
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const uint64_t o0 = (uint64_t)t * kN;
    const uint32_t n_out = (uint32_t)min((uint64_t)kN, U.total_out - o0);
    const PlanEntry *ent = entries + ((size_t)U.tile_off * kTileOut + o0);
    const int32_t H = (int32_t)U.hist_len0;
    const int32_t v0 = ent[0].v, vl = ent[n_out - 1].v;
    Tc2Tile m;
    m.k0 = v0 - ((((v0 - H) % (int32_t)kChunk) + (int32_t)kChunk) % (int32_t)kChunk);
    uint32_t kt = (uint32_t)(vl + (int32_t)taps - m.k0 + (int32_t)kChunk - 1) & ~(kChunk - 1u);
    if (kt > kt_max) kt = kt_max;    // cannot happen for a plan of this ratio (kt_max_of)
    m.kt = kt;
    m.n_out = n_out;
    m.o_start = (uint32_t)o0;
    m.g_idx = t;
    m.pad[0] = m.pad[1] = m.pad[2] = 0;
    tct[t] = m;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
            !fn) {
            cudaGetLastError();
            return nullptr;
        }
        encode = (EncodeTiledFn)fn;
    }
    return encode;
}

}  // namespace

bool tc2_supported(uint32_t channels, uint32_t taps, double ratio) {
    if (channels != 1 && channels != 2 && channels != 4 && channels != 8) return false;
    if (taps != 16 && taps != 32 && taps != 64 && taps != 128) return false;
    // one tile's K range plus a chunk of slack must fit the ring
    return kt_max_of(taps, ratio) + kChunk <= kRing;
}

uint32_t tc2_kt_extent(uint32_t taps, double ratio) { return kt_max_of(taps, ratio); }

size_t tc2_gmat_bytes_per_tile(uint32_t taps, double ratio) {
    return (size_t)2 * kt_max_of(taps, ratio) * 128u + kListBytes;
}

uint32_t tc2_rows_per_group() { return kRows; }

// 2^-17 undoes the operand prescale (exact)
float tc2_out_scale() { return kScaleOut; }

// Two issuers keep two consecutive tiles in flight: their K ranges (the second starts up to
// floor(64*ratio)+1 frames later, rounded to the chunk grid) must fit the ring together.
uint32_t tc2_issuers(uint32_t taps, double ratio) {
    const uint32_t adv = (uint32_t)(64.0 * ratio) + 1u;
    return kt_max_of(taps, ratio) + adv + 2u * kChunk <= kRing ? 2u : 1u;
}

// Output as a 2-D tensor: rows = members at one constant stride, inner = f32 values
// (frames x channels), clipped to `valid_frames`; box = 32 floats (128 bytes) x 32 / channels
// members, 128B swizzle (the epilogue's staging layout).
bool tc2_make_output_tensor_map(CUtensorMap *out, float *base, uint64_t stride_bytes, uint64_t valid_frames,
                                uint32_t n_members, uint32_t channels) {
    EncodeTiledFn encode = encode_tiled();
    if (!encode) return false;
    if (channels != 1 && channels != 2 && channels != 4 && channels != 8) return false;
    if (valid_frames == 0 || valid_frames * channels >= (1ull << 31)) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (stride_bytes & 15u) || stride_bytes == 0) return false;
    if (stride_bytes < valid_frames * channels * 4ull && !getenv("RSB_DEBUG_FAKE_OUT_STRIDE")) return false;
    cuuint64_t dims[2] = {valid_frames * channels, n_members};
    cuuint64_t strides[1] = {stride_bytes};
    cuuint32_t box[2] = {32u, 32u / channels};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

void launch_tc2_tiles(const UnitDev *units, const PlanEntry *entries, Tc2Tile *tct, uint32_t taps, double ratio,
                      uint32_t tile_cap, cudaStream_t stream) {
    if (tile_cap == 0) return;
    tc2_tiles_kernel<<<(tile_cap + 127) / 128, 128, 0, stream>>>(units, entries, tct, taps,
                                                                   kt_max_of(taps, ratio));
}

void launch_tc2_gmat(const UnitDev *units, const PlanEntry *entries, const float *coeffs, const Tc2Tile *tct,
                     const uint32_t *src_tile, const uint32_t *n_stored, uint8_t *gmat, uint32_t taps,
                     double ratio, uint32_t max_stored, double comp_units, cudaStream_t stream) {
    if (max_stored == 0) return;
    const float comp = (float)(comp_units * 5.9604644775390625e-08);   // units of 2^-24
    const uint32_t kt_max = kt_max_of(taps, ratio);
    switch (taps) {
        case 16: tc2_gmat_kernel<16><<<max_stored, 256, 0, stream>>>(units, entries, coeffs, tct, src_tile, n_stored, gmat, kt_max, comp); break;
        case 32: tc2_gmat_kernel<32><<<max_stored, 256, 0, stream>>>(units, entries, coeffs, tct, src_tile, n_stored, gmat, kt_max, comp); break;
        case 64: tc2_gmat_kernel<64><<<max_stored, 256, 0, stream>>>(units, entries, coeffs, tct, src_tile, n_stored, gmat, kt_max, comp); break;
        default: tc2_gmat_kernel<128><<<max_stored, 256, 0, stream>>>(units, entries, coeffs, tct, src_tile, n_stored, gmat, kt_max, comp); break;
    }
}

// role layouts that are compiled (ST, ET, XS); index = Tc2Params::variant
struct Variant { uint32_t st, et, xs, qs; };
static size_t epi_stage_bytes(uint32_t channels, const Variant &V) {
    const size_t box = (32u / channels) * 128u < 1024u ? 1024u : (32u / channels) * 128u;
    const bool quarter = V.qs && channels >= 2 && V.et == 2;
    return 2 * 4 * (quarter ? (channels / 2) * box : channels * box);
}
static const Variant kVariants[] = {{3, 2, 6, 1}, {2, 1, 4, 0}, {2, 2, 4, 0}, {4, 1, 8, 0}, {2, 1, 8, 0}, {3, 1, 6, 0}, {3, 2, 3, 0}, {3, 2, 6, 0}};
constexpr uint32_t kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

uint32_t tc2_g_stages(uint32_t channels, uint32_t taps, double ratio, uint32_t variant) {
    const Variant V = kVariants[variant < kNumVariants ? variant : 0];
    const size_t fixed = (size_t)V.xs * kXStageBytes + epi_stage_bytes(channels, V) + 2048;   // + static barriers
    const size_t per_stage = tc2_gmat_bytes_per_tile(taps, ratio);
    const size_t budget = 232448;    // 227 KB per CTA
    size_t n = (budget - fixed) / per_stage;
    if (n > kMaxGStages) n = kMaxGStages;
    return (uint32_t)n;
}

static void ensure_hang_buffer();

bool launch_conv_tc2(const Tc2Params &p, const CUtensorMap &tmap_in, const CUtensorMap &tmap_out, int sm_count,
                     bool leave_sm_free, cudaStream_t stream) {
    const uint32_t variant = p.variant < kNumVariants ? p.variant : 0;
    const Variant V = kVariants[variant];
    const size_t smem = (size_t)V.xs * kXStageBytes + epi_stage_bytes(p.channels, V) + (size_t)p.g_stages * (2 * p.kt_max * 128u + kListBytes);
    if (p.g_stages < 2) return false;
    ensure_hang_buffer();
    // one SM is left free when the next submit's (serial) plan kernel may need somewhere to run
    const uint32_t grid = (uint32_t)(leave_sm_free && sm_count > 8 ? sm_count - 1 : sm_count);
    auto launch = [&](auto kern, uint32_t threads) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, threads, smem, stream>>>(p, tmap_in, tmap_out);
    };
    const uint32_t thr = role_threads(V.st, V.et);
    {
        static uint32_t cur[2] = {kSuspendHint, kSuspendHint};
        const uint32_t want[2] = {p.hint_crit, p.hint_other};
        if (want[0] != cur[0] || want[1] != cur[1]) {
            cudaMemcpyToSymbolAsync(g_tc2_hint, want, sizeof(want), 0, cudaMemcpyHostToDevice, stream);
            cur[0] = want[0];
            cur[1] = want[1];
        }
    }
    if (variant != 0) {
        // experiment variants: stereo f32 only
        if (p.channels != 2 || p.raw16) return false;
        switch (variant) {
            case 1: launch(conv_tc2_kernel<2, 0, 2, 2, 1, 4, 0>, thr); break;
            case 2: launch(conv_tc2_kernel<2, 0, 2, 2, 2, 4, 0>, thr); break;
            case 3: launch(conv_tc2_kernel<2, 0, 2, 4, 1, 8, 0>, thr); break;
            case 4: launch(conv_tc2_kernel<2, 0, 2, 2, 1, 8, 0>, thr); break;
            case 6: launch(conv_tc2_kernel<2, 0, 2, 3, 2, 3, 0>, thr); break;
            case 7: launch(conv_tc2_kernel<2, 0, 2, 3, 2, 6, 0>, thr); break;     // round-2 layout before quarter staging
            default: launch(conv_tc2_kernel<2, 0, 2, 3, 1, 6, 0>, thr); break;
        }
        return true;
    }
    switch (p.channels) {
        case 1:
            if (p.raw16 && p.raw_bytes == 3) launch(conv_tc2_kernel<1, 1, 3>, thr);
            else if (p.raw16) launch(conv_tc2_kernel<1, 1>, thr);
            else launch(conv_tc2_kernel<1>, thr);
            break;
        case 2:
            if (p.raw16 == 2 && p.raw_bytes == 3) launch(conv_tc2_kernel<2, 2, 3>, thr);
            else if (p.raw16 == 1 && p.raw_bytes == 3) launch(conv_tc2_kernel<2, 1, 3>, thr);
            else if (p.raw16 == 2) launch(conv_tc2_kernel<2, 2>, thr);
            else if (p.raw16 == 1) launch(conv_tc2_kernel<2, 1>, thr);
            else launch(conv_tc2_kernel<2>, thr);
            break;
        case 4: launch(conv_tc2_kernel<4>, thr); break;
        default: launch(conv_tc2_kernel<8>, thr); break;
    }
    return true;
}

// Watchdog record {tag, block, warp, parity | barrier address << 8, then per warp of that block:
// tag | parity << 8 | barrier address << 12} of a launch that trapped; the
// buffer is pinned host memory mapped into the device (readable after the context has failed).
static unsigned int *g_hang_host = nullptr;
void tc2_hang_record(unsigned int out[32]) {
    for (int i = 0; i < 32; ++i) out[i] = g_hang_host ? g_hang_host[i] : 0u;
    if (g_hang_host) for (int i = 0; i < 32; ++i) g_hang_host[i] = 0u;
}
static void ensure_hang_buffer() {
    if (g_hang_host) return;
    void *h = nullptr;
    if (cudaHostAlloc(&h, 128, cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return; }
    std::memset(h, 0, 128);
    void *d = nullptr;
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(h); return; }
    if (cudaMemcpyToSymbol(g_tc2_hang, &d, sizeof(d)) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(h); return; }
    g_hang_host = static_cast<unsigned int *>(h);
    const int wd = getenv("RSB_TC_WATCHDOG") && atoi(getenv("RSB_TC_WATCHDOG")) ? 1 : 0;
    cudaMemcpyToSymbol(g_tc2_watchdog, &wd, sizeof(int));
}

void tc2_phase_profile(int enable, unsigned long long *out, uint32_t count) {
    unsigned long long tmp[32];
    cudaMemcpyFromSymbol(tmp, g_tc2_cycles, sizeof(tmp));
    if (out) std::memcpy(out, tmp, sizeof(unsigned long long) * (count < 32u ? count : 32u));
    unsigned long long zero[32] = {0};
    cudaMemcpyToSymbol(g_tc2_cycles, zero, sizeof(zero));
    cudaMemcpyToSymbol(g_tc2_prof, &enable, sizeof(int));
}

}  // namespace rsb
