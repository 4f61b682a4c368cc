// fir_tensor.cu -- the tensor-core convolution kernel for sm_100a (tcgen05 + TMEM + TMA).
//
// Formulation.  Same banded product as the FFMA2 kernel (fir_fast.cu): for a tile of 32
// consecutive output frames,  OUT[row][k] = sum_j X[row][j] * G[k][j]  with row = one channel of
// one member stream and G the tile's interpolated, banded filter rows (the reference blends two
// dot products per output, fir/avx512.rs:41-45; blending the rows first is the same linear map).
// Here the product runs on the 5th-generation tensor cores as  D[128 rows x 32 outputs] +=
// A[128 x 8] * B[32 x 8]^T  per K step of 8 input frames (tcgen05.mma kind::tf32, M = 128,
// N = 32, accumulator in tensor memory).
//
// Precision.  TF32 keeps 11 significant bits, so every operand is split into a TF32 "hi" part and
// a TF32 "lo" part (x = hi + lo + O(2^-22 |x|)) and three products are accumulated in fp32:
// lo*hi, hi*lo, then hi*hi ("3xTF32"); the small terms first and the hi*hi K steps outside-in
// (both ends of the band towards its centre) so that the large centre taps are added last.
// The parity bar against the oracle is 1e-6 absolute, as for the FFMA2 kernel.
//
// Data flow of one CTA (persistent, one per SM, 16 warps; every hand-off is an mbarrier):
//   * work item = (run of consecutive tiles) x (group of 128 rows = 128 / CH member streams).
//     A run walks forward in time, so every input frame is fetched from HBM/L2 ONCE per row
//     and kept in a ring while the ~6 tiles whose windows cover it are computed (the FFMA2
//     kernel re-fetches each tile's whole window: 5.4x the input through L2, 1.7x from HBM).
//   * warp 9 (one thread): work scheduler (atomic counter) and TMA producer of the input: one
//     2-D tensor copy per 16-frame chunk (box = 16 frames x 128/CH members; mono / stereo with
//     64B / 128B swizzle, 4 / 8 channels unswizzled).  The chunk grid is anchored at the first
//     new input frame: a TMA box must start 16-byte aligned in global memory.
//   * warps 4-7 and 12-15 ("splitter", thread == row, each warpgroup takes 8 of a chunk's 16
//     frames): read the chunk from shared memory, de-interleave, split into hi/lo and write both
//     into TMEM rings (tcgen05.st): ring column == input frame, TMEM lane == row.  The rings are
//     the A operands, no second shared-memory copy exists.
//   * warp 10 (one thread): TMA bulk copies of the tile's G matrices (hi and lo, prebuilt in the
//     canonical no-swizzle K-major core-matrix layout by tc_gmat_kernel) into a 3-stage ring.
//   * warps 8 and 11: issue the tcgen05.mma (63 per 128-tap tile) for alternate tiles, one
//     accumulator each (one warp alone when two tiles' K ranges do not fit the ring together);
//     tcgen05.commit to mbarriers tells the epilogue, the G producer and the splitter.
//   * warps 0-3 (epilogue): tcgen05.ld of the 128 x 32 accumulator, staging through shared
//     memory, coalesced 16-byte stores.
//
// Raw integer input (RAW / SB template parameters; the CLI batch path, resample/src/main.rs:128-156):
// the TMA producer can stream RAW s16 or packed s24 frames instead of f32 (a second tensor map:
// s16 as 4-byte stereo frames in 64-byte swizzled rows or 2-byte mono frames in 32-byte rows, s24
// as bytes in unswizzled 96- / 48-byte rows) and the splitter converts each sample with
// s * 2^-(bits-1) before the hi/lo split; a mono source of a stereo
// stream is duplicated by letting both channel rows of a member read the same samples.  The
// arithmetic after the conversion is unchanged, so the output is bit-identical to running the
// separate format pass (pcm_ingest.cu) first.
//
// TMEM map (512 columns x 128 lanes): [0,224) X hi ring, [224,448) X lo ring, [448,480) and
// [480,512) the two accumulators.
#include <cstdlib>
#include <cstring>

#include "fir_kernels.h"
#include "sm100_ptx.cuh"

namespace rsb {

// Optional per-role cycle accounting (one thread per role of CTA 0; enabled by a debug call):
// MMA issuer [0] wait x_full [1] wait g_full [2] wait d_empty [3] issue [4] commits + meta;
// splitter [5] wait x_empty [6] wait xs_full [7] loads + split [11] wait::st + arrive [12] tcgen05.st
// [15] loop overhead; epilogue [8] wait d_full [9] tcgen05.ld + staging [10] stores;
// [13] kernel cycles of CTA 0; [14] tiles of CTA 0.
__device__ unsigned long long g_tc_cycles[24];
__device__ int g_tc_prof = 0;

namespace {

using namespace ptx;

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
            atomicAdd(&g_tc_cycles[i], (unsigned long long)(n - t));
            t = n;
        }
    }
    __device__ __forceinline__ void count(int i, unsigned long long n) {
        if (on) atomicAdd(&g_tc_cycles[i], n);
    }
};

constexpr uint32_t kRows = 128;                    // MMA M
constexpr uint32_t kN = kTileOut;                  // MMA N (32 output frames)
constexpr uint32_t kChunk = 16;                    // frames per input chunk / ring slot
constexpr uint32_t kRing = 224;                    // frames in a TMEM ring
constexpr uint32_t kSlots = kRing / kChunk;        // 14
constexpr uint32_t kColHi = 0, kColLo = kRing, kColD = 2 * kRing;
constexpr uint32_t kXStages = 4;                   // TMA landing buffers for input chunks
constexpr uint32_t kXStageBytes = kRows * kChunk * 4;   // 8192 for mono and stereo alike
constexpr uint32_t kGStages = 3;
constexpr uint32_t kTcThreads = 16 * 32;   // warps 0-3 epilogue, 4-7 + 12-15 splitter, 8 + 11 MMA, 9-10 TMA
constexpr uint32_t kKtLimit = 192;                 // largest K extent of a tile (12 + 1 chunks)
constexpr uint32_t kItemSlots = 2;

// tcgen05 instruction descriptor: D = f32, A = B = tf32, both K-major, N = 32, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kN >> 3) << 17) | ((kRows >> 4) << 24);

__host__ __device__ inline uint32_t tc_kt_max(uint32_t taps, double ratio) {
    // first and last output of a tile start at most floor(31*ratio)+1 frames apart; up to 7
    // frames of alignment in front; rounded up to the MMA's K step
    const uint32_t span = (uint32_t)(31.0 * ratio) + 1u;
    return (7u + span + taps + 7u) & ~7u;
}
// floats per member in the epilogue's staging area: 32 frames x channels, + 4 (keeps 16-byte
// alignment, spreads the members over the banks)
__host__ __device__ inline uint32_t tc_stage_pitch(uint32_t ch) { return 32u * ch + 4u; }

struct Item {
    uint32_t t0, t1;       // tiles [t0, t1) of the unit
    uint32_t group;        // member group
    uint32_t n_chunks;     // input chunks of the run
    int32_t vb;            // virtual frame of chunk 0 (multiple of kChunk)
    uint32_t valid;
};

// Shared-memory descriptor of a B operand: K-major, no swizzle; the two 16-byte K halves of a
// K step are 512 bytes apart (LBO), 8-row groups 128 bytes apart (SBO); descriptor version 1.
// The low word holds the start address (>> 4) and the LBO, the high word SBO and version.
__device__ __forceinline__ uint32_t b_desc_lo(uint32_t smem_addr) {
    return ((smem_addr & 0x3ffffu) >> 4) | ((512u >> 4) << 16);
}
constexpr uint32_t kBDescHi = (128u >> 4) | (1u << 14);
constexpr uint32_t kBDescKStep = 1024u >> 4;   // descriptor increment per K step of 8 frames

struct TcSmem {
    uint64_t xs_full[kXStages], xs_empty[kXStages];
    uint64_t g_full[kGStages];
    uint64_t t_done[4];       // tile t's MMAs have completed: barrier t & 3 (count 1, tcgen05.commit)
    uint64_t x_full[kSlots], x_empty[kSlots];
    uint64_t d_empty[2];
    uint64_t item_full[kItemSlots], item_empty[kItemSlots];
    Item item[kItemSlots];
    uint32_t tmem_base;
};

// RAW: 0 = f32 input; 1 = raw s16 frames of CH channels (CH = 1 or 2); 2 = raw MONO s16 frames
// duplicated into both channels of a stereo stream (the CLI's mono -> stereo, main.rs:139-146)
// SB: bytes per raw sample, 2 (s16) or 3 (packed little-endian s24)
template <int CH, int RAW = 0, int SB = 2>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ TcParams P, const __grid_constant__ CUtensorMap tmap) {
    static_assert(CH == 1 || CH == 2 || CH == 4 || CH == 8, "tensor kernel: 1, 2, 4 or 8 channels");
    static_assert(RAW == 0 || (RAW == 1 && CH <= 2) || (RAW == 2 && CH == 2), "raw s16 input: mono / stereo");
    static_assert(SB == 2 || (SB == 3 && RAW != 0), "raw samples: 16 or packed 24 bits");
    constexpr bool kRawStereo = RAW == 1 && CH == 2;   // s16: 4-byte frames, 64-byte rows, 64B swizzle
    constexpr bool kRawMono = (RAW == 1 && CH == 1) || RAW == 2;   // s16: 2-byte frames, 32-byte rows
    // packed s24: 6-byte stereo / 3-byte mono frames in unswizzled 96- / 48-byte rows (the tensor
    // map's element is one byte)
    constexpr uint32_t kRawFrameBytes = RAW == 0 ? 0u : (kRawStereo ? 2u : 1u) * SB;
    // bytes one input chunk lands in shared memory: f32 rows, or one row of 16 raw frames per member
    constexpr uint32_t kXLandBytes = RAW != 0 ? (kRows / CH) * kChunk * kRawFrameBytes : kXStageBytes;
    constexpr uint32_t kMpg = kRows / CH;               // members per group
    extern __shared__ __align__(1024) uint8_t smem_tc[];
    __shared__ TcSmem S;
    const uint32_t g_bytes = P.kt_max * 128u;           // one G half: kt_max/4 K groups x 512 B
    uint8_t *xst = smem_tc;                             // [kXStages][kXStageBytes]
    uint8_t *gst = smem_tc + kXStages * kXStageBytes;   // [kGStages][2][g_bytes]
    float *stage = reinterpret_cast<float *>(gst + kGStages * 2 * g_bytes);

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const bool prof = g_tc_prof != 0 && blockIdx.x == 0;
    const long long t_kernel = prof ? clock64() : 0;
    RoleClock rc;
    rc.start(false);
    if (tid == 0) {
        for (uint32_t i = 0; i < kXStages; ++i) { mbar_init(&S.xs_full[i], 1); mbar_init(&S.xs_empty[i], 2 * kRows); }
        for (uint32_t i = 0; i < kGStages; ++i) mbar_init(&S.g_full[i], 1);
        for (uint32_t i = 0; i < 4; ++i) mbar_init(&S.t_done[i], 1);
        for (uint32_t i = 0; i < kSlots; ++i) { mbar_init(&S.x_full[i], 2 * kRows); mbar_init(&S.x_empty[i], P.issuers); }
        for (uint32_t i = 0; i < 2; ++i) mbar_init(&S.d_empty[i], kRows);
        for (uint32_t i = 0; i < kItemSlots; ++i) {
            mbar_init(&S.item_full[i], 1);
            mbar_init(&S.item_empty[i], 3 + 3 * kRows);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(&S.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    const UnitDev &U = P.units[0];
    const int32_t H = (int32_t)U.hist_len0;
    const uint32_t n_tiles = U.n_tiles;
    const TcTile *tct = P.tct + U.tile_off;
    const uint32_t n_runs = (n_tiles + P.run_tiles - 1) / P.run_tiles;
    const uint32_t n_items = n_runs * P.groups;

    // every consumer role fetches the next item from the scheduler's two-slot queue
    auto get_item = [&](uint32_t it) {
        const uint32_t slot = it % kItemSlots;
        mbar_wait(&S.item_full[slot], (it / kItemSlots) & 1u);
        const Item I = S.item[slot];
        mbar_arrive(&S.item_empty[slot]);
        return I;
    };

    // the same for a whole converged warp (one arrival)
    auto get_item_warp = [&](uint32_t it) {
        const uint32_t slot = it % kItemSlots;
        mbar_wait(&S.item_full[slot], (it / kItemSlots) & 1u);
        const Item I = S.item[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.item_empty[slot]);
        return I;
    };

    if (warp == 9) {
        // ===== scheduler + TMA producer of the input chunks =====
        if (lane == 0) {
            uint32_t xs_seq = 0;
            for (uint32_t it = 0;; ++it) {
                const uint32_t slot = it % kItemSlots;
                mbar_wait(&S.item_empty[slot], ((it / kItemSlots) & 1u) ^ 1u);
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
                    const TcTile first = tct[I.t0], last = tct[I.t1 - 1];
                    // chunk grid anchored at the first frame of the new input (virtual frame H): a
                    // TMA box must start 16-byte aligned in global memory (an odd stereo frame is
                    // an illegal instruction, tools/experiments/tma_swizzle_test.cu), and no
                    // chunk straddles the history / input seam
                    const int32_t rel = first.k0 - H;
                    const int32_t fl = rel >= 0 ? rel / (int32_t)kChunk
                                                : -((-rel + (int32_t)kChunk - 1) / (int32_t)kChunk);
                    I.vb = H + fl * (int32_t)kChunk;
                    I.n_chunks = (uint32_t)(last.k0 + (int32_t)last.kt - I.vb + (int32_t)kChunk - 1) / kChunk;
                }
                S.item[slot] = I;
                mbar_arrive(&S.item_full[slot]);
                if (!I.valid) break;
                const int32_t m0 = (int32_t)(I.group * kMpg);
                for (uint32_t j = 0; j < I.n_chunks; ++j) {
                    const int32_t v = I.vb + (int32_t)(j * kChunk);
                    if (v < H) continue;       // touches the history: the splitter loads it itself
                    const uint32_t s = xs_seq % kXStages;
                    mbar_wait(&S.xs_empty[s], ((xs_seq / kXStages) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&S.xs_full[s], kXLandBytes);
                    // inner coordinate in tensor-map elements: frames (mono f32, stereo 8-byte
                    // frames) or floats (4 / 8 channels)
                    tensor_g2s_2d(xst + s * kXStageBytes, &tmap,
                                  (v - H) * (int32_t)(SB == 3 ? kRawFrameBytes : CH >= 4 ? CH : 1), m0,
                                  &S.xs_full[s]);
                    ++xs_seq;
                }
            }
        }
        __syncwarp();
    } else if (warp == 10) {
        // ===== TMA producer of the G matrices =====
        if (lane == 0) {
            uint32_t g_seq = 0;
            const size_t tile_floats = (size_t)2 * P.kt_max * kN;
            for (uint32_t it = 0;; ++it) {
                const Item I = get_item(it);
                if (!I.valid) break;
                uint32_t kt = tct[I.t0].kt;
                for (uint32_t t = I.t0; t < I.t1; ++t) {
                    const uint32_t kt_next = t + 1 < I.t1 ? tct[t + 1].kt : 0u;
                    const uint32_t s = g_seq % kGStages;
                    if (g_seq >= kGStages) {     // the stage's previous tile has been multiplied
                        const uint32_t prev = g_seq - kGStages;
                        mbar_wait(&S.t_done[prev & 3u], (prev >> 2) & 1u);
                    }
                    const uint32_t bytes = kt * 128u;
                    const float *src = P.gmat + (size_t)(U.tile_off + t) * tile_floats;
                    uint8_t *dst = gst + (size_t)s * 2 * g_bytes;
                    mbar_arrive_expect_tx(&S.g_full[s], 2 * bytes);
                    bulk_g2s(dst, src, bytes, &S.g_full[s]);
                    bulk_g2s(dst + g_bytes, src + (size_t)P.kt_max * kN, bytes, &S.g_full[s]);
                    ++g_seq;
                    kt = kt_next;
                }
            }
        }
        __syncwarp();
    } else if (warp == 8 || warp == 11) {
        // ===== MMA issuers.  The whole warp runs the loop (warp-uniform values), one elected lane
        // issues.  Issuing costs ~30 cycles of a warp's time per MMA (a dozen operand moves into
        // uniform registers per tcgen05.mma, plus waits and commits: ~3400 cycles per tile), about
        // twice what the tensor pipe needs for the tile.  So two warps issue alternate tiles into
        // the two accumulators whenever two consecutive tiles' K ranges fit the ring together
        // (P.issuers == 2); otherwise warp 8 issues every tile.  A ring slot is handed back to
        // the splitter once BOTH issuers have committed it. =====
        const uint32_t mine = warp == 8 ? 0u : 1u;
        const uint32_t step = P.issuers;
        // Ring sequence numbers; their slot (q % kSlots) and phase parity ((q / kSlots) & 1) are
        // carried along incrementally: a division by 14 in every wait of the critical warp is a
        // dozen dependent instructions.
        uint32_t q_base = 0;      // the run's chunk 0
        uint32_t q_waited = 0;    // chunks whose x_full barrier has been consumed
        uint32_t q_rel = 0;       // chunks handed back to the splitter (by this warp)
        uint32_t base_slot = 0, w_slot = 0, w_par = 0, r_slot = 0;
        uint32_t d_seq = 0;       // tiles of this CTA so far; tile -> accumulator d_seq & 1
        rc.start(prof && lane == 0 && mine == 0);
        for (uint32_t it = 0;; ++it) {
            const Item I = get_item_warp(it);
            if (!I.valid) break;
            if (mine >= step) continue;           // single-issuer mode: warp 11 only drains the queue
            rc.count(14, I.t1 - I.t0);
            auto wait_next_chunk = [&]() {
                mbar_wait(&S.x_full[w_slot], w_par);
                ++q_waited;
                if (++w_slot == kSlots) { w_slot = 0; w_par ^= 1u; }
            };
            // a chunk is handed back only after this warp has seen it filled (its arrival then
            // belongs to the slot's current use even if the warp's own tiles never read the chunk)
            auto release_chunk = [&]() {
                while (q_waited <= q_rel) wait_next_chunk();
                tc_commit_elect(&S.x_empty[r_slot]);
                ++q_rel;
                if (++r_slot == kSlots) r_slot = 0;
            };
            // this warp's first tile of the run: tile parity follows the CTA's tile sequence
            const uint32_t d_run = d_seq;
            uint32_t t = I.t0 + (step == 2 ? ((d_seq ^ mine) & 1u) : 0u);
            d_seq += t - I.t0;
            TcTile m = {0, 0, 0, 0};
            if (t < I.t1) m = tct[t];
            for (; t < I.t1; t += step, d_seq += step) {
                // the tile whose first frame bounds what this warp still needs: its own next tile
                TcTile mn = m;
                if (t + step < I.t1) mn = tct[t + step];
                const uint32_t j_last = (uint32_t)(m.k0 + (int32_t)m.kt - 1 - I.vb) / kChunk;
                rc.lap(4);
                // Order matters with two issuers.  The accumulator wait comes FIRST: it passes only
                // after the epilogue has drained tile d_seq - 2, hence after tile d_seq - 3 (the
                // OTHER issuer's, and the previous user of this tile's G stage) has completed.  A
                // parity wait on g_full before that could be looking at the stage's previous phase
                // still in flight and pass early (seen as an mbarrier arrival underflow = illegal
                // instruction in the G producer, with short tiles).
                const uint32_t b = d_seq & 1u;
                mbar_wait(&S.d_empty[b], ((d_seq >> 1) & 1u) ^ 1u);
                rc.lap(2);
                const uint32_t gs = d_seq % kGStages;      // G stages follow the tile sequence
                mbar_wait(&S.g_full[gs], (d_seq / kGStages) & 1u);
                rc.lap(1);
                while (q_waited <= q_base + j_last) wait_next_chunk();
                rc.lap(0);
                tc_fence_after();

                const uint32_t d_tmem = tmem + kColD + b * kN;
                const uint32_t col0 = (base_slot * kChunk + (uint32_t)(m.k0 - I.vb)) % kRing;
                const uint32_t n_ks = m.kt >> 3;
                const uint32_t ghi = b_desc_lo(smem_u32(gst + (size_t)gs * 2 * g_bytes));
                const uint32_t glo = ghi + (g_bytes >> 4);
                // small terms first: x_lo * g_hi and x_hi * g_lo, two K steps per issue block
                const uint32_t ahi = tmem + kColHi, alo = tmem + kColLo;
                auto wrap = [](uint32_t c) { return c >= kRing ? c - kRing : c; };
                uint32_t col = col0, dk = 0, ks = 0;
#pragma unroll 2
                for (; ks + 2 <= n_ks; ks += 2) {
                    const uint32_t c1 = wrap(col + 8);
                    tc_mma_tf32_ts_x4(d_tmem, alo + col, ahi + col, alo + c1, ahi + c1, ghi + dk, glo + dk,
                                      ghi + dk + kBDescKStep, glo + dk + kBDescKStep, kBDescHi, kIdesc,
                                      ks != 0);
                    dk += 2 * kBDescKStep;
                    col = wrap(c1 + 8);
                }
                if (ks < n_ks)
                    tc_mma_tf32_ts_x2(d_tmem, alo + col, ghi + dk, ahi + col, glo + dk, kBDescHi, kIdesc,
                                      ks != 0);
                // x_hi * g_hi, K steps outside-in (front, back, front + 1, back - 1, ...)
                uint32_t cf = col0, cb = wrap(col0 + 8 * (n_ks - 1));
                uint32_t df = 0, db = (n_ks - 1) * kBDescKStep, i = 0;
                auto back = [](uint32_t c) { return c >= 8 ? c - 8 : c + kRing - 8; };
#pragma unroll 2
                for (; i + 4 <= n_ks; i += 4) {
                    const uint32_t cf1 = wrap(cf + 8), cb1 = back(cb);
                    tc_mma_tf32_ts_x4(d_tmem, ahi + cf, ahi + cb, ahi + cf1, ahi + cb1, ghi + df, ghi + db,
                                      ghi + df + kBDescKStep, ghi + db - kBDescKStep, kBDescHi, kIdesc, 1u);
                    df += 2 * kBDescKStep;
                    db -= 2 * kBDescKStep;
                    cf = wrap(cf1 + 8);
                    cb = back(cb1);
                }
                for (; i + 2 <= n_ks; i += 2) {
                    tc_mma_tf32_ts_x2(d_tmem, ahi + cf, ghi + df, ahi + cb, ghi + db, kBDescHi, kIdesc, 1u);
                    df += kBDescKStep;
                    db -= kBDescKStep;
                    cf = wrap(cf + 8);
                    cb = back(cb);
                }
                if (i < n_ks) tc_mma_tf32_ts(d_tmem, ahi + cf, ghi + df, kBDescHi, kIdesc, 1u);
                rc.lap(3);
                // one commit tells the epilogue (accumulator ready) and the G producer (stage free)
                tc_commit_elect(&S.t_done[d_seq & 3u]);
                // Chunks that end at or before the next tile's first frame are free again.  (Handing
                // them back earlier, right after the front K steps of this pass, was measured
                // slower: a tcgen05.commit in the middle of the MMA stream stalls the issue.)
                const bool last = t + step >= I.t1;     // this warp's last tile of the run
                while (q_rel < q_base + I.n_chunks &&
                       (last || I.vb + (int32_t)((q_rel - q_base + 1) * kChunk) <= mn.k0))
                    release_chunk();
                m = mn;
            }
            // a run this warp had no tile in
            while (q_rel < q_base + I.n_chunks) release_chunk();
            d_seq = d_run + (I.t1 - I.t0);
            q_base += I.n_chunks;
            base_slot = (base_slot + I.n_chunks) % kSlots;
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===== splitter: shared memory (TMA landing buffer) -> hi / lo rings in TMEM.  Two
        // warpgroups (warps 4-7 and 12-15: the same TMEM lane quadrants); each takes 8 of a
        // chunk's 16 frames, which halves the time from "ring slot free" to "chunk readable". =====
        constexpr uint32_t kHalf = kChunk / 2;
        const uint32_t wg = warp >= 12 ? 1u : 0u;           // which 8 frames of every chunk
        const uint32_t row = tid & 127u;                    // TMEM lane
        const uint32_t ml = row / CH, c = row % CH;         // member inside the group, channel
        const uint32_t lane_base = ((warp & 3u) * 32u) << 16;
        uint32_t q_seq = 0, xs_seq = 0;
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
                rc.lap(15);
                // Load and split BEFORE waiting for the ring slot: the slot becomes free only when
                // the tile(s) still reading it have completed, and from then on the MMA issuer is
                // waiting for this chunk; with the values already split in registers only the
                // TMEM stores remain on that critical path.
                const int32_t v = I.vb + (int32_t)(j * kChunk);
                float x[kHalf];
                int fast_slot = -1;
                if (v >= H) {
                    const uint32_t s = xs_seq % kXStages;
                    mbar_wait(&S.xs_full[s], (xs_seq / kXStages) & 1u);
                    __syncwarp();
                    rc.lap(6);
                    const uint32_t base = smem_u32(xst + s * kXStageBytes);
                    if (SB == 3 && kRawStereo) {
                        // packed s24 stereo: unswizzled 96-byte rows; this half's 8 frames are 48
                        // bytes = three 16-byte units; sample k of channel c sits at byte 6k + 3c
                        const uint32_t rb = base + ml * 96u + wg * 48u;
                        uint32_t w[13];
#pragma unroll
                        for (uint32_t u = 0; u < 3; ++u) {
                            const float4 q4 = lds128(rb + 16u * u);
                            w[4 * u] = __float_as_uint(q4.x); w[4 * u + 1] = __float_as_uint(q4.y);
                            w[4 * u + 2] = __float_as_uint(q4.z); w[4 * u + 3] = __float_as_uint(q4.w);
                        }
                        w[12] = 0;
#pragma unroll
                        for (uint32_t k = 0; k < 8; ++k) {
                            const uint32_t o0 = 6 * k, o1 = 6 * k + 3;
                            const uint32_t v0 = __funnelshift_r(w[o0 >> 2], w[(o0 >> 2) + 1], (o0 & 3u) * 8u);
                            const uint32_t v1 = __funnelshift_r(w[o1 >> 2], w[(o1 >> 2) + 1], (o1 & 3u) * 8u);
                            const uint32_t vv = c ? v1 : v0;
                            x[k] = (float)((int)(vv << 8) >> 8) * (1.0f / 8388608.0f);
                        }
                    } else if (SB == 3 && kRawMono) {
                        // packed s24 mono: unswizzled 48-byte rows; this half's 8 frames are 24 bytes
                        const uint32_t rb = base + ml * 48u + wg * 24u;
                        uint32_t w[7];
#pragma unroll
                        for (uint32_t u = 0; u < 3; ++u) {
                            const float2 q2 = lds64(rb + 8u * u);
                            w[2 * u] = __float_as_uint(q2.x); w[2 * u + 1] = __float_as_uint(q2.y);
                        }
                        w[6] = 0;
#pragma unroll
                        for (uint32_t k = 0; k < 8; ++k) {
                            const uint32_t o = 3 * k;
                            const uint32_t vv = __funnelshift_r(w[o >> 2], w[(o >> 2) + 1], (o & 3u) * 8u);
                            x[k] = (float)((int)(vv << 8) >> 8) * (1.0f / 8388608.0f);
                        }
                    } else if (kRawMono) {
                        // raw s16 mono: unswizzled 32-byte rows (16 frames), one per member; both
                        // channel rows of a stereo member read the same samples (RAW == 2)
                        const float4 q4 = lds128(base + ml * 32u + wg * 16u);
                        const uint32_t w[4] = {__float_as_uint(q4.x), __float_as_uint(q4.y),
                                               __float_as_uint(q4.z), __float_as_uint(q4.w)};
#pragma unroll
                        for (uint32_t k = 0; k < 4; ++k) {
                            x[2 * k] = (float)(int)(short)(w[k] & 0xffffu) * (1.0f / 32768.0f);
                            x[2 * k + 1] = (float)((int)w[k] >> 16) * (1.0f / 32768.0f);
                        }
                    } else if (kRawStereo) {
                        // raw s16 stereo: 64-byte rows (16 frames of 2 x s16), 64B swizzle: unit u
                        // at u ^ ((row >> 1) & 3); a word is one frame, the channel picks its half;
                        // s / 2^15 is the reference's `s as f32 / 32768.0` (main.rs:131-136)
                        const uint32_t rb = base + ml * 64u;
#pragma unroll
                        for (uint32_t u = 0; u < 2; ++u) {
                            const float4 q4 = lds128(rb + (((u + 2 * wg) ^ ((ml >> 1) & 3u)) << 4));
                            const uint32_t w[4] = {__float_as_uint(q4.x), __float_as_uint(q4.y),
                                                   __float_as_uint(q4.z), __float_as_uint(q4.w)};
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) {
                                const int sv = c ? (int)w[k] >> 16 : (int)(short)(w[k] & 0xffffu);
                                x[4 * u + k] = (float)sv * (1.0f / 32768.0f);
                            }
                        }
                    } else if (CH == 2) {
                        // 128-byte rows (16 stereo frames), 128B swizzle: unit u at u ^ (row & 7)
                        const uint32_t rb = base + ml * 128u;
#pragma unroll
                        for (uint32_t u = 0; u < 4; ++u) {
                            const float4 q4 = lds128(rb + (((u + 4 * wg) ^ (ml & 7u)) << 4));
                            x[2 * u] = c ? q4.y : q4.x;
                            x[2 * u + 1] = c ? q4.w : q4.z;
                        }
                    } else if (CH == 1) {
                        // 64-byte rows (16 mono frames), 64B swizzle: unit u at u ^ ((row >> 1) & 3)
                        const uint32_t rb = base + ml * 64u;
#pragma unroll
                        for (uint32_t u = 0; u < 2; ++u) {
                            const float4 q4 = lds128(rb + (((u + 2 * wg) ^ ((ml >> 1) & 3u)) << 4));
                            x[4 * u] = q4.x; x[4 * u + 1] = q4.y; x[4 * u + 2] = q4.z; x[4 * u + 3] = q4.w;
                        }
                    } else {
                        // 4 / 8 channels: rows of 16 frames x CH floats, no swizzle; one 4-byte load
                        // per frame (the 128/CH members of a warp's lanes collide on the banks
                        // CH-fold at most 4 ways: a few dozen wavefronts per chunk, negligible)
                        const float *rowp = reinterpret_cast<const float *>(xst + s * kXStageBytes) +
                                            ml * (kChunk * CH) + (kHalf * wg) * CH + c;
#pragma unroll
                        for (uint32_t f = 0; f < kHalf; ++f) x[f] = rowp[f * CH];
                    }
                    fast_slot = (int)s;
                    ++xs_seq;
                } else {
                    // the chunk lies in the history (only at the very start of a batch)
#pragma unroll
                    for (uint32_t f = 0; f < kHalf; ++f) {
                        const int64_t vv = (int64_t)v + kHalf * wg + f;
                        float xv = 0.f;
                        if (member_ok && vv >= 0) {
                            if (vv < H) xv = hist[((int64_t)kHistFrames - H + vv) * CH + c];
                            else if ((uint64_t)(vv - H) < U.total_frames) xv = in[(vv - H) * CH + c];
                        }
                        x[f] = xv;
                    }
                }
                uint32_t hi[kHalf], lo[kHalf];
#pragma unroll
                for (uint32_t f = 0; f < kHalf; ++f) {
                    const float hv = to_tf32(x[f]);
                    hi[f] = __float_as_uint(hv);
                    lo[f] = __float_as_uint(to_tf32(__fsub_rn(x[f], hv)));
                }
                rc.lap(7);
                // The landing buffer is handed back only after the loaded values have been consumed:
                // the empty asm below takes every split value as an input, so the loads and the
                // arithmetic are complete before the arrive.  (Arriving right behind the ld.shared
                // instructions -- the SASS was LDS.128, LDS.128, SYNCS.ARRIVE back to back -- let
                // the producer's next TMA copy overwrite the buffer before the loads had read it:
                // intermittent single wrong samples in the first rows of a group.)
                asm volatile("" ::"r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]),
                             "r"(hi[6]), "r"(hi[7]), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]),
                             "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7])
                             : "memory");
                if (fast_slot >= 0) mbar_arrive(&S.xs_empty[fast_slot]);
                mbar_wait(&S.x_empty[rs], par);
                __syncwarp();
                rc.lap(5);
                tc_fence_after();
                const uint32_t colw = rs * kChunk + kHalf * wg;
                tmem_st8(tmem + lane_base + kColHi + colw, hi);
                tmem_st8(tmem + lane_base + kColLo + colw, lo);
                rc.lap(12);
                // The stores must have completed before hi[] / lo[] are overwritten by the next
                // chunk: tcgen05.st reads its source registers asynchronously (publishing one
                // chunk late, after the next chunk's arithmetic, left single stale samples in a
                // few lanes, intermittently).
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(&S.x_full[rs]);
                if (++rs == kSlots) { rs = 0; par ^= 1u; }
                rc.lap(11);
            }
        }
    } else {
        // ===== epilogue: accumulator (TMEM) -> shared-memory staging -> global =====
        const uint32_t row = tid;
        const uint32_t ml = row / CH, c = row % CH;
        const uint32_t lane_base = ((warp & 3u) * 32u) << 16;
        const uint32_t pitch = tc_stage_pitch(CH);
        // A member's tile output is a run of 32 * CH floats, cut into 16-byte pieces.
        constexpr uint32_t kPieces = kN * CH / 4;            // pieces per member and tile
        constexpr uint32_t kIter = kMpg * kPieces / kRows;   // pieces per thread and tile (8)
        constexpr uint32_t kMemStep = kRows / kPieces;       // member stride between a thread's pieces
        static_assert(kIter == 8 && kMemStep >= 1, "epilogue piece mapping");
        const uint32_t p_off = (tid % kPieces) * 4u;         // first float of this thread's pieces
        const uint32_t mem0 = tid / kPieces;
        uint32_t d_seq = 0;
        rc.start(prof && tid == 0);
        for (uint32_t it = 0;; ++it) {
            const Item I = get_item(it);
            if (!I.valid) break;
            const uint32_t m0 = I.group * kMpg;
            const uint32_t nm = min(kMpg, U.n_members - m0);
            // this thread's kIter members: output pointer, capacity, 16-byte alignment
            float *outp[kIter];
            uint32_t capf[kIter];
            uint32_t vec_ok = 0;
#pragma unroll
            for (uint32_t i = 0; i < kIter; ++i) {
                const uint32_t mem = mem0 + kMemStep * i;
                outp[i] = nullptr;
                capf[i] = 0;
                if (mem < nm) {
                    const JobDev *job = P.jobs + U.member_off + m0 + mem;
                    outp[i] = job->out;
                    const uint64_t cap = job->out_capacity;
                    capf[i] = cap > 0xffffffffull ? 0xffffffffu : (uint32_t)cap;
                    if ((reinterpret_cast<uintptr_t>(outp[i]) & 15u) == 0) vec_ok |= 1u << i;
                }
            }
            TcTile m_next = tct[I.t0];
            for (uint32_t t = I.t0; t < I.t1; ++t, ++d_seq) {
                const TcTile m = m_next;
                if (t + 1 < I.t1) m_next = tct[t + 1];
                const uint32_t b = d_seq & 1u;
                rc.lap(10);
                mbar_wait(&S.t_done[d_seq & 3u], (d_seq >> 2) & 1u);
                __syncwarp();     // lanes leave the polling loop at different times
                rc.lap(8);
                tc_fence_after();
                uint32_t acc[kN];
                tmem_ld32(tmem + lane_base + kColD + b * kN, acc);
                tmem_wait_ld();
                tc_fence_before();
                mbar_arrive(&S.d_empty[b]);
                // stage[buffer][member][frame][channel]; two buffers, so ONE barrier per tile: the
                // writes of tile t+1 reuse the buffer of tile t-1, whose reads every thread has
                // finished before it arrives at tile t's barrier
                float *sb = stage + (d_seq & 1u) * (kMpg * pitch);
                float *st = sb + ml * pitch + c;
#pragma unroll
                for (uint32_t o = 0; o < kN; ++o) st[o * CH] = __uint_as_float(acc[o]);
                __syncwarp();     // bar.sync counts whole warps: every lane's staging writes first
                named_bar_sync(1, kRows);
                rc.lap(9);
                const uint32_t valid = m.n_out * CH;               // floats of the run that exist
                if (p_off < valid) {
                    const uint64_t g0 = (uint64_t)m.o_start * CH + p_off;   // float index in the stream
                    const bool piece_full = p_off + 4u <= valid;
                    const float *src = sb + mem0 * pitch + p_off;
#pragma unroll
                    for (uint32_t i = 0; i < kIter; ++i) {
                        if (outp[i] == nullptr) continue;
                        const float4 q4 = *reinterpret_cast<const float4 *>(src + i * (kMemStep * pitch));
                        const uint64_t cap_floats = (uint64_t)capf[i] * CH;
                        if (piece_full && ((vec_ok >> i) & 1u) && g0 + 4u <= cap_floats) {
                            *reinterpret_cast<float4 *>(outp[i] + g0) = q4;
                        } else {
                            const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                            for (uint32_t q = 0; q < 4; ++q)
                                if (p_off + q < valid && g0 + q < cap_floats) outp[i][g0 + q] = qv[q];
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
    if (prof && tid == 0) atomicAdd(&g_tc_cycles[13], (unsigned long long)(clock64() - t_kernel));
}

// One CTA per tile: builds the tile's banded filter matrix, split into TF32 hi and lo parts, in
// the canonical K-major core-matrix layout the MMA reads ([K group of 4][32 rows][4 floats]),
// K counted from k0 = the first needed virtual frame rounded down to the 8-frame grid.  A thread
// owns one output row and walks along its two coefficient rows, one 16-byte piece (4 K values)
// of the hi and of the lo matrix per step; consecutive lanes write consecutive pieces.
template <int TAPS>
__global__ void __launch_bounds__(128)
tc_gmat_kernel(const UnitDev *units, const TileRec *tiles, const PlanEntry *entries,
               const float *coeffs, float *gmat, TcTile *tct, uint32_t kt_max) {
    const UnitDev &U = units[0];
    const uint32_t t = blockIdx.x;
    if (t >= U.n_tiles) return;
    const size_t tile = (size_t)U.tile_off + t;
    const TileRec rec = tiles[tile];
    const uint32_t tid = threadIdx.x;
    const uint32_t half = kt_max * kN;                   // floats in the hi (or lo) matrix
    const PlanEntry e0 = entries[tile * kTileOut];
    const PlanEntry el = entries[tile * kTileOut + rec.n_out - 1];
    // K origin: the first needed frame rounded down to a multiple of 8 counted from the first
    // frame of the new input (virtual frame H), the anchor of the input chunk grid
    const int32_t H = (int32_t)U.hist_len0;
    const int32_t k0 = e0.v - ((((e0.v - H) % 8) + 8) % 8);
    uint32_t kt = (uint32_t)(el.v + TAPS - k0 + 7) & ~7u;
    if (kt > kt_max) kt = kt_max;    // cannot happen for a plan of this ratio (tc_kt_max)
    if (tid == 0) {
        TcTile m;
        m.k0 = k0;
        m.kt = kt;
        m.n_out = rec.n_out;
        m.o_start = rec.o_start;
        tct[tile] = m;
    }
    const uint32_t o = tid & 31u;                        // output row of this thread
    const bool row_ok = o < rec.n_out;
    PlanEntry e = e0;
    if (row_ok) e = entries[tile * kTileOut + o];
    const uint32_t p1 = e.phase1;
    const uint32_t p2 = p1 + 1 < kPhases - 1 ? p1 + 1 : kPhases - 1;
    const float fr = e.frac, omf = __fsub_rn(1.0f, fr);
    const int d = e.v - k0;                              // K index of the row's tap 0
    const float *ca = coeffs + (size_t)p1 * TAPS;
    const float *cb = coeffs + (size_t)p2 * TAPS;
    float4 *dst_hi = reinterpret_cast<float4 *>(gmat + tile * (size_t)2 * half);
    float4 *dst_lo = dst_hi + half / 4;
    for (uint32_t kg = tid >> 5; kg < kt / 4; kg += 4) {
        float hi[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
        const int j0 = (int)(4 * kg) - d;
        if (row_ok && j0 > -4 && j0 < TAPS) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j0 + q;
                if (j >= 0 && j < TAPS) {
                    const float g = __fmaf_rn(__ldg(cb + j), fr, __fmul_rn(__ldg(ca + j), omf));
                    hi[q] = to_tf32(g);
                    lo[q] = to_tf32(__fsub_rn(g, hi[q]));
                }
            }
        }
        dst_hi[kg * kN + o] = make_float4(hi[0], hi[1], hi[2], hi[3]);
        dst_lo[kg * kN + o] = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

}  // namespace

bool tc_supported(uint32_t channels, uint32_t taps, double ratio) {
    if (channels != 1 && channels != 2 && channels != 4 && channels != 8) return false;
    if (taps != 16 && taps != 32 && taps != 64 && taps != 128) return false;
    return tc_kt_max(taps, ratio) <= kKtLimit;
}

uint32_t tc_kt_extent(uint32_t taps, double ratio) { return tc_kt_max(taps, ratio); }

size_t tc_gmat_floats_per_tile(uint32_t taps, double ratio) {
    return (size_t)2 * tc_kt_max(taps, ratio) * kN;
}

bool tc_make_input_tensor_map(CUtensorMap *out, const float *base, uint64_t stride_bytes,
                              uint64_t total_frames, uint32_t n_members, uint32_t channels) {
    typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                      const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                      const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) !=
                cudaSuccess || !fn) {
            cudaGetLastError();
            return false;
        }
        encode = (EncodeTiledFn)fn;
    }
    if (channels != 1 && channels != 2 && channels != 4 && channels != 8) return false;
    if (total_frames == 0 || total_frames * channels >= (1ull << 31)) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (stride_bytes & 15u) || stride_bytes == 0)
        return false;
    if (stride_bytes < total_frames * channels * 4ull && !getenv("RSB_DEBUG_FAKE_IN_STRIDE")) return false;
    // mono / stereo: element = one frame (f32, or an 8-byte stereo frame), 64B / 128B swizzle (the
    // splitter reads 16-byte units, one row per lane).  4 / 8 channels: element = one float, inner
    // dimension = frames x channels, no swizzle (the splitter reads single floats).
    const bool wide = channels >= 4;
    cuuint64_t dims[2] = {wide ? total_frames * channels : total_frames, n_members};
    cuuint64_t strides[1] = {stride_bytes};
    cuuint32_t box[2] = {wide ? kChunk * channels : kChunk, (cuuint32_t)(kRows / channels)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt =
        channels == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const CUtensorMapSwizzle sw = channels == 1   ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : channels == 2 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = encode(out, dt, 2, const_cast<float *>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool tc_make_raw16_tensor_map(CUtensorMap *out, const void *base, uint64_t stride_bytes,
                              uint64_t total_frames, uint32_t n_members, uint32_t src_channels,
                              uint32_t channels, uint32_t sample_bytes) {
    typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                      const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                      const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        !fn) {
        cudaGetLastError();
        return false;
    }
    if (!((src_channels == 2 && channels == 2) || (src_channels == 1 && channels <= 2))) return false;
    if (total_frames == 0 || total_frames >= (1ull << 31)) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (stride_bytes & 15u) || stride_bytes == 0)
        return false;
    if (sample_bytes != 2 && sample_bytes != 3) return false;
    if (stride_bytes < total_frames * (uint64_t)sample_bytes * src_channels) return false;
    cuuint64_t dims[2] = {total_frames, n_members};
    cuuint64_t strides[1] = {stride_bytes};
    cuuint32_t box[2] = {kChunk, (cuuint32_t)(kRows / channels)};
    cuuint32_t estr[2] = {1, 1};
    if (sample_bytes == 3) {
        // packed s24: element = one byte, a frame is 3 * src_channels of them, no swizzle
        const uint32_t fb = 3u * src_channels;
        dims[0] = total_frames * fb;
        box[0] = kChunk * fb;
        const CUresult r3 = ((EncodeTiledFn)fn)(
            out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        return r3 == CUDA_SUCCESS;
    }
    // element = one raw frame: two s16 (32 bits, 64-byte rows, 64B swizzle) or one s16 (32-byte
    // rows, no swizzle); frames past the end read as zeros
    const bool stereo = src_channels == 2;
    const CUresult r = ((EncodeTiledFn)fn)(
        out, stereo ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2,
        const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        stereo ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

void launch_tc_gmat(const UnitDev *units, const TileRec *tiles, const PlanEntry *entries,
                    const float *coeffs, float *gmat, TcTile *tct, uint32_t taps, double ratio,
                    uint32_t tile_cap, cudaStream_t stream) {
    if (tile_cap == 0) return;
    const uint32_t kt_max = tc_kt_max(taps, ratio);
    switch (taps) {
        case 16: tc_gmat_kernel<16><<<tile_cap, 128, 0, stream>>>(units, tiles, entries, coeffs, gmat, tct, kt_max); break;
        case 32: tc_gmat_kernel<32><<<tile_cap, 128, 0, stream>>>(units, tiles, entries, coeffs, gmat, tct, kt_max); break;
        case 64: tc_gmat_kernel<64><<<tile_cap, 128, 0, stream>>>(units, tiles, entries, coeffs, gmat, tct, kt_max); break;
        default: tc_gmat_kernel<128><<<tile_cap, 128, 0, stream>>>(units, tiles, entries, coeffs, gmat, tct, kt_max); break;
    }
}

void launch_conv_tc(const TcParams &p, const CUtensorMap &tmap, int sm_count, bool leave_sm_free,
                    cudaStream_t stream) {
    const size_t smem = (size_t)kXStages * kXStageBytes + (size_t)kGStages * 2 * p.kt_max * 128u +
                        (size_t)2 * (kRows / p.channels) * tc_stage_pitch(p.channels) * sizeof(float);
    // one SM is left free when the next submit's (serial) plan kernel may need somewhere to run
    const uint32_t grid = (uint32_t)(leave_sm_free && sm_count > 8 ? sm_count - 1 : sm_count);
    auto launch = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, kTcThreads, smem, stream>>>(p, tmap);
    };
    switch (p.channels) {
        case 1:
            if (p.raw16 && p.raw_bytes == 3) launch(conv_tc_kernel<1, 1, 3>);
            else if (p.raw16) launch(conv_tc_kernel<1, 1>);
            else launch(conv_tc_kernel<1>);
            break;
        case 2:
            if (p.raw16 == 2 && p.raw_bytes == 3) launch(conv_tc_kernel<2, 2, 3>);
            else if (p.raw16 == 1 && p.raw_bytes == 3) launch(conv_tc_kernel<2, 1, 3>);
            else if (p.raw16 == 2) launch(conv_tc_kernel<2, 2>);
            else if (p.raw16 == 1) launch(conv_tc_kernel<2, 1>);
            else launch(conv_tc_kernel<2>);
            break;
        case 4: launch(conv_tc_kernel<4>); break;
        default: launch(conv_tc_kernel<8>); break;
    }
}

uint32_t tc_rows_per_group() { return kRows; }

// Two issuers keep two consecutive tiles in flight: their K ranges (the second starts up to
// floor(32*ratio)+1 frames later), the 8-frame alignment slack and one chunk of granularity must
// fit the ring together.
uint32_t tc_issuers(uint32_t taps, double ratio) {
    const uint32_t adv = (uint32_t)(32.0 * ratio) + 1u;
    return tc_kt_max(taps, ratio) + 8u + adv + (kChunk - 1u) <= kRing ? 2u : 1u;
}

void tc_phase_profile(int enable, unsigned long long *out16) {
    if (out16) cudaMemcpyFromSymbol(out16, g_tc_cycles, sizeof(unsigned long long) * 24);
    unsigned long long zero[24] = {0};
    cudaMemcpyToSymbol(g_tc_cycles, zero, sizeof(zero));
    cudaMemcpyToSymbol(g_tc_prof, &enable, sizeof(int));
}

}  // namespace rsb
