// fir_kernels.h -- launch interface between the host library (fir_api.cu) and
// the sm_100a kernels (fir_kernels.cu, fir_fast.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fir_common.h"

namespace rsb {

struct ConvParams {
    const UnitDev *units;
    const JobDev *jobs;           // grouped by unit (see JobDev)
    const PlanSeg *segs;
    const TileRec *tiles;
    const PlanEntry *entries;     // [tile][kTileOut]
    const float *gtiles;          // [tile][kTileOut][gs] banded filter rows (fast kernel)
    const uint32_t *tile_total;   // device counter written by the plan kernel
    uint32_t *work_counter;       // zeroed per submit: dynamic work distribution (fast kernel)
    const float *coeffs;          // [1024][taps]
    StreamStateDev st;
    uint32_t channels;
    uint32_t taps;
    uint32_t groups;              // member groups per tile (max over units)
    uint32_t streams_per_group;
    // Work item w -> (tile t, group g).  Items are ordered in super-blocks of `group_block`
    // groups: all tiles of groups [0, gb), then all tiles of [gb, 2gb), ...  Few streams are
    // in flight at a time (their buffers can be gigabytes apart: measured 1.6x slow-down at
    // 60 s x 1024 streams with every stream touched round-robin, TLB reach), while a tile's
    // filter matrix is still reused by gb consecutive items out of L2.
    uint32_t group_block;
    // Optional 2-D TMA tensor map over the new input of a single-unit batch whose member
    // buffers are equally strided in memory: rows = members, inner = frames.  One tensor
    // copy then replaces one bulk copy per member (which cost ~100 cycles each to issue).
    uint32_t tmap_valid;
};

// Builds the tensor map above (host).  Returns false when the layout does not qualify.
bool fast_make_input_tensor_map(CUtensorMap *out, const float *base, uint64_t stride_bytes,
                                uint64_t total_frames, uint32_t n_members, uint32_t channels,
                                uint32_t taps, double ratio);

constexpr uint32_t kExactStreamsPerGroup = 4;

// plan: one thread per unit walks its calls; also assigns tile slices
void launch_plan(UnitDev *units, uint32_t n_units, StreamStateDev st, double ratio, uint32_t taps,
                 PlanSeg *segs, CallCounts *calls, uint32_t tile_out, uint32_t *tile_total,
                 cudaStream_t stream);
// tile records: binary search of each tile's first segment
// (gtiles != nullptr: also builds the fast kernel's banded filter tiles [tile][kTileOut][gs])
void launch_tiles(const UnitDev *units, uint32_t n_units, const PlanSeg *segs, TileRec *tiles,
                  PlanEntry *entries, uint32_t max_tiles_per_unit, uint32_t taps,
                  const float *coeffs, float *gtiles, uint32_t gs, cudaStream_t stream);
// exact (AVX-512 order) convolution
void launch_conv_exact(const ConvParams &p, uint32_t max_items, int sm_count, cudaStream_t stream);
// fast convolution (fir_fast.cu); returns false when the configuration is unsupported
bool fast_supported(uint32_t channels, uint32_t taps, double ratio);
uint32_t fast_streams_per_group(uint32_t channels, uint32_t taps, double ratio);
uint32_t fast_row_stride(uint32_t channels, uint32_t taps, double ratio);   // floats per G row
void launch_conv_fast(const ConvParams &p, const CUtensorMap *tmap, double ratio,
                      uint32_t max_items, int sm_count, cudaStream_t stream);
void fast_set_warp_specialised(int on);
// debug: returns and clears the fast kernel's per-phase cycle counters, sets the enable flag
void fast_phase_profile(int enable, unsigned long long *out8);
// state write-back: new history tail, position, hist_len (one CTA per job)
void launch_update(const UnitDev *units, const JobDev *jobs, uint32_t n_jobs, StreamStateDev st,
                   uint32_t channels, cudaStream_t stream);
void launch_state_scalars(const UnitDev *units, const JobDev *jobs, uint32_t n_jobs,
                          StreamStateDev st, cudaStream_t stream);
void launch_reset(StreamStateDev st, uint32_t first, uint32_t count, cudaStream_t stream);
// expands a unit's plan into per-frame arrays with the kernels' own device code
void launch_expand_plan(const UnitDev *units, uint32_t unit, const PlanEntry *entries,
                        uint32_t n_frames, uint32_t *off, uint32_t *p1, uint32_t *p2, uint32_t *fb,
                        cudaStream_t stream);
void launch_fill_synthetic(float *dst, uint32_t first_stream, uint32_t n_streams, uint64_t frames,
                           uint32_t channels, uint32_t rate_hz, uint64_t seed, cudaStream_t stream);

// ---- tensor-core (tcgen05) convolution, fir_tensor.cu ----
// Per-tile record of the tensor kernel: the tile's K range in virtual frames.
struct TcTile {
    int32_t k0;        // first virtual frame of the tile's K range (multiple of 8)
    uint32_t kt;       // K extent in frames (multiple of 8)
    uint32_t n_out;
    uint32_t o_start;
};
struct TcParams {
    const UnitDev *units;     // exactly one plan unit
    const JobDev *jobs;
    const TcTile *tct;        // [tile]
    const float *gmat;        // [tile][hi, lo][kt_max / 4][32][4]
    uint32_t *work_counter;   // zeroed per submit
    uint32_t channels;
    uint32_t groups;          // member groups of 128 / channels streams
    uint32_t run_tiles;       // consecutive tiles per work item
    uint32_t kt_max;
    uint32_t issuers;         // MMA issuer warps: 2 when two consecutive tiles fit the TMEM ring
    uint32_t raw16;           // mono / stereo: the tensor map covers RAW s16 frames and the splitter
                              // converts them (v / 32768, main.rs:131-136): 1 = frames of `channels`
                              // samples, 2 = mono frames duplicated into both channels of a stereo stream
    uint32_t raw_bytes;       // bytes per raw sample: 2 (s16) or 3 (packed s24)
};
bool tc_supported(uint32_t channels, uint32_t taps, double ratio);
uint32_t tc_kt_extent(uint32_t taps, double ratio);
size_t tc_gmat_floats_per_tile(uint32_t taps, double ratio);
uint32_t tc_rows_per_group();
uint32_t tc_issuers(uint32_t taps, double ratio);
// 2-D tensor map over equally strided member inputs: box = 16 frames x (128 / channels) members
bool tc_make_input_tensor_map(CUtensorMap *out, const float *base, uint64_t stride_bytes,
                              uint64_t total_frames, uint32_t n_members, uint32_t channels);
// The same over raw interleaved s16 frames (element = one frame: 4 bytes stereo with 64B swizzle,
// 2 bytes mono unswizzled; box = 16 frames x 128 / channels members): the format step fused into
// the tensor kernel's loader.  src_channels 1 with channels 2 = mono source of a stereo stream.
bool tc_make_raw16_tensor_map(CUtensorMap *out, const void *base, uint64_t stride_bytes,
                              uint64_t total_frames, uint32_t n_members, uint32_t src_channels,
                              uint32_t channels, uint32_t sample_bytes);
void launch_tc_gmat(const UnitDev *units, const TileRec *tiles, const PlanEntry *entries,
                    const float *coeffs, float *gmat, TcTile *tct, uint32_t taps, double ratio,
                    uint32_t tile_cap, cudaStream_t stream);
void launch_conv_tc(const TcParams &p, const CUtensorMap &tmap, int sm_count, bool leave_sm_free,
                    cudaStream_t stream);
// debug: returns and clears the tensor kernel's per-role cycle counters, sets the enable flag
void tc_phase_profile(int enable, unsigned long long *out16);

// ---- tensor-core convolution, generation 2 (fir_tc2.cu): fp16 hi/lo operands, 64-output tiles ----
constexpr uint32_t kTc2TileOut = 64;   // output frames per tensor tile (two plan tiles)
struct Tc2Tile {
    int32_t k0;        // first virtual frame of the tile's K range (on the 16-frame chunk grid anchored at H)
    uint32_t kt;       // K extent in frames (multiple of 16)
    uint32_t n_out;
    uint32_t o_start;
    uint32_t g_idx;    // index of the tile's [G_hi, G_lo] matrices in the G store
    uint32_t pad[3];
};
struct Tc2Params {
    const UnitDev *units;     // exactly one plan unit
    const JobDev *jobs;
    const Tc2Tile *tct;       // [tile]
    const uint8_t *gmat;      // [stored tile][hi, lo][kt_max / 8][64][8 halves]
    uint32_t *work_counter;   // zeroed per submit
    uint32_t channels;
    uint32_t groups;          // member groups of 128 / channels streams
    uint32_t run_tiles;       // consecutive tiles per work item
    uint32_t kt_max;
    uint32_t issuers;         // MMA issuer warps: 2 when two consecutive tiles fit the TMEM ring
    uint32_t g_stages;        // shared-memory stages of G matrices (2 or 3)
    uint32_t raw16;           // as TcParams::raw16
    uint32_t raw_bytes;       // bytes per raw sample: 2 (s16) or 3 (packed s24)
    uint32_t prefetch_chunks; // input chunks the TMA producer prefetches into L2 ahead of its loads
    uint32_t variant;         // role layout (debug / tuning; 0 = default, see kVariants in fir_tc2.cu)
    uint32_t ablate;          // debug (RSB_TC_ABLATE): 1 no G copies, 2 no input TMA, 4 no output stores,
                              // 8 splitter: handshakes only, 16 epilogue: handshakes only, 32 one MMA per tile,
                              // 64 lo-pass MMAs at full width (no trimming to the active output quarters)
    float out_scale;          // epilogue factor 2^-17 (undoes the operand prescale)
    uint32_t epi_split;       // two epilogue teams: 1 = each drains one column half of every tile, 0 = alternate tiles
    uint32_t hint_crit;       // try_wait suspend-time hint (ns) of the issuers' and the epilogue's waits
    uint32_t hint_other;      // ... of every other role's waits
};
// Expected relative loss (in units of 2^-24) of the tensor core's truncating fp32 accumulation with the
// kernel's issue order, measured on B200 against an f64 evaluation for full-scale noise:
// 0.84 (44.1 -> 48 kHz), 0.87 (48 -> 44.1 kHz); tc2_gmat_kernel folds g * comp * 2^-24 into the G_lo matrices
// (tools/tc2_error_stats.py).
constexpr double kTc2TruncationComp = 0.85;
float tc2_out_scale();
bool tc2_supported(uint32_t channels, uint32_t taps, double ratio);
uint32_t tc2_kt_extent(uint32_t taps, double ratio);
size_t tc2_gmat_bytes_per_tile(uint32_t taps, double ratio);
uint32_t tc2_rows_per_group();
uint32_t tc2_issuers(uint32_t taps, double ratio);
uint32_t tc2_g_stages(uint32_t channels, uint32_t taps, double ratio, uint32_t variant = 0);
// Output of an equally strided batch as a 2-D tensor for the epilogue's TMA stores; inner
// dimension clipped to `valid_frames` (what the plan produces, capped by the smallest capacity).
bool tc2_make_output_tensor_map(CUtensorMap *out, float *base, uint64_t stride_bytes, uint64_t valid_frames,
                                uint32_t n_members, uint32_t channels);
void launch_tc2_tiles(const UnitDev *units, const PlanEntry *entries, Tc2Tile *tct, uint32_t taps, double ratio,
                      uint32_t tile_cap, cudaStream_t stream);
// src_tile / n_stored: optional de-duplication (stored matrix g is built from tile src_tile[g])
void launch_tc2_gmat(const UnitDev *units, const PlanEntry *entries, const float *coeffs, const Tc2Tile *tct,
                     const uint32_t *src_tile, const uint32_t *n_stored, uint8_t *gmat, uint32_t taps,
                     double ratio, uint32_t max_stored, double comp_units, cudaStream_t stream);
bool launch_conv_tc2(const Tc2Params &p, const CUtensorMap &tmap_in, const CUtensorMap &tmap_out, int sm_count,
                     bool leave_sm_free, cudaStream_t stream);
// debug: returns (up to `count` of 24) and clears the kernel's per-role cycle counters, sets the enable flag
void tc2_phase_profile(int enable, unsigned long long *out, uint32_t count);
// watchdog record of a tensor-kernel launch that trapped on a stuck mbarrier wait (zeros: none); clears it
void tc2_hang_record(unsigned int out[32]);

// ---- fused single-launch submit (fir_submit.cu): one resample() call per listed stream ----
struct SubmitJob {
    const float *in;          // interleaved new input (device)
    float *out;               // interleaved output (device)
    const float *hist;        // the stream's live history buffer
    float *hist_next;         // the other one (written)
    uint32_t in_frames;       // frames offered (<= kInputCapacity)
    uint32_t cap_frames;      // output capacity, frames
    uint32_t stream;
    uint32_t pad;
};
struct SubmitResult {
    double position;          // state after the call
    uint32_t copied;          // frames ("consumed", :617-620)
    uint32_t produced;        // frames
    uint32_t available;
    uint32_t status;          // 0 ok, 1 plan segment overflow
    uint32_t hist0;           // buffered frames BEFORE the call (the convolution kernel's window)
    uint32_t n_seg;           // plan segments of the call
};
constexpr uint32_t kSubmitSegs = 128;       // plan segments of ONE call (bound: segs_per_call_bound)
void launch_submit_fused(const SubmitJob *jobs, SubmitResult *results, uint32_t n_jobs, StreamStateDev st,
                         const float *coeffs, double ratio, uint32_t taps, uint32_t channels,
                         uint32_t max_in_frames, PlanSeg *seg_store, cudaStream_t stream);

}  // namespace rsb
