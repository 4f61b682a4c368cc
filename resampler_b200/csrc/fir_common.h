// fir_common.h -- data layout shared by the host library and the sm_100a kernels.
//
// Virtual frame coordinates.  Within one submit, a stream's input is the
// concatenation  [ history (H frames, device state) | new input (caller's buffer) ]
// and frame v of that concatenation is "virtual frame v".  The history is kept
// RIGHT-ALIGNED in a fixed buffer of kInputCapacity frames so that the seam
// (v == H) is 16-byte aligned in both sources; nothing is ever appended or
// compacted (the reference's double buffer, resampler_fir.rs:521-538, 604-615,
// is replaced by this zero-copy view).  A plan segment's `vbase` is the virtual
// frame the reference's read_position points at during that call.
#pragma once
#include <stdint.h>

#include "planner.h"

namespace rsb {

constexpr uint32_t kHistFrames = kInputCapacity;   // max frames carried between calls (:18)

// One job = one stream's work in a submit (device copy).  Jobs are stored grouped by plan
// unit: the members of unit U are jobs[U.member_off .. U.member_off + U.n_members).
struct JobDev {
    const float *in;          // interleaved new input (device), may be null when total_frames == 0
    float *out;               // interleaved output (device)
    const float *hist;        // this stream's live history buffer  [kHistFrames*channels]
    float *hist_next;         // the other history buffer (written by the state update)
    uint64_t out_capacity;    // frames that fit in `out`
    uint32_t stream;          // stream index inside the handle
    uint32_t unit;            // plan unit this job belongs to
};

// One plan unit = all jobs of a submit that share (state cohort, call signature):
// identical state + identical call sequence => identical plan (the plan does not
// depend on sample data).
struct UnitDev {
    // inputs
    uint64_t total_frames;    // frames offered in total
    uint32_t call_frames;     // frames per resample() call (canonical caller loop)
    uint32_t cap_frames;      // output capacity per call, frames
    uint32_t max_calls;       // 1 for a single resample() call
    uint32_t single_call;     // 1: exactly one call even if total_frames == 0
    uint32_t rep_stream;      // a member stream whose state is read
    uint32_t member_off;      // first entry in the member list
    uint32_t n_members;
    uint32_t seg_off, seg_cap;      // segment pool slice
    uint32_t call_off, call_cap;    // per-call count slice (call_cap == 0: not recorded)
    uint32_t tile_off, tile_cap;    // tile record slice
    // results
    uint64_t total_out;       // output frames produced
    uint64_t total_copied;    // input frames consumed ("copied", the API's `consumed`)
    uint32_t n_calls;
    uint32_t n_segs;
    uint32_t n_tiles;
    uint32_t status;          // 0 ok, 1 segment pool overflow, 2 call limit hit before input end
    uint32_t hist_len0;       // history frames at submit start
    uint32_t final_available;
    double final_position;
};

constexpr uint32_t kTileOut = 32;   // output frames per tile (both convolution kernels)

// Tile record: tile `t` of a unit covers outputs [o_start, o_start + n_out).
struct TileRec {
    uint32_t unit;
    uint32_t o_start;
    uint32_t seg;       // absolute index of the segment containing o_start
    uint32_t n_out;
    int32_t v_base;     // first virtual frame of the tile's input window; (v_base - H) % 4 == 0
    uint32_t winp;      // window length in frames, multiple of 4
    // copies of the unit's fields, so that a kernel needs one load instead of a dependent two
    uint32_t n_members, member_off;
    uint32_t hist_len0, pad0;
    uint64_t total_frames;
};

// Per-output-frame plan entry, expanded once per tile by the tile kernel and shared by every
// stream group that processes the tile: entries[tile * kTileOut + k].
struct PlanEntry {
    int32_t v;          // virtual input frame of the window start (vbase + floor(position))
    uint32_t phase1;
    float frac;
    uint32_t off;       // floor(position): the reference's input_offset inside its call
};

struct CallCounts {
    uint32_t copied;    // frames
    uint32_t produced;  // frames
};

// Persistent per-stream state (device).
struct StreamStateDev {
    double *position;        // [n_streams]
    uint32_t *hist_len;      // [n_streams]  == available_frames between calls
    float *hist[2];          // [n_streams][kHistFrames*channels], right-aligned; which of the
                             // two is live for a stream is tracked by the host (it flips at
                             // every submit the stream takes part in)
};

}  // namespace rsb
