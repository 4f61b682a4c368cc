// filter_design_device.h -- the polyphase table of filter_design.h designed ON the GPU (SURVEY.md 8(f)
// row 2), bit-identical to the host design (src/window.rs:17-131).
#pragma once
#include <cstdint>

namespace rsb {

// Designs the [1024][taps] table for (cutoff, taps, beta) on `device` and copies it to `host_out`
// (1024 * taps floats).  `elapsed_ms` (optional) receives the device time of the three kernels.
// Returns false on a CUDA error.
bool design_table_on_device(int device, float cutoff, uint32_t taps, double beta, float *host_out,
                            float *elapsed_ms);

}  // namespace rsb
