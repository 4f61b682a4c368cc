// fir_submit.h -- launchers of the small / divergent submit path (fir_submit.cu)
#pragma once
#include "fir_kernels.h"

namespace rsb {

bool submit_two_stage(uint32_t taps, uint32_t channels, uint32_t max_in_frames);
void launch_submit_plan(const SubmitJob *jobs, SubmitResult *results, uint32_t n_jobs, StreamStateDev st, double ratio,
                        uint32_t taps, PlanSeg *seg_store, cudaStream_t stream);
void launch_submit_conv(const SubmitJob *jobs, const SubmitResult *results, uint32_t n_jobs, const float *coeffs,
                        uint32_t taps, uint32_t channels, uint32_t max_in_frames, const PlanSeg *seg_store,
                        uint32_t in_hz, uint32_t out_hz, cudaStream_t stream);

}  // namespace rsb
