// pcm_ingest.h -- launch interface of the PCM format step (pcm_ingest.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rsb {

// One stream's raw samples and the f32 buffer they are converted into.
struct PcmJob {
    const void *src;       // raw samples (device), format-sized, little endian
    float *dst;            // 16-byte aligned, n_out values
    uint64_t n_out;        // f32 values to write = source values * dup
    uint32_t src_aligned;  // src is 16-byte aligned: vector loads
    uint32_t pad_;
};

uint32_t pcm_bytes_per_sample(int format);   // 0 for an unknown format
uint64_t pcm_chunks(uint64_t n_out);         // 32 KB output chunks of one job
// chunk_first[j] = first chunk of job j, chunk_first[n_jobs] = n_chunks (device array);
// chunks_per_job != 0: every job has exactly that many chunks (no search in the kernel)
bool launch_pcm_ingest(const PcmJob *jobs, const uint64_t *chunk_first, uint32_t n_jobs,
                       uint64_t n_chunks, uint64_t chunks_per_job, int format, uint32_t dup,
                       int sm_count, cudaStream_t stream);

}  // namespace rsb
