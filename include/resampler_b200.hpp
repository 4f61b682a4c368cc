// resampler_b200.hpp -- C++ host-side mirror of the reference's ResamplerFir interface over the
// C ABI (resampler_b200.h).  The reference is Rust (no rustc in this image), so the host side
// above the C ABI is C++; names, argument meaning and error behaviour follow
// hasenbanck/resampler v0.5.1 src/resampler_fir.rs:252, 295, 456, 509, 630, 638.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "resampler_b200.h"

namespace resampler_b200 {

// src/lib.rs:166-254
enum class SampleRate : uint32_t {
    Hz16000 = 16000, Hz22050 = 22050, Hz32000 = 32000, Hz44100 = 44100, Hz48000 = 48000,
    Hz88200 = 88200, Hz96000 = 96000, Hz176400 = 176400, Hz192000 = 192000, Hz384000 = 384000
};
// src/resampler_fir.rs:138-162 (default Sample64)
enum class Latency : int { Sample8 = 0, Sample16 = 1, Sample32 = 2, Sample64 = 3 };
inline constexpr size_t taps(Latency l) { return size_t(16) << static_cast<int>(l); }
// src/resampler_fir.rs:101-123 (default Db120)
enum class Attenuation : int { Db60 = 0, Db90 = 1, Db120 = 2 };

// src/error.rs:1-26
struct ResampleError : std::runtime_error {
    enum Kind { InvalidInputBufferSize = 1, InvalidOutputBufferSize = 2 } kind;
    explicit ResampleError(Kind k)
        : std::runtime_error(k == InvalidInputBufferSize ? "Input buffer size is invalid"
                                                         : "Output buffer size is invalid"),
          kind(k) {}
};

inline void check(int rc) {
    if (rc == RSB_OK) return;
    if (rc == RSB_ERR_INVALID_INPUT_BUFFER_SIZE) throw ResampleError(ResampleError::InvalidInputBufferSize);
    if (rc == RSB_ERR_INVALID_OUTPUT_BUFFER_SIZE) throw ResampleError(ResampleError::InvalidOutputBufferSize);
    // the reference panics with these exact texts (resampler_fir.rs:302-309)
    if (rc == RSB_ERR_ZERO_INPUT_RATE || rc == RSB_ERR_ZERO_OUTPUT_RATE)
        throw std::invalid_argument(rsb_status_string(rc));
    throw std::runtime_error(std::string("resampler_b200: ") + rsb_last_error());
}

// N independent streams with identical parameters on one GPU (the batched entry that is new).
class FirBatch {
  public:
    FirBatch(uint32_t n_streams, uint32_t channels, uint32_t in_hz, uint32_t out_hz,
             Latency latency = Latency::Sample64, Attenuation attenuation = Attenuation::Db120,
             int device = 0) {
        check(rsb_fir_create(&h_, device, n_streams, channels, in_hz, out_hz,
                             static_cast<int>(latency), static_cast<int>(attenuation)));
    }
    ~FirBatch() { rsb_fir_destroy(h_); }
    FirBatch(const FirBatch &) = delete;
    FirBatch &operator=(const FirBatch &) = delete;

    size_t buffer_size_output() const { return rsb_fir_buffer_size_output(h_); }
    size_t delay() const { return rsb_fir_delay(h_); }
    void reset(int64_t stream = -1) { check(rsb_fir_reset(h_, stream)); }

    // one resample() call per listed stream; returns (consumed, produced) per job
    void submit(const std::vector<const float *> &in, const std::vector<size_t> &in_lens,
                const std::vector<float *> &out, const std::vector<size_t> &out_lens,
                std::vector<size_t> &consumed, std::vector<size_t> &produced,
                int memspace = RSB_MEM_HOST, const uint32_t *streams = nullptr) {
        consumed.resize(in.size());
        produced.resize(in.size());
        check(rsb_fir_submit_batch(h_, (uint32_t)in.size(), streams, in.data(), in_lens.data(),
                                   out.data(), out_lens.data(), consumed.data(), produced.data(),
                                   memspace, 0));
    }
    // the canonical caller loop per stream (resample/src/main.rs:226-254) in one launch
    void process(const std::vector<const float *> &in, const std::vector<size_t> &total_lens,
                 size_t call_len, const std::vector<float *> &out,
                 const std::vector<size_t> &out_capacities, std::vector<size_t> &consumed,
                 std::vector<size_t> &produced, int memspace = RSB_MEM_HOST, uint32_t flags = 0) {
        consumed.resize(in.size());
        produced.resize(in.size());
        check(rsb_fir_process_batch(h_, (uint32_t)in.size(), nullptr, in.data(), total_lens.data(),
                                    call_len, 0, out.data(), out_capacities.data(), consumed.data(),
                                    produced.data(), nullptr, memspace, flags));
    }
    // The CLI's batch path (resample/src/main.rs:128-156 + 226-254): raw samples in, the
    // format step and the canonical loop on the GPU.  in_frames are source frames.
    void process_pcm(const std::vector<const void *> &in, const std::vector<size_t> &in_frames,
                     rsb_pcm_format format, uint32_t src_channels, size_t call_len,
                     const std::vector<float *> &out, const std::vector<size_t> &out_capacities,
                     std::vector<size_t> &consumed, std::vector<size_t> &produced,
                     int memspace = RSB_MEM_HOST) {
        consumed.assign(in.size(), 0);
        produced.assign(in.size(), 0);
        check(rsb_fir_process_pcm_batch(h_, (uint32_t)in.size(), nullptr, in.data(), in_frames.data(),
                                        format, src_channels, call_len, 0, out.data(),
                                        out_capacities.data(), consumed.data(), produced.data(),
                                        nullptr, memspace, 0));
    }
    // Opt-in tail handling (not in the reference): delay() frames of silence per stream.
    void flush(const std::vector<float *> &out, const std::vector<size_t> &out_capacities,
               std::vector<size_t> &produced, int memspace = RSB_MEM_HOST) {
        produced.assign(out.size(), 0);
        check(rsb_fir_flush_batch(h_, (uint32_t)out.size(), nullptr, out.data(), out_capacities.data(),
                                  produced.data(), memspace, 0));
    }
    void sync() { check(rsb_fir_sync(h_)); }
    rsb_fir *raw() { return h_; }

  private:
    rsb_fir *h_ = nullptr;
};

// Drop-in mirror of the reference's single-stream type.
class ResamplerFir {
  public:
    // ResamplerFir::new (resampler_fir.rs:252-266)
    ResamplerFir(size_t channels, SampleRate in, SampleRate out, Latency latency = Latency::Sample64,
                 Attenuation attenuation = Attenuation::Db120, int device = 0)
        : b_(1, (uint32_t)channels, static_cast<uint32_t>(in), static_cast<uint32_t>(out), latency,
             attenuation, device) {}
    // ResamplerFir::new_from_hz (resampler_fir.rs:295-404); zero rates throw std::invalid_argument
    // with the reference's panic text
    static ResamplerFir new_from_hz(size_t channels, uint32_t in_hz, uint32_t out_hz,
                                    Latency latency = Latency::Sample64,
                                    Attenuation attenuation = Attenuation::Db120, int device = 0) {
        return ResamplerFir(channels, in_hz, out_hz, latency, attenuation, device, 0);
    }
    size_t buffer_size_output() const { return b_.buffer_size_output(); }   // :456-465
    size_t delay() const { return b_.delay(); }                            // :630-632
    void reset() { b_.reset(0); }                                          // :638-642
    // resample(&input, &mut output) -> (consumed, produced)   (resampler_fir.rs:509-621)
    std::pair<size_t, size_t> resample(const float *input, size_t input_len, float *output,
                                       size_t output_len) {
        size_t c = 0, p = 0;
        check(rsb_fir_resample(b_.raw(), 0, input, input_len, output, output_len, &c, &p));
        return {c, p};
    }

  private:
    ResamplerFir(size_t channels, uint32_t in_hz, uint32_t out_hz, Latency l, Attenuation a, int dev,
                 int)
        : b_(1, (uint32_t)channels, in_hz, out_hz, l, a, dev) {}
    FirBatch b_;
};

}  // namespace resampler_b200
