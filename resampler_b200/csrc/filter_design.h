// filter_design.h -- host-side Kaiser-windowed-sinc polyphase table.
// Construct-time only; stays on the host because the table's bits depend on the
// platform sinf (reference: f32::sin -> libm sinf, src/window.rs:33).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

namespace rsb {

struct FirTable {
    uint32_t taps = 0;
    uint32_t cutoff_bits = 0;
    int attenuation = 0;
    std::vector<float> coeffs;   // [1024][taps] row-major (src/resampler_fir.rs:406-421)
};

int latency_to_taps(int latency);          // src/resampler_fir.rs:153-161; -1 if invalid
double attenuation_to_beta(int attenuation);   // src/resampler_fir.rs:117-123; <0 if invalid
double kaiser_cutoff(uint32_t taps, double beta);   // src/window.rs:114-131
float design_cutoff(uint32_t in_hz, uint32_t out_hz, uint32_t taps, double beta);  // :311-326
double bessel_i0(double x);                                                        // window.rs:96-112

// Process-wide cache keyed by (cutoff bits, taps, attenuation) like the reference's
// FIR_CACHE (src/resampler_fir.rs:91-95, 164-166, 425-443).
// `builder` (optional) designs the table on a cache miss instead of the host code (the device-side
// design of filter_design_device.cu); it returns false to fall back to the host design.
typedef bool (*TableBuilder)(float cutoff, uint32_t taps, double beta, float *out, void *ctx);
std::shared_ptr<const FirTable> get_or_create_table(float cutoff, uint32_t taps, int attenuation,
                                                    TableBuilder builder = nullptr, void *ctx = nullptr);

// Uncached general builder (src/window.rs:17-55), exposed for the known-answer tests.
void make_sincs_for_kaiser(uint32_t sample_count, uint32_t factor, float cutoff, double beta,
                           bool symmetric, float *out /*[factor][sample_count]*/);
void make_kaiser_window(uint32_t n, double beta, bool symmetric, float *out);

// Host twin of the device-side design's sine (sinf_glibc.h): the CPU tests hold it against libm's sinf.
float sinf_restated(float x);

}  // namespace rsb
