// filter_design.cpp -- see filter_design.h.  Built with -ffp-contract=off: the
// reference computes the sinc, product, running sum and normalisation in f32
// with one rounding per operation (src/window.rs:29-52).
#include "filter_design.h"
#include "sinf_glibc.h"

#include <cmath>
#include <map>
#include <mutex>
#include <tuple>

namespace rsb {

float sinf_restated(float x) { return sinf_glibc(x); }

int latency_to_taps(int latency) {
    static const int kTaps[4] = {16, 32, 64, 128};
    return (latency >= 0 && latency < 4) ? kTaps[latency] : -1;
}

double attenuation_to_beta(int attenuation) {
    static const double kBeta[3] = {7.0, 10.0, 13.0};
    return (attenuation >= 0 && attenuation < 3) ? kBeta[attenuation] : -1.0;
}

// I0(x) = sum_k ((x/2)^(2k)) / (k!)^2, summed until the partial sum is stationary
// (src/window.rs:96-112: at most 1499 terms, same term recurrence and order).
double bessel_i0(double x) {
    const double q = x * x / 4.0;
    double t = 1.0, acc = 1.0;
    for (int k = 1; k < 1500; ++k) {
        t = t * q / static_cast<double>(k * k);
        const double before = acc;
        acc += t;
        if (acc == before) break;
    }
    return acc;
}

double kaiser_cutoff(uint32_t taps, double beta) {
    const double n = static_cast<double>(taps);
    const double a_db = beta / 0.1102 + 8.7;
    const double width = (a_db - 7.95) / (14.36 * n);
    double c = 1.0 - (width * 1.005);
    return c < 0.7 ? 0.7 : (c > 1.0 ? 1.0 : c);
}

float design_cutoff(uint32_t in_hz, uint32_t out_hz, uint32_t taps, double beta) {
    const double fi = static_cast<double>(in_hz), fo = static_cast<double>(out_hz);
    const double base = kaiser_cutoff(taps, beta);
    // up-sampling keeps the full input band; down-sampling scales to the output Nyquist
    const double c = (fi <= fo) ? base : base * (fo / fi);
    return static_cast<float>(c);
}

void make_kaiser_window(uint32_t n, double beta, bool symmetric, float *out) {
    const double i0b = bessel_i0(beta);
    for (uint32_t i = 0; i < n; ++i) {
        const double x = static_cast<double>(i);
        const double t = symmetric ? 2.0 * x / static_cast<double>(n - 1) - 1.0
                                   : x / (static_cast<double>(n) / 2.0) - 1.0;
        out[i] = static_cast<float>(bessel_i0(beta * std::sqrt(1.0 - t * t)) / i0b);
    }
}

void make_sincs_for_kaiser(uint32_t sample_count, uint32_t factor, float cutoff, double beta,
                           bool symmetric, float *out) {
    const uint32_t total = sample_count * factor;
    std::vector<float> win(total), proto(total);
    make_kaiser_window(total, beta, symmetric, win.data());
    const float pi32 = 3.14159274101257324219f;
    const int32_t half = static_cast<int32_t>(total / 2);
    volatile float running = 0.0f;   // strictly sequential f32 sum (src/window.rs:27,41)
    for (uint32_t i = 0; i < total; ++i) {
        const float a = static_cast<float>(static_cast<int32_t>(i) - half) * cutoff /
                        static_cast<float>(factor);
        float s = 1.0f;
        if (a != 0.0f) {
            const float ap = a * pi32;
            s = sinf(ap) / ap;
        }
        const float v = win[i] * s;
        running = running + v;
        proto[i] = v;
    }
    float norm = running;
    norm /= static_cast<float>(factor);
    // phase order is reversed: row (factor-1-n) takes prototype samples n, n+factor, ...
    for (uint32_t ph = 0; ph < factor; ++ph) {
        const uint32_t n = factor - 1 - ph;
        float *row = out + static_cast<size_t>(ph) * sample_count;
        for (uint32_t p = 0; p < sample_count; ++p) row[p] = proto[factor * p + n] / norm;
    }
}

std::shared_ptr<const FirTable> get_or_create_table(float cutoff, uint32_t taps, int attenuation,
                                                    TableBuilder builder, void *ctx) {
    static std::mutex mu;
    static std::map<std::tuple<uint32_t, uint32_t, int>, std::shared_ptr<const FirTable>> cache;
    uint32_t bits;
    static_assert(sizeof(bits) == sizeof(cutoff), "f32");
    __builtin_memcpy(&bits, &cutoff, 4);
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple(bits, taps, attenuation);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    auto t = std::make_shared<FirTable>();
    t->taps = taps;
    t->cutoff_bits = bits;
    t->attenuation = attenuation;
    t->coeffs.resize(static_cast<size_t>(1024) * taps);
    if (!builder || !builder(cutoff, taps, attenuation_to_beta(attenuation), t->coeffs.data(), ctx))
        make_sincs_for_kaiser(taps, 1024, cutoff, attenuation_to_beta(attenuation), true,
                              t->coeffs.data());
    cache.emplace(key, t);
    return t;
}

}  // namespace rsb
