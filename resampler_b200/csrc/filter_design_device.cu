// filter_design_device.cu -- Kaiser-windowed-sinc polyphase table on the GPU (SURVEY.md 8(f) row 2;
// reference: make_sincs_for_kaiser / make_kaiser_window / bessel_i0, src/window.rs:17-112).
//
// The table's bits are observable (the EXACT kernel reproduces the reference's samples bit for
// bit), so every operation is the reference's, with explicit round-to-nearest intrinsics:
//   * window: f64 Bessel series with the reference's term recurrence and stopping rule, f64 sqrt and
//     division, one rounding to f32 (IEEE basic operations: identical on any conforming machine);
//   * sinc: the reference calls f32::sin = the platform's sinf.  The device code restates glibc's
//     sinf (sysdeps/ieee754/flt-32/s_sinf.c, the FMA build x86-64 hosts with FMA select): argument
//     reduction and polynomial in f64 with fused multiply-adds where that build has them, the
//     2/pi-table reduction for |x| >= 120 (the outer taps of a 128-tap table reach |x| = 201).
//     Checked on this image against libm's sinf: 0 differences over 2e8 random arguments in
//     [-260, 260] and over every argument of a 128-tap table (a host WITHOUT FMA would select glibc's
//     other build, which differs in ~1 of 1e7 arguments by one ulp);
//   * the normalisation sum is the reference's strictly sequential f32 sum over all 1024 * taps
//     products (src/window.rs:27, 41): one thread adds them in order (0.3 ms for 131 072 terms: the
//     reason a device design cannot be much faster than the host's 1.35 ms, and why both are
//     cached per (cutoff bits, taps, attenuation) like the reference's FIR_CACHE).
#include <cuda_runtime.h>

#include <cstring>

#include "filter_design.h"
#include "filter_design_device.h"
#include "sinf_glibc.h"

namespace rsb {

namespace {

__device__ double bessel_i0_dev(double x) {
    const double q = __ddiv_rn(__dmul_rn(x, x), 4.0);
    double t = 1.0, acc = 1.0;
    for (int k = 1; k < 1500; ++k) {
        t = __ddiv_rn(__dmul_rn(t, q), (double)(k * k));
        const double before = acc;
        acc = __dadd_rn(acc, t);
        if (acc == before) break;
    }
    return acc;
}

// prototype filter: window (symmetric Kaiser) x sinc, src/window.rs:38-43
__global__ void proto_kernel(float *proto, uint32_t total, uint32_t factor, float cutoff, double beta, double i0_beta) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double x = (double)i;
    const double t = __dsub_rn(__ddiv_rn(__dmul_rn(2.0, x), (double)(total - 1)), 1.0);
    const double arg = __dmul_rn(beta, __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(t, t))));
    const float w = __double2float_rn(__ddiv_rn(bessel_i0_dev(arg), i0_beta));
    const float a = __fdiv_rn(__fmul_rn((float)((int32_t)i - (int32_t)(total / 2)), cutoff), (float)factor);
    float s = 1.0f;
    if (a != 0.0f) {
        const float ap = __fmul_rn(a, 3.14159274101257324219f);
        s = __fdiv_rn(sinf_glibc(ap), ap);
    }
    proto[i] = __fmul_rn(w, s);
}

// the reference's strictly sequential f32 sum (one warp stages, lane 0 adds in order)
__global__ void seq_sum_kernel(const float *proto, uint32_t total, uint32_t factor, float *norm) {
    __shared__ __align__(16) float buf[2048];
    const uint32_t lane = threadIdx.x;
    float run = 0.0f;
    for (uint32_t base = 0; base < total; base += 2048) {
        for (uint32_t j = lane; j < 2048; j += 32) buf[j] = base + j < total ? proto[base + j] : 0.0f;
        __syncwarp();
        if (lane == 0) {
            const uint32_t n = min(2048u, total - base);
            uint32_t j = 0;
            // the adds are one dependent chain (4 cycles each); the loads run ahead of it
            for (; j + 8 <= n; j += 8) {
                const float4 a = *reinterpret_cast<const float4 *>(&buf[j]);
                const float4 b = *reinterpret_cast<const float4 *>(&buf[j + 4]);
                run = __fadd_rn(run, a.x); run = __fadd_rn(run, a.y); run = __fadd_rn(run, a.z); run = __fadd_rn(run, a.w);
                run = __fadd_rn(run, b.x); run = __fadd_rn(run, b.y); run = __fadd_rn(run, b.z); run = __fadd_rn(run, b.w);
            }
            for (; j < n; ++j) run = __fadd_rn(run, buf[j]);
        }
        __syncwarp();
    }
    if (lane == 0) *norm = __fdiv_rn(run, (float)factor);
}

// phase rows: row (factor - 1 - n) takes prototype samples n, n + factor, ..., normalised (:47-52)
__global__ void table_kernel(const float *proto, const float *norm, float *table, uint32_t taps, uint32_t factor) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= taps * factor) return;
    const uint32_t ph = i / taps, p = i - ph * taps;
    table[i] = __fdiv_rn(proto[factor * p + (factor - 1 - ph)], *norm);
}

}  // namespace

bool design_table_on_device(int device, float cutoff, uint32_t taps, double beta, float *host_out,
                            float *elapsed_ms) {
    if (cudaSetDevice(device) != cudaSuccess) return false;
    const uint32_t factor = 1024, total = taps * factor;
    float *d_proto = nullptr, *d_table = nullptr, *d_norm = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool ok = cudaMalloc(&d_proto, sizeof(float) * total) == cudaSuccess &&
              cudaMalloc(&d_table, sizeof(float) * total) == cudaSuccess &&
              cudaMalloc(&d_norm, sizeof(float)) == cudaSuccess && cudaEventCreate(&e0) == cudaSuccess &&
              cudaEventCreate(&e1) == cudaSuccess;
    if (ok) {
        const double i0_beta = bessel_i0(beta);          // host f64, same IEEE operations
        cudaEventRecord(e0);
        proto_kernel<<<(total + 255) / 256, 256>>>(d_proto, total, factor, cutoff, beta, i0_beta);
        seq_sum_kernel<<<1, 32>>>(d_proto, total, factor, d_norm);
        table_kernel<<<(total + 255) / 256, 256>>>(d_proto, d_norm, d_table, taps, factor);
        cudaEventRecord(e1);
        ok = cudaMemcpy(host_out, d_table, sizeof(float) * total, cudaMemcpyDeviceToHost) == cudaSuccess &&
             cudaGetLastError() == cudaSuccess;
        if (ok && elapsed_ms) cudaEventElapsedTime(elapsed_ms, e0, e1);
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d_proto);
    cudaFree(d_table);
    cudaFree(d_norm);
    if (!ok) cudaGetLastError();
    return ok;
}

}  // namespace rsb
