// microbench.cu -- device micro-benchmarks that justify the fast kernel's design
// ("measure an FFMA peak on the B200 first").  Diagnostics only;
// exported through rsb_microbench() and run by tools/microbench.py under gpurun.
//
//   id 0   register-only scalar FFMA peak            -> TFLOP/s
//   id 1   register-only packed FFMA2 (fma.rn.f32x2) -> TFLOP/s
//   id 10+ banded-GEMM inner-loop candidates with both operands in shared memory
//          (G = interpolated filter rows, X = staged input columns) -> TFLOP/s issued
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/resampler_b200.h"

namespace {

__device__ __forceinline__ void ffma2(float2 &d, const float2 a, const float2 b) {
    // d = a * b + d, two independent fp32 FMAs in one instruction (sm_100+)
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;"
                 : "+l"(reinterpret_cast<uint64_t &>(d))
                 : "l"(reinterpret_cast<const uint64_t &>(a)),
                   "l"(reinterpret_cast<const uint64_t &>(b)));
}

// ---- id 0: scalar FFMA peak ----
__global__ void __launch_bounds__(256) k_ffma_peak(float *out, int iters, float a0, float b0) {
    float acc[16];
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = a0 + i; b[i] = b0 + 0.5f * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = __fmaf_rn(a[i & 3], b[(i >> 2) & 3], acc[i]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- id 1: packed FFMA2 peak ----
__global__ void __launch_bounds__(256) k_ffma2_peak(float *out, int iters, float a0, float b0) {
    float2 acc[16];
    float2 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2((float)(threadIdx.x + i), (float)i);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] = make_float2(a0 + i, a0 - i);
        b[i] = make_float2(b0 + 0.5f * i, b0 - 0.25f * i);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) ffma2(acc[i], a[i & 3], b[(i >> 2) & 3]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- id 2: warp-level tensor-core mma.sync m16n8k8 TF32 (legacy path) peak ----
__global__ void __launch_bounds__(256) k_mma_tf32_peak(float *out, int iters) {
    float d[8][4];
    uint32_t a[4], b[2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) d[i][j] = (float)(threadIdx.x + i + j);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + 0.125f * i);
    b[0] = __float_as_uint(0.5f);
    b[1] = __float_as_uint(0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile(
                "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 "
                "{%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- id 10+: banded GEMM core candidates ----
// CTA tile: KT output rows x NC columns; window WIN (multiple of 4) shared-memory resident.
// G[KT][GS] (row stride GS words), X[NC][XS] planar columns.
// Thread tile: K rows x C columns; MAP 0: lanes = 32 column groups (G broadcast to the warp);
// MAP 1: lanes = 8 column groups x 4 row groups.
template <int K, int C, int MAP, bool PACKED>
__global__ void __launch_bounds__(256) k_core(float *out, int reps, int win) {
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
    constexpr int WARPS = 8;
    constexpr int ROWS_PER_WARP = MAP == 0 ? K : 4 * K;
    constexpr int COLS_PER_WARP = MAP == 0 ? 32 * C : 8 * C;
    // warps split rows first, then columns
    constexpr int KT = ROWS_PER_WARP * 2;                  // 2 row groups of warps
    constexpr int NC = COLS_PER_WARP * (WARPS / 2);        // 4 column groups of warps
    const int GS = win + 4;                                // row stride: == 4 (mod 32) when win % 32 == 0
    const int XS = win + 4;
    float *G = smem;
    float *X = smem + KT * GS;
    for (int i = threadIdx.x; i < KT * GS + NC * XS; i += blockDim.x)
        smem[i] = 1e-3f * (float)((i * 2654435761u) >> 24);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = warp & 1, wc = warp >> 1;
    // rows/columns of one thread are interleaved across the lanes so that a warp's float4
    // loads hit distinct bank quads (row and column strides are == 4 words mod 32)
    int row0, col0, rstep, cstep;
    if (MAP == 0) {
        row0 = wr * ROWS_PER_WARP;
        rstep = 1;
        col0 = wc * COLS_PER_WARP + lane;
        cstep = 32;
    } else {
        row0 = wr * ROWS_PER_WARP + (lane >> 3);
        rstep = 4;
        col0 = wc * COLS_PER_WARP + (lane & 7);
        cstep = 8;
    }
    float accs[K][C];
    float2 accp[K][C];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int c = 0; c < C; ++c) { accs[k][c] = 0.f; accp[k][c] = make_float2(0.f, 0.f); }

    for (int r = 0; r < reps; ++r) {
        for (int j = 0; j < win; j += 4) {
            float4 xv[C];
#pragma unroll
            for (int c = 0; c < C; ++c)
                xv[c] = *reinterpret_cast<const float4 *>(X + (col0 + c * cstep) * XS + j);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float4 g = *reinterpret_cast<const float4 *>(G + (row0 + k * rstep) * GS + j);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    if (PACKED) {
                        ffma2(accp[k][c], make_float2(g.x, g.y), make_float2(xv[c].x, xv[c].y));
                        ffma2(accp[k][c], make_float2(g.z, g.w), make_float2(xv[c].z, xv[c].w));
                    } else {
                        accs[k][c] = __fmaf_rn(g.x, xv[c].x, accs[k][c]);
                        accs[k][c] = __fmaf_rn(g.y, xv[c].y, accs[k][c]);
                        accs[k][c] = __fmaf_rn(g.z, xv[c].z, accs[k][c]);
                        accs[k][c] = __fmaf_rn(g.w, xv[c].w, accs[k][c]);
                    }
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int c = 0; c < C; ++c) s += PACKED ? accp[k][c].x + accp[k][c].y : accs[k][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int K, int C, int MAP, bool PACKED>
int run_core(int ctas_per_sm, int sm_count, int win, float *d_out, double *tflops, double *aux) {
    constexpr int WARPS = 8;
    constexpr int ROWS_PER_WARP = MAP == 0 ? K : 4 * K;
    constexpr int COLS_PER_WARP = MAP == 0 ? 32 * C : 8 * C;
    constexpr int KT = ROWS_PER_WARP * 2;
    constexpr int NC = COLS_PER_WARP * (WARPS / 2);
    const size_t smem = (size_t)(KT + NC) * (win + 4) * sizeof(float);
    auto kern = k_core<K, C, MAP, PACKED>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return 1;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
    if (occ < 1) return 1;
    if (ctas_per_sm > occ) ctas_per_sm = occ;
    const int grid = sm_count * ctas_per_sm;
    const int reps = 200;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<grid, 256, smem>>>(d_out, 4, win);
    cudaEventRecord(e0);
    kern<<<grid, 256, smem>>>(d_out, reps, win);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) return 2;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double fma = (double)grid * 256.0 * reps * (win / 4) * 4.0 * K * C;
    *tflops = 2.0 * fma / (ms * 1e-3) / 1e12;
    if (aux) *aux = (double)occ;
    return 0;
}

}  // namespace

extern "C" int rsb_microbench(int device, int id, int arg, double *result, double *aux) {
    if (cudaSetDevice(device) != cudaSuccess) return RSB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int sms = prop.multiProcessorCount;
    float *d_out = nullptr;
    if (cudaMalloc(&d_out, (size_t)sms * 16 * 256 * sizeof(float)) != cudaSuccess) return RSB_ERR_OUT_OF_MEMORY;
    int rc = 0;
    double r = 0.0, a = 0.0;
    if (id == 2) {
        const int ctas = sms * (arg > 0 ? arg : 4);
        const int iters = 4096;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k_mma_tf32_peak<<<ctas, 256>>>(d_out, 16);
        cudaEventRecord(e0);
        k_mma_tf32_peak<<<ctas, 256>>>(d_out, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) rc = 2;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        // 8 mma per iteration per warp, 16*8*8 MACs each
        const double mac = (double)ctas * 8.0 * iters * 8.0 * 1024.0;
        r = 2.0 * mac / (ms * 1e-3) / 1e12;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    } else if (id == 0 || id == 1) {
        const int ctas = sms * (arg > 0 ? arg : 4);
        const int iters = 4096;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        if (id == 0) k_ffma_peak<<<ctas, 256>>>(d_out, 16, 1.0f, 2.0f);
        else k_ffma2_peak<<<ctas, 256>>>(d_out, 16, 1.0f, 2.0f);
        cudaEventRecord(e0);
        if (id == 0) k_ffma_peak<<<ctas, 256>>>(d_out, iters, 1.0f, 2.0f);
        else k_ffma2_peak<<<ctas, 256>>>(d_out, iters, 1.0f, 2.0f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) rc = 2;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fma = (double)ctas * 256.0 * iters * 64.0 * (id == 1 ? 2.0 : 1.0);
        r = 2.0 * fma / (ms * 1e-3) / 1e12;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    } else {
        const int win = 160;
        const int cps = arg > 0 ? arg : 2;
        switch (id) {
            case 10: rc = run_core<8, 4, 0, false>(cps, sms, win, d_out, &r, &a); break;
            case 11: rc = run_core<8, 4, 0, true>(cps, sms, win, d_out, &r, &a); break;
            case 12: rc = run_core<8, 4, 1, false>(cps, sms, win, d_out, &r, &a); break;
            case 13: rc = run_core<8, 4, 1, true>(cps, sms, win, d_out, &r, &a); break;
            case 14: rc = run_core<8, 2, 0, false>(cps, sms, win, d_out, &r, &a); break;
            case 15: rc = run_core<8, 2, 0, true>(cps, sms, win, d_out, &r, &a); break;
            case 16: rc = run_core<8, 2, 1, false>(cps, sms, win, d_out, &r, &a); break;
            case 17: rc = run_core<8, 2, 1, true>(cps, sms, win, d_out, &r, &a); break;
            case 18: rc = run_core<4, 4, 1, false>(cps, sms, win, d_out, &r, &a); break;
            case 19: rc = run_core<4, 4, 1, true>(cps, sms, win, d_out, &r, &a); break;
            case 20: rc = run_core<16, 2, 0, false>(cps, sms, win, d_out, &r, &a); break;
            case 21: rc = run_core<16, 2, 0, true>(cps, sms, win, d_out, &r, &a); break;
            case 22: rc = run_core<4, 8, 1, false>(cps, sms, win, d_out, &r, &a); break;
            case 23: rc = run_core<4, 8, 1, true>(cps, sms, win, d_out, &r, &a); break;
            case 24: rc = run_core<8, 8, 1, false>(cps, sms, win, d_out, &r, &a); break;
            case 25: rc = run_core<8, 8, 1, true>(cps, sms, win, d_out, &r, &a); break;
            case 26: rc = run_core<16, 4, 1, false>(cps, sms, win, d_out, &r, &a); break;
            case 27: rc = run_core<16, 4, 1, true>(cps, sms, win, d_out, &r, &a); break;
            default: rc = 3;
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(d_out);
    if (result) *result = r;
    if (aux) *aux = a;
    if (rc == 3) return RSB_ERR_INVALID_ARGUMENT;
    if (rc != 0 || e != cudaSuccess) return RSB_ERR_CUDA;
    return RSB_OK;
}
