// dram_cursor_bench.cu -- how much HBM bandwidth does the resampler's access pattern allow?
// 1024 streams 21 / 23 MB apart; 148 CTAs, each walking 64 streams forward in lock step (like the
// tensor kernel's work items), touching W contiguous bytes per stream per visit.  Plain coalesced
// loads / stores (no TMA): isolates the DRAM side.  Prints GB/s for read-only, write-only and
// read + write, W = 128 ... 4096.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dram_cursor_bench dram_cursor_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kStreams = 1024, kPerCta = 64, kGroups = kStreams / kPerCta;

// mode bit 0: read, bit 1: write.  W = bytes per stream per visit (multiple of 512).
// Each of the CTA's 8 warps owns 8 streams; a lane moves 16 bytes per access (512 B per warp access).
template <int LINES>   // W = LINES * 128 bytes (a lane moves 4 bytes per access)
__global__ void __launch_bounds__(256) walk(const float *in, float *out, size_t in_stride16, size_t out_stride16,
                                            size_t run16, int runs_total, int mode, float *sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int item = blockIdx.x; item < runs_total * kGroups; item += gridDim.x) {
        const int run = item / kGroups, group = item % kGroups;
        const size_t t0 = (size_t)run * run16;
        for (size_t t = 0; t < run16; t += (size_t)LINES * 32) {
            float v[8][LINES];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const size_t stream = (size_t)group * kPerCta + warp * 8 + s;
#pragma unroll
                for (int l = 0; l < LINES; ++l) {
                    const size_t idx = t0 + t + (size_t)l * 32 + lane;
                    if (mode & 1) v[s][l] = __ldcs(in + stream * in_stride16 + idx);
                    else v[s][l] = (float)idx;
                }
            }
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const size_t stream = (size_t)group * kPerCta + warp * 8 + s;
#pragma unroll
                for (int l = 0; l < LINES; ++l) {
                    const size_t idx = t0 + t + (size_t)l * 32 + lane;
                    if (mode & 2) __stcs(out + stream * out_stride16 + idx, v[s][l]);
                    else acc += v[s][l];
                }
            }
        }
    }
    if (acc == 123.456f) *sink = acc;
}

int main() {
    const size_t in_stride = 21168000, out_stride = 23040000;     // bytes per stream (60 s stereo f32)
    float *in, *out;
    float *sink;
    CK(cudaMalloc(&in, in_stride * kStreams));
    CK(cudaMalloc(&out, out_stride * kStreams));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(in, 0, in_stride * kStreams));
    CK(cudaMemset(out, 0, out_stride * kStreams));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const size_t run_bytes = 49152;                       // per stream per work item (~96 tiles)
    const size_t usable = 5 * 1000 * 1000;                  // bytes per stream walked
    const int runs_total = (int)(usable / run_bytes);
    printf("streams %d, stride in %zu out %zu, run %zu B, %d runs x %d groups\n", kStreams, in_stride, out_stride,
           run_bytes, runs_total, kGroups);
    for (int grid : {148 * 2, 148 * 4, 148 * 8}) {
        for (int mode = 1; mode <= 3; ++mode) {
            for (int lines : {1, 2, 4, 8, 16}) {
                auto launch = [&]() {
                    switch (lines) {
                        case 1: walk<1><<<grid, 256>>>(in, out, in_stride / 4, out_stride / 4, run_bytes / 4, runs_total, mode, sink); break;
                        case 2: walk<2><<<grid, 256>>>(in, out, in_stride / 4, out_stride / 4, run_bytes / 4, runs_total, mode, sink); break;
                        case 4: walk<4><<<grid, 256>>>(in, out, in_stride / 4, out_stride / 4, run_bytes / 4, runs_total, mode, sink); break;
                        case 8: walk<8><<<grid, 256>>>(in, out, in_stride / 4, out_stride / 4, run_bytes / 4, runs_total, mode, sink); break;
                        default: walk<16><<<grid, 256>>>(in, out, in_stride / 4, out_stride / 4, run_bytes / 4, runs_total, mode, sink); break;
                    }
                };
                launch();
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                launch();
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                const double bytes = (double)runs_total * run_bytes * kStreams * ((mode & 1 ? 1 : 0) + (mode & 2 ? 1 : 0));
                printf("grid %3d mode %s W %4d B per stream visit: %7.1f GB/s (%.2f ms)\n", grid,
                       mode == 1 ? "read " : mode == 2 ? "write" : "r + w", lines * 128, bytes / ms / 1e6, ms);
            }
        }
    }
    // reference: plain linear copy of the same volume
    {
        const size_t n = (size_t)kStreams * usable;
        CK(cudaMemcpyAsync(out, in, n, cudaMemcpyDeviceToDevice));
        CK(cudaEventRecord(e0));
        CK(cudaMemcpyAsync(out, in, n, cudaMemcpyDeviceToDevice));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("cudaMemcpy D2D %zu B: %.1f GB/s (read + write)\n", n, 2.0 * n / ms / 1e6);
    }
    return 0;
}
