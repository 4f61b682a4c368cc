// tma_cursor_bench.cu -- TMA version of dram_cursor_bench: 148 persistent CTAs stream 64 stereo
// streams (21 / 23 MB apart) through shared memory with tensor copies and (optionally) store the
// same bytes back out with tensor stores.  Compares how the boxes are shaped / ordered:
//   mode 0: 2-D box {16 frames, 64 members} (128 B per member), one per stage, 8 stages
//   mode 1: the same boxes issued 4 at a time (4 consecutive chunks back to back), 2 x 4 stages
//   mode 2: 3-D box {16 frames, 4 pieces, 64 members}: member-major, 512 B per member contiguous
//   mode 3: 3-D box {16 frames, 64 members, 4 pieces}: piece-major (== mode 1 in one instruction)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../resampler_b200/csrc -o tma_cursor_bench tma_cursor_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"

using namespace rsb::ptx;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kStreams = 1024, kMembers = 64, kGroups = kStreams / kMembers;
constexpr uint32_t kPieceBytes = 8192;          // {16 frames x 8 B} x 64 members
constexpr uint32_t kSuper = 4;                  // pieces per super stage
constexpr uint32_t kStages = 8;                 // pieces of shared memory in total (64 KB)

__device__ __forceinline__ void tensor_g2s_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tensor_s2g_2d(const CUtensorMap *tm, int c0, int c1, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tensor_s2g_3d(const CUtensorMap *tm, int c0, int c1, int c2, const void *src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)) : "memory");
}

struct Maps { CUtensorMap in2, out2, in3m, out3m, in3p, out3p, out2q; };

// units: "piece" = 16 frames.  A run is run_pieces consecutive pieces of one member group.
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ Maps M, int mode, int do_store, int run_pieces,
                                                       int runs_total, unsigned *counter) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[kStages], empty[kStages];
    __shared__ int s_item[2];
    __shared__ uint64_t item_full[2], item_empty[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t per = (mode == 0 || mode == 4) ? 1u : kSuper;            // pieces per stage unit
    const uint32_t n_units = kStages / per;                    // stage units in the ring
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&item_full[i], 1); mbar_init(&item_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_items = runs_total * kGroups;
    if (warp == 0) {
        if (lane == 0) {
            uint32_t seq = 0;
            for (int it = 0;; ++it) {
                mbar_wait(&item_empty[it & 1], ((it >> 1) & 1) ^ 1);
                const int idx = (int)atomicAdd(counter, 1u);
                s_item[it & 1] = idx < n_items ? idx : -1;
                mbar_arrive(&item_full[it & 1]);
                if (idx >= n_items) break;
                const int run = idx / kGroups, m0 = (idx % kGroups) * kMembers;
                for (int p = 0; p < run_pieces; p += per, ++seq) {
                    const uint32_t u = seq % n_units;
                    mbar_wait(&empty[u], ((seq / n_units) & 1) ^ 1);
                    uint8_t *dst = smem + u * per * kPieceBytes;
                    const int piece = run * run_pieces + p;
                    mbar_arrive_expect_tx(&full[u], per * kPieceBytes);
                    if (mode == 0 || mode == 4) tensor_g2s_2d(dst, &M.in2, piece * 16, m0, &full[u]);
                    else if (mode == 1) for (uint32_t q = 0; q < kSuper; ++q) tensor_g2s_2d(dst + q * kPieceBytes, &M.in2, (piece + q) * 16, m0, &full[u]);
                    else if (mode == 2) tensor_g2s_3d(dst, &M.in3m, 0, piece, m0, &full[u]);
                    else tensor_g2s_3d(dst, &M.in3p, 0, m0, piece, &full[u]);
                }
            }
        }
    } else {
        uint32_t seq = 0;
        for (int it = 0;; ++it) {
            mbar_wait(&item_full[it & 1], (it >> 1) & 1);
            const int idx = s_item[it & 1];
            __syncwarp();
            if (lane == 0) mbar_arrive(&item_empty[it & 1]);
            if (idx < 0) break;
            const int run = idx / kGroups, m0 = (idx % kGroups) * kMembers;
            for (int p = 0; p < run_pieces; p += per, ++seq) {
                const uint32_t u = seq % n_units;
                mbar_wait(&full[u], (seq / n_units) & 1);
                if (lane == 0) {
                    if (do_store) {
                        const uint8_t *src = smem + u * per * kPieceBytes;
                        const int piece = run * run_pieces + p;
                        if (mode == 0) tensor_s2g_2d(&M.out2, piece * 16, m0, src);
                        else if (mode == 4) for (uint32_t q = 0; q < 4; ++q) tensor_s2g_2d(&M.out2q, piece * 16, m0 + 16 * q, src + q * 2048);
                        else if (mode == 1) for (uint32_t q = 0; q < kSuper; ++q) tensor_s2g_2d(&M.out2, (piece + q) * 16, m0, src + q * kPieceBytes);
                        else if (mode == 2) tensor_s2g_3d(&M.out3m, 0, piece, m0, src);
                        else tensor_s2g_3d(&M.out3p, 0, m0, piece, src);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    mbar_arrive(&empty[u]);
                }
                __syncwarp();
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const size_t in_stride = 21168000, out_stride = 23040000;
    const uint64_t frames = 2500000;            // 8-byte stereo frames per stream that the maps cover
    void *in, *out;
    unsigned *counter;
    CK(cudaMalloc(&in, in_stride * kStreams));
    CK(cudaMalloc(&out, out_stride * kStreams));
    CK(cudaMalloc(&counter, 4));
    CK(cudaMemset(in, 0, in_stride * kStreams));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    Maps M;
    auto mk = [&](CUtensorMap *m, void *base, size_t stride, int kind) {
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r;
        if (kind == 0) {
            cuuint64_t d[2] = {frames, (cuuint64_t)kStreams}; cuuint64_t s[1] = {stride}; cuuint32_t b[2] = {16, (cuuint32_t)kMembers};
            r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else if (kind == 1) {   // member-major: {16, pieces, members}
            cuuint64_t d[3] = {16, frames / 16, (cuuint64_t)kStreams}; cuuint64_t s[2] = {128, stride}; cuuint32_t b[3] = {16, kSuper, (cuuint32_t)kMembers};
            r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {                  // piece-major: {16, members, pieces}
            cuuint64_t d[3] = {16, (cuuint64_t)kStreams, frames / 16}; cuuint64_t s[2] = {stride, 128}; cuuint32_t b[3] = {16, (cuuint32_t)kMembers, kSuper};
            r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) { printf("encode kind %d failed: %d\n", kind, (int)r); exit(1); }
    };
    mk(&M.in2, in, in_stride, 0); mk(&M.out2, out, out_stride, 0);
    mk(&M.in3m, in, in_stride, 1); mk(&M.out3m, out, out_stride, 1);
    mk(&M.in3p, in, in_stride, 2); mk(&M.out3p, out, out_stride, 2);
    {
        cuuint32_t es[2] = {1, 1};
        cuuint64_t d[2] = {frames, (cuuint64_t)kStreams}; cuuint64_t st[1] = {out_stride}; cuuint32_t b[2] = {16, 16};
        if (enc(&M.out2q, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, out, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode q failed\n"); exit(1); }
    }
    const int run_pieces = 352, runs_total = (int)(frames / 16 / run_pieces);
    const size_t smem = kStages * kPieceBytes;
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char *names[5] = {"2-D x1 (128 B/member)", "2-D x4 back to back", "3-D member-major 512 B", "3-D piece-major", "2-D x1, stores 4 x 16 rows"};
    for (int do_store = 0; do_store < 2; ++do_store)
        for (int mode = 0; mode < 5; ++mode) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaMemset(counter, 0, 4));
                CK(cudaEventRecord(e0));
                stream_kernel<<<148, 64, smem>>>(M, mode, do_store, run_pieces, runs_total, counter);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            const double bytes = (double)runs_total * run_pieces * 128.0 * kStreams * (1 + do_store);
            printf("%s  %-26s %7.1f GB/s (%.2f ms)\n", do_store ? "load+store" : "load only ", names[mode], bytes / best / 1e6, best);
        }
    return 0;
}
