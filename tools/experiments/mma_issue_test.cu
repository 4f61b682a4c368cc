// Experiment: how fast can one warp issue small tcgen05.mma (kind::tf32, M=128, N=32/64, A in TMEM)?
// Measures cycles per MMA for several issue styles.  Operand contents are garbage (timing only).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue_test mma_issue_test.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define IDESC(N) ((1u << 4) | (2u << 7) | (2u << 10) | (((N) >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ void mma1(uint32_t d, uint32_t a, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 bd;\n\telect.sync _|e, 0xffffffff;\n\tmov.b64 bd, {%2, %3};\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}"
                 ::"r"(d), "r"(a), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
}
// PTX loop: n MMAs, A column += 8 (wrap at 224), descriptor += 64
__device__ __forceinline__ void mma_loop(uint32_t d, uint32_t a, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t n) {
    asm volatile("{\n\t.reg .pred p, e, w, q;\n\t.reg .b64 bd;\n\t.reg .b32 ra, rb, rn, lim;\n\t"
                 "elect.sync _|e, 0xffffffff;\n\t"
                 "mov.b32 ra, %1;\n\tmov.b32 rb, %2;\n\tmov.b32 rn, %5;\n\t"
                 "add.u32 lim, %1, 224;\n\t"
                 "setp.eq.b32 p, 0, 0;\n\t"
                 "LOOP:\n\t"
                 "mov.b64 bd, {rb, %3};\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [ra], bd, %4, p;\n\t"
                 "add.u32 ra, ra, 8;\n\t"
                 "setp.ge.u32 w, ra, lim;\n\t"
                 "@w sub.u32 ra, ra, 224;\n\t"
                 "add.u32 rb, rb, 64;\n\t"
                 "sub.u32 rn, rn, 1;\n\t"
                 "setp.ne.u32 q, rn, 0;\n\t"
                 "@q bra LOOP;\n\t}"
                 ::"r"(d), "r"(a), "r"(blo), "r"(bhi), "r"(idesc), "r"(n) : "memory");
}
// 4 MMAs in one block, consecutive K steps, no wrap
__device__ __forceinline__ void mma4(uint32_t d, uint32_t a, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 b0, b1, b2, b3;\n\t.reg .b32 t1, t2, t3, a1, a2, a3;\n\t"
                 "elect.sync _|e, 0xffffffff;\n\t"
                 "add.u32 t1, %2, 64;\n\tadd.u32 t2, %2, 128;\n\tadd.u32 t3, %2, 192;\n\t"
                 "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
                 "mov.b64 b0, {%2, %3};\n\tmov.b64 b1, {t1, %3};\n\tmov.b64 b2, {t2, %3};\n\tmov.b64 b3, {t3, %3};\n\t"
                 "setp.eq.b32 p, 0, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], b0, %4, p;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a1], b1, %4, p;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a2], b2, %4, p;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a3], b3, %4, p;\n\t}"
                 ::"r"(d), "r"(a), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
// 4 MMAs in one block, every A address and descriptor word passed in
__device__ __forceinline__ void mma4g(uint32_t d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1,
                                      uint32_t b2, uint32_t b3, uint32_t bhi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, t, e;\n\t.reg .b64 d0, d1, d2, d3;\n\t"
                 "elect.sync _|e, 0xffffffff;\n\t"
                 "mov.b64 d0, {%5, %9};\n\tmov.b64 d1, {%6, %9};\n\tmov.b64 d2, {%7, %9};\n\tmov.b64 d3, {%8, %9};\n\t"
                 "setp.ne.b32 p, %11, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], d0, %10, p;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], d1, %10, t;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%3], d2, %10, t;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%4], d3, %10, t;\n\t}"
                 ::"r"(d), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(b2), "r"(b3), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
}
// 4 MMAs in one block, accumulators alternate d0, d1, d0, d1
__device__ __forceinline__ void mma4alt(uint32_t d0, uint32_t d1, uint32_t a, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 b0, b1, b2, b3;\n\t.reg .b32 t1, t2, t3, a1, a2, a3;\n\t"
                 "elect.sync _|e, 0xffffffff;\n\t"
                 "add.u32 t1, %3, 64;\n\tadd.u32 t2, %3, 128;\n\tadd.u32 t3, %3, 192;\n\t"
                 "add.u32 a1, %2, 8;\n\tadd.u32 a2, %2, 16;\n\tadd.u32 a3, %2, 24;\n\t"
                 "mov.b64 b0, {%3, %4};\n\tmov.b64 b1, {t1, %4};\n\tmov.b64 b2, {t2, %4};\n\tmov.b64 b3, {t3, %4};\n\t"
                 "setp.eq.b32 p, 0, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], b0, %5, p;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [a1], b1, %5, p;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a2], b2, %5, p;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [a3], b3, %5, p;\n\t}"
                 ::"r"(d0), "r"(d1), "r"(a), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__global__ void k(int variant, int n_mma, int reps, long long *out, uint32_t col_arg, int fill) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_base;
    const uint32_t warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        float r = ((h & 0xffffff) / 8388608.0f - 1.0f);
        float v = fill == 0 ? 0.f : fill == 1 ? r : fill == 2 ? r * 2.4e-4f : (i % 5 ? 0.f : r);
        reinterpret_cast<float *>(smem)[i] = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    {   // fill the A region (columns 0..447) of every lane
        const uint32_t lane_base = ((warp & 3u) * 32u) << 16;
        for (uint32_t c0 = 0; c0 < 448; c0 += 8) {
            uint32_t v[8];
            for (int e = 0; e < 8; ++e) {
                uint32_t h = (threadIdx.x * 977u + c0 + e) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
                float r = ((h & 0xffffff) / 8388608.0f - 1.0f);
                float x = fill == 0 ? 0.f : fill == 1 ? r : fill == 2 ? r * 2.4e-4f : r;
                v[e] = __float_as_uint(x) & 0xffffe000u;
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tmem + lane_base + c0),
                         "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (variant == 5 && warp == 1) {
        const uint32_t blo0 = (((smem_u32(smem) + 32768u) & 0x3ffffu) >> 4) | ((512u >> 4) << 16);
        const uint32_t bhi = (128u >> 4) | (1u << 14);
        for (int r = 0; r < reps; ++r) {
            uint32_t col = 224, blo = blo0;
            for (int i = 0; i < n_mma; i += 4) {
                mma4(tmem + 480, tmem + col, blo, bhi, IDESC(32));
                blo += 256; col += 32; if (col >= 416) col -= 192;
                if ((i & 31) == 28) blo = blo0;
            }
        }
    }
    if (warp == 0) {
        const uint32_t blo0 = ((smem_u32(smem) & 0x3ffffu) >> 4) | ((512u >> 4) << 16);
        const uint32_t bhi = (128u >> 4) | (1u << 14);
        const uint32_t idesc = variant >= 20 ? IDESC(16) : variant >= 10 ? IDESC(64) : IDESC(32);
        const int v = variant % 10;
        uint32_t parity = 0;
        long long t_issue = 0, t_total = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            uint32_t col = col_arg, blo = blo0;
            if (v == 0 || false) {
                for (int i = 0; i < n_mma; ++i) {
                    mma1(tmem + 448, tmem + col, blo, bhi, idesc, 1u);
                    blo += 64; col += 8; if (col >= 224) col -= 224;
                    if ((i & 31) == 31) blo = blo0;
                }
            } else if (v == 1) {
                mma_loop(tmem + 448, tmem + col, blo, bhi, idesc, (uint32_t)n_mma);
            } else if (v == 2 || v == 5) {
                for (int i = 0; i < n_mma; i += 4) {
                    mma4(tmem + 448, tmem + col, blo, bhi, idesc);
                    blo += 256; col += 32; if (col >= 192) col -= 192;
                    if ((i & 31) == 28) blo = blo0;
                }
            } else if (v == 4) {
                for (int i = 0; i < n_mma; i += 4) {
                    mma4alt(tmem + 448, tmem + 480, tmem + col, blo, bhi, idesc);
                    blo += 256; col += 32; if (col >= 192) col -= 192;
                    if ((i & 31) == 28) blo = blo0;
                }
            } else if (v == 6) {
                // a commit (to a second barrier nobody waits on) after every 4 MMAs
                for (int i = 0; i < n_mma; i += 4) {
                    mma4(tmem + 448, tmem + col, blo, bhi, idesc);
                    commit(&bar2);
                    blo += 256; col += 32; if (col >= 192) col -= 192;
                    if ((i & 31) == 28) blo = blo0;
                }
            } else if (v == 3) {
                const uint32_t glo = blo0 + 1408;
                uint32_t dk = 0;
                for (int i = 0; i < n_mma; i += 4) {
                    uint32_t c1 = col + 8; if (c1 >= 224) c1 -= 224;
                    mma4g(tmem + 448, tmem + 224 + col, tmem + col, tmem + 224 + c1, tmem + c1, blo0 + dk, glo + dk,
                          blo0 + dk + 64, glo + dk + 64, bhi, idesc, i != 0);
                    dk += 128; if (dk >= 1280) dk = 0;
                    col = c1 + 8; if (col >= 224) col -= 224;
                }
            }
            const long long t1 = clock64();
            commit(&bar);
            mbar_wait(&bar, parity);
            parity ^= 1u;
            const long long t2 = clock64();
            t_issue += t1 - t0;
            t_total += t2 - t0;
        }
        if (threadIdx.x == 0) { out[0] = t_issue; out[1] = t_total; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main(int argc, char **argv) {
    long long *d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int n_mma = 64, reps = 200;
    const int fill = argc > 2 ? atoi(argv[2]) : 0;
    for (int variant : {2, 6, 12, 16, 22}) {
        k<<<1, 128, 65536>>>(variant, n_mma, reps, d, (uint32_t)(argc > 1 ? atoi(argv[1]) : 0), fill);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("fill %d variant %2d (N=%d, %s): err=%s issue %.1f cyc/MMA, issue+drain %.1f cyc/MMA\n", fill, variant, variant >= 20 ? 16 : variant >= 10 ? 64 : 32,
               variant % 10 == 6 ? "4 per asm + one commit per 4 MMAs" : variant % 10 == 4 ? "4 per asm, accumulators alternate" : variant % 10 == 5 ? "4 per asm, a second warp issues into another accumulator" : variant % 10 == 2 ? "4 per asm" : "other", cudaGetErrorString(e),
               (double)h[0] / (n_mma * reps), (double)h[1] / (n_mma * reps));
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
