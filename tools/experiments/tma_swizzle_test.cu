// Experiment: which (data type, swizzle, box) combinations does a 2-D tiled TMA copy accept, and
// where do the 16-byte units land?  One variant per process (an illegal instruction is sticky).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_swizzle_test tma_swizzle_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k(const __grid_constant__ CUtensorMap tm, float *out, int c0, int c1, int n, int nbytes) {
    extern __shared__ __align__(1024) float smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(nbytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(&tm), "r"(c0), "r"(c1),
                       "r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = smem[i];
}

int main(int argc, char **argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int frames = (variant == 8) ? 3000 : 1000, streams = 200, ch = (variant == 2 || variant == 3) ? 1 : 2;
    const int stride = (variant == 8) ? 6000 : 2048;   // floats per stream
    std::vector<float> h((size_t)streams * stride);
    for (int s = 0; s < streams; ++s)
        for (int f = 0; f < frames * ch; ++f) h[(size_t)s * stride + f] = 10000.f * s + f;   // value = member*10000 + float index
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (!fn) { printf("no entry point\n"); return 1; }
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)frames, (cuuint64_t)streams};
    cuuint64_t strides[1] = {(cuuint64_t)stride * 4};
    cuuint32_t box[2] = {16, 64};
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
    switch (variant) {
        case 0: break;
        case 1: sw = CU_TENSOR_MAP_SWIZZLE_NONE; break;
        case 2: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; box[0] = 32; break;
        case 3: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; sw = CU_TENSOR_MAP_SWIZZLE_64B; box[0] = 16; box[1] = 128; break;
        case 4: box[1] = 8; break;
        case 5: sw = CU_TENSOR_MAP_SWIZZLE_64B; box[0] = 8; break;
        case 6: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; dims[0] = 2 * frames; box[0] = 32; break;   // stereo as 2 floats
    }
    int c0 = 32, c1 = 3;
    if (variant == 7) { dims[1] = 70; c1 = 64; }
    if (variant == 8) { dims[1] = 70; c1 = 64; dims[0] = 3000; }
    if (variant == 9) c0 = 33;
    if (variant == 10) c0 = 992;
    if (variant == 11) c0 = -5;
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tm, dt, 2, d, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d encode rc=%d\n", variant, (int)r);
    if (r) return 1;
    const int esz = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT64 ? 8 : 4;
    const int nbytes = box[0] * box[1] * esz, n = nbytes / 4;
    cudaMalloc(&o, n * 4);
    k<<<1, 128, 16384>>>(tm, o, c0, c1, n, nbytes);
    if (cudaGetLastError() != cudaSuccess) { printf("launch failed\n"); return 3; }
    cudaError_t e = cudaDeviceSynchronize();
    printf("err=%s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<float> ho(n);
    cudaMemcpy(ho.data(), o, n * 4, cudaMemcpyDeviceToHost);
    const int row_floats = box[0] * esz / 4;
    for (int row = 0; row < 10; ++row) {
        printf("row %d:", row);
        for (int i = 0; i < row_floats; i += 4) printf(" %g", ho[row * row_floats + i]);
        printf("\n");
    }
    return 0;
}
