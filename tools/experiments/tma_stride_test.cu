// Experiment: can a tiled TMA tensor copy de-interleave stereo frames (elementStrides = {2,1})?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_stride_test tma_stride_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k(const __grid_constant__ CUtensorMap tm, float *out, int c0, int c1, int n, int nbytes) {
    extern __shared__ __align__(128) float smem[];
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

int main() {
    const int frames = 1000, streams = 8, stride = 2048;   // floats per stream
    std::vector<float> h(streams * stride);
    for (int s = 0; s < streams; ++s)
        for (int f = 0; f < frames; ++f) {
            h[s * stride + 2 * f] = 1000.f * s + f;            // L
            h[s * stride + 2 * f + 1] = -(1000.f * s + f);     // R
        }
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (!fn) { printf("no entry point\n"); return 1; }
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)2 * frames, (cuuint64_t)streams};
    cuuint64_t strides[1] = {(cuuint64_t)stride * 4};
    const int span = 168;                       // source elements spanned
    cuuint32_t box[2] = {span, 4};              // 4 streams
    cuuint32_t estr[2] = {2, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    if (r) return 1;
    const int per_row = span / 2, n = per_row * 4;
    cudaMalloc(&o, n * 4);
    for (int nbytes : {n * 4, n * 8})
    for (int c0 : {0, 1, 7, 1990}) {
        k<<<1, 128, 16384>>>(tm, o, c0, 2, n, nbytes);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> ho(n);
        cudaMemcpy(ho.data(), o, n * 4, cudaMemcpyDeviceToHost);
        printf("nbytes=%d c0=%d err=%s row0: %g %g %g %g ... %g | row1: %g %g\n", nbytes, c0, cudaGetErrorString(e), ho[0], ho[1], ho[2], ho[3], ho[per_row - 1], ho[per_row], ho[per_row + 1]);
    }
    return 0;
}
