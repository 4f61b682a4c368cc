// sm100_ptx.cuh -- thin inline-PTX wrappers shared by the sm_100a kernels: mbarriers, TMA bulk and
// tensor copies, cp.async, and the tcgen05 (5th-generation tensor core / tensor memory) family.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rsb {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)      // suspend-time hint: sleep instead of re-polling
        : "memory");
}
// Non-blocking test of the phase with the given parity.
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// TMA bulk copy global -> shared (SASS: UBLKCP), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
// TMA tiled tensor copy global -> shared of one 2-D box (SASS: UTMALDG)
__device__ __forceinline__ void tensor_g2s_2d(void *dst, const CUtensorMap *tm, int c0, int c1,
                                              uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
        "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- named barriers (subsets of a CTA) ----
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- tcgen05: tensor memory (TMEM) management ----
// Allocates `cols` (power of two >= 32) TMEM columns; the base address is written to *dst (shared
// memory).  Executed by one full warp, which also has to free the columns before the CTA exits.
__device__ __forceinline__ void tmem_alloc(uint32_t *dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// All prior tcgen05.mma of this thread arrive (once) on `bar` when they have completed.
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor], kind::tf32.  Called by ALL lanes of a converged warp
// with warp-uniform arguments; one elected lane issues.  (Issuing from inside an `if (lane == 0)`
// makes ptxas wrap every MMA in an ELECT / BRA.U.ANY loop: measured 66 cycles per MMA, four
// times the 16 cycles an M=128, N=32 MMA keeps the tensor pipe busy.)  The descriptor is passed
// as (lo, hi) halves: only `lo` (the start address field) changes between K steps.
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo,
                                               uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        ".reg .b64 bd;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Two MMAs into the same accumulator with one election: the first honours `accumulate`, the
// second always accumulates.
__device__ __forceinline__ void tc_mma_tf32_ts_x2(uint32_t d_tmem, uint32_t a0, uint32_t b0_lo,
                                                  uint32_t a1, uint32_t b1_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t, e;\n\t"
        ".reg .b64 bd0, bd1;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 bd0, {%2, %5};\n\t"
        "mov.b64 bd1, {%4, %5};\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd0, %6, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%3], bd1, %6, t;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a0), "r"(b0_lo), "r"(a1), "r"(b1_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Four MMAs into the same accumulator with one election (measured 23 cycles per MMA against 36
// for one asm block per MMA, tools/experiments/mma_issue_test.cu): the first honours
// `accumulate`, the others always accumulate.
__device__ __forceinline__ void tc_mma_tf32_ts_x4(uint32_t d_tmem, uint32_t a0, uint32_t a1,
                                                  uint32_t a2, uint32_t a3, uint32_t b0_lo,
                                                  uint32_t b1_lo, uint32_t b2_lo, uint32_t b3_lo,
                                                  uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t, e;\n\t"
        ".reg .b64 d0, d1, d2, d3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 d0, {%5, %9};\n\t"
        "mov.b64 d1, {%6, %9};\n\t"
        "mov.b64 d2, {%7, %9};\n\t"
        "mov.b64 d3, {%8, %9};\n\t"
        "setp.ne.b32 p, %11, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], d0, %10, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], d1, %10, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%3], d2, %10, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%4], d3, %10, t;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0_lo), "r"(b1_lo), "r"(b2_lo), "r"(b3_lo), "r"(b_hi),
        "r"(idesc), "r"(accumulate)
        : "memory");
}
// All prior tcgen05.mma of the electing warp arrive (once) on `bar` when they have completed;
// called by all lanes of a converged warp.
__device__ __forceinline__ void tc_commit_elect(uint64_t *bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar))
        : "memory");
}
// Each thread of the warp writes 16 consecutive 32-bit columns of its own TMEM lane.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
        "%12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
// The same for 8 columns.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// Each thread of the warp reads 32 consecutive 32-bit columns of its own TMEM lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
        "%12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, "
        "%30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// Round-to-nearest conversion to TF32 (result keeps the f32 container, low 13 bits zero).
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

}  // namespace ptx
}  // namespace rsb
