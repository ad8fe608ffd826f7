// tcgen05 / TMA / mbarrier PTX wrappers and descriptor helpers shared by the tensor-core kernels
// (conv_tc.cu: forward / data gradient; wgrad_tc.cu: weight gradient).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace svb {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    // bounded spin: a lost arrival (bad descriptor, wrong byte count) must surface as a launch
    // failure through the C ABI, never as a hung GPU
    const long long t_start = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3FFu) == 0 && clock64() - t_start > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// converged-warp forms (every lane executes them, the elect.sync lane acts): UBLKCP takes uniform-register operands, and a
// divergent single-lane producer pays a R2UR waterfall loop per copy exactly like a divergent MMA issuer (see umma below)
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes, uint32_t elected) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(
                     smem_u32(bar)),
                 "r"(bytes), "r"(elected)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "r"(elected)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_mcast(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// converged-warp form of tcgen05.commit (see the issue discipline above umma)
__device__ __forceinline__ void umma_commit(uint64_t *bar, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)),
        "r"(elected)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1).
// Rows are 128 bytes; 8-row groups (1024 B atoms) are SBO apart; within an atom the 16-byte chunk
// index is XORed with (row & 7).  hi word is constant per operand; lo word carries the address.
__device__ __forceinline__ uint32_t desc_hi_sw128(uint32_t base_offset) {
    // bits [32,46) SBO>>4 = 64 ; [46,48) version = 1 ; [49,52) base offset ; [61,64) layout = 2 (SWIZZLE_128B)
    return 64u | (1u << 14) | ((base_offset & 7u) << 17) | (2u << 29);
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }

// cute::UMMA::InstrDescriptor: fp32 accumulate, both operands K-major; fmt 1 = BF16 (kind::f16), 2 = TF32
__host__ __device__ inline uint32_t umma_idesc(int fmt, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Issue discipline (measured with tools/mma_probe.cu, profiles/r02_mma_probe.txt): a tcgen05.mma issued from a
// DIVERGENT single thread (`if (elected) { ... }`) makes the compiler wrap every UTCHMMA in a per-operand R2UR "waterfall"
// loop (ELECT / R2UR.BROADCAST / BRA.U.ANY): 65-82 cycles per MMA whatever its size, i.e. slower than the tensor
// pipe for every N < 256.  Issued warp-CONVERGED -- all 32 lanes execute the same instruction stream and the MMA itself
// is predicated on the elect.sync lane -- descriptors and addresses live in uniform registers and one warp reaches the
// hardware limits (N = 32: 40 clk = shared-memory operand fetch, N = 64: 48, N = 128: 64 = tensor floor).
// The MMA warp must therefore be converged whenever it calls umma / umma_commit (hence __syncwarp() after every wait).
//
// COLL = use of the tensor core's A-operand collector buffer (PTX .collector::a::*): 0 discard (read A from shared
// memory, keep nothing), 1 fill (read and keep), 2 lastuse (take A from the collector: no shared-memory read of A),
// 3 use (take and keep).  Consecutive MMAs with the SAME A descriptor form a fill / use... / lastuse run.
#define SVB_UMMA_ASM(KIND, COLLECT)                                                                         \
    asm volatile(                                                                                           \
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"                                                   \
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"                                                \
        "setp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %7, 0;\n\t"                                                \
        "@q tcgen05.mma.cta_group::1.kind::" KIND COLLECT " [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),        \
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc), "r"(elected)                      \
        : "memory")

// `elected` = elect_one_sync() of the converged issuing warp (non-zero in exactly one lane)
template <bool BF, int COLL = 0>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                     uint32_t idesc, uint32_t acc, uint32_t elected) {
    if (BF) {
        if (COLL == 1) SVB_UMMA_ASM("f16", ".collector::a::fill");
        else if (COLL == 2) SVB_UMMA_ASM("f16", ".collector::a::lastuse");
        else if (COLL == 3) SVB_UMMA_ASM("f16", ".collector::a::use");
        else SVB_UMMA_ASM("f16", "");
    } else {
        if (COLL == 1) SVB_UMMA_ASM("tf32", ".collector::a::fill");
        else if (COLL == 2) SVB_UMMA_ASM("tf32", ".collector::a::lastuse");
        else if (COLL == 3) SVB_UMMA_ASM("tf32", ".collector::a::use");
        else SVB_UMMA_ASM("tf32", "");
    }
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// split x into bf16 hi + bf16 lo (x ~= hi + lo to 2^-17 relative); returns packed pairs
__device__ __forceinline__ void bf16_split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
    hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

}  // namespace svb
