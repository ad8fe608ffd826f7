// Micro-probe: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, SWIZZLE_128B K-major operands in shared memory)
// as issued by ONE thread, for N in {32, 64, 128, 256}, 1 / 2 / 4 accumulator tiles in rotation, with and without
// the A-collector reuse hint, and with 1 or 2 issuing threads (different warps, disjoint accumulators).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/mma_probe tools/mma_probe.cu && tools/mma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = 64u | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
template <int COLL>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc = 1) {
    if (COLL == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else if (COLL == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// warp-converged issue: every lane executes the instruction stream, the MMA itself is predicated on the elect.sync lane, so all
// operands stay in uniform registers (no per-MMA R2UR "waterfall" loop as in a divergent single-thread issuer)
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred;
}
template <int COLL>
__device__ __forceinline__ void mma_w(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t el, uint32_t acc = 1) {
    if (COLL == 1)
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %5, 0;\n\tsetp.ne.b32 q, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(el), "r"(acc) : "memory");
    else if (COLL == 2)
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %5, 0;\n\tsetp.ne.b32 q, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(el), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %5, 0;\n\tsetp.ne.b32 q, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(el), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
}

// mode 0: every MMA reads A and B from shared memory; mode 1: triples (discard, fill, lastuse) with the same A in the last two
// a_step: byte distance between the A operands of consecutive MMAs (0 = same rows: L1-like reuse is impossible in smem anyway)
__global__ void __launch_bounds__(128, 1) probe(int N, int nd, int iters, int mode, int issuers, int a_bytes, long long *out, float *result) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(smem + 64);
    unsigned char *A = smem + 1024, *Bm = smem + 1024 + 96 * 1024;
    for (int i = threadIdx.x; i < (96 + 64) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(A)[i] = 0x3c003c00u;   // small bf16 values
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *tptr;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (mode >= 2 && warp < issuers) {
        const uint32_t el = elect_one();
        const uint32_t t0c = __shfl_sync(0xffffffffu, tbase, 0) + (uint32_t)warp * 256;
        const uint32_t a0 = s32(A), b0 = s32(Bm);
        const int n_a = a_bytes / (128 * 128);
        for (int j = 0; j < nd; ++j) mma_w<0>(t0c + (uint32_t)(j * N), mk_desc(a0), mk_desc(b0), idesc, el, 0);   // overwrite: accumulators start at one MMA
        long long t0 = clock64();
        int di = 0, ai = 0;
        if (mode >= 4) {
            // the conv kernel's order: taps x 2 M tiles x 2 k-blocks x 3 split products; a tap is a row shift of A
            const int dil = mode == 4 ? 1 : 5, KS = 11;
            const uint32_t tap_step = (uint32_t)dil * 128;
            for (int i = 0; i < iters; i += KS * 12) {
                uint32_t a_tap = a0, b_tap = b0;
                for (int k = 0; k < KS; ++k) {
#pragma unroll 1
                    for (int m = 0; m < 2; ++m) {
                        const uint32_t d = t0c + (uint32_t)(m * N);
                        const uint32_t a_row = a_tap + (uint32_t)m * 16384;
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) {
                            const uint64_t ah = mk_desc(a_row + kb * 32), al = mk_desc(a_row + 64 + kb * 32);
                            const uint64_t bh = mk_desc(b_tap + kb * 32), bl = mk_desc(b_tap + 64 + kb * 32);
                            mma_w<0>(d, al, bh, idesc, el);
                            if (mode == 6) {
                                mma_w<0>(d, ah, bl, idesc, el);
                                mma_w<0>(d, ah, bh, idesc, el);
                            } else {
                                mma_w<1>(d, ah, bl, idesc, el);
                                mma_w<2>(d, ah, bh, idesc, el);
                            }
                        }
                    }
                    a_tap += tap_step, b_tap += (uint32_t)N * 128;
                    if (b_tap + (uint32_t)N * 128 > b0 + 64 * 1024) b_tap = b0;
                }
            }
        } else if (mode == 2) {
#pragma unroll 4
            for (int i = 0; i < iters; ++i) {
                mma_w<0>(t0c + (uint32_t)(di * N), mk_desc(a0 + (uint32_t)ai * 16384 + (i & 1) * 32), mk_desc(b0 + (i & 3) * 32), idesc, el);
                if (++di == nd) di = 0;
                if (++ai == n_a) ai = 0;
            }
        } else {
#pragma unroll 2
            for (int i = 0; i < iters; i += 3) {
                const uint32_t d = t0c + (uint32_t)(di * N);
                const uint64_t ah = mk_desc(a0 + (uint32_t)ai * 16384), al = mk_desc(a0 + (uint32_t)ai * 16384 + 64);
                const uint64_t bh = mk_desc(b0), bl = mk_desc(b0 + 64);
                mma_w<0>(d, al, bh, idesc, el);
                mma_w<1>(d, ah, bl, idesc, el);
                mma_w<2>(d, ah, bh, idesc, el);
                if (++di == nd) di = 0;
                if (++ai == n_a) ai = 0;
            }
        }
        long long t1 = clock64();
        if (el) {
            commit(bar + warp);
            wait(bar + warp, 0);
            long long t2 = clock64();
            out[(blockIdx.x * 2 + warp) * 2 + 0] = t1 - t0;
            out[(blockIdx.x * 2 + warp) * 2 + 1] = t2 - t0;
        }
    } else if ((threadIdx.x & 31) == 0 && warp < issuers) {
        // this issuer's accumulators: nd tiles of N columns, issuers use disjoint halves of TMEM
        const uint32_t t0c = tbase + (uint32_t)warp * 256;
        const uint32_t a0 = s32(A), b0 = s32(Bm);
        const int n_a = a_bytes / (128 * 128);           // distinct 128-row A operands (16 KB apart)
        for (int j = 0; j < nd; ++j) mma<0>(t0c + (uint32_t)(j * N), mk_desc(a0), mk_desc(b0), idesc, 0);
        long long t0 = clock64();
        int di = 0, ai = 0;
        if (mode == 0) {
#pragma unroll 4
            for (int i = 0; i < iters; ++i) {
                mma<0>(t0c + (uint32_t)(di * N), mk_desc(a0 + (uint32_t)ai * 16384 + (i & 1) * 32), mk_desc(b0 + (i & 3) * 32), idesc);
                if (++di == nd) di = 0;
                if (++ai == n_a) ai = 0;
            }
        } else {
#pragma unroll 2
            for (int i = 0; i < iters; i += 3) {
                const uint32_t d = t0c + (uint32_t)(di * N);
                const uint64_t ah = mk_desc(a0 + (uint32_t)ai * 16384), al = mk_desc(a0 + (uint32_t)ai * 16384 + 64);
                const uint64_t bh = mk_desc(b0), bl = mk_desc(b0 + 64);
                mma<0>(d, al, bh, idesc);
                mma<1>(d, ah, bl, idesc);
                mma<2>(d, ah, bh, idesc);
                if (++di == nd) di = 0;
                if (++ai == n_a) ai = 0;
            }
        }
        long long t1 = clock64();
        commit(bar + warp);
        wait(bar + warp, 0);
        long long t2 = clock64();
        out[(blockIdx.x * 2 + warp) * 2 + 0] = t1 - t0;
        out[(blockIdx.x * 2 + warp) * 2 + 1] = t2 - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tbase));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (threadIdx.x == 0 && blockIdx.x == 0) result[0] = __uint_as_float(v);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}

int main() {
    long long *out;
    float *result;
    cudaMallocManaged(&out, 148 * 4 * sizeof(long long));
    cudaMallocManaged(&result, 64);
    const size_t smem = 1024 + (96 + 64) * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    printf("grid  N  accs mode issuers a_ops | issue clk/MMA  total clk/MMA   (tensor floor N/2)\n");
    const int iters = 11 * 12 * 24;
    for (int grid : {148})
        for (int N : {32, 64, 128})
            for (int nd : {2})
                for (int mode : {3, 4, 5, 6})
                    for (int issuers : {1})
                        for (int n_a : {4}) {
                            if (nd * N > 256) continue;
                            for (int rep = 0; rep < 2; ++rep) {
                                probe<<<grid, 128, smem>>>(N, nd, iters, mode, issuers, n_a * 16384, out, result);
                                cudaError_t e = cudaDeviceSynchronize();
                                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                            }
                            double mi = 0, mt = 0;
                            for (int c = 0; c < grid; ++c)
                                for (int w = 0; w < issuers; ++w) {
                                    mi = out[(c * 2 + w) * 2] > mi ? out[(c * 2 + w) * 2] : mi;
                                    mt = out[(c * 2 + w) * 2 + 1] > mt ? out[(c * 2 + w) * 2 + 1] : mt;
                                }
                            printf("%4d %3d %4d %4d %7d %5d | %12.1f %14.1f   (%d)  acc[0][0] %.6f expect %.6f\n", grid, N, nd, mode, issuers, n_a, mi / iters, mt / iters / 1.0, N / 2,
                                   result[0], (1 + (iters + nd - 1) / nd) / 1024.0);
                        }
    return 0;
}
