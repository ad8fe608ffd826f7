// Micro-probe: latency / throughput of cp.async.bulk (TMA) global->shared on B200, one issuing
// thread per CTA, 148 CTAs.  Usage: tma_probe        (prints a table)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const unsigned char *src, size_t per_cta_stride, int bytes, int depth, int iters, long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);          // [depth]
    unsigned char *buf = smem + 1024;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const unsigned char *base = src + (size_t)blockIdx.x * per_cta_stride;
    long long t0 = clock64();
    int issued = 0, done = 0;
    size_t off = 0;
    const size_t span = per_cta_stride ? per_cta_stride : (size_t)8 << 20;
    while (done < iters) {
        while (issued < iters && issued - done < depth) {
            const int s = issued % depth;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + s)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             s32(buf + (size_t)s * bytes)),
                         "l"(base + off), "r"(bytes), "r"(s32(bar + s))
                         : "memory");
            off += bytes;
            if (off + bytes > span) off = 0;
            ++issued;
        }
        const int s = done % depth;
        const uint32_t parity = (done / depth) & 1;
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok)
                         : "r"(s32(bar + s)), "r"(parity)
                         : "memory");
        ++done;
    }
    cycles[blockIdx.x] = clock64() - t0;
}

int main() {
    const size_t total = (size_t)1 << 30;
    unsigned char *src;
    cudaMalloc(&src, total);
    cudaMemset(src, 1, total);
    long long *cyc;
    cudaMallocManaged(&cyc, 148 * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("mode      bytes depth |  us/copy  GB/s/SM  TB/s total   (clock %.0f MHz)\n", clk_khz / 1e3);
    const int sizes[] = {2048, 4096, 16384, 32768, 49152};
    const int depths[] = {1, 2, 4, 8};
    for (int mode = 0; mode < 3; ++mode)          // 0: all CTAs same addresses (L2 hot), 1: distinct 4 MB windows (L2), 2: distinct, streaming 6 MB (HBM)
        for (int bytes : sizes)
            for (int depth : depths) {
                if ((size_t)depth * bytes > 200 * 1024) continue;
                const size_t stride = mode == 0 ? 0 : (mode == 1 ? (size_t)512 << 10 : (size_t)6 << 20);
                const int iters = 400;
                for (int rep = 0; rep < 2; ++rep) {
                    probe<<<148, 32, 1024 + (size_t)depth * bytes>>>(src, stride, bytes, depth, iters, cyc);
                    cudaDeviceSynchronize();
                }
                double mx = 0;
                for (int i = 0; i < 148; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
                const double us = mx / (clk_khz / 1e3) / iters;
                printf("%s %6d %5d | %8.3f %8.1f %8.2f\n", mode == 0 ? "same    " : mode == 1 ? "l2-dist " : "hbm-dist", bytes, depth, us,
                       bytes / us / 1e3, 148.0 * bytes / us / 1e6);
            }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return 0;
}
