// tcgen05 implicit-GEMM 1-D convolution on the C4T layout (sm_100a).
//
//   D[128 time rows x N columns] (fp32, TMEM) += A[128 x 8] (tf32, smem) * B[8 x N] (tf32, smem)
//
// GEMM mapping: M = time (128 rows per CTA), N = output channels (<= 256 per CTA), K = taps x Cin.
// * A operand = the activation slab.  C4T keeps, per channel quad, consecutive time steps as
//   consecutive 16-byte rows, so a [rows x 4 ch] slab is one contiguous span: it is fetched with ONE
//   cp.async.bulk (TMA, UBLKCP) per quad and it already IS the K-major no-swizzle UMMA core-matrix
//   layout (8 rows x 16 B = 128 B, SBO = 128 B, LBO = quad stride).  A conv tap at dilation d is a
//   row shift of the same slab = +16*k*d bytes on the descriptor start address, so one slab load
//   (128 + halo rows) feeds all KS taps.
// * B operand = weights, packed on the host per (column block, input-channel chunk, tap) in the
//   same core-matrix order, streamed through a ring of bulk copies.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
//   warps 2..5 = operand transform during the main loop (leaky-relu pre-activation of the ResBlock
//   + round-to-nearest tf32, optional hi/lo split for 3xTF32, in place in shared memory) and
//   epilogue afterwards (tcgen05.ld -> bias / residual / scale / accumulate -> 128-bit stores).
// * HBM tensors stay exact fp32; tf32 rounding happens only on the operand copy in shared memory.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "conv_tc.cuh"

namespace svb {

constexpr int kTcM = 128;       // rows per CTA (UMMA M)
constexpr int kTcCK = 32;       // input channels per chunk = 4 MMAs of K = 8
constexpr int kTcThreads = 192;

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
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
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

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// start address, leading byte offset (between the two 16-byte K halves of one MMA), stride byte
// offset (between 8-row core matrices), all in 16-byte units.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
    return d;                   // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ inline uint32_t umma_idesc_tf32(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                   // c_format = F32
    d |= 2u << 7;                   // a_format = TF32
    d |= 2u << 10;                  // b_format = TF32
    d |= (uint32_t)(N >> 3) << 17;  // n_dim
    d |= (uint32_t)(M >> 4) << 24;  // m_dim
    return d;
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

struct TcArgs {
    ConvArgs a;
    const float *w_hi, *w_lo;
    int n_tile, n_chunks, R, x3, nW, nA, tmem_cols;
    uint32_t slab_bytes, wtile_bytes;
};

__global__ void __launch_bounds__(kTcThreads) conv1d_c4_tc_kernel(TcArgs p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const ConvArgs &a = p.a;
    // ---- shared memory carve-up
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint64_t *a_full = bars, *a_ready = bars + 2, *a_empty = bars + 4;
    uint64_t *w_full = bars + 6, *w_empty = bars + 6 + 8;
    uint64_t *acc_full = bars + 6 + 16;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 6 + 16 + 1);
    unsigned char *slab0 = smem + 256;                               // [nA][(1 + x3)][slab_bytes]
    const uint32_t slab_slot = p.slab_bytes * (1 + p.x3);
    unsigned char *wring = slab0 + p.nA * slab_slot;                 // [nW][(1 + x3)][wtile_bytes]
    const uint32_t w_slot = p.wtile_bytes * (1 + p.x3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, nblk = blockIdx.y, t0 = blockIdx.x * kTcM;
    const int halo = (a.KS - 1) / 2 * a.dil;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) mbar_init(a_full + i, 1), mbar_init(a_ready + i, 128), mbar_init(a_empty + i, 1);
        for (int i = 0; i < 8; ++i) mbar_init(w_full + i, 1), mbar_init(w_empty + i, 1);
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            const int cin_q = a.Cin >> 2;
            const float *in_b = a.in + ((size_t)b * cin_q * a.in_Tp + (kPad + t0 - halo)) * 4;
            const size_t wtile_floats = (size_t)p.wtile_bytes / 4;
            for (int c = 0; c < p.n_chunks; ++c) {
                const int sA = c & (p.nA - 1);
                mbar_wait(a_empty + sA, ((c / p.nA) & 1) ^ 1);
                mbar_expect_tx(a_full + sA, p.slab_bytes);
                unsigned char *dst = slab0 + sA * slab_slot;
                const uint32_t qbytes = (uint32_t)p.R * 16;
                for (int q = 0; q < 8; ++q)
                    bulk_g2s(dst + q * qbytes, in_b + (size_t)(c * 8 + q) * a.in_Tp * 4, qbytes, a_full + sA);
                for (int k = 0; k < a.KS; ++k) {
                    const int it = c * a.KS + k, sW = it % p.nW;
                    mbar_wait(w_empty + sW, ((it / p.nW) & 1) ^ 1);
                    mbar_expect_tx(w_full + sW, w_slot);
                    const size_t off = (((size_t)nblk * p.n_chunks + c) * a.KS + k) * wtile_floats;
                    bulk_g2s(wring + sW * w_slot, p.w_hi + off, p.wtile_bytes, w_full + sW);
                    if (p.x3) bulk_g2s(wring + sW * w_slot + p.wtile_bytes, p.w_lo + off, p.wtile_bytes, w_full + sW);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(kTcM, p.n_tile);
            const uint32_t a_lbo = (uint32_t)p.R * 16, b_lbo = (uint32_t)p.n_tile * 16;
            uint32_t acc = 0;
            for (int c = 0; c < p.n_chunks; ++c) {
                const int sA = c & (p.nA - 1);
                mbar_wait(a_ready + sA, (c / p.nA) & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(slab0 + sA * slab_slot), a_lo = a_hi + p.slab_bytes;
                for (int k = 0; k < a.KS; ++k) {
                    const int it = c * a.KS + k, sW = it % p.nW;
                    mbar_wait(w_full + sW, (it / p.nW) & 1);
                    tc_fence_after();
                    const uint32_t b_hi = smem_u32(wring + sW * w_slot), b_lo = b_hi + p.wtile_bytes;
                    const uint32_t row_off = (uint32_t)(k * a.dil) * 16;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t ao = row_off + kk * 2 * a_lbo, bo = kk * 2 * b_lbo;
                        const uint64_t adh = umma_desc(a_hi + ao, a_lbo, 128), bdh = umma_desc(b_hi + bo, b_lbo, 128);
                        if (p.x3) {     // small terms first
                            umma_tf32(tmem_base, umma_desc(a_lo + ao, a_lbo, 128), bdh, idesc, acc);
                            acc = 1;
                            umma_tf32(tmem_base, adh, umma_desc(b_lo + bo, b_lbo, 128), idesc, acc);
                        }
                        umma_tf32(tmem_base, adh, bdh, idesc, acc);
                        acc = 1;
                    }
                    umma_commit(w_empty + sW);      // frees the weight slot when these MMAs retire
                }
                umma_commit(a_empty + sA);
            }
            umma_commit(acc_full);
        }
    } else {
        // ====================== transform (main loop) + epilogue warps ================
        const int tid = threadIdx.x - 64;                       // 0..127
        const int n4 = 8 * p.R;                                 // float4 rows in a slab
        for (int c = 0; c < p.n_chunks; ++c) {
            const int sA = c & (p.nA - 1);
            mbar_wait(a_full + sA, (c / p.nA) & 1);
            float4 *hi = reinterpret_cast<float4 *>(slab0 + sA * slab_slot);
            float4 *lo = reinterpret_cast<float4 *>(slab0 + sA * slab_slot + p.slab_bytes);
            for (int i = tid; i < n4; i += 128) {
                float4 v = lrelu4(hi[i], a.in_slope);
                const float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
                hi[i] = h;
                if (p.x3) lo[i] = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
            }
            fence_proxy_async();                                // generic-proxy writes -> visible to the tensor core
            mbar_arrive(a_ready + sA);
        }
        // ---- epilogue: TMEM lane = time row, column = output channel
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int lane_base = 32 * (warp & 3);                  // a warp may only touch its own TMEM lane quarter
        const int q = t0 + lane_base + lane;                    // GEMM row of this thread
        const int out_q = a.Cout >> 2;
        for (int j = 0; j < p.n_tile / 32; ++j) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(j * 32), v);
            if (q < a.Tq) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int cop = nblk * p.n_tile + j * 32 + 4 * g;
                    int phi = 0, co = cop;
                    if (a.ups_u > 0) { phi = cop / a.Cout; co = cop - phi * a.Cout; }
                    const float4 bv = __ldg(reinterpret_cast<const float4 *>(a.bias + co));
                    const size_t row = ((size_t)b * out_q + (co >> 2)) * a.out_Tp + kPad +
                                       (a.ups_u > 0 ? (size_t)q * a.ups_u + phi : (size_t)q);
                    float4 o = make_float4(v[4 * g] + bv.x, v[4 * g + 1] + bv.y, v[4 * g + 2] + bv.z, v[4 * g + 3] + bv.w);
                    if (a.res) {
                        const float4 rv = __ldg(reinterpret_cast<const float4 *>(a.res) + row);
                        o.x += rv.x, o.y += rv.y, o.z += rv.z, o.w += rv.w;
                    }
                    o.x *= a.out_scale, o.y *= a.out_scale, o.z *= a.out_scale, o.w *= a.out_scale;
                    float4 *op = reinterpret_cast<float4 *>(a.out) + row;
                    if (a.accumulate) {
                        const float4 old = *op;
                        o.x += old.x, o.y += old.y, o.z += old.z, o.w += old.w;
                    }
                    *op = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ------------------------------------------------------------------ host side
static float host_tf32(float x) {   // round to nearest, ties away (cvt.rna.tf32.f32)
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return x;
    u += 0x1000u;
    u &= 0xFFFFE000u;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}

static int pick_n_tile(int CoutP) {
    if (CoutP % 16 != 0) return 0;
    if (CoutP <= 256) return CoutP;
    for (int n = 256; n >= 16; n -= 16)
        if (CoutP % n == 0) return n;
    return 0;
}

int tc_pack_weights(const float *packed, int KS, int Cin, int CoutP, TcWeights *out, std::vector<void *> *allocs) {
    out->KS = KS, out->Cin = Cin, out->CoutP = CoutP, out->ok = false;
    const int n_tile = pick_n_tile(CoutP);
    if (n_tile == 0 || Cin % kTcCK != 0) return SVB_OK;           // FFMA handles it
    out->n_tile = n_tile;
    const int n_chunks = Cin / kTcCK, n_blk = CoutP / n_tile;
    const size_t tile = (size_t)n_tile * kTcCK;                    // floats per (nblk, chunk, tap)
    std::vector<float> hi((size_t)n_blk * n_chunks * KS * tile), lo(hi.size());
    for (int nb = 0; nb < n_blk; ++nb)
        for (int c = 0; c < n_chunks; ++c)
            for (int k = 0; k < KS; ++k) {
                float *th = hi.data() + (((size_t)nb * n_chunks + c) * KS + k) * tile;
                float *tl = lo.data() + (((size_t)nb * n_chunks + c) * KS + k) * tile;
                for (int q = 0; q < 8; ++q)                        // K quads of the chunk
                    for (int n = 0; n < n_tile; ++n)
                        for (int e = 0; e < 4; ++e) {
                            const int ci = c * kTcCK + q * 4 + e, co = nb * n_tile + n;
                            const float w = packed[((size_t)k * Cin + ci) * CoutP + co];
                            const float h = host_tf32(w);
                            th[((size_t)q * n_tile + n) * 4 + e] = h;
                            tl[((size_t)q * n_tile + n) * 4 + e] = host_tf32(w - h);
                        }
            }
    SVB_CUDA(cudaMalloc((void **)&out->hi, hi.size() * 4));
    allocs->push_back(out->hi);
    SVB_CUDA(cudaMalloc((void **)&out->lo, lo.size() * 4));
    allocs->push_back(out->lo);
    SVB_CUDA(cudaMemcpy(out->hi, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice));
    SVB_CUDA(cudaMemcpy(out->lo, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice));
    out->ok = true;
    return SVB_OK;
}

bool tc_supported(const TcWeights &w, const ConvArgs &a) {
    return w.ok && a.Cin % kTcCK == 0 && a.bias != nullptr && (a.KS - 1) / 2 * a.dil <= kPad &&
           (a.ups_u == 0 || a.Cout % 4 == 0);
}

int launch_conv_tc(const TcWeights &w, const ConvArgs &a, int precision, cudaStream_t st) {
    TcArgs p;
    p.a = a, p.w_hi = w.hi, p.w_lo = w.lo;
    p.n_tile = w.n_tile, p.n_chunks = a.Cin / kTcCK;
    const int halo = (a.KS - 1) / 2 * a.dil;
    p.R = kTcM + 2 * halo;
    p.x3 = precision == SVB_PREC_TF32X3 ? 1 : 0;
    p.slab_bytes = (uint32_t)8 * p.R * 16;
    p.wtile_bytes = (uint32_t)p.n_tile * kTcCK * 4;
    int cols = 32;
    while (cols < p.n_tile) cols <<= 1;
    p.tmem_cols = cols;
    p.nA = std::min(2, p.n_chunks);
    const size_t fixed = 256 + (size_t)p.nA * p.slab_bytes * (1 + p.x3);
    const size_t wslot = (size_t)p.wtile_bytes * (1 + p.x3);
    // wide tiles take the whole SM; narrow (bandwidth-bound) layers keep 2-3 CTAs per SM resident
    const size_t budget = p.n_tile > 128 ? 224 * 1024 : p.n_tile > 64 ? 110 * 1024 : 74 * 1024;
    int nW = fixed + 2 * wslot <= budget ? (int)((budget - fixed) / wslot) : 2;
    nW = std::max(1, std::min(std::min(nW, 8), p.n_chunks * a.KS));
    p.nW = nW;
    const size_t smem = fixed + (size_t)nW * wslot;
    SVB_CHECK(smem <= 227 * 1024, SVB_ERR_INVALID, "tc conv: tile does not fit shared memory (N %d, %zu B)", p.n_tile, smem);
    static size_t configured = 0;
    if (smem > configured) {
        SVB_CUDA(cudaFuncSetAttribute(conv1d_c4_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((a.Tq + kTcM - 1) / kTcM, a.CoutP / p.n_tile, a.B);
    conv1d_c4_tc_kernel<<<grid, kTcThreads, smem, st>>>(p);
    SVB_CUDA(cudaGetLastError());
    return SVB_OK;
}

}  // namespace svb
