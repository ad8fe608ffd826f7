// Weight gradient of the stride-1 (dilated) convolutions on tcgen05:
//     dW[ci][co][k] += sum_b sum_t act(A[b][t + k*dil - pad][ci]) * G[b][t][co]
// The reduction runs over TIME, so both MMA operands are "MN-major": in G32T a row is one time step holding 32
// channels, and after the in-place split a 128-byte row is [32 x bf16 hi | 32 x bf16 lo] = 64 M (or N) elements of
// ONE K index -- exactly the canonical MN-major SWIZZLE_128B atom (8 K-rows x 128 B, chunk ^ (row & 7)) that the
// forward kernel's operand transform already produces.  One MMA (M = 128: two channel groups of A, N = 64 or 128:
// one or two groups of G, K = 16 time steps) therefore yields all four split cross terms
//     [a_hi ; a_lo] x [g_hi | g_lo]  ->  hi*hi, hi*lo, lo*hi, lo*lo
// in separate accumulator blocks; the epilogue adds them.  A conv tap is a row (= K) offset on the A descriptor.
// CTA = (pair of A groups, pair of G groups, set of <= 4 taps) x a strided share of the (clip, 128-row) chunks;
// accumulators (<= 4 taps x 128 columns) stay in TMEM for the CTA's lifetime, one epilogue with fp32 atomics.
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..9 = operand transform, 2..5 also the epilogue.
#include <algorithm>
#include <cstdlib>

#include "tc_ptx.cuh"
#include "train_ops.cuh"

namespace svb {

namespace {

constexpr int kWgChunk = 128;               // time rows per pipeline stage (8 MMAs of K = 16 per tap)
constexpr int kWgTaps = 4;                  // taps per CTA (4 x 128 accumulator columns = the whole TMEM)
constexpr int kWgStages = 2;
constexpr int kWgThreads = 320;

struct WgTcArgs {
    WgradArgs a;
    int gA, gG, n_pairs, n_cosets, n_tapsets, splits, chunks_per_b;
    int n_eg, gA_g, gG_g;                   // grouped conv: effective groups, G32T channel groups of A / G per effective group
    int Rx, Rxp;                            // rows of an A slab (chunk + tap reach), rounded up to whole 8-row atoms
    uint32_t x_tile_bytes, stage_bytes;
};

// [32 fp32] row -> [32 bf16 hi | 32 bf16 lo] in place, chunk c stored at c ^ (row & 7)   (as conv_tc.cu's transform)
__device__ __forceinline__ void split_rows_inplace(unsigned char *tile, int rows, float slope, int tid) {
    uint4 *op = reinterpret_cast<uint4 *>(tile);
    const int cl = tid & 7;
    const bool odd = cl & 1;
    for (int r0 = 0; r0 < rows; r0 += 32) {
        const int r = r0 + (tid >> 3);
        const bool ok = r < rows;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = *reinterpret_cast<const float4 *>(op + (size_t)r * 8 + cl);
        v.x = fmaxf(v.x, v.x * slope), v.y = fmaxf(v.y, v.y * slope), v.z = fmaxf(v.z, v.z * slope), v.w = fmaxf(v.w, v.w * slope);
        const __nv_bfloat162 hA = __floats2bfloat162_rn(v.x, v.y), hB = __floats2bfloat162_rn(v.z, v.w);
        const uint32_t h0 = *reinterpret_cast<const uint32_t *>(&hA), h1 = *reinterpret_cast<const uint32_t *>(&hB);
        const __nv_bfloat162 lA = __floats2bfloat162_rn(v.x - __uint_as_float(h0 << 16), v.y - __uint_as_float(h0 & 0xffff0000u));
        const __nv_bfloat162 lB = __floats2bfloat162_rn(v.z - __uint_as_float(h1 << 16), v.w - __uint_as_float(h1 & 0xffff0000u));
        const uint32_t l0 = *reinterpret_cast<const uint32_t *>(&lA), l1 = *reinterpret_cast<const uint32_t *>(&lB);
        // even lane keeps hi and receives the neighbour's hi; odd lane keeps lo (the shuffles also order reads before writes)
        const uint32_t g0 = __shfl_xor_sync(0xffffffffu, odd ? h0 : l0, 1);
        const uint32_t g1 = __shfl_xor_sync(0xffffffffu, odd ? h1 : l1, 1);
        if (ok) {
            uint4 *row = op + (size_t)r * 8;
            const int sw = r & 7;
            if (!odd) row[(cl >> 1) ^ sw] = make_uint4(h0, h1, g0, g1);
            else row[(4 + (cl >> 1)) ^ sw] = make_uint4(g0, g1, l0, l1);
        }
    }
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(WgTcArgs p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const WgradArgs &a = p.a;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint64_t *full = bars, *ready = bars + kWgStages, *empty = bars + 2 * kWgStages, *done = bars + 3 * kWgStages;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(done + 1);
    unsigned char *stage0 = smem + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- this CTA's class and share of the chunks
    int cls = blockIdx.x / p.splits;
    const int sp = blockIdx.x - cls * p.splits;
    const int tapset = cls % p.n_tapsets;
    cls /= p.n_tapsets;
    const int coset = cls % p.n_cosets;
    cls /= p.n_cosets;
    const int pair = cls % p.n_pairs, eg = cls / p.n_pairs;          // eg: effective conv group (0 when dense)
    const int k0 = tapset * kWgTaps, ntaps = min(kWgTaps, a.K - k0);
    const int nxa = min(2, p.gA_g - pair * 2), ng = min(2, p.gG_g - coset * 2);    // real channel groups of this CTA
    const int ga0 = eg * p.gA_g + pair * 2, gg0 = eg * p.gG_g + coset * 2;          // first G32T channel group of A / G
    const int N = 64 * ng;
    const int n_units = a.B * p.chunks_per_b;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWgStages; ++i) mbar_init(full + i, 1), mbar_init(ready + i, 8), mbar_init(empty + i, 1);
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto x_tile = [&](int s, int g) { return stage0 + (size_t)s * p.stage_bytes + (size_t)g * p.x_tile_bytes; };
    auto g_tile = [&](int s, int j) { return stage0 + (size_t)s * p.stage_bytes + 2 * (size_t)p.x_tile_bytes + (size_t)j * (kWgChunk * 128); };

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int s = 0, ph = 1;                                          // "empty" barriers start free
            const uint32_t bytes = (uint32_t)nxa * p.Rx * 128 + (uint32_t)ng * kWgChunk * 128;
            for (int u = sp; u < n_units; u += p.splits) {
                const int b = u / p.chunks_per_b, t0 = (u - b * p.chunks_per_b) * kWgChunk;
                mbar_wait(empty + s, ph);
                mbar_expect_tx(full + s, bytes);
                for (int g = 0; g < nxa; ++g)
                    bulk_g2s(x_tile(s, g), a.A + (((size_t)b * p.gA + ga0 + g) * a.TpA + kPad + t0 - a.pa) * 32, (uint32_t)p.Rx * 128,
                             full + s);
                for (int j = 0; j < ng; ++j)
                    bulk_g2s(g_tile(s, j), a.G + (((size_t)b * p.gG + gg0 + j) * a.TpG + kPad + t0) * 32, kWgChunk * 128, full + s);
                if (++s == kWgStages) s = 0, ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        const uint32_t elected = elect_one_sync();
        // instruction descriptor: bf16 x bf16 -> fp32, BOTH operands MN-major (bits 15 / 16)
        const uint32_t idesc = umma_idesc(1, 128, N) | (1u << 15) | (1u << 16);
        const uint32_t hi_word = desc_hi_sw128(0);                      // SBO = 1024 B between 8-row K atoms
        const uint32_t lbo_a = nxa == 2 ? (p.x_tile_bytes >> 4) : 0u;   // second 64-row M block = next channel group (or the same)
        const uint32_t lbo_b = (uint32_t)(kWgChunk * 128) >> 4;
        int s = 0, ph = 0;
        uint32_t fresh = 1;
        bool any = false;
        // converged-warp issue (tc_ptx.cuh, "issue discipline"): every lane walks the loops, MMAs / commits are predicated
        for (int u = sp; u < n_units; u += p.splits) {
            mbar_wait(ready + s, ph);
            __syncwarp();
            tc_fence_after();
            const uint32_t xa = smem_u32(x_tile(s, 0)), gb = smem_u32(g_tile(s, 0));
#pragma unroll 1
            for (int ks = 0; ks < kWgChunk / 16; ++ks) {
                const uint32_t b_lo = (((gb + (uint32_t)ks * 16 * 128) >> 4) & 0x3FFFu) | (lbo_b << 16);
#pragma unroll 1
                for (int tk = 0; tk < ntaps; ++tk) {
                    const uint32_t row = (uint32_t)ks * 16 + (uint32_t)((k0 + tk) * a.da);
                    const uint32_t a_lo = (((xa + row * 128) >> 4) & 0x3FFFu) | (lbo_a << 16);
                    umma<true>(tmem_base + (uint32_t)(tk * 128), a_lo, hi_word, b_lo, hi_word, idesc, fresh ^ 1u, elected);
                }
                fresh = 0;
            }
            umma_commit(empty + s, elected);
            any = true;
            if (++s == kWgStages) s = 0, ph ^= 1;
        }
        umma_commit(done, elected);
        (void)any;
    } else {
        // ====================== operand transform warps (2..9), then epilogue (2..5) ==================
        const int tid = threadIdx.x - 64;                               // 0..255
        int s = 0, ph = 0;
        for (int u = sp; u < n_units; u += p.splits) {
            mbar_wait(full + s, ph);
            for (int g = 0; g < nxa; ++g) split_rows_inplace(x_tile(s, g), p.Rx, a.slope, tid);
            for (int j = 0; j < ng; ++j) split_rows_inplace(g_tile(s, j), kWgChunk, 1.f, tid);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(ready + s);
            if (++s == kWgStages) s = 0, ph ^= 1;
        }
        if (warp < 6 && sp < n_units) {
            mbar_wait(done, 0);
            tc_fence_after();
            const int lane_base = 32 * (warp & 3);
            const int m = lane_base + lane;                             // accumulator row
            const int gm = m >> 6, ci = ((ga0 + gm) << 5) + (m & 31);
            const bool row_ok = gm < nxa && ci < a.Ca;
            const int ca_g = a.Ca / a.cgroups, cg_g = a.Cg / a.cgroups;            // channels per conv group
            const int cgrp = ci / ca_g, ci_l = ci - cgrp * ca_g;                   // conv group / local channel of this row
            for (int tk = 0; tk < ntaps; ++tk)
                for (int j = 0; j < ng; ++j) {
                    float vh[32], vl[32];
                    const uint32_t col = (uint32_t)(tk * 128 + j * 64);
                    tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + col, vh);
                    tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + col + 32, vl);
                    if (!row_ok) continue;
                    if (a.cgroups == 1) {
                        float *o = a.out + (long long)ci * a.s_ci + (long long)(k0 + tk) * a.s_k;
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const int co = ((gg0 + j) << 5) + c;
                            if (co < a.Cg) atomicAdd(o + (long long)co * a.s_co, vh[c] + vl[c]);
                        }
                    } else {
                        // grouped: only the block diagonal exists; polyphase rows map back to natural taps
                        int c_nat = ci_l, k_nat = k0 + tk;
                        if (a.poly_s > 0) {
                            c_nat = ci_l / a.poly_s;
                            k_nat = (k0 + tk) * a.poly_s + (ci_l - c_nat * a.poly_s);
                            if (k_nat >= a.K_nat) continue;
                        }
                        const int cin_nat = a.poly_s > 0 ? ca_g / a.poly_s : ca_g, K_out = a.poly_s > 0 ? a.K_nat : a.K;
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const int co = ((gg0 + j) << 5) + c;
                            if (co < a.Cg && co / cg_g == cgrp)
                                atomicAdd(a.out + ((long long)co * cin_nat + c_nat) * K_out + k_nat, vh[c] + vl[c]);
                        }
                    }
                }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool wgrad_tc_supported(const WgradArgs &a) {
    static int on = -1;
    if (on < 0) on = getenv("SVB_WGRAD_TC") ? atoi(getenv("SVB_WGRAD_TC")) : 1;
    if (!on) return false;
    const int reach = (a.K - 1) * a.da;
    return a.sa == 1 && a.sb == 1 && a.db == 0 && a.pb == 0 && a.da >= 1 && a.pa >= 0 && a.pa <= kPad && reach - a.pa <= kPad &&
           a.slope >= 0.f && a.slope <= 1.f && a.Ca >= 32 && a.Cg >= 32 &&
           (a.cgroups == 1 || (a.Ca % a.cgroups == 0 && a.Cg % a.cgroups == 0 && a.Ca % 32 == 0 && a.Cg % 32 == 0));
}

int launch_wgrad_tc(const WgradArgs &a, cudaStream_t st) {
    WgTcArgs p;
    p.a = a;
    p.gA = c4t_groups(a.Ca), p.gG = c4t_groups(a.Cg);
    p.n_eg = 1, p.gA_g = p.gA, p.gG_g = p.gG;
    if (a.cgroups > 1) {
        // effective groups: merge conv groups until both sides own whole 32-channel G32T groups (the epilogue masks the
        // cross-group products of a merged block)
        int merge = 1;
        while ((a.Ca / a.cgroups * merge) % 32 != 0 || (a.Cg / a.cgroups * merge) % 32 != 0) merge *= 2;
        SVB_CHECK(a.cgroups % merge == 0, SVB_ERR_INVALID, "wgrad_tc: %d conv groups of %d x %d channels do not tile", a.cgroups,
                  a.Ca / a.cgroups, a.Cg / a.cgroups);
        p.n_eg = a.cgroups / merge;
        p.gA_g = p.gA / p.n_eg, p.gG_g = p.gG / p.n_eg;
    }
    p.n_pairs = (p.gA_g + 1) / 2, p.n_cosets = (p.gG_g + 1) / 2, p.n_tapsets = (a.K + kWgTaps - 1) / kWgTaps;
    p.chunks_per_b = (a.Tq + kWgChunk - 1) / kWgChunk;
    const int classes = p.n_eg * p.n_pairs * p.n_cosets * p.n_tapsets;
    const int units = a.B * p.chunks_per_b;
    p.splits = std::max(1, std::min(units, 148 / std::max(1, classes)));
    p.Rx = kWgChunk + (a.K - 1) * a.da;
    p.Rxp = round_up(p.Rx, 8);
    p.x_tile_bytes = (uint32_t)p.Rxp * 128;
    p.stage_bytes = 2 * p.x_tile_bytes + 2 * kWgChunk * 128;
    const size_t smem = 1024 + (size_t)kWgStages * p.stage_bytes;
    SVB_CHECK(smem <= 227 * 1024, SVB_ERR_INVALID, "wgrad_tc: slab of %d rows does not fit shared memory", p.Rx);
    static size_t configured = 0;
    if (smem > configured) {
        SVB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    wgrad_tc_kernel<<<classes * p.splits, kWgThreads, smem, st>>>(p);
    SVB_CUDA(cudaGetLastError());
    return SVB_OK;
}

}  // namespace svb
