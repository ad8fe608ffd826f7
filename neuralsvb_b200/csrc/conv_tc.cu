// tcgen05 implicit-GEMM 1-D convolution on the G32T activation layout (sm_100a).
//
//   D[128 time rows x N columns] (fp32, TMEM) += A[128 x 16] (bf16, smem) * B[16 x N] (bf16, smem)
//
// GEMM mapping: M = time (128 rows per MMA), N = output channels (<= 128 per CTA), K = taps x Cin.
// * A operand = the activation slab.  G32T keeps, per 32-channel group, consecutive time steps as
//   consecutive 128-byte rows, so the [rows x 32 ch] slab is one contiguous span fetched with ONE
//   cp.async.bulk (TMA, UBLKCP) straight into the operand slot.  The transform warps rewrite it IN
//   PLACE into the MMA operand: leaky-relu pre-activation, hi/lo split (or tf32 rounding) and the
//   SWIZZLE_128B chunk permutation; one 128-byte K-major row per time step.  A conv tap at
//   dilation d is a ROW SHIFT of that operand = +128*k*d bytes on the descriptor start address,
//   so one slab (MT*128 + halo rows) feeds all KS taps of MT accumulator tiles.
// * B operand = weights, packed and pre-swizzled on the host per (column block, input-channel
//   chunk, tap), streamed through a ring of bulk copies (or kept resident for narrow layers).
// * Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + elected-lane
//   MMA issuer, warps 2..5 = operand transform, warps 6..13 = epilogue (tcgen05.ld -> bias /
//   residual / scale / accumulate -> 128-byte row stores) on the accumulator set the MMAs are not
//   writing.  HBM tensors stay exact fp32; rounding happens only on the operand copy in smem.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>

#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace svb {

constexpr int kTcM = 128;       // rows per CTA (UMMA M)
constexpr int kTcCK = 32;       // input channels per chunk = 4 MMAs of K = 8


// Operand modes (svb_precision): 1 = 1xTF32, 2 = 3xTF32, 3 = 3xBF16 (hi/lo split, 16-bit mantissa).
// One operand row = 32 input channels = 128 bytes:  TF32: 32 x tf32 ; BF16x3: [32 x bf16 hi | 32 x bf16 lo].
struct TcArgs {
    ConvArgs a;
    const unsigned char *w;     // packed, pre-swizzled weight tiles of this mode
    int n_tile, n_chunks, MT, R, Rp, nW, nA, tmem_cols;
    int n_sets;                 // accumulator sets in TMEM: 2 = epilogue of group g overlaps MMAs of g+1; 1 = MT can be twice as large
    int col_blocks, groups_per_b, total_groups;
    int tps, n_st, w_resident;  // taps per weight stage, stages per chunk, whole layer resident in smem
    int blk_chunk_step;         // grouped conv: first input chunk of column block nblk = nblk * blk_chunk_step (0: dense)
    uint32_t wstage_bytes;      // ring slot = tps weight tiles
    int pdl;                    // launched with programmatic stream serialization
    int collect;                // A-operand collector reuse between the two products of a_hi (SVB_TC_COLLECT, default on)
    long long *stats;           // SVB_TC_STATS: [grid][kTcStatSlots] blocked-cycle counters (diagnostics build of the kernel)
    int dbg;                    // SVB_TC_DBG bit mask: 1 no MMAs, 2 hi*hi only, 4 no transform, 8 no epilogue ld/st
    uint32_t raw_bytes;         // fp32 slab as TMA delivers it: R rows x 128 B
    uint32_t op_bytes;          // operand slot: Rp rows x 128 B (x2 with the 3xTF32 lo plane)
    uint32_t wtile_bytes;       // one (column block, chunk, tap) weight tile (x2 with the 3xTF32 lo plane)
    uint32_t off_op, off_w, off_stage, off_bias;   // byte offsets of operand slots / weight ring / epilogue tiles / bias in dynamic smem
};

constexpr int kMaxW = 8;
constexpr int kMaxDevices = 64;     // per-device caches of function attributes / SM counts
// Warp roles.  The SM's issue arbiter favours HIGH warp ids, so the two latency-critical single-thread
// roles get the highest ids: warps 0-7 epilogue, 8-15 operand transform, 16 TMA producer, 17 MMA issuer.
constexpr int kTcThreadsP = 576;
constexpr int kWarpProducer = 16, kWarpMma = 17, kWarpTransform0 = 8;
constexpr int kStageBytes = 8 * 4096;   // epilogue transpose tiles: 32 rows x 128 B per epilogue warp
constexpr int kBiasBytes = 4096;        // bias of the layer (Cout <= 1024 floats; wider layers take 12 KB)

// Persistent kernel: one CTA per SM walks "groups" (MT consecutive 128-row tiles of one clip for
// one column block).  Every stage is decoupled by mbarriers, so the TMA producer runs ahead into
// the next group, the transform warps prepare operands while the tensor core works on the
// previous chunk, and the epilogue drains accumulator set (g & 1) from TMEM while the MMAs of
// group g + 1 fill the other set.
// STATS (SVB_TC_STATS=1, diagnostics only): every role accumulates the cycles it spends blocked on each of its
// barriers; the launcher prints the per-CTA mean / max.  Says which stage of the pipeline the others wait for.
template <bool ST>
__device__ __forceinline__ void mbar_wait_t(uint64_t *bar, uint32_t parity, long long &acc) {
    if (ST) {
        const long long t = clock64();
        mbar_wait(bar, parity);
        acc += clock64() - t;
    } else {
        mbar_wait(bar, parity);
    }
}
constexpr int kTcStatSlots = 16;

template <int MODE, bool ST = false>
__global__ void __launch_bounds__(kTcThreadsP, 1) conv1d_c4_tc_kernel(TcArgs p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr bool BF = MODE == SVB_PREC_BF16X3;
    constexpr bool X3 = MODE == SVB_PREC_TF32X3;
    const ConvArgs &a = p.a;
    // ---- shared memory carve-up (operand slots and weight ring are 1024-byte aligned: swizzle atoms)
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint64_t *raw_full = bars, *a_ready = bars + 4, *a_empty = bars + 8;      // up to 4 operand slots
    uint64_t *w_full = bars + 12, *w_empty = bars + 12 + kMaxW;
    uint64_t *acc_full = bars + 12 + 2 * kMaxW, *acc_empty = acc_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_empty + 2);
    unsigned char *op0 = smem + p.off_op;                            // [nA][op_bytes]: TMA target AND MMA operand
    unsigned char *wring = smem + p.off_w;                           // [nW][wtile_bytes]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int halo = (a.KS - 1) / 2 * a.dil;
    const int acc_cols = p.MT * p.n_tile;                            // columns of one accumulator set
    long long st_w0 = 0, st_w1 = 0, st_w2 = 0;                       // STATS: cycles blocked on up to three barriers
    const long long st_t0 = ST ? clock64() : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(raw_full + i, 1), mbar_init(a_ready + i, 8), mbar_init(a_empty + i, 1);
        for (int i = 0; i < 2; ++i) mbar_init(acc_full + i, 1), mbar_init(acc_empty + i, 8);
        for (int i = 0; i < kMaxW; ++i) mbar_init(w_full + i, 1), mbar_init(w_empty + i, 1);
        fence_barrier_init();
    }
    if (warp == kWarpMma) tmem_alloc(tmem_ptr, p.tmem_cols);
    {   // bias is constant data (not produced by the previous kernel): stage it once, before the PDL wait
        float *sb = reinterpret_cast<float *>(smem + p.off_bias);
        for (int i = threadIdx.x; i < a.Cout; i += kTcThreadsP) sb[i] = __ldg(a.bias + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation) overlapped the
    // tail of the previous kernel in the stream; its results are needed from here on.  The next
    // kernel may begin ITS prologue as soon as every CTA of this grid has reached this point.
    if (p.pdl) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }

    // group id -> (column block, clip, first row); consecutive ids are neighbours in time
    auto decode = [&](int g, int &nblk, int &b, int &t0) {
        const int tg = g % p.groups_per_b;
        const int r = g / p.groups_per_b;
        b = r % a.B, nblk = r / a.B, t0 = tg * (kTcM * p.MT);
    };

    // The two single-thread roles below are latency-critical (every instruction they execute sits
    // between two TMA copies or two MMAs), so all ring indices / phases are kept as incrementing
    // counters -- no runtime integer divisions in the loops.
    if (warp == kWarpProducer) {
        // ================================ TMA producer ================================
        // One bulk copy per slab (R rows x 128 B, contiguous in G32T) and one per weight stage.
        if (lane == 0) {
            const int gin = c4t_groups(a.Cin);
            const size_t group_stride = (size_t)a.in_Tp * 32;                     // floats between channel groups
            const size_t chunk_w_bytes = (size_t)a.KS * p.wtile_bytes;
            int sA = 0, phA = 1, sW = 0, phW = 1;                                 // "empty" barriers start free
            int tg = blockIdx.x % p.groups_per_b, rb = blockIdx.x / p.groups_per_b;   // group -> (time group, b + B*nblk)
            const int step_tg = gridDim.x % p.groups_per_b, step_rb = gridDim.x / p.groups_per_b;
            bool first_group = true;
            for (int g = blockIdx.x; g < p.total_groups; g += gridDim.x) {
                const int b = rb % a.B, nblk = rb / a.B;
                const int t0 = tg * (kTcM * p.MT);
                const float *in_c = a.in + (((size_t)b * gin + (size_t)nblk * p.blk_chunk_step) * a.in_Tp + (kPad + t0 - halo)) * 32;
                const unsigned char *w_c = p.w + (size_t)nblk * p.n_chunks * chunk_w_bytes;
                for (int c = 0; c < p.n_chunks; ++c) {
                    mbar_wait_t<ST>(a_empty + sA, phA, st_w0);                    // MMAs of the slot's previous slab retired
                    mbar_expect_tx(raw_full + sA, p.raw_bytes);
                    bulk_g2s(op0 + sA * p.op_bytes, in_c, p.raw_bytes, raw_full + sA);
                    in_c += group_stride;
                    if (++sA == p.nA) sA = 0, phA ^= 1;
                    if (!(p.w_resident && !first_group)) {
                        const unsigned char *w_k = w_c;
                        for (int st = 0; st < p.n_st; ++st) {
                            const int nt = min(p.tps, a.KS - st * p.tps);
                            const uint32_t bytes = (uint32_t)nt * p.wtile_bytes;
                            mbar_wait_t<ST>(w_empty + sW, phW, st_w1);
                            mbar_expect_tx(w_full + sW, bytes);
                            bulk_g2s(wring + sW * p.wstage_bytes, w_k, bytes, w_full + sW);
                            w_k += bytes;
                            if (++sW == p.nW) sW = 0, phW ^= 1;
                        }
                    }
                    w_c += chunk_w_bytes;
                }
                first_group = false;
                tg += step_tg, rb += step_rb;
                if (tg >= p.groups_per_b) tg -= p.groups_per_b, ++rb;
            }
            if (ST) {
                long long *o = p.stats + (size_t)blockIdx.x * kTcStatSlots;
                o[0] = st_w0, o[1] = st_w1, o[2] = clock64() - st_t0;
            }
        }
    } else if (warp == kWarpMma) {
        // ================================ MMA issuer ==================================
        // The whole warp walks the loop CONVERGED and every MMA / commit is predicated on the elected lane, so the
        // descriptors stay in uniform registers (tc_ptx.cuh, "issue discipline": a divergent single-thread issuer
        // costs 65-82 cycles per MMA, more than the MMA itself for every N below 256).
        const uint32_t elected = elect_one_sync();
        const uint32_t idesc = umma_idesc(BF ? 1 : 2, kTcM, p.n_tile);
        const uint32_t hi_word = desc_hi_sw128(0);
        // Descriptors are advanced in their ENCODED form (address >> 4 in the low 14 bits; every offset below is a
        // multiple of 16 bytes and the sum stays inside shared memory, so the field cannot overflow): one uniform
        // add per operand and MMA -- the issue loop, not the tensor pipe, sets the pace of the narrow layers.
        const uint32_t a_lo_plane = (X3 ? p.op_bytes / 2 : 64u) >> 4;      // lo plane / half-row
        const uint32_t b_lo_plane = (X3 ? p.wtile_bytes / 2 : 64u) >> 4;
        const uint32_t op_base = desc_lo(smem_u32(op0)), w_base = desc_lo(smem_u32(wring));
        const uint32_t tap_step = (uint32_t)a.dil * (128 >> 4);            // a tap is a row shift of the operand
        const uint32_t op_step = p.op_bytes >> 4, wstage_step = p.wstage_bytes >> 4, wtile_step = p.wtile_bytes >> 4;
        // the ablation switches exist only in the diagnostics instantiation
        const bool mma_on = ST ? !(p.dbg & 1) : true, hh_only = ST ? (p.dbg & 2) != 0 : false, collect = ST ? p.collect != 0 : true;
        int sA = 0, phA = 0, sW = 0, phW = 0, as = 0, phE = 1;
        bool first_group = true;
        for (int g = blockIdx.x; g < p.total_groups; g += gridDim.x) {
            mbar_wait_t<ST>(acc_empty + as, phE, st_w0);            // epilogue has drained this accumulator set
            __syncwarp();
            tc_fence_after();
            const uint32_t d_set = tmem_base + (uint32_t)(as * acc_cols);
            uint32_t fresh = 1;                                     // first MMA of the group overwrites
            for (int c = 0; c < p.n_chunks; ++c) {
                mbar_wait_t<ST>(a_ready + sA, phA, st_w1);
                __syncwarp();
                tc_fence_after();
                uint32_t a_tap = op_base + sA * op_step;
                int in_stage = 0;
                uint32_t b_tap = 0;
                for (int k = 0; k < a.KS; ++k) {
                    if (in_stage == 0) {                            // first tap of a weight stage
                        if (!(p.w_resident && !first_group)) {
                            mbar_wait_t<ST>(w_full + sW, phW, st_w2);
                            __syncwarp();
                            tc_fence_after();
                        }
                        b_tap = w_base + sW * wstage_step;
                    }
                    if (mma_on) {
                        uint32_t d = d_set, a_row = a_tap;
#pragma unroll 1
                        for (int m = 0; m < p.MT; ++m) {
                            const uint32_t acc = fresh ^ 1u;
                            if (BF) {                               // 2 x 16 channels; small cross terms first
                                if (hh_only) {
                                    umma<true>(d, a_row, hi_word, b_tap, hi_word, idesc, acc, elected);
                                    umma<true>(d, a_row + 2, hi_word, b_tap + 2, hi_word, idesc, 1u, elected);
                                } else if (collect) {               // a_hi is read from shared memory once for its two products
                                    umma<true>(d, a_row + a_lo_plane, hi_word, b_tap, hi_word, idesc, acc, elected);
                                    umma<true, 1>(d, a_row, hi_word, b_tap + b_lo_plane, hi_word, idesc, 1u, elected);
                                    umma<true, 2>(d, a_row, hi_word, b_tap, hi_word, idesc, 1u, elected);
                                    umma<true>(d, a_row + a_lo_plane + 2, hi_word, b_tap + 2, hi_word, idesc, 1u, elected);
                                    umma<true, 1>(d, a_row + 2, hi_word, b_tap + b_lo_plane + 2, hi_word, idesc, 1u, elected);
                                    umma<true, 2>(d, a_row + 2, hi_word, b_tap + 2, hi_word, idesc, 1u, elected);
                                } else {
                                    umma<true>(d, a_row + a_lo_plane, hi_word, b_tap, hi_word, idesc, acc, elected);
                                    umma<true>(d, a_row, hi_word, b_tap + b_lo_plane, hi_word, idesc, 1u, elected);
                                    umma<true>(d, a_row, hi_word, b_tap, hi_word, idesc, 1u, elected);
                                    umma<true>(d, a_row + a_lo_plane + 2, hi_word, b_tap + 2, hi_word, idesc, 1u, elected);
                                    umma<true>(d, a_row + 2, hi_word, b_tap + b_lo_plane + 2, hi_word, idesc, 1u, elected);
                                    umma<true>(d, a_row + 2, hi_word, b_tap + 2, hi_word, idesc, 1u, elected);
                                }
                            } else {
#pragma unroll
                                for (uint32_t kb = 0; kb < 8; kb += 2) {   // 4 x 8 channels (32 bytes = 2 descriptor units each)
                                    const uint32_t acc_k = kb == 0 ? acc : 1u;
                                    if (X3) {
                                        umma<false>(d, a_row + a_lo_plane + kb, hi_word, b_tap + kb, hi_word, idesc, acc_k, elected);
                                        umma<false, 1>(d, a_row + kb, hi_word, b_tap + b_lo_plane + kb, hi_word, idesc, 1u, elected);
                                        umma<false, 2>(d, a_row + kb, hi_word, b_tap + kb, hi_word, idesc, 1u, elected);
                                    } else {
                                        umma<false>(d, a_row + kb, hi_word, b_tap + kb, hi_word, idesc, acc_k, elected);
                                    }
                                }
                            }
                            d += (uint32_t)p.n_tile, a_row += (kTcM * 128) >> 4;
                        }
                    }
                    fresh = 0;
                    a_tap += tap_step, b_tap += wtile_step;
                    if (++in_stage == p.tps || k + 1 == a.KS) {     // stage consumed: release its ring slot
                        in_stage = 0;
                        if (!p.w_resident) {
                            umma_commit(w_empty + sW, elected);
                            if (++sW == p.nW) sW = 0, phW ^= 1;
                        }
                    }
                }
                umma_commit(a_empty + sA, elected);
                if (++sA == p.nA) sA = 0, phA ^= 1;
            }
            umma_commit(acc_full + as, elected);
            if (++as == p.n_sets) as = 0, phE ^= 1;
            first_group = false;
        }
        if (ST && elected) {
            long long *o = p.stats + (size_t)blockIdx.x * kTcStatSlots;
            o[3] = st_w0, o[4] = st_w1, o[5] = st_w2, o[6] = clock64() - st_t0;
        }
    } else if (warp >= kWarpTransform0) {
        // ====================== operand transform warps (256 threads) =================
        // In place, 8 lanes per 128-byte row (one 16-byte chunk = 4 channels each): conflict-free
        // LDS.128 / STS.128.  Row r, logical chunk j is stored at chunk position j ^ (r & 7)
        // (SWIZZLE_128B).  bf16 mode: lane pairs exchange halves so the even lane writes the
        // 8-channel hi chunk (j = c/2) and the odd lane the lo chunk (4 + c/2).
        const int tid = threadIdx.x - kWarpTransform0 * 32;     // 0..255
        const int cl = tid & 7;                                 // chunk of the row this lane reads
        const bool odd = cl & 1;
        const float slope = a.in_slope;                         // 0 <= slope <= 1: lrelu(x) = max(x, slope * x)
        int sA = 0, phA = 0;
        for (int g = blockIdx.x; g < p.total_groups; g += gridDim.x) {
            for (int c = 0; c < p.n_chunks; ++c) {
                mbar_wait_t<ST>(raw_full + sA, phA, st_w0);
                uint4 *op = reinterpret_cast<uint4 *>(op0 + sA * p.op_bytes);
                // 64 rows per pass over the 256 threads: two independent rows per thread for ILP
                for (int r0 = 0; r0 < ((p.dbg & 4) ? 0 : p.R); r0 += 64) {
                    int rr[2];
                    bool ok[2];
                    float4 v[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        rr[u] = r0 + 32 * u + (tid >> 3);
                        ok[u] = rr[u] < p.R;
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok[u]) v[u] = *reinterpret_cast<const float4 *>(op + (size_t)rr[u] * 8 + cl);
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        float4 &x = v[u];
                        x.x = fmaxf(x.x, x.x * slope), x.y = fmaxf(x.y, x.y * slope);
                        x.z = fmaxf(x.z, x.z * slope), x.w = fmaxf(x.w, x.w * slope);
                    }
                    if (BF) {
                        uint32_t h0[2], h1[2], l0[2], l1[2], g0[2], g1[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const __nv_bfloat162 hA = __floats2bfloat162_rn(v[u].x, v[u].y), hB = __floats2bfloat162_rn(v[u].z, v[u].w);
                            h0[u] = *reinterpret_cast<const uint32_t *>(&hA), h1[u] = *reinterpret_cast<const uint32_t *>(&hB);
                            const __nv_bfloat162 lA = __floats2bfloat162_rn(v[u].x - __uint_as_float(h0[u] << 16),
                                                                           v[u].y - __uint_as_float(h0[u] & 0xffff0000u));
                            const __nv_bfloat162 lB = __floats2bfloat162_rn(v[u].z - __uint_as_float(h1[u] << 16),
                                                                           v[u].w - __uint_as_float(h1[u] & 0xffff0000u));
                            l0[u] = *reinterpret_cast<const uint32_t *>(&lA), l1[u] = *reinterpret_cast<const uint32_t *>(&lB);
                        }
                        // even lane keeps hi and receives the neighbour's hi; odd lane keeps lo (shuffles also order reads before writes)
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            g0[u] = __shfl_xor_sync(0xffffffffu, odd ? h0[u] : l0[u], 1);
                            g1[u] = __shfl_xor_sync(0xffffffffu, odd ? h1[u] : l1[u], 1);
                        }
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            if (!ok[u]) continue;
                            uint4 *row = op + (size_t)rr[u] * 8;
                            const int sw = rr[u] & 7;
                            if (!odd) row[(cl >> 1) ^ sw] = make_uint4(h0[u], h1[u], g0[u], g1[u]);          // channels 8j..8j+7 hi
                            else row[(4 + (cl >> 1)) ^ sw] = make_uint4(g0[u], g1[u], l0[u], l1[u]);          // channels 8j..8j+7 lo
                        }
                    } else {
                        __syncwarp();
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            if (!ok[u]) continue;
                            uint4 *row = op + (size_t)rr[u] * 8;
                            const int sw = rr[u] & 7;
                            const float4 h = make_float4(to_tf32(v[u].x), to_tf32(v[u].y), to_tf32(v[u].z), to_tf32(v[u].w));
                            row[cl ^ sw] = make_uint4(__float_as_uint(h.x), __float_as_uint(h.y), __float_as_uint(h.z), __float_as_uint(h.w));
                            if (X3)
                                (row + p.op_bytes / 32)[cl ^ sw] =
                                    make_uint4(__float_as_uint(to_tf32(v[u].x - h.x)), __float_as_uint(to_tf32(v[u].y - h.y)),
                                               __float_as_uint(to_tf32(v[u].z - h.z)), __float_as_uint(to_tf32(v[u].w - h.w)));
                        }
                    }
                }
                fence_proxy_async();                                // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready + sA);           // one arrival per warp (8 per slab)
                if (++sA == p.nA) sA = 0, phA ^= 1;
            }
        }
        if (ST && tid == 0) {
            long long *o = p.stats + (size_t)blockIdx.x * kTcStatSlots;
            o[7] = st_w0, o[8] = clock64() - st_t0;
        }
    } else {
        // ================================ epilogue warps (256 threads) ================
        // TMEM lane = time row, column = output channel; a warp may only touch lanes 32*(warp%4)..+31,
        // so two warps share each lane quarter and alternate over the 32-column blocks.  A block is
        // 32 rows x 128 B, contiguous in G32T.  Global accesses use the "wide" mapping (lane l, step i
        // -> row 4i + l/8, chunk l%8: 512 contiguous bytes per instruction); TMEM hands each thread
        // one whole row, so a per-warp 4 KB swizzled smem tile transposes between the two views.
        // The residual block for step n+1 is requested before block n is drained.
        const int ew = warp;                                        // 0..7
        const int lane_base = 32 * (warp & 3);
        const int half = ew >> 2;                                   // 0 / 1: which blocks this warp takes
        const int gout = c4t_groups(a.Cout);
        const float4 *res4 = reinterpret_cast<const float4 *>(a.res);
        float4 *out4 = reinterpret_cast<float4 *>(a.out);
        float4 *tile = reinterpret_cast<float4 *>(smem + p.off_stage + ew * 4096);   // [32 rows][8 chunks], chunk ^ (row & 7)
        const int jb = p.n_tile / 32, nblocks = p.MT * jb;
        const int wr = lane >> 3, wc = lane & 7;                    // wide mapping: row offset / chunk
        const bool no_mem = p.dbg & 8;
        int gi = 0;
        const size_t rstep = (size_t)(a.ups_u > 0 ? a.ups_u : 1) * 8;      // float4 between consecutive GEMM rows
        float4 rres[8];
        // (group, block) -> float4 index of row (q_base + 0) of its 32-row x 128-byte tile; rows are rstep apart;
        // q_base = first GEMM row of this warp in the block
        auto block_base_g = [&](int g, int blk, int &co0, int &q_base) -> size_t {
            int nblk, b, t0;
            decode(g, nblk, b, t0);
            const int m = blk / jb, j = blk - m * jb;
            q_base = t0 + m * kTcM + lane_base;
            const int cop0 = nblk * p.n_tile + j * 32;              // 32 columns never straddle an upsampler phase
            int phi = 0;
            co0 = cop0;
            if (a.ups_u > 0) { phi = cop0 / a.Cout; co0 = cop0 - phi * a.Cout; }
            return (((size_t)b * gout + (co0 >> 5)) * a.out_Tp + kPad +
                    (a.ups_u > 0 ? (size_t)q_base * a.ups_u + phi : (size_t)q_base)) * 8;
        };
        // residual rows of one block in the wide mapping -> registers (latency hidden behind the current block)
        auto fetch_res = [&](int g, int blk) {
            if (!res4 || no_mem || g >= p.total_groups) return;
            int co0, q_base;
            const size_t base = block_base_g(g, blk, co0, q_base);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = 4 * i + wr;
                rres[i] = (q_base + rr < a.Tq) ? __ldg(res4 + base + (size_t)rr * rstep + wc) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        if (half < nblocks) fetch_res(blockIdx.x, half);
        for (int g = blockIdx.x; g < p.total_groups; g += gridDim.x, ++gi) {
            int nblk, b, t0;
            decode(g, nblk, b, t0);
            const int as = p.n_sets == 2 ? (gi & 1) : 0;
            auto block_base = [&](int blk, int &co0, int &q_base) -> size_t {
                const int m = blk / jb, j = blk - m * jb;
                q_base = t0 + m * kTcM + lane_base;
                const int cop0 = nblk * p.n_tile + j * 32;          // 32 columns never straddle an upsampler phase
                int phi = 0;
                co0 = cop0;
                if (a.ups_u > 0) { phi = cop0 / a.Cout; co0 = cop0 - phi * a.Cout; }
                return (((size_t)b * gout + (co0 >> 5)) * a.out_Tp + kPad +
                        (a.ups_u > 0 ? (size_t)q_base * a.ups_u + phi : (size_t)q_base)) * 8;
            };
            mbar_wait_t<ST>(acc_full + as, (p.n_sets == 2 ? (gi >> 1) : gi) & 1, st_w0);
            tc_fence_after();
            for (int blk = half; blk < nblocks; blk += 2) {
                const int m = blk / jb, j = blk - m * jb;
                int co0, q_base;
                const size_t base = block_base(blk, co0, q_base);
                if (res4) {                                          // residual: wide registers -> tile
#pragma unroll
                    for (int i = 0; i < 8; ++i) tile[(4 * i + wr) * 8 + (wc ^ ((4 * i + wr) & 7))] = rres[i];
                    __syncwarp();
                }
                // prefetch the residual of this warp's NEXT block: same group, or the first one of its next group
                if (blk + 2 < nblocks) fetch_res(g, blk + 2);
                else fetch_res(g + gridDim.x, half);
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(as * acc_cols + m * p.n_tile + j * 32), v);
                // own row: accumulator + bias (+ residual) -> tile
#pragma unroll
                for (int gq = 0; gq < 8; ++gq) {
                    const float4 bv = *reinterpret_cast<const float4 *>(smem + p.off_bias + (size_t)(co0 + 4 * gq) * 4);
                    float4 o = make_float4(v[4 * gq] + bv.x, v[4 * gq + 1] + bv.y, v[4 * gq + 2] + bv.z, v[4 * gq + 3] + bv.w);
                    float4 *slot = tile + lane * 8 + (gq ^ (lane & 7));
                    if (res4) {
                        const float4 rv = *slot;
                        o.x += rv.x, o.y += rv.y, o.z += rv.z, o.w += rv.w;
                    }
                    o.x *= a.out_scale, o.y *= a.out_scale, o.z *= a.out_scale, o.w *= a.out_scale;
                    *slot = o;
                }
                __syncwarp();
                // tile -> global, wide mapping (512 contiguous bytes per store instruction)
                if (!no_mem) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = 4 * i + wr;
                        if (q_base + rr >= a.Tq) continue;
                        float4 o = tile[rr * 8 + (wc ^ (rr & 7))];
                        float4 *dst = out4 + base + (size_t)rr * rstep + wc;
                        if (a.accumulate) {
                            const float4 old = *dst;
                            o.x += old.x, o.y += old.y, o.z += old.z, o.w += old.w;
                        }
                        *dst = o;
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + as);             // this accumulator set may be overwritten
        }
        if (ST && threadIdx.x == 0) {
            long long *o = p.stats + (size_t)blockIdx.x * kTcStatSlots;
            o[9] = st_w0, o[10] = clock64() - st_t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (ST && threadIdx.x == 0) p.stats[(size_t)blockIdx.x * kTcStatSlots + 11] = clock64() - st_t0;
    if (warp == kWarpMma) {
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
static uint16_t host_bf16(float x) {   // round to nearest even (__float2bfloat16_rn)
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float bf16_to_float(uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}

static int pick_n_tile(int CoutP) {      // columns per CTA: a multiple of 32 (one G32T channel group per epilogue block)
    if (CoutP % 32 != 0) return 0;
    if (CoutP <= 128) return CoutP;
    for (int n = 128; n >= 32; n -= 32)
        if (CoutP % n == 0) return n;
    return 0;
}

// Weight tiles in K-major SWIZZLE_128B order: row n (output column) = 128 bytes holding the 32 input
// channels of the chunk; the 16-byte chunk c of row n is stored at chunk position c ^ (n & 7).
int tc_pack_weights(const float *packed, int KS, int Cin, int CoutP, TcWeights *out, std::vector<void *> *allocs, int force_n_tile) {
    out->KS = KS, out->Cin = Cin, out->CoutP = CoutP, out->ok = false;
    const int n_tile = force_n_tile > 0 ? force_n_tile : pick_n_tile(CoutP);
    if (force_n_tile > 0 && (CoutP % force_n_tile != 0 || force_n_tile % 32 != 0 || force_n_tile > 128)) return SVB_OK;
    if (n_tile == 0 || n_tile % 8 != 0 || Cin % 4 != 0) return SVB_OK;    // CUDA cores handle it
    out->n_tile = n_tile;
    // input channels are padded to whole 32-channel groups (zero weights against the zero pad channels of G32T)
    const int n_chunks = (Cin + kTcCK - 1) / kTcCK, n_blk = CoutP / n_tile;
    const size_t tile_b = (size_t)n_tile * 128;                    // bytes per tile plane
    const size_t n_tiles = (size_t)n_blk * n_chunks * KS;
    std::vector<unsigned char> tf(n_tiles * tile_b), tf3(n_tiles * tile_b * 2), bf(n_tiles * tile_b);
    for (int nb = 0; nb < n_blk; ++nb)
        for (int c = 0; c < n_chunks; ++c)
            for (int k = 0; k < KS; ++k) {
                const size_t t = ((size_t)nb * n_chunks + c) * KS + k;
                float *t1 = reinterpret_cast<float *>(tf.data() + t * tile_b);
                float *t3h = reinterpret_cast<float *>(tf3.data() + t * tile_b * 2);
                float *t3l = reinterpret_cast<float *>(tf3.data() + t * tile_b * 2 + tile_b);
                uint16_t *tb = reinterpret_cast<uint16_t *>(bf.data() + t * tile_b);
                for (int n = 0; n < n_tile; ++n)
                    for (int ci = 0; ci < kTcCK; ++ci) {
                        const int cin_i = c * kTcCK + ci;
                        const float w = cin_i < Cin ? packed[((size_t)k * Cin + cin_i) * CoutP + nb * n_tile + n] : 0.f;
                        const float h = host_tf32(w);
                        const int c16 = ci >> 2;                   // fp32: 4 channels per 16-byte chunk
                        const size_t i4 = (size_t)n * 32 + ((c16 ^ (n & 7)) << 2) + (ci & 3);
                        t1[i4] = h, t3h[i4] = h, t3l[i4] = host_tf32(w - h);
                        const uint16_t bh = host_bf16(w), bl = host_bf16(w - bf16_to_float(bh));
                        const int ch = ci >> 3, cl = 4 + (ci >> 3); // bf16: 8 channels per chunk; hi chunks 0-3, lo 4-7
                        tb[(size_t)n * 64 + ((ch ^ (n & 7)) << 3) + (ci & 7)] = bh;
                        tb[(size_t)n * 64 + ((cl ^ (n & 7)) << 3) + (ci & 7)] = bl;
                    }
            }
    auto up = [&](const void *src, size_t bytes, void **dst) -> int {
        SVB_CUDA(cudaMalloc(dst, bytes));
        allocs->push_back(*dst);
        SVB_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return SVB_OK;
    };
    SVB_TRY(up(tf.data(), tf.size(), &out->blob[SVB_PREC_TF32]));
    SVB_TRY(up(tf3.data(), tf3.size(), &out->blob[SVB_PREC_TF32X3]));
    SVB_TRY(up(bf.data(), bf.size(), &out->blob[SVB_PREC_BF16X3]));
    out->ok = true;
    return SVB_OK;
}

// device twin of the loops in tc_pack_weights: one thread per (tile, output column n, input channel ci)
__global__ void tc_repack_kernel(const float *__restrict__ packed, int KS, int Cin, int CoutP, int n_tile, int n_chunks,
                                 long long total, unsigned char *__restrict__ tf, unsigned char *__restrict__ tf3,
                                 unsigned char *__restrict__ bf) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ci = (int)(i % kTcCK);
    const int n = (int)((i / kTcCK) % n_tile);
    const long long t = i / ((long long)kTcCK * n_tile);
    const int k = (int)(t % KS);
    const int c = (int)((t / KS) % n_chunks);
    const int nb = (int)(t / ((long long)KS * n_chunks));
    const int cin_i = c * kTcCK + ci;
    const float w = cin_i < Cin ? packed[((size_t)k * Cin + cin_i) * CoutP + nb * n_tile + n] : 0.f;
    auto tf32 = [](float x) {      // host_tf32: round half up in magnitude to 10 mantissa bits
        uint32_t u = __float_as_uint(x);
        if ((u & 0x7F800000u) == 0x7F800000u) return x;
        u += 0x1000u;
        u &= 0xFFFFE000u;
        return __uint_as_float(u);
    };
    const size_t tile_b = (size_t)n_tile * 128;
    const float h = tf32(w);
    const size_t i4 = (size_t)n * 32 + (((ci >> 2) ^ (n & 7)) << 2) + (ci & 3);
    reinterpret_cast<float *>(tf + t * tile_b)[i4] = h;
    reinterpret_cast<float *>(tf3 + t * tile_b * 2)[i4] = h;
    reinterpret_cast<float *>(tf3 + t * tile_b * 2 + tile_b)[i4] = tf32(w - h);
    const __nv_bfloat16 bh = __float2bfloat16_rn(w);
    const __nv_bfloat16 bl = __float2bfloat16_rn(w - __bfloat162float(bh));
    uint16_t *tb = reinterpret_cast<uint16_t *>(bf + t * tile_b);
    const int ch = ci >> 3, cl = 4 + (ci >> 3);
    tb[(size_t)n * 64 + ((ch ^ (n & 7)) << 3) + (ci & 7)] = __bfloat16_as_ushort(bh);
    tb[(size_t)n * 64 + ((cl ^ (n & 7)) << 3) + (ci & 7)] = __bfloat16_as_ushort(bl);
}

int tc_repack_weights_dev(const float *packed_dev, const TcWeights &w, cudaStream_t st) {
    if (!w.ok) return SVB_OK;
    const int n_chunks = (w.Cin + kTcCK - 1) / kTcCK, n_blk = w.CoutP / w.n_tile;
    const long long total = (long long)n_blk * n_chunks * w.KS * w.n_tile * kTcCK;
    tc_repack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        packed_dev, w.KS, w.Cin, w.CoutP, w.n_tile, n_chunks, total, reinterpret_cast<unsigned char *>(w.blob[SVB_PREC_TF32]),
        reinterpret_cast<unsigned char *>(w.blob[SVB_PREC_TF32X3]), reinterpret_cast<unsigned char *>(w.blob[SVB_PREC_BF16X3]));
    SVB_CUDA(cudaGetLastError());
    return SVB_OK;
}

bool tc_supported(const TcWeights &w, const ConvArgs &a) {
    if (a.cin_blk > 0 && (a.cin_blk % 32 != 0 || a.ups_u != 0 || (a.CoutP / w.n_tile) * a.cin_blk != a.Cin || w.Cin != a.cin_blk)) return false;
    return w.ok && a.Cin % 4 == 0 && a.bias != nullptr && (a.KS - 1) / 2 * a.dil <= kPad &&
           (a.ups_u == 0 || a.Cout % 32 == 0) && a.Cout <= 3072;
}

template <int MODE, bool ST>
static int launch_mode(const TcArgs &p, int grid, size_t smem, cudaStream_t st) {
    auto kern = conv1d_c4_tc_kernel<MODE, ST>;
    // function attributes are per device: one process may drive several GPUs (the reference's mp.spawn gives one each,
    // but nothing in the C ABI forbids a handle per device in one process)
    static size_t configured[kMaxDevices] = {};
    static bool carve[kMaxDevices] = {};
    int dev = 0;
    SVB_CUDA(cudaGetDevice(&dev));
    dev = dev < kMaxDevices ? dev : kMaxDevices - 1;
    if (smem > configured[dev] || dev == kMaxDevices - 1) {
        SVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    if (!carve[dev]) {   // keep the SM's smem/L1 split fixed across the differently-sized launches of a forward
        SVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        carve[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kTcThreadsP), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = p.pdl ? 1 : 0;
    cfg.attrs = attr, cfg.numAttrs = 1;
    SVB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return SVB_OK;
}

static int sm_count() {
    static int n[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev < kMaxDevices ? dev : kMaxDevices - 1;
    if (n[slot] == 0 || slot == kMaxDevices - 1) {
        cudaDeviceGetAttribute(&n[slot], cudaDevAttrMultiProcessorCount, dev);
        if (n[slot] <= 0) n[slot] = 148;
    }
    return n[slot];
}

int launch_conv_tc(const TcWeights &w, const ConvArgs &a, int precision, cudaStream_t st, int max_ctas) {
    SVB_CHECK(precision >= SVB_PREC_TF32 && precision <= SVB_PREC_BF16X3, SVB_ERR_INVALID, "tc conv: bad precision %d",
              precision);
    TcArgs p;
    p.a = a, p.w = reinterpret_cast<const unsigned char *>(w.blob[precision]);
    const int cin_eff = a.cin_blk > 0 ? a.cin_blk : a.Cin;     // input channels one column block contracts over
    p.n_tile = w.n_tile, p.n_chunks = (cin_eff + kTcCK - 1) / kTcCK;
    p.blk_chunk_step = a.cin_blk > 0 ? p.n_chunks : 0;
    const int halo = (a.KS - 1) / 2 * a.dil;
    const int tiles = (a.Tq + kTcM - 1) / kTcM;              // 128-row tiles per clip
    p.col_blocks = a.CoutP / p.n_tile;
    const int planes = precision == SVB_PREC_TF32X3 ? 2 : 1;
    p.wtile_bytes = (uint32_t)p.n_tile * 128 * planes;
    p.dbg = 0;
    if (const char *e = getenv("SVB_TC_DBG")) p.dbg = atoi(e);
    p.collect = 1;
    if (const char *e = getenv("SVB_TC_COLLECT")) p.collect = atoi(e) != 0;
    p.pdl = 1;
    if (const char *e = getenv("SVB_TC_PDL")) p.pdl = atoi(e) != 0;
    int force_mt = 0;
    if (const char *e = getenv("SVB_TC_MT")) force_mt = atoi(e);
    // ---- M tiles per group: each weight tile fetched from L2 feeds MT accumulators.  Two accumulator
    // sets live in TMEM (2 * MT * N <= 512 columns); slabs and the weight ring must fit shared memory.
    size_t smem = 0;
    const size_t bias_bytes = a.Cout <= 1024 ? kBiasBytes : 3 * kBiasBytes;
    const size_t budget = 226 * 1024 - kStageBytes - bias_bytes;
    int force_sets = 0;
    if (const char *e = getenv("SVB_TC_SETS")) force_sets = atoi(e);
    const long long tile_units = (long long)tiles * p.col_blocks * a.B;
    for (int MT : {4, 2, 1}) {
        if (force_mt && MT != force_mt && MT != 1) continue;
        // Two accumulator sets (epilogue overlapped) need 2*MT*N <= 512 TMEM columns.  Wide layers
        // (N = 128) that are bound by weight streaming from L2 take MT = 4 with ONE set instead:
        // every weight tile then feeds 512 rows.  Needs enough groups to keep all SMs busy.
        int sets = 2 * MT * p.n_tile <= 512 ? 2 : 1;
        if (force_sets) sets = force_sets;
        if (sets * MT * p.n_tile > 512) continue;
        if (!force_mt) {
            if (MT == 4) continue;      // measured: MT = 4 (one or two accumulator sets) is slower than MT = 2 on every layer
            (void)tile_units;
            if (MT == 2 && tiles < 2) continue;
        }
        p.n_sets = sets;
        if ((tiles + MT - 1) / MT * MT * kTcM > round_up(a.Tq, kTileT) && MT != 1) continue;   // stay inside the allocation
        p.MT = MT;
        p.R = MT * kTcM + 2 * halo;
        p.Rp = round_up(p.R, 8);
        p.raw_bytes = (uint32_t)p.R * 128;
        p.op_bytes = (uint32_t)p.Rp * 128 * planes;
        // operand slots double as TMA targets: 3 of them keep two slab copies in flight behind the MMAs
        p.nA = 3;
        if (const char *e = getenv("SVB_TC_NA")) p.nA = std::max(2, std::min(4, atoi(e)));
        p.off_op = 1024;
        p.off_w = p.off_op + (uint32_t)p.nA * p.op_bytes;
        if (p.off_w + 2 * (size_t)p.wtile_bytes > budget) {
            p.nA = 2;
            p.off_w = p.off_op + (uint32_t)p.nA * p.op_bytes;
        }
        if (p.off_w + 2 * (size_t)p.wtile_bytes > budget && MT != 1) continue;
        SVB_CHECK(p.off_w + (size_t)p.wtile_bytes <= budget, SVB_ERR_INVALID, "tc conv: tile does not fit shared memory (N %d)",
                  p.n_tile);
        // weight stage = as many taps of one chunk as fit in ~48 KB, fetched by a single bulk copy
        const size_t avail = budget - p.off_w;
        // (measured: one bulk copy per tap beats multi-tap stages -- a single large copy streams slowly)
        int tps = 1;
        if (const char *e = getenv("SVB_TC_TPS")) tps = std::max(1, std::min(a.KS, atoi(e)));
        while (tps > 1 && (size_t)tps * p.wtile_bytes * 2 > avail) --tps;
        if ((size_t)a.KS * p.wtile_bytes <= avail && p.n_chunks == 1 && p.col_blocks == 1) tps = a.KS;   // whole layer fits
        p.tps = tps;
        p.n_st = (a.KS + tps - 1) / tps;
        p.wstage_bytes = (uint32_t)tps * p.wtile_bytes;
        p.w_resident = (p.n_chunks * p.n_st == 1 && p.col_blocks == 1) ? 1 : 0;
        int nW = (int)(avail / p.wstage_bytes);
        nW = std::max(1, std::min(std::min(nW, kMaxW), p.w_resident ? 1 : 1 << 30));
        p.nW = nW;
        p.off_stage = p.off_w + (uint32_t)nW * p.wstage_bytes;
        p.off_bias = p.off_stage + kStageBytes;
        smem = (size_t)p.off_bias + bias_bytes;
        break;
    }
    if (getenv("SVB_TC_VERBOSE"))
        fprintf(stderr, "[tc] Cin %d CoutP %d KS %d dil %d Tq %d | n_tile %d MT %d sets %d R %d nA %d tps %d n_st %d nW %d resident %d smem %zu\n",
                a.Cin, a.CoutP, a.KS, a.dil, a.Tq, p.n_tile, p.MT, p.n_sets, p.R, p.nA, p.tps, p.n_st, p.nW, p.w_resident, smem);
    int cols = 32;
    while (cols < p.n_sets * p.MT * p.n_tile) cols <<= 1;
    p.tmem_cols = cols;
    p.groups_per_b = (tiles + p.MT - 1) / p.MT;
    p.total_groups = p.groups_per_b * a.B * p.col_blocks;
    const int grid = std::min(p.total_groups, max_ctas > 0 ? std::min(max_ctas, sm_count()) : sm_count());
    p.stats = nullptr;
    const bool want_stats = getenv("SVB_TC_STATS") != nullptr;
    if ((want_stats || p.dbg != 0 || !p.collect) && precision == SVB_PREC_BF16X3) {
        // diagnostics: run the instrumented instantiation, wait for it and print where each role was blocked
        const size_t n = (size_t)grid * kTcStatSlots;
        SVB_CUDA(cudaMalloc((void **)&p.stats, n * sizeof(long long)));
        SVB_CUDA(cudaMemsetAsync(p.stats, 0, n * sizeof(long long), st));
        SVB_TRY((launch_mode<SVB_PREC_BF16X3, true>(p, grid, smem, st)));
        SVB_CUDA(cudaStreamSynchronize(st));
        std::vector<long long> h(n);
        SVB_CUDA(cudaMemcpy(h.data(), p.stats, n * sizeof(long long), cudaMemcpyDeviceToHost));
        SVB_CUDA(cudaFree(p.stats));
        double mean[kTcStatSlots] = {};
        long long mx[kTcStatSlots] = {};
        for (int c = 0; c < grid; ++c)
            for (int i = 0; i < kTcStatSlots; ++i) mean[i] += (double)h[(size_t)c * kTcStatSlots + i] / grid, mx[i] = std::max(mx[i], h[(size_t)c * kTcStatSlots + i]);
        if (want_stats) fprintf(stderr,
                "[tc-stats] C%d>%d k%d d%d Tq %d grid %d groups %d MT %d | kcycles mean (max): total %.0f (%.0f) | producer: a_empty %.0f w_empty %.0f of %.0f | "
                "mma: acc_empty %.0f a_ready %.0f w_full %.0f of %.0f | transform: raw_full %.0f of %.0f | epilogue: acc_full %.0f of %.0f\n",
                a.Cin, a.CoutP, a.KS, a.dil, a.Tq, grid, p.total_groups, p.MT, mean[11] / 1e3, mx[11] / 1e3, mean[0] / 1e3, mean[1] / 1e3, mean[2] / 1e3,
                mean[3] / 1e3, mean[4] / 1e3, mean[5] / 1e3, mean[6] / 1e3, mean[7] / 1e3, mean[8] / 1e3, mean[9] / 1e3, mean[10] / 1e3);
        return SVB_OK;
    }
    switch (precision) {
        case SVB_PREC_TF32: return launch_mode<SVB_PREC_TF32, false>(p, grid, smem, st);
        case SVB_PREC_TF32X3: return launch_mode<SVB_PREC_TF32X3, false>(p, grid, smem, st);
        default: return launch_mode<SVB_PREC_BF16X3, false>(p, grid, smem, st);
    }
}

}  // namespace svb
