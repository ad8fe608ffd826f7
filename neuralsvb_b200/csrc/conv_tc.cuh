// tcgen05 (5th-gen tensor core) implicit-GEMM convolution on the C4T layout.  See conv_tc.cu.
#pragma once
#include <vector>

#include "conv_ffma.cuh"

namespace svb {

struct TcWeights {
    void *blob[4] = {nullptr, nullptr, nullptr, nullptr};   // per svb_precision: weight tiles in UMMA core-matrix order
    int KS = 0, Cin = 0, CoutP = 0;
    int n_tile = 0;         // GEMM columns per CTA (UMMA N)
    bool ok = false;
};

// packed_ffma: [KS][Cin][CoutP] fp32 (the FFMA packing).  Allocations are appended to `allocs`.
// force_n_tile > 0: GEMM columns per CTA fixed by the caller (grouped layers: one tile = whole conv groups; `Cin` is then
// the number of input channels ONE column block contracts over and packed_ffma is [KS][Cin][CoutP])
int tc_pack_weights(const float *packed_ffma, int KS, int Cin, int CoutP, TcWeights *out, std::vector<void *> *allocs,
                    int force_n_tile = 0);
// same packing on the device, from a device copy of the FFMA packing into the blobs tc_pack_weights allocated
// (training: the weights change after every optimizer step)
int tc_repack_weights_dev(const float *packed_ffma_dev, const TcWeights &w, cudaStream_t st);
bool tc_supported(const TcWeights &w, const ConvArgs &a);
// max_ctas > 0 caps the persistent grid (used to run independent ResBlock chains side by side on SM subsets)
int launch_conv_tc(const TcWeights &w, const ConvArgs &a, int precision, cudaStream_t st, int max_ctas = 0);

}  // namespace svb
