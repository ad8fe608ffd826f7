// placeholder until the tcgen05 path lands: nothing is packed, nothing is supported, FFMA runs.
#include "conv_tc.cuh"

namespace svb {
int tc_pack_weights(const float *, int KS, int Cin, int CoutP, TcWeights *out, std::vector<void *> *) {
    out->KS = KS, out->Cin = Cin, out->CoutP = CoutP, out->ok = false;
    return SVB_OK;
}
bool tc_supported(const TcWeights &w, const ConvArgs &) { return w.ok; }
int launch_conv_tc(const TcWeights &, const ConvArgs &, int, cudaStream_t) {
    set_error("tensor-core conv path not built");
    return SVB_ERR_STATE;
}
}  // namespace svb
