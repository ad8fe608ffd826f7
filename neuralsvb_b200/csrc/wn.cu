// WN gated dilated-conv stack (first piece of the SVB acoustic step): the decoder / encoder body of the
// GlobalFVAE that produces the mel `spec2wav` consumes.  Reference: modules/fastspeech/fs2_vae.py:19-91 (class WN),
// used by modules/voice_conversion/vae_models.py:81-146 with hidden 192, kernel 5, dilation_rate 1, 4 / 8 layers.
//
//   g' = cond_layer(g)                                       1x1 conv, gin -> 2*H*L            (:73-74)
//   for i in layers:  x_in = in_layers[i](x)                 Conv1d(H, 2H, k, dilation d^i)    (:77)
//                     acts = tanh(x_in[:H] + g'_i[:H]) * sigmoid(x_in[H:] + g'_i[H:])          (:11-17, :80-86)
//                     rs   = res_skip_layers[i](acts)        1x1 conv, H -> 2H (H on the last) (:88)
//                     x = (x + rs[:H]) * mask ; out += rs[H:]      (last: out += rs)           (:89-93)
//   return out * mask                                                                          (:94)
//
// All three convolutions run on the generator's tcgen05 kernel (conv_tc.cu) over the G32T layout; the gate and the
// residual / skip update are element-wise kernels on the same layout.  Inference only (eval mode: dropout is identity).
#include <map>
#include <string>
#include <vector>

#include "generator.cuh"

using namespace svb;

namespace {

struct WnConv {
    int Cin = 0, Cout = 0, K = 1, dil = 1;
    float *w = nullptr, *b = nullptr;
    TcWeights tc;
};

// acts[c] = tanh(xin[c] + cond[c]) * sigmoid(xin[H + c] + cond[H + c])   (G32T: channel groups of 32)
__global__ void wn_gate_kernel(const float *__restrict__ xin, const float *__restrict__ cond, int cond_g0, int cond_groups, int gh, int T,
                               int Tp, float *__restrict__ acts) {
    const int lane = threadIdx.x & 31, b = blockIdx.z, grp = blockIdx.y;
    for (int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < T; t += gridDim.x * (blockDim.x >> 5)) {
        const size_t row = (size_t)(kPad + t) * 32 + lane;
        float a = xin[((size_t)b * 2 * gh + grp) * Tp * 32 + row], s = xin[((size_t)b * 2 * gh + gh + grp) * Tp * 32 + row];
        if (cond) {
            a += cond[((size_t)b * cond_groups + cond_g0 + grp) * Tp * 32 + row];
            s += cond[((size_t)b * cond_groups + cond_g0 + gh + grp) * Tp * 32 + row];
        }
        acts[((size_t)b * gh + grp) * Tp * 32 + row] = tanhf(a) * (1.f / (1.f + expf(-s)));
    }
}

// not last: x = (x + rs[:H]) * mask ; out += rs[H:]      last: out = (out + rs) * mask
__global__ void wn_res_skip_kernel(const float *__restrict__ rs, float *__restrict__ x, float *__restrict__ out, const float *__restrict__ mask,
                                   int gh, int T, int Tp, int last) {
    const int lane = threadIdx.x & 31, b = blockIdx.z, grp = blockIdx.y;
    const int grs = last ? gh : 2 * gh;
    for (int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < T; t += gridDim.x * (blockDim.x >> 5)) {
        const size_t row = (size_t)(kPad + t) * 32 + lane;
        const float m = mask ? mask[(size_t)b * T + t] : 1.f;
        const size_t ih = ((size_t)b * gh + grp) * Tp * 32 + row;
        if (!last) {
            x[ih] = (x[ih] + rs[((size_t)b * grs + grp) * Tp * 32 + row]) * m;
            out[ih] += rs[((size_t)b * grs + gh + grp) * Tp * 32 + row];
        } else {
            out[ih] = (out[ih] + rs[((size_t)b * grs + grp) * Tp * 32 + row]) * m;
        }
    }
}

// x *= mask   (FVAEDecoder: pre_net output, fs2_vae.py:149)
__global__ void wn_mask_kernel(float *__restrict__ x, const float *__restrict__ mask, int gh, int T, int Tp) {
    const int lane = threadIdx.x & 31, b = blockIdx.z, grp = blockIdx.y;
    for (int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < T; t += gridDim.x * (blockDim.x >> 5))
        x[((size_t)b * gh + grp) * Tp * 32 + (size_t)(kPad + t) * 32 + lane] *= mask[(size_t)b * T + t];
}

}  // namespace

struct svb_wn {
    int device = 0, precision = SVB_PREC_BF16X3;
    int H = 0, K = 1, dil_rate = 1, L = 0, gin = 0;
    bool finalized = false;
    std::map<std::string, HostTensor> host_w;
    std::vector<WnConv> in_layers, rs_layers;
    WnConv cond;
    // FVAEDecoder wrapper (fs2_vae.py:130-152): pre_net = ConvTranspose1d(latent, H, k = s, stride = s), out_proj = Conv1d(H, out, 1)
    int latent = 0, out_ch = 0, stride = 0;
    WnConv pre, post;
    std::vector<void *> allocs;
    // workspace (G32T): x, xin (2H), acts, rs (2H), out, g (gin), cond (2HL)
    char *ws = nullptr;
    size_t ws_cap = 0;
    int ws_B = 0, ws_T = 0;
};

namespace {

int wn_get(svb_wn *w, const std::string &name, std::vector<int64_t> want, const HostTensor **out) {
    auto it = w->host_w.find(name);
    SVB_CHECK(it != w->host_w.end(), SVB_ERR_MISSING, "wn: weight '%s' was never set", name.c_str());
    bool ok = it->second.shape.size() == want.size();
    for (size_t i = 0; ok && i < want.size(); ++i) ok = it->second.shape[i] == want[i];
    SVB_CHECK(ok, SVB_ERR_INVALID, "wn: weight '%s' has the wrong shape", name.c_str());
    *out = &it->second;
    return SVB_OK;
}

int wn_pack(svb_wn *w, const std::string &prefix, int Cin, int Cout, int K, int dil, WnConv *c) {
    const HostTensor *wt, *bt;
    SVB_TRY(wn_get(w, prefix + ".weight", {Cout, Cin, K}, &wt));
    SVB_TRY(wn_get(w, prefix + ".bias", {Cout}, &bt));
    const std::vector<float> p = pack_conv_weights(wt->data.data(), Cout, Cin, K);
    c->Cin = Cin, c->Cout = Cout, c->K = K, c->dil = dil;
    SVB_CUDA(cudaMalloc((void **)&c->w, p.size() * 4));
    w->allocs.push_back(c->w);
    SVB_CUDA(cudaMemcpy(c->w, p.data(), p.size() * 4, cudaMemcpyHostToDevice));
    SVB_CUDA(cudaMalloc((void **)&c->b, (size_t)Cout * 4));
    w->allocs.push_back(c->b);
    SVB_CUDA(cudaMemcpy(c->b, bt->data.data(), (size_t)Cout * 4, cudaMemcpyHostToDevice));
    SVB_TRY(tc_pack_weights(p.data(), K, Cin, Cout, &c->tc, &w->allocs));
    return SVB_OK;
}

int wn_conv(const svb_wn *w, const WnConv &c, const float *in, float *out, int B, int T, int Tp, cudaStream_t st) {
    ConvArgs a;
    a.in = in, a.w = c.w, a.bias = c.b, a.res = nullptr, a.out = out;
    a.B = B, a.Cin = c.Cin, a.in_Tp = Tp, a.Cout = c.Cout, a.out_Tp = Tp, a.CoutP = c.Cout, a.Tq = T;
    a.KS = c.K, a.dil = c.dil, a.ups_u = 0, a.in_slope = 1.f, a.out_scale = 1.f, a.accumulate = 0;
    if (w->precision != SVB_PREC_FP32 && tc_supported(c.tc, a)) return launch_conv_tc(c.tc, a, w->precision, st);
    return launch_conv_ffma(a, st);
}

}  // namespace

extern "C" int svb_wn_create(int32_t hidden, int32_t kernel_size, int32_t dilation_rate, int32_t n_layers, int32_t gin_channels,
                             int32_t precision, int32_t device, svb_wn_t **out) {
    SVB_CHECK(out && hidden > 0 && hidden % 32 == 0, SVB_ERR_INVALID, "wn_create: hidden_channels %d must be a multiple of 32", hidden);
    SVB_CHECK(kernel_size % 2 == 1 && kernel_size >= 1 && kernel_size <= 11, SVB_ERR_INVALID, "wn_create: kernel_size %d", kernel_size);
    SVB_CHECK(n_layers >= 1 && n_layers <= 16 && dilation_rate >= 1, SVB_ERR_INVALID, "wn_create: n_layers %d dilation_rate %d", n_layers, dilation_rate);
    SVB_CHECK(gin_channels >= 0 && gin_channels % 4 == 0, SVB_ERR_INVALID, "wn_create: gin_channels %d must be a multiple of 4", gin_channels);
    SVB_CHECK(precision >= 0 && precision <= 3, SVB_ERR_INVALID, "wn_create: bad precision %d", precision);
    int dil = 1;
    for (int i = 0; i < n_layers; ++i, dil *= dilation_rate)
        SVB_CHECK((kernel_size - 1) / 2 * dil <= kPad, SVB_ERR_INVALID, "wn_create: layer %d reaches %d rows (halo limit %d)", i,
                  (kernel_size - 1) / 2 * dil, kPad);
    int count = 0;
    SVB_CUDA(cudaGetDeviceCount(&count));
    SVB_CHECK(device >= 0 && device < count, SVB_ERR_INVALID, "wn_create: device %d of %d", device, count);
    svb_wn *w = new (std::nothrow) svb_wn();
    SVB_CHECK(w, SVB_ERR_NOMEM, "wn_create: out of host memory");
    w->device = device, w->precision = precision, w->H = hidden, w->K = kernel_size, w->dil_rate = dilation_rate, w->L = n_layers;
    w->gin = gin_channels;
    *out = w;
    return SVB_OK;
}

extern "C" void svb_wn_destroy(svb_wn_t *w) {
    if (!w) return;
    cudaSetDevice(w->device);
    for (void *p : w->allocs) cudaFree(p);
    if (w->ws) cudaFree(w->ws);
    delete w;
}

extern "C" int svb_wn_set_weight(svb_wn_t *w, const char *name, const float *data, const int64_t *shape, int32_t ndim) {
    SVB_CHECK(w && name && data && shape && ndim >= 1 && ndim <= 3, SVB_ERR_INVALID, "wn_set_weight: bad argument");
    SVB_CHECK(!w->finalized, SVB_ERR_STATE, "wn_set_weight('%s') after finalize", name);
    HostTensor t;
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        SVB_CHECK(shape[i] > 0, SVB_ERR_INVALID, "wn_set_weight('%s'): non-positive dim", name);
        t.shape.push_back(shape[i]);
        n *= (size_t)shape[i];
    }
    t.data.assign(data, data + n);
    w->host_w[name] = std::move(t);
    return SVB_OK;
}

// folded weights (remove_weight_norm applied, fs2_vae.py:96-103): in_layers.{i}.weight [2H, H, K] / .bias,
// res_skip_layers.{i}.weight [2H or H, H, 1] / .bias, cond_layer.weight [2HL, gin, 1] / .bias
extern "C" int svb_wn_finalize(svb_wn_t *w) {
    SVB_CHECK(w && !w->finalized, SVB_ERR_STATE, "wn_finalize: null handle or called twice");
    SVB_CUDA(cudaSetDevice(w->device));
    w->in_layers.resize(w->L), w->rs_layers.resize(w->L);
    int dil = 1;
    for (int i = 0; i < w->L; ++i, dil *= w->dil_rate) {
        SVB_TRY(wn_pack(w, "in_layers." + std::to_string(i), w->H, 2 * w->H, w->K, dil, &w->in_layers[i]));
        SVB_TRY(wn_pack(w, "res_skip_layers." + std::to_string(i), w->H, i + 1 < w->L ? 2 * w->H : w->H, 1, 1, &w->rs_layers[i]));
    }
    if (w->gin > 0) SVB_TRY(wn_pack(w, "cond_layer", w->gin, 2 * w->H * w->L, 1, 1, &w->cond));
    if (w->stride > 0) {        // decoder wrapper: pre_net.0 (ConvTranspose1d weight [latent, H, s]) and out_proj
        const HostTensor *wt, *bt;
        SVB_TRY(wn_get(w, "pre_net.0.weight", {w->latent, w->H, w->stride}, &wt));
        SVB_TRY(wn_get(w, "pre_net.0.bias", {w->H}, &bt));
        int KS = 0;
        const std::vector<float> p = pack_convT_weights(wt->data.data(), w->latent, w->H, w->stride, w->stride, 0, &KS);
        WnConv &c = w->pre;
        c.Cin = w->latent, c.Cout = w->H, c.K = KS, c.dil = 1;
        SVB_CUDA(cudaMalloc((void **)&c.w, p.size() * 4));
        w->allocs.push_back(c.w);
        SVB_CUDA(cudaMemcpy(c.w, p.data(), p.size() * 4, cudaMemcpyHostToDevice));
        SVB_CUDA(cudaMalloc((void **)&c.b, (size_t)w->H * 4));
        w->allocs.push_back(c.b);
        SVB_CUDA(cudaMemcpy(c.b, bt->data.data(), (size_t)w->H * 4, cudaMemcpyHostToDevice));
        SVB_TRY(tc_pack_weights(p.data(), KS, w->latent, w->stride * w->H, &c.tc, &w->allocs));
        SVB_TRY(wn_pack(w, "out_proj", w->H, w->out_ch, 1, 1, &w->post));
    }
    w->finalized = true;
    return SVB_OK;
}

namespace {

// workspace layout for (B, T): x, acts, out (H each), xin, rs (2H each), g (gin), cond (2HL), z (latent, T / stride), y (out_ch)
struct WnBufs {
    float *x, *acts, *out, *xin, *rs, *g, *cond, *z, *y;
};

int wn_workspace(svb_wn *w, int B, int T, WnBufs *bf, cudaStream_t st) {
    const int H = w->H;
    const size_t nH = c4t_floats(B, H, T), n2H = c4t_floats(B, 2 * H, T);
    const size_t nG = w->gin > 0 ? c4t_floats(B, w->gin, T) : 0, nC = w->gin > 0 ? c4t_floats(B, 2 * H * w->L, T) : 0;
    const size_t nZ = w->stride > 0 ? c4t_floats(B, w->latent, T / w->stride) : 0, nY = w->out_ch > 0 ? c4t_floats(B, w->out_ch, T) : 0;
    const size_t need = (3 * nH + 2 * n2H + nG + nC + nZ + nY) * 4;
    if (need > w->ws_cap) {
        if (w->ws) SVB_CUDA(cudaFree(w->ws));
        w->ws = nullptr, w->ws_cap = 0;
        SVB_CUDA(cudaMalloc((void **)&w->ws, need));
        w->ws_cap = need, w->ws_B = 0;
    }
    if (w->ws_B != B || w->ws_T != T) {     // new layout: rebuild the zero padding rows
        SVB_CUDA(cudaMemsetAsync(w->ws, 0, need, st));
        w->ws_B = B, w->ws_T = T;
    }
    bf->x = reinterpret_cast<float *>(w->ws), bf->acts = bf->x + nH, bf->out = bf->acts + nH, bf->xin = bf->out + nH, bf->rs = bf->xin + n2H;
    bf->g = bf->rs + n2H, bf->cond = bf->g + nG, bf->z = bf->cond + nC, bf->y = bf->z + nZ;
    return SVB_OK;
}

// WN.forward on G32T buffers: bf.x holds x, the result lands in bf.out
int wn_core(svb_wn *w, const WnBufs &bf, const float *mask_dev, const float *g_dev, int B, int T, cudaStream_t st) {
    const int H = w->H, Tp = c4t_rows(T), gh = H / 32;
    SVB_CUDA(cudaMemsetAsync(bf.out, 0, c4t_floats(B, H, T) * 4, st));
    const int cond_groups = c4t_groups(2 * H * w->L);
    if (g_dev) {
        SVB_TRY(launch_nct_to_c4t(g_dev, bf.g, B, w->gin, T, Tp, st));
        SVB_TRY(wn_conv(w, w->cond, bf.g, bf.cond, B, T, Tp, st));
    }
    const dim3 grid((unsigned)std::min(148 * 4, (T + 7) / 8), (unsigned)gh, (unsigned)B);
    for (int i = 0; i < w->L; ++i) {
        SVB_TRY(wn_conv(w, w->in_layers[i], bf.x, bf.xin, B, T, Tp, st));
        wn_gate_kernel<<<grid, 256, 0, st>>>(bf.xin, g_dev ? bf.cond : nullptr, i * 2 * gh, cond_groups, gh, T, Tp, bf.acts);
        SVB_TRY(wn_conv(w, w->rs_layers[i], bf.acts, bf.rs, B, T, Tp, st));
        wn_res_skip_kernel<<<grid, 256, 0, st>>>(bf.rs, bf.x, bf.out, mask_dev, gh, T, Tp, i + 1 == w->L ? 1 : 0);
    }
    SVB_CUDA(cudaGetLastError());
    return SVB_OK;
}

}  // namespace

extern "C" int svb_wn_forward(svb_wn_t *w, const float *x_dev, const float *mask_dev, const float *g_dev, int32_t B, int32_t T,
                              float *out_dev, void *stream) {
    SVB_CHECK(w && w->finalized, SVB_ERR_STATE, "wn_forward: handle not finalized");
    SVB_CHECK(x_dev && out_dev && B > 0 && T > 0, SVB_ERR_INVALID, "wn_forward: null buffer or empty input (B %d T %d)", B, T);
    SVB_CHECK(!g_dev || w->gin > 0, SVB_ERR_INVALID, "wn_forward: conditioning given but gin_channels is 0");
    SVB_CUDA(cudaSetDevice(w->device));
    cudaStream_t st = as_stream(stream);
    WnBufs bf;
    SVB_TRY(wn_workspace(w, B, T, &bf, st));
    SVB_TRY(launch_nct_to_c4t(x_dev, bf.x, B, w->H, T, c4t_rows(T), st));
    SVB_TRY(wn_core(w, bf, mask_dev, g_dev, B, T, st));
    return launch_c4t_to_nct(bf.out, out_dev, B, w->H, T, c4t_rows(T), st);
}

// ---- FVAEDecoder / GlobalFVAEDecoder (fs2_vae.py:130-152, vae_models.py:108-128): the mel decoder whose output spec2wav consumes
extern "C" int svb_fvae_decoder_create(int32_t latent_channels, int32_t hidden, int32_t out_channels, int32_t kernel_size, int32_t n_layers,
                                       int32_t gin_channels, int32_t stride, int32_t precision, int32_t device, svb_wn_t **out) {
    SVB_CHECK(latent_channels > 0 && latent_channels % 4 == 0 && out_channels > 0 && out_channels % 4 == 0 && stride >= 1 && stride <= 8,
              SVB_ERR_INVALID, "fvae_decoder_create: latent %d / out %d channels must be multiples of 4, stride %d in 1..8", latent_channels,
              out_channels, stride);
    SVB_TRY(svb_wn_create(hidden, kernel_size, 1, n_layers, gin_channels, precision, device, out));
    (*out)->latent = latent_channels, (*out)->out_ch = out_channels, (*out)->stride = stride;
    return SVB_OK;
}

/* z [B, latent, T / stride], mask [B, T] or NULL, g [B, gin, T] or NULL -> out [B, out_channels, T]  (T a multiple of stride) */
extern "C" int svb_fvae_decoder_forward(svb_wn_t *w, const float *z_dev, const float *mask_dev, const float *g_dev, int32_t B, int32_t T,
                                        float *out_dev, void *stream) {
    SVB_CHECK(w && w->finalized && w->stride > 0, SVB_ERR_STATE, "fvae_decoder_forward: not a finalized decoder handle");
    SVB_CHECK(z_dev && out_dev && B > 0 && T > 0 && T % w->stride == 0, SVB_ERR_INVALID,
              "fvae_decoder_forward: T %d must be a positive multiple of the stride %d", T, w->stride);
    SVB_CHECK(!g_dev || w->gin > 0, SVB_ERR_INVALID, "fvae_decoder_forward: conditioning given but gin_channels is 0");
    SVB_CUDA(cudaSetDevice(w->device));
    cudaStream_t st = as_stream(stream);
    const int Tz = T / w->stride, Tpz = c4t_rows(Tz), Tp = c4t_rows(T), gh = w->H / 32;
    WnBufs bf;
    SVB_TRY(wn_workspace(w, B, T, &bf, st));
    SVB_TRY(launch_nct_to_c4t(z_dev, bf.z, B, w->latent, Tz, Tpz, st));
    {   // pre_net: ConvTranspose1d(k = s, stride = s) = s phase 1x1 convolutions written interleaved (:138-143,148)
        ConvArgs a;
        a.in = bf.z, a.w = w->pre.w, a.bias = w->pre.b, a.res = nullptr, a.out = bf.x;
        a.B = B, a.Cin = w->latent, a.in_Tp = Tpz, a.Cout = w->H, a.out_Tp = Tp, a.CoutP = w->stride * w->H, a.Tq = Tz;
        a.KS = w->pre.K, a.dil = 1, a.ups_u = w->stride, a.in_slope = 1.f, a.out_scale = 1.f, a.accumulate = 0;
        if (w->precision != SVB_PREC_FP32 && tc_supported(w->pre.tc, a)) SVB_TRY(launch_conv_tc(w->pre.tc, a, w->precision, st));
        else SVB_TRY(launch_conv_ffma(a, st));
    }
    const dim3 grid((unsigned)std::min(148 * 4, (T + 7) / 8), (unsigned)gh, (unsigned)B);
    if (mask_dev) wn_mask_kernel<<<grid, 256, 0, st>>>(bf.x, mask_dev, gh, T, Tp);                       // x * x_mask   :149
    SVB_TRY(wn_core(w, bf, mask_dev, g_dev, B, T, st));                                                  // wn(x, mask, g) * mask   :150 (mask is 0/1)
    SVB_TRY(wn_conv(w, w->post, bf.out, bf.y, B, T, Tp, st));                                            // out_proj   :151
    return launch_c4t_to_nct(bf.y, out_dev, B, w->out_ch, T, Tp, st);
}
