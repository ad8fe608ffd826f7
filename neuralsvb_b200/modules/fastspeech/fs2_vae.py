"""WN gated dilated-conv stack of the GlobalFVAE -- first piece of the SVB acoustic step.

Mirror of the reference class ``modules/fastspeech/fs2_vae.py:19-103`` (same constructor, ``state_dict`` names with
the weight-norm ``weight_g`` / ``weight_v`` pairs, ``remove_weight_norm``, ``forward(x, x_mask, g)``); the module only
holds parameters, the arithmetic is ``libsvb_vocoder.so`` (``svb_wn_*``, csrc/wn.cu: three tcgen05 convolutions per
layer plus the gate / residual-skip kernels).  Inference only (eval mode, no autograd): the acoustic model's training
step is not part of this package yet.
"""
import ctypes

import torch
from torch import nn

from neuralsvb_b200 import _native


class _WNConv(nn.Module):
    """Parameter holder with the names ``torch.nn.utils.weight_norm(Conv1d)`` gives: bias, weight_g, weight_v."""

    def __init__(self, cin, cout, k):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(cout))
        self.weight_g = nn.Parameter(torch.ones(cout, 1, 1))
        self.weight_v = nn.Parameter(torch.randn(cout, cin, k) * 0.01)
        self.weight = None                      # set by remove_weight_norm

    def folded(self):
        if self.weight is not None:
            return self.weight
        return torch._weight_norm(self.weight_v, self.weight_g, 0)

    def remove_weight_norm(self):
        if self.weight is None:
            w = self.folded().detach()
            del self.weight_g, self.weight_v
            self.weight = nn.Parameter(w)


class WN(nn.Module):
    def __init__(self, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0, p_dropout=0,
                 share_cond_layers=False, is_BTC=False, precision='bf16x3'):
        super().__init__()
        assert kernel_size % 2 == 1 and hidden_channels % 2 == 0          # fs2_vae.py:25-26
        if share_cond_layers:
            raise NotImplementedError('WN(share_cond_layers=True): the shared cond layer lives outside this module '
                                      '(glow-style use); the FVAE path of SURVEY 8(f) does not use it')
        self.is_BTC, self.hidden_channels, self.kernel_size = is_BTC, hidden_channels, kernel_size
        self.dilation_rate, self.n_layers, self.gin_channels, self.p_dropout = dilation_rate, n_layers, gin_channels, p_dropout
        self.share_cond_layers, self.precision = share_cond_layers, precision
        if gin_channels:
            self.cond_layer = _WNConv(gin_channels, 2 * hidden_channels * n_layers, 1)
        self.in_layers = nn.ModuleList(_WNConv(hidden_channels, 2 * hidden_channels, kernel_size) for _ in range(n_layers))
        self.res_skip_layers = nn.ModuleList(
            _WNConv(hidden_channels, 2 * hidden_channels if i < n_layers - 1 else hidden_channels, 1) for i in range(n_layers))
        self._handle, self._versions = None, None

    def remove_weight_norm(self):
        for m in self.modules():
            if isinstance(m, _WNConv):
                m.remove_weight_norm()

    def _convs(self):
        out = [(f'in_layers.{i}', m) for i, m in enumerate(self.in_layers)]
        out += [(f'res_skip_layers.{i}', m) for i, m in enumerate(self.res_skip_layers)]
        if self.gin_channels:
            out.append(('cond_layer', self.cond_layer))
        return out

    def _native_handle(self, device):
        versions = tuple(p._version for p in self.parameters()) + (str(device),)
        if self._handle is not None and versions == self._versions:
            return self._handle
        self._free()
        lib = _native.lib()
        h = ctypes.c_void_p()
        _native.check(lib.svb_wn_create(self.hidden_channels, self.kernel_size, self.dilation_rate, self.n_layers, self.gin_channels,
                                        _native.PREC[self.precision], device.index or 0, ctypes.byref(h)), 'wn_create')
        for name, m in self._convs():
            for suffix, t in (('weight', m.folded()), ('bias', m.bias)):
                t = t.detach().float().cpu().contiguous()
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                _native.check(lib.svb_wn_set_weight(h, f'{name}.{suffix}'.encode(), ctypes.c_void_p(t.data_ptr()), shape, t.dim()),
                              'wn_set_weight')
        _native.check(lib.svb_wn_finalize(h), 'wn_finalize')
        self._handle, self._versions = h, versions
        return h

    def _free(self):
        if self._handle is not None:
            _native.lib().svb_wn_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def forward(self, x, x_mask=None, g=None, **kwargs):
        """fs2_vae.py:62-94.  x [B, H, T] ([B, T, H] with is_BTC), x_mask [B, 1, T], g [B, gin, T]."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise RuntimeError('neuralsvb_b200 WN is inference only: call .eval() and run under torch.no_grad()')
        if not x.is_cuda:
            raise RuntimeError('neuralsvb_b200 has no CPU path: move the module and its inputs to a CUDA device')
        if self.is_BTC:
            x = x.transpose(1, 2)
            x_mask = x_mask.transpose(1, 2) if x_mask is not None else None
        x = x.float().contiguous()
        B, H, T = x.shape
        assert H == self.hidden_channels, (H, self.hidden_channels)
        mask = None if x_mask is None else x_mask.float().expand(B, 1, T).reshape(B, T).contiguous()
        gc = None if g is None else g.float().expand(B, self.gin_channels, T).contiguous()
        out = torch.empty_like(x)
        h = self._native_handle(x.device)
        st = torch.cuda.current_stream(x.device).cuda_stream
        _native.check(_native.lib().svb_wn_forward(h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(mask.data_ptr()) if mask is not None else None,
                                                   ctypes.c_void_p(gc.data_ptr()) if gc is not None else None, B, T,
                                                   ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(st)), 'wn_forward')
        return out.transpose(1, 2) if self.is_BTC else out


class FVAEDecoder(nn.Module):
    """Mirror of ``modules/fastspeech/fs2_vae.py:130-152``: pre_net ConvTranspose1d(k = s, stride = s) -> * x_mask -> WN -> * x_mask
    -> out_proj.  One native call (``svb_fvae_decoder_forward``); ``pre_net`` / ``out_proj`` are parameter holders with the
    reference's ``state_dict`` names.  Inference only."""

    def __init__(self, latent_channels, hidden_channels, out_channels, kernel_size, n_layers, gin_channels=0, p_dropout=0,
                 strides=[4], precision='bf16x3'):
        super().__init__()
        if len(strides) != 1:
            raise NotImplementedError('FVAEDecoder: one up-sampling stride (the reference configs use strides=[4])')
        self.strides, self.hidden_size = list(strides), hidden_channels
        self.latent_channels, self.out_channels, self.precision = latent_channels, out_channels, precision
        self.pre_net = nn.Sequential(nn.ConvTranspose1d(latent_channels, hidden_channels, kernel_size=strides[0], stride=strides[0]))
        self.wn = WN(hidden_channels, kernel_size, 1, n_layers, gin_channels, p_dropout, precision=precision)
        self.out_proj = nn.Conv1d(hidden_channels, out_channels, 1)
        self._handle, self._versions = None, None

    def _native_handle(self, device):
        versions = tuple(p._version for p in self.parameters()) + (str(device),)
        if self._handle is not None and versions == self._versions:
            return self._handle
        self._free()
        lib = _native.lib()
        h = ctypes.c_void_p()
        wn = self.wn
        _native.check(lib.svb_fvae_decoder_create(self.latent_channels, wn.hidden_channels, self.out_channels, wn.kernel_size, wn.n_layers,
                                                  wn.gin_channels, self.strides[0], _native.PREC[self.precision], device.index or 0,
                                                  ctypes.byref(h)), 'fvae_decoder_create')
        tensors = [(f'{name}.{sfx}', t) for name, m in wn._convs() for sfx, t in (('weight', m.folded()), ('bias', m.bias))]
        tensors += [('pre_net.0.weight', self.pre_net[0].weight), ('pre_net.0.bias', self.pre_net[0].bias),
                    ('out_proj.weight', self.out_proj.weight), ('out_proj.bias', self.out_proj.bias)]
        for name, t in tensors:
            t = t.detach().float().cpu().contiguous()
            shape = (ctypes.c_int64 * t.dim())(*t.shape)
            _native.check(lib.svb_wn_set_weight(h, name.encode(), ctypes.c_void_p(t.data_ptr()), shape, t.dim()), 'wn_set_weight')
        _native.check(lib.svb_wn_finalize(h), 'wn_finalize')
        self._handle, self._versions = h, versions
        return h

    def _free(self):
        if self._handle is not None:
            _native.lib().svb_wn_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def forward(self, x, x_mask, g):
        """x [B, latent, T / s]; x_mask [B, 1, T] (or the scalar 1 the reference passes at inference); g [B, gin, T] or None."""
        if torch.is_grad_enabled() and self.training:
            raise RuntimeError('neuralsvb_b200 FVAEDecoder is inference only: call .eval() and run under torch.no_grad()')
        if not x.is_cuda:
            raise RuntimeError('neuralsvb_b200 has no CPU path: move the module and its inputs to a CUDA device')
        x = x.float().contiguous()
        B, _, Tz = x.shape
        T = Tz * self.strides[0]
        mask = None
        if torch.is_tensor(x_mask):
            mask = x_mask.float().expand(B, 1, T).reshape(B, T).contiguous()
        gc = None if g is None else g.float().expand(B, self.wn.gin_channels, T).contiguous()
        out = torch.empty(B, self.out_channels, T, device=x.device, dtype=torch.float32)
        st = torch.cuda.current_stream(x.device).cuda_stream
        _native.check(_native.lib().svb_fvae_decoder_forward(
            self._native_handle(x.device), ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(mask.data_ptr()) if mask is not None else None,
            ctypes.c_void_p(gc.data_ptr()) if gc is not None else None, B, T, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(st)),
            'fvae_decoder_forward')
        return out


class GlobalFVAEDecoder(FVAEDecoder):
    """``modules/voice_conversion/vae_models.py:108-128``: one latent vector per utterance, repeated over T / 4 frames."""

    def forward(self, x, x_mask, g):
        x = x.repeat(1, 1, g.shape[-1] // self.strides[0])          # [B, latent, 1] -> [B, latent, T // 4]   :122
        return super().forward(x, x_mask, g)


class GlobalFVAEEncoder(nn.Module):
    """``modules/voice_conversion/vae_models.py:81-106`` (FVAEEncoder, fs2_vae.py:106-127, plus the global pooling head):
    pre_net strided conv -> * mask -> WN (8 layers) -> * mask -> out_proj -> poolings (3 strided convs, ReLU, BatchNorm in eval
    mode) -> mean over time -> (m, logs) -> z = m + eps * exp(logs).  Every convolution is a native kernel (``svb_conv_nct_forward``
    for the strided / 1x1 layers, ``svb_wn_forward`` for the stack); the BatchNorm affine and the mean act on [B, 256, T / 32]
    tensors.  Inference only.  ``forward(..., eps=...)`` injects the posterior noise (parity tests); default: torch.randn_like."""

    def __init__(self, in_channels, hidden_channels, latent_channels, kernel_size, n_layers, gin_channels=0, p_dropout=0, strides=[4],
                 precision='bf16x3'):
        super().__init__()
        if len(strides) != 1:
            raise NotImplementedError('GlobalFVAEEncoder: one down-sampling stride (the reference configs use strides=[4])')
        s0 = strides[0]
        self.strides, self.hidden_size, self.latent_channels = list(strides), hidden_channels, latent_channels
        self.pre_net = nn.Sequential(nn.Conv1d(in_channels, hidden_channels, kernel_size=s0 * 2, stride=s0, padding=s0 // 2))
        self.wn = WN(hidden_channels, kernel_size, 1, n_layers, gin_channels, p_dropout, precision=precision)
        self.out_proj = nn.Conv1d(hidden_channels, latent_channels * 2, 1)
        c2 = latent_channels * 2
        self.poolings = nn.Sequential(nn.Conv1d(c2, c2, kernel_size=3, stride=2), nn.ReLU(), nn.BatchNorm1d(c2),
                                      nn.Conv1d(c2, c2, kernel_size=3, stride=2), nn.ReLU(), nn.BatchNorm1d(c2),
                                      nn.Conv1d(c2, c2, kernel_size=3, stride=2))

    @staticmethod
    def _bn_eval(bn, x):
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        return x * scale[None, :, None] + (bn.bias - bn.running_mean * scale)[None, :, None]

    def forward(self, x, x_mask, g, eps=None):
        from neuralsvb_b200.modules.hifigan.discriminators import conv_nct
        if torch.is_grad_enabled() and self.training:
            raise RuntimeError('neuralsvb_b200 GlobalFVAEEncoder is inference only: call .eval() and run under torch.no_grad()')
        s0 = self.strides[0]
        c = self.pre_net[0]
        x = conv_nct(x.float(), c.weight, c.bias, 2 * s0, stride=s0, pad=s0 // 2)                       # vae_models.py:97
        x_mask = x_mask[:, :, ::s0][:, :, :x.shape[-1]]                                                 # :98
        x = x * x_mask                                                                                  # :99
        x = self.wn(x, x_mask, g) * x_mask                                                              # :100
        x = conv_nct(x, self.out_proj.weight, self.out_proj.bias, 1)                                    # :101
        p = self.poolings
        x = self._bn_eval(p[2], conv_nct(x, p[0].weight, p[0].bias, 3, stride=2, slope=0.0))            # conv + ReLU (slope 0), BatchNorm (eval)
        x = self._bn_eval(p[5], conv_nct(x, p[3].weight, p[3].bias, 3, stride=2, slope=0.0))
        x = conv_nct(x, p[6].weight, p[6].bias, 3, stride=2)
        x = torch.mean(x, dim=-1, keepdim=True)                                                         # :102
        m, logs = torch.split(x, self.latent_channels, dim=1)                                           # :103
        z = m + (torch.randn_like(m) if eps is None else eps) * torch.exp(logs)                         # :104
        return z, m, logs, x_mask


class GlobalFVAE(nn.Module):
    """``modules/voice_conversion/vae_models.py:130-146`` (GlobalFVAE over TMPFVAE :11-52, FVAE fs2_vae.py:155-178) without the
    optional prior flow (``use_prior_glow: false`` in vae_global_mle_eng): ``g_pre_net`` strided conv on the condition, the global
    posterior encoder and the mel decoder -- every convolution native.  ``forward(x, x_mask, g, infer)`` returns what the reference
    returns; the KL term is computed in torch on the [B, latent, 1] statistics.  Inference / evaluation only."""

    def __init__(self, in_out_channels, hidden_channels, latent_size, kernel_size, enc_n_layers, dec_n_layers, gin_channels, strides,
                 use_prior_glow=False, glow_hidden=None, glow_kernel_size=None, glow_n_blocks=None, precision='bf16x3'):
        super().__init__()
        if use_prior_glow:
            raise NotImplementedError('GlobalFVAE(use_prior_glow=True): the prior flow is not part of the SVB configuration')
        self.strides, self.hidden_size, self.latent_size, self.use_prior_glow = list(strides), hidden_channels, latent_size, False
        self.g_pre_net = nn.Sequential(*[nn.Conv1d(gin_channels, gin_channels, kernel_size=s * 2, stride=s, padding=s // 2) for s in strides])
        self.encoder = GlobalFVAEEncoder(in_out_channels, hidden_channels, latent_size, kernel_size, enc_n_layers, gin_channels,
                                         strides=strides, precision=precision)
        self.decoder = GlobalFVAEDecoder(latent_size, hidden_channels, in_out_channels, kernel_size, dec_n_layers, gin_channels,
                                         strides=strides, precision=precision)

    def forward(self, x=None, x_mask=None, g=None, infer=False, eps=None):
        from neuralsvb_b200.modules.hifigan.discriminators import conv_nct
        g_sqz = g
        for c in self.g_pre_net:                                                                        # vae_models.py:20
            g_sqz = conv_nct(g_sqz.float(), c.weight, c.bias, c.kernel_size[0], stride=c.stride[0], pad=c.padding[0])
        if infer:                                                                                       # :45-52
            # the reference samples [B, latent, T / 4] here and its GlobalFVAEDecoder then repeats that T / 4 times (a shape
            # error: the SVB task never takes this branch, it calls vae_model.decoder directly, svb_vae.py:303); a GLOBAL
            # latent has one time step
            z_p = torch.randn(g_sqz.shape[0], self.latent_size, 1, device=g.device) if eps is None else eps
            return self.decoder(z_p, 1, g), z_p
        z_q, m_q, logs_q, x_mask_sqz = self.encoder(x, x_mask, g_sqz, eps=eps)                          # :22
        x_recon = self.decoder(z_q, x_mask, g)                                                          # :23
        # validate_args=False: the argument check of torch.distributions is a device -> host synchronisation per call
        q_dist = torch.distributions.Normal(m_q, logs_q.exp(), validate_args=False)
        loss_kl = torch.distributions.kl_divergence(q_dist, torch.distributions.Normal(0.0, 1.0, validate_args=False))   # :38-39
        loss_kl = (loss_kl * x_mask_sqz).sum() / x_mask_sqz.sum() / z_q.shape[1]
        return x_recon, loss_kl, None, m_q, logs_q, x_mask_sqz, z_q
