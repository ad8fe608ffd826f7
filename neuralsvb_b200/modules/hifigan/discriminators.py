"""MultiPeriodDiscriminator / MultiScaleDiscriminator and the GAN losses, forward pass on the CUDA
operators of csrc/disc_ops.cu (reference: modules/hifigan/hifigan.py:181-365).

Same constructors, ``forward(y, y_hat, mel=None) -> (y_d_rs, y_d_gs, fmap_rs, fmap_gs)`` and
state_dict names (weight-norm ``weight_g/weight_v``; spectral-norm ``weight_orig/weight_u/weight_v``
for MSD[0]) as the reference.  Forward only this round (no autograd): used for evaluation and as
the parity-checked building block of the vocoder training step.
``use_cond=True`` (mel-conditioned discriminators, off in the shipped config) is not implemented.
"""
import ctypes

import torch
from torch import nn

from neuralsvb_b200 import _native
from neuralsvb_b200.utils.synthetic import MPD_PERIODS, MSD_LAYERS

LRELU_SLOPE = 0.1


def _cuda(t):
    if not t.is_cuda:
        raise RuntimeError('discriminators need CUDA tensors: there is no CPU fallback')
    return t.contiguous().float()


def conv_nct(x, w, b, K, stride=1, dil=1, pad=0, groups=1, slope=1.0, W=1):
    """x [B, Cin, T(, W)] -> [B, Cout, Tout(, W)] through svb_conv_nct_forward."""
    lib = _native.lib()
    x = _cuda(x)
    B, Cin, Tin = x.shape[0], x.shape[1], x.shape[2]
    Cout = w.shape[0]
    Tout = (Tin + 2 * pad - dil * (K - 1) - 1) // stride + 1
    y = torch.empty((B, Cout, Tout) + ((W,) if x.dim() == 4 else ()), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _native.check(lib.svb_conv_nct_forward(_native.ptr(x), _native.ptr(w), _native.ptr(b), _native.ptr(y), B, Cin, Cout, Tin,
                                               W, K, stride, dil, pad, groups, ctypes.c_float(slope),
                                               _native.current_stream_ptr(x.device)), 'conv_nct_forward')
    return y


class _NormConv(nn.Module):
    """Parameter container for a weight- or spectral-normalised conv; effective weight computed on the device."""

    def __init__(self, shape, spectral=False):
        super().__init__()
        cout = shape[0]
        inner = 1
        for s in shape[1:]:
            inner *= s
        self.spectral = spectral
        self.bias = nn.Parameter(torch.zeros(cout))
        v = torch.randn(*shape) * 0.02
        if spectral:
            self.weight_orig = nn.Parameter(v)
            self.register_buffer('weight_u', nn.functional.normalize(torch.randn(cout), dim=0))
            self.register_buffer('weight_v', nn.functional.normalize(torch.randn(inner), dim=0))
        else:
            self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).view(-1, *([1] * (len(shape) - 1))))
            self.weight_v = nn.Parameter(v)
        self._cache = None

    def effective(self, device):
        """[Cout, Cin/groups, K] effective weight + bias on `device` (cached until parameters change)."""
        key = (str(device), self.bias._version, (self.weight_orig if self.spectral else self.weight_v)._version)
        if self._cache is not None and self._cache[0] == key:
            return self._cache[1], self._cache[2]
        lib = _native.lib()
        dev_idx = device.index if device.index is not None else torch.cuda.current_device()
        if self.spectral:
            w = self.weight_orig.detach().float().cpu().contiguous()
            u, v = self.weight_u.float().cpu().contiguous(), self.weight_v.float().cpu().contiguous()
            sigma = ctypes.c_float()
            _native.check(lib.svb_spectral_sigma_host(_native.ptr(w), _native.ptr(u), _native.ptr(v), w.shape[0], w[0].numel(),
                                                      dev_idx, ctypes.byref(sigma)), 'spectral_sigma')
            eff = w / sigma.value
        else:
            v = self.weight_v.detach().float().cpu().contiguous()
            g = self.weight_g.detach().float().cpu().contiguous().view(-1)
            eff = torch.empty_like(v)
            _native.check(lib.svb_fold_weight_norm_host(_native.ptr(v), _native.ptr(g), v.shape[0], v[0].numel(), _native.ptr(eff),
                                                        dev_idx), 'fold_weight_norm')
        eff = eff.reshape(eff.shape[0], eff.shape[1], -1).contiguous().to(device)
        b = self.bias.detach().float().to(device)
        self._cache = (key, eff, b)
        return eff, b


class DiscriminatorP(nn.Module):
    def __init__(self, period, kernel_size=5, stride=3, use_spectral_norm=False, use_cond=False, c_in=1):
        super().__init__()
        if use_cond:
            raise NotImplementedError('use_cond discriminators are not on the shipped path')
        self.period, self.kernel_size, self.stride = period, kernel_size, stride
        ch = [(c_in, 32), (32, 128), (128, 512), (512, 1024), (1024, 1024)]
        self.convs = nn.ModuleList([_NormConv((co, ci, kernel_size, 1), use_spectral_norm) for ci, co in ch])
        self.conv_post = _NormConv((1, 1024, 3, 1), use_spectral_norm)

    def forward(self, x, mel=None):
        lib = _native.lib()
        x = _cuda(x)
        b, c, t = x.shape
        p = self.period
        if t % p != 0:                                   # reflect pad to a multiple of the period (:209-212)
            tp = t + (p - t % p)
            xp = torch.empty(b, c, tp, device=x.device)
            with torch.cuda.device(x.device):
                _native.check(lib.svb_pad_reflect_right(_native.ptr(x), _native.ptr(xp), b * c, t, tp,
                                                        _native.current_stream_ptr(x.device)), 'pad_reflect_right')
            x, t = xp, tp
        x = x.view(b, c, t // p, p)
        fmap = []
        for i, l in enumerate(self.convs):
            w, bias = l.effective(x.device)
            x = conv_nct(x, w, bias, self.kernel_size, stride=(self.stride if i < 4 else 1), pad=2, slope=LRELU_SLOPE, W=p)
            fmap.append(x)
        w, bias = self.conv_post.effective(x.device)
        x = conv_nct(x, w, bias, 3, pad=1, W=p)
        fmap.append(x)
        return torch.flatten(x, 1, -1), fmap


class MultiPeriodDiscriminator(nn.Module):
    def __init__(self, use_cond=False, c_in=1):
        super().__init__()
        self.discriminators = nn.ModuleList([DiscriminatorP(p, use_cond=use_cond, c_in=c_in) for p in MPD_PERIODS])

    def forward(self, y, y_hat, mel=None):
        y_d_rs, y_d_gs, fmap_rs, fmap_gs = [], [], [], []
        for d in self.discriminators:
            r, fr = d(y, mel)
            g, fg = d(y_hat, mel)
            y_d_rs.append(r), fmap_rs.append(fr), y_d_gs.append(g), fmap_gs.append(fg)
        return y_d_rs, y_d_gs, fmap_rs, fmap_gs


class DiscriminatorS(nn.Module):
    def __init__(self, use_spectral_norm=False, use_cond=False, upsample_rates=None, c_in=1):
        super().__init__()
        if use_cond:
            raise NotImplementedError('use_cond discriminators are not on the shipped path')
        self.convs = nn.ModuleList([_NormConv((co, (c_in if i == 0 else ci) // g, k), use_spectral_norm)
                                    for i, (ci, co, k, _, g, _) in enumerate(MSD_LAYERS)])
        self.conv_post = _NormConv((1, 1024, 3), use_spectral_norm)

    def forward(self, x, mel=None):
        fmap = []
        for l, (_, _, k, s, g, p) in zip(self.convs, MSD_LAYERS):
            w, bias = l.effective(x.device)
            x = conv_nct(x, w, bias, k, stride=s, pad=p, groups=g, slope=LRELU_SLOPE)
            fmap.append(x)
        w, bias = self.conv_post.effective(x.device)
        x = conv_nct(x, w, bias, 3, pad=1)
        fmap.append(x)
        return torch.flatten(x, 1, -1), fmap


def avg_pool_4_2_1(x):
    lib = _native.lib()
    x = _cuda(x)
    b, c, t = x.shape
    y = torch.empty(b, c, (t + 2 - 4) // 2 + 1, device=x.device)
    with torch.cuda.device(x.device):
        _native.check(lib.svb_avgpool1d_4_2_1(_native.ptr(x), _native.ptr(y), b * c, t, _native.current_stream_ptr(x.device)),
                      'avgpool')
    return y


class MultiScaleDiscriminator(nn.Module):
    def __init__(self, use_cond=False, c_in=1):
        super().__init__()
        self.discriminators = nn.ModuleList([DiscriminatorS(use_spectral_norm=True, c_in=c_in), DiscriminatorS(c_in=c_in),
                                             DiscriminatorS(c_in=c_in)])

    def forward(self, y, y_hat, mel=None):
        y_d_rs, y_d_gs, fmap_rs, fmap_gs = [], [], [], []
        for i, d in enumerate(self.discriminators):
            if i != 0:
                y, y_hat = avg_pool_4_2_1(y), avg_pool_4_2_1(y_hat)
            r, fr = d(y, mel)
            g, fg = d(y_hat, mel)
            y_d_rs.append(r), fmap_rs.append(fr), y_d_gs.append(g), fmap_gs.append(fg)
        return y_d_rs, y_d_gs, fmap_rs, fmap_gs


# ------------------------------------------------------------------ losses (device reductions)
def pair_stats(a, b=None, want_log=False):
    """float64 [6] on the host: sum (a-b)^2, sum a^2, sum |ln a - ln b|, sum |a-b|, sum (1-a)^2, sum b^2."""
    lib = _native.lib()
    a = _cuda(a)
    b = None if b is None else _cuda(b)
    out = torch.empty(6, device=a.device, dtype=torch.float64)
    with torch.cuda.device(a.device):
        _native.check(lib.svb_pair_stats(_native.ptr(a), _native.ptr(b), a.numel(), int(want_log), ctypes.c_void_p(out.data_ptr()),
                                         _native.current_stream_ptr(a.device)), 'pair_stats')
    return out.cpu()


def feature_loss(fmap_r, fmap_g):
    """2 * sum over discriminators and layers of mean |r - g|  (hifigan.py:328-334)."""
    loss = 0.0
    for dr, dg in zip(fmap_r, fmap_g):
        for rl, gl in zip(dr, dg):
            loss += float(pair_stats(rl, gl)[3]) / rl.numel()
    return loss * 2


def discriminator_loss(disc_real_outputs, disc_generated_outputs):
    """(mean over discriminators of mean (1 - dr)^2, of mean dg^2)  (hifigan.py:337-347)."""
    r, g = 0.0, 0.0
    for dr, dg in zip(disc_real_outputs, disc_generated_outputs):
        s = pair_stats(dr, dg)
        r += float(s[4]) / dr.numel()
        g += float(s[5]) / dg.numel()
    n = len(disc_real_outputs)
    return r / n, g / n


def generator_loss(disc_outputs):
    """mean over discriminators of mean (1 - dg)^2  (hifigan.py:359-365)."""
    return sum(float(pair_stats(dg)[4]) / dg.numel() for dg in disc_outputs) / len(disc_outputs)
