"""STFT magnitudes of the multi-resolution STFT loss on the fused CUDA front end
(reference: modules/parallel_wavegan/losses/stft_loss.py:12-31)."""
import ctypes

import torch

from neuralsvb_b200 import _native


def stft(x, fft_size, hop_size, win_length, window=None):
    """x [B, T] (CUDA) -> magnitude [B, frames, fft_size//2+1] = sqrt(clamp(re^2+im^2, 1e-7)),
    torch.stft defaults (center=True, reflect), hann window of win_length.  ``window`` is
    accepted for signature compatibility; the kernel generates the hann window itself."""
    if not x.is_cuda:
        raise RuntimeError('stft needs a CUDA tensor: there is no CPU fallback')
    from neuralsvb_b200.modules.hifigan.mel_utils import StftFn
    lib = _native.lib()
    B, n = x.shape
    cfg = (int(fft_size), int(hop_size), int(win_length), _native.PAD_CENTER_REFLECT, _native.OUT_MAG, 0, 0, 1, 1e-7)
    frames = int(lib.svb_stft_num_frames(ctypes.byref(_native.StftConfig(*cfg)), n))
    return StftFn.apply(x, cfg, None, (B, frames, fft_size // 2 + 1))


_mel_cache = {}


def stft_mel(x, fft_size, hop_size, win_length):
    """Mel-projected STFT magnitudes [B, frames, 80] of the ``use_mel_loss`` variant
    (modules/parallel_wavegan/stft_loss.py:40-47: ``mag @ librosa.filters.mel(22050, fft_size, 80).T``, no log),
    fused in the STFT kernel (output kind MEL_MAG)."""
    from neuralsvb_b200.modules.hifigan.mel_utils import StftFn
    from neuralsvb_b200.utils import audio
    if not x.is_cuda:
        raise RuntimeError('stft needs a CUDA tensor: there is no CPU fallback')
    lib = _native.lib()
    key = (int(fft_size), str(x.device))
    if key not in _mel_cache:
        hp = {'audio_sample_rate': 22050, 'fft_size': int(fft_size), 'audio_num_mel_bins': 80, 'fmin': 0, 'fmax': 22050 // 2}
        _mel_cache[key] = torch.from_numpy(audio.build_mel_basis(hp).copy()).float().to(x.device)
    B, n = x.shape
    cfg = (int(fft_size), int(hop_size), int(win_length), _native.PAD_CENTER_REFLECT, _native.OUT_MEL_MAG, 0, 80, 1, 1e-7)
    frames = int(lib.svb_stft_num_frames(ctypes.byref(_native.StftConfig(*cfg)), n))
    return StftFn.apply(x, cfg, _mel_cache[key], (B, frames, 80))


MR_STFT = ((1024, 120, 600), (2048, 240, 1200), (512, 50, 240))       # losses/stft_loss.py:113-115


class _StftLossFn(torch.autograd.Function):
    """(spectral convergence, log-magnitude L1) of predicted magnitudes x against target magnitudes y
    (SpectralConvergengeLoss / LogSTFTMagnitudeLoss, losses/stft_loss.py:34-73): float64 reductions by svb_pair_stats,
    gradient w.r.t. x by svb_loss_grad (the target carries none)."""

    @staticmethod
    def forward(ctx, x_mag, y_mag):
        from neuralsvb_b200.modules.hifigan.discriminators import pair_stats_dev
        s = pair_stats_dev(y_mag, x_mag, want_log=True)           # stays on the device: no host synchronisation
        dnorm, ynorm = s[0].sqrt(), s[1].sqrt()
        ctx.save_for_backward(x_mag, y_mag, (dnorm * ynorm).clamp_min(1e-30))
        return (dnorm / ynorm).float(), (s[2] / y_mag.numel()).float()

    @staticmethod
    def backward(ctx, g_sc, g_mag):
        lib = _native.lib()
        x, y, norms = ctx.saved_tensors
        n = x.numel()
        dx = torch.empty_like(x)
        s_sc = (g_sc.double() / norms).float().contiguous()      # device scalars: the kernels read them in place
        s_mag = g_mag.detach().float().contiguous()
        with torch.cuda.device(x.device):
            st = _native.current_stream_ptr(x.device)
            # d/dx ||y - x|| / ||y|| = (x - y) / (||y - x|| ||y||) ;  d/dx mean |ln y - ln x| = sign(ln x - ln y) / (n x)
            _native.check(lib.svb_loss_grad_dev(_native.ptr(x), _native.ptr(y), 3, ctypes.c_float(1.0), _native.ptr(s_sc),
                                                _native.ptr(dx), n, 0, st), 'loss_grad')
            _native.check(lib.svb_loss_grad_dev(_native.ptr(x), _native.ptr(y), 4, ctypes.c_float(1.0 / n), _native.ptr(s_mag),
                                                _native.ptr(dx), n, 1, st), 'loss_grad')
        return dx, None


def stft_loss(x, y, fft_size, shift_size, win_length, use_mel_loss=False):
    """(spectral convergence ||Y - X||_F / ||Y||_F, log-magnitude L1 mean |ln Y - ln X|) of one resolution
    (STFTLoss.forward, losses/stft_loss.py:89-106); magnitudes from the fused STFT kernel, reductions on the device.
    Python floats without grad; differentiable scalars w.r.t. the predicted signal ``x`` when it requires grad."""
    from neuralsvb_b200.modules.hifigan.discriminators import pair_stats
    f = stft_mel if use_mel_loss else stft
    x_mag, y_mag = f(x, fft_size, shift_size, win_length), f(y.detach(), fft_size, shift_size, win_length)
    if torch.is_grad_enabled() and x_mag.requires_grad:
        return _StftLossFn.apply(x_mag, y_mag)
    s = pair_stats(y_mag, x_mag, want_log=True)
    return float(s[0]) ** 0.5 / float(s[1]) ** 0.5, float(s[2]) / y_mag.numel()


def multi_resolution_stft_loss(x, y, resolutions=MR_STFT, use_mel_loss=False):
    """MultiResolutionSTFTLoss.forward (losses/stft_loss.py:130-153; modules/parallel_wavegan/stft_loss.py:55-100 with
    ``use_mel_loss``): mean over resolutions of (sc, mag)."""
    sc, mag = 0.0, 0.0
    for fs, ss, wl in resolutions:
        s, m = stft_loss(x, y, fs, ss, wl, use_mel_loss)
        sc, mag = sc + s, mag + m
    return sc / len(resolutions), mag / len(resolutions)
