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
    lib = _native.lib()
    x = x.contiguous().float()
    B, n = x.shape
    c = _native.StftConfig(int(fft_size), int(hop_size), int(win_length), _native.PAD_CENTER_REFLECT,
                           _native.OUT_MAG, 0, 0, 1, 1e-7)
    frames = int(lib.svb_stft_num_frames(ctypes.byref(c), n))
    out = torch.empty(B, frames, fft_size // 2 + 1, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _native.check(lib.svb_stft_forward(ctypes.byref(c), _native.ptr(x), B, n, None, _native.ptr(out),
                                           _native.current_stream_ptr(x.device)), 'stft_forward')
    return out


MR_STFT = ((1024, 120, 600), (2048, 240, 1200), (512, 50, 240))       # losses/stft_loss.py:113-115


def stft_loss(x, y, fft_size, shift_size, win_length):
    """(spectral convergence ||Y - X||_F / ||Y||_F, log-magnitude L1 mean |ln Y - ln X|) of one resolution
    (STFTLoss.forward, losses/stft_loss.py:89-106); magnitudes from the fused STFT kernel, reductions on the device."""
    from neuralsvb_b200.modules.hifigan.discriminators import pair_stats
    x_mag, y_mag = stft(x, fft_size, shift_size, win_length), stft(y, fft_size, shift_size, win_length)
    s = pair_stats(y_mag, x_mag, want_log=True)
    return float(s[0]) ** 0.5 / float(s[1]) ** 0.5, float(s[2]) / y_mag.numel()


def multi_resolution_stft_loss(x, y, resolutions=MR_STFT):
    """MultiResolutionSTFTLoss.forward (losses/stft_loss.py:130-153): mean over resolutions of (sc, mag)."""
    sc, mag = 0.0, 0.0
    for fs, ss, wl in resolutions:
        s, m = stft_loss(x, y, fs, ss, wl)
        sc, mag = sc + s, mag + m
    return sc / len(resolutions), mag / len(resolutions)
