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
