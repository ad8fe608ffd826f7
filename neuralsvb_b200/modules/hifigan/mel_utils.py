"""Training-time mel spectrogram of the vocoder losses on the fused CUDA front end
(reference: modules/hifigan/mel_utils.py:45-80)."""
import ctypes

import torch

from neuralsvb_b200 import _native
from neuralsvb_b200.utils import audio

_basis_cache = {}


def _basis(hp, device):
    key = (hp['audio_sample_rate'], hp['fft_size'], hp['audio_num_mel_bins'], hp['fmin'], hp['fmax'], str(device))
    if key not in _basis_cache:
        _basis_cache[key] = torch.from_numpy(audio.build_mel_basis(hp).copy()).to(device)
    return _basis_cache[key]


def mel_spectrogram(y, hparams, center=False, complex=False):
    """y [B, T_wav] on a CUDA device -> ln-mel [B, n_mels, T_wav / hop]:
    clamp(-1, 1), reflect-pad (n_fft-hop)/2, STFT (hann(win), center=False),
    sqrt(re^2+im^2+1e-9), mel matmul, ln(clamp(., 1e-5)) -- one fused kernel."""
    if complex or center:
        raise NotImplementedError('only the non-complex, center=False branch is on the vocoder path')
    if not y.is_cuda:
        raise RuntimeError('mel_spectrogram needs a CUDA tensor: there is no CPU fallback')
    lib = _native.lib()
    y = y.contiguous().float()
    B, n = y.shape
    c = _native.StftConfig(int(hparams['fft_size']), int(hparams['hop_size']), int(hparams['win_size']),
                           _native.PAD_HALF_REFLECT, _native.OUT_LN_MEL, 1, int(hparams['audio_num_mel_bins']), 0, 1e-5)
    frames = int(lib.svb_stft_num_frames(ctypes.byref(c), n))
    out = torch.empty(B, c.n_mels, frames, device=y.device, dtype=torch.float32)
    with torch.cuda.device(y.device):
        _native.check(lib.svb_stft_forward(ctypes.byref(c), _native.ptr(y), B, n, _native.ptr(_basis(hparams, y.device)),
                                           _native.ptr(out), _native.current_stream_ptr(y.device)), 'stft_forward')
    return out
