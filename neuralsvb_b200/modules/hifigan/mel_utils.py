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


class StftFn(torch.autograd.Function):
    """svb_stft_forward as an autograd node; backward = svb_stft_backward (the spectrum is recomputed per frame).
    ``cfg`` = (n_fft, hop, win, pad_mode, out_kind, clamp_input, n_mels, frames_major, eps)."""

    @staticmethod
    def forward(ctx, y, cfg, basis, out_shape):
        lib = _native.lib()
        y = y.contiguous().float()
        c = _native.StftConfig(*cfg)
        out = torch.empty(out_shape, device=y.device, dtype=torch.float32)
        with torch.cuda.device(y.device):
            _native.check(lib.svb_stft_forward(ctypes.byref(c), _native.ptr(y), y.shape[0], y.shape[1], _native.ptr(basis),
                                               _native.ptr(out), _native.current_stream_ptr(y.device)), 'stft_forward')
        ctx.cfg, ctx.basis = cfg, basis
        ctx.save_for_backward(y)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _native.lib()
        (y,) = ctx.saved_tensors
        c = _native.StftConfig(*ctx.cfg)
        dout = dout.contiguous().float()
        dy = torch.zeros_like(y)
        with torch.cuda.device(y.device):
            _native.check(lib.svb_stft_backward(ctypes.byref(c), _native.ptr(y), y.shape[0], y.shape[1], _native.ptr(ctx.basis),
                                                _native.ptr(dout), _native.ptr(dy), _native.current_stream_ptr(y.device)),
                          'stft_backward')
        return dy, None, None, None


def mel_spectrogram(y, hparams, center=False, complex=False):
    """y [B, T_wav] on a CUDA device -> ln-mel [B, n_mels, T_wav / hop]:
    clamp(-1, 1), reflect-pad (n_fft-hop)/2, STFT (hann(win), center=False),
    sqrt(re^2+im^2+1e-9), mel matmul, ln(clamp(., 1e-5)) -- one fused kernel."""
    if complex or center:
        raise NotImplementedError('only the non-complex, center=False branch is on the vocoder path')
    if not y.is_cuda:
        raise RuntimeError('mel_spectrogram needs a CUDA tensor: there is no CPU fallback')
    lib = _native.lib()
    B, n = y.shape
    cfg = (int(hparams['fft_size']), int(hparams['hop_size']), int(hparams['win_size']), _native.PAD_HALF_REFLECT,
           _native.OUT_LN_MEL, 1, int(hparams['audio_num_mel_bins']), 0, 1e-5)
    frames = int(lib.svb_stft_num_frames(ctypes.byref(_native.StftConfig(*cfg)), n))
    return StftFn.apply(y, cfg, _basis(hparams, y.device), (B, cfg[6], frames))


def wav2spec_mel(y, hparams, frames=None):
    """y [B, T_wav] on a CUDA device -> log10-mel [B, n_mels, frames] exactly as the binarizer stores it and as the
    plugin's ``spec2wav`` receives it: ``process_utterance`` (data_gen/tts/data_gen_utils.py:123-134) =
    librosa.stft(center=True, pad_mode='constant') -> |.| -> mel basis -> log10(max(wav2spec_eps, .)).
    This is the generator's CONDITIONING input in training (``mel_spectrogram`` above is only the mel-L1 loss
    transform: natural log, clamp 1e-5, half-reflect padding).  ``frames`` crops to the clip's own frame count
    (len // hop, the binarized ``mel[:T]`` aligned with ``wav[:T * hop]``); no gradient flows through it."""
    if not y.is_cuda:
        raise RuntimeError('wav2spec_mel needs a CUDA tensor: there is no CPU fallback')
    lib = _native.lib()
    B, n = y.shape
    cfg = (int(hparams['fft_size']), int(hparams['hop_size']), int(hparams['win_size']), _native.PAD_CENTER_ZERO,
           _native.OUT_LOG10_MEL, 0, int(hparams['audio_num_mel_bins']), 0, float(hparams.get('wav2spec_eps', 1e-10)))
    total = int(lib.svb_stft_num_frames(ctypes.byref(_native.StftConfig(*cfg)), n))
    with torch.no_grad():
        mel = StftFn.apply(y.detach(), cfg, _basis(hparams, y.device), (B, cfg[6], total))
    frames = n // cfg[1] if frames is None else frames
    return mel[:, :, :frames].contiguous()
