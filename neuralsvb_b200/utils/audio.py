"""Host helpers of the wav -> mel front end (reference: utils/audio.py:57-117 and the call
sites of librosa.filters.mel in data_gen/tts/data_gen_utils.py:130, modules/hifigan/mel_utils.py:62).
The STFT / magnitude / mel projection / log themselves run in the fused CUDA kernel
(csrc/frontend.cu); what stays on the host is the constant mel filterbank and index arithmetic."""
import functools

import numpy as np


def _slaney_hz_to_mel(hz):
    hz = np.atleast_1d(np.asarray(hz, dtype=np.float64))
    lin = hz * (3.0 / 200.0)
    log = 15.0 + np.log(np.maximum(hz, 1e-30) / 1000.0) * (27.0 / np.log(6.4))
    return np.where(hz >= 1000.0, log, lin)


def _slaney_mel_to_hz(mel):
    mel = np.atleast_1d(np.asarray(mel, dtype=np.float64))
    lin = mel * (200.0 / 3.0)
    log = 1000.0 * np.exp((np.log(6.4) / 27.0) * (mel - 15.0))
    return np.where(mel >= 15.0, log, lin)


@functools.lru_cache(maxsize=32)
def _mel_filterbank_cached(sr, n_fft, n_mels, fmin, fmax):
    bins = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    edges = _slaney_mel_to_hz(np.linspace(_slaney_hz_to_mel(fmin)[0], _slaney_hz_to_mel(fmax)[0], n_mels + 2))
    lo, ce, hi = edges[:-2, None], edges[1:-1, None], edges[2:, None]
    up = (bins[None, :] - lo) / (ce - lo)
    down = (hi - bins[None, :]) / (hi - ce)
    tri = np.clip(np.minimum(up, down), 0.0, None).astype(np.float32)
    area = (2.0 / (hi - lo)).astype(np.float64)
    fb = (tri * area).astype(np.float32)            # float32 storage then Slaney area norm, like librosa
    fb.setflags(write=False)
    return fb


def mel_filterbank(sr, n_fft, n_mels=80, fmin=0.0, fmax=None):
    """Slaney-scale, area-normalised triangular filterbank [n_mels, n_fft//2+1] float32 -- the
    matrix ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` (0.8 defaults) returns."""
    fmax = sr / 2.0 if fmax is None else fmax
    return _mel_filterbank_cached(int(sr), int(n_fft), int(n_mels), float(fmin), float(fmax))


def build_mel_basis(hparams):
    """utils/audio.py:98-101."""
    assert hparams['fmax'] <= hparams['audio_sample_rate'] // 2
    return mel_filterbank(hparams['audio_sample_rate'], hparams['fft_size'], hparams['audio_num_mel_bins'],
                          hparams['fmin'], hparams['fmax'])


def librosa_pad_lr(x, fsize, fshift, pad_sides=1):
    """Right (or both-sides) padding that makes len(x) = (len(x)//hop + 1) * hop (utils/audio.py:67-76)."""
    assert pad_sides in (1, 2)
    pad = (x.shape[0] // fshift + 1) * fshift - x.shape[0]
    return (0, pad) if pad_sides == 1 else (pad // 2, pad - pad // 2)


def amp_to_db(x):
    return 20 * np.log10(np.maximum(1e-5, x))


def normalize(S, hparams):
    return (S - hparams['min_level_db']) / -hparams['min_level_db']


def save_wav(wav, path, sr, norm=False):
    """float waveform -> int16 wav file (utils/audio.py:11-16).  An int16 array (the device-side conversion of
    ``HifiGAN.spec2wav_batch(int16=True)``) is written as is."""
    from scipy.io import wavfile
    wav = np.asarray(wav)
    if wav.dtype == np.int16:
        wavfile.write(path, sr, wav)
        return
    wav = wav.astype(np.float32)
    if norm:
        wav = wav / np.abs(wav).max()
    wavfile.write(path, sr, (wav * np.float32(32767)).astype(np.int16))
