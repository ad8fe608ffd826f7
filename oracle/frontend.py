"""ORACLE (test infrastructure only; see oracle/__init__.py).

numpy restatement of the reference's wav -> log-mel front end.

Follows, in order:
  data_gen/tts/data_gen_utils.py:93-147   process_utterance
  utils/audio.py:57-59,67-76,98-117       _stft, librosa_pad_lr, _build_mel_basis, amp_to_db, normalize
  vocoders/pwg.py:105-122                 PWG.wav2spec (argument wiring, eps, transposes)

Third-party arithmetic restated from its published algorithm: librosa==0.8.0
(Requirements.txt:41; not vendored under /root/reference, not installable
here).  ``librosa.stft`` = centre zero-pad n_fft//2, periodic hann(win_length)
zero-padded to n_fft, frames at stride hop, float64 window * float32 frame,
rfft, cast to complex64.  ``librosa.filters.mel`` = Slaney mel scale,
triangular filters, Slaney area normalisation, float32 storage.
"""
import numpy as np


# ---------------------------------------------------------------- librosa.filters.mel
def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        log_t = f >= min_log_hz
        mels[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if m.ndim:
        log_t = m >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
    elif m >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (m - min_log_mel))
    return freqs


def mel_filterbank(sr, n_fft, n_mels=80, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with htk=False,
    norm='slaney', dtype=float32 (call sites: data_gen_utils.py:130,
    utils/audio.py:100, mel_utils.py:62, parallel_wavegan/stft_loss.py:45)."""
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = 1 + n_fft // 2
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


# ---------------------------------------------------------------- librosa.stft
def hann_periodic(win_length):
    """scipy.signal.get_window('hann', N, fftbins=True), float64."""
    n = np.arange(win_length, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)


def stft_librosa(y, n_fft, hop_length, win_length=None, pad_mode='constant'):
    """librosa.stft(y, n_fft, hop_length, win_length, window='hann',
    center=True, pad_mode=...) -> complex64 [1 + n_fft//2, 1 + len(y)//hop]."""
    if win_length is None:
        win_length = n_fft
    win = hann_periodic(win_length)
    lpad = (n_fft - win_length) // 2
    win = np.pad(win, (lpad, n_fft - win_length - lpad))
    y = np.asarray(y)
    yp = np.pad(y, n_fft // 2, mode=pad_mode)
    n_frames = 1 + (len(yp) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    frames = yp[idx]                                   # [n_fft, n_frames], dtype of y
    return np.fft.rfft(win[:, None] * frames, axis=0).astype(np.complex64)


def istft_librosa(stft_matrix, hop_length, win_length=None):
    """librosa.istft(stft_matrix, hop_length, win_length, window='hann', center=True, length=None) (librosa 0.8.0,
    core/spectrum.py): per-frame irfft * padded window, overlap-add, division by the window sum-square where it is
    above float32 tiny, centre padding n_fft//2 trimmed from both ends."""
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    if win_length is None:
        win_length = n_fft
    win = hann_periodic(win_length)
    lpad = (n_fft - win_length) // 2
    win = np.pad(win, (lpad, n_fft - win_length - lpad))
    n_frames = stft_matrix.shape[1]
    expected = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(expected, dtype=np.float32)
    wss = np.zeros(expected, dtype=np.float32)
    frames = np.fft.irfft(stft_matrix, n=n_fft, axis=0)
    for i in range(n_frames):
        y[i * hop_length:i * hop_length + n_fft] += (win * frames[:, i]).astype(np.float32)
        wss[i * hop_length:i * hop_length + n_fft] += (win ** 2).astype(np.float32)
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2:expected - n_fft // 2]


def denoise(wav, v, fft_size, hop_size, win_size):
    """vocoders/vocoder_utils.py:7-15."""
    spec = stft_librosa(wav, fft_size, hop_size, win_size, pad_mode='constant')
    spec_m = np.clip(np.abs(spec) - v, a_min=0, a_max=None)
    spec_a = np.angle(spec)
    return istft_librosa(spec_m * np.exp(1j * spec_a), hop_size, win_size)


# ---------------------------------------------------------------- utils/audio.py helpers
def librosa_pad_lr(x, fsize, fshift, pad_sides=1):
    """utils/audio.py:67-76."""
    assert pad_sides in (1, 2)
    pad = (x.shape[0] // fshift + 1) * fshift - x.shape[0]
    if pad_sides == 1:
        return 0, pad
    return pad // 2, pad // 2 + pad % 2


def amp_to_db(x):
    """utils/audio.py:104-105."""
    return 20 * np.log10(np.maximum(1e-5, x))


def normalize(S, min_level_db):
    """utils/audio.py:112-113."""
    return (S - min_level_db) / -min_level_db


# ---------------------------------------------------------------- process_utterance
def process_utterance(wav, fft_size=1024, hop_size=256, win_length=1024, num_mels=80, fmin=80, fmax=7600,
                      eps=1e-6, sample_rate=22050, min_level_db=-100, return_linear=False):
    """data_gen/tts/data_gen_utils.py:93-147 for an in-memory waveform
    (``wav_path`` not a str, :112-113), loud_norm off, vocoder='pwg'.
    Returns (wav[T*hop], mel[num_mels, T]) (+ normalised dB linear spec)."""
    wav = np.asarray(wav)
    x_stft = stft_librosa(wav, n_fft=fft_size, hop_length=hop_size, win_length=win_length, pad_mode='constant')
    spc = np.abs(x_stft)                                # (n_bins, T) float32
    fmin = 0 if fmin == -1 else fmin
    fmax = sample_rate / 2 if fmax == -1 else fmax
    mel_basis = mel_filterbank(sample_rate, fft_size, num_mels, fmin, fmax)
    mel = mel_basis @ spc
    mel = np.log10(np.maximum(eps, mel))
    l_pad, r_pad = librosa_pad_lr(wav, fft_size, hop_size, 1)
    wav = np.pad(wav, (l_pad, r_pad), mode='constant', constant_values=0.0)
    wav = wav[:mel.shape[1] * hop_size]
    if not return_linear:
        return wav, mel
    spc = normalize(amp_to_db(spc), min_level_db)
    return wav, mel, spc


def wav2spec(wav, hp, return_linear=False):
    """vocoders/pwg.py:105-122 argument wiring (eps default 1e-10, outputs
    transposed to [T, 80] / [T, n_bins])."""
    res = process_utterance(
        wav, fft_size=hp['fft_size'], hop_size=hp['hop_size'], win_length=hp['win_size'],
        num_mels=hp['audio_num_mel_bins'], fmin=hp['fmin'], fmax=hp['fmax'],
        sample_rate=hp['audio_sample_rate'], min_level_db=hp.get('min_level_db', -100),
        return_linear=return_linear, eps=float(hp.get('wav2spec_eps', 1e-10)))
    if return_linear:
        return res[0], res[1].T, res[2].T
    return res[0], res[1].T


def float_to_int16(wav, norm=False):
    """save_wav's sample conversion (utils/audio.py:11-16): optional peak normalisation, ``wav * 32767``, then
    numpy's float -> int16 cast (truncation toward zero)."""
    wav = np.asarray(wav, np.float32)
    if norm:
        wav = wav / np.abs(wav).max()
    return (wav * np.float32(32767)).astype(np.int16)
