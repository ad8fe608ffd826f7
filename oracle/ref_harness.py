"""ORACLE tooling (test infrastructure only; see oracle/__init__.py).

Imports the UNMODIFIED reference modules from /root/reference so that golden
fixtures can be generated from the reference itself.  Only usable in the build
container: /root/reference does not exist on the GPU box and nothing in the
-m gpu tests, smoke() or bench.py may import this module.

Shims -- none of them change reference arithmetic:
  * sys.modules stubs for wheels that are absent here (chardet, h5py, librosa,
    pyloudnorm, parselmouth, webrtcvad, skimage, ...).  The librosa stub
    implements ``stft`` with torch.stft(center=True, pad_mode='constant') and
    ``filters.mel`` with torchaudio's Slaney filterbank -- implementations that
    are independent of oracle/frontend.py, so they pin it.
  * torch.stft called without ``return_complex`` (torch<=1.9 idiom,
    mel_utils.py:70, losses/stft_loss.py:26) is answered with
    view_as_real(stft(..., return_complex=True)).
  * the three RNG draws of the NSF source (source.py:53,132,397) are replaced
    by injected tensors while the generator runs.
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'modules', 'hifigan'))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _torch_stft_librosa(y, n_fft=2048, hop_length=None, win_length=None, window='hann', center=True,
                        pad_mode='reflect', **_):
    """librosa.stft stand-in built on torch.stft (float64 like librosa 0.8)."""
    win_length = n_fft if win_length is None else win_length
    hop_length = win_length // 4 if hop_length is None else hop_length
    yt = torch.from_numpy(np.asarray(y)).double()
    w = torch.hann_window(win_length, periodic=True, dtype=torch.float64)
    s = torch.stft(yt, n_fft, hop_length, win_length, w, center=center, pad_mode=pad_mode, return_complex=True)
    return s.numpy().astype(np.complex64)


def _torch_istft_librosa(stft_matrix, hop_length=None, win_length=None, window='hann', center=True, length=None, **_):
    """librosa.istft stand-in built on torch.istft (window sum-square normalisation, centre padding trimmed) --
    independent of oracle/frontend.istft_librosa, so it pins it."""
    s = torch.from_numpy(np.asarray(stft_matrix)).to(torch.complex128)
    n_fft = 2 * (s.shape[0] - 1)
    win_length = n_fft if win_length is None else win_length
    hop_length = win_length // 4 if hop_length is None else hop_length
    w = torch.hann_window(win_length, periodic=True, dtype=torch.float64)
    y = torch.istft(s, n_fft, hop_length, win_length, w, center=center, length=length)
    return y.numpy().astype(np.float32)


def _torchaudio_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **_):
    import torchaudio
    fmax = sr / 2.0 if fmax is None else fmax
    fb = torchaudio.functional.melscale_fbanks(1 + n_fft // 2, float(fmin), float(fmax), n_mels, sr,
                                               norm='slaney', mel_scale='slaney')
    return fb.T.contiguous().numpy().astype(np.float32)


_installed = False


def install():
    """Put /root/reference on sys.path with the stubs above.  Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError('/root/reference is not mounted (ref_harness is build-container only)')
    sys.dont_write_bytecode = True            # the mount is read-only
    for name in ['chardet', 'h5py', 'pyloudnorm', 'parselmouth', 'webrtcvad', 'skimage', 'skimage.transform',
                 'pycwt', 'pycwt.wavelet', 'resemblyzer', 'textgrid']:
        if name not in sys.modules:
            _stub(name)
    sys.modules['skimage.transform'].resize = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    sys.modules['skimage'].transform = sys.modules['skimage.transform']
    sys.modules['textgrid'].TextGrid = object
    lib = _stub('librosa', stft=_torch_stft_librosa, istft=_torch_istft_librosa)
    lib.filters = _stub('librosa.filters', mel=_torchaudio_mel)
    lib.core = _stub('librosa.core')
    sys.path.insert(0, REF_ROOT)
    _installed = True


@contextlib.contextmanager
def legacy_stft():
    """torch.stft without return_complex -> real-stacked output (torch 1.9 behaviour)."""
    orig = torch.stft

    def shim(*a, **k):
        if k.get('return_complex') is None and len(a) < 10:
            k['return_complex'] = True
            return torch.view_as_real(orig(*a, **k))
        return orig(*a, **k)
    torch.stft = shim
    try:
        yield
    finally:
        torch.stft = orig


@contextlib.contextmanager
def cpu_cuda():
    """``Tensor.cuda()`` -> identity while the block runs: modules/parallel_wavegan/stft_loss.py:45 hard-codes
    ``.cuda()`` on its mel basis; the arithmetic is device independent."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


@contextlib.contextmanager
def injected_noise(rand_ini, noise):
    """Replace torch.rand / torch.randn_like by the pre-drawn tensors in the
    reference's draw order: rand(B,9) -> randn_like([B,T,9]) -> randn_like([B,T,1])."""
    orig_rand, orig_randn_like = torch.rand, torch.randn_like
    state = {'n': 0}

    def rand(*shape, **k):
        assert tuple(shape) == tuple(rand_ini.shape), (shape, rand_ini.shape)
        return rand_ini.clone()

    def randn_like(t, **k):
        state['n'] += 1
        if state['n'] == 1:
            assert t.shape == noise.shape, (t.shape, noise.shape)
            return noise.clone()
        return torch.zeros_like(t)            # noise branch: returned by the reference, never used
    torch.rand, torch.randn_like = rand, randn_like
    try:
        yield
    finally:
        torch.rand, torch.randn_like = orig_rand, orig_randn_like


def build_generator(h, state_dict, fold=True):
    """The reference HifiGanGenerator loaded the way vocoders/hifigan.py:17-33 does."""
    install()
    import io
    from modules.hifigan.hifigan import HifiGanGenerator
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        model = HifiGanGenerator(h)
    model.load_state_dict(state_dict, strict=True)
    if fold:
        with contextlib.redirect_stdout(io.StringIO()):
            model.remove_weight_norm()
    return model.eval()


def run_generator(model, mel, f0, rand_ini=None, noise=None):
    with torch.no_grad():
        if f0 is None:
            return model(mel)
        with injected_noise(rand_ini, noise):
            return model(mel, f0)
