"""Shared helpers of the -m gpu parity tests (CUDA path vs the CPU oracle / golden fixtures)."""
import numpy as np
import torch

from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator
from neuralsvb_b200.utils import synthetic as S

SEED = 1234
_models = {}


def config(name, nsf):
    return S.small_config(nsf) if name == 'small' else S.hifigan_config(nsf)


def cuda_generator(cfg_name, nsf, precision='fp32'):
    """Product generator on cuda:0, loaded the way vocoders/hifigan.py:17-33 does."""
    key = (cfg_name, nsf)
    if key not in _models:
        h = config(cfg_name, nsf)
        m = HifiGanGenerator(h, precision=precision)
        m.load_state_dict(S.make_generator_state_dict(h, SEED), strict=True)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            m.remove_weight_norm()
        _models[key] = m.eval().to('cuda:0')
    m = _models[key]
    m.set_precision(precision)
    return m


def inputs(cfg_name, nsf, B, T):
    h = config(cfg_name, nsf)
    hop = int(np.prod(h['upsample_rates']))
    mel, f0 = S.make_mel_f0(B, T, SEED)
    ri = nz = None
    if nsf:
        ri, nz = S.make_nsf_noise(B, T * hop, SEED)
    return h, hop, mel, (f0 if nsf else None), ri, nz


def rms(a, b=None):
    a = np.asarray(a, np.float64)
    if b is not None:
        a = a - np.asarray(b, np.float64)
    return float(np.sqrt((a ** 2).mean()))
