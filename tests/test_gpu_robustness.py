"""-m gpu: shapes and call patterns a drop-in vocoder meets in the reference's callers: ragged lengths,
long single clips (cfg 5: 8 s), the hop-128 singing configuration, ResBlock2, changing shapes on one
handle, two handles in one process."""
import contextlib
import io

import numpy as np
import pytest
import torch

from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator
from neuralsvb_b200.utils import synthetic as S
from oracle import hifigan as O
from tests import gpu_util as U

pytestmark = pytest.mark.gpu


def _model(h, precision='bf16x3', seed=U.SEED):
    m = HifiGanGenerator(h, precision=precision)
    m.load_state_dict(S.make_generator_state_dict(h, seed), strict=True)
    with contextlib.redirect_stdout(io.StringIO()):
        m.remove_weight_norm()
    return m.eval().cuda()


def _check(h, B, T, precision='bf16x3', tol=1e-4):
    hop = int(np.prod(h['upsample_rates']))
    mel, f0 = S.make_mel_f0(B, T, U.SEED)
    nsf = h['use_pitch_embed']
    w = O.fold_weight_norm(S.make_generator_state_dict(h, U.SEED))
    m = _model(h, precision)
    if nsf:
        ri, nz = S.make_nsf_noise(B, T * hop, U.SEED)
        y = m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda())
        with torch.no_grad():
            ref = O.generator_forward(w, h, mel, f0, ri, nz)
    else:
        y = m(mel.cuda())
        with torch.no_grad():
            ref = O.generator_forward(w, h, mel)
    assert tuple(y.shape) == (B, 1, T * hop)
    err = U.rms(y.cpu().numpy(), ref.numpy())
    assert err < tol, err
    return m


@pytest.mark.parametrize('T', [1, 7, 129, 257, 689])
def test_ragged_and_long_single_clips(T):
    """T = 689 frames is BASELINE config 5's 8 s clip; tiny T exercises the padding rows."""
    _check(S.hifigan_config(), 1, T)


def test_hop128_singing_architecture():
    _check(S.hifigan_config(hop=128), 2, 75)


def test_resblock2_and_two_dilations():
    h = S.hifigan_config()
    h['resblock'] = '2'
    h['resblock_dilation_sizes'] = [[1, 3], [1, 3], [1, 3]]
    _check(h, 2, 40)


def test_shape_changes_and_second_handle_do_not_interfere():
    h = S.hifigan_config()
    m1 = _check(h, 2, 64)
    _check(h, 1, 24)                          # another handle in the same process
    mel, f0 = S.make_mel_f0(2, 64, U.SEED)
    a = m1(mel.cuda(), f0.cuda(), seed=3)
    m1(mel[:1, :, :17].contiguous().cuda(), f0[:1, :17].contiguous().cuda(), seed=3)     # smaller shape: workspace re-zeroed
    m1(torch.cat([mel, mel], 0).cuda(), torch.cat([f0, f0], 0).cuda(), seed=3)            # larger shape: workspace regrown
    b = m1(mel.cuda(), f0.cuda(), seed=3)
    assert torch.equal(a, b)


def test_errors_surface_as_exceptions_not_fallbacks():
    h = S.hifigan_config(nsf=False)
    m = _model(h)
    mel, f0 = S.make_mel_f0(1, 8, U.SEED)
    with pytest.raises(RuntimeError, match='use_pitch_embed'):
        m(mel.cuda(), f0.cuda())              # f0 given to a non-NSF model: the reference would crash too
    with pytest.raises(RuntimeError, match='CUDA tensors'):
        m(mel)


def test_long_batch_falls_back_from_merged_launches(monkeypatch):
    """8 clips x 689 frames (cfg 5's clip length) has more work items per SM than a merged launch's list holds in the last stage:
    that stage runs one launch per layer, the others stay merged -- and the result is BIT-IDENTICAL to the all-unmerged schedule
    (the chain-ordered accumulation adds the three ResBlocks in the same order as three consecutive launches)."""
    h = S.hifigan_config()
    B, T = 8, 689
    mel, f0 = S.make_mel_f0(B, T, U.SEED)
    ri, nz = S.make_nsf_noise(B, T * 256, U.SEED)
    args = (mel.cuda(), f0.cuda())
    kw = dict(rand_ini=ri.cuda(), noise=nz.cuda())
    with torch.no_grad():
        y_merged = _model(h)(*args, **kw).clone()
        monkeypatch.setenv('SVB_MERGE', '0')
        y_plain = _model(h)(*args, **kw).clone()
    assert torch.isfinite(y_merged).all()
    assert torch.equal(y_merged, y_plain)
