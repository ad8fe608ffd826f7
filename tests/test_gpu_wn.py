"""WN gated-conv stack (first piece): CUDA path through the C ABI vs the reference-generated fixture
and the oracle.  Reference: modules/fastspeech/fs2_vae.py:19-94."""
import os

import numpy as np
import pytest
import torch

from neuralsvb_b200.utils import synthetic as S

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'wn.npz')
CASES = ['fvae_dec', 'fvae_enc_cond', 'dilated_cond']


def _run(name, precision):
    from neuralsvb_b200.modules.fastspeech.fs2_vae import WN
    g = np.load(GOLDEN)
    H, K, dr, L, gin, B, T = [int(v) for v in g[f'{name}/params']]
    sd = S.make_wn_state_dict(H, K, L, gin, 1234)
    x, mask, cond = S.make_wn_inputs(B, T, H, gin, 1234)
    m = WN(H, K, dr, L, gin_channels=gin, precision=precision)
    m.load_state_dict(sd, strict=True)
    m = m.eval().cuda()
    with torch.no_grad():
        y = m(x.cuda(), mask.cuda(), None if cond is None else cond.cuda()).cpu().numpy()
    return y, g[f'{name}/y']


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('precision', ['fp32', 'bf16x3'])
def test_wn_matches_reference_fixture(name, precision):
    y, ref = _run(name, precision)
    assert y.shape == ref.shape
    rel = float(np.abs(y - ref).max() / np.abs(ref).max())
    # north-star tolerance for spectral features: 1e-3 relative L-inf; the split-bf16 tensor-core path sits at ~1e-5
    assert rel < 1e-3 and rel < (2e-5 if precision == 'fp32' else 1e-4), rel


def test_wn_weight_norm_removed_and_btc_layout():
    from neuralsvb_b200.modules.fastspeech.fs2_vae import WN
    g = np.load(GOLDEN)
    H, K, dr, L, gin, B, T = [int(v) for v in g['fvae_dec/params']]
    x, mask, _ = S.make_wn_inputs(B, T, H, gin, 1234)
    m = WN(H, K, dr, L, gin_channels=gin, is_BTC=True)
    m.load_state_dict(S.make_wn_state_dict(H, K, L, gin, 1234), strict=True)
    m.remove_weight_norm()
    assert 'in_layers.0.weight' in m.state_dict() and 'in_layers.0.weight_g' not in m.state_dict()
    m = m.eval().cuda()
    with torch.no_grad():
        y = m(x.transpose(1, 2).cuda(), mask.transpose(1, 2).cuda()).transpose(1, 2).cpu().numpy()
    assert float(np.abs(y - g['fvae_dec/y']).max() / np.abs(g['fvae_dec/y']).max()) < 1e-4


def test_wn_refuses_training_and_cpu():
    from neuralsvb_b200.modules.fastspeech.fs2_vae import WN
    m = WN(64, 3, 1, 2)
    with pytest.raises(RuntimeError):
        m.eval()(torch.zeros(1, 64, 16))
    with pytest.raises(RuntimeError):
        m.train().cuda()(torch.zeros(1, 64, 16).cuda())


@pytest.mark.parametrize('name', ['global_dec', 'local_dec_nocond_mask1'])
@pytest.mark.parametrize('precision', ['fp32', 'bf16x3'])
def test_fvae_decoder_matches_reference_fixture(name, precision):
    """FVAEDecoder / GlobalFVAEDecoder (fs2_vae.py:130-152, vae_models.py:108-128): the mel that spec2wav consumes."""
    from neuralsvb_b200.modules.fastspeech.fs2_vae import FVAEDecoder, GlobalFVAEDecoder
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fvae_decoder.npz'))
    lat, H, oc, K, L, gin, B, T, glob = [int(v) for v in g[f'{name}/params']]
    m = (GlobalFVAEDecoder if glob else FVAEDecoder)(lat, H, oc, K, L, gin, strides=[4], precision=precision)
    m.load_state_dict(S.make_fvae_decoder_state_dict(lat, H, oc, K, L, gin, 4, 1234), strict=True)
    m = m.eval().cuda()
    _, mask, cond = S.make_wn_inputs(B, T, H, gin, 1234)
    z = torch.from_numpy(np.random.RandomState(1234 + 5).randn(B, lat, 1 if glob else T // 4).astype(np.float32))
    with torch.no_grad():
        y = m(z.cuda(), mask.cuda() if glob else 1, None if cond is None else cond.cuda()).cpu().numpy()
    ref = g[f'{name}/y']
    assert y.shape == ref.shape == (B, oc, T)
    rel = float(np.abs(y - ref).max() / np.abs(ref).max())
    assert rel < 1e-3 and rel < (2e-5 if precision == 'fp32' else 1e-4), rel      # north-star: 1e-3 relative L-inf on mel frames


@pytest.mark.parametrize('precision', ['fp32', 'bf16x3'])
def test_global_fvae_encoder_matches_reference_fixture(precision):
    """GlobalFVAEEncoder (vae_models.py:81-106): posterior mean / log-sigma of an utterance, eval mode."""
    from neuralsvb_b200.modules.fastspeech.fs2_vae import GlobalFVAEEncoder
    from tests.test_oracle_golden import _fvae_encoder_inputs
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fvae_encoder.npz'))
    (cin, H, lat, K, L, gin, B, T), x, mask, cond = _fvae_encoder_inputs(g)
    m = GlobalFVAEEncoder(cin, H, lat, K, L, gin, strides=[4], precision=precision)
    m.load_state_dict(S.make_fvae_encoder_state_dict(cin, H, lat, K, L, gin, 4, 1234), strict=True)
    m = m.eval().cuda()
    with torch.no_grad():
        z, mq, logs, xm = m(x.cuda(), mask.cuda(), cond.cuda(), eps=torch.zeros(B, lat, 1).cuda())
    for name, t in (('m', mq), ('logs', logs)):
        ref = g[name]
        rel = float(np.abs(t.cpu().numpy() - ref).max() / np.abs(ref).max())
        assert rel < (5e-5 if precision == 'fp32' else 2e-4), (name, rel)
    assert torch.equal(z, mq) and np.array_equal(xm.sum(-1).cpu().numpy(), g['mask_len'])


def test_global_fvae_matches_reference_fixture():
    """The whole GlobalFVAE (vae_models.py:130-146) at the vae_global_mle_eng sizes: reconstruction, KL, posterior statistics."""
    from neuralsvb_b200.modules.fastspeech.fs2_vae import GlobalFVAE
    from tests.test_oracle_golden import _global_fvae_inputs
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'global_fvae.npz'))
    (io_c, H, lat, K, Le, Ld, gin, B, T), x, mask, cond = _global_fvae_inputs(g)
    m = GlobalFVAE(io_c, H, lat, K, Le, Ld, gin, [4], False)
    m.load_state_dict(S.make_global_fvae_state_dict(io_c, H, lat, K, Le, Ld, gin, 4, 1234), strict=True)
    m = m.eval().cuda()
    with torch.no_grad():
        xr, kl, _, m_q, logs_q, xm, z_q = m(x.cuda(), mask.cuda(), cond.cuda(), infer=False, eps=torch.zeros(B, lat, 1).cuda())
        mel, z_p = m(g=cond.cuda(), infer=True)
    rel = float(np.abs(xr.cpu().numpy() - g['x_recon']).max() / np.abs(g['x_recon']).max())
    assert rel < 1e-3 and rel < 3e-4, rel                      # north-star: 1e-3 relative L-inf on mel frames
    assert abs(float(kl) - float(g['loss_kl'])) < 1e-3 * abs(float(g['loss_kl']))
    assert float(np.abs(m_q.cpu().numpy() - g['m_q']).max() / np.abs(g['m_q']).max()) < 2e-4
    assert tuple(mel.shape) == (B, io_c, T) and tuple(z_p.shape) == (B, lat, 1) and torch.isfinite(mel).all()
