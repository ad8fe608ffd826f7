"""The SVB acoustic model on the native path: MleSVBVAE.forward(infer=False, a2a + p2p + a2p) against the fixture
generated from the reference class, and the cfg-5 chain PPG extractor -> SVB mel decode -> HiFi-GAN-NSF spec2wav."""
import os

import numpy as np
import pytest
import torch

from neuralsvb_b200.utils import synthetic as S
from tests.test_oracle_golden import SVB_HP

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'svb_vae.npz')


def _model():
    from neuralsvb_b200.modules.voice_conversion.svb_vae import MleSVBVAE
    m = MleSVBVAE(80, hp=SVB_HP)
    m.load_state_dict(S.make_svb_state_dict(1234), strict=True)
    return m.eval().cuda()


def test_mle_svb_vae_matches_reference_fixture():
    g = np.load(GOLDEN)
    batch = {k: v.cuda() for k, v in S.make_svb_batch(2, 96, 120, 1234).items()}
    with torch.no_grad():
        r = _model()(**batch, infer=False, concurrent_ways=['a2a', 'p2p', 'a2p'], eps=torch.zeros(2, 128, 1).cuda())
    for name, t in (('a2p_mel', r['a2p']['mel_out']), ('a2a_mel', r['a2a']['mel_out']), ('p2p_m_q', r['p2p']['m_q'])):
        ref = g[name]
        rel = float(np.abs(t.cpu().numpy() - ref).max() / np.abs(ref).max())
        assert rel < 1e-3 and rel < 5e-4, (name, rel)          # north-star: 1e-3 relative L-inf on mel frames
    assert abs(float(r['a2p']['mle']) - float(g['a2p_mle'])) < 1e-3 * abs(float(g['a2p_mle']))
    assert abs(float(r['a2a']['kl']) - float(g['a2a_kl'])) < 1e-3 * abs(float(g['a2a_kl']))


def test_cfg5_chain_ppg_to_waveform():
    """BASELINE cfg 5 in one process: amateur mel -> VCASR PPG -> MleSVBVAE a2p mel -> HifiGAN.spec2wav (hop 256, NSF)."""
    from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator
    from neuralsvb_b200.vocoders.hifigan import HifiGAN
    batch = {k: v.cuda() for k, v in S.make_svb_batch(2, 96, 120, 1234).items()}
    with torch.no_grad():
        mel = _model()(**batch, infer=False, concurrent_ways=['a2a', 'p2p', 'a2p'])['a2p']['mel_out']        # [B, T, 80]
    assert tuple(mel.shape) == (2, 120, 80) and torch.isfinite(mel).all()
    h = S.hifigan_config()
    gen = HifiGanGenerator(h, precision='bf16x3')
    gen.load_state_dict(S.make_generator_state_dict(h, 1234), strict=True)
    gen.remove_weight_norm()
    voc = HifiGAN.from_model(gen.eval().cuda(), h)
    _, f0 = S.make_mel_f0(2, 120, 1234)
    wav = voc.spec2wav(mel[0].cpu().numpy(), f0=f0[0].numpy())
    assert wav.shape == (120 * 256,) and wav.dtype == np.float32 and np.isfinite(wav).all()
