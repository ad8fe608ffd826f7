"""The oracle is pinned against fixtures produced by the UNMODIFIED reference
(oracle/gen_golden.py, build container).  CPU only."""
import os

import numpy as np
import pytest
import torch

from neuralsvb_b200.utils import synthetic as S
from oracle import frontend as FE
from oracle import hifigan as O
from oracle.gen_golden import FRONTEND_CASES, GEN_CASES, SEED, _cfg


@pytest.fixture(scope='module')
def gold(golden_dir):
    return {n: np.load(os.path.join(golden_dir, f'{n}.npz')) for n in ('frontend', 'generator', 'losses')}


@pytest.mark.parametrize('name', list(FRONTEND_CASES))
def test_frontend_matches_reference(gold, name):
    n, fft, hop, win, fmin, fmax = FRONTEND_CASES[name]
    wav = S.make_clip(n, seed=SEED)
    hp = dict(fft_size=fft, hop_size=hop, win_size=win, audio_num_mel_bins=80, fmin=fmin, fmax=fmax,
              audio_sample_rate=22050, min_level_db=-100)
    w2, mel, lin = FE.wav2spec(wav, hp, return_linear=True)
    g = gold['frontend']
    ref = g[f'{name}/mel']
    assert mel.shape == ref.shape == (n // hop + 1, 80)                 # frame indexing: bit exact
    assert len(w2) == int(g[f'{name}/wav_len']) == mel.shape[0] * hop
    # tolerance of the north star: 1e-3 relative L-inf on mel frames (measured ~1e-6)
    assert np.abs(mel - ref).max() / np.abs(ref).max() < 1e-5
    assert np.abs(lin[::7, ::5] - g[f'{name}/lin_sub']).max() < 1e-4


@pytest.mark.parametrize('name', list(GEN_CASES))
def test_generator_matches_reference(gold, name):
    cfg, B, T, nsf, stride = GEN_CASES[name]
    if B * T > 256 and os.environ.get('SVB_FULL_ORACLE', '0') != '1' and (os.cpu_count() or 1) < 4:
        pytest.skip('full-size oracle case needs a few cores')
    h = _cfg(cfg, nsf)
    hop = int(np.prod(h['upsample_rates']))
    w = O.fold_weight_norm(S.make_generator_state_dict(h, SEED))
    mel, f0 = S.make_mel_f0(B, T, SEED)
    g = gold['generator']
    with torch.no_grad():
        if nsf:
            ri, nz = S.make_nsf_noise(B, T * hop, SEED)
            taps = {}
            y = O.generator_forward(w, h, mel, f0, ri, nz, taps)
            har = taps['har_source'][:, 0].numpy()[:, ::stride]
            assert np.abs(har - g[f'{name}/har_sub']).max() < 1e-6
        else:
            y = O.generator_forward(w, h, mel)
    y = y.numpy()[:, 0]
    assert y.shape == (B, T * hop)                                      # T*hop samples: bit exact
    ref = g[f'{name}/y_sub']
    rms = float(np.sqrt(((y[:, ::stride] - ref) ** 2).mean()))
    assert rms < 1e-6, rms                                              # north star: 1e-4 RMS on waveform
    np.testing.assert_allclose(np.sqrt((y.astype(np.float64) ** 2).mean(axis=1)), g[f'{name}/rms'], rtol=1e-5)


def test_mel_spectrogram_and_stft_losses_match_reference(gold):
    g = gold['losses']
    h = S.hifigan_config()
    y = S.make_wave_batch(2, 8192, seed=SEED)
    x = (y + 0.05 * S.make_wave_batch(2, 8192, seed=SEED + 1)).clamp(-1, 1)
    m = O.mel_spectrogram(y, h).numpy()
    ref = g['mel_spectrogram/y']
    assert m.shape == ref.shape == (2, 80, 8192 // 256)
    assert np.abs(m - ref).max() / np.abs(ref).max() < 1e-5
    sc, mag = O.mr_stft_loss(x, y)
    np.testing.assert_allclose([float(sc), float(mag)], g['mr_stft/sc_mag'], rtol=1e-5)
    for fs, ss, wl in O.MR_STFT:
        mg = O.stft_mag(x, fs, ss, wl).numpy()
        assert tuple(mg.shape) == tuple(g[f'stft_mag/{fs}_shape']) == (2, 1 + 8192 // ss, fs // 2 + 1)
        assert np.abs(mg[:, ::3, ::7] - g[f'stft_mag/{fs}']).max() < 1e-5


def test_mel_filterbank_matches_torchaudio_slaney():
    torchaudio = pytest.importorskip('torchaudio')
    for sr, nfft, fmin, fmax in ((22050, 1024, 80, 7600), (22050, 512, 50, 11025), (22050, 2048, 0, 11025)):
        fb = torchaudio.functional.melscale_fbanks(1 + nfft // 2, float(fmin), float(fmax), 80, sr,
                                                   norm='slaney', mel_scale='slaney').T.numpy()
        mine = FE.mel_filterbank(sr, nfft, 80, fmin, fmax)
        assert np.abs(fb - mine).max() < 1e-6
        assert (mine.sum(axis=1) > 0).all()                             # no empty filters


@pytest.mark.parametrize('name', ['mpd', 'msd'])
def test_discriminators_and_gan_losses_match_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, 'discriminators.npz'))
    sd = S.make_mpd_state_dict(SEED) if name == 'mpd' else S.make_msd_state_dict(SEED)
    w = O.fold_discriminator_weights(sd)
    y = S.make_wave_batch(2, 8192, seed=SEED)[:, None]
    y_hat = (y + 0.1 * S.make_wave_batch(2, 8192, seed=SEED + 5)[:, None]).clamp(-1, 1)
    with torch.no_grad():
        rs, gs, fr, fg = (O.mpd_forward if name == 'mpd' else O.msd_forward)(y, y_hat, w)
        losses = [float(O.feature_loss(fr, fg)), *[float(v) for v in O.discriminator_loss(rs, gs)], float(O.generator_loss(gs))]
    np.testing.assert_allclose(losses, g[f'{name}/losses'], rtol=2e-5)
    for i, (r, gg) in enumerate(zip(rs, gs)):
        assert r.shape == g[f'{name}/logit_r{i}'].shape
        assert np.abs(r.numpy() - g[f'{name}/logit_r{i}']).max() < 2e-5 * max(1.0, np.abs(g[f'{name}/logit_r{i}']).max())
        assert np.abs(gg.numpy() - g[f'{name}/logit_g{i}']).max() < 2e-5 * max(1.0, np.abs(g[f'{name}/logit_g{i}']).max())
        for j, f in enumerate(fr[i]):
            assert tuple(f.shape) == tuple(g[f'{name}/fmap_r{i}_{j}_shape'])


def test_oracle_istft_round_trip_and_denoise_shape():
    """oracle/frontend.istft_librosa (librosa 0.8.0 istft restated): STFT -> iSTFT reproduces the signal away from the
    edges, and the denoise post-filter (vocoders/vocoder_utils.py:7-15) keeps the length hop * (len // hop)."""
    import numpy as np
    from neuralsvb_b200.utils import synthetic as S
    from oracle import frontend as FE
    w = S.make_clip(256 * 20, seed=1)
    y = FE.istft_librosa(FE.stft_librosa(w, 1024, 256, 512), 256, 512)
    assert y.shape == w.shape and np.abs(y[512:-512] - w[512:-512]).max() < 1e-6
    d = FE.denoise(w, 0.1, 1024, 256, 512)
    assert d.shape == w.shape and np.abs(d).max() < np.abs(w).max()


@pytest.mark.parametrize('name', ['mpd', 'msd'])
def test_cond_discriminators_match_reference(golden_dir, name):
    """use_cond=True (mel-conditioned cond_net + 2 input channels): oracle vs the fixture the reference modules wrote."""
    g = np.load(os.path.join(golden_dir, 'discriminators_cond.npz'))
    sd = S.make_mpd_state_dict(SEED, use_cond=True) if name == 'mpd' else S.make_msd_state_dict(SEED, use_cond=True)
    w = O.fold_discriminator_weights(sd)
    y = S.make_wave_batch(2, 8192, seed=SEED)[:, None]
    y_hat = (y + 0.1 * S.make_wave_batch(2, 8192, seed=SEED + 5)[:, None]).clamp(-1, 1)
    mel, _ = S.make_mel_f0(2, 32, SEED)
    with torch.no_grad():
        rs, gs, fr, fg = (O.mpd_forward if name == 'mpd' else O.msd_forward)(y, y_hat, w, mel=mel)
        losses = [float(O.feature_loss(fr, fg)), *[float(v) for v in O.discriminator_loss(rs, gs)], float(O.generator_loss(gs)),
                  float(sum((dg ** 2).mean() for dg in gs) / len(gs))]
    np.testing.assert_allclose(losses, g[f'{name}/losses'], rtol=2e-5)
    for i, (r, gg) in enumerate(zip(rs, gs)):
        assert np.abs(r.numpy() - g[f'{name}/logit_r{i}']).max() < 2e-5 * max(1.0, np.abs(g[f'{name}/logit_r{i}']).max())
        assert np.abs(gg.numpy() - g[f'{name}/logit_g{i}']).max() < 2e-5 * max(1.0, np.abs(g[f'{name}/logit_g{i}']).max())
