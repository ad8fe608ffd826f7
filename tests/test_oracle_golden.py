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


# ---------------------------------------------------------------------------------- round 2 fixtures (previously unpinned branches)
from oracle.gen_golden import DISC_GRAD_STRIDE, GEN_EXTRA_CASES, GRAD_STRIDE, extra_config, grad_stride  # noqa: E402


def _check_grads(g, prefix, named, stride, rtol):
    """Fixture = exact L2 norm + every `stride`-th element of each gradient tensor (gen_golden._pack_grads)."""
    worst = 0.0
    for k, t in named:
        t = t.detach().double().reshape(-1)
        ref_n, ref_s = float(g[f'{prefix}/{k}/norm']), g[f'{prefix}/{k}/sub'].astype(np.float64)
        sub = t[::grad_stride(t.numel(), stride)].numpy()
        assert sub.shape == ref_s.shape, k
        scale = max(ref_n / np.sqrt(t.numel()), 1e-30)                      # RMS element of the reference gradient
        err = float(np.abs(sub - ref_s).max() / scale)
        worst = max(worst, err)
        assert abs(float(t.norm()) - ref_n) <= rtol * max(ref_n, 1e-30), (k, float(t.norm()), ref_n)
        assert err < rtol * 50, (k, err)                                    # element-wise, in units of the RMS element
    return worst


@pytest.mark.parametrize('name', list(GEN_EXTRA_CASES))
def test_generator_extra_architectures_match_reference(golden_dir, name):
    """ResBlock2 (hifigan.py:70-91) and the hop-128 singing architecture, from the reference modules."""
    g = np.load(os.path.join(golden_dir, 'generator_extra.npz'))
    cfg, B, T, nsf, stride = GEN_EXTRA_CASES[name]
    h = extra_config(cfg, nsf)
    hop = int(np.prod(h['upsample_rates']))
    w = O.fold_weight_norm(S.make_generator_state_dict(h, SEED))
    mel, f0 = S.make_mel_f0(B, T, SEED)
    with torch.no_grad():
        if nsf:
            ri, nz = S.make_nsf_noise(B, T * hop, SEED)
            y = O.generator_forward(w, h, mel, f0, ri, nz)
        else:
            y = O.generator_forward(w, h, mel)
    y = y.numpy()[:, 0]
    assert y.shape == (B, T * hop) and tuple(g[f'{name}/meta']) == (B, T, int(nsf), stride, hop)
    assert float(np.sqrt(((y[:, ::stride] - g[f'{name}/y_sub']) ** 2).mean())) < 1e-6


@pytest.mark.parametrize('name,cfg', [('small_nsf', None), ('small_rb2', 'small_rb2')])
def test_generator_gradients_match_reference_autograd(golden_dir, name, cfg):
    """torch autograd through the oracle == torch autograd through the reference modules (weight norm live)."""
    g = np.load(os.path.join(golden_dir, 'generator_grads.npz'))
    h = S.small_config(True) if cfg is None else extra_config(cfg, True)
    B, T = 2, 24
    hop = int(np.prod(h['upsample_rates']))
    p = {k: v.clone().requires_grad_(True) for k, v in S.make_generator_state_dict(h, SEED).items()}
    mel, f0 = S.make_mel_f0(B, T, SEED)
    ri, nz = S.make_nsf_noise(B, T * hop, SEED)
    cot = torch.randn(B, 1, T * hop, generator=torch.Generator().manual_seed(7))
    y = O.generator_forward(O.fold_weight_norm(p), h, mel, f0, ri, nz)
    (y * cot).sum().backward()
    assert np.abs(y.detach().numpy()[:, 0, ::3] - g[f'{name}/y_sub']).max() < 1e-6
    _check_grads(g, name, [(k, v.grad) for k, v in p.items()], GRAD_STRIDE, 1e-4)


def test_mel_stft_loss_denoise_and_int16_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'losses_extra.npz'))
    y = S.make_wave_batch(2, 8192, seed=SEED)
    x = (y + 0.05 * S.make_wave_batch(2, 8192, seed=SEED + 1)).clamp(-1, 1)
    # use_mel_loss (modules/parallel_wavegan/stft_loss.py:43-47): values, per resolution, and d/dx
    xg = x.clone().requires_grad_(True)
    sc, mag = O.mr_stft_loss(xg, y, use_mel_loss=True)
    np.testing.assert_allclose([float(sc), float(mag)], g['mr_stft_mel/sc_mag'], rtol=2e-5)
    per = []
    for fs, ss, wl in O.MR_STFT:
        mb = torch.from_numpy(FE.mel_filterbank(22050, fs, 80)).T
        s1, m1 = O.stft_loss(x, y, fs, ss, wl, mb)
        per += [float(s1), float(m1)]
    np.testing.assert_allclose(per, g['mr_stft_mel/per_resolution'], rtol=2e-5)
    (sc + mag).backward()
    assert abs(float(xg.grad.double().norm()) - float(g['mr_stft_mel/dx_norm'])) < 1e-4 * float(g['mr_stft_mel/dx_norm'])
    assert np.abs(xg.grad.numpy()[:, ::5] - g['mr_stft_mel/dx_sub']).max() < 1e-3 * np.abs(g['mr_stft_mel/dx_sub']).max()
    # denoise (vocoders/vocoder_utils.py:7-15; librosa istft restated in oracle/frontend.py)
    for win in (512, 1024):
        wav = S.make_clip(256 * 40, seed=SEED + 3)
        d = FE.denoise(wav, 0.1, 1024, 256, win)
        ref = g[f'denoise/win{win}']
        assert d.shape == ref.shape and np.abs(d - ref).max() < 2e-6
    # save_wav float -> int16 (utils/audio.py:11-16): bit exact
    wav = np.clip(S.make_clip(4000, seed=SEED + 4) * 1.7, -1.0, 1.0).astype(np.float32)
    for norm in (0, 1):
        assert np.array_equal(FE.float_to_int16(wav, bool(norm)), g[f'save_wav/int16_norm{norm}'])


@pytest.mark.parametrize('name', ['mpd', 'msd'])
def test_training_mode_discriminators_match_reference(golden_dir, name):
    """Training-mode MPD / MSD from the reference modules: two consecutive forwards (spectral-norm power iteration of
    MSD[0]: logits + u buffers), D-loss parameter gradients, d(G adversarial + feature loss)/d y_hat."""
    g = np.load(os.path.join(golden_dir, 'discriminators_train.npz'))
    sd0 = S.make_mpd_state_dict(SEED) if name == 'mpd' else S.make_msd_state_dict(SEED)
    is_buf = lambda k: k.endswith('weight_u') or (k.endswith('weight_v') and k[:-1] + 'orig' in sd0)
    sd = {k: (v.clone() if is_buf(k) else v.clone().requires_grad_(True)) for k, v in sd0.items()}
    y = S.make_wave_batch(2, 8192, seed=SEED)[:, None]
    y_hat = (y + 0.1 * S.make_wave_batch(2, 8192, seed=SEED + 5)[:, None]).clamp(-1, 1)

    def fwd():
        if name == 'mpd':
            return O.mpd_forward(y, y_hat, O.fold_discriminator_weights(sd))
        return O.msd_forward_train(y, y_hat, sd)
    rs, gs, _, _ = fwd()
    r_loss, g_loss = O.discriminator_loss(rs, gs)
    np.testing.assert_allclose([float(r_loss), float(g_loss)], g[f'{name}/d_loss'], rtol=2e-5)
    (r_loss + g_loss).backward()
    tol = lambda ref: 3e-5 * max(1.0, np.abs(ref).max())
    for i, (r, gg) in enumerate(zip(rs, gs)):
        assert np.abs(r.detach().numpy() - g[f'{name}/fwd1/logit_r{i}']).max() < tol(g[f'{name}/fwd1/logit_r{i}'])
        assert np.abs(gg.detach().numpy() - g[f'{name}/fwd1/logit_g{i}']).max() < tol(g[f'{name}/fwd1/logit_g{i}'])
    if name == 'msd':
        for k in [k for k in sd if k.endswith('weight_u')]:
            assert np.abs(sd[k].numpy() - g[f'{name}/fwd1/{k}']).max() < 1e-5, k
    _check_grads(g, f'{name}/d_grad', [(k, v.grad) for k, v in sd.items() if not is_buf(k)], DISC_GRAD_STRIDE, 2e-4)
    with torch.no_grad():
        rs2, gs2, _, _ = fwd()
    for i, (r, gg) in enumerate(zip(rs2, gs2)):
        assert np.abs(r.numpy() - g[f'{name}/fwd2/logit_r{i}']).max() < tol(g[f'{name}/fwd2/logit_r{i}'])
        assert np.abs(gg.numpy() - g[f'{name}/fwd2/logit_g{i}']).max() < tol(g[f'{name}/fwd2/logit_g{i}'])
    if name == 'msd':
        for k in [k for k in sd if k.endswith('weight_u')]:
            assert np.abs(sd[k].numpy() - g[f'{name}/fwd2/{k}']).max() < 1e-5, k
    # generator side (eval-mode weights, discriminator frozen)
    w = O.fold_discriminator_weights(sd0)
    yh = y_hat.clone().requires_grad_(True)
    rs, gs, fr, fg = (O.mpd_forward if name == 'mpd' else O.msd_forward)(y, yh, w)
    lg = O.generator_loss(gs) + O.feature_loss(fr, fg)
    lg.backward()
    np.testing.assert_allclose(float(lg), float(g[f'{name}/g_loss']), rtol=2e-5)
    assert abs(float(yh.grad.double().norm()) - float(g[f'{name}/g_dyhat_norm'])) < 2e-4 * float(g[f'{name}/g_dyhat_norm'])
    assert np.abs(yh.grad.numpy()[:, 0, ::5] - g[f'{name}/g_dyhat_sub']).max() < 2e-3 * np.abs(g[f'{name}/g_dyhat_sub']).max()


def test_wn_oracle_matches_reference_fixture(golden_dir):
    """oracle/fs2_vae.py against the reference WN class (tests/golden/wn.npz, written by oracle/gen_golden.py gen_wn)."""
    from oracle import fs2_vae as OW
    g = np.load(os.path.join(golden_dir, 'wn.npz'))
    for name in ('fvae_dec', 'fvae_enc_cond', 'dilated_cond'):
        H, K, dr, L, gin, B, T = [int(v) for v in g[f'{name}/params']]
        w = OW.fold_weight_norm(S.make_wn_state_dict(H, K, L, gin, 1234))
        x, mask, cond = S.make_wn_inputs(B, T, H, gin, 1234)
        with torch.no_grad():
            y = OW.wn_forward(w, H, K, dr, L, x, mask, cond).numpy()
        assert float(np.abs(y - g[f'{name}/y']).max()) <= 1e-5 * float(np.abs(g[f'{name}/y']).max()), name


def test_fvae_decoder_oracle_matches_reference_fixture(golden_dir):
    """oracle/fs2_vae.py:fvae_decoder_forward against the reference FVAEDecoder / GlobalFVAEDecoder (tests/golden/fvae_decoder.npz)."""
    from oracle import fs2_vae as OW
    g = np.load(os.path.join(golden_dir, 'fvae_decoder.npz'))
    for name in ('global_dec', 'local_dec_nocond_mask1'):
        lat, H, oc, K, L, gin, B, T, glob = [int(v) for v in g[f'{name}/params']]
        w = OW.fold_weight_norm(S.make_fvae_decoder_state_dict(lat, H, oc, K, L, gin, 4, 1234))
        _, mask, cond = S.make_wn_inputs(B, T, H, gin, 1234)
        z = torch.from_numpy(np.random.RandomState(1234 + 5).randn(B, lat, 1 if glob else T // 4).astype(np.float32))
        with torch.no_grad():
            y = OW.fvae_decoder_forward(w, H, K, L, 4, z, mask if glob else 1, cond, bool(glob)).numpy()
        assert float(np.abs(y - g[f'{name}/y']).max()) <= 1e-5 * float(np.abs(g[f'{name}/y']).max()), name


def _fvae_encoder_inputs(g):
    cin, H, lat, K, L, gin, B, T = [int(v) for v in g['params']]
    rs = np.random.RandomState(1234 + 11)
    x = torch.from_numpy(rs.randn(B, cin, T).astype(np.float32))
    mask = torch.ones(B, 1, T)
    mask[1, :, T - 36:] = 0
    cond = torch.from_numpy(rs.randn(B, gin, T // 4).astype(np.float32))
    return (cin, H, lat, K, L, gin, B, T), x * mask, mask, cond


def test_fvae_encoder_oracle_matches_reference_fixture(golden_dir):
    """oracle/fs2_vae.py:global_fvae_encoder_forward against the reference GlobalFVAEEncoder (tests/golden/fvae_encoder.npz)."""
    from oracle import fs2_vae as OW
    g = np.load(os.path.join(golden_dir, 'fvae_encoder.npz'))
    (cin, H, lat, K, L, gin, B, T), x, mask, cond = _fvae_encoder_inputs(g)
    w = OW.fold_weight_norm(S.make_fvae_encoder_state_dict(cin, H, lat, K, L, gin, 4, 1234))
    with torch.no_grad():
        z, m, logs, xm = OW.global_fvae_encoder_forward(w, H, lat, K, L, 4, x, mask, cond, torch.zeros(B, lat, 1))
    for name, t in (('m', m), ('logs', logs)):
        assert float((t - torch.from_numpy(g[name])).abs().max()) <= 2e-5 * float(np.abs(g[name]).max()), name
    assert np.array_equal(xm.sum(-1).numpy(), g['mask_len'])


def _global_fvae_inputs(g):
    io_c, H, lat, K, Le, Ld, gin, B, T = [int(v) for v in g['params']]
    rs = np.random.RandomState(1234 + 23)
    x = torch.from_numpy(rs.randn(B, io_c, T).astype(np.float32))
    mask = torch.ones(B, 1, T)
    mask[1, :, T - 40:] = 0
    cond = torch.from_numpy(rs.randn(B, gin, T).astype(np.float32))
    return (io_c, H, lat, K, Le, Ld, gin, B, T), x * mask, mask, cond


def test_global_fvae_oracle_matches_reference_fixture(golden_dir):
    """oracle/fs2_vae.py:global_fvae_forward against the reference GlobalFVAE (tests/golden/global_fvae.npz)."""
    from oracle import fs2_vae as OW
    g = np.load(os.path.join(golden_dir, 'global_fvae.npz'))
    (io_c, H, lat, K, Le, Ld, gin, B, T), x, mask, cond = _global_fvae_inputs(g)
    w = OW.fold_weight_norm(S.make_global_fvae_state_dict(io_c, H, lat, K, Le, Ld, gin, 4, 1234))
    with torch.no_grad():
        xr, kl, m, logs = OW.global_fvae_forward(w, io_c, H, lat, K, Le, Ld, 4, x, mask, cond, torch.zeros(B, lat, 1))
    assert float((xr - torch.from_numpy(g['x_recon'])).abs().max()) <= 2e-5 * float(np.abs(g['x_recon']).max())
    assert abs(float(kl) - float(g['loss_kl'])) <= 1e-5 * abs(float(g['loss_kl']))


def test_vc_asr_oracle_matches_reference_fixture(golden_dir):
    """oracle/vc_asr.py against h_content of the reference VCASR (tests/golden/vc_asr.npz): Prenet + 2 conformer layers."""
    from oracle import vc_asr as OV
    g = np.load(os.path.join(golden_dir, 'vc_asr.npz'))
    B, T = [int(v) for v in g['params']]
    with torch.no_grad():
        h = OV.vc_asr_h_content(S.make_vc_asr_state_dict(1234), S.make_vc_asr_mel(B, T, 1234))
    assert float((h - torch.from_numpy(g['h_content'])).abs().max()) <= 1e-5 * float(np.abs(g['h_content']).max())


SVB_HP = {'hidden_size': 256, 'audio_num_mel_bins': 80, 'asr_enc_layers': 2, 'mel_strides': [2, 1, 1], 'asr_last_norm': False, 'latent_size': 128,
          'fvae_enc_dec_hidden': 192, 'fvae_kernel_size': 5, 'fvae_enc_n_layers': 8, 'fvae_dec_n_layers': 4}


def test_svb_vae_oracle_matches_reference_fixture(golden_dir):
    """oracle/svb_vae.py (conditions + GlobalFVAE + latent map + a2p decode) against the reference MleSVBVAE (tests/golden/svb_vae.npz)."""
    from oracle import svb_vae as OS
    g = np.load(os.path.join(golden_dir, 'svb_vae.npz'))
    with torch.no_grad():
        r = OS.mle_svb_vae_forward(S.make_svb_state_dict(1234), S.make_svb_batch(2, 96, 120, 1234))
    for name, t in (('a2p_mel', r['a2p']['mel_out']), ('a2a_mel', r['a2a']['mel_out']), ('p2p_m_q', r['p2p']['m_q'])):
        assert float((t - torch.from_numpy(g[name])).abs().max()) <= 1e-5 * float(np.abs(g[name]).max()), name
    assert abs(float(r['a2p']['mle']) - float(g['a2p_mle'])) <= 1e-5 * abs(float(g['a2p_mle']))


def test_svb_vae_drop_in_exposes_the_reference_state_dict(golden_dir):
    """Same parameter / buffer names and shapes as the reference MleSVBVAE (minus the ASR token decoder, a training-time head)."""
    from neuralsvb_b200.modules.voice_conversion.svb_vae import MleSVBVAE
    g = np.load(os.path.join(golden_dir, 'svb_vae.npz'))
    ref = {str(k): str(s) for k, s in zip(g['state_keys'], g['state_shapes'])}
    mine = {k: ','.join(str(d) for d in v.shape) for k, v in MleSVBVAE(80, hp=SVB_HP).state_dict().items()}
    assert mine == ref, (sorted(set(mine) ^ set(ref))[:10], [k for k in mine if k in ref and mine[k] != ref[k]][:10])
    assert MleSVBVAE(80, hp=SVB_HP).load_state_dict(S.make_svb_state_dict(1234), strict=True) is not None
