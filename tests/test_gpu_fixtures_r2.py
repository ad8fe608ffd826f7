"""-m gpu: the CUDA path against the round-2 fixtures written by the UNMODIFIED reference modules
(oracle/gen_golden.py: generator_extra / generator_grads / losses_extra / discriminators_train) -- the branches that
round 1 checked against the oracle only: ResBlock2, the hop-128 architecture, the use_mel_loss STFT loss, the
denoise post-filter, save_wav's int16 conversion, training-mode spectral norm, and parameter gradients taken by
autograd through the reference modules themselves."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator
from neuralsvb_b200.utils import synthetic as S
from oracle.gen_golden import DISC_GRAD_STRIDE, GEN_EXTRA_CASES, GRAD_STRIDE, SEED, extra_config, grad_stride
from tests import gpu_util as U

pytestmark = pytest.mark.gpu


def _rel_l2_sub(t, g, prefix, k, stride):
    """Relative L2 error of a gradient on the fixture's subsample + relative error of its full L2 norm."""
    t = t.detach().double().reshape(-1).cpu()
    ref_s = torch.from_numpy(g[f'{prefix}/{k}/sub'].astype(np.float64))
    sub = t[::grad_stride(t.numel(), stride)]
    assert sub.shape == ref_s.shape, k
    ref_n = float(g[f'{prefix}/{k}/norm'])
    e_sub = float((sub - ref_s).norm() / ref_s.norm().clamp_min(1e-30))
    e_norm = abs(float(t.norm()) - ref_n) / max(ref_n, 1e-30)
    return e_sub, e_norm


@pytest.mark.parametrize('prec', ['fp32', 'bf16x3'])
@pytest.mark.parametrize('name', list(GEN_EXTRA_CASES))
def test_generator_extra_architectures_match_reference_fixture(golden_dir, name, prec):
    """a8 ResBlock2 (hifigan.py:70-91) + the hop-128 singing architecture: waveform <= 1e-4 RMS vs the reference."""
    g = np.load(os.path.join(golden_dir, 'generator_extra.npz'))
    cfg, B, T, nsf, stride = GEN_EXTRA_CASES[name]
    h = extra_config(cfg, nsf)
    hop = int(np.prod(h['upsample_rates']))
    m = HifiGanGenerator(h, precision=prec)
    m.load_state_dict(S.make_generator_state_dict(h, SEED), strict=True)
    with contextlib.redirect_stdout(io.StringIO()):
        m.remove_weight_norm()
    m = m.eval().cuda()
    mel, f0 = S.make_mel_f0(B, T, SEED)
    with torch.no_grad():
        if nsf:
            ri, nz = S.make_nsf_noise(B, T * hop, SEED)
            y = m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda())
        else:
            y = m(mel.cuda())
    y = y.cpu().numpy()[:, 0]
    assert y.shape == (B, T * hop)                                       # sample indexing: bit exact
    err = U.rms(y[:, ::stride], g[f'{name}/y_sub'])
    assert err < 1e-4, err
    np.testing.assert_allclose(np.sqrt((y.astype(np.float64) ** 2).mean(axis=1)), g[f'{name}/rms'], rtol=1e-3)


@pytest.mark.parametrize('name,cfg', [('small_nsf', None), ('small_rb2', 'small_rb2')])
def test_generator_backward_matches_reference_autograd_fixture(golden_dir, name, cfg):
    """svb_gen_backward (fp32 mode pins the logic) against gradients taken by torch autograd through the reference
    HifiGanGenerator with live weight norm: every parameter tensor, relative L2 on the stored subsample + norm."""
    g = np.load(os.path.join(golden_dir, 'generator_grads.npz'))
    h = S.small_config(True) if cfg is None else extra_config(cfg, True)
    B, T = 2, 24
    hop = int(np.prod(h['upsample_rates']))
    m = HifiGanGenerator(h, precision='fp32')
    m.load_state_dict(S.make_generator_state_dict(h, SEED), strict=True)
    m = m.cuda().train()
    mel, f0 = S.make_mel_f0(B, T, SEED)
    ri, nz = S.make_nsf_noise(B, T * hop, SEED)
    cot = torch.randn(B, 1, T * hop, generator=torch.Generator().manual_seed(7))
    y = m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda())
    assert np.abs(y.detach().cpu().numpy()[:, 0, ::3] - g[f'{name}/y_sub']).max() < 1e-5
    (y * cot.cuda()).sum().backward()
    worst = 0.0
    for k, p in m.named_parameters():
        e_sub, e_norm = _rel_l2_sub(p.grad, g, name, k, GRAD_STRIDE)
        worst = max(worst, e_sub, e_norm)
        assert e_sub < 2e-4 and e_norm < 2e-4, (k, e_sub, e_norm)
    print(f'{name}: worst relative gradient error vs the reference autograd fixture {worst:.2e}')


def test_eval_between_training_steps_keeps_backward_weights_current():
    """train forward -> optimizer step -> eval forward -> train forward/backward must differentiate the
    CURRENT weights (set_training(1) used to rebuild the data-gradient packings from the host copy of handle
    creation).  Checker: torch autograd through the oracle on the updated parameters."""
    from oracle import hifigan as O
    h = S.small_config(True)
    B, T, hop = 1, 16, 16
    m = HifiGanGenerator(h, precision='fp32')
    m.load_state_dict(S.make_generator_state_dict(h, SEED), strict=True)
    m = m.cuda().train()
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    mel, f0 = S.make_mel_f0(B, T, SEED)
    ri, nz = S.make_nsf_noise(B, T * hop, SEED)
    cot = torch.randn(B, 1, T * hop, generator=torch.Generator().manual_seed(7))
    kw = dict(rand_ini=ri.cuda(), noise=nz.cuda())
    (m(mel.cuda(), f0.cuda(), **kw) * cot.cuda()).sum().backward()
    opt.step(), opt.zero_grad()                                          # weights move a lot (lr 0.05)
    (m(mel.cuda(), f0.cuda(), **kw) * cot.cuda()).sum().backward()       # device-side re-pack happens here
    opt.zero_grad()
    m.eval()
    with torch.no_grad():
        m(mel.cuda(), f0.cuda(), **kw)                                   # validation between two optimizer steps
    m.train()
    (m(mel.cuda(), f0.cuda(), **kw) * cot.cuda()).sum().backward()
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    (O.generator_forward(O.fold_weight_norm(p), h, mel, f0, ri, nz) * cot).sum().backward()
    worst = max(float((q.grad.cpu() - p[k].grad).norm() / p[k].grad.norm().clamp_min(1e-30)) for k, q in m.named_parameters())
    assert worst < 2e-4, worst


def test_mel_stft_loss_matches_reference_fixture(golden_dir):
    """a14: STFTLoss(use_mel_loss=True) (modules/parallel_wavegan/stft_loss.py:13-100): value and d/dx."""
    from neuralsvb_b200.modules.parallel_wavegan.losses.stft_loss import multi_resolution_stft_loss
    g = np.load(os.path.join(golden_dir, 'losses_extra.npz'))
    y = S.make_wave_batch(2, 8192, seed=SEED)
    x = (y + 0.05 * S.make_wave_batch(2, 8192, seed=SEED + 1)).clamp(-1, 1)
    xc = x.cuda().requires_grad_(True)
    sc, mag = multi_resolution_stft_loss(xc, y.cuda(), use_mel_loss=True)
    np.testing.assert_allclose([float(sc), float(mag)], g['mr_stft_mel/sc_mag'], rtol=1e-3)   # north star: 1e-3 on spectral features
    (sc + mag).backward()
    dx = xc.grad.cpu()
    ref = g['mr_stft_mel/dx_sub']
    assert abs(float(dx.double().norm()) - float(g['mr_stft_mel/dx_norm'])) < 2e-3 * float(g['mr_stft_mel/dx_norm'])
    assert np.linalg.norm(dx.numpy()[:, ::5] - ref) < 2e-3 * np.linalg.norm(ref)


@pytest.mark.parametrize('win', [512, 1024])
def test_denoise_matches_reference_fixture(golden_dir, win):
    """N4: vocoder_denoise_c post-filter (vocoders/vocoder_utils.py:7-15) against the reference function itself."""
    from neuralsvb_b200.vocoders.vocoder_utils import denoise
    g = np.load(os.path.join(golden_dir, 'losses_extra.npz'))
    hp = dict(S.hifigan_config(), win_size=win)
    wav = S.make_clip(256 * 40, seed=SEED + 3)
    got = denoise(wav, v=0.1, hp=hp)
    ref = g[f'denoise/win{win}']
    assert got.shape == ref.shape
    assert U.rms(got, ref) < 1e-5 and np.abs(got - ref).max() < 1e-4


@pytest.mark.parametrize('name', ['mpd', 'msd'])
def test_training_mode_discriminators_match_reference_fixture(golden_dir, name):
    """a11 gap of round 1: MSD in train() mode -- two consecutive forwards (four power iterations of the spectral
    norm): logits and u buffers; D-loss parameter gradients and d(G loss)/d y_hat vs autograd through the reference."""
    from neuralsvb_b200.modules.hifigan import discriminators as D
    g = np.load(os.path.join(golden_dir, 'discriminators_train.npz'))
    if name == 'mpd':
        mk, sd = D.MultiPeriodDiscriminator, S.make_mpd_state_dict(SEED)
    else:
        mk, sd = D.MultiScaleDiscriminator, S.make_msd_state_dict(SEED)
    y = S.make_wave_batch(2, 8192, seed=SEED)[:, None]
    y_hat = (y + 0.1 * S.make_wave_batch(2, 8192, seed=SEED + 5)[:, None]).clamp(-1, 1)
    tol = lambda ref: 5e-4 * max(1.0, np.abs(ref).max())
    for use_tc, gtol in ((False, 1e-3), (True, 2e-2)):
        D.USE_TC = use_tc
        try:
            m = mk()
            m.load_state_dict(sd, strict=True)
            m = m.cuda().train()
            rs, gs, _, _ = m(y.cuda(), y_hat.cuda())
            r_loss, g_loss = D.discriminator_loss(rs, gs)
            np.testing.assert_allclose([float(r_loss), float(g_loss)], g[f'{name}/d_loss'], rtol=5e-4)
            (r_loss + g_loss).backward()
            for i, (r, gg) in enumerate(zip(rs, gs)):
                assert np.abs(r.detach().cpu().numpy() - g[f'{name}/fwd1/logit_r{i}']).max() < tol(g[f'{name}/fwd1/logit_r{i}'])
                assert np.abs(gg.detach().cpu().numpy() - g[f'{name}/fwd1/logit_g{i}']).max() < tol(g[f'{name}/fwd1/logit_g{i}'])
            bufs = dict(m.named_buffers())
            for k in [k for k in bufs if k.endswith('weight_u')]:
                assert np.abs(bufs[k].cpu().numpy() - g[f'{name}/fwd1/{k}']).max() < 1e-4, k
            errs = {}
            for k, p in m.named_parameters():
                e_sub, e_norm = _rel_l2_sub(p.grad, g, f'{name}/d_grad', k, DISC_GRAD_STRIDE)
                errs[k] = max(e_sub, e_norm)
            med, worst = float(np.median(list(errs.values()))), max(errs.values())
            print(f'{name} {"tc" if use_tc else "fp32"}: D-loss gradient error vs reference autograd: median {med:.1e}, worst {worst:.1e}')
            assert med < gtol / 4 and worst < gtol * 4, sorted(errs.items(), key=lambda kv: -kv[1])[:4]
            with torch.no_grad():
                rs2, gs2, _, _ = m(y.cuda(), y_hat.cuda())
            for i, (r, gg) in enumerate(zip(rs2, gs2)):
                assert np.abs(r.cpu().numpy() - g[f'{name}/fwd2/logit_r{i}']).max() < tol(g[f'{name}/fwd2/logit_r{i}'])
                assert np.abs(gg.cpu().numpy() - g[f'{name}/fwd2/logit_g{i}']).max() < tol(g[f'{name}/fwd2/logit_g{i}'])
            for k in [k for k in bufs if k.endswith('weight_u')]:
                assert np.abs(bufs[k].cpu().numpy() - g[f'{name}/fwd2/{k}']).max() < 1e-4, k
            # generator side: eval-mode weights, discriminator frozen
            m2 = mk()
            m2.load_state_dict(sd, strict=True)
            m2 = m2.cuda().eval()
            for p in m2.parameters():
                p.requires_grad_(False)
            yh = y_hat.cuda().requires_grad_(True)
            rs, gs, fr, fg = m2(y.cuda(), yh)
            lg = D.generator_loss(gs) + D.feature_loss(fr, fg)
            lg.backward()
            np.testing.assert_allclose(float(lg), float(g[f'{name}/g_loss']), rtol=5e-4)
            ref = g[f'{name}/g_dyhat_sub']
            e = np.linalg.norm(yh.grad.cpu().numpy()[:, 0, ::5] - ref) / np.linalg.norm(ref)
            # the feature loss is an L1: its gradient is sign(r - g) pulled back through leaky-relu masks, so forward
            # differences of 1e-6 flip a few signs / masks (fp32 kernels measured 7.6e-3 on MPD) -- a property of the loss
            assert e < 2e-2, e
        finally:
            D.USE_TC = True


def test_int16_conversion_matches_reference_save_wav_fixture(golden_dir):
    """N4: save_wav's float -> int16 (utils/audio.py:11-16) on the device, bit exact against the reference's wav file."""
    from neuralsvb_b200.vocoders.vocoder_utils import wav_to_int16
    g = np.load(os.path.join(golden_dir, 'losses_extra.npz'))
    wav = np.clip(S.make_clip(4000, seed=SEED + 4) * 1.7, -1.0, 1.0).astype(np.float32)
    for norm in (0, 1):
        got = wav_to_int16(torch.from_numpy(wav).cuda(), bool(norm)).cpu().numpy()
        assert got.dtype == np.int16 and np.array_equal(got, g[f'save_wav/int16_norm{norm}']), norm
    # batched, per-clip peak
    w2 = np.stack([wav, 0.25 * wav[::-1]])
    got = wav_to_int16(torch.from_numpy(w2.copy()).cuda(), True).cpu().numpy()
    from oracle import frontend as FE
    assert np.array_equal(got[0], FE.float_to_int16(w2[0], True)) and np.array_equal(got[1], FE.float_to_int16(w2[1], True))


def test_spec2wav_int16_output_is_the_converted_float_output():
    from neuralsvb_b200.vocoders.hifigan import HifiGAN
    from oracle import frontend as FE
    h = S.hifigan_config()
    m = U.cuda_generator('hop256', True, 'bf16x3')
    voc = HifiGAN.from_model(m, h)
    mel, f0 = S.make_mel_f0(2, 24, SEED)
    mels, f0s = np.ascontiguousarray(mel.permute(0, 2, 1).numpy()), f0.numpy()
    yf = voc.spec2wav_batch(mels, f0s, seed=11)
    for norm in (False, True):
        yi = voc.spec2wav_batch(mels, f0s, seed=11, int16=True, norm=norm)
        assert yi.dtype == np.int16 and yi.shape == yf.shape
        for b in range(2):
            assert np.array_equal(yi[b], FE.float_to_int16(yf[b], norm)), (norm, b)


def test_wav2spec_batch_is_the_per_clip_call_and_matches_the_reference_fixture(golden_dir):
    """N2: a ragged batch through ONE device call == wav2spec clip by clip == the reference's process_utterance."""
    from neuralsvb_b200.vocoders.hifigan import HifiGAN
    g = np.load(os.path.join(golden_dir, 'frontend.npz'))
    hp = dict(S.hifigan_config(), min_level_db=-100)
    lens = [44100, 1000, 255, 256, 7777, 1]
    wavs = [S.make_clip(n, seed=SEED) for n in lens]
    got = HifiGAN.wav2spec_batch(wavs, hp=hp)
    assert len(got) == len(wavs)
    for n, w, (wo, mel) in zip(lens, wavs, got):
        w1, m1 = HifiGAN.wav2spec(w, hp=hp)
        assert mel.shape == (n // 256 + 1, 80) and len(wo) == mel.shape[0] * 256          # frame indexing: bit exact
        assert np.array_equal(mel, m1) and np.array_equal(wo, w1)
    for name, idx in (('cfg1_win512', 0), ('ragged_short', 1)):
        ref = g[f'{name}/mel']
        assert np.abs(got[idx][1] - ref).max() / np.abs(ref).max() < 1e-3                   # north star: 1e-3 rel L-inf on mel frames


def test_binarizer_dataset_loader_and_training_step(tmp_path):
    """N2 + N3 end to end: VocoderBinarizer (batched wav2spec on the device, the reference's IndexedDataset on disk)
    -> VocoderBatchLoader (pinned ring, its own H2D stream) -> one G + D training step of HifiGanTask's wiring."""
    from neuralsvb_b200.data_gen.tts.base_binarizer import VocoderBinarizer
    from neuralsvb_b200.modules.hifigan import discriminators as D
    from neuralsvb_b200.tasks.vocoder.dataset_utils import VocoderBatchLoader
    from neuralsvb_b200.tasks.vocoder.hifigan import HifiGanTask, vocoder_losses
    from neuralsvb_b200.utils.hparams import hparams
    from neuralsvb_b200.utils.indexed_datasets import IndexedDataset
    from neuralsvb_b200.vocoders.hifigan import HifiGAN
    hp = dict(S.hifigan_config(), vocoder='neuralsvb_b200.vocoders.hifigan.HifiGAN', binary_data_dir=str(tmp_path), seed=SEED,
              lambda_mel=5.0, lambda_adv=1.0, use_fm_loss=False, use_ms_stft=False)
    hparams.clear()
    hparams.update(hp)
    try:
        clips = [S.make_clip(256 * (40 + 3 * i) + 17 * i, seed=SEED + i) for i in range(6)]
        items = {'train': [(f'clip{i}', c, 0) for i, c in enumerate(clips)]}
        f0_of = lambda wav, mel: np.full(len(mel), 220.0, np.float32)
        stats = VocoderBinarizer(items, pitch_fn=f0_of, batch_seconds=2.0).process()
        assert stats['train']['items'] == 6
        ds = IndexedDataset(str(tmp_path / 'train'))
        it = ds[2]
        w1, m1 = HifiGAN.wav2spec(clips[2], hp=hp)
        assert it['wav'].dtype == np.float16 and np.array_equal(it['mel'], m1) and it['len'] == len(m1)     # base_binarizer.py:176
        assert np.array_equal(it['wav'], w1.astype(np.float16))
        ld = VocoderBatchLoader(str(tmp_path / 'train'), 256, 8192, 3, seed=SEED, device='cuda:0')
        batches = list(ld)
        assert len(batches) == 2 and batches[0]['wavs'].is_cuda and batches[0]['mels'].shape == (3, 32, 80)
        b = batches[0]
        i = int(b['item_names'][0][4:])
        full = ds[i]
        # the crop is frame aligned: its mel rows are rows of the stored mel
        hit = [s for s in range(len(full['mel']) - 31) if np.array_equal(full['mel'][s:s + 32], b['mels'][0].cpu().numpy())]
        assert len(hit) == 1
        s = hit[0]
        assert np.array_equal(b['wavs'][0, 0].cpu().numpy(), full['wav'][s * 256:(s + 32) * 256].astype(np.float32))
        # one G + D step on the loaded batch (conditioning mel = the binarized log10-mel, [B, T, 80] -> [B, 80, T])
        gen = HifiGanGenerator(hp, precision='bf16x3').cuda().train()
        mpd, msd = D.MultiPeriodDiscriminator().cuda().train(), D.MultiScaleDiscriminator().cuda().train()
        mel = HifiGanTask._cond_mel(b, b['wavs'])
        assert mel.shape == (3, 80, 32)
        lg, _, yh = vocoder_losses(gen, mpd, msd, b['wavs'], mel, b['f0'], hp, 0)
        lg.backward()
        ld_, _, _ = vocoder_losses(None, mpd, msd, b['wavs'], mel, b['f0'], hp, 1, y_hat=yh)
        ld_.backward()
        assert np.isfinite(float(lg)) and np.isfinite(float(ld_))
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in gen.parameters())
    finally:
        hparams.clear()
