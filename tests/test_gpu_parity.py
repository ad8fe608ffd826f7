"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures that
the unmodified reference produced.  Tolerances are the north star's: waveform <= 1e-4 RMS,
mel <= 1e-3 relative L-inf, frame / sample indexing bit exact."""
import os

import numpy as np
import pytest
import torch

from neuralsvb_b200.utils import synthetic as S
from oracle import frontend as FE
from oracle import hifigan as O
from oracle.gen_golden import FRONTEND_CASES, GEN_CASES
from tests import gpu_util as U

pytestmark = pytest.mark.gpu

WAV_RMS_TOL = 1e-4          # BASELINE.json north_star
MEL_REL_TOL = 1e-3
PRECISIONS = ['fp32', 'bf16x3', 'tf32x3']   # 1xTF32 ('tf32') is an opt-in fast mode, see test_tf32_fast_mode_error



@pytest.fixture(scope='module')
def gold(golden_dir):
    return {n: np.load(os.path.join(golden_dir, f'{n}.npz')) for n in ('frontend', 'generator', 'losses')}


# ------------------------------------------------------------------ front end
@pytest.mark.parametrize('name', list(FRONTEND_CASES))
def test_wav2spec_matches_reference_fixture(gold, name):
    from neuralsvb_b200.vocoders.hifigan import HifiGAN
    n, fft, hop, win, fmin, fmax = FRONTEND_CASES[name]
    wav = S.make_clip(n, seed=U.SEED)
    hp = dict(fft_size=fft, hop_size=hop, win_size=win, audio_num_mel_bins=80, fmin=fmin, fmax=fmax,
              audio_sample_rate=22050, min_level_db=-100)
    w2, mel, lin = HifiGAN.wav2spec(wav, return_linear=True, hp=hp)
    g = gold['frontend']
    ref = g[f'{name}/mel']
    assert mel.shape == ref.shape == (n // hop + 1, 80)               # bit-exact frame count
    assert len(w2) == int(g[f'{name}/wav_len']) and np.array_equal(w2[:n], wav) and not w2[n:].any()
    assert np.abs(mel - ref).max() / np.abs(ref).max() < MEL_REL_TOL
    assert np.abs(lin[::7, ::5] - g[f'{name}/lin_sub']).max() < 1e-3
    # and against the oracle run on this box
    _, mel_o = FE.wav2spec(wav, hp)
    assert np.abs(mel - mel_o).max() / np.abs(mel_o).max() < MEL_REL_TOL


def test_mel_spectrogram_and_stft_magnitudes_match_reference(gold):
    from neuralsvb_b200.modules.hifigan.mel_utils import mel_spectrogram
    from neuralsvb_b200.modules.parallel_wavegan.losses.stft_loss import stft
    g = gold['losses']
    h = S.hifigan_config()
    y = S.make_wave_batch(2, 8192, seed=U.SEED).cuda()
    x = (S.make_wave_batch(2, 8192, seed=U.SEED) + 0.05 * S.make_wave_batch(2, 8192, seed=U.SEED + 1)).clamp(-1, 1).cuda()
    m = mel_spectrogram(y, h).cpu().numpy()
    ref = g['mel_spectrogram/y']
    assert m.shape == ref.shape
    assert np.abs(m - ref).max() / np.abs(ref).max() < MEL_REL_TOL
    for fs, ss, wl in O.MR_STFT:
        mg = stft(x, fs, ss, wl, None).cpu().numpy()
        assert tuple(mg.shape) == tuple(g[f'stft_mag/{fs}_shape'])
        refm = g[f'stft_mag/{fs}']
        assert np.abs(mg[:, ::3, ::7] - refm).max() / np.abs(refm).max() < MEL_REL_TOL


# ------------------------------------------------------------------ NSF source
@pytest.mark.parametrize('cfg,B,T', [('small', 2, 24), ('small', 1, 37), ('hop256', 2, 40)])
def test_nsf_source_matches_oracle(cfg, B, T):
    h, hop, mel, f0, ri, nz = U.inputs(cfg, True, B, T)
    m = U.cuda_generator(cfg, True)
    m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda())
    har = m.get_tap('har_source').cpu().numpy()[:, 0]
    w = O.fold_weight_norm(S.make_generator_state_dict(h, U.SEED))
    f0_up = torch.repeat_interleave(f0, hop, dim=1)[:, :, None]
    ref, _ = O.source_module(f0_up, w, ri, nz, sr=h['audio_sample_rate'])
    ref = ref[:, :, 0].numpy()
    assert har.shape == ref.shape
    # the phase accumulation is reproduced exactly; what remains is sinf/tanhf ulp noise
    assert np.abs(har - ref).max() < 5e-6, np.abs(har - ref).max()


# ------------------------------------------------------------------ generator vs reference fixtures
@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('name', list(GEN_CASES))
def test_generator_matches_reference_fixture(gold, name, precision):
    cfg, B, T, nsf, stride = GEN_CASES[name]
    h, hop, mel, f0, ri, nz = U.inputs(cfg, nsf, B, T)
    m = U.cuda_generator(cfg, nsf, precision)
    with torch.no_grad():
        if nsf:
            y = m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda())
        else:
            y = m(mel.cuda())
    torch.cuda.synchronize()
    assert tuple(y.shape) == (B, 1, T * hop)                            # sample indexing: bit exact
    y = y.cpu().numpy()[:, 0]
    g = gold['generator']
    err = U.rms(y[:, ::stride], g[f'{name}/y_sub'])
    assert err < WAV_RMS_TOL, f'{name} {precision}: waveform RMS error {err:.3e}'
    if precision == 'fp32':
        assert err < 5e-6, err
    if nsf:
        har = m.get_tap('har_source').cpu().numpy()[:, 0]
        assert np.abs(har[:, ::stride] - g[f'{name}/har_sub']).max() < 5e-6


@pytest.mark.parametrize('precision', PRECISIONS)
def test_generator_layers_match_oracle(precision):
    """Layer-level parity on the hop-256 model: every stage boundary against the oracle."""
    cfg, B, T = 'hop256', 2, 24
    h, hop, mel, f0, ri, nz = U.inputs(cfg, True, B, T)
    m = U.cuda_generator(cfg, True, precision)
    y = m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda()).cpu().numpy()
    w = O.fold_weight_norm(S.make_generator_state_dict(h, U.SEED))
    taps = {}
    with torch.no_grad():
        y_o = O.generator_forward(w, h, mel, f0, ri, nz, taps).numpy()
    tol = 5e-5
    for name, ref in taps.items():
        got = m.get_tap(name).cpu().numpy()
        assert got.shape == tuple(ref.shape), name
        rel = np.abs(got - ref.numpy()).max() / np.abs(ref.numpy()).max()
        assert rel < tol, f'{name}: rel L-inf {rel:.3e}'
    assert U.rms(y, y_o) < WAV_RMS_TOL


def test_generator_full_size_against_oracle():
    """BASELINE config 2 (B=16 x 128 frames) against the oracle run on this box's CPU."""
    cfg, B, T = 'hop256', 16, 128
    h, hop, mel, f0, ri, nz = U.inputs(cfg, True, B, T)
    w = O.fold_weight_norm(S.make_generator_state_dict(h, U.SEED))
    with torch.no_grad():
        y_o = O.generator_forward(w, h, mel, f0, ri, nz).numpy()[:, 0]
    for precision in PRECISIONS:
        m = U.cuda_generator(cfg, True, precision)
        y = m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda()).cpu().numpy()[:, 0]
        err = U.rms(y, y_o)
        print(f'cfg2 {precision}: waveform RMS error {err:.3e} (signal RMS {U.rms(y_o):.3f})')
        assert err < WAV_RMS_TOL, (precision, err)


# ------------------------------------------------------------------ size-independent properties
@pytest.mark.parametrize('precision', PRECISIONS)
def test_batch_independence_and_causal_extent(precision):
    """Clips are independent units (the sharding unit of the multi-GPU path): a clip's waveform
    does not depend on its batch neighbours; and samples further than the receptive field from
    a cut do not depend on what follows the cut."""
    cfg, B, T = 'hop256', 4, 64
    h, hop, mel, f0, ri, nz = U.inputs(cfg, True, B, T)
    m = U.cuda_generator(cfg, True, precision)
    melc, f0c, ric, nzc = mel.cuda(), f0.cuda(), ri.cuda(), nz.cuda()
    y = m(melc, f0c, rand_ini=ric, noise=nzc)
    y1 = m(melc[2:3].contiguous(), f0c[2:3].contiguous(), rand_ini=ric[2:3].contiguous(), noise=nzc[2:3].contiguous())
    assert torch.equal(y[2:3], y1)
    Th = 40
    yh = m(melc[:, :, :Th].contiguous(), f0c[:, :Th].contiguous(), rand_ini=ric,
           noise=nzc[:, :Th * hop].contiguous())
    safe = (Th - 12) * hop       # receptive field of the stack is < 12 frames
    assert torch.allclose(y[:, :, :safe], yh[:, :, :safe], atol=1e-5, rtol=0)
    # determinism of the in-kernel Philox noise for a fixed seed; different seeds differ
    a = m(melc, f0c, seed=7)
    b = m(melc, f0c, seed=7)
    c = m(melc, f0c, seed=8)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert U.rms(a.cpu().numpy(), y.cpu().numpy()) < 0.2 and float(a.abs().max()) <= 1.0


def test_spec2wav_plugin_end_to_end():
    """HifiGAN.spec2wav through host buffers (the reference-facing call) == device-side forward."""
    from neuralsvb_b200.vocoders.hifigan import HifiGAN
    cfg, T = 'hop256', 33
    h, hop, mel, f0, _, _ = U.inputs(cfg, True, 1, T)
    m = U.cuda_generator(cfg, True, 'fp32')
    voc = HifiGAN.from_model(m, h)
    wav = voc.spec2wav(mel[0].T.numpy(), f0=f0[0].numpy(), seed=11)
    assert wav.dtype == np.float32 and wav.shape == (T * hop,)
    y = m(mel.cuda(), f0.cuda(), seed=11).cpu().numpy()[0, 0]
    assert np.array_equal(wav, y)
    wav2 = voc.spec2wav(mel[0].T.numpy())                    # non-NSF call path model(c) with an NSF model
    assert wav2.shape == (T * hop,) and np.isfinite(wav2).all()
    with pytest.raises(ValueError, match='mel'):
        m(torch.zeros(1, 64, 8, device='cuda'))


def test_tf32_fast_mode_error(gold):
    """1xTF32 is NOT parity-grade on every input (operand rounding 2^-11): it is an opt-in fast
    mode.  Record its error and check it stays within 5x of the bar; the default (bf16x3) and
    tf32x3 modes are the ones held to 1e-4."""
    cfg, B, T, nsf, stride = GEN_CASES['hop256_t16']
    h, hop, mel, f0, ri, nz = U.inputs(cfg, nsf, B, T)
    m = U.cuda_generator(cfg, nsf, 'tf32')
    y = m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda()).cpu().numpy()[:, 0]
    err = U.rms(y[:, ::stride], gold['generator']['hop256_t16/y_sub'])
    print(f'1xTF32 waveform RMS error {err:.3e}')
    assert err < 5 * WAV_RMS_TOL


# ------------------------------------------------------------------ post-filter
@pytest.mark.parametrize('win', [512, 1024])
def test_denoise_matches_oracle(win):
    """vocoder_denoise_c post-filter (vocoders/vocoder_utils.py:7-15) on the fused STFT / inverse-FFT kernels."""
    from neuralsvb_b200.vocoders.vocoder_utils import denoise
    wav = S.make_clip(256 * 60, seed=U.SEED)
    hp = dict(fft_size=1024, hop_size=256, win_size=win)
    got = denoise(wav, v=0.1, hp=hp)
    ref = FE.denoise(wav, 0.1, 1024, 256, win)
    assert got.shape == ref.shape == (256 * 60,)
    assert np.abs(got - ref).max() < 2e-5, np.abs(got - ref).max()
    assert U.rms(got, wav) > 1e-3            # the filter did change the signal
