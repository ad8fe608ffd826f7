"""-m gpu: the native generator backward (svb_gen_backward through HifiGanGenerator's autograd node) against
torch autograd through the CPU oracle (the reference's forward, oracle/hifigan.py) on the same weights, inputs and
cotangent.  Gradients are compared per parameter tensor (weight-norm g / v, biases, noise convs, l_linear) in
relative L2; the reference has no training fixtures for the vocoder, so the oracle is the checker."""
import numpy as np
import pytest
import torch

from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator
from neuralsvb_b200.utils import synthetic as S
from oracle import hifigan as O
from tests import gpu_util as U

pytestmark = pytest.mark.gpu

# fp32 mode (every kernel on CUDA cores) pins the backward LOGIC: relative L2 per parameter tensor.
# The tensor-core modes move the forward by ~5e-6 relative, and the gradient of this leaky-relu network is NOT a
# continuous function of that: a pre-activation that changes sign flips its mask.  The CPU oracle shows it on itself
# -- perturbing its mel input by 1e-5 (forward moves 3e-6) moves its own gradients by 1.4e-3 (median over tensors)
# and 1.5e-2 (worst).  So for those modes the oracle's self-sensitivity at a matched perturbation is measured in the
# test and the CUDA gradients must stay within 3x of it (median, worst tensor and global).
REL_TOL = {'fp32': 2e-4}
SENS_EPS, SENS_FACTOR, SENS_FLOOR = 3e-5, 3.0, 2e-4


def _cfg(name, nsf):
    if name == 'small_rb2':
        h = S.small_config(nsf)
        h['resblock'] = '2'
        h['resblock_dilation_sizes'] = [[1, 3], [1, 3], [1, 3]]
        return h
    return U.config(name, nsf)


def _stats(g, g_ref):
    errs = {k: float((g[k].double() - g_ref[k].double()).norm() / g_ref[k].double().norm().clamp_min(1e-30)) for k in g_ref}
    num = sum(float((g[k].double() - g_ref[k].double()).pow(2).sum()) for k in g_ref)
    den = sum(float(g_ref[k].double().pow(2).sum()) for k in g_ref)
    return errs, (num / den) ** 0.5, float(np.median(list(errs.values()))), max(errs.values())


def _oracle_grads(h, sd, mel, f0, ri, nz, cot):
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    w = O.fold_weight_norm(p)
    y = O.generator_forward(w, h, mel, f0, ri, nz)
    (y * cot).sum().backward()
    return y.detach(), {k: v.grad for k, v in p.items()}


@pytest.mark.parametrize('cfg,nsf,B,T,prec', [
    ('small', True, 2, 24, 'fp32'),
    ('small', True, 2, 24, 'bf16x3'),
    ('small', False, 1, 37, 'bf16x3'),
    ('small_rb2', True, 2, 24, 'bf16x3'),
    ('hop256', True, 1, 12, 'bf16x3'),
])
def test_generator_backward_matches_oracle_autograd(cfg, nsf, B, T, prec):
    h = _cfg(cfg, nsf)
    hop = int(np.prod(h['upsample_rates']))
    sd = S.make_generator_state_dict(h, U.SEED)
    mel, f0 = S.make_mel_f0(B, T, U.SEED)
    f0 = f0 if nsf else None
    ri = nz = None
    if nsf:
        ri, nz = S.make_nsf_noise(B, T * hop, U.SEED)
    cot = torch.randn(B, 1, T * hop, generator=torch.Generator().manual_seed(7))
    y_ref, g_ref = _oracle_grads(h, sd, mel, f0, ri, nz, cot)

    m = HifiGanGenerator(h, precision=prec)
    m.load_state_dict(sd, strict=True)
    m = m.to('cuda:0').train()
    cu = lambda t: None if t is None else t.cuda()
    y = m(cu(mel), cu(f0), rand_ini=cu(ri), noise=cu(nz))
    assert y.requires_grad
    assert U.rms(y.detach().cpu().numpy(), y_ref.numpy()) < 1e-4
    (y * cot.cuda()).sum().backward()
    names = dict(m.named_parameters())
    assert set(names) == set(g_ref)
    errs, glob, med, worst = _stats({k: p.grad.cpu() for k, p in names.items()}, g_ref)
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print(f'{cfg} nsf={nsf} {prec}: global {glob:.1e} median {med:.1e} worst {[(k, f"{e:.1e}") for k, e in top]}')
    if prec == 'fp32':
        for k, e in errs.items():
            assert e < REL_TOL[prec], (k, e)
        return
    ref = [0.0, 0.0, 0.0]
    for s_ in range(3):
        noise = torch.randn(mel.shape, generator=torch.Generator().manual_seed(100 + s_))
        _, g_pert = _oracle_grads(h, sd, mel + SENS_EPS * noise, f0, ri, nz, cot)
        ref = [max(a, b) for a, b in zip(ref, _stats(g_pert, g_ref)[1:])]
    print(f'   oracle self-sensitivity at eps {SENS_EPS}: global {ref[0]:.1e} median {ref[1]:.1e} worst {ref[2]:.1e}')
    for got, lim, what in zip((glob, med, worst), ref, ('global', 'median', 'worst')):
        assert got < SENS_FACTOR * lim + SENS_FLOOR, (what, got, lim)


def test_training_step_updates_native_weights():
    """An SGD step changes the parameters; the next forward must see them (re-packed weights)."""
    h = S.small_config(True)
    sd = S.make_generator_state_dict(h, U.SEED)
    mel, f0 = S.make_mel_f0(2, 24, U.SEED)
    ri, nz = S.make_nsf_noise(2, 24 * 16, U.SEED)
    m = HifiGanGenerator(h, precision='bf16x3')
    m.load_state_dict(sd, strict=True)
    m = m.to('cuda:0').train()
    opt = torch.optim.SGD(m.parameters(), lr=1e-2)
    y0 = m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda())
    (y0 ** 2).mean().backward()
    opt.step()
    y1 = m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda()).detach()
    p = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    y_ref = O.generator_forward(O.fold_weight_norm(p), h, mel, f0, ri, nz)
    assert U.rms(y1.cpu().numpy(), y_ref.numpy()) < 1e-4
    assert U.rms(y1.cpu().numpy(), y0.detach().cpu().numpy()) > 1e-6        # the step did change the output
    m.eval()
    with torch.no_grad():
        y2 = m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda())
    assert U.rms(y2.cpu().numpy(), y_ref.numpy()) < 1e-4


# ------------------------------------------------------------------ discriminators + GAN losses
def _d_signals():
    y = S.make_wave_batch(2, 8192, seed=U.SEED)[:, None]
    y_hat = (y + 0.1 * S.make_wave_batch(2, 8192, seed=U.SEED + 5)[:, None]).clamp(-1, 1)
    return y, y_hat


@pytest.mark.parametrize('name', ['mpd', 'msd'])
def test_discriminator_backward_matches_oracle_autograd(name):
    """feature + generator + discriminator losses through MPD / MSD: gradients w.r.t. every discriminator parameter
    (weight-norm g / v, spectral-norm weight_orig, biases) and w.r.t. the generated waveform."""
    from neuralsvb_b200.modules.hifigan import discriminators as D
    if name == 'mpd':
        m, sd, fwd = D.MultiPeriodDiscriminator(), S.make_mpd_state_dict(U.SEED), O.mpd_forward
    else:
        m, sd, fwd = D.MultiScaleDiscriminator(), S.make_msd_state_dict(U.SEED), O.msd_forward
    y, y_hat = _d_signals()
    # oracle (eval-mode spectral norm: no power iteration)
    is_buf = lambda k: k.endswith('weight_u') or (k.endswith('weight_v') and k[:-1] + 'orig' in sd)
    p = {k: (v.clone() if is_buf(k) else v.clone().requires_grad_(True)) for k, v in sd.items()}
    yh = y_hat.clone().requires_grad_(True)
    rs, gs, fr, fg = fwd(y, yh, O.fold_discriminator_weights(p))
    dr, dg = O.discriminator_loss(rs, gs)
    (O.feature_loss(fr, fg) + O.generator_loss(gs) + dr + dg).backward()
    g_ref = {k: v.grad for k, v in p.items() if not is_buf(k)}
    g_ref['y_hat'] = yh.grad

    m.load_state_dict(sd, strict=True)
    m = m.eval().cuda()                                   # eval: spectral norm without power iteration, as the oracle
    yh2 = y_hat.cuda().requires_grad_(True)
    rs, gs, fr, fg = m(y.cuda(), yh2)
    dr, dg = D.discriminator_loss(rs, gs)
    loss = D.feature_loss(fr, fg) + D.generator_loss(gs) + dr + dg
    loss.backward()
    got = {k: q.grad.cpu() for k, q in m.named_parameters()}
    got['y_hat'] = yh2.grad.cpu()
    assert set(got) == set(g_ref)
    errs, glob, med, worst = _stats(got, g_ref)
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print(f'{name}: global {glob:.1e} median {med:.1e} worst {[(k, f"{e:.1e}") for k, e in top]}')
    assert glob < 3e-3 and med < 2e-3 and worst < 2e-2, (glob, med, worst)      # split-bf16 dense layers + a few flipped leaky-relu masks


# ------------------------------------------------------------------ spectral losses
def test_mel_and_mr_stft_loss_backward_match_oracle_autograd():
    """d/dx of  L1(mel_spectrogram(x), mel_spectrogram(y)) + sc + mag  (multi-resolution STFT loss): the fused STFT
    kernel's adjoint (svb_stft_backward) against torch autograd through torch.stft in the oracle."""
    import torch.nn.functional as F
    from neuralsvb_b200.modules.hifigan import discriminators as D
    from neuralsvb_b200.modules.hifigan.mel_utils import mel_spectrogram
    from neuralsvb_b200.modules.parallel_wavegan.losses.stft_loss import multi_resolution_stft_loss
    hp = S.hifigan_config()
    y = S.make_wave_batch(2, 8192, seed=U.SEED)
    x0 = (y + 0.05 * S.make_wave_batch(2, 8192, seed=U.SEED + 1)).clamp(-1, 1)
    for which in ('mel', 'stft', 'stft_mel'):
        xr = x0.clone().requires_grad_(True)
        if which == 'mel':
            ref = F.l1_loss(O.mel_spectrogram(xr, hp), O.mel_spectrogram(y, hp))
        else:
            sc, mag = O.mr_stft_loss(xr, y, use_mel_loss=which == 'stft_mel')
            ref = sc + mag
        ref.backward()
        xc = x0.cuda().requires_grad_(True)
        if which == 'mel':
            got = D.l1_loss(mel_spectrogram(xc, hp), mel_spectrogram(y.cuda(), hp))
        else:
            sc, mag = multi_resolution_stft_loss(xc, y.cuda(), use_mel_loss=which == 'stft_mel')
            got = sc + mag
        got.backward()
        assert abs(float(got) - float(ref)) <= 1e-3 * abs(float(ref))
        err = float((xc.grad.cpu().double() - xr.grad.double()).norm() / xr.grad.double().norm())
        print(f'{which} loss {float(got):.5f} (oracle {float(ref):.5f}), d/dx relative L2 error {err:.1e}')
        assert err < 2e-3, (which, err)


# ------------------------------------------------------------------ the composed G / D step
def test_vocoder_train_step_matches_oracle_wiring():
    """One generator step and one discriminator step of tasks/vocoder/hifigan.vocoder_losses (all loss terms on) against
    the same wiring written in eager PyTorch from the oracle's components (SURVEY 8(d) cfg 3: the reference ships no
    wiring, so results are pinned per component and the composition is checked here)."""
    import torch.nn.functional as F
    from neuralsvb_b200.modules.hifigan import discriminators as D
    from neuralsvb_b200.tasks.vocoder.hifigan import vocoder_losses
    h = S.small_config(True)
    hp = dict(S.hifigan_config(), lambda_mel=5.0, lambda_adv=1.0, use_fm_loss=True, use_ms_stft=True)
    B, T, hop = 2, 512, 16
    sd_g, sd_p, sd_s = S.make_generator_state_dict(h, U.SEED), S.make_mpd_state_dict(U.SEED), S.make_msd_state_dict(U.SEED)
    mel, f0 = S.make_mel_f0(B, T, U.SEED)
    ri, nz = S.make_nsf_noise(B, T * hop, U.SEED)
    y = S.make_wave_batch(B, T * hop, seed=U.SEED)[:, None]

    # ---- oracle wiring
    is_buf = lambda sd, k: k.endswith('weight_u') or (k.endswith('weight_v') and k[:-1] + 'orig' in sd)
    pg = {k: v.clone().requires_grad_(True) for k, v in sd_g.items()}
    pp = {k: v.clone().requires_grad_(True) for k, v in sd_p.items()}
    ps = {k: (v.clone() if is_buf(sd_s, k) else v.clone().requires_grad_(True)) for k, v in sd_s.items()}
    wp, ws = O.fold_discriminator_weights(pp), O.fold_discriminator_weights(ps)
    y_hat = O.generator_forward(O.fold_weight_norm(pg), h, mel, f0, ri, nz)
    _, gp, frp, fgp = O.mpd_forward(y, y_hat, wp)
    _, gs, frs, fgs = O.msd_forward(y, y_hat, ws)
    sc, mag = O.mr_stft_loss(y_hat.squeeze(1), y.squeeze(1))
    loss_g = (5.0 * F.l1_loss(O.mel_spectrogram(y_hat.squeeze(1), hp), O.mel_spectrogram(y.squeeze(1), hp))
              + O.generator_loss(gp) + O.generator_loss(gs) + O.feature_loss(frp, fgp) + O.feature_loss(frs, fgs) + sc + mag)
    gg_ref = dict(zip(pg, torch.autograd.grad(loss_g, list(pg.values()))))
    rp, gp2, _, _ = O.mpd_forward(y, y_hat.detach(), wp)
    rs, gs2, _, _ = O.msd_forward(y, y_hat.detach(), ws)
    loss_d = sum(O.discriminator_loss(rp, gp2)) + sum(O.discriminator_loss(rs, gs2))
    d_params = {**{'mpd.' + k: v for k, v in pp.items()}, **{'msd.' + k: v for k, v in ps.items() if not is_buf(sd_s, k)}}
    gd_ref = dict(zip(d_params, torch.autograd.grad(loss_d, list(d_params.values()))))

    # ---- CUDA path
    gen = HifiGanGenerator(h, precision='fp32')
    gen.load_state_dict(sd_g, strict=True)
    gen = gen.cuda().train()
    mpd, msd = D.MultiPeriodDiscriminator(), D.MultiScaleDiscriminator()
    mpd.load_state_dict(sd_p, strict=True), msd.load_state_dict(sd_s, strict=True)
    mpd, msd = mpd.cuda().eval(), msd.cuda().eval()          # eval: spectral norm without power iteration, as the oracle
    for q in list(mpd.parameters()) + list(msd.parameters()):
        q.requires_grad_(False)                              # what the trainer does for the other optimizer's parameters
    lg, logs, yh = vocoder_losses(gen, mpd, msd, y.cuda(), mel.cuda(), f0.cuda(), hp, 0,
                                  gen_kwargs=dict(rand_ini=ri.cuda(), noise=nz.cuda()))
    lg.backward()
    assert abs(float(lg) - float(loss_g)) < 2e-3 * abs(float(loss_g)), (float(lg), float(loss_g))
    errs, glob, med, worst = _stats({k: q.grad.cpu() for k, q in gen.named_parameters()}, gg_ref)
    print(f'G step: loss {float(lg):.4f} (oracle {float(loss_g):.4f}) grads global {glob:.1e} median {med:.1e} worst {worst:.1e}')
    assert glob < 1e-2 and med < 5e-3, (glob, med, worst)
    for q in list(mpd.parameters()) + list(msd.parameters()):
        q.requires_grad_(True)
    ld, _, _ = vocoder_losses(None, mpd, msd, y.cuda(), mel.cuda(), f0.cuda(), hp, 1, y_hat=yh)
    ld.backward()
    assert abs(float(ld) - float(loss_d)) < 2e-3 * abs(float(loss_d)), (float(ld), float(loss_d))
    got = {**{'mpd.' + k: q.grad.cpu() for k, q in mpd.named_parameters()}, **{'msd.' + k: q.grad.cpu() for k, q in msd.named_parameters()}}
    errs, glob, med, worst = _stats(got, gd_ref)
    print(f'D step: loss {float(ld):.4f} (oracle {float(loss_d):.4f}) grads global {glob:.1e} median {med:.1e} worst {worst:.1e}')
    assert glob < 1e-2 and med < 1e-3, (glob, med, worst)


# ------------------------------------------------------------------ call patterns of a training loop
def _train_model(h, prec='fp32'):
    m = HifiGanGenerator(h, precision=prec)
    m.load_state_dict(S.make_generator_state_dict(h, U.SEED), strict=True)
    return m.cuda().train()


@pytest.mark.parametrize('B,T', [(1, 1), (3, 5), (1, 130)])
def test_backward_ragged_shapes_fp32(B, T):
    """One frame, odd batch, more than one 256-row tile per clip at the first stage: exact (fp32) gradient parity."""
    h = S.small_config(True)
    sd = S.make_generator_state_dict(h, U.SEED)
    mel, f0 = S.make_mel_f0(B, T, U.SEED)
    ri, nz = S.make_nsf_noise(B, T * 16, U.SEED)
    cot = torch.randn(B, 1, T * 16, generator=torch.Generator().manual_seed(3))
    _, g_ref = _oracle_grads(h, sd, mel, f0, ri, nz, cot)
    m = _train_model(h)
    (m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda()) * cot.cuda()).sum().backward()
    errs, glob, med, worst = _stats({k: p.grad.cpu() for k, p in m.named_parameters()}, g_ref)
    assert worst < 2e-4, (worst, sorted(errs.items(), key=lambda kv: -kv[1])[:3])


def test_backward_is_repeatable_and_accumulates_like_autograd():
    """Two forward/backward passes without zero_grad double the gradients (torch semantics); the native buffers are
    cleared per backward, so a second call does not see the first one's atomics."""
    h = S.small_config(True)
    mel, f0 = S.make_mel_f0(2, 24, U.SEED)
    ri, nz = S.make_nsf_noise(2, 24 * 16, U.SEED)
    m = _train_model(h)
    args = (mel.cuda(), f0.cuda())
    kw = dict(rand_ini=ri.cuda(), noise=nz.cuda())
    rel = lambda x, y: float((x - y).norm() / y.norm().clamp_min(1e-30))      # fp32 atomics: equal up to summation order
    m(*args, **kw).pow(2).mean().backward()
    g1 = {k: p.grad.clone() for k, p in m.named_parameters()}
    m(*args, **kw).pow(2).mean().backward()
    for k, p in m.named_parameters():
        assert rel(p.grad, 2 * g1[k]) < 1e-4, (k, rel(p.grad, 2 * g1[k]))
    m.zero_grad()
    m(*args, **kw).pow(2).mean().backward()
    for k, p in m.named_parameters():
        assert rel(p.grad, g1[k]) < 1e-4, (k, rel(p.grad, g1[k]))


def test_backward_without_training_forward_fails_loudly():
    import ctypes
    from neuralsvb_b200 import _native
    h = S.small_config(True)
    m = _train_model(h).eval()
    mel, f0 = S.make_mel_f0(1, 8, U.SEED)
    with torch.no_grad():
        m(mel.cuda(), f0.cuda())
    dy = torch.zeros(1, 1, 8 * 16, device='cuda')
    rc = _native.lib().svb_gen_backward(m.native_handle(), _native.ptr(dy), None)
    assert rc != 0 and b'training' in _native.lib().svb_last_error()
