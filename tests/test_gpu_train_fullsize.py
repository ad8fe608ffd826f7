"""-m gpu: the training step that bench.py TIMES (SURVEY 8(d) cfg 3: B = 16 clips x 32768 samples, hop-256 NSF
generator in bf16x3, dense discriminator layers on split-bf16 tensor cores, mel L1 + adversarial + multi-resolution
STFT) against the oracle wiring (the reference's components in eager PyTorch on the CPU, torch autograd) on the SAME
weights, inputs and NSF noise: both losses and the global gradient L2 of every generator / discriminator parameter.
The CPU side takes about a minute on the GPU box's host cores (it is the checker, never the product path)."""
import os
import time

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator
from neuralsvb_b200.utils import synthetic as S
from oracle import hifigan as O

pytestmark = pytest.mark.gpu
SEED = 1234


def _glob(got, ref):
    num = sum(float((got[k].double() - ref[k].double()).pow(2).sum()) for k in ref)
    den = sum(float(ref[k].double().pow(2).sum()) for k in ref)
    errs = {k: float((got[k].double() - ref[k].double()).norm() / ref[k].double().norm().clamp_min(1e-30)) for k in ref}
    return (num / den) ** 0.5, float(np.median(list(errs.values()))), errs


@pytest.mark.timeout(1500)
def test_full_size_cfg3_training_step_matches_oracle_wiring():
    from neuralsvb_b200.modules.hifigan import discriminators as D
    from neuralsvb_b200.tasks.vocoder.hifigan import vocoder_losses
    B, T, hop = 16, 128, 256
    h = S.hifigan_config(True)
    hp = dict(h, lambda_mel=5.0, lambda_adv=1.0, use_fm_loss=False, use_ms_stft=True)
    sd_g, sd_p, sd_s = S.make_generator_state_dict(h, SEED), S.make_mpd_state_dict(SEED), S.make_msd_state_dict(SEED)
    mel, f0 = S.make_mel_f0(B, T, SEED)
    ri, nz = S.make_nsf_noise(B, T * hop, SEED)
    y = S.make_wave_batch(B, T * hop, seed=SEED)[:, None]

    # ---- CUDA path first (frees the GPU memory before the CPU oracle needs the host)
    gen = HifiGanGenerator(h, precision='bf16x3')
    gen.load_state_dict(sd_g, strict=True)
    gen = gen.cuda().train()
    mpd, msd = D.MultiPeriodDiscriminator(), D.MultiScaleDiscriminator()
    mpd.load_state_dict(sd_p, strict=True), msd.load_state_dict(sd_s, strict=True)
    mpd, msd = mpd.cuda().eval(), msd.cuda().eval()          # eval: spectral norm without power iteration, as the oracle wiring
    for q in list(mpd.parameters()) + list(msd.parameters()):
        q.requires_grad_(False)
    lg, logs_g, yh = vocoder_losses(gen, mpd, msd, y.cuda(), mel.cuda(), f0.cuda(), hp, 0,
                                    gen_kwargs=dict(rand_ini=ri.cuda(), noise=nz.cuda()))
    lg.backward()
    got_g = {k: q.grad.cpu() for k, q in gen.named_parameters()}
    for q in list(mpd.parameters()) + list(msd.parameters()):
        q.requires_grad_(True)
    ld, logs_d, _ = vocoder_losses(None, mpd, msd, y.cuda(), mel.cuda(), f0.cuda(), hp, 1, y_hat=yh)
    ld.backward()
    got_d = {**{'mpd.' + k: q.grad.cpu() for k, q in mpd.named_parameters()},
             **{'msd.' + k: q.grad.cpu() for k, q in msd.named_parameters()}}
    lg_v, ld_v, yh_cpu = float(lg), float(ld), yh.detach().cpu()
    terms = {k: float(v) for k, v in {**logs_g, **logs_d}.items()}
    del gen, mpd, msd, lg, ld, yh
    torch.cuda.empty_cache()

    # ---- oracle wiring on the host
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    t0 = time.time()
    is_buf = lambda sd, k: k.endswith('weight_u') or (k.endswith('weight_v') and k[:-1] + 'orig' in sd)
    pg = {k: v.clone().requires_grad_(True) for k, v in sd_g.items()}
    pp = {k: v.clone().requires_grad_(True) for k, v in sd_p.items()}
    ps = {k: (v.clone() if is_buf(sd_s, k) else v.clone().requires_grad_(True)) for k, v in sd_s.items()}
    wp, ws = O.fold_discriminator_weights(pp), O.fold_discriminator_weights(ps)
    y_hat = O.generator_forward(O.fold_weight_norm(pg), h, mel, f0, ri, nz)
    assert float((yh_cpu - y_hat.detach()).pow(2).mean().sqrt()) < 1e-4          # the waveform bar, at the timed size
    _, gp, _, _ = O.mpd_forward(y, y_hat, wp)
    _, gs, _, _ = O.msd_forward(y, y_hat, ws)
    sc, mag = O.mr_stft_loss(y_hat.squeeze(1), y.squeeze(1))
    l_mel = 5.0 * F.l1_loss(O.mel_spectrogram(y_hat.squeeze(1), hp), O.mel_spectrogram(y.squeeze(1), hp))
    l_adv = O.generator_loss(gp) + O.generator_loss(gs)
    loss_g = l_mel + l_adv + sc + mag
    gg_ref = dict(zip(pg, torch.autograd.grad(loss_g, list(pg.values()))))
    rp, gp2, _, _ = O.mpd_forward(y, y_hat.detach(), wp)
    rs, gs2, _, _ = O.msd_forward(y, y_hat.detach(), ws)
    loss_d = sum(O.discriminator_loss(rp, gp2)) + sum(O.discriminator_loss(rs, gs2))
    d_params = {**{'mpd.' + k: v for k, v in pp.items()}, **{'msd.' + k: v for k, v in ps.items() if not is_buf(sd_s, k)}}
    gd_ref = dict(zip(d_params, torch.autograd.grad(loss_d, list(d_params.values()))))
    print(f'oracle wiring (CPU, {torch.get_num_threads()} threads): {time.time() - t0:.1f} s')

    # ---- losses: 1e-3 relative (north star: 1e-3 on spectral features; the adversarial terms are means over logits)
    ref_terms = {'mel': float(l_mel), 'a': float(l_adv), 'sc': float(sc), 'mag': float(mag)}
    for k, v in ref_terms.items():
        assert abs(terms[k] - v) < 1e-3 * max(abs(v), 1e-6), (k, terms[k], v)
    assert abs(lg_v - float(loss_g)) < 1e-3 * abs(float(loss_g)), (lg_v, float(loss_g))
    assert abs(ld_v - float(loss_d)) < 1e-3 * abs(float(loss_d)), (ld_v, float(loss_d))
    # ---- gradients: global relative L2 over all parameters of each optimizer
    glob_g, med_g, errs_g = _glob(got_g, gg_ref)
    glob_d, med_d, errs_d = _glob(got_d, gd_ref)
    print(f'G step: loss {lg_v:.5f} (oracle {float(loss_g):.5f}); {len(gg_ref)} gradient tensors: global {glob_g:.2e} median {med_g:.2e} '
          f'worst {max(errs_g.values()):.2e}')
    print(f'D step: loss {ld_v:.5f} (oracle {float(loss_d):.5f}); {len(gd_ref)} gradient tensors: global {glob_d:.2e} median {med_d:.2e} '
          f'worst {max(errs_d.values()):.2e}')
    # the generator gradient passes through ~10^8 leaky-relu masks computed from a forward that differs by ~4e-6 RMS
    # (bf16x3); flipped masks bound what any implementation with a different summation order can reach
    assert glob_g < 2e-2 and med_g < 1e-2, (glob_g, med_g, sorted(errs_g.items(), key=lambda kv: -kv[1])[:4])
    assert glob_d < 1e-2 and med_d < 3e-3, (glob_d, med_d, sorted(errs_d.items(), key=lambda kv: -kv[1])[:4])
