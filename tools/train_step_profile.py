"""Kernel-time breakdown of one full vocoder training step (cfg 3) with torch.profiler (CUPTI).
    python tools/train_step_profile.py [B] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from neuralsvb_b200.modules.hifigan import discriminators as D
from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator
from neuralsvb_b200.modules.hifigan.mel_utils import mel_spectrogram
from neuralsvb_b200.tasks.vocoder.hifigan import vocoder_losses
from neuralsvb_b200.utils import synthetic as S

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
hp = dict(S.hifigan_config(True), lambda_mel=5.0, lambda_adv=1.0, use_fm_loss=False, use_ms_stft=True)
gen = HifiGanGenerator(hp, precision='bf16x3').cuda().train()
disc = torch.nn.ModuleDict({'mpd': D.MultiPeriodDiscriminator(), 'msd': D.MultiScaleDiscriminator()}).cuda().train()
y = S.make_wave_batch(B, T * 256, seed=1234)[:, None].cuda()
_, f0 = S.make_mel_f0(B, T, 1234)
f0 = f0.cuda()
with torch.no_grad():
    mel = mel_spectrogram(y.squeeze(1), hp)


def grad(mod, on):
    for p in mod.parameters():
        p.requires_grad_(on)


def step():
    grad(gen, True), grad(disc, False)
    lg, _, y_hat = vocoder_losses(gen, disc['mpd'], disc['msd'], y, mel, f0, hp, 0)
    lg.backward()
    grad(gen, False), grad(disc, True)
    ld, _, _ = vocoder_losses(None, disc['mpd'], disc['msd'], y, mel, f0, hp, 1, y_hat=y_hat)
    ld.backward()
    gen.zero_grad(), disc.zero_grad()


step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f'total device time {tot / 1e3:.1f} ms')
for e in rows[:22]:
    print(f'{e.device_time_total / 1e3:9.2f} ms {100 * e.device_time_total / tot:5.1f}% x{e.count:5d}  {e.key[:110]}')
