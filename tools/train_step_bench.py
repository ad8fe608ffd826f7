"""Times the full vocoder training step of SURVEY 8(d) cfg 3 on one GPU: G step (generator forward, mel L1,
adversarial losses through MPD + MSD, backward into G, clip, AdamW) + D step (MPD + MSD on y and y_hat.detach(),
LSGAN losses, backward, clip, AdamW).
    python tools/train_step_bench.py [B] [frames] [iters] [use_ms_stft]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from neuralsvb_b200.modules.hifigan import discriminators as D
from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator
from neuralsvb_b200.modules.hifigan.mel_utils import mel_spectrogram
from neuralsvb_b200.tasks.vocoder.hifigan import vocoder_losses
from neuralsvb_b200.utils import synthetic as S

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
hp = dict(S.hifigan_config(True), lambda_mel=5.0, lambda_adv=1.0, use_fm_loss=False, use_ms_stft=bool(int(sys.argv[4])) if len(sys.argv) > 4 else True)
gen = HifiGanGenerator(hp, precision='bf16x3').cuda().train()
disc = torch.nn.ModuleDict({'mpd': D.MultiPeriodDiscriminator(), 'msd': D.MultiScaleDiscriminator()}).cuda().train()
og = torch.optim.AdamW(gen.parameters(), lr=2e-4, betas=(0.8, 0.99))
od = torch.optim.AdamW(disc.parameters(), lr=2e-4, betas=(0.8, 0.99))
y = S.make_wave_batch(B, T * 256, seed=1234)[:, None].cuda()
_, f0 = S.make_mel_f0(B, T, 1234)
f0 = f0.cuda()
with torch.no_grad():
    mel = mel_spectrogram(y.squeeze(1), hp)


def grad(mod, on):
    for p in mod.parameters():
        p.requires_grad_(on)


ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tg, td, wall = [], [], []
for it in range(iters + 1):
    torch.cuda.synchronize()
    t0 = time.time()
    ev[0].record()
    grad(gen, True), grad(disc, False)
    lg, logs, y_hat = vocoder_losses(gen, disc['mpd'], disc['msd'], y, mel, f0, hp, 0)
    lg.backward()
    torch.nn.utils.clip_grad_norm_(gen.parameters(), 10.0)
    og.step(), og.zero_grad()
    ev[1].record()
    grad(gen, False), grad(disc, True)
    ld, _, _ = vocoder_losses(None, disc['mpd'], disc['msd'], y, mel, f0, hp, 1, y_hat=y_hat)
    ld.backward()
    torch.nn.utils.clip_grad_norm_(disc.parameters(), 1.0)
    od.step(), od.zero_grad()
    ev[2].record()
    torch.cuda.synchronize()
    if it >= 1:
        tg.append(ev[0].elapsed_time(ev[1])), td.append(ev[1].elapsed_time(ev[2])), wall.append(time.time() - t0)
    print(f'it {it}: G loss {float(lg):.4f} D loss {float(ld):.4f}', flush=True)
g, d = float(np.median(tg)), float(np.median(td))
print(f'cfg3 B={B} x {T * 256} samples, ms_stft={hp["use_ms_stft"]}: G step {g:.1f} ms, D step {d:.1f} ms, total {g + d:.1f} ms '
      f'= {1000.0 / (g + d):.3f} steps/s (wall {np.median(wall) * 1000:.0f} ms), peak torch mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB')
