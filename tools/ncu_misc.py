#!/usr/bin/env python
"""One pass over the secondary kernels for ncu captures: the fused front end (stft_mel_v2), the NSF source, one full
G + D training step (weight gradient, discriminator layout kernels, losses).
    ncu --set full --clock-control none -k regex:'stft_mel|nsf_|wgrad_tc|expand_to_g32t|col2im|g32t_to_nctw|conv_fewout|gconv_kernel' \
        -c 60 -o gpurun_out/prof_misc python tools/ncu_misc.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from neuralsvb_b200.modules.hifigan import discriminators as D  # noqa: E402
from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator  # noqa: E402
from neuralsvb_b200.modules.hifigan.mel_utils import mel_spectrogram, wav2spec_mel  # noqa: E402
from neuralsvb_b200.tasks.vocoder.hifigan import vocoder_losses  # noqa: E402
from neuralsvb_b200.utils import synthetic as S  # noqa: E402

hp = dict(S.hifigan_config(True), lambda_mel=5.0, lambda_adv=1.0, use_fm_loss=False, use_ms_stft=True)
wav = torch.from_numpy(np.stack([S.make_clip(44100, seed=1234 + i) for i in range(8)])).repeat(32, 1).cuda()
wav2spec_mel(wav, hp, frames=44100 // 256 + 1)                     # cfg 1 front end, 256 clips in one launch
B, T = 4, 128                                                       # a small batch keeps the ~40 replays per launch short
gen = HifiGanGenerator(hp, precision='bf16x3').cuda().train()
disc = torch.nn.ModuleDict({'mpd': D.MultiPeriodDiscriminator(), 'msd': D.MultiScaleDiscriminator()}).cuda().train()
y = S.make_wave_batch(B, T * 256, seed=1234)[:, None].cuda()
_, f0 = S.make_mel_f0(B, T, 1234)
f0 = f0.cuda()
with torch.no_grad():
    mel = mel_spectrogram(y.squeeze(1), hp)
for p in disc.parameters():
    p.requires_grad_(False)
lg, _, y_hat = vocoder_losses(gen, disc['mpd'], disc['msd'], y, mel, f0, hp, 0)
lg.backward()
for p in disc.parameters():
    p.requires_grad_(True)
ld, _, _ = vocoder_losses(None, disc['mpd'], disc['msd'], y, mel, f0, hp, 1, y_hat=y_hat)
ld.backward()
torch.cuda.synchronize()
