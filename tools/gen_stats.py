#!/usr/bin/env python
"""Per-launch pipeline diagnostics of ONE generator forward at BASELINE config 2 (SVB_TC_STATS=1: every tcgen05 conv
launch prints where its roles were blocked).   python tools/gen_stats.py [ENV=VAL ...] 2> gpurun_out/gen_stats.txt"""
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
for kv in sys.argv[1:]:
    k, v = kv.split('=')
    os.environ[k] = v
from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator  # noqa: E402
from neuralsvb_b200.utils import synthetic as S  # noqa: E402

h = S.hifigan_config()
B, T = 16, 128
m = HifiGanGenerator(h, precision='bf16x3')
m.load_state_dict(S.make_generator_state_dict(h, 1234), strict=True)
with contextlib.redirect_stdout(io.StringIO()):
    m.remove_weight_norm()
m = m.eval().cuda()
mel, f0 = S.make_mel_f0(B, T, 1234)
mel, f0 = mel.cuda(), f0.cuda()
with torch.no_grad():
    for _ in range(2):
        m(mel, f0)
    torch.cuda.synchronize()
    os.environ['SVB_TC_STATS'] = '1'
    m(mel, f0)
    torch.cuda.synchronize()
