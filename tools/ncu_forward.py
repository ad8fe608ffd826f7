#!/usr/bin/env python
"""Three generator forwards at BASELINE config 2 (for ncu captures: skip the first two with -s, keep the third).
Per forward the tcgen05 conv kernel is launched 29 times: conv_pre, then per stage the upsampler + 6 merged ResBlock steps.
    ncu --set full --clock-control none --import-source on -k regex:conv1d_c4_tc -s 58 -c 29 -o gpurun_out/prof_fwd python tools/ncu_forward.py"""
import contextlib
import io
import sys

import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
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
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        m(mel, f0)
        torch.cuda.synchronize()
