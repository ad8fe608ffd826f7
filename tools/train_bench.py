"""Times the native generator forward (training mode: tape kept) + backward on config 2 (B=16 x 128 frames).
    python tools/train_bench.py [precision] [iters]
Weights do not change between iterations, so the host-side re-packing after an optimizer step is NOT in the
number (it is a known first-version cost)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator
from neuralsvb_b200.utils import synthetic as S

prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B, T = 16, 128
h = S.hifigan_config(True)
m = HifiGanGenerator(h, precision=prec)
m.load_state_dict(S.make_generator_state_dict(h, 1234), strict=True)
m = m.to('cuda:0').train()
mel, f0 = S.make_mel_f0(B, T, 1234)
mel, f0 = mel.cuda(), f0.cuda()
cot = torch.randn(B, 1, T * 256, device='cuda')
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
fw, bw = [], []
for it in range(iters + 2):
    m.zero_grad(set_to_none=True)
    ev[0].record()
    y = m(mel, f0, seed=3)
    ev[1].record()
    (y * cot).sum().backward()
    ev[2].record()
    torch.cuda.synchronize()
    if it >= 2:
        fw.append(ev[0].elapsed_time(ev[1])), bw.append(ev[1].elapsed_time(ev[2]))
print(f'{prec}: forward(train) {np.median(fw):.2f} ms, backward {np.median(bw):.2f} ms '
      f'(incl. fetching {len(list(m.parameters()))} parameter gradients), peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB torch')
t0 = time.time()
with torch.no_grad():
    m.conv_pre.bias.add_(0.0)          # bump a version counter: forces the re-pack path
y = m(mel, f0, seed=3)
torch.cuda.synchronize()
print(f'forward after a parameter update (host re-fold + re-pack): {time.time() - t0:.2f} s')
