#!/usr/bin/env python
"""Runs a few representative generator layers once each (for ncu captures):
    ncu --set full --clock-control none --import-source on -k regex:conv1d_c4_tc -o gpurun_out/prof \
        python tools/ncu_layers.py bf16x3"""
import sys

import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from tests.test_gpu_conv_kernels import run_layer  # noqa: E402

B = 16
SHAPES = [(128, 128, 8192, 11, 5, 0, False), (32, 32, 32768, 7, 1, 0, True)]

prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
for Cin, Cout, T, K, dil, u, with_res in SHAPES:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, T, generator=g).cuda()
    w = (torch.randn(Cout, Cin, K, generator=g) * 0.02).cuda()
    b = torch.zeros(Cout).cuda()
    res = torch.randn(B, Cout, T, generator=g).cuda() if with_res else None
    _, ms = run_layer(x, w, b, res, K, dil, u, 0.1, 1.0, prec, iters=1)
    print(Cin, K, dil, with_res, f'{ms*1e3:.1f} us')
