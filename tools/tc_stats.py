#!/usr/bin/env python
"""Where does conv1d_c4_tc_kernel wait?  Runs representative generator layers once each through the
instrumented instantiation (SVB_TC_STATS=1: per-role cycles blocked on each mbarrier, printed by the launcher).
    python tools/tc_stats.py [ENV=VAL ...] 2> gpurun_out/tc_stats.txt"""
import os
import sys

import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
for kv in sys.argv[1:]:
    k, v = kv.split('=')
    os.environ[k] = v
os.environ['SVB_TC_STATS'] = '1'
from tests.test_gpu_conv_kernels import run_layer  # noqa: E402

B = 16
SHAPES = [(256, 256, 1024, 11, 5, 0, False), (256, 256, 1024, 3, 1, 0, True), (128, 128, 8192, 11, 5, 0, False),
          (128, 128, 8192, 3, 1, 0, True), (64, 64, 16384, 11, 5, 0, False), (64, 64, 16384, 7, 1, 0, True),
          (64, 64, 16384, 3, 5, 0, False), (32, 32, 32768, 11, 5, 0, False), (32, 32, 32768, 11, 1, 0, True),
          (32, 32, 32768, 3, 1, 0, True), (256, 128, 1024, 16, 0, 8, False)]
for Cin, Cout, T, K, dil, u, with_res in SHAPES:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, T, generator=g).cuda()
    w = (torch.randn(*((Cin, Cout, K) if u else (Cout, Cin, K)), generator=g) * 0.02).cuda()
    b = torch.zeros(Cout).cuda()
    res = torch.randn(B, Cout, T * u if u else T, generator=g).cuda() if with_res else None
    run_layer(x, w, b, res, K, max(dil, 1), u, 0.1, 1.0, 'bf16x3', iters=1)
