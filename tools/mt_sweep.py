#!/usr/bin/env python
"""Tuning / ablation sweep for the tcgen05 conv kernel at BASELINE config 2.
    python tools/mt_sweep.py <precision> <ENV_VAR> <v1,v2,...> [FIXED=VAL ...]
Prints one row per generator layer shape (us per launch) and one column per value."""
import os
import sys

import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from tests.test_gpu_conv_kernels import run_layer  # noqa: E402
from tools.layer_bench import B, SHAPES  # noqa: E402


def main():
    prec, var, vals = sys.argv[1], sys.argv[2], sys.argv[3].split(',')
    for kv in sys.argv[4:]:
        k, v = kv.split('=')
        os.environ[k] = v
    print('layer'.ljust(24) + ''.join(f'{var[-3:]}={v}'.rjust(9) for v in vals))
    for name, Cin, Cout, T, K, dil, u, with_res in SHAPES:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, Cin, T, generator=g).cuda()
        w = (torch.randn(*((Cin, Cout, K) if u else (Cout, Cin, K)), generator=g) * 0.02).cuda()
        b = torch.zeros(Cout).cuda()
        res = torch.randn(B, Cout, T * u if u else T, generator=g).cuda() if with_res else None
        row = name.ljust(24)
        for v in vals:
            os.environ[var] = v
            _, ms = run_layer(x, w, b, res, K, max(dil, 1), u, 0.1, 1.0, prec, iters=10)
            row += f'{ms * 1e3:9.1f}'
        print(row, flush=True)


if __name__ == '__main__':
    main()
