#!/usr/bin/env python
"""Per-layer timing of the generator's distinct convolution shapes at BASELINE config 2
(B=16 x 128 frames) through svb_conv1d_run: CUDA-core vs tcgen05 kernels, with the achieved
TFLOP/s and the streamed GB/s (input + output [+ residual] bytes of the layer)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from tests.test_gpu_conv_kernels import run_layer  # noqa: E402

B = 16
# (name, Cin, Cout, T, K, dil, u, residual)
SHAPES = []
for C, T in ((256, 1024), (128, 8192), (64, 16384), (32, 32768)):
    for K in (3, 7, 11):
        SHAPES.append((f'rb C{C} k{K} d1 +res', C, C, T, K, 1, 0, True))
        SHAPES.append((f'rb C{C} k{K} d5', C, C, T, K, 5, 0, False))
SHAPES += [('ups0 512>256 k16 s8', 512, 256, 128, 16, 0, 8, False), ('ups1 256>128 k16 s8', 256, 128, 1024, 16, 0, 8, False),
           ('ups2 128>64 k4 s2', 128, 64, 8192, 4, 0, 2, False), ('ups3 64>32 k4 s2', 64, 32, 16384, 4, 0, 2, False)]


def main():
    precisions = sys.argv[1].split(',') if len(sys.argv) > 1 else ['fp32', 'tf32', 'tf32x3']
    rows = []
    for name, Cin, Cout, T, K, dil, u, with_res in SHAPES:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, Cin, T, generator=g).cuda()
        w = (torch.randn(*((Cin, Cout, K) if u else (Cout, Cin, K)), generator=g) * 0.02).cuda()
        b = torch.zeros(Cout).cuda()
        Tout = T * u if u else T
        res = torch.randn(B, Cout, Tout, generator=g).cuda() if with_res else None
        flops = 2.0 * B * T * Cin * Cout * K
        byts = 4.0 * B * (Cin * T + Cout * Tout * (2 if with_res else 1))
        row = {'layer': name, 'gflop': flops / 1e9, 'mbytes': byts / 1e6}
        for p in precisions:
            _, ms = run_layer(x, w, b, res, K, max(dil, 1), u, 0.1, 1.0, p, iters=10)
            row[p] = {'ms': ms, 'tflops': flops / ms / 1e9, 'gbs': byts / ms / 1e6}
        rows.append(row)
        print(name.ljust(24), ' '.join(f'{p}: {row[p]["ms"]*1e3:8.1f} us {row[p]["tflops"]:7.1f} TF/s {row[p]["gbs"]:7.0f} GB/s |'
                                       for p in precisions), flush=True)
    json.dump(rows, open('gpurun_out/layer_bench.json', 'w'), indent=1)


if __name__ == '__main__':
    import os
    os.makedirs('gpurun_out', exist_ok=True)
    main()
