#!/usr/bin/env python
"""Tuning sweep for the tcgen05 conv kernel: cluster size of the weight multicast (SVB_TC_CL) x resident-CTA target
(SVB_TC_CTAS) per generator layer shape at BASELINE config 2.  Prints one row per layer (us)."""
import os
import sys

import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from tests.test_gpu_conv_kernels import run_layer  # noqa: E402
from tools.layer_bench import B, SHAPES  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
    combos = [(cl, c) for cl in (1, 2, 4, 8) for c in (5, 2)]
    print('layer'.ljust(24) + ''.join(f'CL{cl}/c{c}'.rjust(9) for cl, c in combos))
    for name, Cin, Cout, T, K, dil, u, with_res in SHAPES:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, Cin, T, generator=g).cuda()
        w = (torch.randn(*((Cin, Cout, K) if u else (Cout, Cin, K)), generator=g) * 0.02).cuda()
        b = torch.zeros(Cout).cuda()
        res = torch.randn(B, Cout, T * u if u else T, generator=g).cuda() if with_res else None
        row = name.ljust(24)
        for mt, c in combos:
            os.environ['SVB_TC_CL'], os.environ['SVB_TC_CTAS'] = str(mt), str(c)
            _, ms = run_layer(x, w, b, res, K, max(dil, 1), u, 0.1, 1.0, prec, iters=10)
            row += f'{ms * 1e3:9.1f}'
        print(row, flush=True)


if __name__ == '__main__':
    main()
