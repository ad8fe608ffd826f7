"""-m gpu: single conv layers through the C ABI (svb_conv1d_run) -- the CUDA-core kernel and the
tcgen05 kernel (1xTF32, 3xTF32) against a float64 torch convolution of the same layer."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from neuralsvb_b200 import _native

pytestmark = pytest.mark.gpu

# (B, Cin, Cout, T, K, dil, transposed stride u, residual)
LAYERS = [
    (2, 32, 32, 300, 3, 1, 0, True),
    (2, 32, 32, 700, 11, 5, 0, True),
    (1, 64, 64, 513, 7, 3, 0, False),
    (2, 128, 128, 384, 11, 1, 0, True),
    (2, 256, 256, 256, 3, 3, 0, True),
    (1, 256, 256, 1024, 11, 5, 0, True),
    (2, 64, 32, 200, 4, 0, 2, False),
    (1, 512, 256, 96, 16, 0, 8, False),
    (1, 80, 512, 60, 7, 1, 0, False),          # conv_pre shape: Cin padded to 96 on the tensor-core path
]
TOL = {'fp32': 2e-5, 'tf32x3': 3e-5, 'bf16x3': 1e-4, 'tf32': 3e-3}   # relative to max |y|


def run_layer(x, w, b, res, K, dil, u, slope, scale, precision, iters=1):
    lib = _native.lib()
    B, Cin, T = x.shape
    Cout = w.shape[1] if u else w.shape[0]
    Tout = T * u if u else T
    y = torch.empty(B, Cout, Tout, device='cuda')
    ms = ctypes.c_float(0)
    wc, bc = w.cpu().contiguous(), b.cpu().contiguous()
    rc = lib.svb_conv1d_run(_native.ptr(x), _native.ptr(wc), _native.ptr(bc), _native.ptr(res), B, Cin, Cout, T, K,
                            dil, u, slope, scale, _native.PREC[precision], iters, _native.ptr(y), ctypes.byref(ms),
                            _native.current_stream_ptr())
    _native.check(rc, 'conv1d_run')
    return y, ms.value


@pytest.mark.parametrize('precision', ['fp32', 'tf32', 'tf32x3', 'bf16x3'])
@pytest.mark.parametrize('layer', LAYERS)
def test_conv_layer_matches_float64_torch(layer, precision):
    B, Cin, Cout, T, K, dil, u, with_res = layer
    g = torch.Generator().manual_seed(1234 + Cin + K)
    x = torch.randn(B, Cin, T, generator=g).cuda() * 2
    wshape = (Cin, Cout, K) if u else (Cout, Cin, K)
    w = (torch.randn(*wshape, generator=g) / np.sqrt(Cin * K / max(u, 1))).cuda()
    b = torch.randn(Cout, generator=g).cuda() * 0.1
    Tout = T * u if u else T
    res = torch.randn(B, Cout, Tout, generator=g).cuda() if with_res else None
    slope, scale = 0.1, 1.0 / 3
    y, _ = run_layer(x, w, b, res, K, dil, u, slope, scale, precision)
    xd = F.leaky_relu(x.double(), slope)
    if u:
        ref = F.conv_transpose1d(xd, w.double(), b.double(), stride=u, padding=(K - u) // 2)
    else:
        ref = F.conv1d(xd, w.double(), b.double(), dilation=dil, padding=dil * (K - 1) // 2)
    if res is not None:
        ref = ref + res.double()
    ref = ref * scale
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    assert err < TOL[precision], f'{layer} {precision}: rel L-inf {err:.3e}'
