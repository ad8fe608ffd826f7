import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator
from neuralsvb_b200.utils import synthetic as S
from oracle import hifigan as O
h = S.small_config(True); B, T, hop = 2, 24, 16
sd = S.make_generator_state_dict(h, 1234)
mel, f0 = S.make_mel_f0(B, T, 1234); ri, nz = S.make_nsf_noise(B, T * hop, 1234)
cot = torch.randn(B, 1, T * hop, generator=torch.Generator().manual_seed(7))
p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
taps = {}
y_ref = O.generator_forward(O.fold_weight_norm(p), h, mel, f0, ri, nz, taps=taps)
(y_ref * cot).sum().backward()
for prec in ('fp32', 'bf16x3', 'tf32x3'):
    for mode in ('eval', 'train'):
        m = HifiGanGenerator(h, precision=prec); m.load_state_dict(sd, strict=True); m = m.to('cuda:0')
        m.train() if mode == 'train' else m.eval()
        if mode == 'eval':
            with torch.no_grad():
                y = m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda())
        else:
            y = m(mel.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda())
        out = []
        for name in ('conv_pre', 'ups0', 'stage0', 'ups1', 'stage1', 'stage2'):
            a = m.get_tap(name).cpu().double(); r = taps[name].detach().double()
            out.append(f'{name} {float((a - r).norm() / r.norm()):.1e}')
        line = f'{prec} {mode}: y {float((y.detach().cpu().double() - y_ref.detach().double()).norm() / y_ref.detach().double().norm()):.1e} | ' + ' '.join(out)
        if mode == 'train':
            (y * cot.cuda()).sum().backward()
            errs = [float((q.grad.cpu().double() - p[k].grad.double()).norm() / p[k].grad.double().norm()) for k, q in m.named_parameters()]
            line += f' | grad median {np.median(errs):.1e} max {max(errs):.1e}'
        print(line)
