"""Helpers the hot path's callers use (reference: utils/__init__.py)."""
import time

import torch


class Timer:
    """Cumulative wall-clock per name with a device sync on both sides; prints when enabled
    (reference: utils/__init__.py:243-264; wraps the generator call in spec2wav,
    vocoders/hifigan.py:59, gated by hparams['profile_infer'])."""
    timer_map = {}

    def __init__(self, name, enable=False):
        Timer.timer_map.setdefault(name, 0)
        self.name, self.enable = name, enable

    def __enter__(self):
        if self.enable:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            self.t = time.time()

    def __exit__(self, exc_type, exc_val, exc_tb):
        if self.enable:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            Timer.timer_map[self.name] += time.time() - self.t
            print(f'[Timer] {self.name}: {Timer.timer_map[self.name]}')


class AvgrageMeter(object):
    """(reference spelling kept: utils/__init__.py:102-115)"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.avg, self.sum, self.cnt = 0, 0, 0

    def update(self, val, n=1):
        self.sum += val * n
        self.cnt += n
        self.avg = self.sum / self.cnt


def move_to_cuda(batch, gpu_id=0):
    """Recursively move tensors in lists / tuples / dicts (reference: utils/__init__.py:80-99)."""
    if callable(getattr(batch, 'cuda', None)):
        return batch.cuda(gpu_id, non_blocking=True)
    if callable(getattr(batch, 'to', None)):
        return batch.to(torch.device('cuda', gpu_id), non_blocking=True)
    if isinstance(batch, list):
        return [move_to_cuda(x, gpu_id) for x in batch]
    if isinstance(batch, tuple):
        return tuple(move_to_cuda(x, gpu_id) for x in batch)
    if isinstance(batch, dict):
        return {k: move_to_cuda(v, gpu_id) for k, v in batch.items()}
    return batch


def move_to_cpu(tensors):
    if isinstance(tensors, dict):
        return {k: move_to_cpu(v) for k, v in tensors.items()}
    if isinstance(tensors, (list, tuple)):
        return type(tensors)(move_to_cpu(v) for v in tensors)
    if isinstance(tensors, torch.Tensor):
        return tensors.cpu()
    return tensors


def tensors_to_scalars(tensors):
    """(reference: utils/__init__.py:24-43)"""
    if isinstance(tensors, torch.Tensor):
        return tensors.item()
    if isinstance(tensors, dict):
        return {k: tensors_to_scalars(v) for k, v in tensors.items()}
    if isinstance(tensors, list):
        return [tensors_to_scalars(v) for v in tensors]
    return tensors
