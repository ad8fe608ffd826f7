"""Data-parallel plumbing: one process per GPU, ONE flat-buffer all-reduce per optimizer step.

The reference wraps the task in a DistributedDataParallel subclass that copies torch-1.11 private
internals (utils/ddp_utils.py:8-137) and reduces gradients in 25 MB buckets.  The payloads on this
path are small (55.8 MB generator, 283 MB discriminators, 40 MB SVB model), NVSwitch
gives every pair full bandwidth, and the trainer already knows exactly which parameters belong to
the optimizer that is about to step -- so the B200 design is: grads of that optimizer's parameters
are views into one contiguous buffer, and a single NCCL all-reduce (mean) is issued between
backward and ``optimizer.step()``.  ``find_unused_parameters`` semantics come for free: a parameter
that received no gradient contributes zeros.
"""
import os

import torch
import torch.distributed as dist


def dist_env():
    """(rank, world_size, local_rank) from torchrun-style env vars, or (0, 1, 0)."""
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0)))


def shard(items, rank, world_size, drop_uneven=False):
    """Rank r takes items[r::world] (the reference's batch sharding, tasks/tts/tts.py:69-72,93-96).
    With ``drop_uneven`` the tail that does not divide evenly is dropped, as the reference does for
    training batches."""
    items = list(items)
    if drop_uneven:
        items = items[:len(items) // world_size * world_size]
    return items[rank::world_size]


def broadcast_module(module, src=0):
    """Make every rank start from rank ``src``'s parameters and buffers (done once, not per forward)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


class FlatGradReducer:
    """Gradients of ``params`` live as views into one flat buffer; ``reduce()`` = one all-reduce(mean)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad or p.grad is not None or True]
        n = sum(p.numel() for p in self.params)
        if not self.params:
            self.flat = None
            return
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            off += p.numel()

    def rebind(self):
        """Re-attach views after something (e.g. zero_grad(set_to_none=True)) dropped them."""
        off = 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                view.zero_()
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view
            off += p.numel()

    def zero(self):
        if self.flat is not None:
            self.flat.zero_()

    def reduce(self):
        """One collective for the whole optimizer: mean over ranks, in place."""
        if self.flat is None or not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size()
        if world == 1:
            return
        self.rebind()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.div_(world)
