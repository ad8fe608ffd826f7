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
    """(rank, world_size, local_rank): the initialised process group wins (the trainer's mp.spawn path sets no
    torchrun variables before it calls init_process_group), else torchrun-style env vars, else (0, 1, 0)."""
    local = int(os.environ.get('LOCAL_RANK', 0))
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size(), local
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), local


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
    """Gradients of ``params`` live as views into one flat buffer; one all-reduce(mean) per SEGMENT of it.

    ``segments`` (lists of parameters; default: one segment = one collective per optimizer step, the contract of
    DESIGN.md section 6) lets the exchange of a module whose backward finishes early (the MSD of the vocoder D step)
    overlap the rest of backward: after ``arm()`` a post-accumulate-grad hook counts the gradients of each segment and
    launches that segment's all-reduce asynchronously (on the backend's own stream) as soon as the last one landed;
    ``reduce()`` launches whatever has not been launched, waits, and turns sums into means.  Unarmed (gradient
    accumulation micro-batches, or anything unusual) it is the plain blocking exchange.  ``find_unused_parameters``
    semantics come for free: a parameter that received no gradient contributes zeros."""

    def __init__(self, params, segments=None):
        self.params = list(params)
        self.flat, self.bounds, self.works, self.launched = None, [], [], []
        self.armed, self.collectives = False, 0
        if not self.params:
            return
        segs = [list(s) for s in segments] if segments else [self.params]
        order = [p for s in segs for p in s]
        assert len(order) == len(self.params) and {id(p) for p in order} == {id(p) for p in self.params}, \
            'segments must partition the optimizer\'s parameters'
        self.params = order
        ref = self.params[0]
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=ref.dtype, device=ref.device)
        off = 0
        self.seg_of, self.pending, self.seg_params = {}, [], segs
        for si, seg in enumerate(segs):
            a = off
            for p in seg:
                view = self.flat[off:off + p.numel()].view_as(p)
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view
                off += p.numel()
                self.seg_of[id(p)] = si
                if hasattr(p, 'register_post_accumulate_grad_hook'):
                    p.register_post_accumulate_grad_hook(self._on_grad)
            self.bounds.append((a, off))
        self._reset()

    def _reset(self):
        self.pending = [sum(1 for p in seg if p.requires_grad) for seg in self.seg_params]
        self.launched = [False] * len(self.bounds)
        self.works = []

    def arm(self):
        """Call right before the backward whose gradients will be exchanged (the last micro-batch)."""
        if self.flat is None or not _active():
            return
        self.rebind()
        self._reset()
        self.armed = True

    def _on_grad(self, p):
        if not self.armed:
            return
        si = self.seg_of[id(p)]
        self.pending[si] -= 1
        if self.pending[si] == 0 and not self.launched[si]:
            a, b = self.bounds[si]
            base = self.flat.data_ptr()
            off = a
            for q in self.seg_params[si]:                # every gradient must (still) be a view of the flat buffer
                if q.grad is None or q.grad.data_ptr() != base + off * self.flat.element_size():
                    return                               # reduce() re-binds and sends this segment itself
                off += q.numel()
            self._launch(si)

    def _launch(self, si):
        a, b = self.bounds[si]
        self.works.append(dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, async_op=True))
        self.launched[si] = True
        self.collectives += 1

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
        """Finish the exchange of this optimizer step: mean over ranks, in place."""
        if self.flat is None or not _active():
            self.armed = False
            return
        world = dist.get_world_size()
        if not all(self.launched):
            self.rebind()
        for si in range(len(self.bounds)):
            if not self.launched[si]:
                self._launch(si)
        for w in self.works:
            w.wait()
        self.flat.div_(world)
        self.armed = False
        self._reset()


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
