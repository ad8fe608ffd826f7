"""CPU tests of the entry points that stay around the hot path: Trainer (multi-optimizer loop,
checkpoint layout / resume), tasks.run dispatch, and the N > 1 data-parallel path on gloo with
world_size 2 (clip sharding + ONE flat-buffer all-reduce per optimizer step)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from neuralsvb_b200.tasks.base_task import BaseTask
from neuralsvb_b200.utils import ddp_utils
from neuralsvb_b200.utils.ckpt_utils import get_all_ckpts, get_last_checkpoint, load_ckpt
from neuralsvb_b200.utils.hparams import hparams
from neuralsvb_b200.utils.trainer import Trainer

HP = dict(clip_grad_norm=0, seed=1234, accumulate_grad_batches=1)


class ToyGanTask(BaseTask):
    """Two optimizers like the vocoder / SVB tasks (generator step, discriminator step)."""
    calls = []

    def build_model(self):
        torch.manual_seed(7)
        self.model_gen = nn.Linear(4, 4)
        self.model_disc = nn.Linear(4, 1)
        return None

    def configure_optimizers(self):
        return [torch.optim.SGD(self.model_gen.parameters(), lr=0.1), torch.optim.SGD(self.model_disc.parameters(), lr=0.1)]

    def _data(self):
        g = torch.Generator().manual_seed(3)
        return [torch.randn(8, 4, generator=g) for _ in range(6)]

    def train_dataloader(self):
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        return [torch.stack(ddp_utils.shard(list(b), rank, world)) for b in self._data()]

    def val_dataloader(self):
        return self._data()[:2]

    def test_dataloader(self):
        return self._data()[:1]

    def training_step(self, batch, batch_idx, optimizer_idx=-1):
        ToyGanTask.calls.append(('train', batch_idx, optimizer_idx, self.model_gen.weight.requires_grad,
                                 next(self.model_disc.parameters()).requires_grad))
        y = self.model_gen(batch)
        if optimizer_idx == 0:
            loss = (1 - self.model_disc(y)).pow(2).mean()
        else:
            loss = self.model_disc(y.detach()).pow(2).mean() + (1 - self.model_disc(batch)).pow(2).mean()
        return {'loss': loss, 'progress_bar': {f'l{optimizer_idx}': loss.item()}, 'tb_log': {f'l{optimizer_idx}': loss.item()}}

    def validation_step(self, batch, batch_idx):
        return {'val_loss': self.model_gen(batch).pow(2).mean().item()}

    def validation_end(self, outputs):
        v = sum(o['val_loss'] for o in outputs) / len(outputs)
        return {'val_loss': v, 'tb_log': {'val_loss': v}}


@pytest.fixture(autouse=True)
def _hp():
    hparams.clear()
    hparams.update(HP)
    ToyGanTask.calls = []
    yield
    hparams.clear()


def make_trainer(work_dir, **kw):
    return Trainer(work_dir=str(work_dir), val_check_interval=2, tb_log_interval=1, max_updates=5, num_sanity_val_steps=1,
                   num_ckpt_keep=2, monitor_key='val_loss', **kw)


def test_multi_optimizer_loop_checkpoint_layout_and_resume(tmp_path, monkeypatch):
    monkeypatch.setenv('CUDA_VISIBLE_DEVICES', '')
    t = make_trainer(tmp_path)
    t.fit(ToyGanTask)
    steps = [c for c in ToyGanTask.calls if c[0] == 'train']
    # every batch visits optimizer 0 then 1, and only that optimizer's parameters require grad
    assert steps[0][1:] == (0, 0, True, False) and steps[1][1:] == (0, 1, False, True)
    assert t.global_step == 6
    ckpts = get_all_ckpts(str(tmp_path))
    assert [os.path.basename(c) for c in ckpts] == ['model_ckpt_steps_4.ckpt', 'model_ckpt_steps_2.ckpt']   # keep 2, newest first
    assert os.path.exists(tmp_path / 'model_ckpt_best.pt')
    ck, path = get_last_checkpoint(str(tmp_path))
    assert set(ck) == {'epoch', 'global_step', 'checkpoint_callback_best', 'optimizer_states', 'state_dict'}
    assert set(ck['state_dict']) == {'model_gen', 'model_disc'} and 'weight' in ck['state_dict']['model_gen']
    assert len(ck['optimizer_states']) == 2 and ck['global_step'] == 4
    # resume: a new trainer on the same work_dir continues from the newest checkpoint
    t2 = make_trainer(tmp_path)
    t2.max_updates = 7
    t2.fit(ToyGanTask)
    assert t2.global_step == 8 and t2.task.model_gen.weight.requires_grad in (True, False)
    # partial load of one child (vocoders/hifigan.py-style consumers use the same layout)
    lin = nn.Linear(4, 4)
    load_ckpt(lin, str(tmp_path), 'model_gen')
    assert torch.equal(lin.weight, torch.load(get_all_ckpts(str(tmp_path))[0], weights_only=False)['state_dict']['model_gen']['weight'])


def test_test_mode_calls_test_hooks(tmp_path, monkeypatch):
    monkeypatch.setenv('CUDA_VISIBLE_DEVICES', '')
    seen = []

    class T(ToyGanTask):
        def test_start(self):
            seen.append('start')

        def test_step(self, batch, batch_idx):
            seen.append('step')
            return {}

        def test_end(self, outputs):
            seen.append('end')
            return {}
    make_trainer(tmp_path).test(T)
    assert seen == ['start', 'step', 'end']


def test_run_task_dispatches_to_task_cls(monkeypatch):
    from neuralsvb_b200.tasks import run
    started = []

    class Fake:
        @classmethod
        def start(cls):
            started.append(cls)
    mod = type(sys)('fake_task_mod')
    mod.Fake = Fake
    monkeypatch.setitem(sys.modules, 'fake_task_mod', mod)
    hparams['task_cls'] = 'fake_task_mod.Fake'
    run.run_task()
    assert started == [Fake]


def test_shard_matches_reference_batch_sharding():
    items = list(range(10))
    assert ddp_utils.shard(items, 0, 4) == [0, 4, 8] and ddp_utils.shard(items, 3, 4) == [3, 7]
    assert ddp_utils.shard(items, 1, 4, drop_uneven=True) == [1, 5]
    got = sorted(sum((ddp_utils.shard(items, r, 3) for r in range(3)), []))
    assert got == items                                     # every unit exactly once, no overlap


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ddp_worker(rank, world, port, work_dir, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES='')
    hparams.clear()
    hparams.update(HP)
    n_reduce = []
    real = dist.all_reduce

    def counting(t, *a, **k):
        n_reduce.append(t.numel())
        return real(t, *a, **k)
    dist.all_reduce = counting
    t = Trainer(work_dir=os.path.join(work_dir, f'r{rank}'), val_check_interval=100, tb_log_interval=1000, max_updates=3,
                num_sanity_val_steps=0, dist_backend='gloo', debug=True)
    t.fit(ToyGanTask)
    torch.save({'gen': t.task.model_gen.state_dict(), 'disc': t.task.model_disc.state_dict(), 'n_reduce': n_reduce,
                'steps': t.global_step}, os.path.join(out, f'rank{rank}.pt'))
    dist.destroy_process_group()


def test_ddp_world_size_2_gloo_one_allreduce_per_optimizer_step(tmp_path, monkeypatch):
    world = 2
    mp.spawn(_ddp_worker, nprocs=world, args=(world, _free_port(), str(tmp_path), str(tmp_path)), join=True)
    r0, r1 = (torch.load(tmp_path / f'rank{r}.pt', weights_only=False) for r in range(world))
    # replicas stay identical
    for k in ('gen', 'disc'):
        for name in r0[k]:
            assert torch.equal(r0[k][name], r1[k][name])
    # exactly one collective per (batch, optimizer): 4 steps x 2 optimizers, payload = that optimizer's parameters
    assert r0['n_reduce'] == [20, 5] * r0['steps'] and r0['steps'] == 4
    # and equal to single-process training on the un-sharded batches (mean of the shard gradients)
    monkeypatch.setenv('CUDA_VISIBLE_DEVICES', '')
    hparams.clear()
    hparams.update(HP)
    t = Trainer(work_dir=str(tmp_path / 'single'), val_check_interval=100, tb_log_interval=1000, max_updates=3,
                num_sanity_val_steps=0)
    t.fit(ToyGanTask)
    for name, v in t.task.model_gen.state_dict().items():
        assert torch.allclose(v, r0['gen'][name], atol=1e-6), name


class SegTask(ToyGanTask):
    """Discriminator optimizer exchanged in two segments (the vocoder task splits MSD / MPD the same way)."""

    def build_model(self):
        super().build_model()
        self.model_disc = nn.Sequential(nn.Linear(4, 3), nn.Linear(3, 1))
        return None

    def grad_segments(self, opt_idx):
        if opt_idx != 1:
            return None
        return [list(self.model_disc[1].parameters()), list(self.model_disc[0].parameters())]


def _seg_worker(rank, world, port, work_dir, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES='')
    hparams.clear()
    hparams.update(HP)
    t = Trainer(work_dir=os.path.join(work_dir, f'r{rank}'), val_check_interval=100, tb_log_interval=1000, max_updates=3,
                num_sanity_val_steps=0, dist_backend='gloo', debug=True)
    t.fit(SegTask)
    torch.save({'gen': t.task.model_gen.state_dict(), 'disc': t.task.model_disc.state_dict(), 'steps': t.global_step,
                'collectives': [r.collectives for r in t.reducers], 'bounds': [r.bounds for r in t.reducers]},
               os.path.join(out, f'rank{rank}.pt'))
    dist.destroy_process_group()


def test_ddp_segmented_exchange_matches_single_process(tmp_path, monkeypatch):
    """grad_segments: the all-reduce of a segment is launched from the autograd hook of its last gradient (overlapping
    the rest of backward); one collective per segment per optimizer step; same result as un-sharded training."""
    world = 2
    mp.spawn(_seg_worker, nprocs=world, args=(world, _free_port(), str(tmp_path), str(tmp_path)), join=True)
    r0, r1 = (torch.load(tmp_path / f'rank{r}.pt', weights_only=False) for r in range(world))
    for k in ('gen', 'disc'):
        for name in r0[k]:
            assert torch.equal(r0[k][name], r1[k][name])
    assert r0['collectives'] == [r0['steps'], 2 * r0['steps']] and r0['bounds'][1] == [(0, 4), (4, 19)]
    monkeypatch.setenv('CUDA_VISIBLE_DEVICES', '')
    hparams.clear()
    hparams.update(HP)
    t = Trainer(work_dir=str(tmp_path / 'single'), val_check_interval=100, tb_log_interval=1000, max_updates=3, num_sanity_val_steps=0)
    t.fit(SegTask)
    for k, mod in (('gen', t.task.model_gen), ('disc', t.task.model_disc)):
        for name, v in mod.state_dict().items():
            assert torch.allclose(v, r0[k][name], atol=1e-6), (k, name)


class RankProbeTask(ToyGanTask):
    def train_dataloader(self):
        rank, world, local = ddp_utils.dist_env()
        with open(os.path.join(hparams['probe_dir'], f'probe{dist.get_rank()}.txt'), 'w') as f:
            f.write(f'{rank} {world} {local} {os.environ.get("RANK")} {os.environ.get("WORLD_SIZE")}')
        return [torch.stack(ddp_utils.shard(list(b), rank, world)) for b in self._data()]


def _spawn_path_worker(local_idx, work_dir, port):
    # what Trainer.fit's mp.spawn branch runs (utils/trainer.py: reference behaviour keyed on CUDA_VISIBLE_DEVICES):
    # no torchrun variables in the environment
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        os.environ.pop(k, None)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES='')
    t = Trainer(work_dir=os.path.join(work_dir, f'r{local_idx}'), val_check_interval=100, tb_log_interval=1000, max_updates=1,
                num_sanity_val_steps=0, dist_backend='gloo', debug=True)
    t._ddp_worker(local_idx, RankProbeTask, dict(HP, probe_dir=work_dir), None, 2)
    dist.destroy_process_group()


def test_spawn_path_exports_rank_so_shards_are_disjoint(tmp_path):
    """on the mp.spawn path every worker saw dist_env() == (0, 1, 0) and trained on identical shards."""
    mp.spawn(_spawn_path_worker, nprocs=2, args=(str(tmp_path), _free_port()), join=True)
    got = [open(tmp_path / f'probe{r}.txt').read().split() for r in range(2)]
    assert got[0] == ['0', '2', '0', '0', '2'] and got[1] == ['1', '2', '1', '1', '2']
