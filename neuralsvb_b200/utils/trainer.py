"""Training / evaluation driver with the reference's task contract (utils/trainer.py:23-520).

Same constructor arguments, same hooks called on the task, same checkpoint layout and resume rule,
written for torch 2.x / numpy 2 and for the B200 data-parallel design: one process per
GPU (``mp.spawn`` keyed on CUDA_VISIBLE_DEVICES like the reference, or torchrun env vars), gradients
exchanged with ONE flat-buffer all-reduce per optimizer step (utils/ddp_utils.FlatGradReducer)
instead of a DDP wrapper with bucketed reductions.

Hooks called on the task: build_model, configure_optimizers (list, entries may be None),
training_step(batch, batch_idx, optimizer_idx) -> {'loss', 'progress_bar', 'tb_log'},
validation_step / validation_end, test_start ('EXIT' skips) / test_step / test_end,
on_train_start / on_epoch_start / on_epoch_end / on_train_end, on_before_optimization(opt_idx),
on_after_optimization(epoch, batch_idx, optimizer, opt_idx), train/val/test_dataloader,
build_tensorboard, on_keyboard_interrupt.  Attributes set on the task: trainer, global_step,
current_epoch, testing, logger.
"""
import copy
import logging
import math
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuralsvb_b200.utils import move_to_cuda
from neuralsvb_b200.utils.ckpt_utils import get_all_ckpts, get_last_checkpoint
from neuralsvb_b200.utils.ddp_utils import FlatGradReducer, broadcast_module, dist_env
from neuralsvb_b200.utils.hparams import hparams


def _scalar(v):
    return v.item() if isinstance(v, torch.Tensor) else v


class Trainer:
    def __init__(self, work_dir, default_save_path=None, accumulate_grad_batches=1, max_updates=160000,
                 print_nan_grads=False, val_check_interval=2000, num_sanity_val_steps=5, amp=False,
                 log_save_interval=100, tb_log_interval=10, monitor_key='val_loss', monitor_mode='min',
                 num_ckpt_keep=5, save_best=True, resume_from_checkpoint=0, seed=1234, debug=False,
                 dist_backend=None):
        os.makedirs(work_dir, exist_ok=True)
        self.work_dir = work_dir
        self.accumulate_grad_batches = accumulate_grad_batches
        self.max_updates = max_updates
        self.num_sanity_val_steps = num_sanity_val_steps
        self.print_nan_grads = print_nan_grads
        self.default_save_path = default_save_path
        self.resume_from_checkpoint = resume_from_checkpoint if resume_from_checkpoint > 0 else None
        self.seed, self.debug = seed, debug
        self.task, self.optimizers, self.reducers = None, [], []
        self.testing = False
        self.global_step = self.current_epoch = 0
        self.monitor_key, self.monitor_mode = monitor_key, monitor_mode
        self.num_ckpt_keep, self.save_best = num_ckpt_keep, save_best
        self.best_val_results = math.inf if monitor_mode == 'min' else -math.inf
        self.log_save_interval, self.val_check_interval, self.tb_log_interval = log_save_interval, val_check_interval, tb_log_interval
        self.amp = amp
        self.amp_scalar = torch.amp.GradScaler('cuda', enabled=amp and torch.cuda.is_available())
        self.dist_backend = dist_backend
        # devices: the reference keys everything on CUDA_VISIBLE_DEVICES (utils/trainer.py:75-81)
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis is None:
            self.all_gpu_ids = list(range(torch.cuda.device_count())) if torch.cuda.is_available() else []
        else:
            self.all_gpu_ids = [int(x) for x in vis.split(',') if x != ''] if torch.cuda.is_available() else []
        self.num_gpus = len(self.all_gpu_ids)
        self.on_gpu = self.num_gpus > 0
        self.root_gpu = 0
        self.proc_rank, self.world_size = 0, 1
        self.use_ddp = False
        self.first_epoch = True
        logging.info(f'GPU available: {torch.cuda.is_available()}, GPU used: {self.all_gpu_ids}')

    # ------------------------------------------------------------------ entry points
    def test(self, task_cls):
        self.testing = True
        return self.fit(task_cls)

    def fit(self, task_cls):
        rank, world, local_rank = dist_env()
        if world > 1:                                   # launched by torchrun: this process is one rank
            self._ddp_worker(local_rank, task_cls, None, rank, world)
        elif self.num_gpus > 1:                         # reference behaviour: spawn one process per visible GPU
            mp.spawn(self._ddp_worker, nprocs=self.num_gpus, args=(task_cls, copy.deepcopy(dict(hparams)), None, self.num_gpus))
        else:
            self.task = task_cls()
            self.task.trainer = self
            self.run_single_process(self.task)
        return 1

    def _ddp_worker(self, local_idx, task_cls, hparams_, rank=None, world=None):
        if hparams_ is not None:
            hparams.update(hparams_)
        self.proc_rank = local_idx if rank is None else rank
        self.world_size = world
        self.root_gpu = local_idx
        self.use_ddp = True
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29541')
        # the mp.spawn path (the reference's default, utils/trainer.py:94-107) starts without torchrun's variables:
        # export them so everything that shards by rank (task dataloaders, ddp_utils.dist_env) sees this worker's rank
        os.environ['RANK'], os.environ['WORLD_SIZE'] = str(self.proc_rank), str(world)
        os.environ['LOCAL_RANK'] = str(local_idx)
        backend = self.dist_backend or ('nccl' if self.on_gpu else 'gloo')
        if not dist.is_initialized():
            dist.init_process_group(backend, rank=self.proc_rank, world_size=world)
        if self.on_gpu:
            torch.cuda.set_device(local_idx)
        task = task_cls()
        task.trainer = self
        self.task = task
        if self.proc_rank != 0 and not self.debug:
            sys.stdout = open(os.devnull, 'w')
        random.seed(self.seed)
        np.random.seed(self.seed)
        self.run_single_process(task)

    def run_single_process(self, task):
        model = task.build_model()
        if model is not None:
            task.model = model
        checkpoint, _ = get_last_checkpoint(self.work_dir, self.resume_from_checkpoint)
        if checkpoint is not None:
            self.restore_weights(checkpoint)
        if self.on_gpu:
            task.cuda(self.root_gpu)
        if self.use_ddp:
            broadcast_module(task)                      # once; buffers are not re-broadcast every forward
        if not self.testing:
            self.optimizers = task.configure_optimizers()
            # gradient exchange: one flat buffer per optimizer; a task may split it into segments whose all-reduce is
            # launched as soon as their last gradient lands (``grad_segments(opt_idx)`` -> lists of parameters)
            seg_fn = getattr(task, 'grad_segments', lambda i: None)
            self.reducers = [FlatGradReducer([p for g in o.param_groups for p in g['params']], seg_fn(i))
                             if (o is not None and self.use_ddp) else None for i, o in enumerate(self.optimizers)]
            self.first_epoch = True
        if checkpoint is not None:
            self.restore_opt_state(checkpoint)
        del checkpoint
        if self.use_ddp:
            dist.barrier()
        task.trainer, task.testing = self, self.testing
        task.use_ddp = self.use_ddp
        if self.proc_rank == 0:
            task.build_tensorboard(save_dir=self.work_dir, name='lightning_logs', version='lastest')
        else:
            os.makedirs('tmp', exist_ok=True)
            task.build_tensorboard(save_dir='tmp', name='tb_tmp', version='lastest')
        self.logger = getattr(task, 'logger', None)
        try:
            if self.testing:
                self.run_evaluation(test=True)
            else:
                self.train()
        except KeyboardInterrupt:
            task.on_keyboard_interrupt()

    def get_task_ref(self):
        return self.task

    # ------------------------------------------------------------------ evaluation
    def run_evaluation(self, test=False):
        results = self.evaluate(self.task, test, tqdm_desc='test' if test else 'Valid')
        if results is not None and 'tb_log' in results:
            self.log_metrics_to_tb(results['tb_log'])
        if self.proc_rank == 0 and not test:
            self.save_checkpoint(epoch=self.current_epoch, logs=results)

    def evaluate(self, task, test=False, tqdm_desc='Valid', max_batches=None):
        max_batches = None if max_batches == -1 else max_batches
        task.zero_grad()
        task.eval()
        outputs = []
        with torch.no_grad():
            if test and task.test_start() == 'EXIT':
                task.train()
                return None
            loader = task.test_dataloader() if test else task.val_dataloader()
            for batch_idx, batch in enumerate(loader):
                if batch is None:
                    continue
                if max_batches is not None and batch_idx >= max_batches:
                    break
                if self.on_gpu:
                    batch = move_to_cuda(batch, self.root_gpu)
                outputs.append(task.test_step(batch, batch_idx) if test else task.validation_step(batch, batch_idx))
            results = task.test_end(outputs) if test else task.validation_end(outputs)
        task.train()
        return results

    # ------------------------------------------------------------------ training
    def train(self):
        task = self.task
        task.on_train_start()
        if self.num_sanity_val_steps > 0:
            self.evaluate(task, False, 'Sanity Val', max_batches=self.num_sanity_val_steps)
        loader = task.train_dataloader()
        epoch = self.current_epoch
        done = False
        while not done:
            if self.use_ddp and hasattr(getattr(loader, 'sampler', None), 'set_epoch'):
                loader.sampler.set_epoch(epoch)
            task.current_epoch = self.current_epoch = epoch
            task.on_epoch_start()
            n_batches = 0
            for batch_idx, batch in enumerate(loader):
                n_batches += 1
                pbar_metrics, tb_metrics = self.run_training_batch(batch_idx, batch)
                if self.global_step % self.val_check_interval == 0 and not self.first_epoch:
                    self.run_evaluation()
                self.first_epoch = False
                if (self.global_step + 1) % self.tb_log_interval == 0:
                    self.log_metrics_to_tb(tb_metrics)
                    if self.proc_rank == 0 and pbar_metrics:
                        print(f'| step {self.global_step}: ' + ', '.join(f'{k} {float(_scalar(v)):.4f}' for k, v in pbar_metrics.items()), flush=True)
                self.global_step += 1
                task.global_step = self.global_step
                if self.global_step > self.max_updates:
                    print('| Training end..')
                    done = True
                    break
            task.on_epoch_end()
            epoch += 1
            if n_batches == 0:
                break
        if self.proc_rank == 0 and not get_all_ckpts(self.work_dir):
            # the reference saves only after a validation (utils/trainer.py:155-164): a run shorter than
            # val_check_interval would end with nothing for the vocoder plugin to load
            self.save_checkpoint(epoch=self.current_epoch)
        task.on_train_end()

    def run_training_batch(self, batch_idx, batch):
        if batch is None:
            return {}, {}
        task = self.task
        pbar_all, log_all = {}, {}
        for opt_idx, optimizer in enumerate(self.optimizers):
            if optimizer is None:
                continue
            if len(self.optimizers) > 1:      # only this optimizer's parameters collect gradients
                mine = {id(p) for g in optimizer.param_groups for p in g['params']}
                for p in task.parameters():
                    p.requires_grad = id(p) in mine
            with torch.autocast('cuda', enabled=self.amp and self.on_gpu):
                b = move_to_cuda(copy.copy(batch), self.root_gpu) if self.on_gpu else batch
                out = task.training_step(b, batch_idx, opt_idx)
                loss = out['loss']
                if loss is None:
                    continue
                pbar_all.update(out.get('progress_bar', {}))
                log_all.update(out.get('tb_log', {}))
                loss = loss / self.accumulate_grad_batches
            if loss.requires_grad:
                if self.reducers and self.reducers[opt_idx] is not None:
                    if (self.global_step + 1) % self.accumulate_grad_batches == 0:
                        self.reducers[opt_idx].arm()                # exchange overlaps the rest of this backward
                    else:
                        self.reducers[opt_idx].rebind()
                self.amp_scalar.scale(loss).backward() if self.amp else loss.backward()
            if self.print_nan_grads:
                bad = [n for n, p in task.named_parameters() if p.grad is not None and torch.isnan(p.grad.float()).any()]
                if bad:
                    print('| NaN grads: ', bad)
                    sys.exit(0)
            if (self.global_step + 1) % self.accumulate_grad_batches == 0:
                if self.reducers and self.reducers[opt_idx] is not None:
                    self.reducers[opt_idx].reduce()                 # finish this optimizer's exchange (one collective per segment)
                task.on_before_optimization(opt_idx)
                if self.amp:
                    self.amp_scalar.step(optimizer)
                    self.amp_scalar.update()
                else:
                    optimizer.step()
                if self.reducers and self.reducers[opt_idx] is not None:
                    self.reducers[opt_idx].zero()
                else:
                    optimizer.zero_grad()
                task.on_after_optimization(self.current_epoch, batch_idx, optimizer, opt_idx)
        return pbar_all, log_all

    # ------------------------------------------------------------------ checkpoints
    def restore_weights(self, checkpoint):
        task = self.task
        sd = checkpoint['state_dict']
        if any('.' in k for k in sd):
            task.load_state_dict(sd)
        else:
            for name, child_sd in sd.items():
                getattr(task, name).load_state_dict(child_sd)
        self.best_val_results = checkpoint['checkpoint_callback_best']
        self.global_step = task.global_step = checkpoint['global_step']
        self.current_epoch = checkpoint['epoch']

    def restore_opt_state(self, checkpoint):
        if self.testing:
            return
        for optimizer, state in zip([o for o in self.optimizers if o is not None], checkpoint['optimizer_states']):
            try:
                optimizer.load_state_dict(state)
            except ValueError:
                print('| WARMING: optimizer parameters not match !!!')

    def dump_checkpoint(self):
        return {
            'epoch': self.current_epoch, 'global_step': self.global_step, 'checkpoint_callback_best': self.best_val_results,
            'optimizer_states': [o.state_dict() for o in self.optimizers if o is not None],
            'state_dict': {k: v.state_dict() for k, v in self.task.named_children() if len(list(v.parameters())) > 0},
        }

    def _atomic_save(self, filepath):
        tmp = str(filepath) + '.part'
        torch.save(self.dump_checkpoint(), tmp, _use_new_zipfile_serialization=False)
        os.replace(tmp, filepath)

    def save_checkpoint(self, epoch, logs=None):
        path = f'{self.work_dir}/model_ckpt_steps_{self.global_step}.ckpt'
        logging.info(f'Epoch {epoch:05d}@{self.global_step}: saving model to {path}')
        self._atomic_save(path)
        for old in get_all_ckpts(self.work_dir)[self.num_ckpt_keep:]:
            os.remove(old)
            logging.info(f'Delete ckpt: {os.path.basename(old)}')
        current = None if logs is None else logs.get(self.monitor_key)
        if current is not None and self.save_best:
            better = current < self.best_val_results if self.monitor_mode == 'min' else current > self.best_val_results
            if better:
                self.best_val_results = current
                self._atomic_save(f'{self.work_dir}/model_ckpt_best.pt')

    # ------------------------------------------------------------------ logging
    def log_metrics_to_tb(self, metrics, step=None):
        if not metrics or self.logger is None:
            return
        step = self.global_step if step is None else step
        for k, v in metrics.items():
            if k in ('epoch', 'step'):
                continue
            try:
                self.logger.add_scalar(k, _scalar(v), step)
            except Exception:       # non-scalar entries are the task's business
                pass
