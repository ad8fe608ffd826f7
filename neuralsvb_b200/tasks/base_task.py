"""Task base class: the hook contract the Trainer drives (reference: tasks/base_task.py:131-355)."""
import os
import random

import numpy as np
import torch
from torch import nn

from neuralsvb_b200.utils.ckpt_utils import load_ckpt
from neuralsvb_b200.utils.hparams import hparams
from neuralsvb_b200.utils.trainer import Trainer


class BaseTask(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.current_epoch = self.global_step = 0
        self.trainer, self.use_ddp, self.model, self.logger = None, False, None, None
        self.gradient_clip_norm = hparams.get('clip_grad_norm', 0)
        self.gradient_clip_val = hparams.get('clip_grad_value', 0)

    # ---- build model, dataloaders, optimizer, scheduler and tensorboard
    def build_model(self):
        raise NotImplementedError

    def train_dataloader(self):
        raise NotImplementedError

    def val_dataloader(self):
        raise NotImplementedError

    def test_dataloader(self):
        raise NotImplementedError

    def build_scheduler(self, optimizer):
        return None

    def build_optimizer(self, model):
        raise NotImplementedError

    def configure_optimizers(self):
        optm = self.build_optimizer(self.model)
        self.scheduler = self.build_scheduler(optm)
        return list(optm) if isinstance(optm, (list, tuple)) else [optm]

    def build_tensorboard(self, save_dir, name, version, **kwargs):
        log_dir = os.path.join(save_dir, name, f'version_{version}')
        os.makedirs(log_dir, exist_ok=True)
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.logger = SummaryWriter(log_dir=log_dir, **kwargs)
        except Exception:               # tensorboard is optional on the GPU box
            self.logger = None

    # ---- hooks
    def on_train_start(self):
        pass

    def on_train_end(self):
        pass

    def on_epoch_start(self):
        pass

    def on_epoch_end(self):
        pass

    def on_before_optimization(self, opt_idx):
        if self.gradient_clip_norm > 0:
            torch.nn.utils.clip_grad_norm_(self.parameters(), self.gradient_clip_norm)
        if self.gradient_clip_val > 0:
            torch.nn.utils.clip_grad_value_(self.parameters(), self.gradient_clip_val)

    def on_after_optimization(self, epoch, batch_idx, optimizer, optimizer_idx):
        if getattr(self, 'scheduler', None) is not None:
            self.scheduler.step(self.global_step // hparams.get('accumulate_grad_batches', 1))

    def training_step(self, sample, batch_idx, optimizer_idx=-1):
        raise NotImplementedError

    def validation_step(self, sample, batch_idx):
        raise NotImplementedError

    def validation_end(self, outputs):
        return {}

    def test_start(self):
        pass

    def test_step(self, sample, batch_idx):
        return self.validation_step(sample, batch_idx)

    def test_end(self, outputs):
        return self.validation_end(outputs)

    def on_keyboard_interrupt(self):
        pass

    def load_ckpt(self, ckpt_base_dir, current_model_name=None, model_name='model', force=True, strict=True):
        current_model_name = model_name if current_model_name is None else current_model_name
        load_ckpt(getattr(self, current_model_name), ckpt_base_dir, current_model_name, force, strict)

    # ---- start training / testing
    @classmethod
    def start(cls):
        os.environ.setdefault('MASTER_PORT', str(random.randint(15000, 30000)))
        random.seed(hparams['seed'])
        np.random.seed(hparams['seed'])
        trainer = Trainer(
            work_dir=hparams['work_dir'], val_check_interval=hparams['val_check_interval'],
            tb_log_interval=hparams['tb_log_interval'], max_updates=hparams['max_updates'],
            num_sanity_val_steps=hparams['num_sanity_val_steps'] if not hparams['validate'] else 10000,
            accumulate_grad_batches=hparams['accumulate_grad_batches'], print_nan_grads=hparams['print_nan_grads'],
            resume_from_checkpoint=hparams.get('resume_from_checkpoint', 0), amp=hparams['amp'],
            monitor_key=hparams['valid_monitor_key'], monitor_mode=hparams['valid_monitor_mode'],
            num_ckpt_keep=hparams['num_ckpt_keep'], save_best=hparams['save_best'], seed=hparams['seed'],
            debug=hparams['debug'])
        trainer.test(cls) if hparams['infer'] else trainer.fit(cls)
