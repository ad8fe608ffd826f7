"""Checkpoint discovery and partial loading (contract of the reference's utils/ckpt_utils.py:8-68).

On-disk layout (the contract the vocoder loader depends on, vocoders/hifigan.py:19-29):
``{work_dir}/model_ckpt_steps_{N}.ckpt`` = ``{'epoch', 'global_step', 'checkpoint_callback_best',
'optimizer_states': [...], 'state_dict': {child_name: child.state_dict()}}``.
"""
import glob
import logging
import os
import re

import torch

_STEP_RE = re.compile(r'steps_(\d+)\.ckpt$')


def get_all_ckpts(work_dir, steps=None):
    """Checkpoint paths, newest (largest step) first."""
    pattern = f'{work_dir}/model_ckpt_steps_{"*" if steps is None else steps}.ckpt'
    paths = [p for p in glob.glob(pattern) if _STEP_RE.search(p)]
    return sorted(paths, key=lambda p: int(_STEP_RE.search(p).group(1)), reverse=True)


def get_last_checkpoint(work_dir, steps=None):
    paths = get_all_ckpts(work_dir, steps)
    if not paths:
        return None, None
    ckpt = torch.load(paths[0], map_location='cpu', weights_only=False)
    logging.info(f'load module from checkpoint: {paths[0]}')
    return ckpt, paths[0]


def _select(state_dict, model_name):
    """The sub-dict for ``model_name``: either flat 'model.x.y' keys or a nested {child: state_dict};
    dotted names ('model.sub') descend into the child's keys."""
    if any('.' in k for k in state_dict):
        prefix = model_name + '.'
        return {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
    head, _, rest = model_name.partition('.')
    sub = state_dict[head]
    if not rest:
        return sub
    prefix = rest + '.'
    return {k[len(prefix):]: v for k, v in sub.items() if k.startswith(prefix)}


def load_ckpt(cur_model, ckpt_base_dir, model_name='model', force=True, strict=True):
    if os.path.isfile(ckpt_base_dir):
        base_dir, ckpt_path = os.path.dirname(ckpt_base_dir), ckpt_base_dir
        checkpoint = torch.load(ckpt_base_dir, map_location='cpu', weights_only=False)
    else:
        base_dir = ckpt_base_dir
        checkpoint, ckpt_path = get_last_checkpoint(ckpt_base_dir)
    if checkpoint is None:
        msg = f'| ckpt not found in {base_dir}.'
        assert not force, msg
        print(msg)
        return
    state = _select(checkpoint['state_dict'], model_name)
    if not strict:
        own = cur_model.state_dict()
        for key in [k for k, v in state.items() if k in own and own[k].shape != v.shape]:
            print('| Unmatched keys: ', key, own[key].shape, state[key].shape)
            del state[key]
    cur_model.load_state_dict(state, strict=strict)
    print(f"| load '{model_name}' from '{ckpt_path}'.")
