"""Config system with the reference's semantics (reference: utils/hparams.py:8-128).

A global mutable dict ``hparams``; YAML files chained through ``base_config`` (str or list,
paths starting with '.' are relative to the including file, depth-first, later overrides
earlier, nested dicts merged); ``--hparams "a=1,b.c=2,d=[1 2]"`` overrides typed by the
existing value; ``checkpoints/<exp_name>/config.yaml`` is saved on first run and merged back
unless ``--reset``.  CLI flags: --config --exp_name --hparams --infer --validate --reset
--remove --debug.
"""
import argparse
import ast
import os
import shutil

import yaml

hparams = {}
global_print_hparams = True


class Args:
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)


def override_config(old_config, new_config):
    """Recursive dict merge: nested dicts are merged, everything else replaced."""
    for key, val in new_config.items():
        if isinstance(val, dict) and isinstance(old_config.get(key), dict):
            override_config(old_config[key], val)
        else:
            old_config[key] = val


def _load_chain(config_fn, seen, chain):
    if not os.path.exists(config_fn):
        return {}
    with open(config_fn) as f:
        cur = yaml.safe_load(f) or {}
    seen.add(config_fn)
    merged = {}
    bases = cur.get('base_config', [])
    if not isinstance(bases, list):
        bases = [bases]
    cur['base_config'] = bases if 'base_config' in cur else cur.get('base_config')
    for base in bases:
        if base.startswith('.'):
            base = os.path.normpath(os.path.join(os.path.dirname(config_fn), base))
        if base not in seen:
            override_config(merged, _load_chain(base, seen, chain))
    if 'base_config' not in cur or cur['base_config'] is None:
        cur.pop('base_config', None)
    override_config(merged, cur)
    chain.append(config_fn)
    return merged


def _apply_override(node_root, assignment):
    key, val = assignment.split('=', 1)
    val = val.strip('\'" ')
    node = node_root
    parts = key.split('.')
    for p in parts[:-1]:
        node = node[p]
    leaf = parts[-1]
    old = node.get(leaf)
    if val in ('True', 'False') or isinstance(old, (bool, list, dict)):
        if isinstance(old, list):
            val = val.replace(' ', ',')
        node[leaf] = ast.literal_eval(val)
    elif old is None:
        try:
            node[leaf] = ast.literal_eval(val)
        except (ValueError, SyntaxError):
            node[leaf] = val
    else:
        node[leaf] = type(old)(val)


def set_hparams(config='', exp_name='', hparams_str='', print_hparams=True, global_hparams=True):
    if config == '' and exp_name == '':
        parser = argparse.ArgumentParser(description='')
        parser.add_argument('--config', type=str, default='configs/config_base.yaml')
        parser.add_argument('--exp_name', type=str, default='')
        parser.add_argument('--hparams', type=str, default='')
        parser.add_argument('--infer', action='store_true')
        parser.add_argument('--validate', action='store_true')
        parser.add_argument('--reset', action='store_true')
        parser.add_argument('--remove', action='store_true')
        parser.add_argument('--debug', action='store_true')
        args, _ = parser.parse_known_args()
    else:
        args = Args(config=config, exp_name=exp_name, hparams=hparams_str, infer=False, validate=False,
                    reset=False, debug=False, remove=False)
    assert args.config != '' or args.exp_name != ''

    chain = []
    saved = {}
    work_dir = ''
    ckpt_config_path = ''
    if args.exp_name != '':
        work_dir = f'checkpoints/{args.exp_name}'
        ckpt_config_path = f'{work_dir}/config.yaml'
        if os.path.exists(ckpt_config_path):
            with open(ckpt_config_path) as f:
                saved.update(yaml.safe_load(f) or {})
    hp = {}
    if args.config != '':
        hp.update(_load_chain(args.config, set(), chain))
    if not args.reset:
        hp.update(saved)
    hp['work_dir'] = work_dir

    if args.hparams != '':
        # split on commas that are not inside [...] so that list values survive
        depth, cur, items = 0, '', []
        for ch in args.hparams:
            depth += ch == '['
            depth -= ch == ']'
            if ch == ',' and depth == 0:
                items.append(cur)
                cur = ''
            else:
                cur += ch
        items.append(cur)
        for item in items:
            if item.strip():
                _apply_override(hp, item)

    if work_dir != '' and getattr(args, 'remove', False):
        if input('REMOVE old checkpoint? Y/N [Default: N]: ').lower() == 'y':
            shutil.rmtree(work_dir, ignore_errors=True)
    if work_dir != '' and (not os.path.exists(ckpt_config_path) or args.reset) and not args.infer:
        os.makedirs(work_dir, exist_ok=True)
        with open(ckpt_config_path, 'w') as f:
            yaml.safe_dump(hp, f)

    hp['infer'], hp['debug'], hp['validate'], hp['exp_name'] = args.infer, args.debug, args.validate, args.exp_name
    global global_print_hparams
    if global_hparams:
        hparams.clear()
        hparams.update(hp)
    if print_hparams and global_print_hparams and global_hparams:
        print('| Hparams chains: ', chain)
        print('| Hparams: ')
        for i, (k, v) in enumerate(sorted(hp.items())):
            print(f'\033[;33;m{k}\033[0m: {v}, ', end='\n' if i % 5 == 4 else '')
        print('')
        global_print_hparams = False
    return hp
