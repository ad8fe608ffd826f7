"""CLI entry point, same contract as the reference's tasks/run.py:5-14:

    python -m neuralsvb_b200.tasks.run --config <yaml> --exp_name <name> [--reset] [--infer] [--validate]
                                       [--debug] [--hparams "k=v,..."]

``set_hparams()`` resolves the YAML chain, then ``hparams['task_cls']`` (dotted path) is imported and
``.start()`` is called on it.
"""
import importlib

from neuralsvb_b200.utils.hparams import hparams, set_hparams


def run_task():
    assert hparams['task_cls'] != ''
    module_name, _, cls_name = hparams['task_cls'].rpartition('.')
    getattr(importlib.import_module(module_name), cls_name).start()


if __name__ == '__main__':
    set_hparams()
    run_task()
