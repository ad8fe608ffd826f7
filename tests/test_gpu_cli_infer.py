"""-m gpu: the reference entry points end to end -- tasks/run.py --infer -> Task.start -> Trainer.test ->
HifiGAN() (no-arg constructor reading hparams['vocoder_ckpt']: config.yaml + model_ckpt_steps_*.ckpt in the
reference's checkpoint layout) -> spec2wav -> wav files."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import yaml

from neuralsvb_b200.utils import synthetic as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_run_py_infer_with_reference_layout_checkpoint(tmp_path):
    ckpt_dir = tmp_path / 'voc'
    ckpt_dir.mkdir()
    h = S.hifigan_config()
    with open(ckpt_dir / 'config.yaml', 'w') as f:
        yaml.safe_dump(h, f)
    sd = S.make_generator_state_dict(h, 1234)
    for step in (10, 200):         # the loader must pick the numerically newest one
        torch.save({'state_dict': {'model_gen': sd if step == 200 else {}}, 'global_step': step, 'epoch': 0,
                    'checkpoint_callback_best': 0.0, 'optimizer_states': []}, ckpt_dir / f'model_ckpt_steps_{step}.ckpt')
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, '-m', 'neuralsvb_b200.tasks.run', '--config', os.path.join(ROOT, 'egs/vocoder_infer_synthetic.yaml'),
                        '--exp_name', 'cli_demo', '--infer', '--hparams', f'vocoder_ckpt={ckpt_dir},num_test_samples=2,test_frames=64'],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'model_ckpt_steps_200.ckpt' in r.stdout and 'vocoder infer: 2 clips' in r.stdout
    from scipy.io import wavfile
    wavs = sorted((tmp_path / 'checkpoints' / 'cli_demo' / 'generated_0').glob('*.wav'))
    assert len(wavs) == 2
    sr, data = wavfile.read(wavs[0])
    assert sr == 22050 and data.dtype == np.int16 and len(data) == 64 * 256 and np.abs(data).max() > 100


def test_run_py_trains_the_vocoder_for_a_few_steps(tmp_path):
    """tasks/run.py -> HifiGanTask.start -> Trainer.fit: generator + MPD + MSD, two optimizers, synthetic clips.
    The mel loss must go down over the steps and the run must leave a checkpoint the vocoder plugin can load."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, '-m', 'neuralsvb_b200.tasks.run', '--config', os.path.join(ROOT, 'egs/vocoder_train_synthetic.yaml'),
                        '--exp_name', 'cli_train', '--reset', '--hparams',
                        'max_updates=6,max_sentences=2,max_samples=8192,num_train_batches=1,val_check_interval=4,tb_log_interval=1,lr=0.001'],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'Training end' in r.stdout
    import re
    mels = [float(m) for m in re.findall(r"\bmel ([0-9.eE+-]+)", r.stdout)]        # '| step N: mel 1.2345, a 0.5, ...' 
    assert len(mels) >= 4 and mels[-1] < mels[0], mels
    ckpts = list((tmp_path / 'checkpoints' / 'cli_train').glob('model_ckpt_steps_*.ckpt'))
    assert ckpts, r.stdout[-2000:]
    sd = torch.load(ckpts[0], map_location='cpu')['state_dict']
    assert 'model_gen' in sd and 'conv_pre.weight_v' in sd['model_gen']
