"""Vocoder inference task: drives the B200 vocoder plugin through the reference's own entry points
(``tasks/run.py --infer`` -> ``Task.start()`` -> ``Trainer.test`` -> ``test_step`` -> ``spec2wav``),
the way every NeuralSVB task uses the vocoder at test time (``tasks/tts/fs2.py:371-457``,
``tasks/singing/svb_vae_task.py:346-356``; SURVEY D2: in the reference the vocoder is inference-only).

The reference names ``tasks.vocoder.hifigan.HifiGanTask`` in ``egs/egs_bases/tts/vocoder/hifigan.yaml:2``
but ships no such module; the G + MPD + MSD *training* step behind that name needs the
backward kernels that are not built yet, so ``HifiGanTask`` here refuses to train
and ``HifiGanInferTask`` provides the inference half.

Test items: ``hparams['test_input_dir']`` with ``*.npz`` files holding ``mel [T, 80]`` (log10) and
optionally ``f0 [T]``; without it, ``hparams['num_test_samples']`` synthetic clips of
``hparams['test_frames']`` frames (there is no dataset on the build / GPU boxes).
"""
import glob
import os
import time

import numpy as np

from neuralsvb_b200.tasks.base_task import BaseTask
from neuralsvb_b200.utils import audio, ddp_utils
from neuralsvb_b200.utils.hparams import hparams
from neuralsvb_b200.utils.synthetic import make_mel_f0
from neuralsvb_b200.vocoders.base_vocoder import get_vocoder_cls


class HifiGanInferTask(BaseTask):
    def build_model(self):
        self.vocoder = get_vocoder_cls(hparams)()          # e.g. vocoder: neuralsvb_b200.vocoders.hifigan.HifiGAN
        return None

    def configure_optimizers(self):
        return []

    def test_dataloader(self):
        in_dir = hparams.get('test_input_dir', '')
        if in_dir:
            items = []
            for fn in sorted(glob.glob(os.path.join(in_dir, '*.npz'))):
                z = np.load(fn)
                items.append({'name': os.path.splitext(os.path.basename(fn))[0], 'mel': z['mel'].astype(np.float32),
                              'f0': z['f0'].astype(np.float32) if 'f0' in z else None})
        else:
            items = []
            for i in range(int(hparams.get('num_test_samples', 4))):
                mel, f0 = make_mel_f0(1, int(hparams.get('test_frames', 689)), seed=hparams['seed'] + i)
                items.append({'name': f'synthetic_{i:03d}', 'mel': mel[0].T.contiguous().numpy(), 'f0': f0[0].numpy()})
        rank, world, _ = ddp_utils.dist_env()
        return ddp_utils.shard(items, rank, world)          # clips are the sharding unit; no collective

    def test_start(self):
        self.gen_dir = os.path.join(hparams['work_dir'] or '.', f'generated_{self.trainer.global_step}')
        os.makedirs(self.gen_dir, exist_ok=True)

    def test_step(self, sample, batch_idx):
        t0 = time.time()
        wav = self.vocoder.spec2wav(sample['mel'], f0=sample['f0'] if hparams.get('use_pitch_embed', True) else None)
        dt = time.time() - t0
        audio.save_wav(wav.copy(), os.path.join(self.gen_dir, sample['name'] + '.wav'), hparams['audio_sample_rate'])
        return {'audio_s': len(wav) / hparams['audio_sample_rate'], 'wall_s': dt}

    def test_end(self, outputs):
        a, w = sum(o['audio_s'] for o in outputs), sum(o['wall_s'] for o in outputs)
        print(f'| vocoder infer: {len(outputs)} clips, {a:.2f} audio-s in {w:.3f} s -> {a / max(w, 1e-9):.1f} audio-s/s '
              f'(end to end per clip, incl. H2D/D2H)')
        return {'audio_s': a, 'wall_s': w}


class HifiGanTask(HifiGanInferTask):
    """The name the reference config points at.  Inference works; training is not implemented yet."""

    def training_step(self, sample, batch_idx, optimizer_idx=-1):
        raise NotImplementedError('HiFi-GAN G + MPD + MSD training needs the backward kernels (DESIGN.md section 7); '
                                  'run with --infer')
