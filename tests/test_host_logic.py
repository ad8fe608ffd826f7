"""CPU tests of the host-side mirror: registry, config system, checkpoint layout."""
import os

import numpy as np
import pytest
import torch

from neuralsvb_b200.utils import audio, hparams as HP, synthetic as S
from neuralsvb_b200.vocoders.base_vocoder import VOCODERS, BaseVocoder, get_vocoder_cls, register_vocoder


def test_registry_names_and_dotted_path():
    import neuralsvb_b200.vocoders  # noqa: F401
    from neuralsvb_b200.vocoders.hifigan import HifiGAN
    assert VOCODERS['HifiGAN'] is HifiGAN and VOCODERS['hifigan'] is HifiGAN
    assert get_vocoder_cls({'vocoder': 'hifigan'}) is HifiGAN
    assert get_vocoder_cls({'vocoder': 'neuralsvb_b200.vocoders.hifigan.HifiGAN'}) is HifiGAN

    @register_vocoder
    class Dummy(BaseVocoder):
        pass
    assert get_vocoder_cls({'vocoder': 'dummy'}) is Dummy
    with pytest.raises(NotImplementedError):
        Dummy().spec2wav(None)


def test_hparams_chain_override_and_cli(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    (tmp_path / 'a').mkdir()
    (tmp_path / 'base.yaml').write_text('x: 1\nnest: {p: 1, q: 2}\nlst: [1, 2]\nflag: false\nname: abc\n')
    (tmp_path / 'a' / 'mid.yaml').write_text('base_config: ../base.yaml\nx: 2\nnest: {q: 3}\n')
    (tmp_path / 'a' / 'top.yaml').write_text('base_config:\n  - ./mid.yaml\ny: 5.5\n')
    hp = HP.set_hparams('a/top.yaml', hparams_str='lst=[3 4 5],flag=True,nest.p=7,y=1.5,name=zz',
                        print_hparams=False, global_hparams=False)
    assert hp['x'] == 2 and hp['nest'] == {'p': 7, 'q': 3} and hp['lst'] == [3, 4, 5]
    assert hp['flag'] is True and hp['y'] == 1.5 and hp['name'] == 'zz' and hp['work_dir'] == ''
    # exp_name -> config.yaml snapshot written and merged back on the next run
    HP.set_hparams('a/top.yaml', exp_name='e1', hparams_str='x=9', print_hparams=False, global_hparams=True)
    assert os.path.exists('checkpoints/e1/config.yaml') and HP.hparams['x'] == 9
    hp2 = HP.set_hparams('a/top.yaml', exp_name='e1', print_hparams=False, global_hparams=False)
    assert hp2['x'] == 9 and hp2['work_dir'] == 'checkpoints/e1'


@pytest.mark.skipif(not os.path.isdir('/root/reference/egs'), reason='reference mount only in the build container')
def test_hparams_resolves_reference_yaml_like_the_reference(monkeypatch):
    import importlib.util
    monkeypatch.chdir('/root/reference')
    spec = importlib.util.spec_from_file_location('ref_hparams', '/root/reference/utils/hparams.py')
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for cfg in ('egs/egs_bases/tts/vocoder/hifigan.yaml', 'egs/datasets/audio/PopBuTFy/vae_global_mle_eng.yaml'):
        a = HP.set_hparams(cfg, hparams_str='hop_size=128', print_hparams=False, global_hparams=False)
        b = ref.set_hparams(cfg, hparams_str='hop_size=128', print_hparams=False, global_hparams=False)
        assert a == b


def test_generator_state_dict_layout_is_the_checkpoint_contract():
    from neuralsvb_b200.modules.hifigan.hifigan import HifiGanGenerator
    h = S.hifigan_config()
    m = HifiGanGenerator(h)
    keys = list(m.state_dict())
    assert 'conv_pre.weight_g' in keys and 'ups.0.weight_v' in keys and 'm_source.l_linear.weight' in keys
    assert 'resblocks.11.convs2.2.weight_v' in keys and 'noise_convs.3.bias' in keys and len(keys) == 244
    assert tuple(m.state_dict()['ups.0.weight_g'].shape) == (512, 1, 1)     # weight norm over dim 0 = Cin
    sd = S.make_generator_state_dict(h)
    m.load_state_dict(sd, strict=True)
    assert sum(v.numel() for v in sd.values()) == 13_954_154 or sum(v.numel() for v in sd.values()) > 13_900_000


def test_mel_basis_and_pad_helpers():
    fb = audio.mel_filterbank(22050, 1024, 80, 80, 7600)
    assert fb.shape == (80, 513) and fb.dtype == np.float32 and (fb.sum(1) > 0).all()
    x = np.zeros(44100, np.float32)
    assert audio.librosa_pad_lr(x, 1024, 256, 1) == (0, 173 * 256 - 44100)
    l, r = audio.librosa_pad_lr(x, 1024, 256, 2)
    assert l + r == 173 * 256 - 44100
