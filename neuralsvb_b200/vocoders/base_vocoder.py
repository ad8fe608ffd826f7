"""Vocoder plugin registry -- the reference's selection mechanism, same names and behaviour
(reference: vocoders/base_vocoder.py:2-39).  ``hparams['vocoder']`` is either a registered
name (class name or its lower-case form) or a dotted ``package.module.Class`` path, so a
reference checkout selects this implementation with
``vocoder: neuralsvb_b200.vocoders.hifigan.HifiGAN`` and no code change."""
import importlib

VOCODERS = {}


def register_vocoder(cls):
    for key in (cls.__name__, cls.__name__.lower()):
        VOCODERS[key] = cls
    return cls


def get_vocoder_cls(hparams):
    name = hparams['vocoder']
    if name in VOCODERS:
        return VOCODERS[name]
    module_name, _, cls_name = name.rpartition('.')
    return getattr(importlib.import_module(module_name), cls_name)


class BaseVocoder:
    def spec2wav(self, mel):
        """mel [T, 80] -> wav [T * hop]"""
        raise NotImplementedError

    @staticmethod
    def wav2spec(wav_fn):
        """wav file (or array) -> (wav, mel [T, 80])"""
        raise NotImplementedError
