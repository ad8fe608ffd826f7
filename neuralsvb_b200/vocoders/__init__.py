from neuralsvb_b200.vocoders import hifigan  # noqa: F401  (populates the registry, like vocoders/__init__.py:1-2)
