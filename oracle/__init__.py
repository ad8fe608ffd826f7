"""CPU oracle for the NeuralSVB mel-to-waveform hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package, and only as
the checker (or the timed CPU baseline) -- never on the product path.  The
product (``neuralsvb_b200``) fails loudly when its CUDA library is missing; it
never falls back to anything in here.

What is in here
  frontend.py   numpy restatement of the wav->mel front end
                (data_gen/tts/data_gen_utils.py:93-147 + librosa==0.8.0 stft /
                filters.mel semantics; librosa is a third-party dependency that
                is pinned in Requirements.txt:41 and absent from this image).
  hifigan.py    torch-CPU fp32 restatement of the HiFi-GAN-NSF generator,
                SineGen / SourceModuleHnNSF, MPD / MSD, GAN losses, the torch
                mel_spectrogram and the multi-resolution STFT loss
                (modules/hifigan/hifigan.py, modules/parallel_wavegan/models/source.py,
                modules/hifigan/mel_utils.py, modules/parallel_wavegan/losses/stft_loss.py).
  ref_harness.py / gen_golden.py
                import the REAL reference from /root/reference (build container
                only; the mount does not exist on the GPU box) and write the
                golden fixtures under tests/golden/.

Parity pinning: the reference ships no tests, golden vectors or fixtures
 -- so the oracle is pinned against outputs of the reference
itself, generated in the build container by gen_golden.py from the unmodified
reference modules (torch side) and against torch.stft / torchaudio's Slaney
filterbank (librosa side, since librosa cannot be installed).  tests/test_oracle_golden.py
re-checks the oracle against those committed fixtures on every run.
"""
