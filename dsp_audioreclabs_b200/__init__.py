"""B200-native (sm_100a) front end + KNN for DSP-AudioRecLabs.

The compute lives in ``libdspfront.so`` (hand-written CUDA, C ABI in include/dspfront.h).
This package is the host-side mirror of the reference's interface for that path:

  * ``batch``   -- batched NumPy API over the C ABI
  * ``device``  -- the same calls on torch CUDA tensors (no host copies)
  * ``dist``    -- utterance / train-row sharding across the GPUs of one box
  * ``dropin/`` -- ``src`` and ``config`` modules with the reference's import paths and
                   signatures, so run.py / ablation_study.py / the experiments run unchanged

There is no CPU implementation: importing the compute entry points without the built
library, or calling them without a CUDA device, raises.
"""
from ._capi import DspError, LIB_PATH, load_library  # noqa: F401

__all__ = ["DspError", "LIB_PATH", "load_library"]
