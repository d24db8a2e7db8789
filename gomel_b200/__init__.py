"""gomel_b200 -- B200-native (sm_100a) implementation of the neurlang/gomel spectrogram hot path.

Host-side mirrors of the reference's public API on top of the C ABI of libgomelcuda.so
(include/gomel_cuda.h):

    gomel_b200.mel.Mel      <->  Go package mel   (mel/mel.go)
    gomel_b200.phase.Phase  <->  Python phase.py / Go package phase (phase/phase.go)

There is no CPU fallback: without the CUDA library or a GPU every transform raises.
"""
from . import _lib  # noqa: F401
from .mel import Mel, NewMel  # noqa: F401
from .phase import Phase  # noqa: F401

__all__ = ["Mel", "NewMel", "Phase"]
