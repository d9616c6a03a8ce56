"""aewn -- B200 (sm_100a) hot path of hrbigelow/ae-wavenet behind the reference's own nn.Module surface.

Importing the package never needs a GPU; running a module does, and needs the in-tree libaewn.so (no CPU fallback).
"""
from . import _lib  # noqa: F401
from .wavenet import (Conditioning, Conv1dWrap, GatedResidualCondConv, RecLoss, Upsampling, WaveNet)  # noqa: F401

__all__ = ["WaveNet", "GatedResidualCondConv", "Conditioning", "Upsampling", "Conv1dWrap", "RecLoss"]
