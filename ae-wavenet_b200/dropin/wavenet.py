"""Drop-in for the reference's wavenet.py: put this directory BEFORE the reference tree on sys.path
and the reference's callers (`mfcc_inverter.py`, `autoencoder_model.py`, `chassis.py`) run on the B200 kernel path
unchanged.  Everything is re-exported from the `aewn` package next to this directory."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from aewn.wavenet import GatedResidualCondConv, Conditioning, Upsampling, Conv1dWrap, WaveNet, RecLoss  # noqa: E402,F401
