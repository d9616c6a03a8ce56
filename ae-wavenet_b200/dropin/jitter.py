"""Drop-in for the reference's jitter.py (Jitter, jitter.py:3-33): same constructor and `__call__(win_size)` -> int array,
consuming the same numpy RandomState draws; `Jitter.batch(B, win_size)` draws a whole batch on the device.  Put this
directory BEFORE the reference tree on sys.path."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from aewn.loader import Jitter  # noqa: E402,F401
