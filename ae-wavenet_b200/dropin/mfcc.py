"""Drop-in for the reference's mfcc.py (ProcessWav, mfcc.py:27-76): same constructor, `.vc`, `.n_out` and `__call__(wav)`
-> (3 * n_mfcc, frames) numpy array, computed by csrc/loader.cu on the GPU instead of librosa on the host.  Put this
directory BEFORE the reference tree on sys.path.  `ProcessWav.batch` (whole batches, device tensors) is the
call `aewn.loader.Collate` uses; run it in the process that owns the CUDA context (DataLoader with num_workers = 0)."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from aewn.loader import ProcessWav  # noqa: E402,F401
