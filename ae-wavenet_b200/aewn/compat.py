"""Bindings to the caller's environment.

``vconv``: when the reference tree is importable (the drop-in scenario: mfcc_inverter.py / autoencoder_model.py call
``vconv.compute_inputs`` on OUR ``.vc`` objects) the modules must build their VirtualConv chain from the caller's own
``vconv`` module; otherwise they use the in-package restatement (geometry.py), which is checked against the reference
in tests/test_geometry.py.
"""
import importlib
import os
import sys

from torch import nn


def _pick_vconv():
    if os.environ.get("AEWN_FORCE_OWN_GEOMETRY") == "1":
        from . import geometry
        return geometry
    mod = sys.modules.get("vconv")
    if mod is None:
        try:
            mod = importlib.import_module("vconv")
        except ImportError:
            mod = None
    if mod is not None and all(hasattr(mod, n) for n in ("VirtualConv", "GridRange", "compute_inputs", "output_offsets")):
        return mod
    from . import geometry
    return geometry


vconv = _pick_vconv()


def xavier_init(mod):
    """netmisc.xavier_init (netmisc.py:10-14): Xavier-uniform weights, zero biases."""
    if hasattr(mod, "weight") and mod.weight is not None:
        nn.init.xavier_uniform_(mod.weight)
    if hasattr(mod, "bias") and mod.bias is not None:
        nn.init.constant_(mod.bias, 0)
