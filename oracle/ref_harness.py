"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference modules from /root/reference (CPU, PyTorch).

Only usable where /root/reference exists (the build container).  It is used to
  * pin oracle/torch_oracle.py (the travelling CPU restatement) against the real reference, and
  * generate the committed golden vectors under tests/golden/ (see oracle/make_golden.py).
Nothing in the product package (ae-wavenet_b200/) may import this file.

Shims applied to the reference at import time (all oracle-side):
  F5  util.gather_md_scriptable is undefined            -> alias to util.gather_md_jit        (util.py:157-208)
  F4  vq_bn uses StopGrad/ReplaceGrad without importing  -> inject from vqema_bn               (vq_bn.py:14-15)
  F6  VQEMA.forward prints three tensors per call        -> stdout silenced around the call    (vqema_bn.py:204-206)
"""
import contextlib
import io
import os
import sys

REF = os.environ.get("AEWN_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "wavenet.py"))


_mods = None


def load():
    """Import the reference modules (once) and return them as a namespace dict."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    # the reference is a flat directory of modules with generic names; keep it at the FRONT of sys.path only while
    # importing so that `import wavenet` resolves to the reference and not to a drop-in shim
    saved = list(sys.path)
    for name in ("wavenet", "wave_encoder", "vq_bn", "vqema_bn", "vconv", "netmisc", "util", "hparams",
                 "mfcc_inverter", "mfcc", "data", "parse_tools", "jitter"):
        if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith(REF):
            del sys.modules[name]
    sys.path.insert(0, REF)
    try:
        import util, vconv, netmisc, wavenet, wave_encoder, vqema_bn, vq_bn, hparams, mfcc_inverter, jitter
    finally:
        sys.path[:] = saved
    util.gather_md_scriptable = util.gather_md_jit
    vq_bn.StopGrad, vq_bn.ReplaceGrad = vqema_bn.StopGrad, vqema_bn.ReplaceGrad
    _mods = dict(util=util, vconv=vconv, netmisc=netmisc, wavenet=wavenet, wave_encoder=wave_encoder,
                 vqema_bn=vqema_bn, vq_bn=vq_bn, hparams=hparams, mfcc_inverter=mfcc_inverter, jitter=jitter)
    return _mods


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


class HP(dict):
    """attribute dict, same access protocol as hparams.Hyperparams (hparams.py:6-20)"""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


ARCH_BASIC = dict(filter_sz=2, n_lc_out=128, lc_upsample_strides=[5, 4, 4, 4], lc_upsample_filt_sizes=[25, 16, 16, 16],
                  n_res=368, n_dil=256, n_skp=256, n_post=256, n_quant=256, n_blocks=2, n_block_layers=10,
                  n_global_embed=10, n_speakers=40, bias=True, n_lc_in=64)


def standalone_wavenet(hps, n_win_batch, parent_stride=None, WaveNet=None, vconv=None):
    """Build a stand-alone decoder exactly the way MfccInverter._init_geometry does (mfcc_inverter.py:38-65), behind a
    1-tap parent VirtualConv whose stride equals the total upsampling factor (SURVEY.md 8d, cfg2 recipe).
    Works for the reference WaveNet or (WaveNet=..., vconv=...) for the drop-in one."""
    import numpy as np
    import torch
    if WaveNet is None:
        m = load()
        WaveNet, vconv = m["wavenet"].WaveNet, m["vconv"]
    if parent_stride is None:
        parent_stride = int(np.prod(hps.lc_upsample_strides))
    parent = vconv.VirtualConv(filter_info=1, stride=parent_stride, parent=None, name="LC-grid")
    wn = WaveNet(hps, parent_vc=parent)
    end_gr = vconv.GridRange((0, 10 ** 7), (0, n_win_batch), 1)
    vconv.compute_inputs(wn.vc["end_grcc"], end_gr)
    beg = wn.vc["beg_grcc"]
    geo = dict(
        wav_len=parent.in_len(),
        lc_len=parent.child.in_len(),
        dec_in_len=beg.in_len(),
        trim_dec_in=[beg.input_gr.sub[0] - parent.input_gr.sub[0], beg.input_gr.sub[1] - parent.input_gr.sub[0]],
        trim_dec_out=[end_gr.sub[0] - parent.input_gr.sub[0], end_gr.sub[1] - parent.input_gr.sub[0]],
    )
    wn.trim_ups_out = torch.tensor([0, beg.in_len()], dtype=torch.long)
    wn.post_init(n_win_batch)
    geo["wav_cond_offset"] = list(wn.wav_cond_offset)
    geo["leads"] = [layer.leads.tolist() for layer in wn.conv_layers]
    return wn, geo
