"""BASELINE cfg1 on the CUDA path: the MfccInverter geometry (par/arch.mi.json: upsampling strides [5,4,4,2], 39 conditioning
channels, batch 2, window 4096) through the kernel-backed WaveNet, against the result of the reference's UNMODIFIED
mfcc_inverter.MfccInverter.run on CPU (tests/golden/cfg1_mi.pt, written by oracle/make_golden.py cfg1).

The reference's mfcc_inverter.py does not travel to the GPU box, so the test restates its wiring (mfcc_inverter.py:15-65:
an MFCC VirtualConv parent of window 400 / hop 160, WaveNet(hps, parent_vc), _init_geometry) and its run()
(mfcc_inverter.py:89-107: forward, RecLoss on pred[..., :-1] vs wav[trim_dec_out][1:], autograd.grad w.r.t. mel with
retain_graph, then the caller's loss.backward(), chassis.py:157).  The 13.5 M initial parameters come from the same seed on
both sides and are pinned by SHA-256 digests (the CPU suite pins the same digests against the live reference)."""
import hashlib
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


class HP(dict):
    __getattr__ = dict.__getitem__


def build_mfcc_inverter_decoder(g):
    import aewn
    from aewn import geometry as vc
    hps = HP(g["arch"])
    mfcc_vc = vc.VirtualConv(filter_info=hps.mfcc_win_sz, stride=hps.mfcc_hop_sz, parent=None, name="MFCC")
    torch.manual_seed(2507)
    wn = aewn.WaveNet(hps, parent_vc=mfcc_vc)
    # MfccInverter._init_geometry, mfcc_inverter.py:38-65
    end_gr = vc.GridRange((0, 100000), (0, g["W"]), 1)
    vc.compute_inputs(wn.vc["end_grcc"], end_gr)
    beg = wn.vc["beg_grcc"]
    di, wi = beg.input_gr, mfcc_vc.input_gr
    geo = dict(enc_in_len=mfcc_vc.in_len(), embed_len=mfcc_vc.child.in_len(), dec_in_len=beg.in_len(),
               trim_dec_in=[di.sub[0] - wi.sub[0], di.sub[1] - wi.sub[0]],
               trim_dec_out=[end_gr.sub[0] - wi.sub[0], end_gr.sub[1] - wi.sub[0]])
    wn.trim_ups_out = torch.tensor([0, beg.in_len()], dtype=torch.long)
    wn.post_init(g["W"])
    return wn, geo


def test_cfg1_mfcc_inverter_run_matches_the_reference_cpu_result(golden_dir):
    import aewn
    from aewn import ops
    g = torch.load(os.path.join(golden_dir, "cfg1_mi.pt"))
    wn, geo = build_mfcc_inverter_decoder(g)
    for k in ("enc_in_len", "embed_len", "dec_in_len", "trim_dec_out"):
        assert geo[k] == g[k], k
    assert list(wn.wav_cond_offset) == list(g["wav_cond_offset"])
    sd = wn.state_dict()
    assert {k: hashlib.sha256(v.numpy().tobytes()).hexdigest() for k, v in sd.items() if v.dtype == torch.float32} \
        == g["digest"]                                   # bit-identical initial parameters (same RNG stream)
    wn = wn.cuda().train()
    wav, voice, jit = g["wav"].cuda(), g["voice"].cuda(), g["jit"].cuda()
    mel = g["mel"].cuda().requires_grad_(True)
    quant = wn(wav, mel, voice, jit)                     # MfccInverter.forward -> WaveNet.forward_train
    assert quant.shape == (2, 256, g["W"])
    t0, t1 = g["trim_dec_out"]
    pred, target = quant[..., :-1], wav[:, t0:t1][..., 1:]
    loss = aewn.RecLoss()(pred, target)
    (mel_grad,) = torch.autograd.grad(loss, mel, retain_graph=True)          # mfcc_inverter.py:103
    for p in wn.parameters():
        p.grad = None
    loss.backward()                                                          # chassis.py:157
    ops.check_device_errors()
    ref_tail = g["pred_tail"]
    err = float((pred[:, :, -64:].detach().cpu() - ref_tail).abs().max()) / float(ref_tail.abs().max())
    assert err < 5e-3, err                               # fp16/TF32 operands (10-bit mantissa), fp32 accumulation
    assert abs(float(loss) - float(g["loss"])) < 2e-3
    gerr = float((mel_grad.cpu() - g["mel_grad"]).abs().max()) / float(g["mel_grad"].abs().max())
    assert gerr < 5e-2, gerr                             # behind the whole stack and cuDNN's TF32 front-end backward
    a, b = mel_grad.cpu().double().flatten(), g["mel_grad"].double().flatten()
    assert float(a @ b / (a.norm() * b.norm())) > 0.999
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in wn.parameters())
    # inference mode (MfccInverter.forward's eval branch wraps the call in no_grad, mfcc_inverter.py:83-87; the module stays
    # in train() here because eval() selects the incremental sampler): same logits, nothing saved for a backward pass
    with torch.no_grad():
        quant_i = wn(wav, g["mel"].cuda(), voice, jit)
    ops.check_device_errors()
    assert float((quant_i - quant.detach()).abs().max()) <= 1e-5 * float(quant.detach().abs().max())
