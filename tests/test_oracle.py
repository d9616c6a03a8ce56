"""Pins the CPU oracle (oracle/torch_oracle.py, oracle/vq_oracle.c): against golden vectors produced by the unmodified
reference (always) and against the reference modules themselves (when /root/reference is present)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import torch_oracle as orc


def test_wavenet_small_forward_backward_equals_reference_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "wavenet_small.pt"))
    sd = {k: v.clone().requires_grad_(v.dtype == torch.float32 and k != "cond.eye" and not k.endswith("leads"))
          for k, v in g["state_dict"].items()}
    geo = dict(g["geo"], trim_ups_out=g["trim_ups_out"], n_win_batch=g["W"])
    lc = g["lc"].clone().requires_grad_(True)
    loss, quant = orc.decoder_loss(sd, g["hp"], geo, g["wav"], lc, g["spk"], g["jit"])
    assert torch.allclose(quant, g["quant"], rtol=0, atol=1e-5)
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    loss.backward()
    assert torch.allclose(lc.grad, g["lc_grad"], rtol=1e-4, atol=1e-7)
    for k, ref in g["grads"].items():
        assert torch.allclose(sd[k].grad, ref, rtol=1e-3, atol=1e-6), k


@pytest.mark.parametrize("name", ["d1", "d8", "d2_final"])
def test_grcc_layer_equals_reference_golden(golden_dir, name):
    c = torch.load(os.path.join(golden_dir, "grcc_layer.pt"))["cases"][name]
    p = {k: v for k, v in c["state_dict"].items() if k != "leads"}
    sig, skp = orc.grcc_layer(c["x"], c["cond"], p, c["dil"], c["leads"], c["final"])
    assert torch.allclose(sig, c["sig"], atol=1e-6) and torch.allclose(skp, c["skp"], atol=1e-6)


def test_encoder_equals_reference_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "encoder_small.pt"))
    y, fracs = orc.encoder_forward(g["state_dict"], g["x"])
    assert torch.allclose(y, g["y"], atol=1e-6)
    assert np.allclose([float(f) for f in fracs], g["frac_zero"], atol=1e-12)


def test_vqema_equals_reference_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "vq.pt"))["vqema"]
    r = orc.vqema_forward(g["lin_w"], g["emb"], g["z"], g["ema_numer0"], g["ema_denom0"], g["ema_gamma"])
    assert torch.equal(r["min_ind"], g["min_ind"])
    assert torch.allclose(r["min_dist"], g["min_dist"], atol=1e-7)
    assert torch.equal(r["out"], g["out"])
    assert torch.allclose(r["z_sum"], g["z_sum"], atol=1e-5) and torch.equal(r["n_sum"], g["n_sum"])
    assert torch.allclose(r["ema_numer"], g["ema_numer"], atol=1e-6)
    assert torch.allclose(r["ema_denom"], g["ema_denom"], atol=1e-7)


def test_vq_equals_reference_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "vq.pt"))["vq"]
    r = orc.vq_forward(g["lin_w"], g["emb"], g["z"])
    assert torch.equal(r["min_ind"], g["min_ind"]) and torch.equal(r["out"], g["out"])


def _c_oracle():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "liboracle.so")
    if not os.path.isfile(path):
        pytest.skip("oracle/liboracle.so not built (run `make oracle` or __graft_entry__.build())")
    return ctypes.CDLL(path)


@pytest.mark.parametrize("which,metric", [("vqema", 1), ("vq", 0)])
def test_c_vq_oracle_indices_equal_reference_golden(golden_dir, which, metric):
    """The C restatement (fixed summation order) reproduces the reference's argmin on the golden inputs."""
    lib = _c_oracle()
    g = torch.load(os.path.join(golden_dir, "vq.pt"))[which]
    ze = np.ascontiguousarray(g["ze"].numpy())
    emb = np.ascontiguousarray(g["emb"].numpy())
    B, d, N = ze.shape
    K = emb.shape[0]
    ind = np.zeros(B * N, dtype=np.int64)
    dist = np.zeros(B * N, dtype=np.float32)
    lib.vq_assign_oracle(ze.ctypes.data_as(ctypes.c_void_p), emb.ctypes.data_as(ctypes.c_void_p), B, d, N, K, metric,
                         ind.ctypes.data_as(ctypes.c_void_p), dist.ctypes.data_as(ctypes.c_void_p))
    ref = g["min_ind"].numpy().reshape(-1)
    mism = np.nonzero(ind != ref)[0]
    assert mism.size == 0, f"{mism.size} index mismatches vs the reference"
    assert np.allclose(dist, g["min_dist"].numpy().reshape(-1), rtol=1e-6, atol=1e-7)


@pytest.mark.needs_reference
def test_oracle_matches_live_reference_on_fresh_seed():
    from oracle import ref_harness as rh
    from oracle.make_golden import SMALL, synth_inputs
    m = rh.load()
    hp = rh.HP(SMALL)
    torch.manual_seed(99)
    wn, geo = rh.standalone_wavenet(hp, 64)
    wn.train()
    wav, lc, spk, jit = synth_inputs(2, geo, hp.n_lc_in, hp.n_speakers, 4321, jitter=True)
    quant = wn(wav, lc, spk, jit)
    geo = dict(geo, trim_ups_out=wn.trim_ups_out.tolist(), n_win_batch=64)
    mine = orc.wavenet_forward_train(dict(wn.state_dict()), dict(hp), geo, wav, lc, spk, jit)
    assert torch.allclose(mine, quant, atol=1e-5)


def test_forward_test_oracle_equals_reference_golden(golden_dir):
    """The sampler restatement reproduces the reference's own forward_test run (wavenet.py:367-531) draw for draw."""
    g = torch.load(os.path.join(golden_dir, "forward_test.pt"))
    out, probs = orc.wavenet_forward_test(g["state_dict"], g["hp"], g["geo"]["wav_cond_offset"], g["wav"], g["lc"],
                                          g["spk"], g["jit"], g["n_rep"], g["uniforms"])
    assert out.shape == g["out"].shape and out.dtype == g["out"].dtype
    assert torch.equal(out, g["out"])
    assert probs.shape == g["probs"].shape
    assert float((probs - g["probs"].float()).abs().max()) < 1e-3      # golden probabilities are stored in fp16
    rf1 = g["base_global_rf"]
    assert torch.equal(out[1:, :rf1], out[:1, :rf1].expand(2, -1))      # primed with the input
    n_ts = rf1 + probs.shape[0]
    assert torch.equal(out[1:, n_ts:], out[:1, n_ts:].expand(2, -1))    # untouched past the conditioning sequence


def test_autoencoder_step_oracle_equals_reference_golden(golden_dir):
    """cfg3 in miniature: the travelling oracle's Encoder -> VQEMA -> WaveNet step against the golden written by
    oracle/ae_harness.py from the unmodified reference modules (values, VQ indices, EMA statistics, every gradient)."""
    from oracle import torch_oracle as O
    g = torch.load(os.path.join(golden_dir, "autoencoder_small.pt"))
    sd = {part: {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and k in g["grads"][part] else v)
                 for k, v in d.items()} for part, d in g["state_dict"].items()}
    mels = g["mels"].clone().requires_grad_(True)
    r = O.autoencoder_step(sd, g["hp"], g["geo"], mels, g["wav_dec"], g["spk"], g["jit"], 0.25, 0.99)
    assert torch.equal(r["min_ind"], g["min_ind"])
    assert float((r["quant"] - g["quant"]).abs().max()) < 1e-5
    assert abs(float(r["com"]) - float(g["com"])) < 1e-5 and abs(float(r["rec"]) - float(g["rec"])) < 1e-6
    for k in ("z_sum", "n_sum", "ema_numer", "ema_denom"):
        assert float((r[k] - g[k]).abs().max()) < 1e-5, k
    (r["com"] + r["rec"]).backward()
    assert float((mels.grad - g["mel_grad"]).abs().max()) <= 1e-3 * float(g["mel_grad"].abs().max())
    for part, grads in g["grads"].items():
        for k, ref in grads.items():
            got = sd[part][k].grad
            assert got is not None, (part, k)
            assert float((got - ref).abs().max()) <= 1e-3 * max(float(ref.abs().max()), 1e-12), (part, k)
