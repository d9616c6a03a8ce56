"""Drop-in surface checks that need no GPU: constructor parity (same RNG stream => bit-identical initial weights as the
reference), state_dict keys, geometry attributes, stale-caller compatibility and the no-CPU-fallback rule."""
import hashlib
import json
import os

import pytest
import torch


class HP(dict):
    __getattr__ = dict.__getitem__


from test_geometry import ARCH_BASIC  # noqa: E402


def test_wavenet_init_is_bit_identical_to_reference(golden_dir):
    import aewn
    from aewn import geometry as vc
    ref = json.load(open(os.path.join(golden_dir, "init_digest_basic.json")))
    torch.manual_seed(2507)
    parent = vc.VirtualConv(filter_info=1, stride=320, parent=None, name="LC-grid")
    wn = aewn.WaveNet(HP(ARCH_BASIC), parent_vc=parent)
    vc.compute_inputs(wn.vc["end_grcc"], vc.GridRange((0, 10 ** 7), (0, 1024), 1))
    wn.trim_ups_out = torch.tensor([0, wn.vc["beg_grcc"].in_len()], dtype=torch.long)
    wn.post_init(1024)
    sd = wn.state_dict()
    floats = {k: v for k, v in sd.items() if v.dtype == torch.float32}
    assert sorted(floats) == sorted(ref)
    for k, v in floats.items():
        assert hashlib.sha256(v.numpy().tobytes()).hexdigest() == ref[k], k
    assert [k for k in sd if k.endswith("leads")] == [f"conv_layers.{i}.leads" for i in range(20)]
    assert "conv_layers.19.dil_res.weight" not in sd and "conv_layers.18.dil_res.weight" in sd
    assert sum(p.numel() for p in wn.parameters()) == 13508490


def test_old_keyword_constructor_is_accepted():
    """autoencoder_model.py:83-87 still calls WaveNet(**dec_params, parent_vc=..., n_lc_in=...)."""
    import aewn
    from aewn import geometry as vc
    kw = {k: v for k, v in ARCH_BASIC.items() if k != "bias"}
    parent = vc.VirtualConv(filter_info=1, stride=320, parent=None, name="p")
    with torch.device("meta"):
        wn = aewn.WaveNet(parent_vc=parent, **kw)
    assert len(wn.conv_layers) == 20 and wn.bias is True
    with pytest.raises(TypeError):
        aewn.WaveNet(parent_vc=parent, bogus=1, **kw)


def test_no_cpu_fallback():
    import aewn
    from aewn import geometry as vc
    small = dict(ARCH_BASIC, n_res=16, n_dil=16, n_skp=16, n_post=16, n_lc_out=8, n_blocks=1, n_block_layers=2)
    parent = vc.VirtualConv(filter_info=1, stride=320, parent=None, name="p")
    wn = aewn.WaveNet(HP(small), parent_vc=parent)
    vc.compute_inputs(wn.vc["end_grcc"], vc.GridRange((0, 10 ** 7), (0, 8), 1))
    wn.trim_ups_out = torch.tensor([0, wn.vc["beg_grcc"].in_len()], dtype=torch.long)
    wn.post_init(8)
    wn.train()
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        wn(torch.zeros(1, 4000), torch.zeros(1, 64, 20), torch.zeros(1, dtype=torch.long),
           torch.arange(20).unsqueeze(0))


def test_encoder_and_vq_state_dicts(golden_dir):
    from aewn import geometry as vc, wave_encoder, vqema_bn, vq_bn
    g = torch.load(os.path.join(golden_dir, "encoder_small.pt"))
    torch.manual_seed(2507)
    enc = wave_encoder.Encoder(13, 64, vc.VirtualConv(filter_info=400, stride=160, name="MFCC"))
    sd = enc.state_dict()
    assert list(sd) == list(g["state_dict"])
    assert all(torch.equal(sd[k], g["state_dict"][k]) for k in sd)
    gv = torch.load(os.path.join(golden_dir, "vq.pt"))
    torch.manual_seed(2507)
    bn = vqema_bn.VQEMA(96, 32, 0.25, 0.99, 4096, True)
    assert torch.equal(bn.emb, gv["vqema"]["emb"]) and torch.equal(bn.linear.weight, gv["vqema"]["lin_w"])
    assert torch.equal(bn.ema_numer, gv["vqema"]["ema_numer0"])
    torch.manual_seed(2508)
    vq = vq_bn.VQ(96, 64, 0.25, 512)
    assert torch.equal(vq.emb.detach(), gv["vq"]["emb"])
    with pytest.raises(RuntimeError):
        vqema_bn.VQEMA(96, 32, 0.25, 1.0, 64, True)


def test_autoencoder_wiring_geometry_and_state_dicts_match_reference(golden_dir):
    """aewn/autoencoder.py (the product-side stand-in for autoencoder_model.AutoEncoder, which cannot be constructed at
    the reference's HEAD) must derive the same geometry and expose the same state_dict as the reference modules wired
    by oracle/ae_harness.py -- small fixture and the BASELINE cfg3 size."""
    from aewn.autoencoder import AutoEncoder
    g = torch.load(os.path.join(golden_dir, "autoencoder_small.pt"))
    torch.manual_seed(2507)
    ae = AutoEncoder(HP(g["hp"]), g["n_mel"], g["enc_n_out"], "vqvae-ema", g["hp"]["n_lc_in"], 0.25, 0.99, g["K"], True)
    ae.init_geometry(g["W"])
    geo = g["geo"]
    assert (ae.enc_in_len, ae.enc_in_mel_len, ae.embed_len, ae.dec_in_len) == (
        geo["enc_in_len"], geo["enc_in_mel_len"], geo["embed_len"], geo["dec_in_len"])
    assert ae.trim_dec_in.tolist() == geo["trim_dec_in"] and ae.trim_dec_out.tolist() == geo["trim_dec_out"]
    assert ae.decoder.trim_ups_out.tolist() == geo["trim_ups_out"]
    assert [l.leads.tolist() for l in ae.decoder.conv_layers] == geo["leads"]
    for part, mod in (("encoder", ae.encoder), ("bottleneck", ae.bottleneck), ("decoder", ae.decoder)):
        sd, ref = mod.state_dict(), g["state_dict"][part]
        assert list(sd) == list(ref), part
        for k in sd:
            assert sd[k].shape == ref[k].shape, (part, k)
            # same seed, same construction order => bit-identical initial parameters (same RNG stream)
            if k not in ("z_sum", "n_sum"):                      # torch.empty buffers (vqema_bn.py:113-114)
                assert torch.equal(sd[k], ref[k]), (part, k)
    geo3 = json.load(open(os.path.join(golden_dir, "geometry.json")))["cfg3_vqvae_ema_W16384"]
    arch = dict(ARCH_BASIC, n_lc_in=32)
    with torch.device("meta"):
        ae3 = AutoEncoder(HP(arch), 39, 768, "vqvae-ema", 32, 0.25, 0.99, 4096, True)
    ae3.init_geometry(16384)
    assert (ae3.enc_in_len, ae3.enc_in_mel_len, ae3.embed_len, ae3.dec_in_len) == (23280, 144, 65, 18430)
    assert ae3.trim_dec_in.tolist() == geo3["trim_dec_in"] and ae3.trim_dec_out.tolist() == geo3["trim_dec_out"]
    assert ae3.decoder.trim_ups_out.tolist() == geo3["trim_ups_out"]


def test_wide_unit_packing_respects_tmem_and_staging_limits():
    """ops.pack_wide_units (host logic of aewn_wgradw): <= 3 chunks per unit, <= 512 accumulator columns with every chunk
    rounded up to 32, <= 256 staged rows per CTA (64 for a chunk of <= 128 columns, else 128); nothing lost or duplicated."""
    from aewn import ops
    out = torch.zeros(1)

    def ck(n, tag):
        return dict(x_act=1, x_row=0, n_valid=n, shift=0, out=out, out_off=0, out_rs=1, out_cs=1, tag=tag)

    # arch.basic wgrad1: x[t-d] and x[t] in chunks (256, 112) each, plus [cond, 1] = 139
    chunks = [ck(256, "xd0"), ck(112, "xd1"), ck(256, "x0"), ck(112, "x1"), ck(139, "c")]
    units = ops.pack_wide_units(0, 0, 256, 4, 1000, chunks)
    assert [sorted(c["tag"] for c in u["chunks"]) for u in units] == [["x0", "xd0"], ["c", "x1", "xd1"]]
    for widths in ([256, 256, 256, 256, 144], [48], [16] * 7, [130, 130, 130, 130], [256, 100, 16]):
        units = ops.pack_wide_units(0, 0, 200, 0, 640, [ck(n, i) for i, n in enumerate(widths)])
        seen = sorted(c["tag"] for u in units for c in u["chunks"])
        assert seen == list(range(len(widths)))
        for u in units:
            assert 1 <= len(u["chunks"]) <= 3
            assert sum((c["n"] + 31) // 32 * 32 for c in u["chunks"]) <= 512
            assert sum(128 if c["n"] > 128 else 64 for c in u["chunks"]) <= 256
            assert all(c["n"] % 16 == 0 and c["n"] >= c["n_valid"] for c in u["chunks"])


_CFG1_SCRIPT = r"""
import hashlib, json, sys, torch
import wavenet, hparams, mfcc_inverter
hps = hparams.setup_hparams("mfcc_inverter,mfcc,train", dict(n_win_batch=4096, n_batch=2))
torch.manual_seed(2507)
mi = mfcc_inverter.MfccInverter(hps)
sd = mi.wavenet.state_dict()
print(json.dumps(dict(
    module=type(mi.wavenet).__module__, enc_in_len=mi.enc_in_len, embed_len=mi.embed_len, dec_in_len=mi.dec_in_len,
    trim_dec_in=mi.trim_dec_in.tolist(), trim_dec_out=mi.trim_dec_out.tolist(),
    wav_cond_offset=list(mi.wavenet.wav_cond_offset), leads=[l.leads.tolist() for l in mi.wavenet.conv_layers],
    n_params=sum(p.numel() for p in mi.wavenet.parameters()),
    digest={k: hashlib.sha256(v.numpy().tobytes()).hexdigest() for k, v in sd.items() if v.dtype == torch.float32},
    keys=list(sd))))
"""


@pytest.mark.needs_reference
def test_cfg1_unchanged_mfcc_inverter_constructs_on_the_dropins(golden_dir):
    """BASELINE cfg1 (plumbing, CPU): the reference's UNMODIFIED mfcc_inverter.MfccInverter (arch.mi, window 4096), with
    ae-wavenet_b200/dropin in front of the reference tree, builds on the kernel-backed WaveNet with the reference's own
    geometry, state_dict keys and bit-identical initial weights (same RNG stream).  Both variants run in subprocesses so
    that the flat module names (`wavenet`, ...) cannot leak into this test session."""
    import subprocess
    import sys
    root = os.path.dirname(golden_dir.rstrip("/")).rsplit("/tests", 1)[0]
    pkg = os.path.join(root, "ae-wavenet_b200")

    def run(paths):
        env = dict(os.environ, PYTHONPATH=os.pathsep.join(paths), PYTHONWARNINGS="ignore")
        out = subprocess.run([sys.executable, "-c", _CFG1_SCRIPT], env=env, capture_output=True, text=True, cwd="/tmp",
                             timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        return json.loads(out.stdout.strip().splitlines()[-1])

    ours = run([os.path.join(pkg, "dropin"), pkg, "/root/reference"])
    ref = run(["/root/reference"])
    assert ours["module"] == "aewn.wavenet" and ref["module"] == "wavenet"
    geo = json.load(open(os.path.join(golden_dir, "geometry.json")))["cfg1_mi_W4096"]
    for k in ("enc_in_len", "embed_len", "dec_in_len", "trim_dec_in", "trim_dec_out", "wav_cond_offset", "leads"):
        assert ours[k] == ref[k] == geo[k], k
    assert ours["n_params"] == ref["n_params"] == 13498890            # arch.mi decoder, SURVEY.md 8a
    assert ours["keys"] == ref["keys"]
    assert ours["digest"] == ref["digest"]                             # bit-identical initial parameters
