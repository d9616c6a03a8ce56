"""csrc/loader.cu through aewn/loader.py against oracle/loader_oracle.py and the reference goldens (tests/golden/loader.pt,
written from util.py / jitter.py by `oracle/make_golden.py loader`).

Bars.  Integer work is bit-exact: jitter indices (given the same uniforms) and mu-law codes -- except that a code may
differ where the pre-quantisation value lies within 2 float32 ulps of the quantiser's decision point (CUDA's log1pf and
glibc's are both faithfully, not identically, rounded); the test counts those and bounds them.  mu-law decode: 2e-6
absolute (powf).  MFCC rows (dB-scaled cepstra, |values| up to ~500): 2e-3 absolute against the double-precision-FFT oracle,
derivative rows the same."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def test_mu_law_codec_matches_the_reference_golden(golden_dir):
    from aewn import loader
    g = torch.load(os.path.join(golden_dir, "loader.pt"))
    x = g["x"].cuda()
    for name, fn in (("enc_np", loader.mu_encode_np), ("enc_torch", loader.mu_encode_torch)):
        got = fn(x, 256).cpu().long()
        ref = g[name].long()
        bad = (got != ref).nonzero().flatten()
        assert (got - ref).abs().max() <= 1 and len(bad) <= 2, (name, len(bad))
        # every disagreement sits on a decision point of the quantiser
        xs = g["x"][bad].double()
        v = (torch.sign(xs) * torch.log1p(255 * xs.abs()) / np.log1p(255.0) + 1) * 0.5 * 255 + 0.5
        frac = v - torch.floor(v) if name == "enc_np" else (v - torch.floor(v) - 0.5)
        assert bool(((frac.abs() < 1e-4) | ((1 - frac).abs() < 1e-4)).all()), (name, v)
    q = torch.arange(256).cuda()
    dec = loader.mu_decode_torch(q, 256).cpu()
    assert float((dec - g["dec_torch"]).abs().max()) < 2e-6
    assert float((dec - g["dec_np"]).abs().max()) < 2e-6
    # (decode -> encode is not a usable property here: the reference's decoder puts code q at (2q - 1) / mu - 1, i.e. EXACTLY
    # on the encoder's decision point between q - 1 and q, util.py:62-78, so the round trip is decided by the last ulp)


def test_jitter_indices_match_the_reference_draw_for_draw(golden_dir):
    from aewn import loader
    import loader_oracle as lo
    g = torch.load(os.path.join(golden_dir, "loader.pt"))
    for j in g["jitter"]:
        jit = loader.Jitter(j["prob"])
        got = jit._indices(j["uniforms"].cuda().view(1, -1), 1, j["win"])[0].cpu()
        assert torch.equal(got, j["index"]), j["seed"]
        np.random.seed(j["seed"])                  # the host contract: same numpy stream, same array as jitter.Jitter
        assert np.array_equal(jit(j["win"]), j["index"].numpy())
    # a whole batch on the device: same map as the oracle on the uniforms the device drew, and the documented support
    jit = loader.Jitter(0.12)
    gen = torch.Generator(device="cuda").manual_seed(5)
    idx = jit.batch(16, 117, generator=gen).cpu().numpy()
    gen = torch.Generator(device="cuda").manual_seed(5)
    u = torch.rand(16, 115, dtype=torch.float64, device="cuda", generator=gen).cpu().numpy()
    for b in range(16):
        assert np.array_equal(idx[b], lo.jitter_from_uniforms(u[b], 117, 0.12))
    d = idx - np.arange(117)[None]
    assert d.min() == -1 and d.max() == 1
    assert abs((d[:, 2:] != 0).mean() - 0.24) < 0.03


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32])
def test_mfcc_batch_matches_the_oracle(dtype):
    from aewn import loader
    import loader_oracle as lo
    rs = np.random.RandomState(3)
    B, n = 3, 18470                       # a cfg2-sized window: 16384 + receptive field + MFCC wings
    t = np.arange(n)
    audio = 0.4 * np.sin(2 * np.pi * 220.0 * t / 16000)[None] * rs.rand(B, 1) + 0.1 * rs.randn(B, n)
    codes = lo.mu_encode_np(np.clip(audio, -1, 1).astype(np.float32), 256)
    wav = codes.astype(dtype) if dtype != np.float32 else codes.astype(np.float32)
    pw = loader.ProcessWav()
    got = pw.batch(torch.from_numpy(wav).cuda()).cpu().numpy()
    assert got.shape == (B, 39, pw.n_frames(n)[1])
    for b in range(B):
        ref = lo.process_wav(wav[b])
        err = np.abs(got[b] - ref)
        assert err[:13].max() < 2e-3, (b, err[:13].max())           # cepstra
        assert err[13:].max() < 2e-3, (b, err[13:].max())           # first and second derivatives
    # the reference's host contract: one item in, numpy out
    one = pw(wav[0])
    assert one.shape == got[0].shape and np.abs(one - got[0]).max() == 0.0


def test_mfcc_edges_short_input_and_silence():
    from aewn import loader
    import loader_oracle as lo
    pw = loader.ProcessWav()
    # shortest input scipy's derivative filter accepts (9 frames) and a ragged length
    for n in (1720, 1999):
        wav = np.random.RandomState(n).randint(0, 256, n).astype(np.uint8)
        got = pw.batch(torch.from_numpy(wav).cuda().view(1, -1))[0].cpu().numpy()
        assert np.abs(got - lo.process_wav(wav)).max() < 2e-3, n
    with pytest.raises(ValueError):
        pw.batch(torch.zeros(1, 1500, dtype=torch.uint8, device="cuda"))
    # digital silence (a constant code): everything below the clamp, finite output, zero derivatives
    wav = np.full(4000, 128, np.uint8)
    got = pw.batch(torch.from_numpy(wav).cuda().view(1, -1))[0].cpu().numpy()
    ref = lo.process_wav(wav)
    assert np.isfinite(got).all() and np.abs(got - ref).max() < 2e-3


def test_collate_returns_the_reference_tuple_on_the_device():
    """data.Collate.__call__, data.py:223-240: (wav, mel, voice, jitter, position) for a training batch."""
    from aewn import loader
    rs = np.random.RandomState(0)
    items = [((rs.randint(0, 256, 6000).astype(np.uint8), v), 3, 17) for v in (4, 1, 2)]     # ((snd, voice), epoch, step)
    col = loader.Collate(loader.ProcessWav(), loader.Jitter(0.12), train_mode=True)
    wav, mel, voice, jitter, position = col(items)
    assert wav.shape == (3, 6000) and wav.dtype == torch.float32 and wav.is_cuda
    assert mel.shape == (3, 39, 36) and voice.tolist() == [4, 1, 2]
    assert jitter.shape == (3, 36) and jitter.dtype == torch.long
    assert position.tolist() == [3, 17]
