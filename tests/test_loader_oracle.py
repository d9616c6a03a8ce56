"""The loader oracle (oracle/loader_oracle.py) against what exists of the reference for it, and the table builders of
aewn/loader.py against scipy's own primitives (CPU only).

mu-law and jitter: the reference's functions are importable (pure numpy / torch): golden vectors written from them by
`oracle/make_golden.py loader`, plus a live comparison when /root/reference is present.  MFCC: librosa is absent from the
reference tree and this image ("parity unpinned", see the oracle's header); what CAN be pinned here is that the oracle's and
the product's tables agree with scipy (get_window, dct, savgol_filter), which is what librosa calls."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import loader_oracle as lo  # noqa: E402


def test_mu_law_and_jitter_oracle_match_the_reference_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "loader.pt"))
    x = g["x"].numpy()
    assert np.array_equal(lo.mu_encode_np(x, 256), g["enc_np"].numpy())
    q = np.arange(256, dtype=np.int32)
    assert np.array_equal(lo.mu_decode_np(q, 256).astype(np.float32), g["dec_np"].numpy())
    for j in g["jitter"]:
        got = lo.jitter_from_uniforms(j["uniforms"].numpy(), j["win"], j["prob"])
        assert np.array_equal(got, j["index"].numpy()), j["seed"]
        # the reference's contract (jitter.py:4-8): output element t is one of t - 1, t, t + 1
        d = j["index"].numpy() - np.arange(j["win"])
        assert d.min() >= -1 and d.max() <= 1 and d[0] == 0 and d[1] == 0


@pytest.mark.needs_reference
def test_oracle_matches_the_live_reference_functions():
    sys.path.insert(0, "/root/reference")
    import jitter as ref_jitter
    import util as ref_util
    rs = np.random.RandomState(1)
    x = np.clip(rs.randn(10000) * 0.3, -1, 1).astype(np.float32)
    assert np.array_equal(ref_util.mu_encode_np(x, 256), lo.mu_encode_np(x, 256))
    assert np.array_equal(ref_util.mu_decode_np(np.arange(256), 256), lo.mu_decode_np(np.arange(256), 256))
    for seed in range(4):
        np.random.seed(seed)
        ref = ref_jitter.Jitter(0.12)(77)
        np.random.seed(seed)
        assert np.array_equal(ref, lo.jitter_from_uniforms(np.random.random_sample(75), 77, 0.12))
    assert np.array_equal(ref_jitter.Jitter(0.2).cond2d, lo.jitter_probs(0.2))


def test_product_tables_match_scipy_and_the_oracle():
    from scipy.fftpack import dct
    from scipy.signal import get_window, savgol_filter
    from aewn import loader
    # periodic Hann window
    n = 400
    ang = 2.0 * np.pi * np.arange(n) / n
    assert np.allclose(0.5 - 0.5 * np.cos(ang), get_window("hann", n, fftbins=True), atol=1e-15)
    # mel filterbank: product builder == oracle builder, rows are non-negative triangles of area 2 / bandwidth * ...
    w = loader.mel_filterbank(16000, 400, 80)
    assert np.array_equal(w, lo.mel_filterbank(16000, 400, 80))
    assert w.shape == (80, 201) and (w >= 0).all() and (w.sum(1) > 0).all()
    # DCT-II, orthonormal
    rs = np.random.RandomState(0)
    v = rs.randn(80, 7)
    assert np.allclose(loader.dct_matrix(13, 80).astype(np.float64) @ v, dct(v, axis=0, type=2, norm="ortho")[:13], atol=1e-5)
    # Savitzky-Golay derivative rows == scipy.signal.savgol_filter(window 9, mode 'interp') on a random signal
    sig = rs.randn(3, 40)
    for order in (1, 2):
        rows = loader.savgol_rows(order).astype(np.float64)
        ref = savgol_filter(sig, 9, deriv=order, polyorder=order, axis=-1, mode="interp")
        got = np.empty_like(sig)
        F = sig.shape[1]
        for j in range(F):
            if j < 4:
                r, s0 = j, 0
            elif j >= F - 4:
                r, s0 = 9 - (F - j), F - 9
            else:
                r, s0 = 4, j - 4
            got[:, j] = sig[:, s0:s0 + 9] @ rows[r]
        assert np.allclose(got, ref, atol=1e-6), order


def test_process_wav_oracle_geometry_follows_the_reference_formula():
    """mfcc.py:44-71: left pad 40, one frame trimmed on each side for window 400 / hop 160; frames = valid window positions."""
    from aewn import loader
    pw = loader.ProcessWav()
    assert (pw.left_pad, pw.trim_left, pw.trim_right) == (40, 1, 1)
    for n in (4000, 6000, 18470):
        wav = (np.random.RandomState(n).randint(0, 256, n)).astype(np.uint8)
        m = lo.process_wav(wav)            # asserts librosa's output-size formula (mfcc.py:60-69) internally
        assert m.shape == (39, (n + 40) // 160 - 1), (n, m.shape)
        assert m.shape[1] == pw.n_frames(n)[1]
        assert np.isfinite(m).all()
    # an input that ends on a window edge (the sizes the reference's geometry produces, doc/padding_notes.txt): the kept
    # frames are exactly the window positions that lie inside the input
    assert pw.n_frames(6000)[1] == (6000 - 400) // 160 + 1
