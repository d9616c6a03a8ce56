"""GPU versions of the reference's loader arithmetic: drop-in ``ProcessWav`` (mfcc.py:27-76),
``Jitter`` (jitter.py:3-33), ``Collate`` (data.py:217-240) and the mu-law codec (util.py:62-96), on csrc/loader.cu.

The reference runs librosa's MFCC once per item inside ``Collate`` on the host; here a whole batch of windows is turned
into (B, 3 * n_mfcc, frames) features by three small launches on the device that already holds the window.  The tables the
kernels need (DFT twiddles, periodic Hann window, Slaney mel filterbank, orthonormal DCT-II, Savitzky-Golay derivative
rows) are built here with numpy from the published definitions librosa 0.7 / scipy use; oracle/loader_oracle.py restates
the same pipeline on scipy's own primitives and tests/test_gpu_loader.py compares the two."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L
from . import geometry as vconv
from . import ops
from .wavenet import _require_cuda

_DTYPES = {torch.uint8: 0, torch.int16: 1, torch.int32: 2, torch.float32: 3}


# ------------------------------------------------------------------------------------------------ mu-law (util.py:62-96)
def _mu_encode(x, n_quanta, torch_round):
    _require_cuda(x)
    xf = x.detach().to(torch.float32).contiguous()
    out = torch.empty(xf.shape, dtype=torch.int32, device=xf.device)
    L.check(L.lib().aewn_mu_encode(C.c_void_p(xf.data_ptr()), C.c_longlong(xf.numel()), C.c_int(n_quanta),
                                   C.c_int(1 if torch_round else 0), C.c_void_p(out.data_ptr()), ops._stream()),
            "aewn_mu_encode")
    return out


def mu_encode_torch(x, n_quanta):
    """util.mu_encode_torch, util.py:81-86 (rounds, returns int64)."""
    return _mu_encode(x, n_quanta, True).to(torch.long)


def mu_encode_np(x, n_quanta):
    """util.mu_encode_np, util.py:62-67, for a CUDA tensor of float32 audio (truncates, returns int32)."""
    return _mu_encode(x, n_quanta, False)


def mu_decode_torch(quant, n_quanta):
    """util.mu_decode_torch, util.py:88-96."""
    _require_cuda(quant)
    q = quant.detach().to(torch.int32).contiguous()
    out = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    L.check(L.lib().aewn_mu_decode(C.c_void_p(q.data_ptr()), C.c_longlong(q.numel()), C.c_int(n_quanta),
                                   C.c_void_p(out.data_ptr()), ops._stream()), "aewn_mu_decode")
    return out


# ------------------------------------------------------------------------------------------------ jitter (jitter.py:3-33)
class Jitter(object):
    """jitter.Jitter.  ``__call__(win_size)`` keeps the reference's host contract: it consumes win_size - 2 draws of
    numpy's global RandomState (what `np.random.choice(..., 1, False, p)` consumes, one `random_sample()` per step) and
    returns the same int array; ``batch(B, win_size)`` draws on the device instead and returns a (B, win_size) int64
    CUDA tensor for a whole batch."""

    def __init__(self, replace_prob):
        super(Jitter, self).__init__()
        p, s = replace_prob, (1 - 2 * replace_prob)
        self.replace_prob = replace_prob
        self.cond2d = np.tile([p, s, p], 9).reshape(3, 3, 3)
        self.cond2d[2][1] = [0, s / (p + s), p / (p + s)]

    def _indices(self, u, B, win_size):
        out = torch.empty(B, win_size, dtype=torch.long, device=u.device)
        L.check(L.lib().aewn_jitter_indices(C.c_void_p(u.data_ptr() if u.numel() else None), C.c_int(B), C.c_int(win_size),
                                            C.c_double(self.replace_prob), C.c_void_p(out.data_ptr()), ops._stream()),
                "aewn_jitter_indices")
        return out

    def __call__(self, win_size, device="cuda"):
        u = torch.from_numpy(np.random.random_sample(max(win_size - 2, 0))).to(device)
        return self._indices(u.view(1, -1), 1, win_size)[0].to(torch.int32).cpu().numpy()

    def batch(self, B, win_size, device="cuda", generator=None):
        u = torch.rand(B, max(win_size - 2, 0), dtype=torch.float64, device=device, generator=generator)
        return self._indices(u, B, win_size)


# ------------------------------------------------------------------------------------------------ MFCC (mfcc.py:27-76)
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels):
    """Triangular filters on the Slaney mel scale, each normalised to unit area (librosa.filters.mel defaults)."""
    freqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - freqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w


def dct_matrix(n_out, n_in):
    """Rows of the orthonormal DCT-II (scipy.fftpack.dct(type=2, norm='ortho'))."""
    k = np.arange(n_out)[:, None]
    n = np.arange(n_in)[None, :]
    d = np.sqrt(2.0 / n_in) * np.cos(np.pi * k * (2 * n + 1) / (2.0 * n_in))
    d[0] = np.sqrt(1.0 / n_in)
    return d.astype(np.float32)


def savgol_rows(order, width=9):
    """Savitzky-Golay derivative filter of window `width`, polynomial order = derivative order = `order`, as 9 rows over
    9 consecutive samples: rows 0-3 the left edge (the polynomial fitted to the first `width` samples, differentiated and
    evaluated at positions 0-3: scipy's mode='interp'), row 4 the interior filter (evaluated at the centre), rows 5-8 the
    right edge."""
    x = np.arange(width, dtype=np.float64)
    V = np.vander(x, order + 1, increasing=True)             # V[i, p] = x_i ** p
    P = np.linalg.pinv(V)                                    # coefficients c = P @ y
    rows = np.zeros((width, width))
    for pos in range(width):
        dv = np.zeros(order + 1)
        for p in range(order, order + 1):                    # d^order/dx^order of x**p, p <= order: only p == order survives
            dv[p] = math.factorial(p)
        rows[pos] = dv @ P
    return rows.astype(np.float32)


class ProcessWav(object):
    """mfcc.ProcessWav.  ``__call__(wav)`` takes what the reference takes (a 1-D numpy array of samples) and returns the
    (3 * n_mfcc, frames) numpy array; ``batch(wav)`` takes a (B, L) CUDA tensor (uint8 / int16 / int32 codes as stored in
    the dat file, data.py:36-41, or float32) and returns a (B, 3 * n_mfcc, frames) float32 CUDA tensor."""

    def __init__(self, sample_rate=16000, win_sz=400, hop_sz=160, n_mels=80, n_mfcc=13, name=None):
        self.sample_rate = sample_rate
        self.window_sz = win_sz
        self.hop_sz = hop_sz
        self.n_mels = n_mels
        self.n_mfcc = n_mfcc
        self.n_out = n_mfcc * 3
        self.vc = vconv.VirtualConv(filter_info=self.window_sz, stride=self.hop_sz, parent=None, name=name)
        adj = 1 if self.window_sz % 2 == 0 else 0
        adj_l = self.vc.l_wing_sz + adj
        self.left_pad = adj_l % self.hop_sz
        self.trim_left = adj_l // self.hop_sz
        self.trim_right = self.vc.r_wing_sz // self.hop_sz
        self._tables = {}

    def n_frames(self, n_samples):
        """(frames librosa computes, frames kept) for an input of n_samples (mfcc.py:60-71)."""
        n_all = 1 + (n_samples + self.left_pad) // self.hop_sz
        return n_all, n_all - self.trim_left - self.trim_right

    def _dev_tables(self, device):
        key = str(device)
        if key not in self._tables:
            n = self.window_sz
            ang = 2.0 * np.pi * np.arange(n) / n
            tw = np.stack([np.cos(ang), np.sin(ang)], axis=1)
            win = 0.5 - 0.5 * np.cos(ang)                     # periodic Hann (scipy get_window('hann', n, fftbins=True))
            sg = np.stack([savgol_rows(1), savgol_rows(2)])
            t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt)
            self._tables[key] = dict(tw=t(tw, torch.float64), win=t(win, torch.float64),
                                     mel=t(mel_filterbank(self.sample_rate, n, self.n_mels), torch.float32),
                                     dct=t(dct_matrix(self.n_mfcc, self.n_mels), torch.float32), sg=t(sg, torch.float32))
        return self._tables[key]

    def batch(self, wav):
        _require_cuda(wav)
        if wav.dim() != 2 or wav.dtype not in _DTYPES:
            raise ValueError("ProcessWav.batch expects a (B, L) uint8 / int16 / int32 / float32 CUDA tensor")
        wav = wav if wav.stride(1) == 1 else wav.contiguous()
        B, n = wav.shape
        n_all, n_keep = self.n_frames(n)
        if n_keep < 9:
            raise ValueError(f"{n} samples give {n_keep} frames; the derivative filters need at least 9 "
                             "(librosa.feature.delta raises for fewer)")
        tb = self._dev_tables(wav.device)
        d = L.MfccDesc()
        d.n_fft, d.hop, d.n_mels, d.n_mfcc = self.window_sz, self.hop_sz, self.n_mels, self.n_mfcc
        d.left_pad, d.trim_left, d.n_frames_all, d.n_frames, d.top_db = self.left_pad, self.trim_left, n_all, n_keep, 80.0
        d.twiddle, d.window, d.melw, d.dctm, d.sg = (tb["tw"].data_ptr(), tb["win"].data_ptr(), tb["mel"].data_ptr(),
                                                      tb["dct"].data_ptr(), tb["sg"].data_ptr())
        work = torch.empty(B * self.n_mels * n_all, dtype=torch.float32, device=wav.device)
        wmax = torch.empty(B, dtype=torch.int32, device=wav.device)
        out = torch.empty(B, self.n_out, n_keep, dtype=torch.float32, device=wav.device)
        L.check(L.lib().aewn_mfcc(C.c_void_p(wav.data_ptr()), C.c_int(_DTYPES[wav.dtype]), C.c_longlong(wav.stride(0)), C.c_int(B),
                                  C.c_int(n), C.byref(d), C.c_void_p(work.data_ptr()), C.c_void_p(wmax.data_ptr()),
                                  C.c_void_p(out.data_ptr()), C.c_longlong(out.stride(0)), C.c_longlong(out.stride(1)),
                                  ops._stream()), "aewn_mfcc")
        return out

    def __call__(self, wav, device="cuda"):
        a = np.asarray(wav)
        if a.dtype == np.uint8:
            t = torch.from_numpy(a)
        elif a.dtype == np.int16:
            t = torch.from_numpy(a)
        elif np.issubdtype(a.dtype, np.integer):
            t = torch.from_numpy(a.astype(np.int32))
        else:
            t = torch.from_numpy(a.astype(np.float32))
        return self.batch(t.to(device).view(1, -1))[0].cpu().numpy()


class Collate():
    """data.Collate, data.py:217-240, with the MFCC and the jitter indices computed on `device` for the whole batch.
    Returns the reference's tuple; wav, mel, voice and jitter are CUDA tensors (the reference moves them there in
    chassis.py:146-150 after collating on the host)."""

    def __init__(self, mfcc, jitter, train_mode, device="cuda"):
        self.train_mode = train_mode
        self.mfcc = mfcc
        self.jitter = jitter
        self.device = device

    def __call__(self, batch):
        data = [b[0] for b in batch]
        position = torch.tensor(batch[-1][1:])
        codes = torch.stack([torch.from_numpy(np.ascontiguousarray(d[0])) for d in data]).to(self.device)
        if codes.dtype not in _DTYPES:
            codes = codes.to(torch.int32)
        wav = codes.float()
        mel = self.mfcc.batch(codes)
        voice = torch.tensor([d[1] for d in data]).long().to(self.device)
        jitter = self.jitter.batch(len(data), mel.size()[2], device=self.device)
        if self.train_mode:
            return wav, mel, voice, jitter, position
        paths = [b[0][2] for b in batch]
        return wav, mel, voice, jitter, paths, position
