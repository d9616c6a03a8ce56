"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy / scipy) of the reference's data-loader arithmetic:
mu-law codec (util.py:62-96), the jitter index generator (jitter.py:3-33) and the MFCC + delta features of mfcc.py:39-76.

Pinning.
* mu-law and jitter are pinned against the reference's own functions (importable, pure numpy / torch) by
  tests/test_loader_oracle.py when /root/reference is present, and by the golden vectors that oracle/make_golden.py loader
  writes from them.
* MFCC: **parity unpinned**.  mfcc.py calls `librosa.feature.mfcc` / `librosa.feature.delta`; librosa is a third-party
  dependency that is absent from the reference tree and from this image, and the reference pins no version (it uses
  `librosa.output.write_wav`, chassis.py:342, which librosa removed in 0.8, so the 0.6/0.7 series is meant).  What follows
  restates the published algorithm of librosa 0.7.2 (`feature/spectral.py: mfcc, melspectrogram`, `core/spectrum.py: stft,
  _spectrogram, power_to_db`, `filters.py: mel, get_window`, `feature/utils.py: delta`) on the SAME scipy primitives librosa
  itself calls (`scipy.signal.get_window`, `scipy.fftpack.dct`, `scipy.signal.savgol_filter`), anchored on the reference's
  call site (mfcc.py:56-75: n_fft = win_sz, hop_length = hop_sz, n_mels, n_mfcc, defaults otherwise) and on its own
  output-size formula (mfcc.py:60-69).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this file."""
import numpy as np
from scipy.fftpack import dct
from scipy.signal import get_window, savgol_filter


# ------------------------------------------------------------------------------------------------ mu-law (util.py:62-96)
def mu_encode_np(x, n_quanta):
    """util.py:62-67 (numpy keeps the dtype of x: float32 audio is encoded in float32 arithmetic)."""
    mu = n_quanta - 1
    amp = np.sign(x) * np.log1p(mu * np.abs(x)) / np.log1p(mu)
    quant = (amp + 1) * 0.5 * mu + 0.5
    return quant.astype(np.int32)


def mu_decode_np(quant, n_quanta):
    """util.py:70-78."""
    mu = n_quanta - 1
    qf = quant.astype(np.float32)
    inv_mu = 1.0 / mu
    a = (2 * qf - 1) * inv_mu - 1
    return np.sign(a) * ((1 + mu) ** np.fabs(a) - 1) * inv_mu


# ------------------------------------------------------------------------------------------------ jitter (jitter.py:3-33)
def jitter_probs(replace_prob):
    """jitter.py:13-19.  __call__ indexes the table as cond2d[p1][p1] (jitter.py:29; p2 is read but not used), so the row
    that forbids three-in-a-row, cond2d[2][1], is never selected: every step draws from [p, 1 - 2p, p]."""
    p, s = replace_prob, 1 - 2 * replace_prob
    cond2d = np.tile([p, s, p], 9).reshape(3, 3, 3)
    cond2d[2][1] = [0, s / (p + s), p / (p + s)]
    return cond2d


def jitter_from_uniforms(u, win_size, replace_prob):
    """jitter.py:21-33 with the random draws made explicit: `np.random.choice([0, 1, 2], 1, False, pvec)` consumes ONE
    `random_sample()` x and returns searchsorted(cumsum(pvec) / cumsum(pvec)[-1], x, side='right') (numpy legacy
    RandomState.choice, replace=False branch, size 1).  ``u``: win_size - 2 uniforms in [0, 1)."""
    cond2d = jitter_probs(replace_prob)
    index = np.ones((win_size + 1), dtype=np.int32)
    for t in range(2, win_size):
        p1 = index[t - 1]
        cdf = np.cumsum(cond2d[p1][p1])
        cdf /= cdf[-1]
        index[t] = np.searchsorted(cdf, u[t - 2], side="right")
    index[win_size] = 1
    index += np.arange(-1, win_size)
    return index[:-1]


# ------------------------------------------------------------------------------------------------ MFCC (mfcc.py:27-76)
def hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin=0, fmax=sr/2, htk=False, norm=1) -> (n_mels, 1 + n_fft//2) float32."""
    fmax = sr / 2.0
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2, endpoint=True)
    mel_f = mel_to_hz_slaney(np.linspace(hz_to_mel_slaney(0.0), hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def power_spectrogram(y, n_fft, hop):
    """|librosa.stft(y, n_fft, hop, window='hann', center=True, pad_mode='reflect')|**2 -> (1 + n_fft//2, n_frames).
    The FFT runs in double precision and is stored as complex64, like librosa's `stft_matrix`."""
    win = get_window("hann", n_fft, fftbins=True)
    yp = np.pad(np.asarray(y, dtype=np.float64), n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop
    frames = np.stack([yp[i * hop:i * hop + n_fft] for i in range(n_frames)], axis=1)          # (n_fft, n_frames)
    spec = np.fft.rfft(win[:, None] * frames, axis=0).astype(np.complex64)
    return np.abs(spec) ** 2.0


def power_to_db(S, amin=1e-10, top_db=80.0):
    """librosa.power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0)."""
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, 1.0))
    return np.maximum(log_spec, log_spec.max() - top_db)


def librosa_mfcc(y, sr, n_fft, hop, n_mels, n_mfcc):
    """librosa.feature.mfcc(y=y, sr=sr, n_fft=n_fft, hop_length=hop, n_mels=n_mels, n_mfcc=n_mfcc)."""
    mel = np.dot(mel_filterbank(sr, n_fft, n_mels), power_spectrogram(y, n_fft, hop))
    return dct(power_to_db(mel), axis=0, type=2, norm="ortho")[:n_mfcc]


def delta(data, order):
    """librosa.feature.delta(data, width=9, order=order, axis=-1, mode='interp')."""
    return savgol_filter(data, 9, deriv=order, polyorder=order, axis=-1, mode="interp")


def wings(win_sz):
    """vconv.VirtualConv(filter_info=win_sz) wing sizes (vconv.py: l = (f - 1) // 2, r = f - 1 - l)."""
    left = (win_sz - 1) // 2
    return left, win_sz - 1 - left


def process_wav(wav, sample_rate=16000, win_sz=400, hop_sz=160, n_mels=80, n_mfcc=13):
    """mfcc.ProcessWav.__call__, mfcc.py:39-76: left-pad, MFCC, trim the window positions that do not lie fully inside the
    input, append first and second derivatives.  Returns (3 * n_mfcc, n_frames) float64."""
    l_wing, r_wing = wings(win_sz)
    adj_l = l_wing + (1 if win_sz % 2 == 0 else 0)
    left_pad, trim_left, trim_right = adj_l % hop_sz, adj_l // hop_sz, r_wing // hop_sz
    wav_pad = np.concatenate((np.zeros(left_pad), np.asarray(wav)), axis=0)
    m = librosa_mfcc(wav_pad, sample_rate, win_sz, hop_sz, n_mels, n_mfcc)
    n_pos = wav_pad.shape[0] + (1 if win_sz % 2 == 0 else 0)
    assert m.shape[1] == n_pos // hop_sz + (1 if n_pos % hop_sz > 0 else 0)        # mfcc.py:60-69
    m = m[:, trim_left:-trim_right or None]
    return np.concatenate((m, delta(m, 1), delta(m, 2)), axis=0)
