"""NumPy definition of the MFCC + DTW template-matching variant (TEST INFRASTRUCTURE ONLY).

**Self-oracle, parity unpinned.**  BASELINE config 5 / SURVEY.md section 8 rows a11 and f4 ask for an MFCC + DTW
variant "where present" in compare_feature_methods.py; it is NOT present in the reference
(`grep -ri "mfcc|dtw|mel|fft" /root/reference` finds nothing) and librosa is not in the image, so there is no
reference output to pin against.  This file therefore *defines* the variant from the published textbook
formulas, reusing the reference's own pre-processing, endpoint detection and framing for everything that
exists there (oracle/frontend_oracle.py), and the CUDA path (csrc/mfcc_dtw.cu) is checked against it.  Every
report of these numbers must say "self-oracle".

Definition (parameters in DEFAULTS):
  x      = preprocess(pcm / 32768)                       reference src/audio_processing.py:35-38,78-90
  seg    = x[start:end]                                   endpoints of the reference's endpoint_detection (:135-275)
  y[0]   = seg[0];  y[i] = seg[i] - 0.97 * seg[i-1]       pre-emphasis
  frames = frame_signal(y, 1102, 441, 'hamming')          the reference's framing rule (:299-333): last frame zero-padded
  P      = |rfft(frames, 2048)|^2 / 2048                  power spectrum
  mel    = P @ filterbank.T                               26 triangular filters, HTK mel scale 2595 log10(1 + f/700), 0 .. sr/2
  mfcc   = log(max(mel, 1e-10)) @ dct.T                   orthonormal DCT-II, first 13 coefficients
  dtw    = D[n-1][m-1],  D[i][j] = |a_i - b_j|_2 + min(D[i-1][j], D[i][j-1], D[i-1][j-1]),  D[0][0] = |a_0 - b_0|_2
  label  = majority of the k templates with the smallest dtw cost (ties: lower template index / smallest label)
"""
import numpy as np

from . import frontend_oracle as fo

DEFAULTS = dict(frame_length=1102, frame_shift=441, n_fft=2048, n_mels=26, n_ceps=13, pre_emphasis=0.97,
                log_floor=1e-10, sample_rate=44100)


def hz_to_mel(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel_to_hz(m):
    return 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def mel_filterbank(n_mels, n_fft, sample_rate):
    """(n_mels, n_fft//2 + 1) triangular filters with FFT-bin corner points floor((n_fft + 1) * hz / sr)."""
    pts = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sample_rate / 2.0), n_mels + 2))
    bins = np.floor((n_fft + 1) * pts / sample_rate).astype(np.int64)
    fb = np.zeros((n_mels, n_fft // 2 + 1))
    for m in range(1, n_mels + 1):
        lo, ce, hi = bins[m - 1], bins[m], bins[m + 1]
        for b in range(lo, ce):
            fb[m - 1, b] = (b - lo) / (ce - lo)
        for b in range(ce, hi):
            fb[m - 1, b] = (hi - b) / (hi - ce)
    return fb


def dct_matrix(n_ceps, n_mels):
    """Orthonormal DCT-II rows: c_k = s_k * sum_m x_m cos(pi k (2m + 1) / (2M)), s_0 = sqrt(1/M), s_k = sqrt(2/M)."""
    m = np.arange(n_mels)
    d = np.cos(np.pi * np.arange(n_ceps)[:, None] * (2 * m[None, :] + 1) / (2.0 * n_mels))
    d *= np.sqrt(2.0 / n_mels)
    d[0] *= np.sqrt(0.5)
    return d


def mfcc_segment(x, start, end, frame_length=1102, frame_shift=441, n_fft=2048, n_mels=26, n_ceps=13,
                 pre_emphasis=0.97, log_floor=1e-10, sample_rate=44100):
    """MFCC frames (F, n_ceps) of the pre-processed float64 signal x over [start, end)."""
    seg = np.asarray(x, dtype=np.float64)[start:end]
    if len(seg) == 0:
        return np.zeros((0, n_ceps))
    y = seg.copy()
    y[1:] = seg[1:] - pre_emphasis * seg[:-1]
    frames = fo.frame_signal(y, frame_length, frame_shift, "hamming")
    power = np.abs(np.fft.rfft(frames, n_fft, axis=1)) ** 2 / n_fft
    mel = power @ mel_filterbank(n_mels, n_fft, sample_rate).T
    return np.log(np.maximum(mel, log_floor)) @ dct_matrix(n_ceps, n_mels).T


def mfcc_utterance(pcm, **kw):
    """int16 PCM -> (mfcc (F, n_ceps), start, end): reference pre-processing and endpoints, then mfcc_segment."""
    p = dict(DEFAULTS); p.update(kw)
    x = fo.preprocess(fo.pcm_to_float(np.asarray(pcm)))
    start, end, _, _ = fo.endpoint_detection(x, p["frame_length"], p["frame_shift"])
    return mfcc_segment(x, start, end, **p), start, end


def dtw_cost(a, b):
    """Accumulated cost D[n-1][m-1] with Euclidean local distance and steps (1,0), (0,1), (1,1); anti-diagonal sweep."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    n, m = len(a), len(b)
    if n == 0 or m == 0:
        return np.inf
    d = np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(axis=2))
    D = np.full((n + 1, m + 1), np.inf)
    D[0, 0] = 0.0
    for s in range(n + m - 1):                        # cells with i + j == s
        i = np.arange(max(0, s - m + 1), min(n, s + 1))
        j = s - i
        D[i + 1, j + 1] = d[i, j] + np.minimum(np.minimum(D[i, j + 1], D[i + 1, j]), D[i, j])
    return D[n, m]


def dtw_matrix(queries, templates):
    return np.array([[dtw_cost(q, t) for t in templates] for q in queries])


def dtw_topk(queries, templates, k):
    """(idx[nq, k], cost[nq, k]) sorted by (cost, template index)."""
    c = dtw_matrix(queries, templates)
    idx = np.empty((len(queries), k), dtype=np.int64)
    cost = np.empty((len(queries), k))
    for i, row in enumerate(c):
        order = np.lexsort((np.arange(len(row)), row))[:k]
        idx[i], cost[i] = order, row[order]
    return idx, cost
