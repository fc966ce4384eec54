"""CPU: the self-oracle of the MFCC + DTW variant (oracle/mfcc_dtw_oracle.py; NOT in the reference, parity
unpinned) is internally consistent, and the product's host-side tables equal the oracle's independent ones."""
import functools

import numpy as np


def test_host_tables_match_the_oracle():
    from dsp_audioreclabs_b200 import mfcc_dtw
    from oracle import mfcc_dtw_oracle as mo
    for n_mels, n_fft in ((26, 2048), (40, 1024), (13, 512)):
        fb = mo.mel_filterbank(n_mels, n_fft, 44100)
        assert np.allclose(mfcc_dtw.mel_filterbank(n_mels, n_fft, 44100), fb, atol=1e-7)
        assert np.all(fb >= 0) and np.all(fb.max(axis=1) > 0.5)           # every filter has a peak
    d = mfcc_dtw.dct_matrix(13, 26).astype(np.float64)
    assert np.allclose(d, mo.dct_matrix(13, 26), atol=1e-7) and np.allclose(d @ d.T, np.eye(13), atol=1e-6)


def test_dtw_oracle_equals_the_recurrence():
    from oracle import mfcc_dtw_oracle as mo
    rng = np.random.default_rng(0)
    for n, m in ((1, 1), (1, 6), (7, 1), (5, 4), (9, 13)):
        a, b = rng.standard_normal((n, 3)), rng.standard_normal((m, 3))
        d = np.sqrt(((a[:, None] - b[None]) ** 2).sum(2))

        @functools.lru_cache(None)
        def D(i, j):
            if i == 0 and j == 0:
                return d[0, 0]
            c = []
            if i > 0: c.append(D(i - 1, j))
            if j > 0: c.append(D(i, j - 1))
            if i > 0 and j > 0: c.append(D(i - 1, j - 1))
            return d[i, j] + min(c)
        assert mo.dtw_cost(a, b) == D(n - 1, m - 1)
    x = rng.standard_normal((8, 2))
    assert mo.dtw_cost(x, x) == 0.0 and mo.dtw_cost(x, np.repeat(x, 2, axis=0)) == 0.0     # warping absorbs repetition


def test_mfcc_oracle_properties():
    from oracle import mfcc_dtw_oracle as mo, synth
    pcm = synth.utterance_pcm(3, 26000, seed0=5)
    m1, s1, e1 = mo.mfcc_utterance(pcm)
    assert m1.shape[1] == 13 and len(m1) >= 1 and np.all(np.isfinite(m1))
    # peak normalisation makes the features invariant to the recording gain (up to int16 rounding of the halved signal)
    m2, s2, e2 = mo.mfcc_utterance((pcm // 2 * 2).astype(np.int16))
    m3, s3, e3 = mo.mfcc_utterance((pcm // 2).astype(np.int16))
    assert (s2, e2) == (s3, e3) and np.allclose(m2, m3, atol=1e-6)
