"""Drop-in for src/feature_extraction.py (same names, arguments, return values, errors);
arithmetic in libdspfront.so."""
import numpy as np

from dsp_audioreclabs_b200 import batch as _b
from src.audio_processing import (   # noqa: F401  (re-exported like the reference, :5-9)
    compute_short_time_energy,
    compute_short_time_magnitude,
    compute_zero_crossing_rate,
)

_STATS = ('mean', 'std', 'max', 'min', 'median')
_FEATS = ('energy', 'magnitude', 'zcr')


def extract_frame_features(frames):
    """src/feature_extraction.py:12-43 -> {'energy', 'magnitude', 'zcr'} float64 sequences."""
    if len(frames) == 0:
        raise ValueError("No frames provided for feature extraction.")
    cached = getattr(frames, '_dsp_features', None)
    if cached is not None:
        return {k: cached[k].copy() for k in _FEATS}
    e, m, z, _ = _b.frame_features(np.asarray(frames), want_stats=False)
    return {'energy': e, 'magnitude': m, 'zcr': z}


def compute_statistics(sequence):
    """src/feature_extraction.py:46-62."""
    s = _b.sequence_stats(sequence)
    return dict(zip(_STATS, (s[0], s[1], s[2], s[3], s[4])))


def extract_statistical_features(frame_features):
    """src/feature_extraction.py:65-88 -> (15-vector, names)."""
    vec, names = [], []
    for ft in _FEATS:
        st = compute_statistics(frame_features[ft])
        for k in _STATS:
            vec.append(st[k])
            names.append(f'{ft}_{k}')
    return np.array(vec), names


def extract_features_from_frames(frames, method='statistical', use_only_energy_zcr=False):
    """src/feature_extraction.py:91-132."""
    if method not in ('statistical', 'sequence'):
        if len(frames) == 0:
            raise ValueError("No frames provided for feature extraction.")
        raise ValueError(f"不支持的特征提取方法: {method}")
    if len(frames) == 0:
        raise ValueError("No frames provided for feature extraction.")
    cached = getattr(frames, '_dsp_features', None)
    if method == 'statistical':
        names = [f'{f}_{s}' for f in _FEATS for s in _STATS]
        if cached is not None:
            return cached['stats'].copy(), names
        _, _, _, st = _b.frame_features(np.asarray(frames), want_stats=True)
        return st, names
    ff = extract_frame_features(frames)
    keys = ('energy', 'zcr') if use_only_energy_zcr else _FEATS
    return np.stack([ff[k] for k in keys], axis=1), None


def pad_or_truncate_sequence(sequence, target_length):
    """src/feature_extraction.py:135-154 (pure data movement)."""
    current_length = len(sequence)
    if current_length < target_length:
        padding = np.zeros((target_length - current_length, sequence.shape[1]))
        return np.vstack([sequence, padding])
    return sequence[:target_length]


def normalize_features(features, mean=None, std=None):
    """src/feature_extraction.py:157-181 -> (normalized, mean, std)."""
    return _b.zscore(np.asarray(features, dtype=np.float64), mean, std)
