"""Drop-in for src/audio_processing.py: same names, arguments, return values and errors; the
arithmetic runs in the CUDA library (libdspfront.so) -- there is no NumPy fallback.

Every function cites the reference lines it stands in for.  `process_audio_file` keeps the
16-bit PCM as integers and uses the fused kernel; the per-call functions on float arrays use
the float64 replay kernel, which restates NumPy's operation order and is bit-identical to it.
"""
import numpy as np

from dsp_audioreclabs_b200 import batch as _b, wavio as _wavio


class FrameArray(np.ndarray):
    """The (n_frames, frame_length) ndarray `process_audio_file` returns, carrying the per-frame
    features the fused kernel already computed so that `extract_features_from_frames(frames)`
    (the very next call in every caller: run_experiments.py:102, train_model.py:88,
    compare_feature_methods.py:56,96) does not recompute them.  Any view, slice or arithmetic
    result drops the cache."""

    _dsp_features = None

    def __array_finalize__(self, obj):
        self._dsp_features = None


def read_wav_pcm(filepath):
    """WAV -> (pcm array as stored, sample_rate, n_channels): the library's native chunk walk
    (csrc/wavio.cpp), which accepts and rejects what `wave.open` does, and the same ValueError
    for other sample widths as load_wav (src/audio_processing.py:21-40)."""
    return _wavio.read_wav_pcm(filepath)


def load_wav(filepath):
    """src/audio_processing.py:9-46 -- float64 in [-1, 1] and the sample rate.  The conversion
    itself is a cast and one exact division per sample; it is done here on the host because
    its only use is to hand the caller a float array (the fused path never materialises it).
    8-bit data reproduces the reference's uint8 wrap-around (`audio_data - 128` on a uint8
    array, :33-34)."""
    pcm, sample_rate, n_channels = read_wav_pcm(filepath)
    if pcm.dtype == np.uint8:
        audio_data = (pcm - 128) / 128.0
    else:
        audio_data = pcm / 32768.0
    if n_channels == 2:
        audio_data = audio_data.reshape(-1, 2).mean(axis=1)
    return audio_data, sample_rate


def remove_dc(audio_data):
    """src/audio_processing.py:49-59."""
    return _b.preprocess(audio_data, 0)


def normalize_audio(audio_data):
    """src/audio_processing.py:62-75."""
    return _b.preprocess(audio_data, 1)


def preprocess(audio_data):
    """src/audio_processing.py:78-90."""
    return _b.preprocess(audio_data, 2)


def _one_frame(frame):
    f = np.ascontiguousarray(frame, dtype=np.float64).reshape(1, -1)
    if f.shape[1] == 0:
        return 0.0, 0.0, 0.0
    e, m, z, _ = _b.frame_features(f, want_stats=False)
    return e[0], m[0], z[0]


def compute_short_time_energy(frame):
    """src/audio_processing.py:93-103."""
    return _one_frame(frame)[0]


def compute_short_time_magnitude(frame):
    """src/audio_processing.py:106-116."""
    return _one_frame(frame)[1]


def compute_zero_crossing_rate(frame):
    """src/audio_processing.py:119-132."""
    return _one_frame(frame)[2]


def endpoint_detection(audio_data, frame_length, frame_shift,
                       energy_high_ratio=0.5, energy_low_ratio=0.1,
                       zcr_threshold_ratio=1.5):
    """src/audio_processing.py:135-275 -> (start_point, end_point, energy_list, zcr_list)."""
    return _b.endpoint_detection(audio_data, frame_length, frame_shift, energy_high_ratio,
                                 energy_low_ratio, zcr_threshold_ratio)


def create_window(window_type, length):
    """src/audio_processing.py:278-296 (ValueError for an unknown window)."""
    if window_type not in ('rectangular', 'hamming', 'hanning'):
        raise ValueError(f"不支持的窗函数类型: {window_type}")
    return _b.window(window_type, length)


def frame_signal(audio_data, frame_length, frame_shift, window_type='hamming'):
    """src/audio_processing.py:299-333 -> (n_frames, frame_length) float64."""
    if window_type not in ('rectangular', 'hamming', 'hanning'):
        raise ValueError(f"不支持的窗函数类型: {window_type}")
    return _b.frame_signal(audio_data, frame_length, frame_shift, window_type)


def process_audio_file(filepath, frame_length, frame_shift,
                       window_type='hamming',
                       do_endpoint_detection=True,
                       energy_high_ratio=0.5,
                       energy_low_ratio=0.1,
                       zcr_threshold_ratio=1.5):
    """src/audio_processing.py:336-396 -> (frames, sample_rate, metadata).

    One fused front-end launch on the PCM samples gives the endpoints, the EPD lists and the
    per-frame features; the dense `frames` matrix callers expect is produced by the framing
    kernel from the trimmed, pre-processed signal."""
    if window_type not in ('rectangular', 'hamming', 'hanning'):
        raise ValueError(f"不支持的窗函数类型: {window_type}")
    pcm, sample_rate, n_channels = read_wav_pcm(filepath)
    n = pcm.size // max(n_channels, 1)
    if n == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    res = _b.frontend_batch(pcm, np.array([0, pcm.size]), frame_length, frame_shift, window_type,
                            do_endpoint_detection, energy_high_ratio, energy_low_ratio,
                            zcr_threshold_ratio, channels=n_channels, emit_epd_lists=True)
    metadata = {'original_length': n, 'sample_rate': sample_rate}
    start, end = int(res.start[0]), int(res.end[0])
    if do_endpoint_detection:
        el, zl = res.epd_lists(0)
        metadata.update({'start_point': start, 'end_point': end,
                         'energy_list': np.asarray(el, dtype=np.float64),
                         'zcr_list': np.asarray(zl, dtype=np.float64),
                         'segmented_length': end - start})
    if end - start <= 0:
        raise ValueError("No audio remaining after preprocessing and endpoint detection.")
    audio = preprocess(load_wav(filepath)[0])[start:end]
    frames = frame_signal(audio, frame_length, frame_shift, window_type).view(FrameArray)
    e, m, z = res.frames(0)
    frames._dsp_features = {'energy': e.astype(np.float64), 'magnitude': m.astype(np.float64),
                            'zcr': z.astype(np.float64), 'stats': res.stats[0].astype(np.float64)}
    metadata['n_frames'] = len(frames)
    return frames, sample_rate, metadata
