"""Batched replacement for the reference's per-file dataset loops:
`load_dataset` walks the class-per-directory tree exactly like
experiments/run_experiments.py:64-114 / train_model.py:56-98 (sorted class folders, hidden ones
skipped, glob('*.wav') per class), but decodes every file first and runs ONE fused front-end
launch over the whole set instead of one Python call chain per file.  `ablation_sweep` re-runs
the front end once per (frame_length, frame_shift) on the samples already decoded, which is
BASELINE config 4's "one GPU launch per configuration" (ablation_study.py:146-163 re-reads and
re-processes every WAV for every sweep value)."""
import os
from glob import glob

import numpy as np

from . import batch

FEATURE_NAMES = batch.FEATURE_NAMES


def read_wav_pcm(path):
    """(pcm, sample_rate, channels) with the reference's rules (src/audio_processing.py:21-40)."""
    import wave
    with wave.open(path, "rb") as w:
        ch, width, sr = w.getnchannels(), w.getsampwidth(), w.getframerate()
        raw = w.readframes(w.getnframes())
    if width == 1:
        return np.frombuffer(raw, dtype=np.uint8), sr, ch
    if width == 2:
        return np.frombuffer(raw, dtype=np.int16), sr, ch
    raise ValueError(f"unsupported sample width: {width}")


def decode_tree(data_dir):
    """-> (clips [(pcm, channels)], labels, class_names, paths); unreadable files are skipped like the
    reference's try/except does (run_experiments.py:109-111)."""
    class_names = sorted(d for d in os.listdir(data_dir)
                         if os.path.isdir(os.path.join(data_dir, d)) and not d.startswith("."))
    clips, labels, paths = [], [], []
    for ci, name in enumerate(class_names):
        for path in glob(os.path.join(data_dir, name, "*.wav")):
            try:
                pcm, _, ch = read_wav_pcm(path)
            except Exception:
                continue
            clips.append((pcm, ch))
            labels.append(ci)
            paths.append(path)
    return clips, np.array(labels, dtype=np.int64), class_names, paths


def features_for_clips(clips, frame_length, frame_shift, window_type="hamming", do_endpoint_detection=True,
                       energy_high_ratio=0.5, energy_low_ratio=0.1, zcr_threshold_ratio=1.5, ctx=None):
    """15-dim statistical features of every clip: one launch per (dtype, channels) group (16-bit mono
    files -- the normal case -- all go through the fused int16 kernel).  Returns (X[n,15] float64, ok[n])."""
    n = len(clips)
    X = np.zeros((n, 15), dtype=np.float64)
    ok = np.zeros(n, dtype=bool)
    groups = {}
    for i, (pcm, ch) in enumerate(clips):
        groups.setdefault((pcm.dtype.str, ch), []).append(i)
    for (_, ch), idx in groups.items():
        samples, offsets, lengths = batch.pack_aligned([clips[i][0] for i in idx])
        res = batch.frontend_batch(samples, offsets, frame_length, frame_shift, window_type, do_endpoint_detection,
                                   energy_high_ratio, energy_low_ratio, zcr_threshold_ratio, channels=ch,
                                   emit_frames=False, lengths=lengths, ctx=ctx)
        good = (res.status & 0xff) == 0
        X[idx] = res.stats
        ok[idx] = good
    return X, ok


def load_dataset(data_dir, frame_length, frame_shift, window_type="hamming", do_endpoint_detection=True,
                 energy_high_ratio=0.5, energy_low_ratio=0.1, zcr_threshold_ratio=1.5, ctx=None):
    """-> (X, y, class_names, feature_names): what SpeechRecognitionExperiment.load_dataset /
    train_model.load_dataset build, with failed files dropped."""
    clips, labels, class_names, _ = decode_tree(data_dir)
    X, ok = features_for_clips(clips, frame_length, frame_shift, window_type, do_endpoint_detection,
                               energy_high_ratio, energy_low_ratio, zcr_threshold_ratio, ctx)
    return X[ok], labels[ok], class_names, list(FEATURE_NAMES)


def ablation_sweep(data_dir, frame_configs, window_type="hamming", sample_rate=44100, ctx=None, **kw):
    """{(frame_length, frame_shift): (X, y)} for each configuration, decoding the WAVs once.
    `frame_configs` are sample counts, e.g. int(sample_rate * ms / 1000) as train_model.py:45-46."""
    clips, labels, class_names, _ = decode_tree(data_dir)
    out = {}
    for fl, fs in frame_configs:
        X, ok = features_for_clips(clips, int(fl), int(fs), window_type, ctx=ctx, **kw)
        out[(int(fl), int(fs))] = (X[ok], labels[ok])
    return out, class_names
