"""Synthetic 44.1 kHz int16 utterances (SURVEY.md section 8(d) recipe).

Test/bench infrastructure; deterministic per (seed0, index).  Each utterance is
a noise floor, a short louder-noise "unvoiced onset" and a Hann-enveloped
two-partial burst whose pitch depends on the class ``index % 10``, plus a DC
offset, quantised (round toward zero) to int16 PCM -- the only thing the
reference's ``load_wav`` (src/audio_processing.py:9-46) can yield for 16-bit
mono files is ``pcm / 32768.0``.
"""
import numpy as np

SAMPLE_RATE = 44100


def utterance_pcm(index, n_samples=SAMPLE_RATE, seed0=1234, sample_rate=SAMPLE_RATE):
    """Return one utterance as int16[n_samples]."""
    rng = np.random.default_rng(seed0 + int(index))
    cls = int(index) % 10
    n = int(n_samples)
    t = np.arange(n, dtype=np.float64) / sample_rate
    x = rng.standard_normal(n) * 0.005
    b0 = int(rng.uniform(0.15, 0.30) * n)
    b1 = int(rng.uniform(0.60, 0.85) * n)
    f0 = 150.0 + 90.0 * cls + rng.uniform(-10.0, 10.0)
    nb = max(b1 - b0, 1)
    env = np.hanning(nb) if nb > 1 else np.ones(1)
    tb = t[b0:b0 + nb]
    burst = 0.6 * (np.sin(2 * np.pi * f0 * tb) + 0.3 * np.sin(2 * np.pi * (2 * f0 + 5 * cls) * tb))
    x[b0:b0 + nb] += env * burst
    on = int(0.050 * sample_rate)
    o0 = max(b0 - on, 0)
    x[o0:b0] += rng.standard_normal(b0 - o0) * 0.03
    x += 0.01
    x = np.clip(x, -1.0, 32767.0 / 32768.0)
    return np.trunc(x * 32768.0).astype(np.int16)


def ragged_lengths(batch, lo=0.8, hi=1.2, seed=99, sample_rate=SAMPLE_RATE):
    rng = np.random.default_rng(seed)
    return (rng.uniform(lo, hi, batch) * sample_rate).astype(np.int64)


def batch_pcm(batch, n_samples=SAMPLE_RATE, seed0=1234, first_index=0, lengths=None):
    """Return (samples int16[total], offsets int64[batch+1]) for a packed ragged batch."""
    if lengths is None:
        lengths = np.full(batch, int(n_samples), dtype=np.int64)
    offsets = np.zeros(batch + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    samples = np.empty(int(offsets[-1]), dtype=np.int16)
    for b in range(batch):
        samples[offsets[b]:offsets[b + 1]] = utterance_pcm(first_index + b, int(lengths[b]), seed0)
    return samples, offsets


def write_wav(path, pcm, width=2, channels=1, sample_rate=SAMPLE_RATE):
    import wave
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(width)
        w.setframerate(sample_rate)
        w.writeframes(np.ascontiguousarray(pcm).tobytes())


def write_config1_dataset(root, classes=10, clips=20, seed=1234):
    """BASELINE configs[0] / SURVEY.md 8(d) config 1: class folders '0'..'9' x `clips` 16-bit mono WAVs of
    U(0.8, 1.2) s (class c, clip i is utterance 10 i + c of the generator).  Returns the number of files."""
    import os
    rng = np.random.default_rng(seed)
    for c in range(classes):
        d = os.path.join(str(root), str(c))
        os.makedirs(d, exist_ok=True)
        for i in range(clips):
            n = int(rng.uniform(0.8, 1.2) * SAMPLE_RATE)
            write_wav(os.path.join(d, f"clip_{i:02d}.wav"), utterance_pcm(10 * i + c, n, seed0=seed))
    return classes * clips
