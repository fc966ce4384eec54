#!/usr/bin/env python3
"""WAV ingest rate: the reference's per-file `wave` + np.frombuffer + /32768.0 (load_wav,
src/audio_processing.py:9-46) against the library's batched native read into one packed staging buffer.
Usage: wav_ingest_bench.py [files] [seconds_per_file]"""
import os, sys, tempfile, time, wave
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsp_audioreclabs_b200 import wavio
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
rng = np.random.default_rng(0)
pcm = rng.integers(-20000, 20000, int(44100 * sec), dtype=np.int16)
paths = []
for i in range(n):
    p = os.path.join(d, f"{i}.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(44100); w.writeframes(pcm.tobytes())
    paths.append(p)
def ref_load(p):
    with wave.open(p, "rb") as w:
        raw = w.readframes(w.getnframes())
    return np.frombuffer(raw, dtype=np.int16) / 32768.0
t0 = time.perf_counter(); tot = sum(ref_load(p).size for p in paths); t_ref = time.perf_counter() - t0
t0 = time.perf_counter(); groups, info = wavio.read_packed(paths); t_nat = time.perf_counter() - t0
assert sum(int(g.lengths.sum()) for g in groups) == tot
print(f"{n} files x {sec:g} s: reference load_wav loop {t_ref:.3f} s ({n * sec / t_ref:,.0f} audio-s/s); "
      f"native batched read {t_nat:.3f} s ({n * sec / t_nat:,.0f} audio-s/s, {tot * 2 / t_nat / 1e9:.2f} GB/s, "
      f"pinned={groups[0]._buffer.pinned}, threads={min(os.cpu_count() or 1, 16)})")
for p in paths: os.remove(p)
os.rmdir(d)
