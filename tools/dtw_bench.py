#!/usr/bin/env python3
"""MFCC + DTW template matching at a slice of BASELINE config 5 (self-oracle variant): MFCC of N utterances, then
Q queries x T templates under DTW.  Times are host-call wall times (H2D / D2H included).
Usage: dtw_bench.py [utterances] [queries] [templates]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsp_audioreclabs_b200 import batch, mfcc_dtw
from oracle import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
nt = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
ctx = batch.default_context(0)
base = [synth.utterance_pcm(i, int(l), seed0=99) for i, l in enumerate(synth.ragged_lengths(200, 0.5, 1.5, seed=2))]
utts = [base[i % 200] for i in range(n)]
s, o, l = batch.pack_aligned(utts)
res = batch.frontend_batch(s, o, 1102, 441, "hamming", lengths=l, emit_frames=False, ctx=ctx)
mfcc_dtw.mfcc_batch(s[: o[8]], o[:9], res.start[:8], res.end[:8], lengths=l[:8], ctx=ctx)
t0 = time.perf_counter(); mf, off = mfcc_dtw.mfcc_batch(s, o, res.start, res.end, lengths=l, ctx=ctx); t1 = time.perf_counter()
audio_s = float(l.sum()) / 44100
print(f"MFCC: {n} utterances ({audio_s:.0f} audio-s, {off[-1]} frames) in {1e3 * (t1 - t0):.1f} ms = {audio_s / (t1 - t0):,.0f} audio-s/s (host call)")
seqs = [mf[off[b]:off[b + 1]] for b in range(n)]
q = [seqs[i % n] for i in range(nq)]
t = [seqs[(7 * i + 3) % n] for i in range(nt)]
clf = mfcc_dtw.DTWClassifier(3, ctx=ctx).fit(t, np.arange(nt) % 10)
clf.kneighbors(q[:16])
t0 = time.perf_counter(); nc, ni, nl = clf.kneighbors(q); t1 = time.perf_counter()
cells = float(sum(len(a) for a in q)) * float(sum(len(b) for b in t))
print(f"DTW: {nq} x {nt} pairs, mean {np.mean([len(a) for a in q]):.0f} x {np.mean([len(b) for b in t]):.0f} frames, "
      f"{1e3 * (t1 - t0):.1f} ms = {nq * nt / (t1 - t0):,.0f} pairs/s = {cells / (t1 - t0) / 1e9:.1f} G cells/s (host call); "
      f"10k x 10k at this rate: {1e8 / (nq * nt / (t1 - t0)):.1f} s on one GPU")
