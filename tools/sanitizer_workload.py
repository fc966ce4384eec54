#!/usr/bin/env python3
"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitizer_workload.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsp_audioreclabs_b200 import batch, mfcc_dtw
from oracle import synth, frontend_oracle as fo
ctx = batch.default_context(0)
lens = synth.ragged_lengths(40, 0.2, 1.1, seed=3)
utts = [synth.utterance_pcm(i, int(n), seed0=11) for i, n in enumerate(lens)]
s, o, l = batch.pack_aligned(utts)
ctx.set_tuning("pcm_variant", 10)                      # the pipelined kernel for every geometry
for fl, fs, w in ((256, 128, "hamming"), (256, 128, "hanning"), (512, 256, "rectangular"), (1102, 441, "hamming"), (64, 32, "rectangular")):
    r = batch.frontend_batch(s, o, fl, fs, w, emit_epd_lists=True, lengths=l, ctx=ctx)
    ref = [fo.frontend_utterance(u, fl, fs, w) for u in utts[:6]]
    for b, rr in enumerate(ref):
        assert (int(r.start[b]), int(r.end[b])) == (rr["start"], rr["end"]), (fl, fs, b)
ctx.set_tuning("pcm_variant", -1)
r = batch.frontend_batch(s, o, 1102, 441, "hamming", lengths=l, emit_frames=False, ctx=ctx)     # resident kernel
rng = np.random.default_rng(0)
train, q = rng.standard_normal((300, 200)), rng.standard_normal((70, 200))
knn = batch.KNN(3, ctx=ctx).fit(train, rng.integers(0, 5, 300))
knn.predict(q)                                          # tensor-core scan + wide rerank
knn15 = batch.KNN(3, ctx=ctx).fit(train[:, :15].copy(), rng.integers(0, 5, 300))
knn15.predict(q[:, :15].copy())
mf, off = mfcc_dtw.mfcc_batch(s, o, r.start, r.end, lengths=l, ctx=ctx)
seqs = [mf[off[b]:off[b + 1]] for b in range(12)]
mfcc_dtw.DTWClassifier(2, ctx=ctx).fit(seqs[:8], np.arange(8) % 3).kneighbors(seqs[8:])
print("sanitizer workload ok")
