import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from dsp_audioreclabs_b200 import batch
from oracle import synth
ctx = batch.default_context(0)
utts = [synth.utterance_pcm(i, 44100, seed0=5) for i in range(64)] * 32
s, o, l = batch.pack_aligned(utts)
for _ in range(2):
    r = batch.frontend_batch(s, o, 256, 128, "hamming", lengths=l, force_exact=True, ctx=ctx)
print("exact ok", int(r.n_frames.sum()))
