import numpy as np, sys
sys.path.insert(0,'/root/repo')
from dsp_audioreclabs_b200 import batch
from oracle import synth, frontend_oracle as fo
ctx = batch.default_context(0)
lens = synth.ragged_lengths(40, 0.2, 1.1, seed=3)
utts = [synth.utterance_pcm(i, int(n), seed0=11) for i, n in enumerate(lens)]
s, o, l = batch.pack_aligned(utts)
ctx.set_tuning("pcm_variant", 10)
for fl, fs, w in ((256,128,"hamming"),(256,128,"hanning"),(1102,441,"hamming"),(64,32,"rectangular")):
    r = batch.frontend_batch(s, o, fl, fs, w, emit_epd_lists=True, lengths=l, ctx=ctx)
    ref = [fo.frontend_utterance(u, fl, fs, w) for u in utts[:6]]
    for b, rr in enumerate(ref):
        assert (int(r.start[b]), int(r.end[b])) == (rr["start"], rr["end"]), (fl, fs, b)
print("sanitizer workload ok")
