import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from dsp_audioreclabs_b200 import batch
from oracle import synth, frontend_oracle as fo
ctx = batch.default_context(0)
lens = [9001, 12347, 7777, 15003, 8192, 1000, 5000, 30001, 44100, 20011, 333, 25000, 26001, 40000, 12000, 13001]
utts = [synth.utterance_pcm(70 + i, n, seed0=3) for i, n in enumerate(lens)] * 3
samples = np.concatenate(utts); off = np.concatenate([[0], np.cumsum([len(u) for u in utts])]).astype(np.int64)
ctx.set_tuning("pcm_variant", 10)
res = batch.frontend_batch(samples, off, 256, 128, "hamming", emit_epd_lists=True, ctx=ctx)
bad = 0
for b, u in enumerate(utts):
    r = fo.frontend_utterance(u, 256, 128, "hamming")
    ok = (int(res.start[b]), int(res.end[b]), int(res.n_frames[b])) == (r["start"], r["end"], len(r["zcr"]))
    if ok and len(r["zcr"]):
        e, m, z = res.frames(b)
        ok = np.array_equal(z.astype(np.float64), r["zcr"]) and np.allclose(e, r["energy"], rtol=1e-5, atol=0)
    bad += not ok
print("misaligned batch through the pipelined kernel: mismatches", bad, "replayed", int((res.status >= 0x100).sum()), "of", len(utts))
