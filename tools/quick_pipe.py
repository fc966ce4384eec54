#!/usr/bin/env python3
"""Quick check of the pipelined kernel: parity of a few utterances against the oracle, then the kernel time at 256/128.
Usage: quick_pipe.py [utterances]   (DSP_LIB_PATH selects a tuning build)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dsp_audioreclabs_b200 import batch, device as devapi
from oracle import synth, frontend_oracle as fo
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ctx = batch.default_context(0)
lens = synth.ragged_lengths(24, 0.2, 1.1, seed=3)
utts = [synth.utterance_pcm(i, int(m), seed0=11) for i, m in enumerate(lens)]
reps = int(os.environ.get("QUICK_REPS", "30"))          # 24 x 30 = 720 utterances: enough for the split launch
s, o, l = batch.pack_aligned(utts * reps)
ctx.set_tuning("pcm_variant", 10)
bad = 0
for w in ("hamming", "hanning", "rectangular"):
    r = batch.frontend_batch(s, o, 256, 128, w, emit_epd_lists=True, lengths=l, ctx=ctx)
    refs = [fo.frontend_utterance(u, 256, 128, w) for u in utts]
    for b in range(len(utts) * reps):
        u, rr = utts[b % len(utts)], refs[b % len(utts)]
        e, m, z = r.frames(b)
        ok = (int(r.start[b]), int(r.end[b])) == (rr["start"], rr["end"]) and np.array_equal(z.astype(np.float64), rr["zcr"]) \
            and np.allclose(e, rr["energy"], rtol=1e-5, atol=0) and np.allclose(m, rr["magnitude"], rtol=1e-5, atol=0)
        bad += (not ok)
print("parity mismatches:", bad, flush=True)
dev = torch.device("cuda", 0)
samples, row_offsets = bench.synth_batch_device(n, dev, seed=7)
stream = torch.cuda.Stream(device=dev)
for fl, fs in ((256, 128),):
    fe = devapi.DeviceFrontend(row_offsets, fl, fs, "hamming", ctx=ctx, device=dev)
    with torch.cuda.stream(stream):
        for _ in range(3): fe.run(samples, stream=stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5): fe.run(samples, stream=stream)
        e1.record(stream)
    stream.synchronize()
    print(f"fl={fl} fs={fs} pipelined {e0.elapsed_time(e1) / 5:.3f} ms  {n} utterances  replayed {int((fe.status >= 0x100).sum().item())}", flush=True)
