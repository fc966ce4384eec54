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
def timed(tag, offsets, lengths=None, fl=256, fs=128):
    fe = devapi.DeviceFrontend(offsets, fl, fs, "hamming", ctx=ctx, device=dev, lengths=lengths)
    with torch.cuda.stream(stream):
        for _ in range(3): fe.run(samples, stream=stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5): fe.run(samples, stream=stream)
        e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gbs = fe.algorithmic_bytes() / ms / 1e6
    print(f"{tag}: fl={fl} fs={fs} pipelined {ms:.3f} ms  {len(offsets) - 1} utterances  {gbs:.0f} GB/s = {gbs / 6554.2:.3f} of peak  "
          f"replayed {int((fe.status >= 0x100).sum().item())}", flush=True)


L = bench.UTT_LEN
LAYOUTS = os.environ.get("QUICK_LAYOUTS", "packed,aligned,ragged").split(",")
if "packed" in LAYOUTS:
    timed("packed CSR L=44100 (every other start 8 bytes off)", row_offsets)
if "aligned" in LAYOUTS:      # round 1's layout: every utterance padded to 44,104 samples (16-byte aligned starts)
    keep = samples
    bench.UTT_LEN = 44104
    samples, pad_off = bench.synth_batch_device(n, dev, seed=7)
    bench.UTT_LEN = L
    timed("16-byte aligned starts (L=44104)", pad_off)
    samples = keep
if "ragged" not in LAYOUTS:
    sys.exit(0)
rng = np.random.default_rng(99)
lo_s, hi_s = (float(v) for v in os.environ.get("QUICK_RAGGED", "0.8,1.2").split(","))
lens = (rng.uniform(lo_s, hi_s, n) * 44100).astype(np.int64)
samples, r_off = bench.synth_batch_device(n, dev, seed=8, lengths=lens)       # every utterance generated at its own length
timed(f"ragged U({lo_s},{hi_s}) s packed CSR (any alignment)", r_off)
