#!/usr/bin/env python3
"""Front-end kernel time for several (frame_length, frame_shift) configurations and both kernels (BASELINE config 4:
one launch per configuration).  Usage: config_sweep.py [utterances]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dsp_audioreclabs_b200 import batch, device as devapi
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dev = torch.device("cuda", 0)
ctx = batch.default_context(0)
samples, row_offsets = bench.synth_batch_device(n, dev, seed=7)
stream = torch.cuda.Stream(device=dev)
cfgs = [(256, 128), (1102, 441), (64, 32), (128, 64), (512, 256), (1024, 512), (2048, 1024), (1024, 256), (512, 64), (2205, 441), (1102, 1323), (352, 441)]
if len(sys.argv) > 2: cfgs = [tuple(int(v) for v in c.split('/')) for c in sys.argv[2].split(',')]
for fl, fs in cfgs:
    row = []
    for variant in (0, 10):
        ctx.set_tuning("pcm_variant", variant)
        fe = devapi.DeviceFrontend(row_offsets, fl, fs, "hamming", ctx=ctx, device=dev)
        with torch.cuda.stream(stream):
            for _ in range(3): fe.run(samples, stream=stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3): fe.run(samples, stream=stream)
            e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / 3
        row.append((ms, int((fe.status >= 0x100).sum().item())))
    ctx.set_tuning("pcm_variant", -1)
    print(f"fl={fl:5d} fs={fs:5d}  resident {row[0][0]:7.3f} ms (replayed {row[0][1]})   pipelined {row[1][0]:7.3f} ms (replayed {row[1][1]})   {n} utterances", flush=True)
