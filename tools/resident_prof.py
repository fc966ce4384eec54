import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from dsp_audioreclabs_b200 import batch, device as devapi
import bench
n = 20000
dev = torch.device("cuda", 0)
ctx = batch.default_context(0)
samples, row_offsets = bench.synth_batch_device(n, dev, seed=7)
stream = torch.cuda.Stream(device=dev)
for fl, fs in ((1102, 441), (2205, 441), (352, 441), (64, 32)):
    fe = devapi.DeviceFrontend(row_offsets, fl, fs, "hamming", ctx=ctx, device=dev)
    with torch.cuda.stream(stream):
        for _ in range(2): fe.run(samples, stream=stream)
    stream.synchronize()
    print("geometry", fl, fs, flush=True)
