#!/usr/bin/env python3
"""KNN classify at BASELINE config 3 scale: 1M queries x 100k train rows, D = 15, k = 3
(KNN_QUERIES / KNN_TRAIN / KNN_DIM in the environment change the shape).
Single GPU: the whole product; under torchrun: train rows sharded, candidate all-gather."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsp_audioreclabs_b200 import batch, device as devapi, dist as ddist  # noqa: E402


def main():
    m = int(os.environ.get("KNN_QUERIES", 1000000))
    n = int(os.environ.get("KNN_TRAIN", 100000))
    d, k = int(os.environ.get("KNN_DIM", 15)), 3      # KNN_DIM > 63: the tensor-core scan (sequence features)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device=dev).manual_seed(5)
    centers = torch.randn(10, d, device=dev, generator=g, dtype=torch.float64) * 1.5
    ytr = torch.randint(0, 10, (n,), device=dev, generator=g)
    xtr = centers[ytr] + torch.randn(n, d, device=dev, generator=g, dtype=torch.float64)
    yq = torch.randint(0, 10, (m,), device=dev, generator=g)
    xq = centers[yq] + torch.randn(m, d, device=dev, generator=g, dtype=torch.float64)
    ctx = batch.default_context(local)
    xtr_n, mu, sd = devapi.zscore_device(xtr, ctx=ctx)
    xq_n, _, _ = devapi.zscore_device(xq, mu, sd, ctx=ctx)
    labels = ytr.to(torch.int32)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world == 1:
        knn = devapi.DeviceKNN(k, ctx=ctx, device=dev).fit(xtr_n.contiguous(), labels)
        knn.predict(xq_n)                     # full-size warm-up: the library's staging buffers are allocated here
        torch.cuda.synchronize()
        ev0.record()
        pred = knn.predict(xq_n)
        ev1.record()
        torch.cuda.synchronize()
        stats = knn.last_stats()
    else:
        tb = ddist.balanced_bounds(n, world)
        qb = ddist.balanced_bounds(m, world)
        sk = ddist.ShardedKNN(k).fit(xtr_n[tb[rank]:tb[rank + 1]].contiguous(), labels[tb[rank]:tb[rank + 1]].contiguous())
        ql = xq_n[qb[rank]:qb[rank + 1]].contiguous()
        sk.predict(ql[:1000].contiguous())
        torch.distributed.barrier(); torch.cuda.synchronize()
        ev0.record()
        pred = sk.predict(ql)
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        stats = None
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    acc = float((pred.long() == (yq if world == 1 else yq[qb[rank]:qb[rank + 1]])).double().mean())
    if rank == 0:
        pairs = float(m) * n
        print(json.dumps({"knn": f"{m} queries x {n} train, D={d}, k={k}", "n_gpus": world, "ms": ms,
                          "queries_per_s": m / (ms / 1e3), "pair_rate_per_s": pairs / (ms / 1e3),
                          "algorithmic_TFLOPs": 2 * d * pairs / (ms / 1e3) / 1e12, "accuracy_rank0": acc,
                          "rescanned_float64,scan_kind": stats}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
