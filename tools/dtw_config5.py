#!/usr/bin/env python3
"""BASELINE config 5 (self-oracle variant): MFCC + DTW template matching, 10k queries x 10k templates of
variable length (utterances of U(0.5, 1.5) s -> 23..148 MFCC frames), templates sharded by row over the ranks, queries
replicated, ONE all-gather of the packed top-k candidates (dist.ShardedDTW).  One process per GPU:

  python tools/dtw_config5.py [queries] [templates]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dtw_config5.py

Rank 0 prints one JSON line: pairs/s over all ranks (wall time of the sharded kneighbors call incl. its host<->device
copies and the exchange, max over ranks) and a spot check of 8 queries against the single-shard answer."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from dsp_audioreclabs_b200 import batch, mfcc_dtw, dist as ddist
from oracle import synth

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl" , device_id=dev) if world > 1 else dist.init_process_group("gloo", rank=0, world_size=1, init_method="tcp://127.0.0.1:29533")
ctx = batch.default_context(local)
nbase = 400
base = [synth.utterance_pcm(i, int(l), seed0=99) for i, l in enumerate(synth.ragged_lengths(nbase, 0.5, 1.5, seed=2))]
s, o, l = batch.pack_aligned(base)
res = batch.frontend_batch(s, o, 1102, 441, "hamming", lengths=l, emit_frames=False, ctx=ctx)
mf, off = mfcc_dtw.mfcc_batch(s, o, res.start, res.end, lengths=l, ctx=ctx)
seqs = [mf[off[b]:off[b + 1]] for b in range(nbase)]
rng = np.random.default_rng(5)
queries = [seqs[i] + 0.01 * rng.standard_normal(seqs[i].shape).astype(np.float32) for i in rng.integers(0, nbase, nq)]
tsel = rng.integers(0, nbase, nt)
templates = [seqs[i] for i in tsel]
labels = tsel % 10
tb = ddist.balanced_bounds(nt, world)
sd = ddist.ShardedDTW(3, device=(dev if world > 1 else None)).fit(templates[tb[rank]:tb[rank + 1]], labels[tb[rank]:tb[rank + 1]])
sd.kneighbors(queries)                                       # full-size warm-up: the library's device buffers (a 400 MB cost matrix at N = 1) are allocated here
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
cost, idx, lab = sd.kneighbors(queries)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device=dev if world > 1 else "cpu")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
dt = float(t.item())
ok = None
if rank == 0:
    ref = mfcc_dtw.DTWClassifier(3, ctx=ctx).fit(templates, labels).kneighbors(queries[:8])
    ok = bool(np.array_equal(ref[1], idx[:8]) and np.allclose(ref[0], cost[:8], rtol=1e-6))
    qf, tf = np.array([len(a) for a in queries]), np.array([len(a) for a in templates])
    print(json.dumps({"config": "BASELINE configs[4] (MFCC + DTW, self-oracle): %d queries x %d templates, variable length" % (nq, nt),
                      "n_gpus": world, "seconds": dt, "pairs_per_s": nq * nt / dt, "G_cells_per_s": float(qf.sum()) * float(tf.sum()) / dt / 1e9,
                      "query_frames_min_mean_max": [int(qf.min()), float(qf.mean()), int(qf.max())],
                      "template_frames_min_mean_max": [int(tf.min()), float(tf.mean()), int(tf.max())],
                      "exchange": "templates sharded by row x%d, queries replicated, ONE all-gather of the packed top-3 candidates (%s)" % (world, "NCCL" if world > 1 else "single rank"),
                      "first_8_queries_equal_single_shard_answer": ok}), flush=True)
dist.destroy_process_group()
