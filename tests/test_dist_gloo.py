"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (sharding, candidate
exchange, merge).  The compute callbacks are the oracle here -- the product has no CPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_topk(train, labels, queries, k, index_base):
    from oracle import knn_oracle as ko
    idx, d2 = ko.knn_topk(train.numpy(), queries.numpy(), k)
    lab = labels.numpy()[idx]
    return torch.from_numpy(d2), torch.from_numpy(idx + index_base), torch.from_numpy(lab.astype(np.int32))


def _oracle_topk_bounded(train, labels, queries, k, index_base, bound=None):
    """A local_topk that takes the threshold hints and USES them: rows outside a query's radius are dropped, so the
    exchange and the merge see short lists (what the CUDA library returns for a bounded call)."""
    d2, idx, lab = _oracle_topk(train, labels, queries, k, index_base)
    if bound is not None:
        out = d2 > bound[:, None]
        d2 = torch.where(out, torch.full_like(d2, float("inf")), d2)
        idx = torch.where(out, torch.full_like(idx, -1), idx)
        lab = torch.where(out, torch.full_like(lab, -1), lab)
    return d2, idx, lab


def _oracle_merge(cd, ci, cl):
    from oracle import knn_oracle as ko
    mi, _ = ko.merge_candidates(cd.numpy(), ci.numpy(), cd.shape[2])
    flat_i = np.transpose(ci.numpy(), (1, 0, 2)).reshape(ci.shape[1], -1)
    flat_l = np.transpose(cl.numpy(), (1, 0, 2)).reshape(ci.shape[1], -1)
    lab = np.stack([[flat_l[q][np.where(flat_i[q] == i)[0][0]] for i in mi[q]] for q in range(len(mi))]) \
        if len(mi) else np.zeros((0, cd.shape[2]), np.int32)
    classes = np.unique(cl.numpy())
    return torch.from_numpy(ko.vote(lab, classes).astype(np.int32))


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dsp_audioreclabs_b200 import dist as ddist
    k = np.load(os.path.join(GOLDEN, "knn_golden.npz"))
    xn, y, qn = k["d15/train_norm"], k["d15/train_labels"].astype(np.int32), k["d15/query_norm"][:120]
    tb = ddist.balanced_bounds(len(xn), world)
    qb = ddist.balanced_bounds(len(qn), world)
    knn = ddist.ShardedKNN(3, local_topk=_oracle_topk, merge_vote=_oracle_merge)
    knn.fit(torch.from_numpy(xn[tb[rank]:tb[rank + 1]]), torch.from_numpy(y[tb[rank]:tb[rank + 1]]))
    assert knn.index_base == tb[rank]
    q_local = torch.from_numpy(qn[qb[rank]:qb[rank + 1]])
    a = knn.predict_sharded(q_local).numpy()                # the candidate all-gather (sizes exchanged)
    a2 = knn.predict_sharded(q_local, [len(q_local)] * world).numpy()      # equal counts known to the caller: no size exchange
    assert np.array_equal(a, a2)
    assert np.array_equal(knn.predict(q_local).numpy(), a)  # automatic choice (replicated here: the train set is tiny)
    b = knn.predict_replicated(q_local).numpy()
    hinted = ddist.ShardedKNN(3, local_topk=_oracle_topk_bounded, merge_vote=_oracle_merge)
    hinted.fit(torch.from_numpy(xn[tb[rank]:tb[rank + 1]]), torch.from_numpy(y[tb[rank]:tb[rank + 1]]))
    assert hinted._takes_bound and not knn._takes_bound
    assert np.array_equal(hinted.predict_sharded(q_local).numpy(), a)                         # threshold hints travel with the queries
    assert np.array_equal(hinted.predict_sharded(q_local, [len(q_local)] * world).numpy(), a)
    mean, std = ddist.zscore_stats_allreduce(torch.from_numpy(k["d15/train"][tb[rank]:tb[rank + 1]]))
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), a=a, b=b, lo=qb[rank], hi=qb[rank + 1],
             mean=mean.numpy(), std=std.numpy())
    dist.destroy_process_group()


def test_sharded_knn_world2_matches_single_rank(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    k = np.load(os.path.join(GOLDEN, "knn_golden.npz"))
    for r in range(world):
        o = np.load(tmp_path / f"r{r}.npz")
        ref = k["d15/pred"][int(o["lo"]):int(o["hi"])]
        assert np.array_equal(o["a"], ref)          # candidate all-gather path
        assert np.array_equal(o["b"], ref)          # train all-gather fast path
        assert np.allclose(o["mean"], k["d15/mean"], rtol=1e-12, atol=1e-14)
        assert np.allclose(o["std"], k["d15/std"], rtol=1e-12)


def test_utterance_shards_balance_by_samples():
    from dsp_audioreclabs_b200 import dist as ddist
    rng = np.random.default_rng(0)
    lens = rng.integers(30000, 53000, 1000)
    off = np.concatenate([[0], np.cumsum(lens)])
    for world in (1, 2, 4, 8):
        b = ddist.utterance_shards(off, world)
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0)
        per = np.array([off[b[r + 1]] - off[b[r]] for r in range(world)])
        assert per.max() - per.min() <= 2 * lens.max()
    assert list(ddist.balanced_bounds(10, 4)) == [0, 3, 6, 8, 10]


def _oracle_dtw_topk(templates, labels, queries, k, index_base):
    from oracle import mfcc_dtw_oracle as mo
    idx, cost = mo.dtw_topk(queries, templates, min(k, len(templates)))
    return cost, idx + index_base, np.asarray(labels)[idx].astype(np.int32)


def _dtw_data():
    rng = np.random.default_rng(21)
    def seq(n, c): return (rng.standard_normal((n, 4)) * 0.2 + np.cos(np.arange(n)[:, None] * (0.2 + 0.1 * c))).astype(np.float32)
    labels = np.arange(11) % 3
    temps = [seq(int(rng.integers(5, 30)), c) for c in labels]
    quers = [seq(int(rng.integers(5, 30)), c) for c in (0, 1, 2, 1, 0)]
    return temps, labels, quers


def _dtw_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dsp_audioreclabs_b200 import dist as ddist
    temps, labels, quers = _dtw_data()
    tb = ddist.balanced_bounds(len(temps), world)
    sd = ddist.ShardedDTW(3, local_topk=_oracle_dtw_topk).fit(temps[tb[rank]:tb[rank + 1]], labels[tb[rank]:tb[rank + 1]])
    assert sd.index_base == tb[rank]
    cost, idx, lab = sd.kneighbors(quers)
    np.savez(os.path.join(out_dir, f"d{rank}.npz"), cost=cost, idx=idx, lab=lab, pred=sd.predict(quers))
    dist.destroy_process_group()


def test_sharded_dtw_world2_matches_single_rank(tmp_path):
    """Templates sharded by row over two ranks, one candidate all-gather: the merged neighbours equal the
    single-rank result of the (self-)oracle on every rank."""
    from oracle import mfcc_dtw_oracle as mo
    world = 2
    mp.spawn(_dtw_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    temps, labels, quers = _dtw_data()
    ridx, rcost = mo.dtw_topk(quers, temps, 3)
    for r in range(world):
        o = np.load(tmp_path / f"d{r}.npz")
        assert np.array_equal(o["idx"], ridx) and np.allclose(o["cost"], rcost, rtol=1e-12)
        assert np.array_equal(o["lab"], labels[ridx])
        votes = [np.bincount(labels[row], minlength=3).argmax() for row in ridx]
        assert np.array_equal(o["pred"], votes)
