"""One process per GPU: how the path shards across the GPUs of a box.

* Front end: independent utterances -> contiguous utterance ranges per rank, balanced by
  sample count, NO collective on the data path (`utterance_shards`).
* KNN: the train set is sharded by row.  Every rank scores the (all-gathered) queries against
  its own rows, keeps its local top-k (distance, global row, label), and ONE all-gather of
  those candidate lists over NCCL/NVLink lets each rank merge and vote for its own queries
  (`ShardedKNN.predict`).  When the train matrix is small (D = 15) the cheaper equivalent is
  to all-gather the train rows once and classify locally (`ShardedKNN.predict_replicated`).

* DTW template matching (the self-specified MFCC + DTW variant): templates sharded by row, queries replicated,
  the same single all-gather of top-k candidates (`ShardedDTW`).

torch.distributed is plumbing only; the compute callbacks default to the CUDA library and can
be replaced (the CPU gloo tests plug in the oracle, since there is no CPU implementation here).
"""
import numpy as np
import torch
import torch.distributed as dist


def balanced_bounds(n, world):
    """Contiguous ranges [b[r], b[r+1]) of n items, sizes differing by at most one."""
    base, rem = divmod(int(n), int(world))
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def utterance_shards(offsets, world):
    """Contiguous utterance ranges per rank balanced by total samples (ragged batches):
    boundaries at the utterances closest to k/world of the cumulative sample count."""
    offsets = np.asarray(offsets, dtype=np.int64)
    b = len(offsets) - 1
    total = offsets[-1] - offsets[0]
    targets = offsets[0] + (np.arange(1, world) * total) // world
    cuts = np.searchsorted(offsets, targets, side="left")
    bounds = np.concatenate([[0], np.clip(cuts, 0, b), [b]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def _all_gather_rows(x, group=None):
    """all_gather of [n_r, ...] tensors with different n_r: returns (list of per-rank tensors)."""
    world = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return [o[:s] for o, s in zip(out, sizes)]


class ShardedKNN:
    """Row-sharded KNN classify across the ranks of a process group.

    local_topk(train, labels, queries, k, index_base) -> (sqdist[m,k] f64, idx[m,k] i64, label[m,k] i32)
    merge_vote(cand_sqdist[R,m,k], cand_idx[R,m,k], cand_label[R,m,k]) -> labels[m] i32
    Both default to the CUDA library (device tensors)."""

    def __init__(self, n_neighbors=3, group=None, local_topk=None, merge_vote=None):
        self.k = int(n_neighbors)
        self.group = group
        self._topk = local_topk or self._cuda_topk
        self._merge = merge_vote or self._cuda_merge
        self._knn = None

    # -- CUDA callbacks ---------------------------------------------------------------------
    def _cuda_topk(self, train, labels, queries, k, index_base):
        from .device import DeviceKNN
        if self._knn is None or self._knn_key != (train.data_ptr(), train.shape[0], index_base):
            self._knn = DeviceKNN(k, device=train.device, index_base=index_base).fit(train, labels)
            self._knn_key = (train.data_ptr(), train.shape[0], index_base)
        return self._knn.topk(queries)

    def _cuda_merge(self, cd, ci, cl):
        from .device import DeviceKNN
        helper = self._knn or DeviceKNN(self.k, device=cd.device)
        return helper.merge_vote(cd, ci, cl)[0]

    # -- API -------------------------------------------------------------------------------
    def fit(self, train_shard, labels_shard):
        """Keep this rank's rows; the global row index of its first row comes from an
        all-gather of the shard sizes."""
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        n = torch.tensor([train_shard.shape[0]], dtype=torch.int64, device=train_shard.device)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=self.group)
        self.index_base = int(sum(int(s.item()) for s in sizes[:rank]))
        self.train, self.labels = train_shard.contiguous(), labels_shard.contiguous()
        return self

    def predict(self, queries_local):
        """Labels for this rank's queries.  Collectives: one all-gather of the query features, one
        all-gather of the per-shard top-k candidates."""
        rank = dist.get_rank(self.group)
        per_rank_q = _all_gather_rows(queries_local.contiguous(), self.group)
        q_all = torch.cat(per_rank_q, dim=0)
        d2, idx, lab = self._topk(self.train, self.labels, q_all, self.k, self.index_base)
        world = dist.get_world_size(self.group)
        cd = [torch.empty_like(d2) for _ in range(world)]
        ci = [torch.empty_like(idx) for _ in range(world)]
        cl = [torch.empty_like(lab) for _ in range(world)]
        # the single candidate exchange (three dtypes -> three tensors of one logical all-gather)
        dist.all_gather(cd, d2, group=self.group)
        dist.all_gather(ci, idx, group=self.group)
        dist.all_gather(cl, lab, group=self.group)
        lo = sum(t.shape[0] for t in per_rank_q[:rank])
        hi = lo + queries_local.shape[0]
        return self._merge(torch.stack([c[lo:hi] for c in cd]), torch.stack([c[lo:hi] for c in ci]),
                           torch.stack([c[lo:hi] for c in cl]))

    def predict_replicated(self, queries_local):
        """D = 15 fast path: all-gather the (small) train shards once, classify own queries with no
        candidate exchange.  Same result as `predict`."""
        if getattr(self, "_full", None) is None:
            tr = torch.cat(_all_gather_rows(self.train, self.group), dim=0)
            lb = torch.cat(_all_gather_rows(self.labels, self.group), dim=0)
            self._full = (tr.contiguous(), lb.contiguous())
        d2, idx, lab = self._topk(self._full[0], self._full[1], queries_local.contiguous(), self.k, 0)
        return self._merge(d2[None], idx[None], lab[None])


class ShardedDTW:
    """Row-sharded DTW template matching (BASELINE config 5: templates sharded by row, queries replicated, ONE
    all-gather of the per-shard top-k candidates).  Sequences are lists of [frames, dim] float32 arrays.

    local_topk(templates, labels, queries, k, index_base) -> (cost[m,k] f64, idx[m,k] i64, label[m,k] i32), sorted by
    (cost, index); defaults to the CUDA library (mfcc_dtw.DTWClassifier).  The merge is a k-way selection over the
    R sorted lists by (cost, global index) and a vote with ties to the smallest label."""

    def __init__(self, n_neighbors=1, group=None, local_topk=None):
        self.k = int(n_neighbors)
        self.group = group
        self._topk = local_topk or self._cuda_topk
        self._clf = None

    def _cuda_topk(self, templates, labels, queries, k, index_base):
        from .mfcc_dtw import DTWClassifier
        if self._clf is None:
            self._clf = DTWClassifier(k, index_base=index_base).fit(templates, labels)
        nc, ni, nl = self._clf.kneighbors(queries)
        return nc, ni, self._clf.classes_[np.clip(nl, 0, None)].astype(np.int32) * (nl >= 0) - (nl < 0)

    def fit(self, template_shard, labels_shard):
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        n = torch.tensor([len(template_shard)], dtype=torch.int64)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=self.group)
        self.index_base = int(sum(int(s.item()) for s in sizes[:rank]))
        self.templates, self.labels = list(template_shard), np.asarray(labels_shard)
        return self

    def kneighbors(self, queries):
        """(cost[m,k], global template index[m,k], label[m,k]) for the (replicated) queries."""
        world = dist.get_world_size(self.group)
        c, i, l = self._topk(self.templates, self.labels, queries, self.k, self.index_base)
        c, i, l = (torch.from_numpy(np.ascontiguousarray(a)) for a in (c.astype(np.float64), i.astype(np.int64), l.astype(np.int32)))
        cc = [torch.empty_like(c) for _ in range(world)]
        ci = [torch.empty_like(i) for _ in range(world)]
        cl = [torch.empty_like(l) for _ in range(world)]
        dist.all_gather(cc, c, group=self.group)      # the single candidate exchange (three dtypes)
        dist.all_gather(ci, i, group=self.group)
        dist.all_gather(cl, l, group=self.group)
        cost = torch.cat(cc, dim=1).numpy(); idx = torch.cat(ci, dim=1).numpy(); lab = torch.cat(cl, dim=1).numpy()
        idx_key = np.where(idx < 0, np.iinfo(np.int64).max, idx)
        order = np.lexsort((idx_key, cost), axis=1)[:, :self.k]
        take = lambda a: np.take_along_axis(a, order, axis=1)
        return take(cost), take(idx), take(lab)

    def predict(self, queries):
        _, _, lab = self.kneighbors(queries)
        out = np.empty(len(lab), dtype=np.int64)
        for q, row in enumerate(lab):
            vals, cnt = np.unique(row[row >= 0], return_counts=True)
            out[q] = vals[np.argmax(cnt)] if len(vals) else -1      # argmax: first maximum = smallest label
        return out


def zscore_stats_allreduce(x_local, group=None):
    """Train-set mean / population std over the rows of ALL ranks: one tiny all-reduce of
    (count, sum, sum of squares about the global mean) -- float64, two passes like np.std."""
    n = torch.tensor([x_local.shape[0]], dtype=torch.float64, device=x_local.device)
    s = x_local.sum(dim=0)
    buf = torch.cat([n, s])
    dist.all_reduce(buf, group=group)
    mean = buf[1:] / buf[0]
    ss = ((x_local - mean) ** 2).sum(dim=0)
    dist.all_reduce(ss, group=group)
    std = torch.sqrt(ss / buf[0])
    return mean, std
