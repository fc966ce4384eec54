"""NumPy restatement of the KNN classify step (TEST INFRASTRUCTURE ONLY).

The reference wraps scikit-learn (src/models.py:33-35,52-58; pin
scikit-learn==1.7.2, requirements.txt:20 -- a third-party dependency that is
not vendored under /root/reference):
    KNeighborsClassifier(n_neighbors=3): Minkowski p=2, uniform weights,
    algorithm='auto' -> kd_tree when D <= 15 and k < n//2, else brute
    (sklearn/neighbors/_base.py:615-641).
Published algorithm restated here: exact float64 squared Euclidean distance
sum_j (q_j - t_j)^2 accumulated in feature order (kd_tree's reduced distance),
the k smallest with ties resolved towards the lower train index, then a
majority vote where a tie goes to the smallest class label
(sklearn/neighbors/_classification.py:262-309).  Pinned against the live
sklearn in this container by tests/golden/knn_golden.npz (oracle/gen_golden.py).
"""
import numpy as np


def sqdist_rows(q, train):
    """float64 squared distances of one query to every train row, accumulated in
    feature order like the tree's rdist loop."""
    d = np.zeros(len(train))
    for j in range(train.shape[1]):
        t = q[j] - train[:, j]
        d += t * t
    return d


def knn_topk(train, queries, k):
    """(idx[m,k], sqdist[m,k]) sorted by (distance, train index)."""
    train = np.asarray(train, dtype=np.float64)
    queries = np.asarray(queries, dtype=np.float64)
    m = len(queries)
    idx = np.empty((m, k), dtype=np.int64)
    dist = np.empty((m, k))
    for i in range(m):
        d = sqdist_rows(queries[i], train)
        order = np.lexsort((np.arange(len(d)), d))[:k]
        idx[i] = order
        dist[i] = d[order]
    return idx, dist


def vote(neighbor_labels, classes):
    """Majority label; ties -> smallest class (argmax over a histogram ordered by class)."""
    out = np.empty(len(neighbor_labels), dtype=classes.dtype)
    for i, row in enumerate(neighbor_labels):
        counts = np.array([(row == c).sum() for c in classes])
        out[i] = classes[np.argmax(counts)]
    return out


def knn_predict(train, labels, queries, k=3):
    labels = np.asarray(labels)
    idx, _ = knn_topk(train, queries, k)
    return vote(labels[idx], np.unique(labels))


def merge_candidates(cand_dist, cand_idx, k):
    """Merge per-shard candidate lists [R, m, k] -> global top-k by (distance, index):
    the CPU model of the row-sharded multi-GPU exchange."""
    r, m, kk = cand_dist.shape
    d = np.transpose(cand_dist, (1, 0, 2)).reshape(m, r * kk)
    ix = np.transpose(cand_idx, (1, 0, 2)).reshape(m, r * kk)
    out_i = np.empty((m, k), dtype=np.int64)
    out_d = np.empty((m, k))
    for i in range(m):
        order = np.lexsort((ix[i], d[i]))[:k]
        out_i[i] = ix[i][order]
        out_d[i] = d[i][order]
    return out_i, out_d
