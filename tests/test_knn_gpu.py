"""GPU parity of z-score + KNN classify against scikit-learn outputs captured through the
reference's wrapper (tests/golden/knn_golden.npz) and against the NumPy oracle.
Bit-exact: z-scored features (float64), neighbour indices and predicted labels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["d15", "d40"])
def test_zscore_bit_exact(ctx, golden_knn, tag):
    from dsp_audioreclabs_b200 import batch
    k = golden_knn
    xn, mu, sd = batch.zscore(k[f"{tag}/train"], ctx=ctx)
    assert np.array_equal(mu, k[f"{tag}/mean"]) and np.array_equal(sd, k[f"{tag}/std"])
    assert np.array_equal(xn, k[f"{tag}/train_norm"])
    qn, _, _ = batch.zscore(k[f"{tag}/query"], mu, sd, ctx=ctx)
    assert np.array_equal(qn, k[f"{tag}/query_norm"])
    # std == 0 -> 1 (feature_extraction.py:177)
    x = k[f"{tag}/train"][:50].copy()
    x[:, 2] = 3.25
    xn, mu, sd = batch.zscore(x, ctx=ctx)
    assert sd[2] == 1.0 and np.all(xn[:, 2] == 0.0)


@pytest.mark.parametrize("tag", ["d15", "d40"])
def test_knn_matches_sklearn_fixture(ctx, golden_knn, tag):
    from dsp_audioreclabs_b200 import batch
    k = golden_knn
    knn = batch.KNN(3, ctx=ctx).fit(k[f"{tag}/train_norm"], k[f"{tag}/train_labels"])
    dist, idx, lab = knn.kneighbors(k[f"{tag}/query_norm"])
    assert np.array_equal(idx, k[f"{tag}/nbr_idx"])
    assert np.allclose(dist, k[f"{tag}/nbr_dist"], rtol=1e-12, atol=0)
    pred = knn.predict(k[f"{tag}/query_norm"])
    assert np.array_equal(pred, k[f"{tag}/pred"])
    assert np.mean(pred == k[f"{tag}/query_labels"]) == pytest.approx(float(k[f"{tag}/accuracy"]))


def test_knn_ties_duplicates_and_small_sets(ctx):
    from dsp_audioreclabs_b200 import batch
    from oracle import knn_oracle as ko
    rng = np.random.default_rng(3)
    # duplicated train rows: equal distances must resolve to the lower index
    base = rng.standard_normal((40, 15))
    train = np.concatenate([base, base, base[:10]])
    labels = rng.integers(0, 4, len(train))
    q = np.concatenate([base[:20] + 1e-3, rng.standard_normal((30, 15))])
    knn = batch.KNN(3, ctx=ctx).fit(train, labels)
    _, idx, _ = knn.kneighbors(q)
    ridx, _ = ko.knn_topk(train, q, 3)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(knn.predict(q), ko.knn_predict(train, labels, q, 3))
    # fewer train rows than candidates kept by the scan, and k == n
    tiny = rng.standard_normal((3, 15))
    knn = batch.KNN(3, ctx=ctx).fit(tiny, np.array([2, 0, 1]))
    assert np.array_equal(knn.predict(q), ko.knn_predict(tiny, np.array([2, 0, 1]), q, 3))
    # three-way vote tie -> smallest label
    assert set(knn.predict(q)) == {0}
    # other k
    for kk in (1, 5, 8):
        knn = batch.KNN(kk, ctx=ctx).fit(train, labels)
        assert np.array_equal(knn.kneighbors(q)[1], ko.knn_topk(train, q, kk)[0])


def test_knn_near_ties_force_the_float64_rescan(ctx):
    """Train rows packed closer than fp32 can separate: the certificate must reject the fp32
    candidates and the float64 rescan must still return the exact neighbours."""
    from dsp_audioreclabs_b200 import batch
    from oracle import knn_oracle as ko
    rng = np.random.default_rng(11)
    center = rng.standard_normal(15) * 3
    train = center + rng.standard_normal((3000, 15)) * 1e-6
    labels = rng.integers(0, 5, 3000)
    q = center + rng.standard_normal((64, 15)) * 1e-6
    knn = batch.KNN(3, ctx=ctx).fit(train, labels)
    assert np.array_equal(knn.kneighbors(q)[1], ko.knn_topk(train, q, 3)[0])


def test_row_sharded_knn_merge_equals_single_shard(ctx, golden_knn):
    """The multi-GPU exchange emulated on one GPU: per-shard top-k with index_base, candidate
    lists stacked as an all-gather would deliver them, merge + vote kernel."""
    import torch
    from dsp_audioreclabs_b200 import device
    k = golden_knn
    dev = torch.device("cuda", 0)
    xn = torch.from_numpy(k["d15/train_norm"]).to(dev)
    y = torch.from_numpy(k["d15/train_labels"].astype(np.int32)).to(dev)
    q = torch.from_numpy(k["d15/query_norm"]).to(dev)
    for shards in (2, 8):
        bounds = np.linspace(0, xn.shape[0], shards + 1).astype(int)
        cd, ci, cl = [], [], []
        for r in range(shards):
            kn = device.DeviceKNN(3, ctx=ctx, device=dev, index_base=int(bounds[r]))
            kn.fit(xn[bounds[r]:bounds[r + 1]].contiguous(), y[bounds[r]:bounds[r + 1]].contiguous())
            d2, idx, lab = kn.topk(q)
            cd.append(d2); ci.append(idx); cl.append(lab)
        labels, idx, _ = kn.merge_vote(torch.stack(cd), torch.stack(ci), torch.stack(cl))
        torch.cuda.synchronize()
        assert np.array_equal(idx.cpu().numpy(), k["d15/nbr_idx"])
        assert np.array_equal(labels.cpu().numpy(), k["d15/pred"])
