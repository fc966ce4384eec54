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


def test_row_sharded_knn_with_threshold_hints(ctx, golden_knn):
    """dsp_knn_topk_bounded_device: every query arrives with the k-th distance found in the shard that owns it (an upper
    bound on its k-th distance in the whole train set); the other shards return only their rows inside that radius --
    possibly fewer than k -- and the merge is still sklearn's answer bit for bit.  Also: bounds that are exactly the
    global k-th distance (ties at the radius must come back), +inf bounds (= the unbounded call), and the config-3 set."""
    import torch
    from dsp_audioreclabs_b200 import device
    k = golden_knn
    dev = torch.device("cuda", 0)
    xn = torch.from_numpy(k["d15/train_norm"]).to(dev)
    y = torch.from_numpy(k["d15/train_labels"].astype(np.int32)).to(dev)
    q = torch.from_numpy(k["d15/query_norm"]).to(dev)
    m = q.shape[0]
    full_kth = device.DeviceKNN(3, ctx=ctx, device=dev).fit(xn.contiguous(), y.contiguous()).topk(q)[0][:, 2].contiguous()
    for shards in (2, 8):
        bounds = np.linspace(0, xn.shape[0], shards + 1).astype(int)
        knns = [device.DeviceKNN(3, ctx=ctx, device=dev, index_base=int(bounds[r])).fit(
            xn[bounds[r]:bounds[r + 1]].contiguous(), y[bounds[r]:bounds[r + 1]].contiguous()) for r in range(shards)]
        owner = torch.arange(m, device=dev) % shards
        own_kth = torch.stack([kn.topk(q)[0][:, 2] for kn in knns])                     # [shards, m]
        for mode in ("owner", "tight", "inf"):
            if mode == "owner":
                bound = own_kth.gather(0, owner[None, :])[0].contiguous()
            elif mode == "tight":
                bound = full_kth                                                         # the global k-th distance itself
            else:
                bound = torch.full((m,), float("inf"), dtype=torch.float64, device=dev)
            cd, ci, cl = [], [], []
            short = 0
            for kn in knns:
                d2, idx, lab = kn.topk(q, bound)
                short += int((idx < 0).sum())
                assert bool(((idx < 0) == torch.isinf(d2)).all()) and bool(((idx < 0) == (lab < 0)).all())
                cd.append(d2); ci.append(idx); cl.append(lab)
            labels, idx, _ = knns[0].merge_vote(torch.stack(cd), torch.stack(ci), torch.stack(cl))
            torch.cuda.synchronize()
            assert np.array_equal(idx.cpu().numpy(), k["d15/nbr_idx"]), (shards, mode)
            assert np.array_equal(labels.cpu().numpy(), k["d15/pred"]), (shards, mode)
            assert (short > 0) == (mode != "inf"), (shards, mode, short)                  # hinted shards really return short lists


# ---- tensor-core candidate filter for d <= 15 (csrc/knn_tc16.cu) ---------------------------------------------
SCAN_FP32, SCAN_TC16 = 1, 3


def test_tc16_filter_serves_the_15_dim_path(ctx, golden_knn):
    """The sklearn fixture goes through the tcgen05 filter (scan kind 3); the float64 certificate accepts (almost)
    every query, and turning the filter off (fp32 tiled scan) gives the same neighbours."""
    import os
    from dsp_audioreclabs_b200 import batch
    k = golden_knn
    knn = batch.KNN(3, ctx=ctx).fit(k["d15/train_norm"], k["d15/train_labels"])
    _, idx, _ = knn.kneighbors(k["d15/query_norm"])
    rescanned, kind = knn.last_stats()
    assert kind == SCAN_TC16 and rescanned <= 2
    assert np.array_equal(idx, k["d15/nbr_idx"])
    os.environ["DSP_KNN_NO_TC16"] = "1"
    try:
        ref = batch.KNN(3, ctx=ctx).fit(k["d15/train_norm"], k["d15/train_labels"])
    finally:
        del os.environ["DSP_KNN_NO_TC16"]
    assert ref.last_stats()[1] == SCAN_FP32
    assert np.array_equal(ref.kneighbors(k["d15/query_norm"])[1], idx)


@pytest.mark.parametrize("d", [1, 2, 7, 15])
@pytest.mark.parametrize("n", [5, 127, 128, 129, 1000, 4097])
def test_tc16_filter_shapes(ctx, d, n):
    """Tile edges: train rows around multiples of 128, query counts around the 256-query work unit, thin features."""
    from dsp_audioreclabs_b200 import batch
    from oracle import knn_oracle as ko
    rng = np.random.default_rng(100 * d + n)
    train = rng.standard_normal((n, d)) * rng.uniform(0.2, 3.0, d)
    labels = rng.integers(0, 6, n)
    knn = batch.KNN(3, ctx=ctx).fit(train, labels)
    for m in (1, 255, 256, 257, 700):
        q = rng.standard_normal((m, d)) * 1.5
        _, idx, _ = knn.kneighbors(q)
        assert knn.last_stats()[1] == (SCAN_TC16 if n >= 64 else SCAN_FP32)      # under 64 train rows: the fp32 tiled scan
        assert np.array_equal(idx, ko.knn_topk(train, q, 3)[0]), (d, n, m)


def test_tc16_range_gates(ctx):
    """Values the split-fp16 filter does not cover: a query with |q|^2 > 8192 opens the device-side gate and the
    fp32 scan answers the call; a train row out of range switches the filter off at fit.  Exact either way."""
    from dsp_audioreclabs_b200 import batch
    from oracle import knn_oracle as ko
    rng = np.random.default_rng(5)
    train = rng.standard_normal((3000, 15))
    labels = rng.integers(0, 4, 3000)
    knn = batch.KNN(3, ctx=ctx).fit(train, labels)
    q = rng.standard_normal((300, 15))
    q[17] *= 400.0                                        # |q|^2 ~ 2.4e6
    assert np.array_equal(knn.kneighbors(q)[1], ko.knn_topk(train, q, 3)[0])
    assert knn.last_stats()[1] == SCAN_TC16                # the handle still owns a filter; this call was gated
    big = train.copy()
    big[5] *= 1000.0
    knn2 = batch.KNN(3, ctx=ctx).fit(big, labels)
    assert knn2.last_stats()[1] == SCAN_FP32
    assert np.array_equal(knn2.kneighbors(q[:16])[1], ko.knn_topk(big, q[:16], 3)[0])
    # tiny-scale features: the filter's absolute error floor makes the certificate reject, the float64 rescan answers
    small = train * 1e-6
    knn3 = batch.KNN(3, ctx=ctx).fit(small, labels)
    assert np.array_equal(knn3.kneighbors(q[:16] * 1e-6)[1], ko.knn_topk(small, q[:16] * 1e-6, 3)[0])


def test_knn_config3_scale_matches_sklearn_through_the_reference(ctx):
    """10,240 queries against the 100,000-row train set, z-score + KNN, neighbour
    indices and labels bit-exact against sklearn (kd_tree) driven by the reference's create_classifier('knn')
    (fixture: oracle/gen_golden_knn3.py)."""
    import os
    from conftest import GOLDEN
    from dsp_audioreclabs_b200 import batch
    from oracle.gen_golden_knn3 import checksum, config3_data
    g = np.load(os.path.join(GOLDEN, "knn_config3_golden.npz"))
    xtr, ytr, xq, yq = config3_data()
    if not np.array_equal(checksum(xtr, xq), g["checksum"]):
        pytest.skip("numpy's random stream differs from the one the fixture was generated with")
    tn, mu, sd = batch.zscore(xtr, ctx=ctx)
    qn, _, _ = batch.zscore(xq, mu, sd, ctx=ctx)
    knn = batch.KNN(3, ctx=ctx).fit(tn, ytr)
    dist, idx, _ = knn.kneighbors(qn)
    rescanned, kind = knn.last_stats()
    assert kind == SCAN_TC16 and rescanned <= 10, rescanned
    assert np.array_equal(idx, g["nbr_idx"])
    assert np.allclose(dist, g["nbr_dist"], rtol=1e-12, atol=0)
    pred = knn.predict(qn)
    assert np.array_equal(pred, g["pred"])
    assert float(np.mean(pred == yq)) == float(g["accuracy"])
