"""GPU parity of the tensor-core KNN path (csrc/knn_dense.cu): the sequence-feature variant of
compare_feature_methods.py (D = 2 * max_len or 3 * max_len, sklearn brute force) against outputs captured from
the reference (tests/golden/knn_seq_golden.npz) and against the NumPy oracle at larger sizes.
Bit-exact: neighbour indices and predicted labels; distances 1e-12 relative (float64 rerank)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TAGS = ["seq_default", "seq_256", "seq3_default"]


@pytest.mark.parametrize("tag", TAGS)
def test_sequence_knn_matches_sklearn_fixture(ctx, golden_seq, tag):
    from dsp_audioreclabs_b200 import batch
    g = golden_seq
    assert str(g[f"{tag}/fit_method"]) == "brute" and g[f"{tag}/train_norm"].shape[1] > 63
    before = ctx.launch_count
    xn, mu, sd = batch.zscore(g[f"{tag}/train"], ctx=ctx)
    qn, _, _ = batch.zscore(g[f"{tag}/query"], mu, sd, ctx=ctx)
    assert np.array_equal(xn, g[f"{tag}/train_norm"]) and np.array_equal(qn, g[f"{tag}/query_norm"])
    knn = batch.KNN(3, ctx=ctx).fit(xn, g[f"{tag}/train_labels"])
    dist, idx, _ = knn.kneighbors(qn)
    assert np.array_equal(idx, g[f"{tag}/nbr_idx"])
    assert np.allclose(dist, g[f"{tag}/nbr_dist"], rtol=1e-12, atol=0)
    rescanned, kind = knn.last_stats()
    assert kind == 2 and rescanned == 0          # the tensor-core scan ran and every query was certified
    assert np.array_equal(knn.predict(qn), g[f"{tag}/pred"])
    assert ctx.launch_count > before


@pytest.mark.parametrize("n,m,d", [(3000, 300, 1000), (257, 129, 64), (130, 5, 2049), (5, 3, 100)])
def test_dense_scan_equals_the_oracle(ctx, n, m, d):
    """Ragged sizes around the 128 x 128 x 64 tiles, fewer train rows than candidates, several k."""
    from dsp_audioreclabs_b200 import batch
    from oracle import knn_oracle as ko
    rng = np.random.default_rng(n + d)
    centers = rng.standard_normal((6, d))
    ytr = rng.integers(0, 6, n)
    train = centers[ytr] + 0.7 * rng.standard_normal((n, d))
    q = centers[rng.integers(0, 6, m)] + 0.7 * rng.standard_normal((m, d))
    train[n // 2] = train[0]                 # a duplicate row: the tie goes to the lower index
    for k in (1, 3, 7):
        if k > n:
            continue
        knn = batch.KNN(k, ctx=ctx).fit(train, ytr)
        dist, idx, _ = knn.kneighbors(q)
        ref_idx, ref_d2 = ko.knn_topk(train, q, k)
        assert np.array_equal(idx, ref_idx), (n, m, d, k)
        assert np.allclose(dist ** 2, ref_d2, rtol=1e-12, atol=0)
        rescanned, kind = knn.last_stats()
        assert kind == 2 and (rescanned == 0 or n <= 8 or k >= 7), (n, m, d, k, rescanned)
        assert np.array_equal(knn.predict(q), ko.knn_predict(train, ytr, q, k))


def test_near_ties_and_large_values_fall_back_to_float64(ctx):
    """Rows closer than the split-fp16 scan can separate are rejected by the certificate and rescanned in float64;
    values outside the fp16 range switch the whole call to the float64 scan.  Either way the neighbours are exact."""
    from dsp_audioreclabs_b200 import batch
    from oracle import knn_oracle as ko
    rng = np.random.default_rng(3)
    d = 200
    center = rng.standard_normal(d) * 2
    train = center + rng.standard_normal((600, d)) * 1e-7
    labels = rng.integers(0, 4, 600)
    q = center + rng.standard_normal((40, d)) * 1e-7
    knn = batch.KNN(3, ctx=ctx).fit(train, labels)
    assert np.array_equal(knn.kneighbors(q)[1], ko.knn_topk(train, q, 3)[0])
    assert knn.last_stats() == (40, 2)           # nothing could be certified: all 40 queries were rescanned
    big = rng.standard_normal((300, d)) * 1e5
    qb = rng.standard_normal((20, d)) * 1e5
    knn = batch.KNN(3, ctx=ctx).fit(big, labels[:300])
    assert np.array_equal(knn.kneighbors(qb)[1], ko.knn_topk(big, qb, 3)[0])
    assert knn.last_stats() == (20, 0)
    ok = rng.standard_normal((300, d))
    qmix = rng.standard_normal((20, d))
    qmix[7, 11] = 7e4                        # one query value beyond fp16
    knn = batch.KNN(3, ctx=ctx).fit(ok, labels[:300])
    assert np.array_equal(knn.kneighbors(qmix)[1], ko.knn_topk(ok, qmix, 3)[0])


def test_sequence_features_end_to_end(ctx, golden_seq):
    """Front end -> (energy, zcr) sequences -> zero-pad to the longest -> flatten (compare_feature_methods.py:77-123):
    zcr columns exact, energy columns to the fp32 tolerance of the fused kernel."""
    from dsp_audioreclabs_b200 import batch
    from oracle import synth
    g = golden_seq
    tag = "seq_default"
    fl, fs, max_len, _ = (int(v) for v in g[f"{tag}/frame"])
    lens = g[f"{tag}/pcm_lengths"][:40]
    utts = [synth.utterance_pcm(i, int(n), seed0=9000) for i, n in enumerate(lens)]
    s, o, l = batch.pack_aligned(utts)
    res = batch.frontend_batch(s, o, fl, fs, "hamming", lengths=l, ctx=ctx)
    flat = np.zeros((len(utts), max_len, 2))
    for b in range(len(utts)):
        e, _, z = res.frames(b)
        flat[b, :len(e), 0], flat[b, :len(z), 1] = e, z
    ref = g[f"{tag}/flat"][:len(utts)].reshape(len(utts), max_len, 2)
    assert np.array_equal(flat[:, :, 1], ref[:, :, 1])
    assert np.allclose(flat[:, :, 0], ref[:, :, 0], rtol=1e-5, atol=0)
