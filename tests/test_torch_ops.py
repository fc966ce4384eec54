"""torch.library custom ops (dsp_audioreclabs_b200/torch_ops.py): registration and fake (meta) shapes on CPU;
on the GPU the ops give exactly what the NumPy batch API gives and pass torch.library.opcheck."""
import numpy as np
import pytest
import torch


def test_ops_are_registered_with_fake_implementations():
    import dsp_audioreclabs_b200.torch_ops  # noqa: F401
    ns = torch.ops.dsp_audioreclabs
    for name in ("frontend_batch", "zscore_fit", "zscore_apply", "knn_predict", "knn_topk"):
        assert hasattr(ns, name)
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        s = torch.empty(1000, dtype=torch.int16)
        off = torch.empty(3, dtype=torch.int64)
        out = ns.frontend_batch(s, off, off, 2, 600, 9, 256, 128, 1, True, 0.5, 0.1, 1.5)
        assert [tuple(t.shape) for t in out] == [(2,), (2,), (2,), (2,), (2, 15), (9,), (9,), (9,)]
        assert out[4].dtype == torch.float32 and out[0].dtype == torch.int32
        x = torch.empty(7, 15, dtype=torch.float32)
        mu = torch.empty(15, dtype=torch.float64)
        assert ns.zscore_apply(x, mu, mu).dtype == torch.float64
        t = torch.empty(50, 15, dtype=torch.float64)
        assert tuple(ns.knn_predict(t, torch.empty(50, dtype=torch.int32), t[:5], 3).shape) == (5,)
        d2, idx, lab = ns.knn_topk(t, torch.empty(50, dtype=torch.int32), t[:5], 3)
        assert tuple(idx.shape) == (5, 3) and idx.dtype == torch.int64 and lab.dtype == torch.int32


@pytest.mark.gpu
def test_ops_match_the_batch_api(ctx, golden_knn):
    from dsp_audioreclabs_b200 import batch, torch_ops
    from oracle import synth
    ns = torch.ops.dsp_audioreclabs
    lens = synth.ragged_lengths(12, 0.2, 0.5, seed=5)
    samples, off = synth.batch_pcm(12, seed0=77, lengths=lens)
    ref = batch.frontend_batch(samples, off, 256, 128, "hanning", ctx=ctx)
    pl = torch_ops.plan(off, 256, 128)
    d_samples = torch.from_numpy(samples).cuda()
    args = (d_samples, pl.offsets, pl.feat_offsets, pl.n_utts, pl.max_len, pl.total_frames, 256, 128, 2, True, 0.5, 0.1, 1.5)
    start, end, n_frames, status, stats, energy, magnitude, zcr = ns.frontend_batch(*args)
    torch.cuda.synchronize()
    assert np.array_equal(start.cpu().numpy(), ref.start) and np.array_equal(end.cpu().numpy(), ref.end)
    assert np.array_equal(n_frames.cpu().numpy(), ref.n_frames) and np.array_equal(status.cpu().numpy(), ref.status)
    assert np.array_equal(stats.cpu().numpy(), ref.stats)
    # ragged per-frame outputs: utterance b owns [feat_offsets[b], feat_offsets[b] + n_frames[b]) (the rest of its slot is unwritten)
    fo, nf = pl.h_feat_offsets, ref.n_frames
    for name, mine in (("energy", energy), ("magnitude", magnitude), ("zcr", zcr)):
        mine = mine.cpu().numpy()
        for b in range(len(nf)):
            assert np.array_equal(mine[fo[b]:fo[b] + nf[b]], getattr(ref, name)[fo[b]:fo[b] + nf[b]]), (name, b)
    torch.library.opcheck(ns.frontend_batch, args, test_utils=("test_schema", "test_faketensor"))
    # z-score + KNN on the sklearn fixture
    k = golden_knn
    tr = torch.from_numpy(k["d15/train"]).cuda()
    mean, std = ns.zscore_fit(tr)
    assert np.array_equal(mean.cpu().numpy(), k["d15/mean"]) and np.array_equal(std.cpu().numpy(), k["d15/std"])
    tn = ns.zscore_apply(tr, mean, std)
    qn = ns.zscore_apply(torch.from_numpy(k["d15/query"]).cuda(), mean, std)
    assert np.array_equal(tn.cpu().numpy(), k["d15/train_norm"]) and np.array_equal(qn.cpu().numpy(), k["d15/query_norm"])
    y = torch.from_numpy(k["d15/train_labels"].astype(np.int32)).cuda()
    pred = ns.knn_predict(tn, y, qn, 3)
    d2, idx, lab = ns.knn_topk(tn, y, qn, 3)
    assert np.array_equal(pred.cpu().numpy(), k["d15/pred"]) and np.array_equal(idx.cpu().numpy(), k["d15/nbr_idx"])
    # float32 statistics -> float64 z-scores: the widening is exact, the arithmetic is the reference's
    s32 = torch.from_numpy(k["d15/query"].astype(np.float32)).cuda()
    z = ns.zscore_apply(s32, mean, std).cpu().numpy()
    assert np.array_equal(z, (k["d15/query"].astype(np.float32).astype(np.float64) - k["d15/mean"]) / k["d15/std"])
    torch.library.opcheck(ns.knn_predict, (tn, y, qn, 3), test_utils=("test_schema", "test_faketensor"))
