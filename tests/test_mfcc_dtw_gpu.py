"""GPU checks of the MFCC + DTW variant (BASELINE config 5) against its SELF-ORACLE (oracle/mfcc_dtw_oracle.py):
the reference has no MFCC / DTW code, so parity here is unpinned by construction.
Tolerances (north star): MFCC 1e-4 relative per coefficient (floor: 5 % of the frame's largest coefficient), DTW costs 1e-5 relative (fp32)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_mfcc_matches_the_self_oracle(ctx):
    from dsp_audioreclabs_b200 import batch, mfcc_dtw
    from oracle import mfcc_dtw_oracle as mo, synth
    lens = synth.ragged_lengths(14, 0.25, 1.0, seed=8)
    utts = [synth.utterance_pcm(i, int(n), seed0=321) for i, n in enumerate(lens)]
    utts.append(np.full(3000, 1234, dtype=np.int16))              # constant signal: every mel energy hits the log floor
    utts.append(utts[0][:700].copy())                             # shorter than one frame: a single zero-padded frame
    s, o, l = batch.pack_aligned(utts)
    res = batch.frontend_batch(s, o, 1102, 441, "hamming", lengths=l, emit_frames=False, ctx=ctx)
    mf, off = mfcc_dtw.mfcc_batch(s, o, res.start, res.end, lengths=l, ctx=ctx)
    for b, u in enumerate(utts):
        ref, st, en = mo.mfcc_utterance(u)
        assert (st, en) == (int(res.start[b]), int(res.end[b]))
        got = mf[off[b]:off[b + 1]]
        assert got.shape == ref.shape, b
        # north star: 1e-4 relative per coefficient.  Floor for near-zero coefficients: a coefficient is a signed sum of
        # 26 log-mel energies, so below 5 % of the frame's largest coefficient it is held to 1e-4 of that floor instead
        scale = np.abs(ref).max(axis=1, keepdims=True) if len(ref) else 1.0
        tol = 1e-4 * np.maximum(np.abs(ref), 0.05 * scale)
        assert np.all(np.abs(got - ref) <= tol), (b, float((np.abs(got - ref) / tol).max()))


@pytest.mark.parametrize("dim", [13, 1, 16, 17, 39, 64, 100])
def test_dtw_costs_and_neighbours(ctx, dim):
    from dsp_audioreclabs_b200 import mfcc_dtw
    from oracle import mfcc_dtw_oracle as mo
    rng = np.random.default_rng(dim)
    def seq(n, c): return (rng.standard_normal((n, dim)) * 0.3 + np.sin(np.arange(n)[:, None] * (0.1 + 0.05 * c)) * 2).astype(np.float32)
    t_lens = [1, 2, 31, 32, 33, 64, 65, 100, 128, 129, 200, 256, 40, 57, 90, 7, 300, 513]
    # queries longer than one strip of 32 R rows (R = 8 / 4 / 2 / 1 for dim <= 16 / 32 / 64 / 128) cross strip boundaries
    q_lens = [1, 5, 32, 33, 64, 97, 128, 130, 255, 256, 77, 257, 300, 511, 512, 513, 700]
    labels = np.array([i % 4 for i in range(len(t_lens))])
    temps = [seq(n, labels[i]) for i, n in enumerate(t_lens)]
    quers = [seq(n, i % 4) for i, n in enumerate(q_lens)]
    clf = mfcc_dtw.DTWClassifier(3, ctx=ctx).fit(temps, labels)
    nc, ni, nl, cost = clf.kneighbors(quers, return_matrix=True)
    ref = mo.dtw_matrix(quers, temps)
    assert np.allclose(cost, ref, rtol=1e-5, atol=0)
    ridx, rcost = mo.dtw_topk(quers, temps, 3)
    assert np.allclose(nc, rcost, rtol=1e-5, atol=0)
    for i in range(len(quers)):                                   # same neighbours wherever the oracle's costs are not within fp32 of a tie
        gaps = np.diff(np.sort(ref[i])[:4]) / np.sort(ref[i])[:3]
        if np.all(gaps > 1e-4):
            assert np.array_equal(ni[i], ridx[i]), i
            assert np.array_equal(nl[i], labels[ridx[i]])
    pred = clf.predict(quers)
    assert pred.shape == (len(quers),)


def test_mfcc_dtw_template_matching_end_to_end(ctx):
    """Config 5 in miniature: synthetic utterances of 5 classes, MFCC sequences, 1-NN under DTW = the self-oracle's."""
    from dsp_audioreclabs_b200 import batch, mfcc_dtw
    from oracle import mfcc_dtw_oracle as mo, synth
    lens = synth.ragged_lengths(30, 0.3, 0.7, seed=4)
    utts = [synth.utterance_pcm(i, int(n), seed0=777) for i, n in enumerate(lens)]
    labels = np.array([i % 10 for i in range(30)])
    s, o, l = batch.pack_aligned(utts)
    res = batch.frontend_batch(s, o, 1102, 441, "hamming", lengths=l, emit_frames=False, ctx=ctx)
    mf, off = mfcc_dtw.mfcc_batch(s, o, res.start, res.end, lengths=l, ctx=ctx)
    seqs = [mf[off[b]:off[b + 1]] for b in range(30)]
    clf = mfcc_dtw.DTWClassifier(1, ctx=ctx).fit(seqs[:20], labels[:20])
    nc, ni, nl = clf.kneighbors(seqs[20:])
    ref_seqs = [mo.mfcc_utterance(u)[0] for u in utts]
    ridx, rcost = mo.dtw_topk(ref_seqs[20:], ref_seqs[:20], 1)
    assert np.allclose(nc, rcost, rtol=2e-4, atol=0)             # fp32 MFCC features feed the costs
    agree = np.mean(ni[:, 0] == ridx[:, 0])
    assert agree >= 0.9, agree


def test_sharded_dtw_with_the_cuda_callback(ctx):
    """dist.ShardedDTW on a one-rank gloo group with its default (CUDA) scoring: the exchange logic itself is covered
    for two ranks on CPU (tests/test_dist_gloo.py)."""
    import socket
    import torch.distributed as tdist
    from dsp_audioreclabs_b200 import dist as ddist
    from oracle import mfcc_dtw_oracle as mo
    rng = np.random.default_rng(2)
    def seq(n, c): return (rng.standard_normal((n, 13)) * 0.2 + np.cos(np.arange(n)[:, None] * (0.2 + 0.1 * c))).astype(np.float32)
    labels = np.arange(12) % 4 + 10
    temps = [seq(int(rng.integers(8, 60)), c) for c in labels]
    quers = [seq(int(rng.integers(8, 60)), c) for c in (10, 11, 12, 13, 11)]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    tdist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        sd = ddist.ShardedDTW(3).fit(temps, labels)
        cost, idx, lab = sd.kneighbors(quers)
        ridx, rcost = mo.dtw_topk(quers, temps, 3)
        assert np.array_equal(idx, ridx) and np.allclose(cost, rcost, rtol=1e-5)
        assert np.array_equal(lab, labels[ridx])
        assert np.array_equal(sd.predict(quers), [np.bincount(labels[r]).argmax() for r in ridx])
    finally:
        tdist.destroy_process_group()
