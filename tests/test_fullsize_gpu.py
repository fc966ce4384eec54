"""GPU, BASELINE config 2 at full size (100,000 utterances of 44,104 samples, frame 256 / shift
128): size-independent properties, plus the float64 replay kernel and the NumPy oracle on
sub-samples."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

N_UTTS = int(os.environ.get("DSP_FULLSIZE_UTTS", "100000"))
FL, FS = 256, 128


@pytest.fixture(scope="module")
def big(ctx):
    import torch
    sys.path.insert(0, ROOT)
    import bench
    dev = torch.device("cuda", 0)
    samples, offsets = bench.synth_batch_device(N_UTTS, dev, seed=2024)
    return torch, dev, samples, offsets


def run(devapi, ctx, dev, samples, offsets, window, **kw):
    import torch
    fe = devapi.DeviceFrontend(offsets, FL, FS, window, ctx=ctx, device=dev, **kw)
    fe.run(samples)
    torch.cuda.synchronize()
    return fe


@pytest.mark.parametrize("window", ["rectangular", "hamming", "hanning"])
def test_invariants_determinism_permutation(ctx, big, window):
    torch, dev, samples, offsets = big
    from dsp_audioreclabs_b200 import device as devapi
    L = bench_len = int(offsets[1] - offsets[0])
    a = run(devapi, ctx, dev, samples, offsets, window)
    st, en, nf, n1 = a.start.long(), a.end.long(), a.n_frames.long(), a.n_epd_frames.long()
    assert int((a.status & 0xff).abs().sum()) == 0
    assert bool(((st >= 0) & (st < en) & (en <= L)).all())
    assert bool((st % FS == 0).all())
    assert bool((n1 == (L - FL) // FS + 1).all())
    assert bool((nf == (en - st - FL) // FS + 1).all())          # trimmed segment: whole EPD frames
    assert float((nf * 1.0).mean()) < 0.8 * ((L - FL) // FS + 1)  # endpoints really trim
    fo = torch.from_numpy(a.h_feat_offsets[:-1]).to(dev)
    valid = (torch.arange(a.energy.numel(), device=dev) - torch.repeat_interleave(fo, torch.from_numpy(np.diff(a.h_feat_offsets)).to(dev))) \
        < torch.repeat_interleave(nf, torch.from_numpy(np.diff(a.h_feat_offsets)).to(dev))
    e, m, z = a.energy[valid], a.magnitude[valid], a.zcr[valid]
    assert bool((e >= 0).all()) and bool((m >= 0).all()) and bool(torch.isfinite(e).all())
    assert bool(((z >= 0) & (z <= FL - 1) & (z == z.round())).all())
    assert bool((e <= FL + 1e-3).all())                         # |z| <= 1 after peak normalisation
    s = a.stats
    for k in range(3):
        mean, sd, mx, mn, med = (s[:, 5 * k + j] for j in range(5))
        assert bool(((mx >= med) & (med >= mn) & (mx >= mean - 1e-6 * mx.abs()) & (mean >= mn - 1e-6 * mx.abs()) & (sd >= 0)).all())
    # determinism: a second launch is bit-identical
    b = run(devapi, ctx, dev, samples, offsets, window)
    for name in ("start", "end", "n_frames", "status", "energy", "magnitude", "zcr", "stats"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    # permutation: results do not depend on which CTA / in which order an utterance is processed.  Integer results
    # (endpoints, frame counts, crossing counts) never depend on where an utterance sits; the fp32 windowed sums are
    # bit-identical for utterances that keep their address modulo 16 bytes (the window pass reads 16-byte aligned
    # vectors, so the association of its partial sums follows the alignment) and agree to a few ulp otherwise.
    cap = int(a.h_feat_offsets[1] - a.h_feat_offsets[0])
    step16 = 8 // int(np.gcd(L, 8))                       # utterances i and i + step16 share their alignment class
    g = torch.Generator(device=dev).manual_seed(1)
    for keep_alignment in (True, False):
        if keep_alignment:
            assert N_UTTS % step16 == 0
            perm = (torch.randperm(N_UTTS // step16, device=dev, generator=g)[:, None] * step16
                    + torch.arange(step16, device=dev)[None, :]).reshape(-1)
        else:
            perm = torch.randperm(N_UTTS, device=dev, generator=g)
        shuffled = torch.zeros_like(samples)
        shuffled[: N_UTTS * L] = samples[: N_UTTS * L].view(N_UTTS, L)[perm].reshape(-1)
        c = run(devapi, ctx, dev, shuffled, offsets, window)
        assert torch.equal(c.start, a.start[perm]) and torch.equal(c.end, a.end[perm]) and torch.equal(c.n_frames, a.n_frames[perm])
        assert torch.equal(c.zcr[: N_UTTS * cap].view(N_UTTS, cap), a.zcr[: N_UTTS * cap].view(N_UTTS, cap)[perm])
        assert torch.equal(c.stats[:, 10:], a.stats[perm][:, 10:])
        ce, ae = c.energy[: N_UTTS * cap].view(N_UTTS, cap), a.energy[: N_UTTS * cap].view(N_UTTS, cap)[perm]
        if keep_alignment:
            assert torch.equal(c.stats, a.stats[perm]) and torch.equal(ce, ae)
        else:
            assert bool(((ce - ae).abs() <= 2e-6 * ae.abs()).all())
            sc = a.stats[perm][:, [2, 2, 2, 2, 2, 7, 7, 7, 7, 7]].abs()          # scale of a sequence: its maximum
            assert bool(((c.stats[:, :10] - a.stats[perm][:, :10]).abs() <= 2e-6 * sc).all())


def test_fast_kernel_equals_float64_replay_and_oracle_on_subsamples(ctx, big):
    torch, dev, samples, offsets = big
    from dsp_audioreclabs_b200 import device as devapi
    from oracle import frontend_oracle as fo
    L = int(offsets[1] - offsets[0])
    a = run(devapi, ctx, dev, samples, offsets, "hamming")
    rng = np.random.default_rng(7)
    pick = np.sort(rng.choice(N_UTTS, size=min(1500, N_UTTS), replace=False))
    sub = torch.cat([samples[i * L:(i + 1) * L] for i in pick.tolist()] + [torch.zeros(64, dtype=torch.int16, device=dev)])
    sub_off = np.arange(len(pick) + 1, dtype=np.int64) * L
    x = run(devapi, ctx, dev, sub, sub_off, "hamming", force_exact=True)
    pk = torch.from_numpy(pick).to(dev)
    assert bool((x.status >= 0x100).all())
    assert torch.equal(x.start, a.start[pk]) and torch.equal(x.end, a.end[pk]) and torch.equal(x.n_frames, a.n_frames[pk])
    cap = int(a.h_feat_offsets[1])
    zx, za = x.zcr[: len(pick) * cap].view(-1, cap), a.zcr[: N_UTTS * cap].view(-1, cap)[pk]
    ex, ea = x.energy[: len(pick) * cap].view(-1, cap), a.energy[: N_UTTS * cap].view(-1, cap)[pk]
    live = torch.arange(cap, device=dev)[None, :] < x.n_frames[:, None]
    assert torch.equal(zx[live], za[live])
    assert bool(((ex[live] - ea[live]).abs() <= 1e-5 * ex[live].abs()).all())
    # NumPy oracle on a few of the same utterances
    host = sub.cpu().numpy()
    for j in range(0, 40):
        r = fo.frontend_utterance(host[j * L:(j + 1) * L], FL, FS, "hamming")
        i = int(pick[j])
        assert (int(a.start[i]), int(a.end[i]), int(a.n_frames[i])) == (r["start"], r["end"], r["n_frames"])
        assert np.array_equal(za[j, : r["n_frames"]].cpu().numpy().astype(np.float64), r["zcr"])
        assert np.allclose(ea[j, : r["n_frames"]].cpu().numpy(), r["energy"], rtol=1e-5, atol=0)
