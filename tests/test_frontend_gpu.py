"""GPU parity of the front end (through the C ABI) against the reference's golden
fixtures and the NumPy oracle.

Bars (BASELINE.json north_star): bit-exact zero-crossing counts, endpoint indices and frame
counts; energy / magnitude within 1e-5 relative (fp32 feature pass); float64 EPD energies of
the fast kernel within 1e-12 relative; everything the float64 replay kernel produces in
float64 is bit-identical to NumPy.
"""
import numpy as np
import pytest

from conftest import golden_names, golden_pcm

pytestmark = pytest.mark.gpu

RTOL_F32 = 1e-5          # north-star tolerance for energy / magnitude
RTOL_EPD = 1e-12         # fast-kernel float64 EPD energies (exact-integer formulation)


def pack(utts):
    lens = np.array([len(u) for u in utts], dtype=np.int64)
    off = np.zeros(len(utts) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return np.concatenate(utts) if len(utts) else np.zeros(0, np.int16), off


def assert_stats_close(got, ref, seqs):
    """North star: 1e-5 relative.  mean / max / min / median are order statistics or an average of per-frame values
    that are each within 1e-5 relative, so they are held to rtol = 1e-5 with NO absolute term.  The one exception is
    std: a perturbation of relative size eps per frame moves the standard deviation by up to eps * rms(sequence),
    which is not small next to a std that is itself tiny next to the mean -- its absolute term is 1e-5 * rms."""
    for s, key in enumerate(("energy", "magnitude", "zcr")):
        rms = float(np.sqrt(np.mean(np.square(seqs[key].astype(np.float64))))) if len(seqs[key]) else 0.0
        g, r = got[5 * s:5 * s + 5].astype(np.float64), ref[5 * s:5 * s + 5]
        for i, stat in enumerate(("mean", "std", "max", "min", "median")):
            atol = RTOL_F32 * rms if stat == "std" else 0.0
            assert np.isclose(g[i], r[i], rtol=RTOL_F32, atol=atol), (key, stat, g[i], r[i])


def check_against_golden(g, res, names, ci, win, rtol_feat):
    for b, name in enumerate(names):
        base = f"fe/{ci}/{name}/{win}"
        err = int(g[base + "/error"])
        st = int(res.status[b]) & 0xff
        if err:
            assert st != 0, (base, st)
            continue
        assert st == 0, (base, st)
        assert int(res.start[b]) == int(g[f"epd/{ci}/{name}/start"]), base
        assert int(res.end[b]) == int(g[f"epd/{ci}/{name}/end"]), base
        el, zl = res.epd_lists(b)
        rel, rzl = g[f"epd/{ci}/{name}/energy_list"], g[f"epd/{ci}/{name}/zcr_list"]
        assert len(el) == len(rel)
        assert np.array_equal(zl.astype(np.float64), rzl), base
        assert np.allclose(el, rel, rtol=RTOL_EPD, atol=0), (base, np.max(np.abs(el - rel) / np.maximum(rel, 1e-300)))
        assert int(res.n_frames[b]) == int(g[base + "/n_frames"]), base
        e, m, z = res.frames(b)
        assert np.array_equal(z.astype(np.float64), g[base + "/zcr"]), base
        assert np.allclose(e, g[base + "/energy"], rtol=rtol_feat, atol=0), base
        assert np.allclose(m, g[base + "/magnitude"], rtol=rtol_feat, atol=0), base
        assert_stats_close(res.stats[b], g[base + "/stats"],
                           {"energy": g[base + "/energy"], "magnitude": g[base + "/magnitude"], "zcr": g[base + "/zcr"]})


RESIDENT, STREAM, PIPE = 0, 2, 10      # builds of the fused kernel (frontend_pcm.cu kVariants); 10 = frontend_pipe.cu


@pytest.mark.parametrize("variant", [RESIDENT, STREAM, PIPE])
@pytest.mark.parametrize("ci", range(7))
@pytest.mark.parametrize("win", ["rectangular", "hamming", "hanning"])
def test_fast_kernel_matches_reference_fixtures(ctx, golden_fe, ci, win, variant):
    """Both builds of the fused kernel: samples resident in shared memory (TMA bulk loads, any
    alignment) and streaming from global memory / L2 (16-byte aligned utterances)."""
    from dsp_audioreclabs_b200 import batch
    g = golden_fe
    fl, fs = (int(v) for v in g["configs"][ci])
    names = golden_names(g)
    utts = [golden_pcm(g, n) for n in names]
    ctx.set_tuning("pcm_variant", variant)
    try:
        if variant in (STREAM, PIPE):
            samples, off, lengths = batch.pack_aligned(utts)          # 16-byte aligned starts + explicit lengths
            res = batch.frontend_batch(samples, off, fl, fs, win, emit_epd_lists=True, lengths=lengths, ctx=ctx)
            # the streaming kernel really ran: nothing except certified-margin flags was replayed
            assert int(((res.status & 0x100) != 0).sum()) <= 4   # exact-tie fixtures (zeros, constant, square steps)
        else:
            samples, off = pack(utts)
            res = batch.frontend_batch(samples, off, fl, fs, win, emit_epd_lists=True, ctx=ctx)
    finally:
        ctx.set_tuning("pcm_variant", -1)
    check_against_golden(g, res, names, ci, win, RTOL_F32)


@pytest.mark.parametrize("ci", [0, 1, 3])
@pytest.mark.parametrize("win", ["rectangular", "hamming", "hanning"])
def test_float64_replay_kernel_matches_reference_fixtures(ctx, golden_fe, ci, win):
    from dsp_audioreclabs_b200 import batch
    g = golden_fe
    fl, fs = (int(v) for v in g["configs"][ci])
    names = golden_names(g)
    samples, off = pack([golden_pcm(g, n) for n in names])
    res = batch.frontend_batch(samples, off, fl, fs, win, emit_epd_lists=True, force_exact=True, ctx=ctx)
    assert np.all((res.status & 0x100) != 0)
    check_against_golden(g, res, names, ci, win, 2e-7)       # float32 storage of float64-exact values
    for b, name in enumerate(names):                          # float64 outputs are bit-identical
        el, _ = res.epd_lists(b)
        assert np.array_equal(el, g[f"epd/{ci}/{name}/energy_list"]), name


def test_no_endpoint_detection_zero_padded_last_frame(ctx, golden_fe):
    from dsp_audioreclabs_b200 import batch
    g = golden_fe
    for ci in range(4):
        fl, fs = (int(v) for v in g["configs"][ci])
        names = ["syn0", "syn3", "short100", "len256"]
        samples, off = pack([g[f"pcm/{n}"] for n in names])
        for win in g["windows"]:
            for exact in (False, True):
                res = batch.frontend_batch(samples, off, fl, fs, str(win), do_endpoint_detection=False,
                                           force_exact=exact, ctx=ctx)
                for b, name in enumerate(names):
                    base = f"noepd/{ci}/{name}/{win}"
                    assert int(res.start[b]) == 0 and int(res.end[b]) == len(g[f"pcm/{name}"])
                    assert int(res.n_frames[b]) == int(g[base + "/n_frames"])
                    e, m, z = res.frames(b)
                    assert np.array_equal(z.astype(np.float64), g[base + "/zcr"]), (base, exact)
                    assert np.allclose(e, g[base + "/energy"], rtol=RTOL_F32, atol=0)
                    assert np.allclose(m, g[base + "/magnitude"], rtol=RTOL_F32, atol=0)


@pytest.mark.parametrize("fl,fs", [(256, 128), (1102, 441)])
def test_batch_against_numpy_oracle(ctx, fl, fs):
    """Seeded synthetic utterances (SURVEY.md 8(d) generator), ragged lengths, three windows."""
    from dsp_audioreclabs_b200 import batch
    from oracle import frontend_oracle as fo, synth
    nb = 96
    lens = synth.ragged_lengths(nb, 0.5, 1.2, seed=5)
    samples, off = synth.batch_pcm(nb, seed0=4321, lengths=lens)
    trimmed = 0
    for win in ("rectangular", "hamming", "hanning"):
        res = batch.frontend_batch(samples, off, fl, fs, win, emit_epd_lists=True, ctx=ctx)
        ref = fo.frontend_batch(samples, off, fl, fs, win)
        for b in range(nb):
            r = ref[b]
            assert res.ok(b) and r is not None
            assert (int(res.start[b]), int(res.end[b])) == (r["start"], r["end"]), b
            trimmed += r["start"] > 0 and r["end"] < lens[b]
            el, zl = res.epd_lists(b)
            assert np.array_equal(zl.astype(np.float64), r["zcr_list"])
            assert np.allclose(el, r["energy_list"], rtol=RTOL_EPD, atol=0)
            e, m, z = res.frames(b)
            assert len(e) == r["n_frames"]
            assert np.array_equal(z.astype(np.float64), r["zcr"])
            assert np.allclose(e, r["energy"], rtol=RTOL_F32, atol=0)
            assert np.allclose(m, r["magnitude"], rtol=RTOL_F32, atol=0)
            assert_stats_close(res.stats[b], r["stats"], r)
    assert trimmed > nb            # the generator yields real endpoints, not whole clips


def test_misaligned_and_odd_offsets(ctx):
    """Utterances that start at odd sample offsets take the non-TMA load path; results must not
    depend on where an utterance sits in the buffer."""
    from dsp_audioreclabs_b200 import batch
    from oracle import synth
    utts = [synth.utterance_pcm(50 + i, n, seed0=9) for i, n in enumerate([9001, 12345, 7777, 15001, 8192])]
    a_s, a_o = pack(utts)
    ra = batch.frontend_batch(a_s, a_o, 256, 128, "hamming", emit_epd_lists=True, ctx=ctx)
    # same utterances behind a 3-sample pad and in reverse order
    rev = utts[::-1]
    b_s, b_o = pack(rev)
    b_s = np.concatenate([np.zeros(3, np.int16), b_s])
    b_o = b_o + 3
    rb = batch.frontend_batch(b_s, b_o, 256, 128, "hamming", emit_epd_lists=True, ctx=ctx)
    for i in range(len(utts)):
        j = len(utts) - 1 - i
        assert (ra.start[i], ra.end[i], ra.n_frames[i]) == (rb.start[j], rb.end[j], rb.n_frames[j])
        (ea, ma, za), (eb, mb, zb) = ra.frames(i), rb.frames(j)
        assert np.array_equal(za, zb)
        # fp32 windowed sums: the window pass reads 16-byte aligned vectors, so the association of its partial sums
        # follows the utterance's address modulo 16 bytes -- a few ulp between placements, never an integer result
        assert np.allclose(ea, eb, rtol=2e-6, atol=0) and np.allclose(ma, mb, rtol=2e-6, atol=0)
        for x, y in zip(ra.epd_lists(i), rb.epd_lists(j)):
            assert np.array_equal(x, y)
        assert np.array_equal(ra.stats[i][10:], rb.stats[j][10:])
        assert np.allclose(ra.stats[i][:10], rb.stats[j][:10], rtol=0, atol=2e-6 * float(np.abs(ra.stats[i][:10]).max()))


def test_padded_layout_with_explicit_lengths_equals_packed(ctx):
    """offsets + lengths (aligned, padded storage) must give exactly what the packed CSR layout gives."""
    from dsp_audioreclabs_b200 import batch
    from oracle import synth
    utts = [synth.utterance_pcm(70 + i, n, seed0=3) for i, n in enumerate([9001, 12347, 7777, 15003, 8192, 100, 0, 5000])]
    a_s, a_o = pack(utts)
    b_s, b_o, b_l = batch.pack_aligned(utts)
    assert np.all(b_o % 8 == 0) and list(b_l) == [len(u) for u in utts]
    for exact in (False, True):
        ra = batch.frontend_batch(a_s, a_o, 256, 128, "hanning", emit_epd_lists=True, force_exact=exact, ctx=ctx)
        rb = batch.frontend_batch(b_s, b_o, 256, 128, "hanning", emit_epd_lists=True, force_exact=exact, lengths=b_l, ctx=ctx)
        for name in ("start", "end", "n_frames", "n_epd_frames"):
            assert np.array_equal(getattr(ra, name), getattr(rb, name)), name
        assert np.array_equal(ra.status & 0xff, rb.status & 0xff)
        # the packed layout runs the shared-memory-resident build, the aligned one the pipelined kernel:
        # integer outputs and float64 replay results are identical, fp32 sums may differ in the last bits
        for i in range(len(utts)):
            for k, (x, y) in enumerate(zip(ra.frames(i), rb.frames(i))):
                assert np.array_equal(x, y) if (exact or k == 2) else np.allclose(x, y, rtol=2e-6, atol=0)
            for k, (x, y) in enumerate(zip(ra.epd_lists(i), rb.epd_lists(i))):
                assert np.array_equal(x, y) if (exact or k == 1) else np.allclose(x, y, rtol=1e-13, atol=0)
    with pytest.raises(ValueError):
        batch.frontend_batch(b_s, b_o, 256, 128, lengths=b_l + 9, ctx=ctx)     # length exceeds its slot


def test_empty_and_degenerate_inputs(ctx):
    from dsp_audioreclabs_b200 import batch
    z = np.zeros(0, np.int16)
    res = batch.frontend_batch(z, np.array([0]), 256, 128, ctx=ctx)      # empty batch
    assert len(res) == 0
    # a zero-length utterance between two real ones: the reference raises ValueError for it
    from oracle import synth
    u = synth.utterance_pcm(1, 5000)
    samples = np.concatenate([u, u])
    off = np.array([0, 5000, 5000, 10000])
    for exact in (False, True):
        res = batch.frontend_batch(samples, off, 256, 128, force_exact=exact, ctx=ctx)
        assert res.ok(0) and res.ok(2) and not res.ok(1)
        assert int(res.n_frames[1]) == 0
        assert np.array_equal(res.stats[0], res.stats[2])
    with pytest.raises(ValueError):
        batch.frontend_batch(u, np.array([0, 5000]), 256, 128, "blackman", ctx=ctx)
    with pytest.raises(ValueError):
        batch.frontend_batch(u, np.array([0, 5000]), 0, 128, ctx=ctx)


def test_per_call_surface_float64(ctx, golden_fe):
    """remove_dc / normalize / preprocess / endpoint_detection / frame_signal /
    extract_frame_features on a non-PCM float64 signal: bit-identical to NumPy except the
    window's cos() (last-ulp) in frame_signal."""
    from dsp_audioreclabs_b200 import batch
    from oracle import frontend_oracle as fo
    g = golden_fe
    x = g["float/x"]
    assert np.array_equal(batch.preprocess(x, 0, ctx=ctx), x - np.mean(x))
    assert np.array_equal(batch.preprocess(x, 1, ctx=ctx), fo.normalize_audio(x))
    z = batch.preprocess(x, 2, ctx=ctx)
    assert np.array_equal(z, g["float/preprocessed"])
    s, e, el, zl = batch.endpoint_detection(z, 256, 128, ctx=ctx)
    assert (s, e) == (int(g["float/start"]), int(g["float/end"]))
    assert np.array_equal(el, g["float/energy_list"]) and np.array_equal(zl, g["float/zcr_list"])
    fr = batch.frame_signal(z[s:e], 256, 128, "hamming", ctx=ctx)
    assert fr.shape == g["float/frames"].shape
    assert np.allclose(fr, g["float/frames"], rtol=1e-14, atol=1e-17)
    en, mg, zc, st = batch.frame_features(g["float/frames"], ctx=ctx)
    assert np.array_equal(en, g["float/energy"]) and np.array_equal(mg, g["float/magnitude"])
    assert np.array_equal(zc, g["float/zcr"])
    assert np.array_equal(st, g["float/stats"])
    assert np.array_equal(batch.sequence_stats(g["float/energy"], ctx=ctx), g["float/stats"][:5])
    for win in g["windows"]:
        for n in (1, 2, 3, 64, 255, 256, 1102):
            assert np.allclose(batch.window(str(win), n), g[f"window/{win}/{n}"], rtol=1e-14, atol=1e-16)
    # a constant float signal: the result hinges on NumPy's exact summation order
    c = np.full(5000, 0.1)
    assert np.array_equal(batch.preprocess(c, 2, ctx=ctx), fo.preprocess(c))
    with pytest.raises(ValueError):
        batch.frame_features(np.zeros((0, 256)), ctx=ctx)


def test_wav_encodings_through_the_batch_api(ctx, golden_fe):
    """8-bit and stereo PCM (load_wav's other branches) go through the float64 replay kernel."""
    from dsp_audioreclabs_b200 import batch
    g = golden_fe
    for key in ("m16", "s16", "m8", "s8"):
        raw = g[f"wav/{key}/raw"]
        width, ch = (int(v) for v in g[f"wav/{key}/width_channels"])
        res = batch.frontend_batch(raw, np.array([0, raw.size]), 1102, 441, "hamming", channels=ch, ctx=ctx)
        assert [int(res.start[0]), int(res.end[0])] == list(g[f"wav/{key}/start_end"]), key
        assert np.allclose(res.stats[0], g[f"wav/{key}/stats"], rtol=2e-5, atol=0), key


@pytest.mark.parametrize("n_utts", [1, 6, 7, 8, 149, 1036, 2100])
def test_batch_handoff_sizes(ctx, n_utts):
    """The pipelined kernel hands records to its tail warps 7 at a time, double-buffered, with padded and closing
    batches at the end of a CTA's work: batch sizes around those boundaries (per CTA: 1 utterance ... ~14), ragged
    lengths incl. utterances shorter than a frame, must give what the oracle gives for every utterance."""
    from dsp_audioreclabs_b200 import batch
    from oracle import frontend_oracle as fo, synth
    lens = list(synth.ragged_lengths(22, 0.15, 1.2, seed=n_utts)) + [100, 255]
    base = [synth.utterance_pcm(i, int(n), seed0=400 + n_utts) for i, n in enumerate(lens)]
    utts = [base[(5 * i) % len(base)] for i in range(n_utts)]
    s, o, l = batch.pack_aligned(utts)
    ctx.set_tuning("pcm_variant", PIPE)
    try:
        res = batch.frontend_batch(s, o, 256, 128, "hanning", emit_epd_lists=True, lengths=l, ctx=ctx)
    finally:
        ctx.set_tuning("pcm_variant", -1)
    refs = [fo.frontend_utterance(u, 256, 128, "hanning") for u in base]
    for b in range(n_utts):
        r = refs[(5 * b) % len(base)]
        assert (int(res.start[b]), int(res.end[b]), int(res.n_frames[b])) == (r["start"], r["end"], len(r["zcr"])), b
        if len(r["zcr"]):
            e, m, z = res.frames(b)
            assert np.array_equal(z.astype(np.float64), r["zcr"]), b
            assert np.allclose(e, r["energy"], rtol=1e-5, atol=0) and np.allclose(m, r["magnitude"], rtol=1e-5, atol=0), b


@pytest.mark.parametrize("lead", [0, 1, 2, 3, 4, 5, 6, 7])
def test_pipelined_kernel_takes_misaligned_utterances(ctx, lead):
    """A packed (CSR) batch whose utterances start at ARBITRARY sample offsets (odd lengths, `lead` samples in front of
    the first one): the pipelined kernel streams each utterance from the 16-byte boundary below its start and carries
    the 0..7-sample offset through its group sums -- no utterance is handed to the float64 replay because of its address
    (the replayed set equals that of the 16-byte aligned packing, i.e. the degenerate one-frame
    utterances whose only frame sits ON its own thresholds), and every result equals the oracle's.  Also without the
    pcm_variant knob: the automatic choice for 256 / 128 is the pipelined kernel whatever the alignment."""
    from dsp_audioreclabs_b200 import batch
    from oracle import frontend_oracle as fo, synth
    lens = [9001, 12347, 7777, 15003, 8192, 1000, 5000, 30001, 44100, 20011, 333, 25000, 26001, 40000, 12000, 13001, 2049, 2047,
            4096, 4097, 7, 15, 16, 17, 255, 256, 257, 0, 1, 6143, 6145]
    utts = [synth.utterance_pcm(70 + i, n, seed0=3) for i, n in enumerate(lens)] * 2
    samples, off = pack(utts)
    refs = fo.frontend_batch(samples, off, 256, 128, "hamming")
    s_al, o_al, l_al = batch.pack_aligned(utts)
    ctx.set_tuning("pcm_variant", PIPE)
    try:
        aligned = batch.frontend_batch(s_al, o_al, 256, 128, "hamming", lengths=l_al, ctx=ctx)
    finally:
        ctx.set_tuning("pcm_variant", -1)
    replayed_when_aligned = aligned.status >= 0x100
    assert all(len(u) < 2 * 256 for u, r in zip(utts, replayed_when_aligned) if r)
    samples = np.concatenate([np.full(lead, 12345, np.int16), samples, np.full(9, -4321, np.int16)])
    off = off + lead
    for forced in (True, False):
        if forced:
            ctx.set_tuning("pcm_variant", PIPE)
        try:
            res = batch.frontend_batch(samples, off, 256, 128, "hamming", emit_epd_lists=True, ctx=ctx)
        finally:
            ctx.set_tuning("pcm_variant", -1)
        assert np.array_equal(res.status >= 0x100, replayed_when_aligned)      # nothing replayed because of its address
        for b, r in enumerate(refs):
            if r is None:                                         # the reference raises (empty utterance): status says so
                assert (int(res.status[b]) & 0xff) != 0, (lead, b)
                continue
            assert (int(res.status[b]) & 0xff) == 0, (lead, b)
            assert (int(res.start[b]), int(res.end[b]), int(res.n_frames[b])) == (r["start"], r["end"], len(r["zcr"])), (lead, b)
            e, m, z = res.frames(b)
            assert np.array_equal(z.astype(np.float64), r["zcr"]), (lead, b)
            assert np.allclose(e, r["energy"], rtol=1e-5, atol=0) and np.allclose(m, r["magnitude"], rtol=1e-5, atol=0), (lead, b)
            el, zl = res.epd_lists(b)
            assert np.array_equal(zl.astype(np.float64), r["zcr_list"]) and np.allclose(el, r["energy_list"], rtol=1e-12, atol=0), (lead, b)


# BASELINE configs[3] / SURVEY.md 8(d) config 4: the power-of-two pairs and the REAL grids of ablation_study.py
# (config.py:81,85 x train_model.py:45-46: frame lengths at shift 441, shifts at length 1102 -- up to fl = 2205 and
# fs > fl), one launch per configuration, whichever kernel the library routes the geometry to
CONFIG4 = ([(64, 32), (128, 64), (256, 128), (512, 256), (1024, 512), (2048, 1024)]
           + [(int(44100 * ms / 1000), 441) for ms in (8, 10, 12, 15, 18, 20, 25, 30, 35, 40, 45, 50)]
           + [(1102, int(44100 * ms / 1000)) for ms in (3, 5, 7, 8, 10, 12, 15, 18, 20, 25, 30)])


@pytest.mark.parametrize("fl,fs", CONFIG4)
def test_config4_geometries_match_the_oracle(ctx, fl, fs):
    from dsp_audioreclabs_b200 import batch
    from oracle import frontend_oracle as fo, synth
    assert (1102, 441) in CONFIG4 and (2205, 441) in CONFIG4 and (1102, 1323) in CONFIG4 and (1102, 132) in CONFIG4
    lens = [22050, 30011, 17000, 44100, 9001, 26000]
    utts = [synth.utterance_pcm(300 + i, n, seed0=17) for i, n in enumerate(lens)]
    samples, off = pack(utts)                                     # packed CSR: odd lengths, arbitrary alignment
    win = ("rectangular", "hamming", "hanning")[(fl + fs) % 3]
    refs = [fo.frontend_utterance(u, fl, fs, win) for u in utts]
    for variant in (-1, PIPE, RESIDENT):          # the automatic route, and both kernels forced (PIPE falls back where it does not fit)
        ctx.set_tuning("pcm_variant", variant)
        try:
            res = batch.frontend_batch(samples, off, fl, fs, win, emit_epd_lists=True, ctx=ctx)
        finally:
            ctx.set_tuning("pcm_variant", -1)
        for b, r in enumerate(refs):
            assert (int(res.start[b]), int(res.end[b]), int(res.n_frames[b])) == (r["start"], r["end"], len(r["zcr"])), (fl, fs, b, variant)
            e, m, z = res.frames(b)
            assert np.array_equal(z.astype(np.float64), r["zcr"]), (fl, fs, b, variant)
            assert np.allclose(e, r["energy"], rtol=RTOL_F32, atol=0) and np.allclose(m, r["magnitude"], rtol=RTOL_F32, atol=0), (fl, fs, b, variant)
            el, zl = res.epd_lists(b)
            assert np.array_equal(zl.astype(np.float64), r["zcr_list"]) and np.allclose(el, r["energy_list"], rtol=RTOL_EPD, atol=0)
            assert_stats_close(res.stats[b], r["stats"], r)
