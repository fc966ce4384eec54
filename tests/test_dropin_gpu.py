"""GPU: the drop-in modules behave like the reference's on WAV files (golden fixtures written
back to disk) and on a small class-per-directory data set walked the way the reference's
load_dataset does (experiments/run_experiments.py:64-114)."""
import importlib
import os
import sys
import wave

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
DROPIN = os.path.join(ROOT, "dsp_audioreclabs_b200", "dropin")


@pytest.fixture(scope="module")
def mods(tmp_path_factory):
    os.environ["DSP_RESULTS_DIR"] = str(tmp_path_factory.mktemp("results"))
    sys.path.insert(0, DROPIN)
    for k in list(sys.modules):
        if k == "config" or k == "src" or k.startswith("src."):
            del sys.modules[k]
    m = {n: importlib.import_module(n) for n in ("config", "src.audio_processing", "src.feature_extraction", "src.models")}
    yield m
    sys.path.remove(DROPIN)
    for k in list(sys.modules):
        if k == "config" or k == "src" or k.startswith("src."):
            del sys.modules[k]


def write_wav(path, data, width, channels):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(width)
        w.setframerate(44100)
        w.writeframes(np.ascontiguousarray(data).tobytes())


def test_wav_files_like_the_reference(mods, golden_fe, tmp_path):
    ap, fe = mods["src.audio_processing"], mods["src.feature_extraction"]
    g = golden_fe
    for key in ("m16", "s16", "m8", "s8"):
        width, ch = (int(v) for v in g[f"wav/{key}/width_channels"])
        p = tmp_path / f"{key}.wav"
        write_wav(p, g[f"wav/{key}/raw"], width, ch)
        audio, sr = ap.load_wav(str(p))
        assert sr == 44100 and np.array_equal(audio, g[f"wav/{key}/audio"])
        frames, sr, meta = ap.process_audio_file(str(p), 1102, 441, "hamming")
        assert [meta["start_point"], meta["end_point"]] == list(g[f"wav/{key}/start_end"])
        assert set(meta) == {"original_length", "sample_rate", "start_point", "end_point", "energy_list",
                             "zcr_list", "segmented_length", "n_frames"}
        assert isinstance(frames, np.ndarray) and frames.shape == (meta["n_frames"], 1102)
        vec, names = fe.extract_features_from_frames(frames, method="statistical")
        assert len(names) == 15 and names[0] == "energy_mean" and names[-1] == "zcr_median"
        # float64 replay kernel: the reference's own float64 values (the per-file chain must not be fp32-derived)
        assert vec.dtype == np.float64 and np.array_equal(vec, g[f"wav/{key}/stats"])
        # the dense frames are real data too: recomputing from them gives the same features
        vec2, _ = fe.extract_features_from_frames(np.array(frames), method="statistical")
        assert np.array_equal(vec2, g[f"wav/{key}/stats"])
    bad = tmp_path / "w3.wav"
    write_wav(bad, np.zeros(300, np.uint8), 3, 1)
    with pytest.raises(ValueError):
        ap.load_wav(str(bad))


def test_dataset_walk_matches_oracle_and_knn_pipeline(mods, tmp_path):
    """Config-1 style set (class dirs, ~1 s clips), the reference's per-file loop on top of the
    drop-in functions, then z-score + KNN through create_classifier('knn')."""
    ap, fe, mo, cfg = mods["src.audio_processing"], mods["src.feature_extraction"], mods["src.models"], mods["config"]
    from oracle import frontend_oracle as fo, knn_oracle as ko, synth
    rng = np.random.default_rng(1234)
    X, Xref, y = [], [], []
    for c in range(5):
        d = tmp_path / str(c)
        d.mkdir()
        for i in range(6):
            n = int(rng.uniform(0.8, 1.2) * 44100)
            pcm = synth.utterance_pcm(10 * i + c, n, seed0=1234)
            write_wav(d / f"{i}.wav", pcm, 2, 1)
            frames, _, _ = ap.process_audio_file(str(d / f"{i}.wav"), cfg.FRAME_LENGTH, cfg.FRAME_SHIFT, "hamming", True,
                                                 cfg.ENERGY_HIGH_RATIO, cfg.ENERGY_LOW_RATIO, cfg.ZCR_THRESHOLD_RATIO)
            vec, _ = fe.extract_features_from_frames(frames, method="statistical")
            X.append(vec); y.append(c)
            Xref.append(fo.frontend_utterance(pcm, cfg.FRAME_LENGTH, cfg.FRAME_SHIFT, "hamming")["stats"])
    X, Xref, y = np.array(X), np.array(Xref), np.array(y)
    assert X.shape == (30, 15)
    assert np.array_equal(X, Xref)          # float64 replay: bit-identical to the NumPy path
    # the walk over the 5 class folders cost one launch (6 files each were written class by class: the tree grew, so
    # every new class re-batched; a finished tree is one launch -- tests/test_scripts_gpu.py asserts that)
    assert all(kind == "tree" for kind, _ in ap.launch_log[-5:])
    tr, te = np.arange(0, 30, 2), np.arange(1, 30, 2)
    Xn, mu, sd = fe.normalize_features(Xref[tr])
    Xt, _, _ = fe.normalize_features(Xref[te], mu, sd)
    rn, rmu, rsd = fo.zscore(Xref[tr])
    assert np.array_equal(Xn, rn) and np.array_equal(Xt, fo.zscore(Xref[te], rmu, rsd)[0])
    clf = mo.create_classifier("knn", n_neighbors=cfg.KNN_N_NEIGHBORS)
    clf.fit(Xn, y[tr])
    ev = clf.evaluate(Xt, y[te])
    ref_pred = ko.knn_predict(rn, y[tr], Xt, 3)
    assert np.array_equal(ev["predictions"], ref_pred)
    assert set(ev) == {"accuracy", "predictions", "classification_report", "confusion_matrix"}
    assert ev["confusion_matrix"].sum() == len(te) and ev["accuracy"] == pytest.approx(np.mean(ref_pred == y[te]))


def test_sequence_method_and_per_frame_primitives(mods, golden_fe):
    ap, fe = mods["src.audio_processing"], mods["src.feature_extraction"]
    g = golden_fe
    fr = g["float/frames"]
    seq, none = fe.extract_features_from_frames(fr, method="sequence", use_only_energy_zcr=True)
    assert none is None and np.array_equal(seq, g["float/sequence2"])
    assert fe.extract_features_from_frames(fr, method="sequence")[0].shape == (len(fr), 3)
    assert ap.compute_short_time_energy(fr[3]) == g["float/energy"][3]
    assert ap.compute_short_time_magnitude(fr[3]) == g["float/magnitude"][3]
    assert ap.compute_zero_crossing_rate(fr[3]) == g["float/zcr"][3]
    st = fe.compute_statistics(g["float/energy"])
    assert [st[k] for k in ("mean", "std", "max", "min", "median")] == list(g["float/stats"][:5])


def test_batched_dataset_loader_equals_per_file_loop(mods, tmp_path):
    """one fused launch over the decoded tree gives the per-file results; the sweep
    helper (BASELINE config 4) re-frames without re-decoding."""
    from dsp_audioreclabs_b200 import dataset
    from oracle import frontend_oracle as fo, synth
    ap, fe, cfg = mods["src.audio_processing"], mods["src.feature_extraction"], mods["config"]
    rng = np.random.default_rng(99)
    ref = {}
    for c in range(3):
        d = tmp_path / f"class{c}"
        d.mkdir()
        for i in range(4):
            pcm = synth.utterance_pcm(7 * i + c, int(rng.uniform(0.3, 0.6) * 44100), seed0=55)
            write_wav(d / f"{i}.wav", pcm, 2, 1)
            ref[str(d / f"{i}.wav")] = pcm
    (tmp_path / ".hidden").mkdir()
    (tmp_path / "class0" / "broken.wav").write_bytes(b"not a wav")
    X, y, names, feat = dataset.load_dataset(str(tmp_path), cfg.FRAME_LENGTH, cfg.FRAME_SHIFT, "hamming")
    assert names == ["class0", "class1", "class2"] and X.shape == (12, 15) and len(feat) == 15
    assert list(np.bincount(y)) == [4, 4, 4]
    clips, labels, _, paths = dataset.decode_tree(str(tmp_path))
    for row, path in zip(X, paths):
        r = fo.frontend_utterance(ref[path], cfg.FRAME_LENGTH, cfg.FRAME_SHIFT, "hamming")
        assert np.allclose(row, r["stats"], rtol=2e-5, atol=1e-7 * np.abs(r["stats"]).max())
        frames, _, _ = ap.process_audio_file(path, cfg.FRAME_LENGTH, cfg.FRAME_SHIFT, "hamming")
        assert np.allclose(row, fe.extract_features_from_frames(frames)[0], rtol=2e-5, atol=1e-7 * np.abs(r["stats"]).max())
        assert np.array_equal(fe.extract_features_from_frames(frames)[0], r["stats"])
    sweep, _ = dataset.ablation_sweep(str(tmp_path), [(352, 441), (1102, 132), (2205, 441)])
    for (fl, fs), (Xs, ys) in sweep.items():
        assert Xs.shape == (12, 15)
        r = fo.frontend_utterance(ref[paths[0]], fl, fs, "hamming")
        assert np.allclose(Xs[0], r["stats"], rtol=2e-5, atol=1e-7 * np.abs(r["stats"]).max())
