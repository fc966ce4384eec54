"""CPU: the oracle restatement against fixtures produced by the reference itself
(oracle/gen_golden.py).  Bit-exact: the oracle uses the same NumPy operations."""
import numpy as np
import pytest

from conftest import golden_names, golden_pcm
from oracle import frontend_oracle as fo
from oracle import knn_oracle as ko
from oracle import synth


def test_synth_generator_is_stable(golden_fe):
    g = golden_fe
    for i, crc in zip(g["full_index"], g["full_crc"]):
        pcm = synth.utterance_pcm(int(i)).astype(np.int64)
        assert int(np.bitwise_xor.reduce(pcm * (np.arange(44100) + 1))) == int(crc)


@pytest.mark.parametrize("ci", range(7))
def test_frontend_oracle_matches_reference(golden_fe, ci):
    g = golden_fe
    fl, fs = (int(v) for v in g["configs"][ci])
    nontrivial = 0
    for name in golden_names(g):
        pcm = golden_pcm(g, name)
        for win in g["windows"]:
            base = f"fe/{ci}/{name}/{win}"
            err = int(g[base + "/error"])
            try:
                r = fo.frontend_utterance(pcm, fl, fs, str(win))
            except ValueError:
                assert err != 0, base
                continue
            assert err == 0, base
            assert r["start"] == int(g[f"epd/{ci}/{name}/start"])
            assert r["end"] == int(g[f"epd/{ci}/{name}/end"])
            nontrivial += (r["start"] > 0) or (r["end"] < len(pcm))
            assert np.array_equal(r["energy_list"], g[f"epd/{ci}/{name}/energy_list"])
            assert np.array_equal(r["zcr_list"], g[f"epd/{ci}/{name}/zcr_list"])
            assert r["n_frames"] == int(g[base + "/n_frames"])
            for k in ("energy", "magnitude", "zcr", "stats"):
                assert np.array_equal(r[k], g[f"{base}/{k}"]), (base, k)
    assert nontrivial > 0      # the fixtures do exercise real endpoint trimming


def test_no_epd_framing_with_zero_padding(golden_fe):
    g = golden_fe
    for ci in range(4):
        fl, fs = (int(v) for v in g["configs"][ci])
        for name in ("syn0", "syn3", "short100", "len256"):
            for win in g["windows"]:
                r = fo.frontend_utterance(g[f"pcm/{name}"], fl, fs, str(win), do_endpoint_detection=False)
                base = f"noepd/{ci}/{name}/{win}"
                assert r["n_frames"] == int(g[base + "/n_frames"])
                assert r["n_frames"] == fo.feature_frame_count(len(g[f"pcm/{name}"]), fl, fs)
                for k in ("energy", "magnitude", "zcr", "stats"):
                    assert np.array_equal(r[k], g[f"{base}/{k}"])


def test_float_input_and_sequence_helpers(golden_fe):
    g = golden_fe
    x = g["float/x"]
    z = fo.preprocess(x)
    assert np.array_equal(z, g["float/preprocessed"])
    s, e, el, zl = fo.endpoint_detection(z, 256, 128)
    assert (s, e) == (int(g["float/start"]), int(g["float/end"]))
    assert np.array_equal(el, g["float/energy_list"]) and np.array_equal(zl, g["float/zcr_list"])
    fr = fo.frame_signal(z[s:e], 256, 128, "hamming")
    assert np.array_equal(fr, g["float/frames"])
    ff = fo.frame_features(fr)
    assert np.array_equal(fo.statistical_vector(ff), g["float/stats"])
    seq = fo.sequence_matrix(ff, use_only_energy_zcr=True)
    assert np.array_equal(seq, g["float/sequence2"])
    assert np.array_equal(fo.pad_or_truncate(seq, len(seq) + 7), g["float/sequence2_pad"])
    assert np.array_equal(fo.pad_or_truncate(seq, 5), g["float/sequence2_cut"])
    for win in g["windows"]:
        for n in (1, 2, 3, 64, 255, 256, 1102):
            assert np.array_equal(fo.make_window(str(win), n), g[f"window/{win}/{n}"])
    with pytest.raises(ValueError):
        fo.make_window("blackman", 8)
    with pytest.raises(ValueError):
        fo.frame_features(np.zeros((0, 256)))


def test_wav_decode_rules(golden_fe):
    g = golden_fe
    for key in ("m16", "s16", "m8", "s8"):
        raw = g[f"wav/{key}/raw"]
        width, ch = (int(v) for v in g[f"wav/{key}/width_channels"])
        x = fo.pcm_to_float(raw)
        if ch == 2:
            x = fo.stereo_to_mono(x)
        assert np.array_equal(x, g[f"wav/{key}/audio"])
        z = fo.preprocess(x)
        s, e, _, _ = fo.endpoint_detection(z, 1102, 441)
        assert [s, e] == list(g[f"wav/{key}/start_end"])
        st = fo.statistical_vector(fo.frame_features(fo.frame_signal(z[s:e], 1102, 441, "hamming")))
        assert np.array_equal(st, g[f"wav/{key}/stats"])


@pytest.mark.parametrize("tag", ["d15", "d40"])
def test_knn_oracle_matches_sklearn(golden_knn, tag):
    k = golden_knn
    xn, mu, sd = fo.zscore(k[f"{tag}/train"])
    assert np.array_equal(xn, k[f"{tag}/train_norm"]) and np.array_equal(mu, k[f"{tag}/mean"])
    qn, _, _ = fo.zscore(k[f"{tag}/query"], mu, sd)
    assert np.array_equal(qn, k[f"{tag}/query_norm"])
    idx, d2 = ko.knn_topk(xn, qn, 3)
    assert np.array_equal(idx, k[f"{tag}/nbr_idx"])
    assert np.allclose(np.sqrt(d2), k[f"{tag}/nbr_dist"], rtol=1e-12, atol=0)
    pred = ko.knn_predict(xn, k[f"{tag}/train_labels"], qn, 3)
    assert np.array_equal(pred, k[f"{tag}/pred"])


def test_knn_merge_of_row_shards_equals_single_rank(golden_knn):
    """CPU model of the multi-GPU exchange: partition the train rows, local top-k per shard,
    concatenate, merge."""
    k = golden_knn
    xn, qn, y = k["d15/train_norm"], k["d15/query_norm"][:100], k["d15/train_labels"]
    ref_idx, ref_d = ko.knn_topk(xn, qn, 3)
    for shards in (2, 3, 8):
        bounds = np.linspace(0, len(xn), shards + 1).astype(int)
        cd, ci = [], []
        for r in range(shards):
            i, d = ko.knn_topk(xn[bounds[r]:bounds[r + 1]], qn, 3)
            cd.append(d)
            ci.append(i + bounds[r])
        mi, md = ko.merge_candidates(np.stack(cd), np.stack(ci), 3)
        assert np.array_equal(mi, ref_idx) and np.array_equal(md, ref_d)
        assert np.array_equal(ko.vote(y[mi], np.unique(y)), k["d15/pred"][:100])


def test_sequence_knn_fixture_matches_the_oracle(golden_seq):
    """compare_feature_methods.py's sequence variant (sklearn brute force, D = 76 / 258 / 105): the oracle's
    neighbours, distances and votes equal the reference's; the sequence assembly restates :106-123."""
    from oracle import frontend_oracle as fo, knn_oracle as ko
    g = golden_seq
    for tag in ("seq_default", "seq_256", "seq3_default"):
        assert str(g[f"{tag}/fit_method"]) == "brute"
        xn, mu, sd = fo.zscore(g[f"{tag}/train"])
        qn, _, _ = fo.zscore(g[f"{tag}/query"], mu, sd)
        assert np.array_equal(xn, g[f"{tag}/train_norm"]) and np.array_equal(qn, g[f"{tag}/query_norm"])
        idx, d2 = ko.knn_topk(xn, qn, 3)
        assert np.array_equal(idx, g[f"{tag}/nbr_idx"])
        assert np.allclose(np.sqrt(d2), g[f"{tag}/nbr_dist"], rtol=1e-12, atol=0)
        assert np.array_equal(ko.knn_predict(xn, g[f"{tag}/train_labels"], qn, 3), g[f"{tag}/pred"])
