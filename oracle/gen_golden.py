#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py

The reference is imported from a throw-away copy under /tmp because its
``config.py`` creates a directory at import time (config.py:25-26) and the
mount is read-only.  The reference ships no tests or golden vectors of its
own, so these fixtures -- outputs of the reference's
``src.audio_processing`` / ``src.feature_extraction`` / ``src.models`` on
seeded synthetic inputs -- are what pins the oracle and the CUDA path.
"""
import os
import shutil
import sys
import tempfile
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

# (frame_length, frame_shift): BASELINE config 2, the reference default (config.py:39-40),
# no overlap, a gapped pair and the extremes of the real ablation grids
CONFIGS = [(256, 128), (1102, 441), (64, 64), (1102, 1323), (2205, 441), (352, 441), (1102, 132)]
WINDOWS = ["rectangular", "hamming", "hanning"]


def import_reference():
    tmp = tempfile.mkdtemp(prefix="refcopy_")
    dst = os.path.join(tmp, "ref")
    shutil.copytree(REF, dst)
    sys.path.insert(0, dst)
    import src.audio_processing as ap
    import src.feature_extraction as fe
    import src.models as mo
    return tmp, dst, ap, fe, mo


def edge_cases(rng):
    """Named PCM inputs that exercise the reference's branches."""
    cases = {}
    cases["zeros"] = np.zeros(5000, dtype=np.int16)                       # :207-209, :73-75
    cases["constant"] = np.full(5000, 1234, dtype=np.int16)               # DC only
    cases["short100"] = synth.utterance_pcm(3, 100)                       # L < fl for most configs
    cases["len256"] = synth.utterance_pcm(4, 256)                         # exactly one 256-frame
    cases["len2000"] = synth.utterance_pcm(5, 2000)                       # F1 < 10 -> noise_frames 0
    t = np.arange(12000)
    tone = 0.5 * np.sin(2 * np.pi * 441.0 * t / 44100.0)                  # period 100 samples
    tone[:3000] = 0
    tone[9000:] = 0
    cases["gated_tone"] = np.trunc(tone * 32768).astype(np.int16)
    sq = np.zeros(12000)
    sq[2000:6000] = np.where((t[2000:6000] // 64) % 2 == 0, 0.5, -0.5)    # exact-tie energies
    sq[6000:9000] = np.where((t[6000:9000] // 64) % 2 == 0, 0.25, -0.25)
    cases["square_steps"] = np.trunc(sq * 32768).astype(np.int16)
    cases["loud_noise"] = np.clip(rng.standard_normal(9000) * 9000, -32768, 32767).astype(np.int16)
    cases["lsb_noise"] = rng.integers(-1, 2, 9000).astype(np.int16)        # +-1 LSB around zero
    big_dc = (rng.integers(-3, 4, 9000) + 20000).astype(np.int16)          # tiny signal on huge DC
    big_dc[3000:5000] += (200 * np.sin(np.arange(2000) * 0.2)).astype(np.int16)
    cases["big_dc"] = big_dc
    return cases


def run_reference(ap, fe, audio, fl, fs, window):
    """process_audio_file minus load_wav (src/audio_processing.py:364-394) + features."""
    x = ap.preprocess(audio)
    s, e, el, zl = ap.endpoint_detection(x, fl, fs, 0.5, 0.1, 1.5)
    seg = x[s:e]
    rec = {"start": s, "end": e, "energy_list": np.asarray(el, dtype=np.float64),
           "zcr_list": np.asarray(zl, dtype=np.float64)}
    if len(seg) == 0:
        rec["error"] = 1
        return rec
    frames = ap.frame_signal(seg, fl, fs, window)
    try:
        ff = fe.extract_frame_features(frames)
    except ValueError:
        rec["error"] = 2
        return rec
    vec, _ = fe.extract_statistical_features(ff)
    rec.update(error=0, n_frames=len(frames), energy=ff["energy"], magnitude=ff["magnitude"],
               zcr=ff["zcr"], stats=vec)
    return rec


def main():
    os.makedirs(OUT, exist_ok=True)
    tmp, dst, ap, fe, mo = import_reference()
    rng = np.random.default_rng(20261018)
    store = {}

    # ---- inputs ---------------------------------------------------------------
    utts = {}
    for i, n in enumerate([13230, 11025, 15000, 9999, 17640, 12345]):
        utts[f"syn{i}"] = synth.utterance_pcm(100 + i, n, seed0=777)
    utts.update(edge_cases(rng))
    names = sorted(utts)
    store["names"] = np.array(names)
    store["configs"] = np.array(CONFIGS, dtype=np.int64)
    store["windows"] = np.array(WINDOWS)
    for nme in names:
        store[f"pcm/{nme}"] = utts[nme]
    # full-length seeded utterances: only their CRC is stored (regenerated in the tests)
    full_idx = [0, 1, 2, 3]
    store["full_index"] = np.array(full_idx)
    store["full_crc"] = np.array([int(np.bitwise_xor.reduce(
        synth.utterance_pcm(i).astype(np.int64) * (np.arange(44100) + 1))) for i in full_idx])

    # ---- front end ------------------------------------------------------------
    for ci, (fl, fs) in enumerate(CONFIGS):
        for nme in names + [f"full{i}" for i in full_idx]:
            pcm = synth.utterance_pcm(int(nme[4:])) if nme.startswith("full") else utts[nme]
            audio = pcm / 32768.0                                   # load_wav :37-38
            for wi, win in enumerate(WINDOWS):
                rec = run_reference(ap, fe, audio, fl, fs, win)
                base = f"fe/{ci}/{nme}/{win}"
                for k in ("error", "n_frames", "energy", "magnitude", "zcr", "stats"):
                    if k in rec:
                        store[f"{base}/{k}"] = np.asarray(rec[k])
                if wi == 0:
                    b2 = f"epd/{ci}/{nme}"
                    store[f"{b2}/start"] = np.asarray(rec["start"])
                    store[f"{b2}/end"] = np.asarray(rec["end"])
                    store[f"{b2}/energy_list"] = rec["energy_list"]
                    store[f"{b2}/zcr_list"] = rec["zcr_list"]

    # ---- do_endpoint_detection=False framing (zero-padded last frame) -----------
    for ci, (fl, fs) in enumerate(CONFIGS[:4]):
        for nme in ("syn0", "syn3", "short100", "len256"):
            x = ap.preprocess(utts[nme] / 32768.0)
            for win in WINDOWS:
                frames = ap.frame_signal(x, fl, fs, win)
                ff = fe.extract_frame_features(frames)
                vec, _ = fe.extract_statistical_features(ff)
                base = f"noepd/{ci}/{nme}/{win}"
                store[f"{base}/n_frames"] = np.asarray(len(frames))
                for k in ("energy", "magnitude", "zcr"):
                    store[f"{base}/{k}"] = ff[k]
                store[f"{base}/stats"] = vec

    # ---- float64 (non-PCM) input through the per-call surface -------------------
    xf = rng.standard_normal(8000) * np.hanning(8000) * 0.3 + 0.05
    store["float/x"] = xf
    store["float/preprocessed"] = ap.preprocess(xf)
    s, e, el, zl = ap.endpoint_detection(ap.preprocess(xf), 256, 128)
    store["float/start"], store["float/end"] = np.asarray(s), np.asarray(e)
    store["float/energy_list"], store["float/zcr_list"] = el, zl
    fr = ap.frame_signal(ap.preprocess(xf)[s:e], 256, 128, "hamming")
    store["float/frames"] = fr
    ff = fe.extract_frame_features(fr)
    store["float/energy"], store["float/magnitude"], store["float/zcr"] = ff["energy"], ff["magnitude"], ff["zcr"]
    store["float/stats"] = fe.extract_statistical_features(ff)[0]
    seq, _ = fe.extract_features_from_frames(fr, method="sequence", use_only_energy_zcr=True)
    store["float/sequence2"] = seq
    store["float/sequence2_pad"] = fe.pad_or_truncate_sequence(seq, len(seq) + 7)
    store["float/sequence2_cut"] = fe.pad_or_truncate_sequence(seq, 5)
    for win in WINDOWS:
        for n in (1, 2, 3, 64, 255, 256, 1102):
            store[f"window/{win}/{n}"] = ap.create_window(win, n)

    # ---- WAV decode: 8-bit, 16-bit, stereo (src/audio_processing.py:9-46) -------
    wavdir = os.path.join(tmp, "wav")
    os.makedirs(wavdir)

    def write_wav(name, data, width, channels):
        with wave.open(os.path.join(wavdir, name), "wb") as w:
            w.setnchannels(channels)
            w.setsampwidth(width)
            w.setframerate(44100)
            w.writeframes(data.tobytes())

    mono16 = utts["syn1"]
    st16 = np.stack([utts["syn1"][:9000], utts["syn2"][:9000]], axis=1).reshape(-1)
    mono8 = ((utts["syn0"].astype(np.int32) >> 8) + 128).astype(np.uint8)
    st8 = np.stack([mono8[:9000], mono8[1000:10000]], axis=1).reshape(-1)
    for name, data, width, ch in (("m16.wav", mono16, 2, 1), ("s16.wav", st16, 2, 2),
                                  ("m8.wav", mono8, 1, 1), ("s8.wav", st8, 1, 2)):
        write_wav(name, data, width, ch)
        audio, sr = ap.load_wav(os.path.join(wavdir, name))
        key = name[:-4]
        store[f"wav/{key}/raw"] = data
        store[f"wav/{key}/width_channels"] = np.array([width, ch])
        store[f"wav/{key}/audio"] = audio
        frames, _, meta = ap.process_audio_file(os.path.join(wavdir, name), 1102, 441, "hamming")
        vec, _ = fe.extract_features_from_frames(frames, method="statistical")
        store[f"wav/{key}/start_end"] = np.array([meta["start_point"], meta["end_point"]])
        store[f"wav/{key}/stats"] = vec

    np.savez_compressed(os.path.join(OUT, "frontend_golden.npz"), **store)

    # ---- KNN (sklearn through the reference's wrapper, src/models.py:226-246) ---
    ks = {}
    for tag, (ntr, nq, d, ncls) in {"d15": (4000, 600, 15, 10), "d40": (1500, 300, 40, 7)}.items():
        centers = rng.standard_normal((ncls, d)) * 1.5
        ytr = rng.integers(0, ncls, ntr)
        xtr = centers[ytr] + rng.standard_normal((ntr, d))
        yq = rng.integers(0, ncls, nq)
        xq = centers[yq] + rng.standard_normal((nq, d))
        xtr_n, mu, sd = fe.normalize_features(xtr)
        xq_n, _, _ = fe.normalize_features(xq, mu, sd)
        clf = mo.create_classifier("knn", n_neighbors=3)
        clf.fit(xtr_n, ytr)
        pred = clf.predict(xq_n)
        dist, idx = clf.model.kneighbors(xq_n)
        ev = clf.evaluate(xq_n, yq)
        ks[f"{tag}/train"], ks[f"{tag}/train_labels"] = xtr, ytr
        ks[f"{tag}/query"], ks[f"{tag}/query_labels"] = xq, yq
        ks[f"{tag}/mean"], ks[f"{tag}/std"] = mu, sd
        ks[f"{tag}/train_norm"], ks[f"{tag}/query_norm"] = xtr_n, xq_n
        ks[f"{tag}/pred"], ks[f"{tag}/nbr_idx"], ks[f"{tag}/nbr_dist"] = pred, idx, dist
        ks[f"{tag}/accuracy"] = np.asarray(ev["accuracy"])
        ks[f"{tag}/confusion"] = ev["confusion_matrix"]
        ks[f"{tag}/fit_method"] = np.array(clf.model._fit_method)
    np.savez_compressed(os.path.join(OUT, "knn_golden.npz"), **ks)

    # ---- the Python call surface itself -------------------------
    import inspect
    import json
    import config as ref_config
    sig = {}
    for mod, names in ((ap, ["load_wav", "remove_dc", "normalize_audio", "preprocess", "compute_short_time_energy",
                             "compute_short_time_magnitude", "compute_zero_crossing_rate", "endpoint_detection",
                             "create_window", "frame_signal", "process_audio_file"]),
                       (fe, ["extract_frame_features", "compute_statistics", "extract_statistical_features",
                             "extract_features_from_frames", "pad_or_truncate_sequence", "normalize_features"]),
                       (mo, ["create_classifier"])):
        for n in names:
            sig[f"{mod.__name__}.{n}"] = str(inspect.signature(getattr(mod, n)))
    cfg = {k: getattr(ref_config, k) for k in dir(ref_config)
           if k.isupper() and k not in ("BASE_DIR", "DATA_DIR", "RESULTS_DIR", "DATASET_PATHS", "DATASET_TYPE")}
    with open(os.path.join(OUT, "surface.json"), "w") as f:
        json.dump({"signatures": sig, "config": cfg,
                   "config_names": sorted(k for k in dir(ref_config) if k.isupper())}, f, indent=1, sort_keys=True)

    shutil.rmtree(tmp, ignore_errors=True)
    for f in ("frontend_golden.npz", "knn_golden.npz", "surface.json"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
