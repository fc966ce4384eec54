#!/usr/bin/env python3
"""Generate tests/golden/knn_seq_golden.npz by running the UNMODIFIED reference: the sequence-feature
variant of compare_feature_methods.py:77-176 -- per-frame (energy, zcr)
sequences, zero-padded to the longest one, flattened to D = 2 * max_len, z-scored with the train
statistics, classified by the reference's KNN wrapper (sklearn picks brute force for D > 15).

Build container only (needs /root/reference):   python oracle/gen_golden_seq.py
"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    tmp = tempfile.mkdtemp(prefix="refcopy_")
    dst = os.path.join(tmp, "ref")
    shutil.copytree(REF, dst)
    sys.path.insert(0, dst)
    from src import audio_processing as ap, feature_extraction as fe, models as mo
    from sklearn.model_selection import train_test_split

    rng = np.random.default_rng(2024)
    ks = {}
    for tag, (n_utts, fl, fs, both) in {"seq_default": (240, 1102, 441, True), "seq_256": (200, 256, 128, True),
                                        "seq3_default": (120, 1102, 441, False)}.items():
        lens = rng.uniform(0.55, 0.95, n_utts)
        seqs, labels, pcms = [], [], []
        for i in range(n_utts):
            pcm = synth.utterance_pcm(i, int(lens[i] * 44100), seed0=9000)
            x = ap.preprocess(pcm / 32768.0)                                   # load_wav's scaling (:35-38) + preprocess
            start, end, _, _ = ap.endpoint_detection(x, fl, fs)
            frames = ap.frame_signal(x[start:end], fl, fs, "hamming")
            seq, _ = fe.extract_features_from_frames(frames, method="sequence", use_only_energy_zcr=both)
            seqs.append(seq)
            labels.append(i % 10)
            pcms.append(pcm)
        max_len = max(len(s) for s in seqs)                                     # compare_feature_methods.py:106-113
        flat = np.array([fe.pad_or_truncate_sequence(s, max_len) for s in seqs]).reshape(n_utts, -1)
        y = np.array(labels)
        xtr, xte, ytr, yte = train_test_split(flat, y, test_size=0.2, random_state=42, stratify=y)   # :128-134
        xtr_n, mu, sd = fe.normalize_features(xtr)                              # :140-144
        xte_n, _, _ = fe.normalize_features(xte, mu, sd)
        clf = mo.create_classifier("knn", n_neighbors=3)                        # :165,173
        clf.fit(xtr_n, ytr)
        pred = clf.predict(xte_n)
        dist, idx = clf.model.kneighbors(xte_n)
        ks[f"{tag}/frame"] = np.array([fl, fs, max_len, int(both)])
        ks[f"{tag}/pcm_lengths"] = np.array([len(p) for p in pcms])
        ks[f"{tag}/flat"], ks[f"{tag}/labels"] = flat, y
        ks[f"{tag}/train"], ks[f"{tag}/train_labels"] = xtr, ytr
        ks[f"{tag}/query"], ks[f"{tag}/query_labels"] = xte, yte
        ks[f"{tag}/train_norm"], ks[f"{tag}/query_norm"] = xtr_n, xte_n
        ks[f"{tag}/pred"], ks[f"{tag}/nbr_idx"], ks[f"{tag}/nbr_dist"] = pred, idx, dist
        ks[f"{tag}/fit_method"] = np.array(clf.model._fit_method)
        print(tag, "D =", flat.shape[1], "fit_method =", clf.model._fit_method, "accuracy =", float(np.mean(pred == yte)))
    np.savez_compressed(os.path.join(OUT, "knn_seq_golden.npz"), **ks)
    shutil.rmtree(tmp, ignore_errors=True)
    print("knn_seq_golden.npz", os.path.getsize(os.path.join(OUT, "knn_seq_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
