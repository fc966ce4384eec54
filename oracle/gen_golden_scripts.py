#!/usr/bin/env python3
"""Golden outputs of the reference's own driver scripts on BASELINE configs[0].

Run in the build container only:   python oracle/gen_golden_scripts.py

Writes tests/golden/scripts_golden.npz:
  X_sorted / y_sorted   the (200, 15) matrix SpeechRecognitionExperiment.load_dataset('hamming') builds
                        (experiments/run_experiments.py:45-126), rows sorted inside each class so that the
                        comparison does not depend on glob()'s directory order
  feature_analysis      text of results/exp3_feature_analysis/feature_analysis.txt written by
                        `run.py --experiment feature --window-type hamming` (run.py:128-130)
  ablation_json         results.json of `ablation_study.py --experiment frame_length --classifier knn`
                        (ablation_study.py:112-193), timestamp removed
The scripts run UNMODIFIED from a throw-away copy of /root/reference (config.py:25-26 writes next to
itself), with tests/stubs (no-op matplotlib / seaborn) on PYTHONPATH.  Accuracies in ablation_json depend on
glob order (train_test_split sees the rows in file order), so the GPU test compares them only against a
live run of oracle/_ref on the same directory; they are stored here for the record.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

REF = "/root/reference"
STUBS = os.path.join(ROOT, "tests", "stubs")


def sorted_rows(X, y):
    order = np.lexsort(tuple(X[:, j] for j in range(X.shape[1] - 1, -1, -1)) + (y,))
    return X[order], y[order]


def main():
    tmp = tempfile.mkdtemp(prefix="refscripts_")
    ref = os.path.join(tmp, "ref")
    shutil.copytree(REF, ref)
    data = os.path.join(tmp, "data")
    synth.write_config1_dataset(data)
    env = dict(os.environ, PYTHONPATH=STUBS, SPEECH_DATA_DIR=data, OMP_NUM_THREADS="1")
    subprocess.run([sys.executable, "run.py", "--experiment", "feature", "--window-type", "hamming"], cwd=ref, env=env, check=True,
                   stdout=subprocess.DEVNULL)
    text = open(os.path.join(ref, "results", "exp3_feature_analysis", "feature_analysis.txt"), encoding="utf-8").read()
    subprocess.run([sys.executable, "ablation_study.py", "--experiment", "frame_length", "--classifier", "knn"], cwd=ref, env=env,
                   check=True, stdout=subprocess.DEVNULL)
    abl = json.load(open(os.path.join(ref, "results", "ablation_frame_length", "results.json"), encoding="utf-8"))
    abl.pop("timestamp", None)
    # X through the reference's own class, in this process
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r); import config; "
            "from experiments.run_experiments import SpeechRecognitionExperiment as E; "
            "e = E(config.DATA_DIR, %r); X, y, names = e.load_dataset('hamming'); np.savez(%r, X=X, y=y, names=np.array(names))")
    xp = os.path.join(tmp, "x.npz")
    subprocess.run([sys.executable, "-c", code % (ref, STUBS, os.path.join(tmp, "res"), xp)], env=env, check=True, stdout=subprocess.DEVNULL)
    z = np.load(xp)
    Xs, ys = sorted_rows(z["X"], z["y"])
    out = os.path.join(ROOT, "tests", "golden", "scripts_golden.npz")
    np.savez_compressed(out, X_sorted=Xs, y_sorted=ys, names=z["names"], feature_analysis=np.array(text),
                        ablation_json=np.array(json.dumps(abl, ensure_ascii=False)))
    print("wrote", out, Xs.shape, "ablation keys:", list(abl.get("results", abl).keys())[:4])
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
