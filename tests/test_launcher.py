"""CPU: `python -m dsp_audioreclabs_b200.run <script>` makes the reference's import statements resolve to the
drop-in for the hot-path modules and to the reference's own files for everything else (VERDICT r1 weak #1:
with bare PYTHONPATH the script directory shadows the drop-in, and `src.visualization` did not resolve)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

DROPIN = os.path.join(ROOT, "dsp_audioreclabs_b200", "dropin")
STUBS = os.path.join(ROOT, "tests", "stubs")


def make_reference_shaped_tree(base):
    """A tree with the reference's layout and import statements (run.py:52-54, experiments/run_experiments.py:14-24),
    written here -- no reference code."""
    (base / "src").mkdir()
    (base / "experiments").mkdir()
    (base / "src" / "__init__.py").write_text("__version__ = 'ref'\n")
    (base / "src" / "audio_processing.py").write_text("WHO = 'reference'\n")
    (base / "src" / "feature_extraction.py").write_text("WHO = 'reference'\n")
    (base / "src" / "models.py").write_text("WHO = 'reference'\ndef create_classifier(t, **kw):\n    return ('reference', t)\n")
    (base / "src" / "visualization.py").write_text("import matplotlib.pyplot as plt\nimport seaborn as sns\nWHO = 'reference'\n")
    (base / "config.py").write_text("WHO = 'reference'\n")
    (base / "experiments" / "run_experiments.py").write_text(
        "import os, sys\n"
        "sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))\n"
        "import config\n"
        "import src.audio_processing, src.feature_extraction, src.models, src.visualization\n"
        "WHERE = {m: sys.modules[m].__file__ for m in ('config', 'src.audio_processing', 'src.feature_extraction', 'src.models', 'src.visualization')}\n")
    (base / "run.py").write_text(
        "import json, os, sys\n"
        "sys.path.append(os.path.dirname(os.path.abspath(__file__)))\n"
        "import config\n"
        "from experiments.run_experiments import WHERE\n"
        "import src.models\n"
        "WHERE['delegated'] = repr(src.models.create_classifier('svm'))\n"
        "WHERE['argv'] = sys.argv[1:]\n"
        "print('WHERE=' + json.dumps(WHERE))\n")


def run(cmd, cwd, env):
    out = subprocess.run(cmd, cwd=cwd, env=env, check=True, capture_output=True, text=True).stdout
    return json.loads([ln for ln in out.splitlines() if ln.startswith("WHERE=")][-1][6:])


def test_launcher_puts_the_dropin_ahead_of_the_script_directory(tmp_path):
    make_reference_shaped_tree(tmp_path)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, STUBS]), DSP_RESULTS_DIR=str(tmp_path / "results"))
    where = run([sys.executable, "-m", "dsp_audioreclabs_b200.run", "run.py", "--experiment", "feature"], str(tmp_path), env)
    for m in ("config", "src.audio_processing", "src.feature_extraction", "src.models"):
        assert where[m].startswith(DROPIN), (m, where[m])
    assert where["src.visualization"] == str(tmp_path / "src" / "visualization.py")     # outside the hot path: the reference's file
    assert where["delegated"] == "('reference', 'svm')"                                   # non-KNN classifiers: the reference's models.py
    assert where["argv"] == ["--experiment", "feature"]


def test_bare_pythonpath_is_shadowed_by_the_script_directory(tmp_path):
    """The failure mode the launcher exists for: documented so nobody goes back to it."""
    make_reference_shaped_tree(tmp_path)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([DROPIN, ROOT, STUBS]))
    where = run([sys.executable, "run.py"], str(tmp_path), env)
    assert where["config"] == str(tmp_path / "config.py")
    assert where["src.audio_processing"] == str(tmp_path / "src" / "audio_processing.py")


def test_launcher_usage_errors():
    r = subprocess.run([sys.executable, "-m", "dsp_audioreclabs_b200.run"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 2 and "run.py" in r.stdout
    r = subprocess.run([sys.executable, "-m", "dsp_audioreclabs_b200.run", "/nonexistent/x.py"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 2 and "no such script" in r.stderr


def test_bench_arms_are_quoted_on_the_same_config():
    """bench.py: the CUDA arm and the --impl reference arm print the same `config` dict for a given N (the reference arm
    says which bounded sample of it a step times in cpu_baseline.sample)."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    args = types.SimpleNamespace(utts=100000, train_utts=100000, knn_path="sharded")
    c1, c8 = bench.bench_config(args, 1), bench.bench_config(args, 8)
    assert c1["workload"] == bench.WORKLOAD and c1["samples_per_utterance"] == 44100 and c1["frame_length"] == 256
    assert "NCCL" in c8["parallelism"] and "NCCL" not in c1["parallelism"]
    assert bench.METRIC.endswith("(features+endpoints+KNN)")


def test_numa_binding_is_best_effort():
    """dist.bind_to_gpu_numa never raises: without a GPU / sysfs entry it reports what it could not do."""
    from dsp_audioreclabs_b200 import dist as ddist
    before = os.sched_getaffinity(0)
    info = ddist.bind_to_gpu_numa(0)
    assert set(info) >= {"gpu", "numa_node", "cpus", "bound"} and info["bound"] in (True, False)
    os.sched_setaffinity(0, before)
