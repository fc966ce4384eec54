"""Run one of the reference's scripts, unchanged, on the CUDA drop-in modules:

    python -m dsp_audioreclabs_b200.run /path/to/DSP-AudioRecLabs/run.py --experiment feature --window-type hamming
    python -m dsp_audioreclabs_b200.run /path/to/DSP-AudioRecLabs/ablation_study.py --experiment frame_length --classifier knn

Why a launcher and not PYTHONPATH: Python puts the script's own directory at sys.path[0], ahead
of PYTHONPATH, so `import config` / `from src.audio_processing import ...` (run.py:53,
experiments/run_experiments.py:17-20, train_model.py:15-18) would resolve to the reference's
files.  Here sys.path becomes [dropin, script directory, ...]: `config`, `src.audio_processing`,
`src.feature_extraction` and `src.models` come from dsp_audioreclabs_b200/dropin (CUDA), while
`experiments.*`, `train_model`, `src.visualization` and the script itself are the reference's own,
untouched files.  DSP_RUN_REPORT=<file> appends one JSON line with the CUDA launch count of the
process when the script ends (what the end-to-end tests assert on).
"""
import atexit
import json
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DROPIN = os.path.join(HERE, "dropin")


def install(script_dir=None):
    """Put the drop-in directory first on sys.path (and the script directory right after it)."""
    for p in (DROPIN, script_dir):
        while p and p in sys.path:
            sys.path.remove(p)
    if script_dir:
        sys.path.insert(0, script_dir)
    sys.path.insert(0, DROPIN)
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.append(root)          # the dsp_audioreclabs_b200 package itself
    for name in [m for m in sys.modules if m == "config" or m == "src" or m.startswith("src.")]:
        del sys.modules[name]


def _report(path, script):
    try:
        from dsp_audioreclabs_b200 import batch
        launches = sum(c.launch_count for c in batch._default.values())
        import src.audio_processing as ap
        line = {"script": script, "gpu_launches": int(launches), "frontend_launches": list(ap.launch_log),
                "src.audio_processing": getattr(ap, "__file__", None),
                "config": getattr(sys.modules.get("config"), "__file__", None)}
    except Exception as exc:                       # the report must never change the script's exit status
        line = {"script": script, "error": repr(exc)}
    with open(path, "a") as f:
        f.write(json.dumps(line) + "\n")


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        print(f"dsp_audioreclabs_b200.run: no such script: {script}", file=sys.stderr)
        return 2
    install(os.path.dirname(script))
    report = os.environ.get("DSP_RUN_REPORT")
    if report:
        atexit.register(_report, report, script)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
