"""Drop-in `src` package: the reference's import paths (src.audio_processing,
src.feature_extraction, src.models) backed by libdspfront.so.

With this directory's parent (dsp_audioreclabs_b200/dropin) AHEAD of the reference checkout on
sys.path -- `python -m dsp_audioreclabs_b200.run <script.py> ...` arranges exactly that --
run.py / ablation_study.py / train_model.py / compare_feature_methods.py run unchanged.

The package path is extended with the reference's own `src/` directory (found on sys.path, or
under DSP_REFERENCE_ROOT): the three hot-path modules resolve here first, everything that is
outside the hot path (`src.visualization`) resolves to the reference's
file, so `from src.visualization import ...` (experiments/run_experiments.py:20-24) keeps working.
"""
import os
import sys

__version__ = '1.0.0'

_HERE = os.path.dirname(os.path.abspath(__file__))


def _reference_src_dirs():
    roots = []
    env = os.environ.get('DSP_REFERENCE_ROOT')
    if env:
        roots.append(env)
    roots.extend(p or os.getcwd() for p in sys.path)
    out = []
    for r in roots:
        d = os.path.join(os.path.abspath(r), 'src')
        if d != _HERE and d not in out and os.path.isfile(os.path.join(d, '__init__.py')) \
                and os.path.isfile(os.path.join(d, 'audio_processing.py')):
            out.append(d)
    return out


REFERENCE_SRC_DIRS = _reference_src_dirs()
__path__ = [_HERE] + REFERENCE_SRC_DIRS
