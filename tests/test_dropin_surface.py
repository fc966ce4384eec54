"""CPU: the drop-in `src` / `config` modules expose the reference's Python call surface
(names, signatures, config constants) captured from the reference in tests/golden/surface.json."""
import importlib
import inspect
import json
import os
import sys

import pytest

from conftest import GOLDEN, ROOT

DROPIN = os.path.join(ROOT, "dsp_audioreclabs_b200", "dropin")


@pytest.fixture(scope="module")
def dropin(tmp_path_factory):
    os.environ["DSP_RESULTS_DIR"] = str(tmp_path_factory.mktemp("results"))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "config" or k == "src" or k.startswith("src.")}
    sys.path.insert(0, DROPIN)
    mods = {n: importlib.import_module(n) for n in ("config", "src.audio_processing", "src.feature_extraction", "src.models")}
    yield mods
    sys.path.remove(DROPIN)
    for k in list(sys.modules):
        if k == "config" or k == "src" or k.startswith("src."):
            del sys.modules[k]
    sys.modules.update(saved)


def test_signatures_match_the_reference(dropin):
    surf = json.load(open(os.path.join(GOLDEN, "surface.json")))
    for qual, sig in surf["signatures"].items():
        mod, name = qual.rsplit(".", 1)
        fn = getattr(dropin[mod], name)
        assert str(inspect.signature(fn)) == sig, qual


def test_config_constants_match_the_reference(dropin):
    surf = json.load(open(os.path.join(GOLDEN, "surface.json")))
    cfg = dropin["config"]
    for name in surf["config_names"]:
        assert hasattr(cfg, name), name
    for name, val in surf["config"].items():
        got = getattr(cfg, name)
        assert (list(got) if isinstance(val, list) else got) == val, name
    assert cfg.FRAME_LENGTH == 1102 and cfg.FRAME_SHIFT == 441


def test_errors_raised_before_any_gpu_work(dropin):
    ap, fe, mo = dropin["src.audio_processing"], dropin["src.feature_extraction"], dropin["src.models"]
    import numpy as np
    with pytest.raises(ValueError):
        ap.create_window("blackman", 16)
    with pytest.raises(ValueError):
        ap.frame_signal(np.zeros(10), 4, 2, "blackman")
    with pytest.raises(ValueError):
        fe.extract_frame_features(np.zeros((0, 256)))
    with pytest.raises(ValueError):
        fe.extract_features_from_frames(np.zeros((3, 8)), method="mfcc")
    with pytest.raises(ValueError):
        mo.create_classifier("random_forest")
    seq = np.arange(12.0).reshape(6, 2)
    assert fe.pad_or_truncate_sequence(seq, 9).shape == (9, 2) and fe.pad_or_truncate_sequence(seq, 4).shape == (4, 2)
