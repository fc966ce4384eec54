"""Test-only stand-in for seaborn (see tests/stubs/matplotlib)."""
from matplotlib import Stub


def __getattr__(name):
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    return Stub()
