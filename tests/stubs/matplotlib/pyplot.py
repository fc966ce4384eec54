"""Stub pyplot: figures and axes that accept every call; savefig writes an empty placeholder so
callers that list their outputs still find the file."""
from . import AxesArray, Stub, rcParams  # noqa: F401


def subplots(nrows=1, ncols=1, *a, **k):
    n = int(nrows) * int(ncols)
    return Stub(), (Stub() if n == 1 else AxesArray(Stub() for _ in range(n)))


def savefig(path, *a, **k):
    try:
        with open(path, "ab"):
            pass
    except (OSError, TypeError):
        pass


def __getattr__(name):
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    return Stub()
