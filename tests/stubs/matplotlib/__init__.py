"""Test-only stand-in for matplotlib (absent from this image and from /opt/wheelhouse, SURVEY.md
section 0 item 6): lets the reference's driver scripts import and call their plotting code, which
is outside the hot path, without drawing anything.  Only on sys.path of the subprocesses that
tests/test_scripts_gpu.py and oracle/gen_golden_scripts.py start."""


class Stub:
    """Absorbs any attribute access, call, indexing or assignment."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return Stub()

    def __call__(self, *a, **k):
        return Stub()

    def __getitem__(self, key):
        return Stub()

    def __setitem__(self, key, value):
        pass

    def __iter__(self):
        return iter(())

    def __len__(self):
        return 0

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class AxesArray(list):
    def flatten(self):
        return AxesArray(self)

    ravel = flatten

    @property
    def flat(self):
        return iter(self)


rcParams = {}
__version__ = "0.0-stub"


def use(*a, **k):
    pass


def __getattr__(name):
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    return Stub()
