from . import Stub

FontProperties = Stub
fontManager = Stub()


def __getattr__(name):
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    return Stub()
