"""No-op matplotlib.pyplot (TEST INFRASTRUCTURE ONLY) — see matplotlib/__init__.py."""
from . import _Noop

rcParams = {}
_noop = _Noop()


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _noop
