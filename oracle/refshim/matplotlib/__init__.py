"""No-op stand-in for matplotlib (TEST INFRASTRUCTURE ONLY).

The reference imports matplotlib at module top (spectrogram_analyse.py:3,
modulator.py:2, frequency_correction.py:3) and inline (ft8_decode.py:343) purely
for plotting side effects.  This container has no matplotlib, so the golden-vector
generator puts this directory on sys.path to import the UNMODIFIED reference.
Nothing in the product package imports this.
"""


class _Noop:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __getitem__(self, key):
        return self

    def __setitem__(self, key, value):
        pass

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


rcParams = {}
_noop = _Noop()


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _noop
