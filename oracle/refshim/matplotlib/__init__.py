"""Minimal stand-in for matplotlib so the UNMODIFIED reference modules can be imported
in a container that has no matplotlib. TEST INFRASTRUCTURE ONLY (used by
oracle/gen_golden.py to produce tests/golden/*.npz). Every plotting call is swallowed."""


def use(*_a, **_k):
    return None


class _Sink:
    def __getattr__(self, _name):
        return _Sink()

    def __call__(self, *_a, **_k):
        return _Sink()

    def __iter__(self):
        return iter(())

    def __getitem__(self, _i):
        return _Sink()


rcParams = {}
