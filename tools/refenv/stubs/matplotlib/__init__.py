"""Minimal stand-in so that the reference's `import matplotlib.pyplot as plt` succeeds (test tooling only)."""
class _Dummy:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Dummy()
    def __getattr__(self, name): return _Dummy()
def __getattr__(name):
    return _Dummy
