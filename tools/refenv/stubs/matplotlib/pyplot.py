from . import _Dummy
def __getattr__(name):
    return _Dummy
