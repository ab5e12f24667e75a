"""TEST-ONLY stand-in for matplotlib.pyplot: every function is accepted and recorded in `matplotlib.calls`."""
from . import Anything


def __getattr__(name):
    if name.startswith('__'):
        raise AttributeError(name)
    return Anything('plt.' + name)
