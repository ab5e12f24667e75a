"""TEST-ONLY stand-in for matplotlib.cm."""
from . import Anything


def __getattr__(name):
    if name.startswith('__'):
        raise AttributeError(name)
    return Anything('cm.' + name)
