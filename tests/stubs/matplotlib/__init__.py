"""TEST-ONLY stand-in for matplotlib (not installed in this image; plotting is out of scope).

The reference's drivers (src/experiments.py, src/visualizations_utils.py) import matplotlib at module scope and
draw figures after their simulation loops. This stub accepts every call and RECORDS it (`matplotlib.calls`:
list of (dotted name, args, kwargs)), so a test can run the unmodified drivers on this package and still look at
what they plotted (e.g. the measured viscosities of `plot_measured_viscosity_vs_omega`)."""
calls = []


class Anything:
    """Callable, indexable, iterable (as a pair: `fig, ax = plt.subplots()`), attribute access never fails."""

    def __init__(self, name='matplotlib'):
        object.__setattr__(self, '_name', name)

    def __getattr__(self, key):
        if key.startswith('__') and key.endswith('__'):
            raise AttributeError(key)
        return Anything(self._name + '.' + key)

    def __setattr__(self, key, value):
        pass

    def __call__(self, *args, **kwargs):
        calls.append((self._name, args, kwargs))
        return Anything(self._name + '()')

    def __getitem__(self, key):
        return Anything(f'{self._name}[{key!r}]')

    def __setitem__(self, key, value):
        pass

    def __iter__(self):
        return iter((Anything(self._name + '[0]'), Anything(self._name + '[1]')))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def update(self, *args, **kwargs):
        pass


rcParams = Anything('matplotlib.rcParams')


def use(*args, **kwargs):
    calls.append(('matplotlib.use', args, kwargs))


def __getattr__(name):
    if name.startswith('__'):
        raise AttributeError(name)
    if name in ('cm', 'pyplot'):          # `from matplotlib import cm` must find the submodule, not a recorder
        import importlib
        return importlib.import_module('.' + name, __name__)
    return Anything('matplotlib.' + name)
