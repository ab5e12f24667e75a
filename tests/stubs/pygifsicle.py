"""TEST-ONLY stand-in for pygifsicle (src/visualizations_utils.py:9 imports `optimize` at module scope)."""


def optimize(*args, **kwargs):
    pass
