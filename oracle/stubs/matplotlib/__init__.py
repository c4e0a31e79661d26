"""Import stub: the reference hard-imports matplotlib (ultralytics/utils/__init__.py:24) which this image lacks.
Only used by oracle/ref_shim.py when the reference tree itself is imported; plotting is never called."""
__version__ = "0.0-stub"
rcParams = {}


def use(*a, **k):
    pass


def rc(*a, **k):
    pass


def get_backend():
    return "agg"
