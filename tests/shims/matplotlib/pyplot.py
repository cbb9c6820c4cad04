"""TEST SHIM: matplotlib.pyplot (see matplotlib/__init__.py)."""
import sys

from . import Null


class _Image(Null):
    def __init__(self, data, extent=None):
        self._shape, self._extent = getattr(data, "shape", (1, 1)), extent

    def get_extent(self):
        if self._extent is not None:
            return tuple(self._extent)
        h, w = self._shape[:2]
        return (-0.5, w - 0.5, -0.5, h - 0.5)          # matplotlib's extent of an origin='lower' image


def imshow(data, *a, extent=None, **k):
    return _Image(data, extent)


def subplots(*a, **k):
    return Null(), Null()


def figure(*a, **k):
    return Null()


def __getattr__(name):                                  # every other pyplot function: absorbed
    if name.startswith("__"):
        raise AttributeError(name)
    return Null()
