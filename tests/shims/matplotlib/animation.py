"""TEST SHIM: matplotlib.animation (VideoWriter.py:1 imports FFMpegWriter; never used when args.video is False)."""
from . import Null


class FFMpegWriter(Null):
    pass
