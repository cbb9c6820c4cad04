"""TEST SHIM (not product code): a do-nothing stand-in for matplotlib, which is absent from this image.  The reference's
scripts import it for rendering / plotting only (assembly.py:7-8,90,668-747; assembly_cfg.py:15,103-124; eval_assembly.py:5,
241-297).  Every call is absorbed; the one value the reference computes FROM matplotlib — `imshow(...).get_extent()` in
assembly_cfg.py:106-108 — is reproduced: for origin='lower' it is (-0.5, W-0.5, -0.5, H-0.5)."""


class Null:
    """Absorbs attribute access, calls, indexing, iteration and context management."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return Null()

    def __call__(self, *a, **k):
        return Null()

    def __getitem__(self, k):
        return Null()

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __bool__(self):
        return False


def use(*a, **k):
    return None


rcParams = {}
__version__ = "0.0-shim"
