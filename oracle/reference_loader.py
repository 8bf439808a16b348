"""TEST INFRASTRUCTURE ONLY.  Import the unmodified reference modules from /root/reference.

The reference's ``utils/utils.py`` imports five packages that are not installed here
(tensorboardX, plyfile, skimage, plotly, trimesh; utils/utils.py:4-16).  None of them is
touched by the hot path, so empty stub modules are enough.

``/root/reference`` exists only in the build container; ``available()`` is False on the
GPU box and every caller must then fall back to the committed golden fixtures.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DIGS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "DiGS.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """Returns (models.DiGS module, models.losses module, utils.utils module)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _stub("tensorboardX", SummaryWriter=object)
    _stub("plyfile", PlyData=object)
    _stub("trimesh")
    sk = _stub("skimage")
    sk.measure = _stub("skimage.measure")
    pl = _stub("plotly")
    pl.graph_objs = _stub("plotly.graph_objs")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models.DiGS as ref_digs          # noqa: E402
    import models.losses as ref_losses      # noqa: E402
    import utils.utils as ref_utils         # noqa: E402
    _loaded = (ref_digs, ref_losses, ref_utils)
    return _loaded
