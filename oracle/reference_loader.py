"""TEST INFRASTRUCTURE ONLY.  Import the unmodified reference modules from /root/reference.

The reference's ``utils/utils.py`` imports five packages that are not installed here
(tensorboardX, plyfile, skimage, plotly, trimesh; utils/utils.py:4-16).  None of them is
touched by the hot path, so empty stub modules are enough.

``/root/reference`` exists only in the build container.  ``__graft_entry__.build()`` installs the three
module directories the hot path touches (models/, utils/, and surface_reconstruction/ whose training script serves as an
integration harness -- all unmodified) into the git-ignored
``baseline/_ref/`` so that the *reference arm* of bench.py can time the real reference modules on the GPU
box's host cores (and, for context, through stock PyTorch CUDA).  The -m gpu tests and smoke() never import
it: they use the committed golden fixtures.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
INSTALLED_ROOT = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")


def _pick_root():
    env = os.environ.get("DIGS_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", INSTALLED_ROOT):
        if os.path.isfile(os.path.join(cand, "models", "DiGS.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def install(dst=INSTALLED_ROOT, src="/root/reference"):
    """Copy the unmodified reference modules of the hot path (models/, utils/) into baseline/_ref (git-ignored)."""
    import shutil
    if not os.path.isfile(os.path.join(src, "models", "DiGS.py")):
        return False
    for sub in ("models", "utils", "surface_reconstruction"):      # the last one: the training script used as a harness
        shutil.copytree(os.path.join(src, sub), os.path.join(dst, sub), dirs_exist_ok=True,
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "scripts", "camera_params", "log"))
    for f in ("LICENCE", "README.md"):
        if os.path.isfile(os.path.join(src, f)):
            shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    return True


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
