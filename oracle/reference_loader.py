"""Load the real reference modules (y2w-oc/CMR-Agent), from /root/reference in the build container or from
the copy staged by oracle/make_ref.py (oracle/_ref, git-ignored, shipped to the GPU box).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): tests, smoke() and bench.py's baseline legs.
Callers must check ``available()`` and skip when neither root exists.
"""
import importlib.util
import os
import sys
import types

from . import shims

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for cand in (os.environ.get("CMR_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "environment", "environment.py")):
            return cand
    return os.environ.get("CMR_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "environment", "environment.py"))


def _load(name, relpath):
    shims.install()
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def environment():
    """The reference's environment/environment.py under a private module name (never the drop-in)."""
    if "env" not in _cache:
        _cache["env"] = _load("_cmr_reference_environment", "environment/environment.py")
    return _cache["env"]


def pointnet_util():
    """The reference's models/pointnet_util.py under a private module name."""
    if "pn" not in _cache:
        _cache["pn"] = _load("_cmr_reference_pointnet_util", "models/pointnet_util.py")
    return _cache["pn"]


def stub_missing_third_party():
    """Empty stand-ins for plotting/logging packages the reference's drivers and datasets import but
    this image lacks; nothing under test ever calls them."""
    shims.install()
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image", "tensorboardX"):
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            m.__cmr_shim__ = True
            sys.modules[name] = m
    mpl = sys.modules["matplotlib"]
    if getattr(mpl, "__cmr_shim__", False):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
        mpl.image = sys.modules["matplotlib.image"]
    tbx = sys.modules["tensorboardX"]
    if getattr(tbx, "__cmr_shim__", False) and not hasattr(tbx, "SummaryWriter"):
        class SummaryWriter:   # Train_Agent.py:10 imports the name; the tests never construct it
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass
        tbx.SummaryWriter = SummaryWriter


def put_on_path():
    """Make ``import models`` / ``import environment`` / ``import config`` resolve to the reference tree,
    as they do when its drivers are started from the repository root (Train_Agent.py:13-16)."""
    stub_missing_third_party()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return REFERENCE_ROOT


def purge_packages():
    """Forget every reference package imported by name (so a test can re-import with/without install())."""
    for name in list(sys.modules):
        head = name.split(".")[0]
        if head in ("models", "environment", "config", "dataset", "utils"):
            mod = sys.modules[name]
            f = getattr(mod, "__file__", None) or ""
            if f.startswith(REFERENCE_ROOT) or getattr(mod, "__name__", "").startswith("cmr_agent_b200") \
                    or not f:
                del sys.modules[name]


def kitti_dataset():
    """The module at dataset/KittiDataset.py (FarthestSampler lives there).  It imports plotting
    and logging packages this image lacks; they are never called by the code under test and are stubbed."""
    if "kitti" not in _cache:
        put_on_path()
        _cache["kitti"] = _load("_cmr_reference_kitti_dataset", "dataset/KittiDataset.py")
    return _cache["kitti"]
