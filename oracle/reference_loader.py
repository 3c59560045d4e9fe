"""Load the real reference modules by file path (build container only).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  /root/reference does not
exist on the GPU box, so callers must check ``available()`` and skip.
"""
import importlib.util
import os
import sys

from . import shims

REFERENCE_ROOT = os.environ.get("CMR_REFERENCE_ROOT", "/root/reference")


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
    """The module at /root/reference/environment/environment.py."""
    if "env" not in _cache:
        _cache["env"] = _load("_cmr_reference_environment", "environment/environment.py")
    return _cache["env"]


def pointnet_util():
    """The module at /root/reference/models/pointnet_util.py."""
    if "pn" not in _cache:
        _cache["pn"] = _load("_cmr_reference_pointnet_util", "models/pointnet_util.py")
    return _cache["pn"]


def kitti_dataset():
    """The module at /root/reference/dataset/KittiDataset.py (FarthestSampler lives there).  It imports plotting
    and logging packages this image lacks; they are never called by the code under test and are stubbed."""
    if "kitti" not in _cache:
        import types
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image", "tensorboardX", "torchvision",
                     "torchvision.transforms"):
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
        if not hasattr(sys.modules["matplotlib"], "pyplot"):
            sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
            sys.modules["matplotlib"].image = sys.modules["matplotlib.image"]
        if not hasattr(sys.modules["torchvision"], "transforms"):
            sys.modules["torchvision"].transforms = sys.modules["torchvision.transforms"]
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        _cache["kitti"] = _load("_cmr_reference_kitti_dataset", "dataset/KittiDataset.py")
    return _cache["kitti"]
