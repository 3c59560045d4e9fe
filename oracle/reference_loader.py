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
