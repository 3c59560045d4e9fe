"""Make the reference's own import statements resolve to the B200 drop-ins.

The reference imports the hot path as
    from environment import environment as env        (Train_Agent.py:15, Test_Agent.py:18)
    from environment.buffer import Buffer             (Train_Agent.py:16)
    from .pointnet_util import index_points, square_distance      (models/PointNN.py:7)

``install()`` may be called before or after those imports:

* ``environment``: the reference's real package (its ``__init__`` is empty) is imported if it can be found,
  so ``environment.buffer`` keeps resolving; only the submodule ``environment.environment`` is replaced, and
  every already-imported module that holds the original under some name (``env``) is re-pointed.
* ``models.pointnet_util``: the real module is kept (its ``nn.Module`` classes, ``pc_normalize``, ``timeit``
  stay where callers expect them) and the hot-path FUNCTIONS are patched onto it - the classes then call the
  drop-ins through their module globals.  If ``models`` has not been imported yet the real file is loaded
  under its final name first, so ``models/PointNN.py`` binds the drop-ins when it is imported later; if it
  has, the names PointNN (or anybody else in ``models.*``) already bound are re-pointed.
* With no reference tree on ``sys.path`` the drop-in modules are registered under those names as they are.

``uninstall()`` restores what ``install()`` changed.
"""
import importlib
import importlib.util
import os
import sys
import types

_PN_FUNCS = ("square_distance", "index_points", "farthest_point_sample", "query_ball_point", "sample_and_group",
             "sample_and_group_all")
_PN_EXTRAS = ("knn_point", "group_points", "farthest_point_sample_from")

_undo = []   # closures, run in reverse by uninstall()


def _ours(mod):
    return getattr(mod, "__name__", "").startswith("cmr_agent_b200")


def _find_package_dir(name):
    """Directory of the top-level package ``name`` without importing it; None if absent (or ours)."""
    try:
        spec = importlib.util.find_spec(name)
    except (ImportError, ValueError):
        return None
    if spec is None or not spec.submodule_search_locations:
        return None
    for loc in spec.submodule_search_locations:
        if os.path.isdir(loc) and "cmr_agent_b200" not in os.path.abspath(loc):
            return loc
    return None


def _repoint(old, new, prefixes=None):
    """Every loaded module that holds ``old`` as a global now holds ``new``."""
    if old is None or old is new:
        return
    for mname, mod in list(sys.modules.items()):
        if mod is None or _ours(mod) or (prefixes and not mname.startswith(prefixes)):
            continue
        d = getattr(mod, "__dict__", None)
        if not d:
            continue
        for key, val in list(d.items()):
            if val is old:
                d[key] = new
                _undo.append(lambda d=d, key=key, old=old, new=new: d.__setitem__(key, old) if d.get(key) is new else None)


def _install_environment(_env):
    pkg = sys.modules.get("environment")
    if pkg is None or not hasattr(pkg, "__path__"):
        if _find_package_dir("environment") is not None:
            pkg = importlib.import_module("environment")        # the reference's package: empty __init__
        else:
            pkg = types.ModuleType("environment")
            pkg.__path__ = []                                   # no reference tree: a bare namespace
            sys.modules["environment"] = pkg
            _undo.append(lambda: sys.modules.pop("environment", None) if sys.modules.get("environment") is pkg else None)
    original = sys.modules.get("environment.environment")
    if original is _env:
        return
    had_attr = "environment" in pkg.__dict__
    old_attr = pkg.__dict__.get("environment")
    sys.modules["environment.environment"] = _env
    pkg.environment = _env

    def restore():
        if original is not None:
            sys.modules["environment.environment"] = original
        elif sys.modules.get("environment.environment") is _env:
            del sys.modules["environment.environment"]
        if had_attr:
            pkg.environment = old_attr
        elif pkg.__dict__.get("environment") is _env:
            del pkg.__dict__["environment"]
    _undo.append(restore)
    if original is not None and not _ours(original):
        _repoint(original, _env)


def _install_pointnet_util(_pn):
    real = sys.modules.get("models.pointnet_util")
    if real is _pn:
        return
    if real is None:
        pkg_dir = _find_package_dir("models")
        path = os.path.join(pkg_dir, "pointnet_util.py") if pkg_dir else None
        if path and os.path.isfile(path):
            # load the real file under its final name WITHOUT importing the `models` package (whose __init__
            # imports PointNN, which would bind the original functions before we can patch them)
            spec = importlib.util.spec_from_file_location("models.pointnet_util", path)
            real = importlib.util.module_from_spec(spec)
            sys.modules["models.pointnet_util"] = real
            try:
                spec.loader.exec_module(real)
            except Exception:
                del sys.modules["models.pointnet_util"]
                raise
            _undo.append(lambda real=real: sys.modules.pop("models.pointnet_util", None)
                         if sys.modules.get("models.pointnet_util") is real else None)
    if real is None:
        # no reference tree: the drop-in itself answers to the name
        sys.modules["models.pointnet_util"] = _pn
        models_pkg = sys.modules.get("models")
        if models_pkg is not None:
            models_pkg.pointnet_util = _pn
        _undo.append(lambda: sys.modules.pop("models.pointnet_util", None)
                     if sys.modules.get("models.pointnet_util") is _pn else None)
        return
    for name in _PN_FUNCS + _PN_EXTRAS:
        new = getattr(_pn, name)
        old = real.__dict__.get(name)
        if old is new:
            continue
        real.__dict__[name] = new
        if old is None:
            _undo.append(lambda name=name, new=new: real.__dict__.pop(name, None) if real.__dict__.get(name) is new else None)
        else:
            _undo.append(lambda name=name, old=old: real.__dict__.__setitem__(name, old))
            _repoint(old, new, prefixes=("models",))     # `from .pointnet_util import index_points` already executed
    real.__cmr_b200_patched__ = True
    _undo.append(lambda: real.__dict__.pop("__cmr_b200_patched__", None))


def install(environment=True, pointnet_util=True):
    from . import environment as _env
    from . import pointnet_util as _pn

    if environment:
        _install_environment(_env)
    if pointnet_util:
        _install_pointnet_util(_pn)
    return _env, _pn


def uninstall():
    while _undo:
        _undo.pop()()
