"""Make the reference's own import statements resolve to the B200 drop-ins.

The reference imports the hot path as
    from environment import environment as env        (Train_Agent.py:15, Test_Agent.py:18)
    from .pointnet_util import index_points, square_distance      (models/PointNN.py:7)
``install()`` registers the drop-in modules under those names in ``sys.modules`` BEFORE the
reference's drivers/models are imported, so they run unchanged.
"""
import sys
import types


def install(environment=True, pointnet_util=True):
    from . import environment as _env
    from . import pointnet_util as _pn

    if environment:
        pkg = sys.modules.get("environment")
        if pkg is None or not hasattr(pkg, "__path__"):
            pkg = types.ModuleType("environment")
            pkg.__path__ = []          # a namespace the import system treats as a package
            sys.modules["environment"] = pkg
        pkg.environment = _env
        sys.modules["environment.environment"] = _env
    if pointnet_util:
        # `from .pointnet_util import ...` inside the `models` package looks up this key first
        sys.modules["models.pointnet_util"] = _pn
        models_pkg = sys.modules.get("models")
        if models_pkg is not None:
            models_pkg.pointnet_util = _pn
    return _env, _pn


def uninstall():
    for name in ("environment.environment", "models.pointnet_util"):
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__name__", "").startswith("cmr_agent_b200"):
            del sys.modules[name]
    pkg = sys.modules.get("environment")
    if pkg is not None and getattr(getattr(pkg, "environment", None), "__name__", "").startswith("cmr_agent_b200"):
        del sys.modules["environment"]
