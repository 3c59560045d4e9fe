"""cmr_agent_b200 - B200-native (sm_100a) geometric hot path of CMR-Agent.

Drop-in modules (same function signatures as the reference):
    cmr_agent_b200.environment     <- environment/environment.py
    cmr_agent_b200.pointnet_util   <- models/pointnet_util.py
``install()`` aliases them into ``sys.modules`` under the names the reference imports.
See DESIGN.md and INTEGRATION.md.
"""
__version__ = "0.1.0"

from .install import install, uninstall  # noqa: F401
