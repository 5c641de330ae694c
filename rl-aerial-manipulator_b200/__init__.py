"""B200-native batched WaypointQuadEnv (LahiruCooray/rl-aerial-manipulator hot path).

Import as `rl_aerial_manipulator_b200` (alias module at the repo root).
"""
from .quad_constants import QUAD  # noqa: F401
from ._cabi import make_config, load_library, QuadsimError  # noqa: F401

__all__ = ["QUAD", "make_config", "load_library", "QuadsimError"]
