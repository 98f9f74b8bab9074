"""adacharge_b200 — B200-native drop-in for adacharge's per-step MPC path.

Import surface mirrors reference adacharge/__init__.py:1-3 (star re-exports).
"""
from .interface import *  # noqa: F401,F403
from .adaptive_charging_optimization import *  # noqa: F401,F403
from .postprocessing import *  # noqa: F401,F403
from .adacharge import *  # noqa: F401,F403
