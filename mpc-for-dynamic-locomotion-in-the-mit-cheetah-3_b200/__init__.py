"""B200-native batched convex-MPC solver for the Lite3 quadruped (hot path of the
reference's ``src/mpc.py``).  Import through :mod:`mpc_b200` (the directory name is
not a Python identifier) or ``importlib.import_module``."""
from .gait import GaitPlan, GAITS, LEGS, stance_bits          # noqa: F401
from .assembly import (desired_trajectory, assemble_tick,      # noqa: F401
                       reference_velocity, pack_problem)
from . import problems                                          # noqa: F401
from . import _capi                                             # noqa: F401
from ._capi import CmpcError                                    # noqa: F401
from .solver import BatchedMPC, MPC, SolveStats                 # noqa: F401

__version__ = "0.3.0"
from . import sharding                                           # noqa: F401,E402
from .rollout import ClosedLoopRollout                           # noqa: F401,E402
from . import logexport                                          # noqa: F401,E402
from .controllers import BatchedLegController                    # noqa: F401,E402
from . import kinematics                                         # noqa: F401,E402
