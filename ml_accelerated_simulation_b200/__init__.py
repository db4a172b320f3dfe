"""B200-native rollout hot path of BigBalloon8/ml-accelerated-simulation:
finite-volume Navier-Stokes RK4 step + learned-correction CNN, hand-written sm_100a kernels behind
the C ABI of include/mlsim.h.  Module names mirror the ``torch_cfd`` / ``models`` names the reference
scripts import, so ``from ml_accelerated_simulation_b200 import grids, boundaries`` etc. are drop-in.
"""
from . import boundaries, equations, forcings, grids  # noqa: F401  (pure-Python, import cheaply)
from .equations import stable_time_step  # noqa: F401
from .plan import Plan, PlanConfig  # noqa: F401

__all__ = ["Plan", "PlanConfig", "grids", "boundaries", "equations", "forcings", "stable_time_step"]
