"""manipulation_planning_suite_b200 — B200-native kinodynamic-RRT extend step.

The hot path of JoshuaHaustein/manipulation_planning_suite (K push rollouts with fused validity /
goal / slice-ball tests, and the weighted SE(2)^n nearest-neighbour search over the tree) as
hand-written sm_100a CUDA behind a C ABI (include/mps_b200.h).  See DESIGN.md.
"""
from . import abi  # noqa: F401
from .abi import load_library  # noqa: F401

__all__ = ["abi", "load_library"]
