"""B200-native batched BoudaryEnv (quad-mesh generation RL environment).

Hot path = hand-written sm_100a CUDA kernels behind a C ABI (include/meshgen_b200.h,
lib/libmeshgen_b200.so); this package is the thin Python host mirror of the reference's
interface (BoudaryEnv / read_polygon / boundary + an SB3-style VecEnv)."""
from ._lib import MeshgenError, build, load  # noqa: F401
from .batched_env import ACTION_HIGH, ACTION_LOW, BatchedBoudaryEnv, StepResult, as_xy, poly_area  # noqa: F401

__all__ = ["BatchedBoudaryEnv", "StepResult", "MeshgenError", "ACTION_LOW", "ACTION_HIGH", "build", "load",
           "as_xy", "poly_area"]
