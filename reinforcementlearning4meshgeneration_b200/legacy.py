"""Legacy (v1) flavour of the drop-in module: mount it as ``rl.boundary_env`` and the reference's v1 trainers /
testbed run unchanged on the CUDA env (rl/baselines/RL_Mesh.py:6,24-25, rl/baselines/testbed.py:2,68):

    import sys, reinforcementlearning4meshgeneration_b200.legacy as m
    sys.modules["rl.boundary_env"] = m          # before `from rl.boundary_env import BoudaryEnv, read_polygon, boundary`

``BoudaryEnv`` here speaks the legacy gym API of rl/boundary_env.py:67-84 and :263: ``reset() -> obs`` and
``step(action) -> (obs, reward, done, info)`` with ``info = {"is_complete": bool}``; everything else
(``generated_meshes``, ``save_meshes``, ``boundary``, ``original_vertices``, ``updated_boundary``, ``write_2_file``)
is shared with the Gymnasium flavour in ``boundary_env.py``.
"""
from __future__ import annotations

from .boundary_env import Boundary2D, Vertex, boundary, read_polygon  # noqa: F401
from .boundary_env import BoudaryEnv as _GymnasiumBoudaryEnv


class BoudaryEnv(_GymnasiumBoudaryEnv):
    API = "gym"


__all__ = ["BoudaryEnv", "read_polygon", "boundary", "Boundary2D", "Vertex"]
