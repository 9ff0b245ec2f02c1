"""CPU: the C restatement of BoudaryEnv.move() (E:459-594) against the golden traces recorded from the
live reference by oracle/sweep_move_vs_reference.py --record, and against the live reference itself when its tree is
present (the build container).  The smooth_pave branch (every candidate excluded) is outside the restated path: the
traces mark the step where the reference reaches it and restart the episode."""
import os

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import ref_loader
from oracle.c_oracle import OracleEnv

NAMES = ["boundary0", "dolphine3", "easy1_1"]


def load_move(name):
    z = np.load(os.path.join(GOLDEN, f"move_{name}.npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", NAMES)
def test_oracle_move_reproduces_golden_trace(name):
    tr = load_move(name)
    o = OracleEnv(tr["xy0"], original_area=float(tr["original_area"]))
    T = len(tr["type"])
    n_ok = 0
    for i in range(T):
        obs, rew, done, info, smooth = o.move([float(tr["polar"][i, 0]), float(tr["polar"][i, 1])], float(tr["type"][i]))
        assert rew == 0 and smooth == bool(tr["smooth"][i]) and done == bool(tr["done"][i]), f"{name}: flags differ at move {i}"
        assert info["is_complete"] == bool(tr["complete"][i])
        assert (obs is None) == bool(tr["obs_none"][i])
        if obs is not None:
            assert np.array_equal(obs, tr["obs"][i]), f"{name}: observation differs at move {i}"
            assert obs[1] == 0.0            # static point environment: area-ratio slot is 0 (C:1209-1214)
        assert o.n_elements == int(tr["n_elements"][i]) and o.n == int(tr["n_boundary"][i]) and o.ref_index == int(tr["ref_index"][i])
        n_ok += 1
        if tr["reset_after"][i]:
            o.reset()
    assert n_ok == T and int(tr["n_elements"].max()) > 20


@pytest.mark.skipif(not ref_loader.reference_available(), reason="the reference tree is only present in the build container")
def test_oracle_move_against_the_live_reference():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import sweep_move_vs_reference as sw
    for k, name in enumerate(["tool", "half_wheel", "boundary16"]):
        xy = ref_loader.load_domain_xy(name)
        mism, episodes, smooths, elems = sw.run_domain(name, xy, 150, 900 + k)
        assert mism == 0


SMOOTH_NAMES = ["boundary0", "tool", "bird"]


def load_smooth(name):
    z = np.load(os.path.join(GOLDEN, f"smooth_{name}.npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", SMOOTH_NAMES)
def test_oracle_smooth_pave_reproduces_golden_trace(name):
    """move() WITH the reference's smooth_pave (M:816-1140, M:1284-1316; recorded by oracle/sweep_smooth_vs_reference.py
    --record): the oracle calls the same libm, so every vertex coordinate -- front and interior -- is compared exactly."""
    tr = load_smooth(name)
    o = OracleEnv(tr["xy0"], original_area=float(tr["original_area"]))
    o.set_smoothing(True)
    T = len(tr["type"])
    for i in range(T):
        obs, rew, done, info, smooth = o.move([float(tr["polar"][i, 0]), float(tr["polar"][i, 1])], float(tr["type"][i]))
        assert smooth == bool(tr["smooth"][i]) and done == bool(tr["done"][i]) and info["is_complete"] == bool(tr["complete"][i]), (name, i)
        assert (obs is None) == bool(tr["obs_none"][i])
        if obs is not None:
            assert np.array_equal(obs, tr["obs"][i]), f"{name}: observation differs at move {i}"
        assert o.n_elements == int(tr["n_elements"][i]) and o.n == int(tr["n_boundary"][i]) and o.ref_index == int(tr["ref_index"][i])
        nv = int(tr["n_vertices"][i])
        assert np.array_equal(o.vertex_xy(), tr["vertex_xy"][i, :nv]), f"{name}: vertex coordinates differ at move {i}"
        ids, _ = o.boundary()
        assert ids.tolist() == tr["boundary_ids"][i, :o.n].tolist()
        if tr["reset_after"][i]:
            o.reset()
    assert int(tr["smooth"].sum()) >= 1 and o.n_smoothings == int(tr["smooth"].sum())


@pytest.mark.skipif(not ref_loader.reference_available(), reason="the reference tree is only present in the build container")
def test_oracle_smooth_pave_against_the_live_reference():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import sweep_smooth_vs_reference as sw
    total = 0
    for k, name in enumerate(["star", "half_wheel", "fat"]):
        xy = ref_loader.load_domain_xy(name)
        mism, episodes, smooths = sw.run_domain(name, xy, 120, 700 + k)
        assert mism == 0
        total += smooths
    assert total > 10
