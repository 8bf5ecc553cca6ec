"""CPU tests of the search oracle (oracle/kino_ref.cpp) and of the input generators, against the committed golden vectors
(tests/golden/kino_golden.json, produced by tests/golden/make_golden.py).  The reference has no expected outputs for this
path: these goldens freeze the oracle; tests/test_kino_reference_build.py pins it to the reference's own
kino_astar.cpp compiled unmodified (oracle/_ref/libkino_ref.so)."""
import ctypes as C
import hashlib
import json
import math
import os

import numpy as np
import pytest

import oracle_lib
import uav_motion_planning_b200 as u
from uav_motion_planning_b200 import _lib

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kino_golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def setup_case(case, libm_mode=0):
    world = u.make_world(*case["dims"], seed=case["map_seed"], map_type=case["map_type"])
    p = _lib.KinoParams()
    (u.load().uavmp_kino_params_launch if case["launch_params"] else u.load().uavmp_kino_params_default)(C.byref(p))
    p.collision_check_type = case["collision_check_type"]
    if case["allocated_node_num"]:
        p.allocated_node_num = case["allocated_node_num"]
    return world, p, oracle_lib.KinoOracle(world, p, libm_mode=libm_mode)


@pytest.mark.parametrize("case", GOLD, ids=lambda c: c["name"])
def test_golden(case):
    world, p, orc = setup_case(case)
    assert sha(world.occ) == case["occ_sha256"] and sha(world.cloud) == case["cloud_sha256"]   # generator is frozen too
    for q in case["queries"]:
        r = orc.search(q["start_pt"], q["start_vel"], q["end_pt"], q["end_vel"])
        assert (r["status"], r["use_node_num"], r["n_pop"], str(r["pop_hash"]), r["n_path"]) == \
            (q["status"], q["use_node_num"], q["n_pop"], q["pop_hash"], q["n_path"])
        assert sha(r["path"]) == q["path_sha256"]
        assert r["counters"] == q["counters"]


def test_glibc_mode_gives_the_same_searches():
    """libm_mode=1 calls glibc cbrt/acos/cos/pow like the reference does; the expanded-node sequence must not depend on
    the <=1 ulp differences to csrc/fpmath.h on these cases (if it ever does, the digest differs and this documents it)."""
    case = GOLD[0]
    world, p, orc = setup_case(case, libm_mode=1)
    same = 0
    for q in case["queries"]:
        r = orc.search(q["start_pt"], q["start_vel"], q["end_pt"], q["end_vel"])
        assert r["status"] == q["status"]
        same += int(str(r["pop_hash"]) == q["pop_hash"] and r["n_path"] == q["n_path"])
    assert same >= len(case["queries"]) - 1


def test_path_properties():
    case = GOLD[0]
    world, p, orc = setup_case(case)
    for q in case["queries"]:
        r = orc.search(q["start_pt"], q["start_vel"], q["end_pt"], q["end_vel"])
        if r["status"] != 1:
            continue
        path = r["path"]
        assert np.allclose(path[0], q["start_pt"], atol=1e-12)               # samplePath starts at the start node
        assert np.linalg.norm(path[-1] - np.array(q["end_pt"])) < 0.8        # shot trajectory ends near the goal
        idx = np.floor((path - world.origin) / world.resolution).astype(int)
        assert (world.occ3[idx[:, 0], idx[:, 1], idx[:, 2]] != 1).mean() > 0.98   # sampled points avoid inflated voxels
        assert np.abs(np.diff(path, axis=0)).max() < p.max_velocity * p.time_step_size * 1.8


def ulp_dist(a, b):
    ia, ib = np.asarray(a, np.float64).view(np.int64), np.asarray(b, np.float64).view(np.int64)
    return np.abs(ia - ib)


def test_fpmath_within_one_ulp_of_glibc():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1e6, 1e6, 4000), rng.normal(size=4000), 10.0 ** rng.uniform(-300, 300, 2000)])
    got = oracle_lib.fp_eval(0, x)
    # glibc's own cbrt is only good to a few ulp (its libm-err table lists up to 4): compare with the EXACT cube root
    # (rational arithmetic) for tightness, and with glibc for the distance a reference build would see
    assert ulp_dist(got, [math.cbrt(v) for v in x]).max() <= 4
    from fractions import Fraction
    for v, y in zip(x[:3000], got[:3000]):
        fx, lo, hi = Fraction(v), Fraction(np.nextafter(y, -np.inf)), Fraction(np.nextafter(y, np.inf))
        if v < 0:
            lo, hi = hi, lo
        assert abs(lo) ** 3 < abs(fx) < abs(hi) ** 3     # the true root lies strictly within one ulp either side
    xa = rng.uniform(-1, 1, 8000)
    assert ulp_dist(oracle_lib.fp_eval(1, xa), [math.acos(v) for v in xa]).max() <= 1
    xc = rng.uniform(0, 2 * math.pi, 8000)      # cubic() calls cos((theta + 2k pi)/3): arguments in [0, 5pi/3]
    got, ref = oracle_lib.fp_eval(2, xc), np.array([math.cos(v) for v in xc])
    assert np.abs(got - ref).max() < 3e-16
    t = rng.uniform(0.01, 3.0, 4000)
    for n in range(0, 12):
        assert ulp_dist(oracle_lib.fp_eval(3, t, n_pow=n), [math.pow(v, n) for v in t]).max() <= 1


def test_inflation_rule_grid_map_cpp_733_785():
    """Every cloud point marks its 3x3x3 voxel neighbourhood (inflation 0.099 -> 1 voxel, inf_step_z = 1)."""
    world = u.make_world(20, 20, 5, seed=1)
    occ = np.zeros(world.dims, np.int8)
    pts = world.cloud.astype(np.float64)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                p = pts + np.array([dx, dy, dz]) * world.resolution
                idx = np.floor((p - world.origin) * (1.0 / world.resolution)).astype(int)
                ok = ((idx >= 0) & (idx < np.array(world.dims))).all(1)
                occ[idx[ok, 0], idx[ok, 1], idx[ok, 2]] = 1
    assert np.array_equal(occ.reshape(-1), world.occ)


def test_wall_map_has_the_gap():
    world = u.make_world(20, 20, 5, seed=1, map_type=2)     # random_forest.cpp:347-351, two slabs with a 0.5 m gap
    o = world.occ3
    ix = int((0.0 - world.origin[0]) / world.resolution)
    assert o[ix, int((5.0 - world.origin[1]) / 0.1), 10] == 1
    assert o[ix, int((0.0 - world.origin[1]) / 0.1), 10] == 0
