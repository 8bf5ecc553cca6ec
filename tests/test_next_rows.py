"""SURVEY.md §8(f) "next" rows built so far: device-side map inflation (f-2) and batched PolyTraj evaluation (f-1)."""
import numpy as np
import pytest

import uav_motion_planning_b200 as u
from uav_motion_planning_b200.poly_traj import PolyTraj, evaluate_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dims,mtype", [((20, 20, 5), 0), ((50, 50, 10), 0), ((20, 20, 5), 2)])
def test_device_inflation_is_bit_identical(gpu_ctx, dims, mtype):
    """cloud -> occupancy_buffer_inflate_ on the GPU == the host restatement of grid_map.cpp:733-785 (integer/byte work:
    bit-exact), and a search on the device-built map reproduces the search on the host-built one."""
    world = u.make_world(*dims, seed=1, map_type=mtype)      # world.occ is the host inflation of world.cloud
    ka = u.KinoAstar(gpu_ctx)
    ka.setLaunchParams()
    ka.setGridMapFromCloud(world)
    assert np.array_equal(ka.occupancy(), world.occ)
    qs = u.sample_queries(world, 16, seed=41, min_dist=8.0)
    a = {k: (v.copy() if v is not None else None) for k, v in ka.search_batch(*qs).items()}
    ka.setGridMap(world)
    b = ka.search_batch(*qs)
    assert np.array_equal(a["pop_hash"], b["pop_hash"]) and np.array_equal(a["paths"], b["paths"])


def eigen_dot(a, b):
    """numpy restatement of Eigen 3.3's 2-wide linear vectorised reduction (believed; Eigen is unavailable here)."""
    p = a * b
    n = len(p)
    al2, al = (n // 4) * 4, (n // 2) * 2
    if al == 0:
        r = p[0]
        for k in range(1, n):
            r += p[k]
        return r
    r0 = np.array([p[0], p[1]])
    if al > 2:
        r1 = np.array([p[2], p[3]])
        for k in range(4, al2, 4):
            r0 = r0 + p[k:k + 2]
            r1 = r1 + p[k + 2:k + 4]
        r0 = r0 + r1
        if al > al2:
            r0 = r0 + p[al2:al2 + 2]
    r = r0[0] + r0[1]
    for k in range(al, n):
        r += p[k]
    return r


def ref_eval(coef, times, t, deriv):
    """poly_traj.hpp:74-168, statement by statement (one trajectory)."""
    S, nc = coef.shape[1], coef.shape[2]
    idx = 0
    while idx < S and t > times[idx] + 1e-4:
        t -= times[idx]
        idx += 1
    if idx == S:
        idx -= 1
        t = times[idx]
    n = nc - deriv
    tv = np.zeros(n)
    for i in range(n):
        tv[i] = 1.0 if i == 0 else tv[i - 1] * t
    out = np.zeros(3)
    for ax in range(3):
        c = coef[ax, idx]
        if deriv == 0:
            cv = c[:n].copy()
        elif deriv == 1:
            cv = np.array([float(i + 1) * c[i + 1] for i in range(n)])
        else:
            cv = np.array([float((i + 2) * (i + 1)) * c[i + 2] for i in range(n)])
        out[ax] = eigen_dot(tv, cv)
    return out


@pytest.mark.parametrize("order,S", [(5, 4), (7, 8)])
def test_polytraj_eval_batch(gpu_ctx, order, S):
    rng = np.random.default_rng(order)
    B = 33
    coef = rng.normal(size=(B, 3, S, order + 1))
    times = rng.uniform(0.5, 2.0, size=(B, S))
    t = np.concatenate([[0.0], rng.uniform(0, times.sum(1).min(), 40), [times.sum(1).max() + 1.0]])  # incl. past the end
    for deriv in (0, 1, 2):
        got = evaluate_batch(gpu_ctx, coef, times, t, deriv)
        for b in (0, 7, B - 1):
            for k in range(len(t)):
                ref = ref_eval(coef[b], times[b], t[k], deriv)
                # floating point: tolerance 1e-12 relative (the association is Eigen's as far as can be told without Eigen)
                assert np.allclose(got[b, k], ref, rtol=1e-12, atol=1e-12)


def test_polytraj_consumes_plan_output(gpu_ctx):
    """search -> QP -> evaluate: the trajectory starts at the query's start point and ends where the path ended."""
    from uav_motion_planning_b200.planner import plan_batch
    world = u.make_world(20, 20, 5, seed=1)
    ka = u.KinoAstar(gpu_ctx)
    ka.setLaunchParams()
    ka.setGridMap(world)
    sp, sv, ep, ev = u.sample_queries(world, 8, seed=43, min_dist=8.0)
    S, order = 8, 7
    r = plan_batch(gpu_ctx, sp, sv, ep, ev, order=order, S=S)
    ok = r["qp_solved"] == 1
    coef = r["coef"].reshape(-1, 3, S, order + 1)
    pos = evaluate_batch(gpu_ctx, coef, np.ones((len(sp), S)), np.array([0.0, float(S)]), 0)
    assert np.abs(pos[ok, 0] - sp[ok]).max() < 5e-2
    pt = PolyTraj(gpu_ctx)
    for s in range(S):
        pt.addSegment(coef[0, 0, s], coef[0, 1, s], coef[0, 2, s], 1.0)
    pt.init()
    assert np.allclose(pt.evaluatePos(0.0), pos[0, 0]) and pt.getTotalTIme() == float(S)
