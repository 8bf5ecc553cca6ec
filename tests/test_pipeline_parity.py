"""GPU parity of the chained pipeline (uavmp_plan_batch): search -> waypoints -> 3 x QP vs the same chain on the oracle."""
import numpy as np
import pytest

import oracle_lib
import uav_motion_planning_b200 as u
from pipeline_ref import plan_one
from uav_motion_planning_b200.planner import plan_batch

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.mark.parametrize("order,S", [(5, 4), (7, 8)])
def test_plan_batch_matches_oracle(gpu_ctx, order, S):
    world = u.make_world(20, 20, 5, seed=1)
    ka = u.KinoAstar(gpu_ctx)
    ka.setLaunchParams()
    ka.setGridMap(world)
    sp, sv, ep, ev = u.sample_queries(world, 24, seed=21, min_dist=8.0)
    sv[:, 0] = 0.5  # non-zero boundary velocity reaches the QP's start row
    got = plan_batch(gpu_ctx, sp, sv, ep, ev, order=order, S=S)
    orc = oracle_lib.KinoOracle(world, ka.params)
    n_checked = 0
    for q in range(24):
        st, solved, coef, _ = plan_one(orc, sp[q], sv[q], ep[q], ev[q], order, S, 1.0)
        assert st == got["search_status"][q]
        assert solved == got["qp_solved"][q]
        if solved:
            for ax in range(3):
                scale = max(np.abs(coef[ax]).max(), 1e-12)
                assert np.abs(coef[ax] - got["coef"][q, ax]).max() / scale < RTOL
            n_checked += 1
    assert n_checked >= 8


def test_full_batch_properties(gpu_ctx):
    """BASELINE.json configs[1] at full size (4096 queries, 50x50x10 m map): size-independent properties.

    * every solved trajectory interpolates its boundary conditions: segment 0 starts at the query's start point with
      the start velocity, the last segment ends at the last sampled path point;
    * C0..C3 continuity at every interior knot (the equality rows of the QP), to the OSQP tolerance;
    * idempotence: running the same batch twice gives bit-identical coefficients.
    """
    world = u.make_world(50, 50, 10, seed=1)
    ka = u.KinoAstar(gpu_ctx)
    ka.setLaunchParams()
    ka.setGridMap(world)
    B, S, order = 4096, 8, 7
    sp, sv, ep, ev = u.sample_queries(world, B, seed=11)
    a = plan_batch(gpu_ctx, sp, sv, ep, ev, order=order, S=S)
    b = plan_batch(gpu_ctx, sp, sv, ep, ev, order=order, S=S)
    assert np.array_equal(a["search_status"], b["search_status"])
    assert np.array_equal(a["coef"].view(np.uint64), b["coef"].view(np.uint64))
    ok = a["qp_solved"] == 1
    assert ok.mean() > 0.5
    c = a["coef"][ok].reshape(-1, 3, S, order + 1)
    # start position / velocity (local time 0 of segment 0)
    assert np.abs(c[:, :, 0, 0] - sp[ok]).max() < 5e-2   # eps_abs + eps_rel * |x|, |x| <= 25 m
    assert np.abs(c[:, :, 0, 1] - sv[ok]).max() < 5e-2
    # continuity of derivatives 0..3 at interior knots, T = 1: sum_j j!/(j-r)! c_j  ==  r! c'_r
    from math import factorial
    for r in range(4):
        lhs = sum(factorial(j) / factorial(j - r) * c[:, :, :-1, j] for j in range(r, order + 1))
        rhs = factorial(r) * c[:, :, 1:, r]
        scale = max(1.0, np.abs(rhs).max())
        assert np.abs(lhs - rhs).max() / scale < 5e-2, r


def test_sequential_and_overlapped_pipelines_agree_bitwise(gpu_ctx, monkeypatch):
    """UAVMP_NO_OVERLAP=1 runs search -> k_waypoints -> QP kernel -> k_scatter_plan back to back; the default runs the QP kernel
    on a second stream under the search kernel with per-query completion flags.  Same bits either way."""
    world = u.make_world(20, 20, 5, seed=1)
    ka = u.KinoAstar(gpu_ctx)
    ka.setLaunchParams()
    ka.setGridMap(world)
    sp, sv, ep, ev = u.sample_queries(world, 200, seed=51, min_dist=8.0)
    a = plan_batch(gpu_ctx, sp, sv, ep, ev, order=7, S=8)
    monkeypatch.setenv("UAVMP_NO_OVERLAP", "1")
    b = plan_batch(gpu_ctx, sp, sv, ep, ev, order=7, S=8)
    monkeypatch.setenv("UAVMP_QP_THREAD", "1")   # and with the thread-per-problem QP kernel
    c = plan_batch(gpu_ctx, sp, sv, ep, ev, order=7, S=8)
    for other in (b, c):
        assert np.array_equal(a["search_status"], other["search_status"]) and np.array_equal(a["qp_solved"], other["qp_solved"])
        ok = a["qp_solved"] == 1
        assert np.array_equal(a["coef"][ok], other["coef"][ok])
