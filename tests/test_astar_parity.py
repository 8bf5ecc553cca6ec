"""GPU parity: batched CUDA Astar::search (uavmp_astar_search_batch) vs the CPU oracle (oracle/astar_ref.cpp, itself pinned to the
reference's a_star.cpp): status, use_node_num, number and digest of the ordered expansions (exact position and g bits of every popped
node) and every path point, bit for bit."""
import numpy as np
import pytest

import oracle_lib
import uav_motion_planning_b200 as u

pytestmark = pytest.mark.gpu


def compare(ctx, world, n, seed, min_dist, **kw):
    a = u.Astar(ctx)
    a.setParam(**kw)
    a.setGridMap(world)
    sp, _, ep, _ = u.sample_queries(world, n, seed=seed, min_dist=min_dist)
    got = a.search_batch(sp, ep)
    okw = {k: v for k, v in kw.items() if k in ("lambda_heu", "allocated_node_num")}
    for q in range(n):
        ref = oracle_lib.astar_search(world, sp[q], ep[q], **okw)
        o0, o1 = got["path_offsets"][q], got["path_offsets"][q + 1]
        assert (ref["status"], ref["use_node_num"], ref["n_pop"], ref["pop_hash"]) == \
            (got["status"][q], got["use_node_num"][q], got["n_pop"][q], int(got["pop_hash"][q])), q
        assert ref["n_path"] == o1 - o0 and np.array_equal(ref["path"].view(np.uint64), got["paths"][o0:o1].view(np.uint64)), q
    return got


def test_small_map(gpu_ctx):
    world = u.make_world(20, 20, 5, seed=1)
    got = compare(gpu_ctx, world, 96, seed=3, min_dist=5.0)
    assert (got["status"] == 1).mean() > 0.8


def test_weighted_heuristic_and_pool_exhaustion(gpu_ctx):
    world = u.make_world(20, 20, 5, seed=1)
    compare(gpu_ctx, world, 32, seed=5, min_dist=8.0, lambda_heu=2.5)
    got = compare(gpu_ctx, world, 24, seed=4, min_dist=8.0, allocated_node_num=500)
    assert (got["status"] == 2).all() and (got["use_node_num"] == 500).all()


def test_goal_outside_the_map_and_reference_class_call(gpu_ctx):
    world = u.make_world(20, 20, 5, seed=1)
    a = u.Astar(gpu_ctx)
    a.setGridMap(world)
    a.init()
    sp, _, ep, _ = u.sample_queries(world, 2, seed=6, min_dist=5.0)
    path = []
    assert a.search(sp[0], np.array([100.0, 0.0, 1.0]), path) == 2 and not path   # a_star.cpp:52-56
    st = a.search(sp[0], ep[0], path)
    ref = oracle_lib.astar_search(world, sp[0], ep[0])
    assert st == ref["status"] and np.array_equal(np.array(path).view(np.uint64), ref["path"].view(np.uint64))


def test_big_map_batch(gpu_ctx):
    """1 024 queries on the 50 x 50 x 10 m map: properties over the whole batch, oracle parity on a subset."""
    world = u.make_world(50, 50, 10, seed=1)
    a = u.Astar(gpu_ctx)
    a.setGridMap(world)
    sp, _, ep, _ = u.sample_queries(world, 1024, seed=8)
    got = a.search_batch(sp, ep)
    again = a.search_batch(sp, ep, want_paths=False)
    assert np.array_equal(got["pop_hash"], again["pop_hash"]) and np.array_equal(got["use_node_num"], again["use_node_num"])
    ok = np.nonzero(got["status"] == 1)[0]
    off = got["path_offsets"]
    assert len(ok) > 900
    assert np.array_equal(got["paths"][off[ok]], sp[ok])                                   # every path starts at its start point
    assert np.abs(got["paths"][off[ok + 1] - 1] - ep[ok]).max() < world.resolution          # and ends within one cell of the goal
    for q in range(0, 1024, 64):
        ref = oracle_lib.astar_search(world, sp[q], ep[q])
        o0, o1 = off[q], off[q + 1]
        assert (ref["status"], ref["use_node_num"], ref["pop_hash"]) == (got["status"][q], got["use_node_num"][q], int(got["pop_hash"][q]))
        assert np.array_equal(ref["path"].view(np.uint64), got["paths"][o0:o1].view(np.uint64))


def test_golden_vectors_of_the_reference_build(gpu_ctx):
    """tests/golden/astar_golden.json: outputs of the reference's own a_star.cpp build (make_f4_golden.py)"""
    import hashlib
    import json
    import os
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "astar_golden.json")))["cases"]
    for case in cases:
        world = u.make_world(*case["dims"], seed=case["map_seed"])
        a = u.Astar(gpu_ctx)
        a.setParam(**case["params"])
        a.setGridMap(world)
        sp = np.array([q["start_pt"] for q in case["queries"]])
        ep = np.array([q["end_pt"] for q in case["queries"]])
        got = a.search_batch(sp, ep)
        for i, q in enumerate(case["queries"]):
            o0, o1 = got["path_offsets"][i], got["path_offsets"][i + 1]
            assert (q["status"], q["use_node_num"], q["n_path"]) == (got["status"][i], got["use_node_num"][i], o1 - o0), (case["name"], i)
            assert hashlib.sha256(np.ascontiguousarray(got["paths"][o0:o1]).tobytes()).hexdigest() == q["path_sha256"], (case["name"], i)
    u.Astar(gpu_ctx).setParam(lambda_heu=1.0, allocated_node_num=100000)  # back to the defaults for the shared context
