"""The CPU restatements of the two row-f4 searches (oracle/astar_ref.cpp, oracle/rrt_star_ref.cpp) against golden vectors produced by
the REFERENCE's own a_star.cpp / rrt_star.cpp + kdtree.cpp builds (tests/golden/make_f4_golden.py, run where /root/reference exists).
Runs everywhere, also where oracle/_ref cannot be built."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib
import uav_motion_planning_b200 as u

G = os.path.join(os.path.dirname(__file__), "golden")
ASTAR = json.load(open(os.path.join(G, "astar_golden.json")))["cases"]
RRT = json.load(open(os.path.join(G, "rrt_star_golden.json")))["cases"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", ASTAR, ids=[c["name"] for c in ASTAR])
def test_astar_restatement_matches_the_reference_vectors(case):
    world = u.make_world(*case["dims"], seed=case["map_seed"])
    assert sha(world.occ) == case["occ_sha256"]
    for q in case["queries"]:
        r = oracle_lib.astar_search(world, np.array(q["start_pt"]), np.array(q["end_pt"]), **case["params"])
        assert (r["status"], r["use_node_num"], r["n_path"], str(r["lookup_digest"]), r["n_in_map_calls"]) == \
            (q["status"], q["use_node_num"], q["n_path"], q["lookup_digest"], q["n_in_map_calls"])
        assert sha(r["path"]) == q["path_sha256"]


@pytest.mark.parametrize("case", RRT, ids=[c["name"] for c in RRT])
def test_rrt_star_restatement_matches_the_reference_vectors(case):
    world = u.make_world(*case["dims"], seed=case["map_seed"])
    assert sha(world.occ) == case["occ_sha256"]
    for q in case["queries"]:
        r = oracle_lib.rrt_search(world, np.array(q["start_pt"]), np.array(q["end_pt"]), q["query_seed"], **case["params"])
        assert (r["status"], r["use_node_num"], r["n_samples"], r["reach_goal"], str(r["tree_digest"]), r["n_opt_path"]) == \
            (q["status"], q["use_node_num"], q["n_samples"], q["reach_goal"], q["tree_digest"], q["n_opt_path"])
        assert str(np.float64(r["goal_g_cost"]).view(np.uint64)) == q["goal_g_cost_bits"]
        assert sha(r["opt_path"]) == q["opt_path_sha256"]
