"""Generates tests/golden/astar_golden.json and rrt_star_golden.json from the REFERENCE ITSELF (oracle/_ref/libastar_ref.so and
librrt_ref.so: the reference's a_star.cpp / rrt_star.cpp + kdtree.cpp compiled unmodified, see oracle/Makefile), run here where
/root/reference exists.  Committed together with its outputs; tests never regenerate them.  tests/test_f4_golden.py checks the CPU
restatements against these files everywhere (the GPU box has no /root/reference), tests/test_astar_parity.py / test_rrt_star_parity.py
check the CUDA kernels against them.

  python tests/golden/make_f4_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib  # noqa: E402
import uav_motion_planning_b200 as u  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def astar_cases():
    out = []
    for name, dims, nq, qseed, md, kw in [("small", (20, 20, 5), 10, 3, 5.0, {}),
                                          ("small_lambda2.5", (20, 20, 5), 6, 5, 8.0, dict(lambda_heu=2.5)),
                                          ("small_pool500", (20, 20, 5), 6, 4, 8.0, dict(allocated_node_num=500)),
                                          ("big", (50, 50, 10), 4, 8, 10.0, {})]:
        world = u.make_world(*dims, seed=1)
        sp, _, ep, _ = u.sample_queries(world, nq, seed=qseed, min_dist=md)
        if name == "small":
            ep[0] = (100.0, 0.0, 1.0)  # end point outside the map (a_star.cpp:52-56)
        qs = []
        for q in range(nq):
            r = oracle_lib.astar_search_reference(world, sp[q], ep[q], **kw)
            qs.append(dict(start_pt=sp[q].tolist(), end_pt=ep[q].tolist(), status=r["status"], use_node_num=r["use_node_num"],
                           n_path=r["n_path"], path_sha256=sha(r["path"]), lookup_digest=str(r["lookup_digest"]),
                           n_in_map_calls=r["n_in_map_calls"]))
        out.append(dict(name=name, dims=list(dims), map_seed=1, params=kw, occ_sha256=sha(world.occ), queries=qs))
    return out


def rrt_cases():
    out = []
    for name, nq, qseed, md, kw in [("loop8000", 6, 3, 5.0, dict(max_tree_node_num=8000, sample_budget=8000)),
                                    ("budget3000", 6, 4, 4.0, dict(max_tree_node_num=20000, sample_budget=3000)),
                                    ("tree300", 4, 5, 8.0, dict(max_tree_node_num=300, sample_budget=300)),
                                    ("step0.4_radius0.9_res0.01", 4, 9, 5.0, dict(max_tree_node_num=4000, sample_budget=4000, step_length=0.4,
                                                                                    search_radius=0.9, collision_check_resolution=0.01))]:
        world = u.make_world(20, 20, 5, seed=1)
        sp, _, ep, _ = u.sample_queries(world, nq, seed=qseed, min_dist=md)
        qs = []
        for q in range(nq):
            seed = 7919 * q + qseed
            r = oracle_lib.rrt_search_reference(world, sp[q], ep[q], seed, **kw)
            qs.append(dict(start_pt=sp[q].tolist(), end_pt=ep[q].tolist(), query_seed=seed, status=r["status"], use_node_num=r["use_node_num"],
                           n_samples=r["n_samples"], reach_goal=r["reach_goal"], goal_g_cost_bits=str(np.float64(r["goal_g_cost"]).view(np.uint64)),
                           tree_digest=str(r["tree_digest"]), n_opt_path=r["n_opt_path"], opt_path_sha256=sha(r["opt_path"])))
        out.append(dict(name=name, dims=[20, 20, 5], map_seed=1, params=kw, occ_sha256=sha(world.occ), queries=qs))
    return out


if __name__ == "__main__":
    assert oracle_lib.have_astar_ref() and oracle_lib.have_rrt_ref(), "build oracle/_ref first (make -C oracle; needs /root/reference)"
    json.dump(dict(source="oracle/_ref/libastar_ref.so = the reference's a_star.cpp compiled unmodified", cases=astar_cases()),
              open(os.path.join(HERE, "astar_golden.json"), "w"), indent=1)
    json.dump(dict(source="oracle/_ref/librrt_ref.so = the reference's rrt_star.cpp + kdtree.cpp compiled unmodified (seeded sample stream, sample budget)",
                   cases=rrt_cases()), open(os.path.join(HERE, "rrt_star_golden.json"), "w"), indent=1)
    print("written")
