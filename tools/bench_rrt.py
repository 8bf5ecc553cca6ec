"""Timing of the batched RRT* (K4) on the 50 x 50 x 10 m map, with the CPU oracle timed beside it on a bounded sample.
Not the bench line (bench.py is the kino-A* + QP pipeline); prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."), os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")]
import uav_motion_planning_b200 as u  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    world = u.make_world(50, 50, 10, seed=1)
    r = u.RRTStar()
    r.setParam(max_tree_node_num=nodes, sample_budget=nodes)
    r.setGridMap(world)
    sp, _, ep, _ = u.sample_queries(world, B, seed=8)
    seeds = np.arange(B, dtype=np.uint64) + np.uint64(1)
    r.search_batch(sp[:64], ep[:64], seeds[:64], want_paths=False)
    t0 = time.perf_counter()
    got = r.search_batch(sp, ep, seeds)
    t = time.perf_counter() - t0
    out = dict(workload=f"{B} RRT* queries x {nodes} samples, 50x50x10 m map", queries_per_s=B / t, ms=t * 1e3,
               samples_per_s=float(got["n_samples"].sum() / t), tree_nodes=int(got["use_node_num"].sum()),
               reach_end_frac=float((got["status"] == 1).mean()))
    try:
        import oracle_lib
        n = 8
        t0 = time.perf_counter()
        ns = 0
        for q in range(n):
            ns += oracle_lib.rrt_search(world, sp[q], ep[q], int(seeds[q]), max_tree_node_num=nodes, sample_budget=nodes)["n_samples"]
        tc = time.perf_counter() - t0
        out["cpu_oracle_queries_per_s_1core"] = n / tc
        out["cpu_oracle_samples_per_s_1core"] = ns / tc
    except Exception as e:  # noqa: BLE001
        out["cpu_oracle"] = repr(e)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
