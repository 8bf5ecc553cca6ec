"""Timing of the batched grid A* (K3) on the 50 x 50 x 10 m map, with the CPU oracle timed beside it on a bounded sample.
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
    world = u.make_world(50, 50, 10, seed=1)
    a = u.Astar()
    a.setGridMap(world)
    sp, _, ep, _ = u.sample_queries(world, B, seed=8)
    a.search_batch(sp, ep, want_paths=False)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        r = a.search_batch(sp, ep)
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    out = dict(workload=f"{B} grid A* queries, 50x50x10 m map", queries_per_s=B / t, ms=t * 1e3, expansions=int(r["n_pop"].sum()),
               expansions_per_s=float(r["n_pop"].sum() / t), reach_end_frac=float((r["status"] == 1).mean()))
    try:
        import oracle_lib
        n = 64
        t0 = time.perf_counter()
        for q in range(n):
            oracle_lib.astar_search(world, sp[q], ep[q])
        tc = time.perf_counter() - t0
        out["cpu_oracle_queries_per_s_1core"] = n / tc
    except Exception as e:  # noqa: BLE001
        out["cpu_oracle"] = repr(e)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
