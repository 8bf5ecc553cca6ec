"""BASELINE.md run C1: BASELINE.json configs[0] on the CPU path, single thread — one start->goal at a time on the 20x20x5 m random
map @0.1 m (seed 1): kino-A* (launch-file parameters) + 4-segment min-jerk x 3 axes, T_i = 1.0.  Prints latency statistics."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
import uav_motion_planning_b200 as u  # noqa: E402
from pipeline_ref import waypoints_from_path  # noqa: E402
from uav_motion_planning_b200 import _lib  # noqa: E402

world = u.make_world(20, 20, 5, seed=1)
p = _lib.KinoParams()
u.load().uavmp_kino_params_launch(C.byref(p))
orc = oracle_lib.KinoOracle(world, p)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
sp, sv, ep, ev = u.sample_queries(world, n, seed=2, min_dist=10.0)
ts, tq, pops, iters, reached = [], [], [], [], 0
for q in range(n):
    t0 = time.perf_counter()
    r = orc.search(sp[q], sv[q], ep[q], ev[q])
    t1 = time.perf_counter()
    ts.append(t1 - t0)
    pops.append(r["n_pop"])
    if r["status"] == 1:
        reached += 1
        wp = waypoints_from_path(r["path"], 4)
        t2 = time.perf_counter()
        for ax in range(3):
            ok, c, info = oracle_lib.minctrl_solve(5, 4, wp[:, ax], [sv[q][ax], ev[q][ax]], [0, 0], np.ones(4))
            iters.append(info["iter"])
        tq.append(time.perf_counter() - t2)
out = dict(config="configs[0]: single start->goal, 20x20x5 m random map @0.1 m, kino-A* + 4-seg min-jerk (3 axes), 1 CPU thread",
           queries=n, reach_end=reached, search_ms_mean=1e3 * float(np.mean(ts)), search_ms_median=1e3 * float(np.median(ts)),
           search_ms_p95=1e3 * float(np.percentile(ts, 95)), expansions_mean=float(np.mean(pops)),
           qp3_ms_mean=1e3 * float(np.mean(tq)), admm_iters_mean=float(np.mean(iters)),
           plans_per_s_single_thread=n / (sum(ts) + sum(tq)), cpu=os.uname().machine, note="oracle = restated search + the reference's OSQP; "
           "stdout dumps of the reference (minimum_control.cpp:154-158, OSQP verbose) are not part of the timed region")
print(json.dumps(out, indent=1))
