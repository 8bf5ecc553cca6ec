import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import uav_motion_planning_b200 as u, oracle_lib
from test_kino_parity import run_case
ctx = u.Context(0)
w = u.make_world(20, 20, 5, seed=1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bad, got, ka, orc, qs = run_case(ctx, w, n, seed=2, ctype=1, min_dist=8.0)
print("mismatches", len(bad), bad[:4], "status", got["status"][:8], "pops", got["n_pop"][:8], flush=True)
