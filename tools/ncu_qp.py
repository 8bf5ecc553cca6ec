"""one QP launch for ncu: python tools/ncu_qp.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import uav_motion_planning_b200 as u
from uav_motion_planning_b200.minimum_control import MinimumControl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 12288
ctx = u.Context(0)
rng = np.random.default_rng(3)
S = 8
pos = np.cumsum(rng.normal(size=(B, S + 1)), axis=1)
z = np.zeros((B, 2))
mc = MinimumControl(ctx, order=7)
r = mc.solve_batch(pos, z, z, np.ones((B, S)), bound_jerk=z)
print("qp_ms", ctx.timings()["qp_ms"], "iters mean", r["iters"].mean(), "solved", r["solved"].mean())
