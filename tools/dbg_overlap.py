import os, sys, ctypes as C
os.environ["UAVMP_DBG_OVERLAP"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import uav_motion_planning_b200 as u
from uav_motion_planning_b200.planner import plan_batch
ctx = u.Context(0)
world = u.make_world(50, 50, 10, seed=1)
ka = u.KinoAstar(ctx); ka.setLaunchParams(); ka.setGridMap(world)
sp, sv, ep, ev = u.sample_queries(world, 4096, seed=11)
for it in range(3):
    r = plan_batch(ctx, sp, sv, ep, ev, order=7, S=8)
    out = (C.c_ulonglong * 4)()
    ctx.lib.uavmp_debug_overlap.argtypes = [C.c_void_p, C.c_void_p]
    ctx.lib.uavmp_debug_overlap(ctx.h, out)
    q0, q1, s0, s1 = [int(x) for x in out]
    print(f"iter {it}: search CTA exits first {0:.1f} last {(s1 - s0) / 1e6:.1f} ms | QP first CTA start {(q0 - s0) / 1e6:.1f} ms, last CTA start {(q1 - s0) / 1e6:.1f} ms (relative to the first search CTA exit)", ctx.timings())
