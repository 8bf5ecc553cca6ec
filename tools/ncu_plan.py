"""one launch of the bench's kernel (search + in-kernel QP, 4 096 queries) for ncu: python tools/ncu_plan.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import uav_motion_planning_b200 as u
from uav_motion_planning_b200.planner import plan_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = u.Context(0)
world = u.make_world(50, 50, 10, seed=1)
ka = u.KinoAstar(ctx); ka.setLaunchParams(); ka.setGridMap(world)
sp, sv, ep, ev = u.sample_queries(world, B, seed=11)
plan_batch(ctx, sp, sv, ep, ev, order=7, S=8)  # warm-up (ncu -s 1 skips it)
out = plan_batch(ctx, sp, sv, ep, ev, order=7, S=8)
print("kernel_ms", ctx.timings()["search_ms"], "solved", int(out["qp_solved"].sum()))
