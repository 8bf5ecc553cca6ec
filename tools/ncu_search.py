"""one search launch for ncu: python tools/ncu_search.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import uav_motion_planning_b200 as u
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = u.Context(0)
world = u.make_world(50, 50, 10, seed=1)
ka = u.KinoAstar(ctx); ka.setLaunchParams(); ka.setGridMap(world)
sp, sv, ep, ev = u.sample_queries(world, B, seed=11)
ka.search_batch(sp, sv, ep, ev, want_paths=False)
print("search_ms", ctx.timings()["search_ms"], "pops", int(ka.last["n_pop"].sum()))
