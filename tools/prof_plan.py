"""In-kernel phase profile of the PLAN pipeline (search + in-kernel QP) on the bench workload (run on the GPU box).
   python tools/prof_plan.py [B]   phase 'setup' then also holds the QP rounds (qp_round) of the CTAs"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import uav_motion_planning_b200 as u  # noqa: E402
from uav_motion_planning_b200.planner import plan_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = u.Context(0)
world = u.make_world(50, 50, 10, seed=1)
ka = u.KinoAstar(ctx)
ka.setLaunchParams()
ka.setGridMap(world)
sp, sv, ep, ev = u.sample_queries(world, B, seed=11)
plan_batch(ctx, sp, sv, ep, ev, order=7, S=8)  # warm-up
ka.setProfile(True)
out = plan_batch(ctx, sp, sv, ep, ev, order=7, S=8)
t = ctx.timings()
pr = ka.profile(B)
c = ka.counters()
tot = sum(v for k, v in pr["phase_cycles"].items() if k in ("pop", "shot_path", "tables_tile_grid", "cloud_ellipsoid", "dedup_probe_heuristic",
                                                            "node_write", "heap_commit", "setup", "cloud_staging"))
res = dict(B=B, kernel_ms=t["search_ms"], n_pop=c["n_pop"], qp_solved=int(out["qp_solved"].sum()),
           phase_cycles_per_pop={k: v / c["n_pop"] for k, v in pr["phase_cycles"].items()},
           setup_and_qp_cycles_per_query=pr["phase_cycles"]["setup"] / B, total_cta_cycles=tot,
           setup_and_qp_share=pr["phase_cycles"]["setup"] / tot)
print(json.dumps(res, indent=1))
