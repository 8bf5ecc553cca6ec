"""In-kernel phase profile of the search kernel on the bench workload (run on the GPU box).
   python tools/prof_search.py [B]            -> gpurun_out/prof_search.json + a text summary"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import uav_motion_planning_b200 as u  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = u.Context(0)
world = u.make_world(50, 50, 10, seed=1)
ka = u.KinoAstar(ctx)
ka.setLaunchParams()
ka.setGridMap(world)
sp, sv, ep, ev = u.sample_queries(world, B, seed=11)
ka.search_batch(sp, sv, ep, ev, want_paths=False)  # warm-up
ka.setProfile(True)
r = ka.search_batch(sp, sv, ep, ev, want_paths=False)
t = ctx.timings()
pr = ka.profile(B)
c = ka.counters()
qc = pr["query_cycles"].astype(np.float64)
pops = r["n_pop"].astype(np.float64)
diag = {k: pr["phase_cycles"].pop(k) for k in ("n_staged", "commit_replay", "sum_npts", "sum_flagged_prims", "commit_closure_io", "commit_slow_updates", "commit_deferred_writes")}
tot = sum(pr["phase_cycles"].values())
clk_ghz = qc.max() / (t["search_ms"] * 1e6)  # lower bound on the SM clock: the longest query cannot outlast the kernel
out = dict(B=B, diag=diag, search_ms=t["search_ms"], grid=pr["grid"], counters=c,
           phase_share={k: v / tot for k, v in pr["phase_cycles"].items()},
           phase_cycles_per_pop={k: v / c["n_pop"] for k, v in pr["phase_cycles"].items()},
           cycles_per_pop_mean=qc.sum() / pops.sum(), longest_query_cycles=qc.max(), longest_query_pops=float(pops[qc.argmax()]),
           sum_query_cycles=qc.sum(), cta_busy_frac=qc.sum() / (pr["grid"] * qc.max()),
           implied_clock_ghz_lower_bound=clk_ghz,
           pops_percentiles={str(p): float(np.percentile(pops, p)) for p in (50, 90, 99, 99.9, 100)},
           status_hist=np.bincount(r["status"], minlength=3).tolist())
top = np.argsort(-qc)[:3]
out["longest_queries"] = [dict(q=int(q), pops=int(pops[q]), cycles=float(qc[q]), status=int(r["status"][q]),
                              per_pop={n: float(pr["query_phase"][q][k]) / max(pops[q], 1) for k, n in enumerate(pr["names"])})
                         for q in top]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"prof_search_B{B}.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
