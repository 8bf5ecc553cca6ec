"""first-contact diagnostics on the GPU box (not a pytest file)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uav_motion_planning_b200 as u, oracle_lib
from test_kino_parity import run_case
ctx = u.Context(0)
for (X, Y, Z, n, ctype, md) in [(20, 20, 5, 16, 1, 8.0), (20, 20, 5, 16, 2, 8.0), (50, 50, 10, 64, 1, 10.0)]:
    w = u.make_world(X, Y, Z, seed=1)
    t0 = time.time()
    bad, got, ka, orc, qs = run_case(ctx, w, n, seed=2, ctype=ctype, min_dist=md)
    print(f"map {X}x{Y}x{Z} ctype {ctype}: {n} queries, mismatches {len(bad)}, timings {ctx.timings()}", flush=True)
    print("  status", np.bincount(got["status"], minlength=3), "pops", got["n_pop"][:16], flush=True)
    print("  counters", ka.counters())
    for b in bad[:8]:
        print("  BAD (q, st_ref, st_gpu, use_ref, use_gpu, npop_ref, npop_gpu, npath_ref, npath_gpu)", b)
# throughput probe
w = u.make_world(50, 50, 10, seed=1)
ka = u.KinoAstar(ctx); ka.setLaunchParams(); ka.setGridMap(w)
for B in (256, 1024, 4096):
    sp, sv, ep, ev = u.sample_queries(w, B, seed=11)
    ka.search_batch(sp, sv, ep, ev, want_paths=False)
    t = ctx.timings()
    print(f"B={B}: search {t['search_ms']:.2f} ms -> {B / t['search_ms'] * 1e3:.0f} searches/s; total {t['total_ms']:.2f} ms; pops mean {ka.last['n_pop'].mean():.1f} max {ka.last['n_pop'].max()} status {np.bincount(ka.last['status'], minlength=3)}", flush=True)
    print("   counters", ka.counters())
