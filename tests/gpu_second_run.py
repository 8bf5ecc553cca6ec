"""second-contact diagnostics (not a pytest file): closure fix for K1 + first K2 parity numbers"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uav_motion_planning_b200 as u, oracle_lib
from uav_motion_planning_b200.minimum_control import MinimumControl, default_settings
from test_kino_parity import run_case
from test_qp_parity import make_problems
ctx = u.Context(0)
for (X, Y, Z, n, ctype, md) in [(20, 20, 5, 32, 1, 8.0), (20, 20, 5, 32, 2, 8.0), (50, 50, 10, 96, 1, 10.0)]:
    w = u.make_world(X, Y, Z, seed=1)
    bad, got, ka, orc, qs = run_case(ctx, w, n, seed=2, ctype=ctype, min_dist=md)
    print(f"map {X}x{Y}x{Z} ctype {ctype}: {n} queries, mismatches {len(bad)}", flush=True)
    for b in bad[:6]:
        print("  BAD (q, st_ref, st_gpu, use_ref, use_gpu, npop_ref, npop_gpu, npath_ref, npath_gpu)", b)
    if bad:
        # first divergence of the pop traces of the first bad query
        q = bad[0][0]
        ka.setTrace(256); ka.setGridMap(w)
        sp, sv, ep, ev = qs
        ka.search_batch(sp, sv, ep, ev)
        ref = orc.search(sp[q], sv[q], ep[q], ev[q], pop_cap=256)
        tr = ka.pop_trace(q, 256)
        nn = min(len(ref["trace"]), 256)
        d = np.nonzero((tr[:nn] != ref["trace"][:nn]).any(1))[0]
        print("  first trace divergence at pop", d[:1], "of", ref["n_pop"]); ka.setTrace(0)
for order, S in [(5, 3), (5, 4), (7, 8), (7, 16)]:
    B = 64
    pos, bv, ba, bj, T = make_problems(B, S, 1)
    mc = MinimumControl(ctx, order=order)
    got = mc.solve_batch(pos, bv, ba, T, bound_jerk=bj)
    errs, it_bad = [], 0
    for b in range(B):
        ok, coef, info = oracle_lib.minctrl_solve(order, S, pos[b], bv[b], ba[b], T[b], bound_jerk=bj[b])
        errs.append(np.abs(coef - got["coef"][b]).max() / np.abs(coef).max())
        it_bad += int(info["iter"] != got["iters"][b] or info["status_val"] != got["status"][b])
    print(f"QP order {order} S {S}: max rel err {max(errs):.3e}, iter/status mismatches {it_bad}, iters {np.bincount(got['iters'])[::25]}, qp_ms {ctx.timings()['qp_ms']:.3f}", flush=True)
for order, S, B in [(7, 8, 12288), (5, 4, 12288)]:
    pos, bv, ba, bj, T = make_problems(B, S, 3)
    mc = MinimumControl(ctx, order=order)
    for _ in range(2):
        got = mc.solve_batch(pos, bv, ba, T, bound_jerk=bj)
    print(f"QP throughput order {order} S {S} B {B}: qp_ms {ctx.timings()['qp_ms']:.3f} iters mean {got['iters'].mean():.1f} solved {got['solved'].mean():.3f}")
w = u.make_world(50, 50, 10, seed=1)
ka = u.KinoAstar(ctx); ka.setLaunchParams(); ka.setGridMap(w)
sp, sv, ep, ev = u.sample_queries(w, 4096, seed=11)
ka.search_batch(sp, sv, ep, ev, want_paths=False)
t = ctx.timings(); print(f"B=4096: search {t['search_ms']:.2f} ms; pops max {ka.last['n_pop'].max()}")
