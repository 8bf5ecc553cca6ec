"""Generates tests/golden/*.json from the CPU oracle (run here, where /root/reference exists so that oracle/_ref holds
the reference's own OSQP).  Committed together with its outputs; tests never regenerate them.

  python tests/golden/make_golden.py
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib  # noqa: E402
import uav_motion_planning_b200 as u  # noqa: E402
from uav_motion_planning_b200 import _lib  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def kino_cases():
    out = []
    for name, dims, seed, ctype, launch, nq, qseed, md, alloc in [
            ("small_type1", (20, 20, 5), 1, 1, True, 12, 2, 8.0, None),
            ("small_type2", (20, 20, 5), 1, 2, True, 8, 3, 8.0, None),
            ("small_defaults", (20, 20, 5), 1, 1, False, 8, 6, 8.0, None),
            ("small_pool3000", (20, 20, 5), 1, 1, True, 6, 5, 8.0, 3000),
            ("wall_type1", (20, 20, 5), 1, 1, True, 6, 9, 8.0, None),
            ("big_type1", (50, 50, 10), 1, 1, True, 6, 4, 10.0, None)]:
        world = u.make_world(*dims, seed=seed, map_type=2 if name.startswith("wall") else 0)
        p = _lib.KinoParams()
        (u.load().uavmp_kino_params_launch if launch else u.load().uavmp_kino_params_default)(C.byref(p))
        p.collision_check_type = ctype
        if alloc:
            p.allocated_node_num = alloc
        sp, sv, ep, ev = u.sample_queries(world, nq, seed=qseed, min_dist=md)
        if name == "small_type1":  # the reference's only fixed query (test_kino_astar.cpp:21-25)
            sp[0], sv[0], ep[0], ev[0] = (0, 0, 1), (1, 1, 0), (5, 5, 0.8), (0, 0, 0)
        orc = oracle_lib.KinoOracle(world, p)
        qs = []
        for q in range(nq):
            r = orc.search(sp[q], sv[q], ep[q], ev[q])
            qs.append(dict(start_pt=sp[q].tolist(), start_vel=sv[q].tolist(), end_pt=ep[q].tolist(),
                           end_vel=ev[q].tolist(), status=r["status"], use_node_num=r["use_node_num"],
                           n_pop=r["n_pop"], pop_hash=str(r["pop_hash"]), n_path=r["n_path"],
                           path_sha256=sha(r["path"]), path_first=r["path"][0].tolist() if r["n_path"] else None,
                           path_last=r["path"][-1].tolist() if r["n_path"] else None, counters=r["counters"]))
        out.append(dict(name=name, dims=list(dims), map_seed=seed, map_type=2 if name.startswith("wall") else 0,
                        launch_params=launch, collision_check_type=ctype, allocated_node_num=alloc,
                        occ_sha256=sha(world.occ), cloud_sha256=sha(world.cloud), n_cloud=len(world.cloud),
                        queries=qs))
    return out


def qp_cases():
    out = []
    rng = np.random.default_rng(1234)
    for order, S, unit, kw in [(5, 3, True, {}), (5, 4, True, {}), (5, 8, False, {}), (7, 8, True, {}),
                               (7, 12, False, {}), (7, 16, False, dict(eps_abs=1e-5, eps_rel=1e-5, max_iter=4000))]:
        probs = []
        for b in range(4):
            pos = np.cumsum(rng.normal(size=S + 1))
            if order == 5 and S == 3 and b == 0:  # test_qpsolve.cpp:10-18
                pos = np.array([1.0, 2.0, 3.0, 4.0])
                bv, ba = np.zeros(2), np.zeros(2)
            else:
                bv, ba = rng.normal(size=2) * 0.5, rng.normal(size=2) * 0.2
            bj = np.zeros(2)
            T = np.ones(S) if unit else rng.uniform(0.5, 2.0, size=S)
            ok, coef, info = oracle_lib.minctrl_solve(order, S, pos, bv, ba, T, bound_jerk=bj,
                                                      settings=oracle_lib.osqp_settings(**kw))
            asm = oracle_lib.minctrl_assemble(order, S, pos, bv, ba, T, bound_jerk=bj)
            probs.append(dict(pos=pos.tolist(), bound_vel=bv.tolist(), bound_acc=ba.tolist(), bound_jerk=bj.tolist(),
                              T=T.tolist(), solved=int(ok), status_val=info["status_val"], iter=info["iter"],
                              rho_updates=info["rho_updates"], coef=coef.tolist(), nnzP=int(len(asm["Px"])),
                              nnzA=int(len(asm["Ax"])), P_sha256=sha(asm["Px"]), A_sha256=sha(asm["Ax"])))
        out.append(dict(order=order, S=S, settings=kw, problems=probs))
    return out


if __name__ == "__main__":
    assert oracle_lib.have_ref(), "oracle/_ref/libosqp_ref.so missing: run make -C oracle where /root/reference exists"
    json.dump(kino_cases(), open(os.path.join(HERE, "kino_golden.json"), "w"), indent=1)
    json.dump(qp_cases(), open(os.path.join(HERE, "minctrl_golden.json"), "w"), indent=1)
    print("wrote golden vectors")
