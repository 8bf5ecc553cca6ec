"""Builds and binds tests/host/qp_host.cpp: the product's QP kernel body compiled for the host (test harness only)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "qp_host.cpp")
CS = os.path.join(ROOT, "uav_motion_planning_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "host", "libqp_host.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        deps = [SRC] + [os.path.join(CS, f) for f in ("qp_body.h", "qp_body_warp.h", "qp_plan.h", "qp_symbolic.cpp", "fpmath.h", "amd_perm_table.inc")]
        if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
            subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared",
                            "-I" + os.path.join(ROOT, "include"), "-I" + CS, "-o", OUT, SRC,
                            os.path.join(CS, "qp_symbolic.cpp")], check=True)
        _lib = C.CDLL(OUT)
    return _lib


def solve_batch(order, pos, bv, ba, T, bj=None, settings=None, lo=None, hi=None, n_corridor=0):
    from uav_motion_planning_b200.minimum_control import default_settings
    lib = load()
    f = lambda a: np.ascontiguousarray(a, np.float64)
    pos, bv, ba, T = f(pos), f(bv), f(ba), f(T)
    B, S = pos.shape[0], pos.shape[1] - 1
    lo, hi = (f(lo), f(hi)) if n_corridor else (None, None)
    bj = f(np.zeros((B, 2)) if bj is None else bj)
    n = (order + 1) * S
    coef = np.zeros((B, n))
    solved, status, iters = (np.zeros(B, np.int32) for _ in range(3))
    stats = np.zeros(6, np.int32)
    st = settings or default_settings()
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    lib.host_qp_solve_c(order, S, n_corridor, B, p(pos), p(bv), p(ba), p(bj), p(T), p(lo), p(hi), C.byref(st), p(coef), p(solved),
                        p(status), p(iters), p(stats))
    return dict(coef=coef, solved=solved, status=status, iters=iters, stats=stats)


def solve_batch_warp(order, pos, bv, ba, T, bj=None, settings=None, reversed_loops=False, lo=None, hi=None, n_corridor=0):
    """The warp-per-problem body (qp_body_warp.h) run as one lane on the host; reversed_loops flips every parallel loop."""
    from uav_motion_planning_b200.minimum_control import default_settings
    lib = load()
    f = lambda a: np.ascontiguousarray(a, np.float64)
    pos, bv, ba, T = f(pos), f(bv), f(ba), f(T)
    B, S = pos.shape[0], pos.shape[1] - 1
    lo, hi = (f(lo), f(hi)) if n_corridor else (None, None)
    bj = f(np.zeros((B, 2)) if bj is None else bj)
    n = (order + 1) * S
    coef = np.zeros((B, n))
    solved, status, iters = (np.zeros(B, np.int32) for _ in range(3))
    st = settings or default_settings()
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    lib.host_qp_solve_warp_c(int(reversed_loops), order, S, n_corridor, B, p(pos), p(bv), p(ba), p(bj), p(T), p(lo), p(hi),
                             C.byref(st), p(coef), p(solved), p(status), p(iters))
    return dict(coef=coef, solved=solved, status=status, iters=iters)
