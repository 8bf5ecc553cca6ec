"""search -> waypoints -> 3 x QP pipeline over the C-ABI (uavmp_plan_batch / uavmp_plan_batch_dev).

This chaining is an EXTENSION: the reference never feeds KinoAstar's path into MinimumControl (its QP front-end is
RRT*, src/planner/test/src/test_minimum_jerk.cpp:40-75).  The rule used here (and restated for the oracle in
tests/pipeline_ref.py): for a query that reaches the goal with n sampled path points, waypoint k = path[floor(k*(n-1)/S)],
T_i = seg_time (1.0 is the reference's convention, test_minimum_jerk.cpp:66-71), boundary velocity = start/end velocity,
boundary acceleration (and jerk) = 0.
"""
import ctypes as C

import numpy as np

from . import _lib


def plan_batch(ctx, start_pt, start_vel, end_pt, end_vel, order=7, S=8, seg_time=1.0, settings=None):
    """Host buffers in, host buffers out (copies are inside the call)."""
    sp, sv, ep, ev = (_lib.as_f64(a).reshape(-1, 3) for a in (start_pt, start_vel, end_pt, end_vel))
    B = sp.shape[0]
    n = (order + 1) * S
    status = np.zeros(B, np.int32)
    solved = np.zeros(B, np.int32)
    coef = np.zeros((B, 3, n))
    ctx.check(ctx.lib.uavmp_plan_batch(ctx.h, B, _lib.ptr(sp), _lib.ptr(sv), _lib.ptr(ep), _lib.ptr(ev), order, S,
                                       float(seg_time), C.byref(settings) if settings is not None else None,
                                       _lib.ptr(status), _lib.ptr(solved), _lib.ptr(coef)))
    return dict(search_status=status, qp_solved=solved, coef=coef)


def plan_batch_dev(ctx, B, d_sp, d_sv, d_ep, d_ev, d_status, d_solved, d_coef, order=7, S=8, seg_time=1.0,
                   settings=None):
    """Every argument is a raw device pointer (int); asynchronous on the context's stream."""
    vp = C.c_void_p
    ctx.check(ctx.lib.uavmp_plan_batch_dev(ctx.h, B, vp(d_sp), vp(d_sv), vp(d_ep), vp(d_ev), order, S, float(seg_time),
                                           C.byref(settings) if settings is not None else None, vp(d_status),
                                           vp(d_solved), vp(d_coef)))
