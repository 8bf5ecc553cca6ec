"""Host-side mirror of PolyTraj (traj_utils/poly_traj.hpp:1-208) over uavmp_polytraj_eval_batch: the consumer of the QP's
coefficients.  `evaluatePos/Vel/Acc(t)` keep the reference names; `evaluate_batch` is the B200 entry point."""
import numpy as np

from . import _lib


class PolyTraj:
    def __init__(self, ctx=None, device=0):
        self.ctx = ctx if ctx is not None else _lib.Context(device)
        self.coef = None    # [3, S, order+1]
        self.times = None   # [S]

    # void addSegment(cx, cy, cz, t) + init()
    def addSegment(self, cx, cy, cz, t):
        seg = np.stack([np.asarray(cx, float), np.asarray(cy, float), np.asarray(cz, float)])[:, None, :]
        self.coef = seg if self.coef is None else np.concatenate([self.coef, seg], axis=1)
        self.times = np.array([t], float) if self.times is None else np.append(self.times, float(t))

    def init(self):
        self.total_time = float(np.sum(self.times))

    def getTotalTIme(self):  # sic
        return float(np.sum(self.times))

    def _eval(self, t, deriv):
        return evaluate_batch(self.ctx, self.coef[None], self.times[None], np.atleast_1d(np.asarray(t, float)), deriv)[0]

    def evaluatePos(self, t):
        return self._eval(t, 0)[0]

    def evaluateVel(self, t):
        return self._eval(t, 1)[0]

    def evaluateAcc(self, t):
        return self._eval(t, 2)[0]


def evaluate_batch(ctx, coef, times, t, deriv=0):
    """coef [B, 3, S, order+1] (uavmp_plan_batch's layout reshaped), times [B, S], t [n_t] -> [B, n_t, 3]."""
    coef = _lib.as_f64(coef)
    B, _, S, nc = coef.shape
    times = _lib.as_f64(times).reshape(B, S)
    t = _lib.as_f64(t).reshape(-1)
    out = np.zeros((B, len(t), 3))
    ctx.check(ctx.lib.uavmp_polytraj_eval_batch(ctx.h, B, nc - 1, S, _lib.ptr(coef), _lib.ptr(times), len(t), _lib.ptr(t), deriv,
                                                _lib.ptr(out)))
    return out


# ---- quadrotor_msgs/PolynomialTrajectory: the wire layout between the optimiser and traj_server (host-side, no GPU) -------------------
ACTION_ADD, ACTION_ABORT, ACTION_WARN_START, ACTION_WARN_FINAL, ACTION_WARN_IMPOSSIBLE = 1, 2, 3, 4, 5  # PolynomialTrajectory.msg:7-11


def to_polynomial_trajectory(coef, times, trajectory_id=1, stamp=0.0, start_yaw=0.0, final_yaw=0.0, action=ACTION_ADD):
    """One plan of uavmp_plan_batch / MinimumControl (coef [3, n] or [3, S, order+1], segment-major, ascending power, local time) as the
    fields of simulator/utils/quadrotor_msgs/msg/PolynomialTrajectory.msg: coef_x[i * (num_order + 1) + j] is the t^j coefficient of
    segment i — exactly the indexing poly_traj_server.cpp:68-78 (trajCallback) reads back."""
    times = np.asarray(times, float).reshape(-1)
    S = len(times)
    c = np.asarray(coef, float).reshape(3, S, -1)
    order = c.shape[2] - 1
    return dict(header=dict(stamp=float(stamp), frame_id="world"), trajectory_id=int(trajectory_id), action=int(action), num_order=order,
                num_segment=S, start_yaw=float(start_yaw), final_yaw=float(final_yaw), coef_x=c[0].reshape(-1).tolist(),
                coef_y=c[1].reshape(-1).tolist(), coef_z=c[2].reshape(-1).tolist(), time=times.tolist(), mag_coeff=1.0,
                order=[order] * S, debug_info="")


def from_polynomial_trajectory(msg, ctx=None):
    """trajCallback (poly_traj_server.cpp:57-81): per segment i the num_order + 1 coefficients of each axis and time[i] go to
    PolyTraj::addSegment, then init().  Returns (coef [3, S, order+1], times [S]) and, when a context is given, the PolyTraj mirror."""
    n, S = msg["num_order"] + 1, msg["num_segment"]
    coef = np.stack([np.asarray(msg[k], float)[:S * n].reshape(S, n) for k in ("coef_x", "coef_y", "coef_z")])
    times = np.asarray(msg["time"], float)[:S]
    if ctx is None:
        return coef, times
    pt = PolyTraj(ctx)
    for i in range(S):
        pt.addSegment(coef[0, i], coef[1, i], coef[2, i], times[i])
    pt.init()
    return coef, times, pt
