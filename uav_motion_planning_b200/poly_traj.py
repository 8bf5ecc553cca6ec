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
