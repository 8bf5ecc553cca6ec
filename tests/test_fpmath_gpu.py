"""csrc/fpmath.h is the one implementation of cbrt / acos / cos / integer pow used on BOTH sides (g++ -ffp-contract=off in the
oracle, nvcc -fmad=false on the device).  The bit-exact search parity rests on it, so check it directly: identical bits for a
large sample of inputs, including the ranges the OBVP quartic actually feeds it."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
from uav_motion_planning_b200 import _lib

pytestmark = pytest.mark.gpu


def device_eval(ctx, op, x, n_pow=0):
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros_like(x)
    ctx.check(ctx.lib.uavmp_fpmath_eval(ctx.h, op, n_pow, _lib.ptr(x), _lib.ptr(y), x.size))
    return y


def test_device_and_host_fpmath_are_bit_identical(gpu_ctx):
    rng = np.random.default_rng(123)
    wide = np.concatenate([rng.uniform(-1e6, 1e6, 200000), rng.normal(size=200000), 10.0 ** rng.uniform(-300, 300, 50000),
                           -(10.0 ** rng.uniform(-300, 300, 50000)), [0.0, -0.0, 1.0, -1.0, 8.0, 27.0, 1e-310, np.inf]])
    cases = [(0, wide, 0),                                                     # cbrt
             (1, np.concatenate([rng.uniform(-1, 1, 400000), [-1.0, 1.0, 0.0, 1 - 1e-16, -1 + 1e-16]]), 0),   # acos
             (2, np.concatenate([rng.uniform(0, 2 * np.pi, 400000), rng.uniform(-50, 50, 100000)]), 0)]       # cos
    for n in range(0, 8):
        cases.append((3, rng.uniform(1e-3, 10.0, 100000), n))                  # powi(t, n)
    for op, x, n in cases:
        dev = device_eval(gpu_ctx, op, x, n)
        host = oracle_lib.fp_eval(op, x, n_pow=n)
        assert np.array_equal(dev.view(np.uint64), host.view(np.uint64)), (op, n)
