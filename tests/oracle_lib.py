"""ctypes binding of oracle/liboracle.so — the CPU checker.  Imported by tests/, smoke() and bench.py's CPU legs only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")


class KinoParams(C.Structure):
    _fields_ = [("allocated_node_num", C.c_int), ("collision_check_type", C.c_int)] + \
        [(n, C.c_double) for n in ("rou_time", "lambda_heu", "goal_tolerance", "time_step_size", "max_velocity",
                                   "max_acceleration", "acc_resolution", "sample_tau", "robot_r", "robot_h")] + \
        [("libm_mode", C.c_int)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("n_pop", "n_occ_lookup", "n_cloud_pts_tested", "n_hash_probe", "n_insert",
                                            "n_update", "n_heuristic", "n_shot", "heap_len_sum")]


class Result(C.Structure):
    _fields_ = [("status", C.c_int), ("use_node_num", C.c_int), ("n_pop", C.c_int), ("n_path", C.c_int),
                ("n_path_nodes", C.c_int), ("shot_duration", C.c_double), ("pop_hash", C.c_uint64),
                ("counters", Counters)]


_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = C.CDLL(LIB)
        lib.oracle_kino_create.restype = C.c_void_p
        lib.oracle_kino_create.argtypes = [C.POINTER(KinoParams), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_double, C.c_void_p, C.c_int]
        lib.oracle_kino_destroy.argtypes = [C.c_void_p]
        lib.oracle_kino_search.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.POINTER(Result), C.c_void_p, C.c_int,
                                                                            C.c_void_p, C.c_int]
        lib.oracle_fp_eval_n.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
        _lib = lib
    return _lib


def params_from(p, libm_mode=0):
    """uav_motion_planning_b200.KinoParams (or any object with the same fields) -> oracle params."""
    return KinoParams(p.allocated_node_num, p.collision_check_type, p.rou_time, p.lambda_heu, p.goal_tolerance,
                      p.time_step_size, p.max_velocity, p.max_accelration, p.acc_resolution, p.sample_tau, p.robot_r,
                      p.robot_h, libm_mode)


class KinoOracle:
    def __init__(self, world, params, libm_mode=0):
        self.lib = load()
        kp = params_from(params, libm_mode)
        self.occ = np.ascontiguousarray(world.occ, np.int8)
        self.cloud = np.ascontiguousarray(world.cloud, np.float32)
        origin = np.ascontiguousarray(world.origin, np.float64)
        msz = np.ascontiguousarray(world.map_size, np.float64)
        self.h = self.lib.oracle_kino_create(C.byref(kp), self.occ.ctypes.data, *world.dims, origin.ctypes.data,
                                             msz.ctypes.data, world.resolution, self.cloud.ctypes.data,
                                             len(self.cloud))

    def search(self, sp, sv, ep, ev, path_cap=4096, pop_cap=0):
        sp, sv, ep, ev = (np.ascontiguousarray(a, np.float64) for a in (sp, sv, ep, ev))
        res = Result()
        path = np.zeros((path_cap, 3))
        trace = np.zeros((max(pop_cap, 1), 3), np.int32)
        self.lib.oracle_kino_search(self.h, sp.ctypes.data, sv.ctypes.data, ep.ctypes.data, ev.ctypes.data,
                                    C.byref(res), path.ctypes.data, path_cap, trace.ctypes.data if pop_cap else None,
                                    pop_cap)
        c = res.counters
        return dict(status=res.status, use_node_num=res.use_node_num, n_pop=res.n_pop, pop_hash=res.pop_hash,
                    path=path[:min(res.n_path, path_cap)].copy(), n_path=res.n_path,
                    trace=trace[:min(res.n_pop, pop_cap)].copy(),
                    counters={k: getattr(c, k) for k, _ in Counters._fields_})

    def close(self):
        if self.h:
            self.lib.oracle_kino_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fp_eval(op, x, n_pow=0):
    lib = load()
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros_like(x)
    lib.oracle_fp_eval_n(op, x.ctypes.data, n_pow, y.ctypes.data, x.size)
    return y
