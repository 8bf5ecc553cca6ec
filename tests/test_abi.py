"""The C-ABI library loads and exports every symbol include/uavmp.h declares (no compute calls: no GPU here)."""
import ctypes as C
import os
import re

import pytest

import uav_motion_planning_b200 as u
from uav_motion_planning_b200 import _lib
from conftest import have_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "uavmp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(uavmp_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_worldgen_library_is_separate_and_complete():
    """The input generator lives in its own host-only library (include/uavmp_worldgen.h): the CPU reference arm of bench.py
    must not need the product library for its inputs."""
    hdr = open(os.path.join(ROOT, "include", "uavmp_worldgen.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decl = sorted(set(re.findall(r"\b(uavmp_[a-z0-9_]+)\s*\(", hdr)))
    assert decl == sorted(_lib.WORLDGEN_SYMBOLS)
    wg = _lib.load_worldgen()
    for s in decl:
        assert hasattr(wg, s), s
    import subprocess
    out = subprocess.run(["ldd", _lib.WORLDGEN_PATH], capture_output=True, text=True).stdout
    assert "cuda" not in out.lower()
    assert not set(decl) & set(_lib.SYMBOLS)


def test_python_parameter_tables_match_the_c_tables():
    lib = u.load()
    for fn, table in ((lib.uavmp_kino_params_launch, _lib.LAUNCH_PARAMS), (lib.uavmp_kino_params_default, _lib.DEFAULT_PARAMS)):
        p = _lib.KinoParams()
        fn(C.byref(p))
        assert {k: getattr(p, k) for k, _ in _lib.KinoParams._fields_} == table


def test_library_exports_every_declared_symbol():
    lib = u.load()
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert b"sm_100a" in lib.uavmp_version()


def test_struct_sizes_match_header():
    # the ctypes mirrors must have the C layout: 2 ints + 10 doubles; 7 doubles + 5 ints (+pad) + 1 double
    assert C.sizeof(_lib.KinoParams) == 8 + 10 * 8
    assert C.sizeof(_lib.OsqpSettings) == 7 * 8 + 5 * 4 + 4 + 8
    assert C.sizeof(_lib.KinoCounters) == 8 * 8


def test_parameter_tables_match_reference_sources():
    p = _lib.KinoParams()
    u.load().uavmp_kino_params_launch(C.byref(p))     # test_kino_astar_searching.launch:44-57
    assert (p.rou_time, p.lambda_heu, p.allocated_node_num, p.goal_tolerance, p.time_step_size, p.max_velocity,
            p.max_accelration, p.acc_resolution, p.sample_tau, p.robot_r, p.robot_h) == \
        (50.0, 3.0, 100000, 2.0, 0.075, 7.0, 10.0, 4.0, 0.3, 0.4, 0.1)
    u.load().uavmp_kino_params_default(C.byref(p))    # kino_astar.cpp:8-19
    assert (p.rou_time, p.lambda_heu, p.time_step_size, p.max_velocity, p.max_accelration, p.acc_resolution,
            p.sample_tau, p.robot_r) == (1.0, 2.0, 0.1, 5.0, 7.0, 2.0, 0.5, 0.2)
    s = _lib.OsqpSettings()
    u.load().uavmp_osqp_settings_default(C.byref(s))  # osqp_api_constants.h:96-153 + minimum_control.cpp:160-162
    assert (s.rho, s.sigma, s.alpha, s.eps_abs, s.eps_rel, s.eps_prim_inf, s.eps_dual_inf, s.max_iter,
            s.check_termination, s.scaling, s.adaptive_rho, s.adaptive_rho_tolerance) == \
        (0.1, 1e-6, 1.6, 1e-3, 1e-3, 1e-3, 1e-4, 1000, 25, 10, 1, 5.0)


@pytest.mark.skipif(have_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(u.UavmpError):
        u.Context(0)
