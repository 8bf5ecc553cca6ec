"""include/uavmp/*.hpp (the C++ shims with the reference's class signatures) must at least compile and link against libuavmp.so.
Eigen is not installed here, so a minimal stand-in header (tests/host/eigen_stub) provides the few members the shims touch."""
import ctypes as C
import os
import subprocess

from conftest import have_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shims_compile_and_link(tmp_path):
    out = tmp_path / "libshim_check.so"
    pkg = os.path.join(ROOT, "uav_motion_planning_b200")
    subprocess.run(["g++", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "tests", "host", "eigen_stub"), os.path.join(ROOT, "tests", "host", "shim_check.cpp"),
                    "-L" + pkg, "-l:libuavmp.so", "-Wl,-rpath," + pkg, "-o", str(out)], check=True)
    lib = C.CDLL(str(out))
    assert lib.shim_check() == (1 if have_gpu() else 0)   # without a GPU the shim's constructor throws: no CPU fallback
