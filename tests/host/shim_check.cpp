// compile-and-link check of the C++ shims against the Eigen stand-in (tests/test_shims.py); calls nothing that needs a GPU
#include <uavmp/a_star.hpp>
#include <uavmp/kino_astar.hpp>
#include <uavmp/minimum_control.hpp>
#include <uavmp/rrt_star.hpp>
extern "C" int shim_check() {
  uavmp_kino_params p;
  uavmp_kino_params_launch(&p);
  try {
    uavmp::KinoAstar ka(0);  // throws without a CUDA device: "no CPU fallback"
    ka.setParam(p);
    uavmp::MinimumControl mc(ka.context(), 5);
    (void)mc;
    uavmp::Astar as(ka.context());
    as.setParam(1.0, 100000);
    uavmp::RRTStar rs(ka.context());
    rs.setParam();
    return 1;  // a GPU is present
  } catch (const std::exception&) {
    return 0;
  }
}
