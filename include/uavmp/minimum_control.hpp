// include/uavmp/minimum_control.hpp — C++ shim with the reference's class interface over the C-ABI (include/uavmp.h).
//
// Drop-in for traj_optimization::MinimumControl (reference: src/planner/traj_optimization/include/traj_optimization/
// minimum_control.h:10-49): bool solve(VectorXd& pos_1d, Vector2d& bound_vel, Vector2d& bound_acc, VectorXd& time_vec), one axis
// per call, true only when OSQP reports SOLVED (3rd/osqp-eigen/src/Solver.cpp:181-187); getCoef1d() returns the 6S coefficients,
// segment-major, ascending power, local time; a failed solve leaves the previous coefficients in place (minimum_control.cpp:182).
// Compiled only where Eigen exists.
#pragma once
#include <Eigen/Eigen>
#include <iostream>
#include <stdexcept>

#include "../uavmp.h"

namespace uavmp {

class MinimumControl {
 public:
  explicit MinimumControl(uavmp_ctx* ctx, int order = 5) : ctx_(ctx), order_(order) { uavmp_osqp_settings_default(&settings_); }

  bool solve(Eigen::VectorXd& pos_1d, Eigen::Vector2d& bound_vel, Eigen::Vector2d& bound_acc, Eigen::VectorXd& time_vec) {
    const int S = (int)time_vec.size();
    Eigen::VectorXd coef((order_ + 1) * S);
    double bj[2] = {0.0, 0.0};
    int solved = 0;
    int rc = uavmp_minctrl_solve_batch(ctx_, order_, S, 1, pos_1d.data(), bound_vel.data(), bound_acc.data(),
                                       order_ == 7 ? bj : nullptr, time_vec.data(), &settings_, coef.data(), &solved, nullptr,
                                       nullptr);
    if (rc < 0) throw std::runtime_error(uavmp_last_error(ctx_));
    if (!solved) { std::cout << "solver solve failed!" << std::endl; return false; }
    coef_1d_ = coef;
    return true;
  }
  Eigen::VectorXd getCoef1d() { return coef_1d_; }
  void reset() { coef_1d_.setZero(); }
  uavmp_osqp_settings& settings() { return settings_; }

 private:
  uavmp_ctx* ctx_;
  int order_;
  uavmp_osqp_settings settings_;
  Eigen::VectorXd coef_1d_;
};

}  // namespace uavmp
