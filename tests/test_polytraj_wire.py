"""quadrotor_msgs/PolynomialTrajectory wire layout: coef_x[i * (num_order + 1) + j] as poly_traj_server.cpp:68-78
reads it; host-side only."""
import numpy as np

from uav_motion_planning_b200 import poly_traj


def test_layout_and_round_trip():
    rng = np.random.default_rng(0)
    for order, S in ((5, 3), (7, 8)):
        coef = rng.normal(size=(3, S * (order + 1)))       # MinimumControl::getCoef1d layout per axis
        T = rng.uniform(0.5, 2.0, S)
        msg = poly_traj.to_polynomial_trajectory(coef, T, trajectory_id=4, stamp=12.5)
        assert (msg["num_order"], msg["num_segment"], msg["trajectory_id"], msg["action"]) == (order, S, 4, poly_traj.ACTION_ADD)
        assert len(msg["coef_x"]) == S * (order + 1) and msg["order"] == [order] * S
        for i in range(S):
            for j in range(order + 1):
                assert msg["coef_y"][i * (order + 1) + j] == coef[1, i * (order + 1) + j]
        c2, t2 = poly_traj.from_polynomial_trajectory(msg)
        assert np.array_equal(c2.reshape(3, -1), coef) and np.array_equal(t2, T)
        # the polynomial the server evaluates (poly_traj.hpp:74-105): segment i at local time tau = sum_j c_ij tau^j
        tau = 0.3 * T[1]
        p = sum(c2[0, 1, j] * tau ** j for j in range(order + 1))
        assert abs(p - np.polyval(coef[0, (order + 1):2 * (order + 1)][::-1], tau)) < 1e-12
