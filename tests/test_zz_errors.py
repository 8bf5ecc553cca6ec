"""Error behaviour of the C-ABI on the GPU: bad arguments and out-of-contract configurations are refused with a message
(never computed differently), mirroring where the reference would misbehave (e.g. a type-2 search without a cloud crashes in the
KD-tree)."""
import ctypes as C

import numpy as np
import pytest

import uav_motion_planning_b200 as u
from uav_motion_planning_b200 import _lib
from uav_motion_planning_b200.minimum_control import MinimumControl, default_settings

pytestmark = pytest.mark.gpu


def test_search_before_map_is_refused():
    ctx = u.Context(0)
    ka = u.KinoAstar(ctx)
    z = np.zeros((1, 3))
    with pytest.raises(u.UavmpError, match="uavmp_map_set"):
        ka.search_batch(z, z, z + 1, z)
    ctx.close()


def test_bad_parameters_are_refused(gpu_ctx):
    ka = u.KinoAstar(gpu_ctx)
    for kw, msg in [(dict(collision_check_type=3), "collision_check_type"), (dict(acc_resolution=5.0), "lattice|primitives"),
                    (dict(time_step_size=0.001), "checkpoints"), (dict(allocated_node_num=1), "allocated_node_num"),
                    (dict(robot_h=0.6, robot_r=0.4), "robot_h")]:
        with pytest.raises(u.UavmpError, match=msg):
            ka.setParam(**kw)
        ka = u.KinoAstar(gpu_ctx)


def test_type2_without_cloud_is_refused(gpu_ctx):
    world = u.make_world(20, 20, 5, seed=1)
    ka = u.KinoAstar(gpu_ctx)
    ka.setLaunchParams()
    ka.setParam(collision_check_type=2)
    empty = u.mapgen.World(world.occ, world.dims, world.origin, world.map_size, world.resolution, np.zeros((0, 3), np.float32))
    ka.setGridMap(empty)
    z = np.zeros((1, 3))
    with pytest.raises(u.UavmpError, match="cloud"):
        ka.search_batch(z + 1, z, z + 2, z)
    ka.setParam(collision_check_type=1)          # leave the shared context usable
    ka.setGridMap(world)


def test_qp_argument_checks(gpu_ctx):
    mc = MinimumControl(gpu_ctx, order=6)
    with pytest.raises(u.UavmpError, match="order"):
        mc.solve_batch(np.zeros((1, 4)), np.zeros((1, 2)), np.zeros((1, 2)), np.ones((1, 3)))
    mc = MinimumControl(gpu_ctx, order=5)
    with pytest.raises(u.UavmpError, match="settings"):
        mc.solve_batch(np.zeros((1, 4)), np.zeros((1, 2)), np.zeros((1, 2)), np.ones((1, 3)), settings=default_settings(max_iter=0))


def test_start_in_obstacle_or_outside_map(gpu_ctx):
    """A start inside an inflated voxel / outside the map cannot expand: every primitive fails at t = 0 -> NO_PATH_FOUND
    after one pop (the oracle agrees)."""
    import oracle_lib
    world = u.make_world(20, 20, 5, seed=1)
    ka = u.KinoAstar(gpu_ctx)
    ka.setLaunchParams()
    ka.setGridMap(world)
    occ3 = world.occ3
    ix, iy, iz = np.argwhere(occ3 == 1)[1000]
    inside = world.origin + (np.array([ix, iy, iz]) + 0.5) * world.resolution
    sp = np.stack([inside, np.array([50.0, 0.0, 1.0])])
    z = np.zeros((2, 3))
    ep = np.array([[5.0, 5.0, 1.0], [5.0, 5.0, 1.0]])
    got = ka.search_batch(sp, z, ep, z)
    orc = oracle_lib.KinoOracle(world, ka.params)
    for q in range(2):
        ref = orc.search(sp[q], z[q], ep[q], z[q])
        assert (ref["status"], ref["n_pop"], ref["use_node_num"]) == (got["status"][q], got["n_pop"][q], got["use_node_num"][q])
        assert got["status"][q] == 2
