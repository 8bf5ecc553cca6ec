"""Synthetic worlds (host utility, libuavmp_worldgen.so): the reference simulator's obstacle generators restated in
csrc/worldgen.cpp.  Nothing here touches the CUDA library.

random_forest.cpp:55-155 / :286-306 -> cloud; pointcloud_render_node.cpp:84-86 -> 0.1 m voxel centroids;
grid_map.cpp:733-785 -> inflated int8 grid (x-major, z-fastest, grid_map.h:257-260).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass
class World:
    occ: np.ndarray        # int8 [nx*ny*nz]
    dims: tuple            # (nx, ny, nz)
    origin: np.ndarray     # f64[3]   mp_.map_origin_  (grid_map.cpp:53)
    map_size: np.ndarray   # f64[3]
    resolution: float
    cloud: np.ndarray      # float32 [n, 3]

    @property
    def occ3(self):
        return self.occ.reshape(self.dims)


def make_world(x_size, y_size, z_size, seed=1, map_type=0, resolution=0.1, ground_height=0.0,
               obstacles_inflation=0.099, **overrides):
    lib = _lib.load_worldgen()
    mp = _lib.MapgenParams()
    lib.uavmp_mapgen_params_default(C.byref(mp), float(x_size), float(y_size), int(seed))
    mp.map_type = map_type
    for k, v in overrides.items():
        setattr(mp, k, v)
    n = lib.uavmp_mapgen_cloud(C.byref(mp), None, 0)
    if n < 0:
        raise ValueError(f"mapgen failed ({n})")
    cloud = np.zeros((max(n, 1), 3), np.float32)
    lib.uavmp_mapgen_cloud(C.byref(mp), _lib.ptr(cloud), n)
    cloud = cloud[:n]
    # grid_map.cpp:53,69-70
    origin = np.array([-x_size / 2.0, -y_size / 2.0, ground_height], np.float64)
    map_size = np.array([x_size, y_size, z_size], np.float64)
    dims = tuple(int(np.ceil(map_size[i] / resolution)) for i in range(3))
    occ = np.zeros(dims[0] * dims[1] * dims[2], np.int8)
    rc = lib.uavmp_grid_inflate_host(_lib.ptr(cloud), n, _lib.ptr(origin), _lib.ptr(map_size), resolution,
                                     obstacles_inflation, _lib.ptr(occ), *dims)
    if rc != 0:
        raise ValueError("grid inflation failed")
    return World(occ, dims, origin, map_size, resolution, cloud)


def sample_queries(world, n, seed=2, min_dist=10.0, z_lo=0.5, z_hi=2.5, margin=0.5):
    """Start/goal pairs uniform over free (non-inflated) voxels, zero start/goal velocity."""
    rng = np.random.default_rng(seed)
    occ3 = world.occ3
    lo = world.origin + margin
    hi = world.origin + world.map_size - margin
    z_hi = min(z_hi, hi[2])

    def free_pts(m):
        out = np.zeros((0, 3))
        while len(out) < m:
            p = np.stack([rng.uniform(lo[0], hi[0], 2 * m), rng.uniform(lo[1], hi[1], 2 * m),
                          rng.uniform(z_lo, z_hi, 2 * m)], 1)
            i = np.floor((p - world.origin) / world.resolution).astype(int)
            ok = occ3[i[:, 0], i[:, 1], i[:, 2]] == 0
            out = np.concatenate([out, p[ok]])
        return out[:m]

    s_all, g_all = np.zeros((0, 3)), np.zeros((0, 3))
    while len(s_all) < n:
        s, g = free_pts(n), free_pts(n)
        ok = np.linalg.norm(s - g, axis=1) >= min_dist
        s_all = np.concatenate([s_all, s[ok]])
        g_all = np.concatenate([g_all, g[ok]])
    z = np.zeros((n, 3))
    return s_all[:n].copy(), z.copy(), g_all[:n].copy(), z.copy()
