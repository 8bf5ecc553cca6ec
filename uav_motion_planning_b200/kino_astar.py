"""Host-side mirror of path_searching::KinoAstar (kino_astar.h:117-206) over the batched CUDA search.

Same method names and argument meaning as the reference class; `search` also exists in a batched form because the
B200 path only pays off when thousands of independent queries are in flight.
"""
import ctypes as C

import numpy as np

from . import _lib

REACH_END = 1      # kino_astar.h:155-159
NO_PATH_FOUND = 2


class KinoAstar:
    def __init__(self, ctx=None, device=0):
        self.ctx = ctx if ctx is not None else _lib.Context(device)
        self.lib = self.ctx.lib
        self.params = _lib.KinoParams()
        self.lib.uavmp_kino_params_default(C.byref(self.params))
        self._world = None
        self.last = {}

    # -- KinoAstar::setParam(ros::NodeHandle&) : the 12 ROS parameters, by the same names (kino_astar.cpp:8-19)
    def setParam(self, **kw):
        alias = {"max_acceleration": "max_accelration"}
        for k, v in kw.items():
            k = alias.get(k, k)
            if not hasattr(self.params, k):
                raise KeyError(k)
            setattr(self.params, k, v)
        self.ctx.check(self.lib.uavmp_kino_set_params(self.ctx.h, C.byref(self.params)))

    def setLaunchParams(self):
        """test/launch/test_kino_astar_searching.launch:44-57"""
        self.lib.uavmp_kino_params_launch(C.byref(self.params))
        self.ctx.check(self.lib.uavmp_kino_set_params(self.ctx.h, C.byref(self.params)))

    # -- KinoAstar::setGridMap + localCloudCallback (kino_astar.cpp:38-55,76-79)
    def setGridMap(self, world):
        self._world = world
        occ = np.ascontiguousarray(world.occ, np.int8)
        cloud = np.ascontiguousarray(world.cloud, np.float32)
        origin, msz = _lib.as_f64(world.origin), _lib.as_f64(world.map_size)
        self.ctx.check(self.lib.uavmp_map_set(self.ctx.h, _lib.ptr(occ), *world.dims, _lib.ptr(origin),
                                              _lib.ptr(msz), world.resolution, _lib.ptr(cloud) if len(cloud) else None,
                                              len(cloud)))

    def init(self):
        """KinoAstar::init (kino_astar.cpp:57-74): pools are allocated lazily on the device; nothing to do."""

    def reset(self):
        """KinoAstar::reset (kino_astar.cpp:274-300): per-query state lives in per-CTA arenas; nothing to do."""

    def setTrace(self, pop_cap):
        self.ctx.check(self.lib.uavmp_kino_set_trace(self.ctx.h, pop_cap))

    # -- batched search ------------------------------------------------------------------------------------
    def search_batch(self, start_pt, start_vel, end_pt, end_vel, want_paths=True):
        sp, sv, ep, ev = (_lib.as_f64(a).reshape(-1, 3) for a in (start_pt, start_vel, end_pt, end_vel))
        B = sp.shape[0]
        status = np.zeros(B, np.int32)
        use = np.zeros(B, np.int32)
        off = np.zeros(B + 1, np.int64)
        ph = np.zeros(B, np.uint64)
        npop = np.zeros(B, np.int32)
        total = self.ctx.check(self.lib.uavmp_kino_search_batch(
            self.ctx.h, B, _lib.ptr(sp), _lib.ptr(sv), _lib.ptr(ep), _lib.ptr(ev), _lib.ptr(status), _lib.ptr(use),
            _lib.ptr(off), _lib.ptr(ph), _lib.ptr(npop)))
        paths = None
        if want_paths:
            paths = np.zeros((max(total, 1), 3), np.float64)
            self.ctx.check(self.lib.uavmp_kino_get_paths(self.ctx.h, _lib.ptr(paths), max(total, 1)))
            paths = paths[:total]
        self.last = dict(status=status, use_node_num=use, path_offsets=off, pop_hash=ph, n_pop=npop, paths=paths)
        return self.last

    def pop_trace(self, q, cap):
        tr = np.zeros((cap, 3), np.int32)
        self.ctx.check(self.lib.uavmp_kino_get_trace(self.ctx.h, q, _lib.ptr(tr), cap))
        return tr

    def counters(self):
        c = _lib.KinoCounters()
        self.ctx.check(self.lib.uavmp_kino_get_counters(self.ctx.h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in _lib.KinoCounters._fields_}

    # -- int search(start_pt, start_vel, end_pt, end_vel, std::vector<Vector3d>& path)  (kino_astar.h:197-198)
    def search(self, start_pt, start_vel, end_pt, end_vel, path):
        r = self.search_batch(np.asarray(start_pt)[None], np.asarray(start_vel)[None], np.asarray(end_pt)[None],
                              np.asarray(end_vel)[None])
        path.extend(np.asarray(p) for p in r["paths"])  # the reference only push_back's (caller clears)
        return int(r["status"][0])

    # -- optional in-kernel profile (uavmp_kino_set_profile / uavmp_kino_get_profile) ---------------------------
    def setProfile(self, on=True):
        self.ctx.check(self.lib.uavmp_kino_set_profile(self.ctx.h, int(on)))

    def profile(self, B):
        ph = np.zeros(16, np.uint64)
        qc = np.zeros(17 * B, np.int64)
        grid = C.c_int()
        self.ctx.check(self.lib.uavmp_kino_get_profile(self.ctx.h, _lib.ptr(ph), _lib.ptr(qc), 17 * B, C.byref(grid)))
        qphase = qc[B:].reshape(B, 16).copy()
        qc = qc[:B].copy()
        names = ["pop", "shot_path", "tables_tile_grid", "cloud_ellipsoid", "dedup_probe_heuristic", "node_write", "heap_commit", "setup", "cloud_staging", "n_staged", "commit_replay", "sum_npts", "sum_flagged_prims", "commit_closure_io", "commit_slow_updates", "commit_deferred_writes"]
        return dict(phase_cycles=dict(zip(names, ph.tolist())), query_cycles=qc, query_phase=qphase, names=names, grid=grid.value)

    # -- GridMap::cloudCallback on the device (grid_map.cpp:733-785): only the cloud is uploaded -----------------
    def setGridMapFromCloud(self, world, obstacles_inflation=0.099):
        self._world = world
        cloud = np.ascontiguousarray(world.cloud, np.float32)
        origin, msz = _lib.as_f64(world.origin), _lib.as_f64(world.map_size)
        self.ctx.check(self.lib.uavmp_map_set_from_cloud(self.ctx.h, _lib.ptr(cloud), len(cloud), *world.dims, _lib.ptr(origin),
                                                         _lib.ptr(msz), world.resolution, obstacles_inflation))

    def occupancy(self):
        nx, ny, nz = self._world.dims
        occ = np.zeros(nx * ny * nz, np.int8)
        self.ctx.check(self.lib.uavmp_map_get_occupancy(self.ctx.h, _lib.ptr(occ), occ.size))
        return occ
