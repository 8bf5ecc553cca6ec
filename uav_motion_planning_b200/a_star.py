"""Host-side mirror of path_searching::Astar (a_star.h:106-152) over the batched CUDA grid A* (SURVEY.md §8(f) row 4).

Same method names and argument meaning as the reference class (`setParam`, `setGridMap`, `init`, `search(start_pt, end_pt, path)
-> int`, `reset`); `search_batch` is the form that pays off on the GPU.  The map is the context's (uavmp_map_set), shared with
KinoAstar.
"""
import numpy as np

from . import _lib

REACH_END = 1      # a_star.h:122-126
NO_PATH_FOUND = 2


class Astar:
    def __init__(self, ctx=None, device=0):
        self.ctx = ctx if ctx is not None else _lib.Context(device)
        self.lib = self.ctx.lib
        self.lambda_heu, self.allocated_node_num, self.path_cap = 1.0, 100000, 4096  # a_star.cpp:8-10
        self._push()
        self.last = {}

    def _push(self):
        self.ctx.check(self.lib.uavmp_astar_set_params(self.ctx.h, float(self.lambda_heu), int(self.allocated_node_num), int(self.path_cap)))

    # -- Astar::setParam(ros::NodeHandle&): astar/lambda_heu, astar/allocated_node_num (astar/resolution is overwritten by the map's)
    def setParam(self, lambda_heu=None, allocated_node_num=None, path_cap=None):
        if lambda_heu is not None:
            self.lambda_heu = lambda_heu
        if allocated_node_num is not None:
            self.allocated_node_num = allocated_node_num
        if path_cap is not None:
            self.path_cap = path_cap
        self._push()

    def setGridMap(self, world):
        occ = np.ascontiguousarray(world.occ, np.int8)
        cloud = np.ascontiguousarray(world.cloud, np.float32)
        origin, msz = _lib.as_f64(world.origin), _lib.as_f64(world.map_size)
        self.ctx.check(self.lib.uavmp_map_set(self.ctx.h, _lib.ptr(occ), *world.dims, _lib.ptr(origin), _lib.ptr(msz), world.resolution,
                                              _lib.ptr(cloud) if len(cloud) else None, len(cloud)))

    def init(self):
        """Astar::init (a_star.cpp:13-38): the node pools are allocated lazily on the device; nothing to do."""

    def reset(self):
        """Astar::reset (a_star.cpp:196-213): per-query state lives in per-warp arenas; nothing to do."""

    def search_batch(self, start_pt, end_pt, want_paths=True):
        sp, ep = (_lib.as_f64(a).reshape(-1, 3) for a in (start_pt, end_pt))
        B = sp.shape[0]
        status, use, npop = np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32)
        off, ph = np.zeros(B + 1, np.int64), np.zeros(B, np.uint64)
        total = self.ctx.check(self.lib.uavmp_astar_search_batch(self.ctx.h, B, _lib.ptr(sp), _lib.ptr(ep), _lib.ptr(status), _lib.ptr(use),
                                                                 _lib.ptr(off), _lib.ptr(ph), _lib.ptr(npop)))
        paths = None
        if want_paths:
            paths = np.zeros((max(total, 1), 3), np.float64)
            self.ctx.check(self.lib.uavmp_astar_get_paths(self.ctx.h, _lib.ptr(paths), max(total, 1)))
            paths = paths[:total]
        self.last = dict(status=status, use_node_num=use, path_offsets=off, pop_hash=ph, n_pop=npop, paths=paths)
        return self.last

    # -- int search(Eigen::Vector3d start_pt, Eigen::Vector3d end_pt, std::vector<Eigen::Vector3d>& path)  (a_star.h:147)
    def search(self, start_pt, end_pt, path):
        r = self.search_batch(np.asarray(start_pt)[None], np.asarray(end_pt)[None])
        path.extend(np.asarray(p) for p in r["paths"])
        return int(r["status"][0])
