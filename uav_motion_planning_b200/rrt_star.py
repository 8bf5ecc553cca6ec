"""Host-side mirror of path_searching::RRTStar (rrt_star.h:30-101) over the batched CUDA RRT* (SURVEY.md §8(f) row 4).

Same method names and argument meaning as the reference class (`setParam`, `setGridMap`, `init`, `search(start, end, path) -> int`,
`getOptimalPath`, `reset`).  Two things the reference leaves to chance are explicit here (include/uavmp.h): every query carries a
`query_seed` (the reference seeds each sample from std::random_device), and `max_tolerance_time` is `sample_budget`, a number of
drawn samples.  As in the reference, `search` leaves `path` empty (rrt_star.cpp:366,403 clear it) — the result is getOptimalPath().
"""
import numpy as np

from . import _lib

REACH_END = 1      # rrt_star.h:62-65
NO_PATH_FOUND = 2


class RRTStar:
    def __init__(self, ctx=None, device=0):
        self.ctx = ctx if ctx is not None else _lib.Context(device)
        self.lib = self.ctx.lib
        # rrt_star.cpp:7-11 (max_tolerance_time 2.0 s outlasts 100 000 samples on the reference's CPU path: the tree limit ends the search)
        self.params = dict(max_tree_node_num=100000, step_length=0.5, search_radius=0.5, collision_check_resolution=0.05,
                           sample_budget=100000.0, path_cap=4096)
        self._push()
        self.last = {}
        self._optimal_path = np.zeros((0, 3))

    def _push(self):
        p = self.params
        self.ctx.check(self.lib.uavmp_rrt_set_params(self.ctx.h, int(p["max_tree_node_num"]), float(p["step_length"]), float(p["search_radius"]),
                                                     float(p["collision_check_resolution"]), float(p["sample_budget"]), int(p["path_cap"])))

    # -- RRTStar::setParam(ros::NodeHandle&): rrt_star/max_tree_node_num, step_length, search_radius, collision_check_resolution, max_tolerance_time
    def setParam(self, **kw):
        for k, v in kw.items():
            if k not in self.params:
                raise KeyError(k)
            self.params[k] = v
        self._push()

    def setGridMap(self, world):
        occ = np.ascontiguousarray(world.occ, np.int8)
        cloud = np.ascontiguousarray(world.cloud, np.float32)
        origin, msz = _lib.as_f64(world.origin), _lib.as_f64(world.map_size)
        self.ctx.check(self.lib.uavmp_map_set(self.ctx.h, _lib.ptr(occ), *world.dims, _lib.ptr(origin), _lib.ptr(msz), world.resolution,
                                              _lib.ptr(cloud) if len(cloud) else None, len(cloud)))

    def init(self):
        """RRTStar::init (rrt_star.cpp:66-84): node pools and kd-trees are per-warp arenas on the device; nothing to do."""

    def reset(self):
        """RRTStar::reset (rrt_star.cpp:86-101): nothing to do (note: the reference does not clear optimal_path_ either)."""

    def sample_seed(self, query_seed, i):
        return int(self.lib.uavmp_rrt_sample_seed(int(query_seed), int(i)))

    def search_batch(self, start_pt, end_pt, query_seed, want_paths=True):
        sp, ep = (_lib.as_f64(a).reshape(-1, 3) for a in (start_pt, end_pt))
        B = sp.shape[0]
        seeds = np.ascontiguousarray(np.broadcast_to(np.asarray(query_seed, np.uint64), (B,)))
        status, use = np.zeros(B, np.int32), np.zeros(B, np.int32)
        ns, off = np.zeros(B, np.int64), np.zeros(B + 1, np.int64)
        gg, dg = np.zeros(B, np.float64), np.zeros(B, np.uint64)
        total = self.ctx.check(self.lib.uavmp_rrt_search_batch(self.ctx.h, B, _lib.ptr(sp), _lib.ptr(ep), _lib.ptr(seeds), _lib.ptr(status),
                                                               _lib.ptr(use), _lib.ptr(ns), _lib.ptr(gg), _lib.ptr(dg), _lib.ptr(off)))
        paths = None
        if want_paths:
            paths = np.zeros((max(total, 1), 3), np.float64)
            self.ctx.check(self.lib.uavmp_rrt_get_paths(self.ctx.h, _lib.ptr(paths), max(total, 1)))
            paths = paths[:total]
        self.last = dict(status=status, use_node_num=use, n_samples=ns, goal_g_cost=gg, tree_digest=dg, path_offsets=off, paths=paths)
        return self.last

    # -- int search(Eigen::Vector3d start, Eigen::Vector3d end, std::vector<Eigen::Vector3d>& path)  (rrt_star.h:93)
    def search(self, start, end, path, query_seed=0):
        r = self.search_batch(np.asarray(start)[None], np.asarray(end)[None], query_seed)
        if len(r["paths"]):  # the reference keeps a stale optimal_path_ when a search does not rewrite it
            self._optimal_path = r["paths"].copy()
        return int(r["status"][0])

    def getOptimalPath(self):
        return [np.asarray(p) for p in self._optimal_path]
