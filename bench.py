#!/usr/bin/env python
"""bench.py — plans/sec of the batched kino-A* + minimum-snap QP hot path (BASELINE.json metric, configs[1]).

  python bench.py --gpus N --steps K --warmup W            one rank per GPU (under torchrun for N > 1)
  python bench.py --impl reference [...]                   the CPU path (oracle restatement of KinoAstar::search + the
                                                           reference's own OSQP C code), all host threads, rank 0 only

A "step" = one pass of the hot path over one batch of B = 4096 synthetic start->goal queries on the 50 x 50 x 10 m
random-obstacle map @ 0.1 m: KinoAstar::search (launch-file parameters, collision_check_type 1) -> S+1 waypoints ->
three 8-segment 7th-order minimum-snap QPs (x, y, z) per query.  Every step uses a different seeded batch and the L2
is flushed (256 MiB write) between steps.  Weak scaling: every rank processes its own B queries per step, the map is
replicated, and (N > 1) the solved trajectories are all-gathered with NCCL inside the timed region.

Keys beyond the base contract: `roofline` (the search kernel, algorithmic bytes of SURVEY.md §8(d) / CUDA-event
duration), `cpu_baseline` (bounded sample of the same workload on the host cores), `e2e` (host buffers through
uavmp_plan_batch, copies inside the timed region), `clocks`, `gpu_launches`.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "plans/sec (kino-A* + min-snap QP) batched queries"
ORDER, SEG, SEG_TIME = 7, 8, 1.0
MAP = (50.0, 50.0, 10.0)
NODE_BYTES, HASH_SLOT, HEAP_SLOT = 72, 16, 4  # DESIGN.md "algorithmic bytes"


def workload_config(B, n_gpus):
    return {"workload": "configs[1]: batch 4096 queries, 50x50x10 m random map @0.1 m, kino-A* + 8-seg 7th-order "
                        "min-snap, per GPU", "batch_per_gpu": B, "global_batch": B * n_gpus, "map": "500x500x100 int8, "
            "random_forest seed 1", "kino": "launch-file params, collision_check_type 1 (grid + ellipsoid)",
            "qp": "order 7, S 8, T_i 1.0, OSQP eps 1e-3, 3 axes per plan; QP kernel (one warp per problem) overlapped with the search "
            "kernel on a second stream, per-query completion flags", "l2": "flushed between steps (256 MiB write) "
            "and a different query batch every step", "parallelism": f"queries sharded x{n_gpus}, map replicated"}


def make_batches(world, B, n, rank):
    import uav_motion_planning_b200 as u
    return [u.sample_queries(world, B, seed=1000 * rank + 11 + i) for i in range(n)]


def search_bytes(c):
    """SURVEY.md §8(d) bytes_a from the search counters (summed over the batch)."""
    return (1 * c["n_occ_lookup"] + 12 * c["n_cloud_pts_tested"] + HASH_SLOT * c["n_hash_probe"] +
            (NODE_BYTES + HASH_SLOT + HEAP_SLOT) * c["n_insert"] + NODE_BYTES * c["n_update"] +
            (HEAP_SLOT + NODE_BYTES) * c["n_pop"])


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for nme, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        # median over the upper half = the samples taken under load
        under = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": under[len(under) // 2] if under else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU path (oracle): used ONLY by the cpu_baseline leg and by --impl reference
# ---------------------------------------------------------------------------------------------------------------
def cpu_plans(world, params, queries, idx, threads):
    """Run the CPU pipeline for queries[idx] on `threads` host threads; returns (seconds, n_reached, n_solved)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from pipeline_ref import plan_one
    sp, sv, ep, ev = queries
    oracles = [oracle_lib.KinoOracle(world, params) for _ in range(threads)]
    oracle_lib.minctrl_solve(ORDER, SEG, np.arange(SEG + 1.0), [0, 0], [0, 0], np.ones(SEG), bound_jerk=[0, 0])  # dlopen
    nxt, lock, res = [0], threading.Lock(), []

    def work(o):
        while True:
            with lock:
                k = nxt[0]
                nxt[0] += 1
            if k >= len(idx):
                return
            q = idx[k]
            res.append(plan_one(o, sp[q], sv[q], ep[q], ev[q], ORDER, SEG, SEG_TIME)[:2])

    t0 = time.perf_counter()
    ts = [threading.Thread(target=work, args=(o,)) for o in oracles]
    [t.start() for t in ts]
    [t.join() for t in ts]
    dt = time.perf_counter() - t0
    for o in oracles:
        o.close()
    return dt, sum(1 for s, _ in res if s == 1), sum(1 for _, k in res if k)


def kind_of_cpu_path():
    return ("port", "search: oracle/kino_ref.cpp (restatement; the reference's kino_astar.cpp needs ROS/Eigen/PCL); "
            "QP: the reference's vendored OSQP C core compiled unmodified (oracle/_ref) + restated QDLDL")


def run_reference(args, rank, world_size):
    if rank != 0:
        return
    import uav_motion_planning_b200 as u
    from uav_motion_planning_b200 import _lib
    world = u.make_world(*MAP, seed=1)  # host-side input generator only
    params = _lib.KinoParams()
    u.load().uavmp_kino_params_launch(C.byref(params))
    threads = os.cpu_count() or 1
    n_s = args.cpu_sample or max(64, 4 * threads)
    batches = make_batches(world, args.batch, args.steps + args.warmup, 0)
    for i in range(args.warmup):
        cpu_plans(world, params, batches[i], list(range(min(n_s, 2 * threads))), threads)
    tot, cnt, reached = 0.0, 0, 0
    for i in range(args.steps):
        dt, nr, _ = cpu_plans(world, params, batches[args.warmup + i], list(range(n_s)), threads)
        tot += dt; cnt += n_s; reached += nr
    val = cnt / tot
    kind, how = kind_of_cpu_path()
    sample = f"first {n_s} queries of each step's 4096-query batch ({args.steps} steps), {threads} threads; {how}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "plans/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.batch, args.gpus),
        "cpu_baseline": {"value": val, "unit": "plans/s", "cores": threads, "kind": kind, "sample": sample,
                         "reach_end_frac": reached / cnt},
        "e2e": {"value": val, "unit": "plans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------------------
def run_gpu(args, rank, world_size, local_rank):
    import torch
    import torch.distributed as dist
    import uav_motion_planning_b200 as u
    from uav_motion_planning_b200.planner import plan_batch_dev

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference "
                         "for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    ctx = u.Context(local_rank)
    lib = ctx.lib
    B, K, W = args.batch, args.steps, args.warmup
    n = (ORDER + 1) * SEG
    world = u.make_world(*MAP, seed=1)
    ka = u.KinoAstar(ctx)
    ka.setLaunchParams()
    ka.setGridMap(world)
    batches = make_batches(world, B, K + W, rank)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(ext):
        d_in = [[torch.from_numpy(a).to(dev) for a in bt] for bt in batches]  # resident in HBM before timing
        d_status = torch.zeros(B, dtype=torch.int32, device=dev)
        d_solved = torch.zeros(B, dtype=torch.int32, device=dev)
        d_coef = torch.zeros(B, 3 * n, dtype=torch.float64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        if world_size > 1:
            g_coef = torch.empty(world_size * B, 3 * n, dtype=torch.float64, device=dev)
            g_stat = torch.empty(world_size * B, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()

        def step_dev(i):
            sp, sv, ep, ev = d_in[i]
            flush.fill_(i & 0xff)  # evict the previous step's working set from L2
            plan_batch_dev(ctx, B, sp.data_ptr(), sv.data_ptr(), ep.data_ptr(), ev.data_ptr(), d_status.data_ptr(),
                           d_solved.data_ptr(), d_coef.data_ptr(), ORDER, SEG, SEG_TIME)
            if world_size > 1:  # "all-gather of solved trajectories only"
                dist.all_gather_into_tensor(g_coef, d_coef)
                dist.all_gather_into_tensor(g_stat, d_solved)

        for i in range(W):
            step_dev(i)
        barrier()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        search_ms, qp_ms, bytes_a, pops, launches = [], [], [], [], 0
        e0.record(ext)
        for i in range(W, W + K):
            step_dev(i)
            t = ctx.timings()  # resolves the library's own CUDA events (search / QP) for this step
            search_ms.append(t["search_ms"]); qp_ms.append(t["qp_ms"])
            launches += t["search_launches"] + t["qp_launches"] + t["aux_launches"]
            c = ka.counters()
            bytes_a.append(search_bytes(c)); pops.append(c["n_pop"])
        e1.record(ext)
        barrier()
        ms = e0.elapsed_time(e1)
        reached = int((d_status == 1).sum().item())
        solved = int(d_solved.sum().item())

        # ---- e2e: host (pinned) buffers through uavmp_plan_batch, copies inside the timed region -------------
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_in = [[pin(a) for a in bt] for bt in batches]
        h_status = torch.zeros(B, dtype=torch.int32).pin_memory()
        h_solved = torch.zeros(B, dtype=torch.int32).pin_memory()
        h_coef = torch.zeros(B, 3 * n, dtype=torch.float64).pin_memory()
        vp = C.c_void_p

        def step_host(i):
            sp, sv, ep, ev = h_in[i]
            flush.fill_(i & 0xff)
            ctx.check(lib.uavmp_plan_batch(ctx.h, B, vp(sp.data_ptr()), vp(sv.data_ptr()), vp(ep.data_ptr()),
                                           vp(ev.data_ptr()), ORDER, SEG, SEG_TIME, None, vp(h_status.data_ptr()),
                                           vp(h_solved.data_ptr()), vp(h_coef.data_ptr())))
            if world_size > 1:
                d_coef.copy_(h_coef, non_blocking=True)
                dist.all_gather_into_tensor(g_coef, d_coef)

        for i in range(min(W, 2)):
            step_host(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(W, W + K):
            step_host(i)
        barrier()
        e2e_s = time.perf_counter() - t0
    clk = clocks.stop() if rank == 0 else None

    t_ms = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t_ms.tolist()
    total_q = B * world_size * K
    value = total_q / (ms / 1e3)
    e2e_val = total_q / (e2e_ms / 1e3)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        s_ms = float(np.mean(search_ms))
        achieved = float(np.mean(bytes_a)) / (s_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("kino_search_kernel_dram_bytes_per_launch")
        out = {
            "metric": METRIC, "value": value, "unit": "plans/s", "n_gpus": world_size, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(B, world_size),
            "e2e": {"value": e2e_val, "unit": "plans/s", "h2d_bytes_per_step": B * 12 * 8,
                    "d2h_bytes_per_step": B * (3 * n * 8 + 8), "ms_per_step": e2e_ms / K},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "kino_search_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": s_ms, "kernel_share_of_step": s_ms / (ms / K),
                         "algorithmic_bytes_per_launch": float(np.mean(bytes_a)),
                         "expansions_per_s": float(np.mean(pops)) / (s_ms * 1e-3), "qp_exposed_ms": float(np.mean(qp_ms))},
            "clocks": clk,
            "result_check": {"reach_end_frac_last_step": reached / B, "qp_solved_frac_last_step": solved / B},
        }
        if world_size == 1 and not args.no_cpu:
            from uav_motion_planning_b200 import _lib
            threads = os.cpu_count() or 1
            n_s = args.cpu_sample or max(64, 4 * threads)
            dt, nr, _ = cpu_plans(world, ka.params, batches[W], list(range(n_s)), threads)
            kind, how = kind_of_cpu_path()
            out["cpu_baseline"] = {"value": n_s / dt, "unit": "plans/s", "cores": threads, "kind": kind,
                                   "sample": f"first {n_s} queries of the first timed batch, {threads} threads; {how}"}
        print(json.dumps(out), flush=True)
    # Orderly teardown.  The tensors above were allocated while the library's stream was torch's current stream: the caching
    # allocator (and NCCL's record_stream) records events on that stream when they are freed, so they must go BEFORE the
    # library context destroys the stream; a 2-GPU run crashed at exit ("context is destroyed") before this ordering existed.
    del d_in, d_status, d_solved, d_coef, flush, h_in, h_status, h_solved, h_coef
    if world_size > 1:
        del g_coef, g_stat
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()
    ka = None
    ctx.close()
    sys.stdout.flush()
    os._exit(0)  # skip interpreter finalisers: nothing left to do, and nothing may touch the destroyed stream afterwards


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="queries per GPU per step (configs[1]: 4096)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0: auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world_size)
    else:
        run_gpu(args, rank, world_size, local_rank)


if __name__ == "__main__":
    main()
