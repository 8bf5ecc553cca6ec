"""Regenerates profiles/r01_* from the artefacts a GPU run left in gpurun_out/ (run HERE, no GPU needed):
   launches.csv, prof_search.ncu-rep (4096-query launch), prof_qp.ncu-rep, prof_search_B4096.json, bench.log, bench_ref.log"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], zip(rows[1], rows[2])))


def launches():
    rows = list(csv.reader(l for l in open(os.path.join(G, "launches.csv")) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "us" else v / 1e6 if r[ui] == "ns" else v * 1e3 if r[ui] == "s" else v
        name = r[ki].split("(")[0].replace("<unnamed>::", "")[:70]
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{TAG}_launches_bench.md"), "w") as f:
        f.write(f"# {TAG} — ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu` (1 x B200)\n\n"
                "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv "
                "python bench.py --steps 2 --warmup 1 --no-cpu` (times under ncu are serialised / cold-cache: compare SHARES; raw csv "
                f"alongside). Covers the map build, 1 warm-up + 2 timed steps of the device path and of the e2e path.\n\n"
                "| kernel | launches | total ms | share | grid | block |\n|---|---:|---:|---:|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]:.3f} | {a[1] / tot:.4f} | {a[2]} | {a[3]} |\n")
        hot = sum(a[1] for k, a in agg.items() if "kino_search" in k or "qp_solve" in k)
        f.write(f"\nTotal {tot:.1f} ms; this repository's hot-path kernels (`kino_search_kernel`, `qp_solve_*`) hold {hot / tot:.4f} of it. "
                "`k_*` are its glue kernels, `cub::*` the radix sort of the query order, `at::*` torch fills of bench.py's L2-flush buffer.\n")
    shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{TAG}_launches_bench.csv"))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def tobytes(u, v):
    return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)


def kernels():
    reps = [("kino_search_kernel (search + in-kernel QP of 4 096 queries: one launch of the bench workload, tools/ncu_plan.py 4096)", "prof_search.ncu-rep"),
            ("qp_solve_warp_kernel (2 400 problems, order 7, S 8, tools/ncu_qp.py)", "prof_qp.ncu-rep")]
    with open(os.path.join(P, f"{TAG}_ncu_metrics.md"), "w") as f:
        f.write(f"# {TAG} — `ncu --set full --clock-control none --import-source on`, one launch per kernel (1 x B200)\n\n"
                "`.ncu-rep` files stay in gpurun_out/ (scratch); read with `ncu -i … --page raw --csv`. Durations under ncu are not bench values.\n\n")
        for name, rep in reps:
            path = os.path.join(G, rep)
            if not os.path.exists(path):
                continue
            d = raw(path)
            f.write(f"## {name}\n\n| metric | value |\n|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| `{k}` | {d[k][1]} {d[k][0]} |\n")
            f.write("\n")
            if "kino" in name:
                rd, wr = tobytes(*d["dram__bytes_read.sum"]), tobytes(*d["dram__bytes_write.sum"])
                json.dump({"kino_search_kernel_dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                           "source": "ncu --set full --clock-control none -k regex:kino_search -s 1 -c 1 python tools/ncu_plan.py 4096 (search + in-kernel QP)",
                           "gpu_time_duration_ms_under_ncu": float(d["gpu__time_duration.sum"][1]),
                           "l2_hit_pct": float(d["lts__t_sector_hit_rate.pct"][1])}, open(os.path.join(P, "traffic.json"), "w"), indent=1)


def phases():
    src = os.path.join(G, "prof_search_B4096.json")
    if not os.path.exists(src):
        return
    prof = json.load(open(src))
    shutil.copy(src, os.path.join(P, f"{TAG}_prof_search_B4096.json"))
    L = prof["longest_queries"]
    names = ["pop", "shot_path", "tables_tile_grid", "cloud_staging", "cloud_ellipsoid", "dedup_probe_heuristic", "node_write", "heap_commit", "setup"]
    with open(os.path.join(P, f"{TAG}_search_phases.md"), "w") as f:
        f.write(f"# {TAG} — in-kernel phase profile of the search (`tools/prof_search.py 4096`, bench workload)\n\n"
                f"search = {prof['search_ms']:.1f} ms for 4 096 queries, grid {prof['grid']}; SM cycles of thread 0 per phase (`uavmp_kino_set_profile`).\n\n"
                f"| phase | cycles / expansion, all queries | longest query (q{L[0]['q']}, {L[0]['pops']} expansions, status {L[0]['status']}) | "
                f"2nd longest (q{L[1]['q']}, {L[1]['pops']} expansions, status {L[1]['status']}) |\n|---|---:|---:|---:|\n")
        for n in names:
            f.write(f"| {n} | {prof['phase_cycles_per_pop'][n]:.0f} | {L[0]['per_pop'][n]:.0f} | {L[1]['per_pop'][n]:.0f} |\n")
        f.write(f"\nLongest query: {L[0]['cycles']:.3e} SM cycles = the whole kernel; CTA busy fraction {prof['cta_busy_frac']:.2f}.\n")


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    launches()
    kernels()
    phases()
    for src, dst in (("bench.log", f"{TAG}_bench_1gpu.json"), ("bench_ref.log", f"{TAG}_bench_reference_arm.json"),
                     ("bench_2gpu.log", f"{TAG}_bench_2gpu.json"), ("bench_c2.log", f"{TAG}_bench_config2.json"),
                     ("bench_c3.log", f"{TAG}_bench_config3_1gpu.json"), ("bench_c3_2gpu.log", f"{TAG}_bench_config3_2gpu.json"),
                     ("bench_c4.log", f"{TAG}_bench_config4.json"), ("prof_plan.log", f"{TAG}_prof_plan_B4096.json")):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    print("profiles refreshed")
