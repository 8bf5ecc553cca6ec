"""Multi-GPU host logic: queries are independent, so a batch is cut into contiguous per-rank slices (no data-path exchange)
and the only collective is an all-gather of the solved trajectories (fixed-stride records: coef, qp_solved, search_status).

Works with any torch.distributed backend: NCCL with device tensors on the GPU box (bench.py), gloo with CPU tensors in the
world_size-2 CPU test, where `solve_local` is a stand-in for the CUDA pipeline.
"""
import numpy as np


def shard_range(n_queries, rank, world_size):
    """Contiguous slice [lo, hi) of rank `rank`: ceil(B / G) queries per rank, the last ranks may be short or empty."""
    per = (n_queries + world_size - 1) // world_size
    lo = min(rank * per, n_queries)
    return lo, min(lo + per, n_queries)


def plan_sharded(solve_local, start_pt, start_vel, end_pt, end_vel, n_coef, device="cpu"):
    """Every rank holds the full query arrays, solves its slice with `solve_local(sp, sv, ep, ev) -> dict(search_status,
    qp_solved, coef[b, 3, n_coef])` and all ranks end up with the full result arrays."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    B = len(start_pt)
    per = (B + world - 1) // world
    lo, hi = shard_range(B, rank, world)
    coef = torch.zeros(per, 3 * n_coef, dtype=torch.float64, device=device)
    flags = torch.zeros(per, 2, dtype=torch.int32, device=device)  # search_status, qp_solved
    if hi > lo:
        r = solve_local(start_pt[lo:hi], start_vel[lo:hi], end_pt[lo:hi], end_vel[lo:hi])
        coef[:hi - lo] = torch.as_tensor(np.ascontiguousarray(r["coef"]).reshape(hi - lo, -1), device=device)
        flags[:hi - lo, 0] = torch.as_tensor(np.asarray(r["search_status"], np.int32), device=device)
        flags[:hi - lo, 1] = torch.as_tensor(np.asarray(r["qp_solved"], np.int32), device=device)
    g_coef = torch.empty(world * per, 3 * n_coef, dtype=torch.float64, device=device)
    g_flags = torch.empty(world * per, 2, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(g_coef, coef)
    dist.all_gather_into_tensor(g_flags, flags)
    g_coef, g_flags = g_coef[:B].cpu().numpy(), g_flags[:B].cpu().numpy()
    return dict(search_status=g_flags[:, 0].copy(), qp_solved=g_flags[:, 1].copy(), coef=g_coef.reshape(B, 3, n_coef))
