"""Gelman-Rubin R-hat from per-chain Welford summaries, and the one collective of the path.

R-hat does not exist in the reference (its only posterior summary is rawstats, Framework.py:11-17); the
definition is pinned in SURVEY.md §8e: per parameter on x = ln(theta), m chains of n kept samples,
W = mean_j s_j^2, B = n var_j(mean_j) (ddof=1), var+ = (n-1)/n W + B/n, R-hat = sqrt(var+/W).
"""
from __future__ import annotations

import numpy as np


def rhat_from_summaries(summaries, n_param):
    """summaries [m, 1+2P] = (count, mean[P], M2[P]) per chain (numpy) -> R-hat [P]."""
    s = np.asarray(summaries, dtype=np.float64)
    n = s[:, 0]
    if not np.all(n == n[0]):
        raise ValueError("chains have different numbers of kept samples")
    n = float(n[0])
    means = s[:, 1:1 + n_param]
    var = s[:, 1 + n_param:1 + 2 * n_param] / (n - 1.0)
    W = var.mean(axis=0)
    B = n * means.var(axis=0, ddof=1)
    with np.errstate(all="ignore"):
        return np.sqrt(((n - 1.0) / n * W + B / n) / W)


def ess_from_summaries(summaries, n_param):
    """Effective sample size per parameter from the same per-chain summaries: n_eff = m n var+ / B (Gelman et al.,
    Bayesian Data Analysis, the between/within-chain estimate that goes with R-hat), capped at the m n kept rows.
    Needs no autocorrelations, so no sample ever leaves the device for it; it is conservative while the chains have
    not mixed (B large) and saturates at m n once the chain means agree.  New functionality like R-hat (SURVEY.md
    §8 f2); not in the reference."""
    s = np.asarray(summaries, dtype=np.float64)
    n = s[:, 0]
    if not np.all(n == n[0]):
        raise ValueError("chains have different numbers of kept samples")
    m, n = float(len(s)), float(n[0])
    means = s[:, 1:1 + n_param]
    var = s[:, 1 + n_param:1 + 2 * n_param] / (n - 1.0)
    W = var.mean(axis=0)
    B = n * means.var(axis=0, ddof=1)
    with np.errstate(all="ignore"):
        ess = m * n * ((n - 1.0) / n * W + B / n) / B
    return np.where(np.isfinite(ess), np.minimum(ess, m * n), m * n)


def pooled_log_stats(summaries, n_param):
    """Per-chain Welford summaries of ln(theta) -> (count, log_mean[P], log_std[P]) of ALL kept rows pooled.

    log_std uses ddof=1 like pandas' Series.std(); these are the two numbers the reference's rawstats
    (Framework.py:11-17) takes from the posterior frame, so the fitting report needs no frame at all."""
    s = np.asarray(summaries, dtype=np.float64)
    n = s[:, 0]
    means = s[:, 1:1 + n_param]
    m2 = s[:, 1 + n_param:1 + 2 * n_param]
    N = n.sum()
    if N < 1:
        return 0.0, np.full(n_param, np.nan), np.full(n_param, np.nan)
    mean = (n[:, None] * means).sum(axis=0) / N
    M2 = m2.sum(axis=0) + (n[:, None] * (means - mean) ** 2).sum(axis=0)
    with np.errstate(all="ignore"):
        std = np.sqrt(M2 / (N - 1.0)) if N > 1 else np.full(n_param, np.nan)
    return float(N), mean, std


def shard_bounds(n_total, world_size, rank):
    """Contiguous block of chains / parameter sets owned by `rank`."""
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_summaries(local, n_total=None, group=None):
    """All-gather per-chain summaries over the process group (NCCL on GPUs, gloo in CPU tests).

    local: torch tensor [m_local, 1+2P].  Shards may differ by one row; they are padded to the largest
    shard for the collective and trimmed afterwards.  Returns a tensor [m_total, 1+2P] on every rank.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    ws = dist.get_world_size(group)
    counts = torch.zeros(ws, dtype=torch.int64, device=local.device)
    counts[dist.get_rank(group)] = local.shape[0]
    dist.all_reduce(counts, group=group)
    mmax = int(counts.max().item())
    pad = torch.zeros((mmax, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((ws * mmax, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = [out[r * mmax: r * mmax + int(counts[r].item())] for r in range(ws)]
    return torch.cat(parts, dim=0)


def _collective_device(group=None):
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def allgather_rows(local, group=None):
    """All-gather an array whose FIRST axis is sharded in contiguous blocks over the ranks (shards may differ in
    length, also be empty).  numpy in -> numpy out, torch in -> torch out (on the collective's device)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    is_np = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if is_np else local
    t = t.to(_collective_device(group)).contiguous()
    tail = tuple(t.shape[1:])
    width = 1
    for d in tail:
        width *= int(d)
    flat = t.reshape(t.shape[0], width)                           # explicit width: an empty shard has no -1 to infer
    out = allgather_summaries(flat, group=group)
    out = out.reshape((out.shape[0],) + tail)
    return out.cpu().numpy() if is_np else out


def broadcast_rows(arr, shape, src=0, group=None):
    """Broadcast a float64 table from rank `src` (numpy or torch there, None elsewhere); every rank gets a torch tensor
    on the collective's device."""
    import torch
    import torch.distributed as dist
    dev = _collective_device(group)
    if dist.get_rank(group) == src:
        t = (torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)) if isinstance(arr, np.ndarray) else arr)
        t = t.to(dev, dtype=torch.float64).contiguous()
        assert tuple(t.shape) == tuple(shape)
    else:
        t = torch.empty(tuple(shape), dtype=torch.float64, device=dev)
    dist.broadcast(t, src=src, group=group)
    return t


def sharded_mcmc(dm, theta0_all, group=None, **mcmc_kw):
    """Run this rank's contiguous block of chains and all-gather the chain summaries for R-hat.

    theta0_all [C_total, P] is the same on every rank; chain c keeps its global index (Philox key / seed), so
    the per-chain results do not depend on the number of GPUs.  Returns (local result dict, R-hat[P] over all
    chains, (lo, hi))."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if ws > 1 else 0
    lo, hi = shard_bounds(len(theta0_all), ws, rank)
    res = dm.mcmc(theta0_all[lo:hi], chain_offset=lo, device_buffers=True, **mcmc_kw)
    # the one collective of the path, behind the C ABI: odl_rhat = ncclAllGather of the chain summaries + reduction on
    # the device (torch.distributed only carried the communicator's id to the ranks)
    dm.comm_init(group)
    rh, _, _ = dm.rhat(res["summaries"])
    return res, rh, (lo, hi)
