"""N>1 host logic on CPU: contiguous sharding and the one collective of the path (all-gather of per-chain
summaries for Gelman-Rubin R-hat) over gloo with world_size 2 and 3."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from odelib_b200.rhat import allgather_rows, allgather_summaries, broadcast_rows, rhat_from_summaries, shard_bounds
from oracle import odelib_oracle as orc


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 65536, 1000003):
        for ws in (1, 2, 3, 8):
            b = [shard_bounds(n, ws, r) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _summaries(logs):
    m, n, P = logs.shape
    mean = logs.mean(axis=1)
    m2 = ((logs - mean[:, None, :]) ** 2).sum(axis=1)
    return np.concatenate([np.full((m, 1), float(n)), mean, m2], axis=1)


def _worker(rank, ws, port, m_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    rng = np.random.default_rng(5)
    logs = rng.standard_normal((m_total, 40, 3)) * np.array([1.0, 0.1, 3.0]) + rng.standard_normal((m_total, 1, 3)) * 0.3
    full = _summaries(logs)
    lo, hi = shard_bounds(m_total, ws, rank)
    gathered = allgather_summaries(torch.from_numpy(full[lo:hi].copy()))
    np.save(os.path.join(out_dir, f"g{rank}.npy"), gathered.numpy())
    # the facade's sharding helpers: ragged / empty first-axis shards of any rank, numpy and torch; table broadcast
    cube = np.arange(m_total * 4 * 3, dtype=np.float64).reshape(m_total, 4, 3)
    counts = np.arange(m_total, dtype=np.int32)
    assert np.array_equal(allgather_rows(cube[lo:hi]), cube)
    assert np.array_equal(allgather_rows(counts[lo:hi]), counts) and allgather_rows(counts[lo:hi]).dtype == np.int32
    assert torch.equal(allgather_rows(torch.from_numpy(cube[lo:hi].copy())), torch.from_numpy(cube))
    few = cube[:1]                                                # fewer rows than ranks: some shards are empty
    flo, fhi = shard_bounds(1, ws, rank)
    assert np.array_equal(allgather_rows(few[flo:fhi]), few)
    table = broadcast_rows(cube[:, :, 0].copy() if rank == 0 else None, (m_total, 4))
    assert np.array_equal(table.numpy(), cube[:, :, 0])
    dist.destroy_process_group()


@pytest.mark.parametrize("ws,m_total", [(2, 10), (2, 11), (3, 7)])
def test_allgather_summaries_and_rhat_over_gloo(tmp_path, ws, m_total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(ws, port, m_total, str(tmp_path)), nprocs=ws, join=True)
    rng = np.random.default_rng(5)
    logs = rng.standard_normal((m_total, 40, 3)) * np.array([1.0, 0.1, 3.0]) + rng.standard_normal((m_total, 1, 3)) * 0.3
    full = _summaries(logs)
    for r in range(ws):
        g = np.load(tmp_path / f"g{r}.npy")
        np.testing.assert_array_equal(g, full)                      # every rank holds all chains, in global order
        np.testing.assert_allclose(rhat_from_summaries(g, 3), orc.rhat(logs), rtol=1e-12)


def test_rhat_requires_equal_chain_lengths():
    s = _summaries(np.random.default_rng(0).standard_normal((4, 10, 2)))
    s[1, 0] = 9
    with pytest.raises(ValueError):
        rhat_from_summaries(s, 2)
