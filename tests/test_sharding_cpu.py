"""world_size-2 gloo test of the N>1 path's host logic (shard bounds, global member numbering, final gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kmc_dn_b200.sharding import ensemble_statistics, gather_tallies, shard_bounds


def test_shard_bounds_cover_everything():
    for B in (0, 1, 7, 16, 1048576, 1048577):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_member(m, P):
    """Deterministic stand-in for a trajectory that depends only on the GLOBAL member index."""
    rng = np.random.default_rng(1000 + m)
    return rng.random() + 1.0, rng.integers(-50, 50, size=P)


def _worker(rank, world, port, B, P, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(B, world, rank)
    t = torch.tensor([_fake_member(m, P)[0] for m in range(lo, hi)], dtype=torch.float64)
    e = torch.tensor(np.array([_fake_member(m, P)[1] for m in range(lo, hi)]).reshape(hi - lo, P), dtype=torch.int64)
    T, E = gather_tallies(t, e, B)
    if rank == 0:
        np.save(out + "_t.npy", T.numpy()); np.save(out + "_e.npy", E.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank(tmp_path):
    B, P = 13, 3  # uneven split: 7 + 6
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "g")
    mp.spawn(_worker, args=(2, port, B, P, out), nprocs=2, join=True)
    T, E = np.load(out + "_t.npy"), np.load(out + "_e.npy")
    np.testing.assert_array_equal(T, [_fake_member(m, P)[0] for m in range(B)])
    np.testing.assert_array_equal(E, np.array([_fake_member(m, P)[1] for m in range(B)]))


def test_ensemble_statistics_groups_seeds():
    t = np.array([1.0, 2.0, 1.0, 1.0]); eo = np.array([[2, -2], [4, -4], [1, 0], [3, 0]])
    mean, sem = ensemble_statistics(t, eo, 2)
    np.testing.assert_allclose(mean, [[2, -2], [2, 0]])
    np.testing.assert_allclose(sem, [[0, 0], [1 / np.sqrt(2), 0]])
