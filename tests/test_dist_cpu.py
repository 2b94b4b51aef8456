"""World-size-2 gloo test of the multi-GPU host logic (pair sharding + the single all-gather), on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vfm_registration_b200 import dist as vdist


def test_shard_range_covers_everything():
    for n in (0, 1, 5, 8, 9, 512):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b, per = vdist.shard_range(n, r, world)
                assert b - a <= per
                seen += list(range(a, b))
            assert seen == list(range(n))


def _solve(pair):
    pid = pair[0]
    t = np.eye(4)
    t[:3, 3] = [pid, 2 * pid, -pid]
    return t, 0.5 + pid, 0.01 * pid, 100 + pid, pid * 3


def _worker(rank, world, port, n_pairs, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pairs = [(i,) for i in range(n_pairs)]
        res = vdist.register_pairs(pairs, solve_fn=_solve, device=torch.device("cpu"))
        ok = res.T.shape == (n_pairs, 4, 4)
        for i in range(n_pairs):
            ok &= np.array_equal(res.T[i, :3, 3], [i, 2 * i, -i]) and np.array_equal(res.stats[i], [0.5 + i, 0.01 * i, 100 + i, 3 * i])
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [8, 5, 1])
def test_register_pairs_gloo_world2(n_pairs):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0, 0])
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1, 1]


def test_register_pairs_single_process():
    res = vdist.register_pairs([(i,) for i in range(3)], solve_fn=_solve)
    assert res.T.shape == (3, 4, 4) and np.array_equal(res.stats[:, 3], [0, 3, 6])
