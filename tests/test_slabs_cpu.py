"""CPU: host-side logic of the z-slab decomposition, including a world_size-2 gloo exchange that runs
the same halo_ops() plumbing the NCCL transport uses on the GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_and_neighbours(built):
    from lbm_b200.slabs import partition, neighbours
    assert partition(512, 1) == [(1, 512)]
    assert partition(4096, 8) == [(1 + 512 * r, 512) for r in range(8)]
    p = partition(13, 4)
    assert [n for _, n in p] == [4, 3, 3, 3] and p[0][0] == 1
    assert all(p[i][0] + p[i][1] == p[i + 1][0] for i in range(3)) and p[-1][0] + p[-1][1] - 1 == 13
    with pytest.raises(ValueError):
        partition(3, 4)
    assert neighbours(0, 4) == (None, 1) and neighbours(3, 4) == (2, None) and neighbours(1, 4) == (0, 2)
    assert neighbours(0, 4, periodic_z=True) == (3, 1) and neighbours(3, 4, periodic_z=True) == (2, 0)
    assert neighbours(0, 1, periodic_z=True) == (None, None)


def _worker(rank, world, port, n_q, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    from lbm_b200.slabs import halo_ops, neighbours, DOWN, UP
    dist.init_process_group("gloo", rank=rank, world_size=world)
    down, up = neighbours(rank, world)
    # send planes carry (rank, side, k); after the exchange the DOWN ghost planes of rank r must hold
    # what rank r-1 sent UP, and the UP ghost planes what rank r+1 sent DOWN
    send = {s: [torch.full((n,), 100.0 * rank + 10.0 * s + k, dtype=torch.float64) for k in range(n_q)] for s in (DOWN, UP)}
    recv = {s: [torch.full((n,), -1.0, dtype=torch.float64) for _ in range(n_q)] for s in (DOWN, UP)}
    ops = halo_ops(dist, send, recv, down, up)
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    ok = True
    for k in range(n_q):
        if down is not None:
            ok &= bool((recv[DOWN][k] == 100.0 * down + 10.0 * UP + k).all())
        else:
            ok &= bool((recv[DOWN][k] == -1.0).all())
        if up is not None:
            ok &= bool((recv[UP][k] == 100.0 * up + 10.0 * DOWN + k).all())
        else:
            ok &= bool((recv[UP][k] == -1.0).all())
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_q", [(2, 5), (3, 9)])
def test_halo_exchange_plumbing_over_gloo(built, world, n_q):
    ret = mp.get_context("spawn").Manager().dict()
    port = 29600 + world
    mp.spawn(_worker, args=(world, port, n_q, 257, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)) and len(ret) == world
