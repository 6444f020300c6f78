"""Host-side logic of the row-stripe partition (lp_mp_b200/stripes.py) on CPU: the partition plan and, with
world_size 2 and 3 over the gloo backend, the neighbour exchange that carries the halo vectors and the partial
lower bounds.  The kernels themselves are covered by the -m gpu suite (two stripes on one device, bit parity)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lp_mp_b200.stripes import BOTTOM, TOP, StripePlan, neighbour_exchange


def test_plan_partitions_rows_contiguously():
    for H, world in ((1024, 8), (10, 3), (7, 7), (4096, 5)):
        plan = StripePlan(H, world)
        nxt = 0
        for r in range(world):
            row0, n = plan.rows(r)
            assert row0 == nxt and n >= 1
            nxt += n
            assert plan.owner(row0) == r and plan.owner(row0 + n - 1) == r
        assert nxt == H
        sizes = [plan.rows(r)[1] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        StripePlan(3, 4)
    assert not StripePlan(8, 2).has_above(0) and StripePlan(8, 2).has_below(0) and not StripePlan(8, 2).has_below(1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = StripePlan(40, world)
        row0, rows = plan.rows(rank)
        W = 6
        mine_bottom = torch.full((W,), float(100 + rank))     # stands for aD of my last row
        mine_top = torch.full((W,), float(200 + rank))        # stands for bU of my first row
        in_top, in_bot = torch.full((W,), -1.0), torch.full((W,), -1.0)
        # forward pass halo: boundary BOTTOM travels down into the lower neighbour's in_top
        neighbour_exchange(dist, rank, world, BOTTOM, mine_bottom if plan.has_below(rank) else None, in_top if plan.has_above(rank) else None)
        # backward pass halo: boundary TOP travels up into the upper neighbour's in_bot
        neighbour_exchange(dist, rank, world, TOP, mine_top if plan.has_above(rank) else None, in_bot if plan.has_below(rank) else None)
        ok = True
        ok &= bool((in_top == (100 + rank - 1)).all()) if plan.has_above(rank) else bool((in_top == -1).all())
        ok &= bool((in_bot == (200 + rank + 1)).all()) if plan.has_below(rank) else bool((in_bot == -1).all())
        # lower bound: partial sums all-reduced in double
        t = torch.tensor([float(rows)], dtype=torch.float64)
        dist.all_reduce(t)
        ok &= t.item() == 40.0
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_neighbour_exchange_over_gloo(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert list(out) == [1] * world
