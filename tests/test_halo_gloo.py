"""CPU tests of the latitude-band plumbing (N > 1 path) with the gloo backend, world_size 2 and 3.
The CUDA kernels are not involved here: what is checked is that every rank ends up with exactly the
rows of the global tensor its windows promise, for the forward (field) and the backward
(grad_out | u | v packed) exchange."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from paradis_model_b200 import halo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, cfl, interp, q, balance=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = torch.randn(2, 3, H, W, generator=g)
        plan = halo.make_plan(H, W, rank, world, cfl, interp, balance=balance)
        own = full[:, :, plan.row0:plan.row0 + plan.rows].contiguous()
        ext = halo.exchange_rows(own, plan)
        ok = torch.equal(ext, full[:, :, plan.ext_row0:plan.ext_row0 + plan.ext_rows])
        # backward exchange: three tensors in one batch
        a, b, c = halo.exchange_rows_multi([own, own * 2, own * 3], plan)
        ref = full[:, :, plan.ext_row0:plan.ext_row0 + plan.ext_rows]
        ok = ok and torch.equal(a, ref) and torch.equal(b, ref * 2) and torch.equal(c, ref * 3)
        (own_w, ext_w) = plan.windows()
        ok = ok and own_w == (plan.row0, plan.rows) and ext_w[0] >= 0 and ext_w[0] + ext_w[1] <= H
        q.put((rank, bool(ok), plan.lo, plan.hi, plan.rows))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W,balance", [(2, 64, 32, False), (3, 91, 32, False), (3, 181, 360, True)])
def test_exchange_rows_gloo(world, H, W, balance):
    """balance=True: bands of unequal height from the cost-levelled split (polar bands thinner)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, 3.0, "bilinear", q, balance)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert res[0][2] == 0 and res[-1][3] == 0           # no halo beyond the poles
    assert all(r[3] == halo.halo_rows(3.0, "bilinear") for r in res[:-1])
    assert sum(r[4] for r in res) == H
    if balance:
        assert res[0][4] < res[1][4] and res[-1][4] < res[1][4]


def test_band_rows_and_plan():
    assert halo.band_rows(721, 8) == [(0, 91)] + [(91 + 90 * k, 90) for k in range(7)]
    assert sum(n for _, n in halo.band_rows(721, 4)) == 721
    p = halo.make_plan(721, 1440, 3, 8, 6.0, "bilinear")
    assert (p.row0, p.rows, p.halo, p.lo, p.hi) == (271, 90, 9, 9, 9)
    assert p.ext_row0 == 262 and p.ext_rows == 108
    assert halo.make_plan(721, 1440, 0, 8, 6.0, "bicubic").lo == 0
    with pytest.raises(ValueError):
        halo.make_plan(32, 64, 0, 8, 6.0)
    single = halo.make_plan(32, 64, 0, 1, 6.0)
    assert single.ext_rows == 32
    # cost-balanced bands: the ranks that own a polar cap get fewer rows, every row is owned once
    cost = halo.row_costs(721, 1440, 6.0)
    bands = halo.band_rows(721, 8, cost)
    assert bands[0][0] == 0 and sum(n for _, n in bands) == 721
    assert all(bands[k][0] + bands[k][1] == bands[k + 1][0] for k in range(7))
    assert bands[0][1] < bands[3][1] and bands[7][1] < bands[4][1]
    # make_plan adds the per-band warm-up overhead and the minimum thickness (halo + 1 rows): symmetric, polar bands thinner still
    plans = [halo.make_plan(721, 1440, r, 8, 6.0, balance=True) for r in range(8)]
    assert sum(p.rows for p in plans) == 721 and all(p.rows > p.halo for p in plans)
    assert plans[0].rows <= bands[0][1] and abs(plans[0].rows - plans[7].rows) <= 1
    assert abs(plans[3].rows - plans[4].rows) <= 1
    # the levelled split never has a band more expensive than the equal-height split's most expensive one
    S = [0.0]
    for c in cost:
        S.append(S[-1] + c)
    worst = lambda bs: max(S[a + n] - S[a] for a, n in bs)
    assert worst(bands) <= worst(halo.band_rows(721, 8))
    with pytest.raises(ValueError):
        halo.band_rows(20, 4, [1.0] * 20, 0.0, 6)
