"""world_size-2 gloo test of the multi-rank sharding / mask gather (host logic only)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mjpl_b200.parallel import gather_masks, reduce_counts, shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 1000, 10**9 + 3):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from mjpl_b200 import models
        from mjpl_b200.engine import sweep_rows_host

        model = models.load("two_dof_ball")
        orc = oracle.Oracle(model)
        lo, hi = shard_range(n_total, rank, world)
        rows = sweep_rows_host(model, 5, lo, hi - lo)              # rows keyed by GLOBAL row id
        local = orc.check(rows.astype(np.float64), 3)              # stand-in for the GPU mask
        full = gather_masks(local.astype(np.uint8), n_total)
        cnt = reduce_counts([int(local.sum()), hi - lo])
        q.put((rank, full.numpy().copy(), cnt.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_mask_gather_matches_single_rank():
    import oracle
    from mjpl_b200 import models
    from mjpl_b200.engine import sweep_rows_host

    n_total, world = 1001, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = models.load("two_dof_ball")
    want = oracle.Oracle(model).check(sweep_rows_host(model, 5, 0, n_total).astype(np.float64), 3)
    for rank, full, cnt in res:
        np.testing.assert_array_equal(full.astype(bool), want)
        assert cnt.tolist() == [int(want.sum()), n_total]


def test_sweep_rows_host_is_deterministic_and_in_range():
    from mjpl_b200 import models
    from mjpl_b200.engine import sweep_rows_host

    m = models.load("franka_scene")
    a = sweep_rows_host(m, 1, 100, 50)
    b = sweep_rows_host(m, 1, 0, 200)[100:150]
    np.testing.assert_array_equal(a, b)  # keyed by global row id, independent of the chunking
    lo, hi = m.jnt_range[:, 0].astype(np.float32), m.jnt_range[:, 1].astype(np.float32)
    big = sweep_rows_host(m, 2, 0, 20000)
    assert (big >= lo).all() and (big <= hi).all()
    assert abs(((big - lo) / (hi - lo)).mean() - 0.5) < 0.01
    assert not np.array_equal(sweep_rows_host(m, 3, 0, 10), sweep_rows_host(m, 4, 0, 10))


def _retake_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench

        # rank 1 alone wants the first pass again, nobody the second: every rank must see (True, False) --
        # bench.py's timed pass contains barriers, so the ranks have to agree on how many passes they take
        seen = [bench.any_rank(rank == 1 and k == 0, dist, torch.device("cpu")) for k in range(2)]
        q.put((rank, seen))
    finally:
        dist.destroy_process_group()


def test_retaking_a_timed_pass_is_a_collective_decision():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_retake_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert got == {0: [True, False], 1: [True, False]}
    import bench

    assert bench.any_rank(True, None, None) is True and bench.any_rank(False, None, None) is False
