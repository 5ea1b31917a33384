"""CPU tests of the multi-GPU plumbing (mvd.dist) with world_size = 2 over gloo: trials shard by
global trial id, each rank tallies its shard (here with the CPU oracle standing in for the
kernel), one all_reduce(SUM) combines them, and the result equals the single-process tallies."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from mvd import dist
    for total in (0, 1, 7, 1000, 10**6 + 3):
        for W in (1, 2, 3, 4, 8):
            cuts = [dist.shard_range(total, r, W, offset=5) for r in range(W)]
            assert cuts[0][0] == 5 and cuts[-1][1] == 5 + total
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dist.shard_range(10, 2, 2)


def test_world_defaults_to_single_process():
    from mvd import dist
    assert dist.world() == (0, 1)
    v = np.array([3, 4], dtype=np.uint64)
    assert np.array_equal(dist.allreduce_sum(v), v)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    for p in (os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200"),
              os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as td
    import c_oracle as co
    from mvd import bitsource, codes, dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert dist.world() == (rank, world)
        gen1 = codes.freeze_generator([[[1, 1, 1]], [[1, 0, 1]]])
        gen2 = codes.freeze_generator([[[1, 1, 0]], [[1, 0, 1]]])
        tab = codes.enumerate_states(gen1, 2, 1, 2)
        t1, t2 = codes.tap_masks(gen1, 2, 1), codes.tap_masks(gen2, 2, 1)
        otab = co.Table(tab.metrics, 2)
        T = bitsource.bsc_threshold(0.1)
        edge, _ = co.learn_chain(t1, t1, 2, 2, 6200, 200, T, 123, bitsource.LEARN_STREAM, 0, otab)   # replicated
        P1 = codes.p1_from_edge_counts(tab, edge, 1.0)
        Tref = codes.tref_half_table(tab)
        total, N = 301, 100
        begin, end = dist.shard_range(total, rank, world)
        mine = np.array([co.run_trials(t1, t1, 2, 2, N, T, 123, 0, begin, end, otab, P1, Tref, 0),
                         co.run_trials(t1, t2, 2, 2, N, T, 123, 1, begin, end, otab, P1, Tref, 1)], dtype=np.int64)
        summed = dist.allreduce_sum(mine)
        q.put((rank, mine.tolist(), summed.tolist()))
    finally:
        td.destroy_process_group()


def test_two_rank_tallies_equal_single_process():
    import multiprocessing as mp
    import c_oracle as co
    from mvd import bitsource, codes
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][2] == res[1][2]                                   # both ranks hold the same sum
    assert [a + b for a, b in zip(res[0][1], res[1][1])] == res[0][2]
    # single-process reference over the whole trial range
    gen1 = codes.freeze_generator([[[1, 1, 1]], [[1, 0, 1]]])
    gen2 = codes.freeze_generator([[[1, 1, 0]], [[1, 0, 1]]])
    tab = codes.enumerate_states(gen1, 2, 1, 2)
    t1, t2 = codes.tap_masks(gen1, 2, 1), codes.tap_masks(gen2, 2, 1)
    otab = co.Table(tab.metrics, 2)
    T = bitsource.bsc_threshold(0.1)
    edge, _ = co.learn_chain(t1, t1, 2, 2, 6200, 200, T, 123, bitsource.LEARN_STREAM, 0, otab)
    P1 = codes.p1_from_edge_counts(tab, edge, 1.0)
    Tref = codes.tref_half_table(tab)
    whole = [co.run_trials(t1, t1, 2, 2, 100, T, 123, 0, 0, 301, otab, P1, Tref, 0),
             co.run_trials(t1, t2, 2, 2, 100, T, 123, 1, 0, 301, otab, P1, Tref, 1)]
    assert whole == res[0][2]


def _lazy_worker(rank, world, port, q):
    """What `torchrun ... Pd_plotter.py` gives a process: the rendezvous environment, but nobody has called
    init_process_group.  mvd.dist must join the job by itself (silent single-process fallback)."""
    sys.path.insert(0, os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), CUDA_VISIBLE_DEVICES="")
    import torch.distributed as td
    from mvd import dist
    assert not td.is_initialized()
    got = dist.world()                                      # joins lazily (gloo: no CUDA device here)
    assert td.is_initialized() and dist.backend() == "gloo"
    begin, end = dist.shard_range(1001, *got)
    total = dist.allreduce_sum(np.array([end - begin, rank + 1], dtype=np.int64))
    q.put((rank, got, total.tolist(), dist.is_rank0()))
    td.destroy_process_group()


def test_world_joins_the_torchrun_job_lazily():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lazy_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [(0, 2), (1, 2)]
    assert res[0][2] == res[1][2] == [1001, 3]
    assert [r[3] for r in res] == [True, False]
