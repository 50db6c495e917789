"""N>1 host logic on CPU: world_size-2 gloo process group exercising the stream dispatcher
(placement, barrier, max-over-ranks timing, record gather) exactly as bench.py uses it."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spleeterrt_b200 import dispatch as D
    mine = D.stream_ids_for_rank(total, world, rank)
    D.barrier()
    step_ms = 10.0 + 5.0 * rank                      # rank 1 is the slow one
    slowest = D.max_over_ranks(step_ms)
    n_all = D.sum_over_ranks(len(mine))
    recs = D.gather_records({"rank": rank, "streams": mine, "checksum": float(sum(mine))})
    q.put((rank, mine, slowest, n_all, recs))
    dist.destroy_process_group()


def test_stream_dispatch_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, total, port = 2, 7, _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    all_streams = sorted(got[0][1] + got[1][1])
    assert all_streams == list(range(total))                 # every stream placed exactly once
    assert got[0][1] == [0, 2, 4, 6] and got[1][1] == [1, 3, 5]
    for rank, mine, slowest, n_all, recs in got:
        assert slowest == 15.0                               # max over ranks, seen by every rank
        assert n_all == total
        assert [r["rank"] for r in recs] == [0, 1]
        assert recs[1]["checksum"] == 9.0


def test_single_process_passthrough():
    from spleeterrt_b200 import dispatch as D
    assert D.stream_ids_for_rank(5, 1, 0) == [0, 1, 2, 3, 4]
    assert D.max_over_ranks(3.5) == 3.5 and D.gather_records({"a": 1}) == [{"a": 1}]


def _fake_separate(streams):
    """Stands in for Separator.separate on a CPU rank: two 'stems' that are cheap functions of the input."""
    import numpy as np
    return [np.stack([np.stack([l + r, l - r]), np.stack([0.5 * l, 2.0 * r])]).astype(np.float32) for l, r in streams]


def _worker_scatter(rank, world, port, q):
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spleeterrt_b200 import dispatch as D
    rng = np.random.default_rng(7)
    lens = [1000, 37, 4096, 5, 2500]
    streams = [(rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)) for n in lens]
    nets = D.broadcast_nets([(np.arange(10, dtype=np.float32), 1), (np.ones(4, np.float32), 0)] if rank == 0 else None)
    mine, ids = D.scatter_streams(streams if rank == 0 else None)
    ok_scatter = all(np.array_equal(mine[k][0], streams[i][0]) and np.array_equal(mine[k][1], streams[i][1]) for k, i in enumerate(ids))
    res = D.separate_across_ranks(_fake_separate, streams if rank == 0 else None)
    ok = None
    if rank == 0:
        want = _fake_separate(streams)
        ok = len(res) == len(want) and all(np.array_equal(a, b) for a, b in zip(res, want))
    q.put((rank, ids, ok_scatter, ok, [(c.tolist(), m) for c, m in nets], res is None))
    dist.destroy_process_group()


def test_scatter_separate_gather_world2():
    """Rank 0 owns the PCM: weights broadcast once, streams scattered round-robin with grouped send/recv, results
    gathered back in global order (the single-entry-point variant)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_worker_scatter, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == [0, 2, 4] and got[1][1] == [1, 3]
    assert got[0][2] and got[1][2]                         # every rank received exactly its streams
    assert got[0][3] is True and got[1][5] is True         # rank 0 has everything, in order; rank 1 gets None
    assert got[0][4] == got[1][4] and got[1][4][0][0] == list(range(10)) and got[1][4][1][1] == 0
