"""Stream dispatcher for the multi-GPU box.

Units (stream, tile, stem) are independent, so the path shards with NO data-path collective:
stream i goes to GPU i mod G, every rank holds a full weight replica, and the process group is only
used for the start/stop barrier, the max-over-ranks timing reduction and gathering per-rank records
(timings, output checksums).  Works with any torch.distributed backend (nccl on the B200 box, gloo in
the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def stream_ids_for_rank(total_streams, world, rank):
    """Round-robin placement: stream i -> rank i mod world (all tiles and stems of a stream stay on one GPU)."""
    return list(range(rank, total_streams, world))


def is_dist():
    return dist.is_available() and dist.is_initialized()


def barrier():
    if is_dist():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """Reduce a python float with MAX over all ranks (the step time of the slowest GPU)."""
    if not is_dist():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not is_dist():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_records(record):
    """All ranks contribute one small picklable record (timing, checksum); every rank gets the list."""
    if not is_dist():
        return [record]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, record)
    return out


# ---------------------------------------------------------------------------------------------
# One entry point for the whole box (second variant): rank 0 owns the PCM of all streams, scatters each
# rank its share with grouped point-to-point transfers (ncclSend / ncclRecv over NVLink on the B200 box), every rank
# separates its streams on its own GPU, and the stems come back to rank 0 the same way.  The weights go out once with
# a broadcast.  Still no collective on the data path between the kernels: streams never meet.
# ---------------------------------------------------------------------------------------------


def _comm_device():
    """NCCL moves device memory; gloo (CPU tests) moves host memory."""
    if is_dist() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _world():
    return (dist.get_world_size(), dist.get_rank()) if is_dist() else (1, 0)


def broadcast_nets(nets, src=0):
    """nets: on `src` a list of (coeff float32[9822725], stemMode); None elsewhere.  Every rank returns the list
    (one ncclBroadcast of the concatenated blobs)."""
    world, rank = _world()
    if world == 1:
        return nets
    meta = [[(int(np.asarray(c).size), int(m)) for c, m in nets]] if rank == src else [None]
    dist.broadcast_object_list(meta, src=src)
    sizes = meta[0]
    dev = _comm_device()
    flat = torch.empty(sum(n for n, _ in sizes), dtype=torch.float32, device=dev)
    if rank == src:
        flat.copy_(torch.from_numpy(np.concatenate([np.asarray(c, np.float32).reshape(-1) for c, _ in nets])))
    dist.broadcast(flat, src=src)
    host = flat.cpu().numpy()
    out, p = [], 0
    for n, m in sizes:
        out.append((np.ascontiguousarray(host[p:p + n]), m))
        p += n
    return out


def _exchange(sends, recvs):
    """sends: [(tensor, peer)], recvs: [(tensor, peer)] as one grouped batch (ncclGroupStart/End under NCCL)."""
    ops = [dist.P2POp(dist.isend, t, peer) for t, peer in sends] + [dist.P2POp(dist.irecv, t, peer) for t, peer in recvs]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def scatter_streams(streams, src=0):
    """streams: on `src` the list of (L, R) float32 arrays of ALL streams in global order; None elsewhere.
    Returns (this rank's [(L, R)], their global ids): stream i lives on rank i mod world."""
    world, rank = _world()
    if world == 1:
        return list(streams), list(range(len(streams)))
    meta = [[int(np.asarray(l).size) for l, _ in streams]] if rank == src else [None]
    dist.broadcast_object_list(meta, src=src)
    lengths = meta[0]
    ids = stream_ids_for_rank(len(lengths), world, rank)
    dev = _comm_device()
    mine = torch.empty(2 * sum(lengths[i] for i in ids), dtype=torch.float32, device=dev)
    sends, recvs = [], []
    if rank == src:
        for r in range(world):
            rid = stream_ids_for_rank(len(lengths), world, r)
            if not rid:
                continue
            pack = np.concatenate([np.concatenate([np.asarray(streams[i][0], np.float32), np.asarray(streams[i][1], np.float32)]) for i in rid])
            if r == src:
                mine.copy_(torch.from_numpy(pack))
            else:
                sends.append((torch.from_numpy(pack).to(dev), r))
    elif ids:
        recvs.append((mine, src))
    _exchange(sends, recvs)
    host, out, p = mine.cpu().numpy(), [], 0
    for i in ids:
        n = lengths[i]
        out.append((np.ascontiguousarray(host[p:p + n]), np.ascontiguousarray(host[p + n:p + 2 * n])))
        p += 2 * n
    return out, ids


def gather_stems(local_outs, n_total, dst=0):
    """local_outs: this rank's results, one float32[pairs][2][n] per local stream (order of scatter_streams).
    On `dst` returns the list for all n_total streams in global order; None elsewhere."""
    world, rank = _world()
    if world == 1:
        return list(local_outs)
    shapes = gather_records([tuple(o.shape) for o in local_outs])          # every rank learns every shape
    dev = _comm_device()
    sends, recvs, bufs = [], [], {}
    if rank == dst:
        for r in range(world):
            if r == dst or not shapes[r]:
                continue
            bufs[r] = torch.empty(sum(int(np.prod(s)) for s in shapes[r]), dtype=torch.float32, device=dev)
            recvs.append((bufs[r], r))
    elif local_outs:
        pack = np.concatenate([np.asarray(o, np.float32).reshape(-1) for o in local_outs])
        sends.append((torch.from_numpy(pack).to(dev), dst))
    _exchange(sends, recvs)
    if rank != dst:
        return None
    out = [None] * n_total
    for r in range(world):
        rid = stream_ids_for_rank(n_total, world, r)
        if r == dst:
            for i, o in zip(rid, local_outs):
                out[i] = np.asarray(o, np.float32)
            continue
        host, p = (bufs[r].cpu().numpy() if r in bufs else np.zeros(0, np.float32)), 0
        for i, s in zip(rid, shapes[r]):
            k = int(np.prod(s))
            out[i] = host[p:p + k].reshape(s).copy()
            p += k
    return out


def separate_across_ranks(process, streams, src=0):
    """`process`: callable taking a list of (L, R) and returning a list of float32[pairs][2][n] (on a GPU rank:
    Separator.separate).  `streams`: all streams on `src`, None elsewhere.  Returns all results on `src`."""
    world, rank = _world()
    n_total = [len(streams)] if rank == src else [None]
    if world > 1:
        dist.broadcast_object_list(n_total, src=src)
    mine, _ = scatter_streams(streams, src=src)
    outs = process(mine) if mine else []
    return gather_stems(outs, n_total[0], dst=src)
