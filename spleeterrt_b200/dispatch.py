"""Stream dispatcher for the multi-GPU box.

Units (stream, tile, stem) are independent, so the path shards with NO data-path collective:
stream i goes to GPU i mod G, every rank holds a full weight replica, and the process group is only
used for the start/stop barrier, the max-over-ranks timing reduction and gathering per-rank records
(timings, output checksums).  Works with any torch.distributed backend (nccl on the B200 box, gloo in
the CPU tests)."""
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
