"""Concurrent host<->device copy rates on 1..N GPUs of one box: is the end-to-end curve of bench.py (452 MB of fp32 stems
back to the host per GPU and step) bounded by the platform or by this library?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/pcie_probe_ranks.py

No kernels of the library run here: every rank owns pinned host buffers of the bench's sizes (113 MB in, 452 MB out per step) and
copies them with cudaMemcpyAsync (torch copy_ on side streams).  For each active-set size k in (1, 2, 4, ..., world) only ranks
< k copy (the others idle at the barriers), in three modes: D2H alone, H2D alone, both directions at once.  Rates are taken with
CUDA events per rank over the same barrier-bracketed window; rank 0 prints one JSON record with the per-rank and the aggregate
GB/s, plus what the OS says about the topology (NUMA node and CPU list per GPU, `nvidia-smi topo -m`).
Optional first argument "pin": each rank binds itself to the CPUs of its GPU's NUMA node before allocating the pinned buffers.
"""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

N = 441000
REPS = 6


def numa_of_gpu(index):
    try:
        bdf = subprocess.check_output(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"], text=True).strip().lower()
        bdf = bdf[-12:] if len(bdf) > 12 else bdf          # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        cpus = open(f"/sys/devices/system/node/node{max(node, 0)}/cpulist").read().strip()
        return {"bdf": bdf, "numa_node": node, "cpulist": cpus}
    except Exception as e:
        return {"error": repr(e)}


def cpus_from_list(s):
    out = []
    for part in s.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def main():
    pin = len(sys.argv) > 1 and sys.argv[1] == "pin"
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    topo = numa_of_gpu(local)
    if pin and "cpulist" in topo:
        try:
            os.sched_setaffinity(0, cpus_from_list(topo["cpulist"]))
        except Exception as e:
            topo["pin_error"] = repr(e)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h_in = torch.empty((32, 2, N), dtype=torch.float32).pin_memory()
    h_out = torch.empty((32, 4, 2, N), dtype=torch.float32).pin_memory()
    h_in.zero_(); h_out.zero_()                                  # touch the pages from this (possibly pinned) process
    d_in, d_out = h_in.cuda(), torch.empty_like(h_out, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(mode, active):
        barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t0 = time.perf_counter()
        if active:
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s_in):
                    e[0].record()
                    for _ in range(REPS):
                        d_in.copy_(h_in, non_blocking=True)
                    e[1].record()
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s_out):
                    e[2].record()
                    for _ in range(REPS):
                        h_out.copy_(d_out, non_blocking=True)
                    e[3].record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        rec = {"h2d_gbs": 0.0, "d2h_gbs": 0.0, "wall_s": wall}
        if active and mode in ("h2d", "both"):
            rec["h2d_gbs"] = h_in.numel() * 4 * REPS / (e[0].elapsed_time(e[1]) * 1e-3) / 1e9
        if active and mode in ("d2h", "both"):
            rec["d2h_gbs"] = h_out.numel() * 4 * REPS / (e[2].elapsed_time(e[3]) * 1e-3) / 1e9
        barrier()
        if world > 1:
            allrec = [None] * world
            dist.all_gather_object(allrec, rec)
        else:
            allrec = [rec]
        return allrec

    run("both", True)                                            # warm-up
    results = []
    k = 1
    sizes = []
    while k <= world:
        sizes.append(k)
        k *= 2
    for k in sizes:
        for mode in ("d2h", "h2d", "both"):
            recs = run(mode, rank < k)
            if rank == 0:
                act = recs[:k]
                results.append({"active_gpus": k, "mode": mode,
                                "d2h_gbs_per_gpu": [round(r["d2h_gbs"], 1) for r in act], "h2d_gbs_per_gpu": [round(r["h2d_gbs"], 1) for r in act],
                                "d2h_gbs_aggregate": round(sum(r["d2h_gbs"] for r in act), 1), "h2d_gbs_aggregate": round(sum(r["h2d_gbs"] for r in act), 1)})
    topos = [None] * world
    if world > 1:
        dist.all_gather_object(topos, topo)
    else:
        topos = [topo]
    if rank == 0:
        try:
            tm = subprocess.check_output(["nvidia-smi", "topo", "-m"], text=True, timeout=20)
        except Exception as e:
            tm = repr(e)
        print(json.dumps({"what": "concurrent pinned-memory copies, bench.py's per-step sizes (113 MB H2D, 452 MB D2H per GPU), no kernels",
                          "world": world, "cpu_pinning": pin, "host_cpus": os.cpu_count(), "gpu_topology": topos, "results": results,
                          "nvidia_smi_topo": tm}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
