#!/usr/bin/env python3
"""bench.py — 4-stem 44.1 kHz stereo separation throughput on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                 # the reference's own CPU path (oracle/_ref), all host threads
    torchrun ... bench.py --gpus 8 --stems 5 --streams 128 --dispatch nccl     # BASELINE.json configs[3]

A step = one pass of the whole hot path (STFT framer -> S U-Nets -> mask*spectrum -> iSTFT/OLA) over a batch of `--streams`
synthetic 10 s stereo streams per GPU (T=512, F=1024: BASELINE.json configs[1] shape, batched as configs[2] does).
`value` is device-resident throughput (inputs in HBM when the clock starts), `e2e` goes through the host-pointer C-ABI call with
pinned host buffers, H2D and D2H inside the timed region.  The default run (N = 1) also carries, as bounded sub-objects, the
other BASELINE.json configurations (single stream, 256 streams, the VST block sweep), the second precision mode, a sustained
run, the measured roofline denominators, the CPU baselines (the reference's naive and OpenBLAS builds) and the parity of the
timed batch against the reference.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SECONDS = 10.0
N_SAMPLES = 441000
T, F = 512, 1024


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock / power / throttle reasons during the timed region (B200_PROFILING.md recipe).  NVML through pynvml when it is
    importable (a query takes ~0.1 ms, so a 100 ms timed region still gets dozens of samples), else one nvidia-smi call per sample."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th, self.via = index, [], False, None, "nvidia-smi"
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.via = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        flags = [bool(r & 0x8), bool(r & 0x40), bool(r & 0x20), bool(r & 0x4)]     # HW slowdown, HW thermal, SW thermal, SW power cap
        return [str(sm), str(mx), f"{pw:.2f}"] + ["Active" if f else "Not Active" for f in flags]

    def _run(self):
        while not self.stop_flag:
            try:
                if self.nvml:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                                   "--format=csv,noheader,nounits"], text=True, timeout=5)
                    self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.005 if self.nvml else 0.1)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for k, nm in enumerate(self.NAMES):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        mx = max((int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=0)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(self.rows), "power_w_max": max(pw) if pw else None, "via": self.via}


def host_threads():
    """threads for the CPU baseline: one per physical core (the naive OpenMP loops lose time on SMT siblings)"""
    try:
        import psutil
        n = psutil.cpu_count(logical=False)
        if n:
            return n
    except Exception:
        pass
    return os.cpu_count() or 1


def set_host_threads(cores):
    """Make the OpenMP runtime(s) of this process use `cores` threads.  The environment variable is only read when a
    libgomp is loaded, and torchrun exports OMP_NUM_THREADS=1 to its workers, so set it through the API as well: on the
    system libgomp the reference build links against and on torch's own copy."""
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(cores))
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(int(cores))
    except Exception:
        pass


def cpu_run(nets, L, R, backend):
    """One 10 s stream through the CPU implementation.  backend: 'openblas' (the reference's real gemm backend, Executable/gemm.c:82-89,
    oracle/_ref/libref_exec_blas.so), 'naive' (the same sources with -DCPU_GEMM=1: the build parity is pinned to) or 'port'
    (oracle/srt_oracle.c when the reference build is absent).  Returns (stems, seconds, kind, description)."""
    from oracle import oracle as O
    cores = host_threads()
    if backend == "openblas":
        r = O.ref_exec(True)
        O.set_blas_threads(cores)
        set_host_threads(cores)
        t0 = time.perf_counter()
        out = r.separate(nets, L, R, T, F, unaffected=0.1)
        return out, time.perf_counter() - t0, "reference+openblas", "reference C sources, gemm.c on cblas_sgemm (OpenBLAS 0.3.15), -O2 -fopenmp"
    if backend == "naive":
        r = O.ref_exec(False)
        set_host_threads(cores)
        t0 = time.perf_counter()
        out = r.separate(nets, L, R, T, F, unaffected=0.1)
        return out, time.perf_counter() - t0, "reference", "reference C sources, -O2 -fopenmp -DCPU_GEMM=1 (naive sgemm)"
    set_host_threads(cores)
    t0 = time.perf_counter()
    out = O.separate(nets, L, R, T, F, unaffected=0.1)
    return out, time.perf_counter() - t0, "port", "oracle port (oracle/srt_oracle.c)"


def cpu_backends():
    from oracle import oracle as O
    b = []
    if O.have_ref_blas():
        b.append("openblas")
    if O.have_ref():
        b.append("naive")
    return b or ["port"]


def reference_arm(args, rank):
    """The reference's own CPU implementation of the path on the box's host cores, bounded sample per step.  The line's value is
    the best backend available (OpenBLAS-linked build of the reference's sources); the naive -DCPU_GEMM build is timed beside it."""
    if rank != 0:
        return
    from oracle import oracle as O
    nets = O.four_stem_weights()
    if args.stems != 4:
        from spleeterrt_b200 import workload as W
        nets = [(np.asarray(c), m) for c, m in W.stem_nets(args.stems)[0]]
    L, R = O.synth_pcm(0, n=N_SAMPLES)
    cores = host_threads()
    backends = cpu_backends()
    res = {}
    for b in backends:
        cpu_run(nets, L, R, b)                                    # warm-up (loads the library, first-touch of its buffers)
        dts = []
        for _ in range(max(args.steps if b == backends[0] else 1, 1)):
            _, dt, kind, desc = cpu_run(nets, L, R, b)
            dts.append(dt)
        res[b] = {"dt": sum(dts) / len(dts), "kind": kind, "desc": desc}
    main_b = backends[0]
    dt = res[main_b]["dt"]
    frames = (4096 * ((N_SAMPLES + 4095) // 4096) + 8192) // 1024
    rtf = SECONDS / dt
    S = len(nets)
    line = {"impl": "reference", "metric": f"realtime_factor_{S}stem_44k1_stereo", "value": rtf, "unit": "x_realtime",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "frames_per_sec": frames / dt,
            "config": {"workload": f"{S}-stem 44.1kHz stereo, T=512 F=1024, one 10 s stream per step (bounded sample of the GPU arm's batch)",
                       "streams_per_step": 1, "seconds_per_stream": SECONDS},
            "cpu_baseline": {"value": rtf, "unit": "x_realtime", "cores": cores, "kind": res[main_b]["kind"],
                             "sample": f"1 stream x 10 s x {S} stems per step; {res[main_b]['desc']}",
                             "other_backends": {b: {"value": SECONDS / res[b]["dt"], "kind": res[b]["kind"], "sample": res[b]["desc"]}
                                                for b in backends[1:]}},
            "e2e": {"value": rtf, "unit": "x_realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_and_parity(nets, pcm, gpu_stems):
    """Bounded CPU samples for the GPU arm's JSON line (one 10 s stream, all stems, all host threads), and - from the naive build,
    the one parity is pinned to - the parity of stream 0 of the timed batch, outside every timed region."""
    L, R = pcm
    S = len(nets)
    cores = host_threads()
    out = {}
    ref = None
    for b in cpu_backends():
        if b == "openblas":
            cpu_run(nets, L, R, b)                               # warm-up: thread pool + page faults of the 300 MB scratch
        stems, dt, kind, desc = cpu_run(nets, L, R, b)
        out[b] = {"value": SECONDS / dt, "unit": "x_realtime", "cores": cores, "kind": kind,
                  "sample": f"1 stream x 10 s x {S} stems, {dt:.2f} s wall; {desc}"}
        if b != "openblas":
            ref = (stems, kind)
    best = max(out.values(), key=lambda r: r["value"])
    base = dict(best)
    base["all_backends"] = out
    parity = None
    if gpu_stems is not None and ref is not None:
        err = [float(np.sqrt(np.mean((gpu_stems[s].astype(np.float64) - ref[0][s]) ** 2))) for s in range(S)]
        lvl = [float(np.sqrt(np.mean(ref[0][s].astype(np.float64) ** 2))) for s in range(S)]
        parity = {"stem_rms_err": err, "stem_rms": lvl, "tolerance": 1e-4, "ok": bool(max(err) < 1e-4), "against": ref[1],
                  "what": "stream 0 of the timed batch (device path), every stem, both channels, all 441000 samples"}
    return base, parity


def numa_pin(local_rank):
    """Bind this rank to the CPUs of its GPU's NUMA node before any pinned allocation (first touch places the pages)."""
    try:
        bdf = subprocess.check_output(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"], text=True).strip().lower()
        bdf = bdf[-12:] if len(bdf) > 12 else bdf
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return {"numa_node": node, "pinned": False}
        cl = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = []
        for part in cl.split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "pinned": True, "cpus": len(cpus)}
    except Exception as e:
        return {"pinned": False, "error": repr(e)[:100]}


def vst_sweep(lib, nets, blocks=(256, 512, 1024, 2048), n_blocks=2000):
    """BASELINE.json configs[4]: the plugin's shape (T=256, F=1536, 4 stems) through the tier-A symbols the JUCE shell calls
    (Spleeter4StemsInit / ProcessSamples, host blocks cut into <= 1024-sample slices like PluginProcessor.cpp:171-181); wall-clock
    latency per host block, p50 / p99 / max over n_blocks blocks per size."""
    lib.Spleeter4StemsInit.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.Spleeter4StemsProcessSamples.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.Spleeter4StemsFree.argtypes = [C.c_void_p]
    coeffs = [np.ascontiguousarray(c, np.float32) for c, _ in nets[:4]]
    cp = (C.c_void_p * 4)(*[c.ctypes.data for c in coeffs])
    obj = C.create_string_buffer(256)
    lib.Spleeter4StemsInit(obj, 1536, 256, cp)
    rng = np.random.default_rng(5)
    out = {}
    for b in blocks:
        x = (0.3 * rng.standard_normal((2, b))).astype(np.float32)
        y = np.zeros((8, b), np.float32)
        lat = np.zeros(n_blocks)
        for k in range(n_blocks):
            t0 = time.perf_counter()
            off = 0
            while off < b:
                m = min(1024, b - off)
                ptr = (C.c_void_p * 8)(*[y[j].ctypes.data + 4 * off for j in range(8)])
                lib.Spleeter4StemsProcessSamples(obj, x[0].ctypes.data + 4 * off, x[1].ctypes.data + 4 * off, m, ptr)
                off += m
            lat[k] = time.perf_counter() - t0
        lat = np.sort(lat[20:]) * 1e6
        out[str(b)] = {"p50_us": float(lat[len(lat) // 2]), "p99_us": float(lat[int(len(lat) * 0.99)]), "max_us": float(lat[-1]),
                       "budget_us": b / 44100.0 * 1e6}
    lib.Spleeter4StemsFree(obj)
    return {"shape": "T=256 F=1536, 4 stems (PluginProcessor.cpp:124)", "blocks_per_size": n_blocks, "through": "Spleeter4StemsProcessSamples (tier A)",
            "latency": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=32, help="10 s stereo streams per GPU per step")
    ap.add_argument("--max-images", type=int, default=0, help="U-Net tiles per pass (0 = min(streams, 32))")
    ap.add_argument("--stems", type=int, default=4, help="nets per stream (4 = the metric's configuration; 5 = BASELINE.json config 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-objects (other configs, probes, sustained run)")
    ap.add_argument("--precision", default="compensated", choices=["compensated", "compensated_bf16", "tf32"],
                    help="srt_config.precision: compensated = TF32 main term + residual term in e5m2 (bf16 in two layers), the default; "
                         "compensated_bf16 = all residuals in bf16; tf32 = single pass")
    ap.add_argument("--dispatch", default="none", choices=["none", "nccl", "peer", "both"],
                    help="rank 0 holds the PCM of ALL streams in its HBM (include/srt_dispatch.h).  nccl: grouped ncclSend / ncclRecv scatter and "
                         "gather; peer: every rank's kernels load / store rank 0's memory over NVLink (CUDA IPC), no copies; both: time both")
    ap.add_argument("--chunks", type=int, default=4, help="pipelined groups of the NCCL dispatcher")
    ap.add_argument("--verbose", action="store_true", help="phase milestones on stderr (every rank)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    t_start = time.perf_counter()

    def log(msg):
        if args.verbose:
            print(f"[bench rank {rank} +{time.perf_counter() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)
    pin_info = numa_pin(local_rank) if world > 1 else {"pinned": False, "why": "single rank"}

    import torch
    import spleeterrt_b200 as srt
    from spleeterrt_b200 import workload as W
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = srt.load_library()
    nets, wdesc = W.stem_nets(args.stems)
    S = len(nets)
    ns = args.streams
    B = args.max_images or min(ns, 32)
    stream = torch.cuda.Stream()            # a real (non-default) stream shared by torch events and the context
    torch.cuda.set_stream(stream)
    sep = srt.Separator(nets, T, F, max_images=B, max_batch_images=ns, device=local_rank, cuda_stream=stream.cuda_stream,
                        precision=args.precision)

    # ---- inputs: stream i of the global batch lives on rank i mod world (dispatch.py) -----------
    from spleeterrt_b200 import dispatch as D
    my_streams = D.stream_ids_for_rank(ns * world, world, rank)
    assert len(my_streams) == ns
    pcm = [W.synth_pcm(sid, n=N_SAMPLES) for sid in my_streams[:4]]
    hin = torch.empty((ns, 2, N_SAMPLES), dtype=torch.float32).pin_memory()
    for i in range(ns):
        hin[i, 0] = torch.from_numpy(pcm[i % len(pcm)][0])
        hin[i, 1] = torch.from_numpy(pcm[i % len(pcm)][1])
    hout = torch.empty((ns, S, 2, N_SAMPLES), dtype=torch.float32).pin_memory()
    din = hin.cuda()
    dout = torch.empty((ns, S, 2, N_SAMPLES), dtype=torch.float32, device="cuda")
    n_arr = (C.c_size_t * ns)(*([N_SAMPLES] * ns))

    def ptrs(t_in, t_out, count=ns):
        pl = (C.c_void_p * count)(*[t_in[i, 0].data_ptr() for i in range(count)])
        pr = (C.c_void_p * count)(*[t_in[i, 1].data_ptr() for i in range(count)])
        po = (C.c_void_p * (count * S * 2))(*[t_out[i, s, c].data_ptr() for i in range(count) for s in range(S) for c in range(2)])
        return pl, pr, po
    dpl, dpr, dpo = ptrs(din, dout)
    hpl, hpr, hpo = ptrs(hin, hout)
    DEPTH = 3                               # batches in flight in the serving loop (= the library's staging slots)
    houts = [hout] + [torch.empty((ns, S, 2, N_SAMPLES), dtype=torch.float32).pin_memory() for _ in range(DEPTH - 1)]
    hpos = [ptrs(hin, h)[2] for h in houts]

    def step_device(s=None):
        (s or sep).separate_raw(dpl, dpr, n_arr, ns, None, dpo, device=True)

    def step_e2e():
        sep.separate_raw(hpl, hpr, n_arr, ns, None, hpo, device=False)

    def barrier():
        torch.cuda.synchronize()
        D.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return D.max_over_ranks(x, device="cuda")

    def timed_device(steps, s=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step_device(s)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    log("context and buffers ready")
    # ---- device-resident throughput ----------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    # (1) per-kernel CUDA events inside the context (same stream), no host sync between steps
    STAGES = list(W.LAYER_FLOP_PER_PIXEL) + ["down1", "up6", "up7", "stft", "istft", "ola"]
    sep.set_timing(True)
    ms_dev = max_over_ranks(timed_device(args.steps))
    layer_ms = {k: sep.timing(k) / args.steps for k in STAGES}
    sep.set_timing(False)
    # (2) the clean pass `value` is taken from (no per-kernel events), clocks sampled during it
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = sep.launch_count()
    ms_step = timed_device(args.steps)
    launches = sep.launch_count() - l0
    barrier()
    ms_step = max_over_ranks(ms_step)
    clocks = sampler.stop() if rank == 0 else None
    gpu0_stems = dout[0].cpu().numpy() if rank == 0 else None           # stream 0 of the timed batch, for the parity object

    log(f"device-resident: {ms_step:.3f} ms/step")
    # ---- end to end through the host-pointer C ABI --------------------------------------------
    # (a) one synchronous call per step: returns after the D2H of every stem
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    ms_e2e_sync = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)

    # (b) the serving loop: srt_separate_batch_async with DEPTH batches in flight.  Every step still uploads its
    # inputs from pinned host memory and downloads every stem; step k+1's upload and step k-1's download overlap
    # step k's kernels.  The clock runs from the first submit to the last wait (pipeline fill and drain included).
    def run_pipelined(steps):
        tickets = []
        for k in range(steps):
            tickets.append(sep.separate_raw_async(hpl, hpr, n_arr, ns, None, hpos[k % DEPTH]))
            if k >= DEPTH - 1:
                sep.wait(tickets[k - (DEPTH - 1)])
        for k in range(max(0, steps - (DEPTH - 1)), steps):
            sep.wait(tickets[k])
    run_pipelined(DEPTH)
    barrier()
    for h in houts:
        h.zero_()
    t0 = time.perf_counter()
    run_pipelined(args.steps)
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    e2e_check = {"slots_identical": bool(torch.equal(houts[0][0], houts[1][0])) if args.steps >= 2 else None,
                 "max_abs_diff_vs_device_path": float((hout[ns - 1] - dout[ns - 1].cpu()).abs().max())}
    if e2e_check["max_abs_diff_vs_device_path"] > 1e-3 or e2e_check["slots_identical"] is False:
        if not os.environ.get("SRT_BENCH_STAGE_SKIPS"):   # tools/up6_sweep.sh: kernels with pipeline stages switched off put out garbage by design
            raise SystemExit(f"bench: host-pointer results are wrong: {e2e_check}")
    # the ceiling of that loop: the same D2H bytes as plain cudaMemcpyAsync, nothing else running (per rank, all ranks at once)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        houts[1].copy_(dout, non_blocking=True)
        ev0.record()
        for _ in range(3):
            houts[1].copy_(dout, non_blocking=True)
        ev1.record()
    torch.cuda.synchronize()
    d2h_ceiling = hout.numel() * 4 * 3 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    d2h_ceiling_min = -max_over_ranks(-d2h_ceiling)
    barrier()
    log(f"e2e: {ms_e2e:.3f} ms/step pipelined, D2H ceiling {d2h_ceiling:.1f} GB/s")

    # ---- BASELINE.json configs[3]: one rank holds every stream, NCCL scatter / gather (include/srt_dispatch.h) -------------------
    dispatch = None
    if args.dispatch != "none" and world > 1:
        def exchange(ident):
            box = [ident]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        nd = srt.NcclDispatcher(world, rank, local_rank, exchange)
        log("NCCL dispatcher up")
        total = ns * world
        n_all = (C.c_size_t * total)(*([N_SAMPLES] * total))
        if rank == 0:
            gin = torch.empty((total, 2, N_SAMPLES), dtype=torch.float32, device="cuda")
            for i in range(total):
                gin[i].copy_(din[i % ns])
            gout = torch.zeros((total, S, 2, N_SAMPLES), dtype=torch.float32, device="cuda")
            gpl, gpr, gpo = ptrs(gin, gout, total)
        else:
            gpl = gpr = gpo = None

        cstream = torch.cuda.ExternalStream(nd.comm_stream())

        def time_dispatch(step_fn):
            for _ in range(2):
                step_fn()
                nd.wait()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_fn()
            e1.record(cstream)
            nd.wait()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3 / args.steps
            return max_over_ranks(e0.elapsed_time(e1) / args.steps), max_over_ranks(wall)
        dispatch = {"streams_total": total, "chunks": args.chunks,
                    "bytes_in_per_step": int((world - 1) * ns * 2 * N_SAMPLES * 4), "bytes_out_per_step": int((world - 1) * ns * S * 2 * N_SAMPLES * 4),
                    "note": "rank 0 holds the PCM of all streams in its HBM and receives all stems there; device-timed from rank 0's first enqueue "
                            "to the completion of the last transfer, max over ranks"}
        if args.dispatch in ("nccl", "both"):
            ms_disp, wall = time_dispatch(lambda: nd.separate_device(sep, 0, gpl, gpr, n_all, total, None, gpo, chunks=args.chunks))
            ok = None
            if rank == 0:
                # every stream of the global batch is one of the four synthetic streams of rank 0's own batch: compare with the local result
                ok = float(max((gout[i] - dout[i % ns]).abs().max() for i in range(0, total, max(1, total // 16))))
            dispatch["nccl_send_recv"] = {"mode": "grouped ncclSend/ncclRecv scatter, per-rank srt_separate_device, gather into rank 0's HBM, pipelined in chunks",
                                          "ms_per_step": ms_disp, "host_wall_ms_per_step": wall, "value": SECONDS * total / (ms_disp * 1e-3),
                                          "unit": "x_realtime", "max_abs_diff_vs_local_path": ok}
            log(f"dispatch (nccl send/recv): {ms_disp:.3f} ms/step")
        if args.dispatch in ("peer", "both"):
            p_in, p_out, in_off, out_off, fin, fout = nd.peer_buffers(0, [N_SAMPLES] * total, S)
            if rank == 0:
                # the batch lives in root's peer-visible buffer in the dispatcher's layout: [stream][L | R], results [stream][stem][channel]
                npad = (N_SAMPLES + 3) & ~3
                assert npad == N_SAMPLES
                cudart = C.CDLL("libcudart.so.12")               # torch's runtime, already in the process
                cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
                rc = cudart.cudaMemcpy(p_in, gin.data_ptr(), total * 2 * N_SAMPLES * 4, 3)      # cudaMemcpyDeviceToDevice = 3
                assert int(rc) == 0, rc
            barrier()
            def check_peer():
                if rank != 0:
                    return None
                chk = torch.empty((S, 2, N_SAMPLES), dtype=torch.float32, device="cuda")
                worst = 0.0
                for i in range(0, total, max(1, total // 16)):
                    rc = cudart.cudaMemcpy(chk.data_ptr(), p_out + out_off[i] * 4, S * 2 * N_SAMPLES * 4, 3)
                    assert int(rc) == 0, rc
                    worst = max(worst, float((chk - dout[i % ns]).abs().max()))
                return worst
            # (a) direct: one srt_separate_device call per rank whose overlap-add kernel stores into rank 0's memory
            ms_peer, wall = time_dispatch(lambda: nd.separate_peer(sep, 0, n_all, total, None, chunks=1))
            dispatch["peer_memory_direct"] = {"mode": "no copies: every rank's STFT kernel loads the PCM from rank 0's HBM and its overlap-add kernel stores the stems "
                                                      "into rank 0's HBM over NVLink (CUDA IPC mappings); NCCL carries the handles and two one-word all-reduces per step",
                                              "ms_per_step": ms_peer, "host_wall_ms_per_step": wall, "value": SECONDS * total / (ms_peer * 1e-3),
                                              "unit": "x_realtime", "max_abs_diff_vs_local_path": check_peer()}
            log(f"dispatch (peer memory, direct stores): {ms_peer:.3f} ms/step")
            if rank == 0:
                rc = cudart.cudaMemset(C.c_void_p(p_out), 0, C.c_size_t(fout * 4))
            barrier()
            # (b) pipelined: PCM still loaded over NVLink by the kernels; a chunk's stems leave by copy engine while the next chunk computes
            ms_peer, wall = time_dispatch(lambda: nd.separate_peer(sep, 0, n_all, total, None, chunks=args.chunks))
            dispatch["peer_memory"] = {"mode": "PCM loaded from rank 0's HBM by the STFT kernel over NVLink; the stems of a chunk go to rank 0's HBM with one "
                                               "copy-engine transfer per stream (cudaMemcpyAsync into the IPC mapping) while the next chunk computes",
                                       "ms_per_step": ms_peer, "host_wall_ms_per_step": wall, "value": SECONDS * total / (ms_peer * 1e-3),
                                       "unit": "x_realtime", "max_abs_diff_vs_local_path": check_peer()}
            log(f"dispatch (peer memory, pipelined copies): {ms_peer:.3f} ms/step")
        best = max((v for k, v in dispatch.items() if isinstance(v, dict)), key=lambda v: v["value"])
        dispatch["ms_per_step"], dispatch["value"], dispatch["unit"] = best["ms_per_step"], best["value"], "x_realtime"
        barrier()
        nd.close()
        if rank == 0:
            del gin, gout

    barrier()
    if rank == 0:
        pk = peaks()
        frames = W.padded_frames(N_SAMPLES)
        tiles = (frames + T - 1) // T
        audio_s = SECONDS * ns * world
        P = T * F
        units = ns * tiles * S                                            # stem-tiles per step per GPU
        # ---- measured roofline denominators, right after the timed region (same clocks / thermal state) -------------
        tf32_peak = bf16_peak = hbm_probe = None
        probe_error = None
        try:
            tf32_peak = srt.probe_tensor_peak("tf32", 0.15, local_rank)
            bf16_peak = srt.probe_tensor_peak("bf16", 0.15, local_rank)
            hbm_probe = srt.probe_copy_bandwidth(1 << 30, local_rank)
        except Exception as e:
            probe_error = repr(e)
        if not tf32_peak:
            tf32_peak, peak_src = pk["bf16_tflops"] / 2.0, f"{pk['source']} bf16 burst {pk['bf16_tflops']} TF/s / 2"
        else:
            peak_src = ("srt_probe_tensor_peak: every SM issuing N=256 kind::tf32 MMAs from shared memory for 0.15 s right after the timed region "
                        f"(kind::f16/bf16: {bf16_peak:.0f} TF/s; {pk['source']} cuBLAS bf16 burst {pk['bf16_tflops']} / sustained {pk['bf16_tflops_sustained']})")
        hbm_peak = pk["hbm_gbs"]
        comp = args.precision != "tf32"
        tc_flop = W.FLOP_PER_PIXEL_TC * P * units
        tc_ms = sum(layer_ms[k] for k in W.LAYER_FLOP_PER_PIXEL)
        ach = tc_flop / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        # executed tensor work: the compensation term repeats every layer's contraction on bf16 operands (half the MMAs, bf16 rate)
        # (the e5m2 blocks of the default mode run at twice the bf16 rate: their layers hold 79 % of the FLOPs)
        lo_rate = (bf16_peak or 2 * tf32_peak) * (1.0 if args.precision == "compensated_bf16" else 1.0 / (0.21 + 0.79 / 2))
        t_at_peak = tc_flop / (tf32_peak * 1e12) + (tc_flop / (lo_rate * 1e12) if comp else 0.0)
        traffic, traffic_src, dram = None, None, {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            key = f"{args.precision}_{ns}streams_{S}stems"
            if key in tj:
                traffic, traffic_src, dram = tj[key]["tc_layers_dram_bytes_per_step"], tj[key]["source"], tj[key].get("per_kernel_dram_bytes", {})
        # ---- every kernel against its roofline (algorithmic work: SURVEY.md 8d) ------------------------------------------------------------------
        rows = []
        kern = {"down2": "conv_rp<32,3>", "down3": "conv_rp<64,3>", "down4": "conv_tc<128,1>", "down5": "conv_tc<256,2>", "down6": "conv_tc<256,2>",
                "up1": "conv_tc<256,1>", "up2": "conv_tc<128,2>", "up3": "conv_tc<64,1>", "up4": "conv_rp<128,2>", "up5": "conv_rp<64,3>"}
        for k, fpp in W.LAYER_FLOP_PER_PIXEL.items():
            fl = fpp * P * units
            a = fl / (layer_ms[k] * 1e-3) / 1e12 if layer_ms[k] > 0 else None
            rows.append({"stage": k, "kernel": kern[k], "ms": layer_ms[k], "work": fl, "work_unit": "FLOP", "achieved": a, "unit": "TFLOP/s",
                         "bound": "tensor", "peak": tf32_peak, "frac": a / tf32_peak if a else None, "dram_bytes": dram.get(k)})
        hb = {"stft": ("stft_kernel", 2 * 24.4e3 * frames * ns),
              "down1": ("conv_rp<16*stems,4> (8-channel k-blocks)", (2 + 4 + 4) * P * 4.0 * units),
              "up6+up7": ("up6_tc_kernel + up7_kernel", (8 + 2) * P * 4.0 * units),
              "istft": ("istft_ola kernel (mask + iFFT + window + overlap-add)", (2 * 16.4e3 + S * 2 * (F * 4 + 4096)) * frames * ns)}
        for k, (kn, by) in hb.items():
            ms = layer_ms["up6"] + layer_ms["up7"] if k == "up6+up7" else layer_ms[k] + (layer_ms["ola"] if k == "istft" else 0.0)
            a = by / (ms * 1e-3) / 1e9 if ms > 0 else None
            rows.append({"stage": k, "kernel": kn, "ms": ms, "work": by, "work_unit": "B", "achieved": a, "unit": "GB/s", "bound": "hbm",
                         "peak": hbm_peak, "frac": a / hbm_peak if a else None, "dram_bytes": dram.get(k)})
        line = {
            "metric": f"realtime_factor_{S}stem_44k1_stereo", "value": audio_s / (ms_step * 1e-3), "unit": "x_realtime",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"compensated": "tf32+e5m2/bf16", "compensated_bf16": "tf32+bf16", "tf32": "tf32"}[args.precision], "data": "synthetic",
            "frames_per_sec": frames * ns * world / (ms_step * 1e-3),
            "config": {"workload": f"{S}-stem 44.1 kHz stereo, {ns} x 10 s streams per GPU per step, T=512 F=1024 (1 tile/stream), "
                                   f"STFT + {S} U-Nets + mask + iSTFT/OLA", "streams_per_gpu": ns, "time_step": T, "bin_limit": F,
                       "stems": S, "unet_tiles_per_pass": B,
                       "precision": args.precision + {"compensated": " (default: tf32(a) x w + residual (a - tf32(a)) x w in e5m2, bf16 where the residual tensor has 64 channels; fp32 accumulate)",
                                                      "compensated_bf16": " (tf32(a) x w + bf16(a - tf32(a)) x bf16(w), fp32 accumulate)",
                                                      "tf32": " (single-pass TF32 operands, fp32 accumulate)"}[args.precision],
                       "weights": wdesc, "l2": "per-step working set (activations, > 9 GB) >> 126 MB L2; no explicit flush",
                       "parallelism": f"streams sharded over {world} GPU(s), no data-path collective", "cpu_pinning": pin_info},
            "e2e": {"value": audio_s / (ms_e2e * 1e-3), "unit": "x_realtime", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(hin.numel() * 4), "d2h_bytes_per_step": int(hout.numel() * 4),
                    "mode": f"srt_separate_batch_async + srt_batch_wait, {DEPTH} batches in flight, pinned host buffers, host wall clock from first submit to last wait",
                    "sync_call": {"value": audio_s / (ms_e2e_sync * 1e-3), "ms_per_step": ms_e2e_sync,
                                  "mode": "one blocking srt_separate_batch per step"},
                    "pcie_d2h_gbs": hout.numel() * 4 / (ms_e2e * 1e-3) / 1e9,
                    "pcie_ceiling_gbs": d2h_ceiling_min,
                    "pcie_ceiling_note": "plain cudaMemcpyAsync of the same bytes of stems from HBM to the same pinned buffers, all ranks at once, slowest rank",
                    "frac_of_pcie_ceiling": (hout.numel() * 4 / (ms_e2e * 1e-3) / 1e9) / d2h_ceiling_min if d2h_ceiling_min else None,
                    "check": e2e_check},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "conv_tc_kernel / conv_rp_kernel (tcgen05 implicit-GEMM conv / tconv, the 10 tensor-core layers)", "bound": "tensor",
                         "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak if tf32_peak else None,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "ms_per_step": tc_ms, "flop_per_step": tc_flop,
                         "algorithmic": "22400 FLOP per mask pixel per stem x T*F x stem-tiles (SURVEY 8d); the compensation term's MMAs are NOT counted",
                         "frac_executed": t_at_peak / (tc_ms * 1e-3) if tc_ms > 0 else None,
                         "frac_executed_note": "time the MMAs actually issued would take at the measured pipe rates (TF32 main term + bf16 compensation term) / measured time"},
            "roofline_all": rows,
            "peaks": {"tf32_tflops_probe": tf32_peak, "bf16_tflops_probe": bf16_peak, "hbm_gbs_probe": hbm_probe, "hbm_gbs": hbm_peak,
                      "bf16_tflops_cublas_burst": pk["bf16_tflops"], "bf16_tflops_cublas_sustained": pk["bf16_tflops_sustained"], "source": pk["source"],
                      "probe_error": probe_error},
            "stage_ms": {k: layer_ms[k] for k in ("down1", "up6", "up7", "stft", "istft", "ola")},
            "timed_with_layer_events_ms": ms_dev,
        }
        if dispatch:
            line["dispatch"] = dispatch

        if not args.no_extras and world == 1:
            # ---- single stream (BASELINE.json configs[1]): one 10 s stereo stream, all stems, latency -----------------------
            sep1 = srt.Separator(nets, T, F, max_images=1, max_batch_images=1, device=local_rank, cuda_stream=stream.cuda_stream,
                                 precision=args.precision)
            n1 = (C.c_size_t * 1)(N_SAMPLES)
            d1l, d1r = (C.c_void_p * 1)(din[0, 0].data_ptr()), (C.c_void_p * 1)(din[0, 1].data_ptr())
            d1o = (C.c_void_p * (S * 2))(*[dout[0, s, c].data_ptr() for s in range(S) for c in range(2)])
            h1l, h1r = (C.c_void_p * 1)(hin[0, 0].data_ptr()), (C.c_void_p * 1)(hin[0, 1].data_ptr())
            h1o = (C.c_void_p * (S * 2))(*[hout[0, s, c].data_ptr() for s in range(S) for c in range(2)])
            for _ in range(3):
                sep1.separate_raw(d1l, d1r, n1, 1, None, d1o, device=True)
            torch.cuda.synchronize()
            reps = 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                sep1.separate_raw(d1l, d1r, n1, 1, None, d1o, device=True)
            e1.record(stream)
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / reps
            for _ in range(2):
                sep1.separate_raw(h1l, h1r, n1, 1, None, h1o, device=False)
            t0 = time.perf_counter()
            for _ in range(reps):
                sep1.separate_raw(h1l, h1r, n1, 1, None, h1o, device=False)
            ms1e = (time.perf_counter() - t0) * 1e3 / reps
            line["single_stream"] = {"ms_device": ms1, "x_realtime_device": SECONDS / (ms1 * 1e-3), "ms_e2e": ms1e,
                                     "x_realtime_e2e": SECONDS / (ms1e * 1e-3), "note": f"BASELINE.json configs[1]: one 10 s stereo stream, {S} stems, batch of 1 tile"}
            sep1.close()
            # ---- the other precision modes, same batch ---------------------------------------------------------------------------------------
            sep.close()
            sep = None
            modes = {args.precision: {"ms_per_step": ms_step, "value": audio_s / (ms_step * 1e-3)}}
            for other in ("compensated", "compensated_bf16", "tf32"):
                if other == args.precision:
                    continue
                sep2 = srt.Separator(nets, T, F, max_images=B, max_batch_images=ns, device=local_rank, cuda_stream=stream.cuda_stream, precision=other)
                for _ in range(3):
                    step_device(sep2)
                ms_other = timed_device(args.steps, sep2)
                modes[other] = {"ms_per_step": ms_other, "value": audio_s / (ms_other * 1e-3)}
                sep2.close()
            modes["note"] = "same batch, device-resident; parity of every mode: tests/test_gpu_headline.py"
            line["precision_modes"] = modes
            # ---- sustained: >= 5 s of back-to-back steps in the default configuration (power cap, clocks) -----------------------------------
            sep3 = srt.Separator(nets, T, F, max_images=B, max_batch_images=ns, device=local_rank, cuda_stream=stream.cuda_stream,
                                 precision=args.precision)
            for _ in range(3):
                step_device(sep3)
            torch.cuda.synchronize()
            n_sus = int(5500.0 / ms_step) + 1
            smp = ClockSampler(local_rank)
            smp.start()
            ms_sus = timed_device(n_sus, sep3)
            ck = smp.stop()
            line["sustained"] = {"steps": n_sus, "seconds": n_sus * ms_sus * 1e-3, "ms_per_step": ms_sus, "value": audio_s / (ms_sus * 1e-3),
                                 "clocks": ck}
            try:
                line["sustained"]["tf32_tflops_probe_2s"] = srt.probe_tensor_peak("tf32", 2.0, local_rank)
            except Exception as e:
                line["sustained"]["probe_error"] = repr(e)[:200]
            sep3.close()
            # ---- BASELINE.json configs[2]: 256 concurrent streams on one GPU (64 tiles per U-Net pass) -------------------------------------
            try:
                ns3 = 256
                din3 = din.repeat(ns3 // ns, 1, 1) if ns3 % ns == 0 else din[:1].repeat(ns3, 1, 1)
                dout3 = torch.empty((ns3, S, 2, N_SAMPLES), dtype=torch.float32, device="cuda")
                sep4 = srt.Separator(nets, T, F, max_images=64, max_batch_images=ns3, device=local_rank, cuda_stream=stream.cuda_stream,
                                     precision=args.precision)
                p3l, p3r, p3o = ptrs(din3, dout3, ns3)
                n3 = (C.c_size_t * ns3)(*([N_SAMPLES] * ns3))
                for _ in range(2):
                    sep4.separate_raw(p3l, p3r, n3, ns3, None, p3o, device=True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(3):
                    sep4.separate_raw(p3l, p3r, n3, ns3, None, p3o, device=True)
                e1.record(stream)
                torch.cuda.synchronize()
                ms3 = e0.elapsed_time(e1) / 3
                same = bool(torch.equal(dout3[ns3 - ns], dout[0])) if ns3 % ns == 0 else None
                line["streams_256"] = {"ms_per_step": ms3, "value": SECONDS * ns3 / (ms3 * 1e-3), "unit": "x_realtime", "steps": 3,
                                       "frames_per_sec": frames * ns3 / (ms3 * 1e-3), "stems_bit_identical_to_32_stream_batch": same,
                                       "note": "BASELINE.json configs[2]: 256 x 10 s streams resident in HBM, 4 U-Net passes of 64 tiles"}
                sep4.close()
                del din3, dout3
            except Exception as e:
                line["streams_256"] = {"error": repr(e)[:200]}
            # ---- BASELINE.json configs[4]: VST block-size sweep -----------------------------------------------------------------------------------
            try:
                line["vst_block_sweep"] = vst_sweep(lib, nets)
            except Exception as e:
                line["vst_block_sweep"] = {"error": repr(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], line["parity"] = cpu_baseline_and_parity([(np.asarray(c), m) for c, m in nets], pcm[0],
                                                                               gpu0_stems if my_streams[0] == 0 else None)
            except Exception as e:  # the checker being unavailable must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "x_realtime", "cores": os.cpu_count(), "kind": "unavailable",
                                        "sample": repr(e)}
        print(json.dumps(line), flush=True)
    if sep is not None:
        sep.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
