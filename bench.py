#!/usr/bin/env python3
"""bench.py — 4-stem 44.1 kHz stereo separation throughput on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own CPU path (oracle/_ref)

A step = one pass of the whole hot path (STFT framer -> 4 U-Nets -> mask*spectrum -> iSTFT/OLA)
over a batch of `--streams` synthetic 10 s stereo streams per GPU (T=512, F=1024: BASELINE.json
configs[1] shape, batched as configs[2] does).  `value` is device-resident throughput (inputs in
HBM when the clock starts), `e2e` goes through the host-pointer C-ABI call with pinned host
buffers, H2D and D2H inside the timed region.
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
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                               "--format=csv,noheader,nounits"], text=True, timeout=5)
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.15)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        mx = max((int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=0)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


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
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(cores))
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(int(cores))
    except Exception:
        pass


def reference_arm(args, rank):
    """The reference's own CPU implementation of the path (oracle/_ref build of the reference's C
    sources, naive CPU_GEMM backend), all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle as O
    have_ref = O.have_ref()
    nets = O.four_stem_weights()
    L, R = O.synth_pcm(0, n=N_SAMPLES)
    cores = host_threads()
    if have_ref:
        O.ref_exec()                  # load the reference build (and its libgomp) before setting the thread count
    set_host_threads(cores)

    def step():
        if have_ref:
            return O.ref_exec().separate(nets, L, R, T, F, unaffected=0.1)
        return O.separate(nets, L, R, T, F, unaffected=0.1)
    for _ in range(max(min(args.warmup, 1), 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    frames = N_SAMPLES and (4096 * ((N_SAMPLES + 4095) // 4096) + 8192) // 1024
    rtf = SECONDS / dt
    kind = "reference" if have_ref else "port"
    line = {"impl": "reference", "metric": "realtime_factor_4stem_44k1_stereo", "value": rtf, "unit": "x_realtime",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "frames_per_sec": frames / dt,
            "config": {"workload": "4-stem 44.1kHz stereo, T=512 F=1024, one 10 s stream per step (bounded sample of the GPU arm's batch)",
                       "streams_per_step": 1, "seconds_per_stream": SECONDS},
            "cpu_baseline": {"value": rtf, "unit": "x_realtime", "cores": cores, "kind": kind,
                             "sample": "1 stream x 10 s x 4 stems per step; reference C sources built -O2 -fopenmp -DCPU_GEMM=1 (naive sgemm)"},
            "e2e": {"value": rtf, "unit": "x_realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(nets=None, pcm=None, gpu_stems=None):
    """Bounded CPU sample for the GPU arm's JSON line: one 10 s stream, 4 stems, all host threads.  The same run is the
    parity check of the bench's own configuration: `gpu_stems` (stream 0 of the GPU batch, float32[S][2][n]) against
    the CPU result, RMS per stem (outside every timed region)."""
    from oracle import oracle as O
    nets = nets if nets is not None else O.four_stem_weights()
    L, R = pcm if pcm is not None else O.synth_pcm(0, n=N_SAMPLES)
    cores = host_threads()
    have_ref = O.have_ref()
    if have_ref:
        O.ref_exec()
    set_host_threads(cores)
    t0 = time.perf_counter()
    if have_ref:
        ref = O.ref_exec().separate(nets, L, R, T, F, unaffected=0.1)
    else:
        ref = O.separate(nets, L, R, T, F, unaffected=0.1)
    dt = time.perf_counter() - t0
    S = len(nets)
    base = {"value": SECONDS / dt, "unit": "x_realtime", "cores": cores, "kind": "reference" if have_ref else "port",
            "sample": f"1 stream x 10 s x {S} stems, {dt:.2f} s wall; reference C sources (oracle/_ref, -O2 -fopenmp -DCPU_GEMM=1 naive sgemm)"
            if have_ref else f"1 stream x 10 s x {S} stems, {dt:.2f} s wall; oracle port (oracle/srt_oracle.c)"}
    parity = None
    if gpu_stems is not None:
        err = [float(np.sqrt(np.mean((gpu_stems[s].astype(np.float64) - ref[s]) ** 2))) for s in range(S)]
        lvl = [float(np.sqrt(np.mean(ref[s].astype(np.float64) ** 2))) for s in range(S)]
        parity = {"stem_rms_err": err, "stem_rms": lvl, "tolerance": 1e-4, "ok": bool(max(err) < 1e-4),
                  "against": base["kind"], "what": "stream 0 of the timed batch, every stem, both channels, all 441000 samples"}
    return base, parity


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=32, help="10 s stereo streams per GPU per step")
    ap.add_argument("--max-images", type=int, default=0, help="U-Net tiles per pass (0 = streams)")
    ap.add_argument("--stems", type=int, default=4, help="nets per stream (4 = the metric's configuration; 5 = BASELINE.json config 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="compensated", choices=["compensated", "tf32"],
                    help="srt_config.precision: compensated = TF32 main term + bf16 residual term (default, fp32-grade), tf32 = single pass")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    args.warmup = max(args.warmup, 3)

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
    srt.load_library()
    nets, wdesc = W.stem_nets(args.stems)
    S = len(nets)
    ns = args.streams
    B = args.max_images or ns
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

    def ptrs(t_in, t_out):
        pl = (C.c_void_p * ns)(*[t_in[i, 0].data_ptr() for i in range(ns)])
        pr = (C.c_void_p * ns)(*[t_in[i, 1].data_ptr() for i in range(ns)])
        po = (C.c_void_p * (ns * S * 2))(*[t_out[i, s, c].data_ptr() for i in range(ns) for s in range(S) for c in range(2)])
        return pl, pr, po
    dpl, dpr, dpo = ptrs(din, dout)
    hpl, hpr, hpo = ptrs(hin, hout)
    DEPTH = 3                               # batches in flight in the serving loop (= the library's staging slots)
    houts = [hout] + [torch.empty((ns, S, 2, N_SAMPLES), dtype=torch.float32).pin_memory() for _ in range(DEPTH - 1)]
    hpos = [ptrs(hin, h)[2] for h in houts]

    def step_device():
        sep.separate_raw(dpl, dpr, n_arr, ns, None, dpo, device=True)

    def step_e2e():
        sep.separate_raw(hpl, hpr, n_arr, ns, None, hpo, device=False)

    def barrier():
        torch.cuda.synchronize()
        D.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return D.max_over_ranks(x, device="cuda")

    # ---- device-resident throughput ----------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # per-kernel CUDA events inside the context (same stream), no host sync between steps
    sep.set_timing(True)
    l0 = sep.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    launches = sep.launch_count() - l0
    ms_dev = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    layer_ms = {k: sep.timing(k) for k in list(W.LAYER_FLOP_PER_PIXEL) + ["down1", "up6", "up7", "stft", "istft", "ola"]}
    sep.set_timing(False)

    # ---- untimed-layer pass for a clean `value` (no per-layer syncs inside) ---------------------
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    clocks = sampler.stop() if rank == 0 else None

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
        raise SystemExit(f"bench: host-pointer results are wrong: {e2e_check}")

    # ---- single stream (BASELINE.json configs[1]): one 10 s stereo stream, 4 stems, latency --------------
    single = None
    if rank == 0:
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
        single = {"ms_device": ms1, "x_realtime_device": SECONDS / (ms1 * 1e-3), "ms_e2e": ms1e,
                  "x_realtime_e2e": SECONDS / (ms1e * 1e-3), "note": "one 10 s stereo stream, 4 stems, batch of 1 tile"}
        sep1.close()

    if rank == 0:
        pk = peaks()
        frames = W.padded_frames(N_SAMPLES)
        tiles = (frames + T - 1) // T
        audio_s = SECONDS * ns * world
        P = T * F
        n_tc_units = ns * tiles * S                               # stem-tiles per step per GPU
        tc_flop = W.FLOP_PER_PIXEL_TC * P * n_tc_units
        tc_ms = sum(layer_ms[k] for k in W.LAYER_FLOP_PER_PIXEL) / args.steps
        tf32_peak = pk["bf16_tflops_sustained"] / 2.0
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and ns == 32:      # the capture was taken at the default batch
            tj = json.load(open(tpath))
            traffic, traffic_src = tj["tc_layers"]["dram_bytes_per_step"], tj["source"]
        ach = tc_flop / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        per_layer = {k: {"ms": layer_ms[k] / args.steps,
                         "tflops": (W.LAYER_FLOP_PER_PIXEL[k] * P * n_tc_units) / (layer_ms[k] / args.steps * 1e-3) / 1e12
                         if layer_ms[k] > 0 else None} for k in W.LAYER_FLOP_PER_PIXEL}
        # HBM-bound stages: algorithmic bytes per hop-frame
        stft_bytes = 2 * 24.4e3 * frames * ns
        istft_bytes = 96.8e3 * frames * ns
        other = {k: layer_ms[k] / args.steps for k in ("down1", "up6", "up7", "stft", "istft", "ola")}
        line = {
            "metric": f"realtime_factor_{S}stem_44k1_stereo", "value": audio_s / (ms_step * 1e-3), "unit": "x_realtime",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32+bf16" if args.precision == "compensated" else "tf32", "data": "synthetic",
            "frames_per_sec": frames * ns * world / (ms_step * 1e-3),
            "config": {"workload": f"{S}-stem 44.1 kHz stereo, {ns} x 10 s streams per GPU per step, T=512 F=1024 (1 tile/stream), "
                                   f"STFT + {S} U-Nets + mask + iSTFT/OLA", "streams_per_gpu": ns, "time_step": T, "bin_limit": F,
                       "stems": S, "precision": args.precision + (" (tf32(a) x w + bf16(a - tf32(a)) x bf16(w), fp32 accumulate)" if args.precision == "compensated"
                                                                    else " (single-pass TF32 operands, fp32 accumulate)"), "weights": wdesc, "l2": "per-step working set (activations) >> 126 MB L2; no explicit flush",
                       "parallelism": f"streams sharded over {world} GPU(s), no data-path collective"},
            "e2e": {"value": audio_s / (ms_e2e * 1e-3), "unit": "x_realtime", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(hin.numel() * 4), "d2h_bytes_per_step": int(hout.numel() * 4),
                    "mode": f"srt_separate_batch_async + srt_batch_wait, {DEPTH} batches in flight, pinned host buffers, host wall clock from first submit to last wait",
                    "sync_call": {"value": audio_s / (ms_e2e_sync * 1e-3), "ms_per_step": ms_e2e_sync,
                                  "mode": "one blocking srt_separate_batch per step"},
                    "pcie_d2h_gbs": hout.numel() * 4 / (ms_e2e * 1e-3) / 1e9, "check": e2e_check},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv/tconv, 10 layers)", "bound": "tensor",
                         "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak if tf32_peak else None,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": f"{pk['source']} bf16 sustained {pk['bf16_tflops_sustained']} TF/s / 2 (TF32 operands)",
                         "ms_per_step": tc_ms, "flop_per_step": tc_flop, "per_layer": per_layer},
            "stage_ms": other,
            "hbm_stages": {"stft_gbs": stft_bytes / (other["stft"] * 1e-3) / 1e9 if other["stft"] > 0 else None,
                           "istft_ola_gbs": istft_bytes / ((other["istft"] + other["ola"]) * 1e-3) / 1e9 if other["istft"] > 0 else None,
                           "peak_gbs": pk["hbm_gbs"]},
            "timed_with_layer_events_ms": ms_dev,
            "single_stream": single,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                gpu0 = dout[0].cpu().numpy() if my_streams[0] == 0 else None      # stream 0 of the timed batch (device path)
                line["cpu_baseline"], line["parity"] = cpu_baseline_sample([(np.asarray(c), m) for c, m in nets], pcm[0], gpu0)
            except Exception as e:  # the checker being unavailable must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "x_realtime", "cores": os.cpu_count(), "kind": "unavailable",
                                        "sample": repr(e)}
        print(json.dumps(line), flush=True)
    sep.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
