"""End-to-end (host-pointer API) timing probe: blocking vs. three-in-flight asynchronous submission for several
SRT_E2E_GROUPS settings.  Run on the GPU box: python tools/e2e_probe.py"""
import ctypes as C, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spleeterrt_b200 as srt
from spleeterrt_b200 import workload as W

T, F, N, ns = 512, 1024, 441000, 32
nets, _ = W.four_stem_nets()
S = len(nets)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sep = srt.Separator(nets, T, F, max_images=ns, max_batch_images=ns, device=0, cuda_stream=stream.cuda_stream)
hin = torch.randn((ns, 2, N), dtype=torch.float32).mul_(0.1).pin_memory()
DEPTH = 3
houts = [torch.empty((ns, S, 2, N), dtype=torch.float32).pin_memory() for _ in range(DEPTH)]
n_arr = (C.c_size_t * ns)(*([N] * ns))
pl = (C.c_void_p * ns)(*[hin[i, 0].data_ptr() for i in range(ns)])
pr = (C.c_void_p * ns)(*[hin[i, 1].data_ptr() for i in range(ns)])
pos = [(C.c_void_p * (ns * S * 2))(*[h[i, s, c].data_ptr() for i in range(ns) for s in range(S) for c in range(2)]) for h in houts]

def blocking(steps):
    for k in range(steps):
        sep.separate_raw(pl, pr, n_arr, ns, None, pos[k % DEPTH])

def pipelined(steps):
    t = []
    for k in range(steps):
        t.append(sep.separate_raw_async(pl, pr, n_arr, ns, None, pos[k % DEPTH]))
        if k >= DEPTH - 1:
            sep.wait(t[k - (DEPTH - 1)])
    for k in range(max(0, steps - (DEPTH - 1)), steps):
        sep.wait(t[k])

def submit_only_cost(steps):
    t0 = time.perf_counter()
    tk = sep.separate_raw_async(pl, pr, n_arr, ns, None, pos[0])
    dt = time.perf_counter() - t0
    sep.wait(tk)
    return dt * 1e3

res = {}
for groups in (1, 2, 4, 8):
    os.environ["SRT_E2E_GROUPS"] = str(groups)
    for name, fn in (("blocking", blocking), ("pipelined", pipelined)):
        fn(3); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(10); torch.cuda.synchronize()
        res[f"{name}_g{groups}_ms"] = (time.perf_counter() - t0) * 1e3 / 10
    res[f"submit_host_ms_g{groups}"] = submit_only_cost(1)
print(json.dumps(res, indent=1))

# ---- does the download slow down while the kernels run?  (device-resident steps on the context's stream,
# one 451 MB pinned D2H per step on a second stream, timed with events on the copy stream)
os.environ["SRT_E2E_GROUPS"] = "1"
din = hin.cuda()
dout = torch.empty((ns, S, 2, N), dtype=torch.float32, device="cuda")
dpl = (C.c_void_p * ns)(*[din[i, 0].data_ptr() for i in range(ns)])
dpr = (C.c_void_p * ns)(*[din[i, 1].data_ptr() for i in range(ns)])
dpo = (C.c_void_p * (ns * S * 2))(*[dout[i, s, c].data_ptr() for i in range(ns) for s in range(S) for c in range(2)])
s_copy = torch.cuda.Stream()
def d2h_rate(with_compute, reps=6):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    if with_compute:
        for _ in range(reps):
            sep.separate_raw(dpl, dpr, n_arr, ns, None, dpo, device=True)
    c1.record(stream)
    with torch.cuda.stream(s_copy):
        e0.record(s_copy)
        for k in range(reps):
            houts[k % DEPTH].copy_(dout, non_blocking=True)
        e1.record(s_copy)
    torch.cuda.synchronize()
    return {"d2h_gbs": dout.numel() * 4 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, "compute_ms_per_step": c0.elapsed_time(c1) / reps}
print(json.dumps({"d2h_alone": d2h_rate(False), "d2h_with_kernels": d2h_rate(True)}))
