"""Per-call wall time of the blocking host-pointer call for one 10 s stream (debug probe)."""
import ctypes as C, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spleeterrt_b200 as srt
from spleeterrt_b200 import workload as W
T, F, N = 512, 1024, 441000
nets, _ = W.four_stem_nets(); S = len(nets)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sep1 = srt.Separator(nets, T, F, max_images=1, max_batch_images=1, device=0, cuda_stream=stream.cuda_stream)
hin = torch.randn((2, N)).mul_(0.1).pin_memory(); hout = torch.empty((S, 2, N)).pin_memory()
n1 = (C.c_size_t * 1)(N)
l, r = (C.c_void_p * 1)(hin[0].data_ptr()), (C.c_void_p * 1)(hin[1].data_ptr())
o = (C.c_void_p * (S * 2))(*[hout[s, c].data_ptr() for s in range(S) for c in range(2)])
ts = []
for k in range(12):
    t0 = time.perf_counter(); sep1.separate_raw(l, r, n1, 1, None, o); ts.append(round((time.perf_counter() - t0) * 1e3, 3))
print(json.dumps({"blocking_ms": ts}))
ts = []
for k in range(6):
    t0 = time.perf_counter(); tk = sep1.separate_raw_async(l, r, n1, 1, None, o); t1 = time.perf_counter(); sep1.wait(tk); t2 = time.perf_counter()
    ts.append([round((t1 - t0) * 1e3, 3), round((t2 - t1) * 1e3, 3)])
print(json.dumps({"async_submit_wait_ms": ts}))
