"""Per-kernel device time for ONE 10 s stream (1 tile x 4 stems): where the single-stream latency goes.
python tools/single_stage_probe.py [n_streams]"""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spleeterrt_b200 as srt
from spleeterrt_b200 import workload as W
T, F, N = 512, 1024, 441000
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nets, _ = W.four_stem_nets(); S = len(nets)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sep = srt.Separator(nets, T, F, max_images=ns, max_batch_images=ns, device=0, cuda_stream=stream.cuda_stream)
din = torch.randn((ns, 2, N), device="cuda") * 0.1
dout = torch.empty((ns, S, 2, N), device="cuda")
n_arr = (C.c_size_t * ns)(*([N] * ns))
pl = (C.c_void_p * ns)(*[din[i, 0].data_ptr() for i in range(ns)]); pr = (C.c_void_p * ns)(*[din[i, 1].data_ptr() for i in range(ns)])
po = (C.c_void_p * (ns * S * 2))(*[dout[i, s, c].data_ptr() for i in range(ns) for s in range(S) for c in range(2)])
for _ in range(5): sep.separate_raw(pl, pr, n_arr, ns, None, po, device=True)
torch.cuda.synchronize()
K = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(K): sep.separate_raw(pl, pr, n_arr, ns, None, po, device=True)
e1.record(stream); torch.cuda.synchronize()
total = e0.elapsed_time(e1) / K
sep.set_timing(True)
for _ in range(K): sep.separate_raw(pl, pr, n_arr, ns, None, po, device=True)
torch.cuda.synchronize()
names = list(W.LAYER_FLOP_PER_PIXEL) + ["down1", "up6", "up7", "stft", "istft"]
spans = {k: round(sep.timing(k) / K * 1e3, 1) for k in names}
print("RESULT", json.dumps({"streams": ns, "ms_per_call_back_to_back": round(total, 4), "kernel_us": spans, "sum_kernel_us": round(sum(spans.values()), 1)}))
