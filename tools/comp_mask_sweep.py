"""Which layers need the compensation term?  SRT_COMP_MASK selects, per tensor-core layer (bit i: 0..4 = down2..down6, 5..9 = up1..up5),
whether the layer contracts the residuals of its inputs; this sweep measures, per mask, the stem error against the reference build
(bench configuration on the -12 dBFS signal, and the real drum / vocal nets on the full-scale clip) and the device time per 32-stream
step.  Run on a B200:   python tools/comp_mask_sweep.py > profiles/r2_comp_mask_sweep.json"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spleeterrt_b200 as srt
from oracle import oracle as O
from spleeterrt_b200 import workload as W

T, F, N = 512, 1024, 441000


def rms(a):
    return float(np.sqrt(np.mean(np.square(a, dtype=np.float64))))


nets, _ = W.stem_nets(4)
L, R = W.synth_pcm(0, n=N)
ref_bench = O.ref_exec().separate(nets, L, R, T, F)
nets_fs = [(nets[0][0], 1), (nets[3][0], 0)]
Lf, Rf = W.synth_pcm_fullscale(0, n=N)
ref_fs = O.ref_exec().separate(nets_fs, Lf, Rf, T, F)

ns = 32
pcm = [W.synth_pcm(i, n=N) for i in range(4)]
din = torch.empty((ns, 2, N), dtype=torch.float32)
for i in range(ns):
    din[i, 0] = torch.from_numpy(pcm[i % 4][0]); din[i, 1] = torch.from_numpy(pcm[i % 4][1])
din = din.cuda()
dout = torch.empty((ns, 4, 2, N), dtype=torch.float32, device="cuda")
n_arr = (C.c_size_t * ns)(*([N] * ns))
pl = (C.c_void_p * ns)(*[din[i, 0].data_ptr() for i in range(ns)]); pr = (C.c_void_p * ns)(*[din[i, 1].data_ptr() for i in range(ns)])
po = (C.c_void_p * (ns * 8))(*[dout[i, s, c].data_ptr() for i in range(ns) for s in range(4) for c in range(2)])
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)

MASKS = [("all ten layers (default)", 0x3ff), ("all but down2 and up5 (the two most expensive layers)", 0x1fe), ("encoder layers only", 0x01f),
         ("decoder layers only", 0x3e0), ("deep layers only: down4..down6, up1..up3", 0x0fc), ("none (= single-pass TF32)", 0x000)]
rows = []
for name, mask in MASKS:
    os.environ["SRT_COMP_MASK"] = hex(mask)
    one = srt.Separator(nets, T, F, max_images=1)
    eb = [rms(one.separate([(L, R)])[0][s] - ref_bench[s]) for s in range(4)]
    one.close()
    two = srt.Separator(nets_fs, T, F, max_images=1)
    ef = [rms(two.separate([(Lf, Rf)])[0][s] - ref_fs[s]) for s in range(2)]
    two.close()
    sep = srt.Separator(nets, T, F, max_images=ns, max_batch_images=ns, cuda_stream=stream.cuda_stream)
    for _ in range(3):
        sep.separate_raw(pl, pr, n_arr, ns, None, po, device=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        sep.separate_raw(pl, pr, n_arr, ns, None, po, device=True)
    e1.record(stream)
    torch.cuda.synchronize()
    sep.close()
    rows.append({"layers_compensated": name, "mask": hex(mask), "ms_per_32_stream_step": e0.elapsed_time(e1) / 10,
                 "stem_rms_err_bench_config": eb, "stem_rms_err_full_scale": ef})
    print(rows[-1], file=sys.stderr)
print(json.dumps({"what": "stem RMS error against the reference build and device time per step, per set of compensated layers (SRT_COMP_MASK); "
                          "default residual formats (e5m2, bf16 for down2 / up5); tolerance 1e-4", "rows": rows}, indent=1))
