"""One profiled step of the bench workload (32 x 10 s streams, 4 stems, T=512, F=1024) for ncu:
3 warm-up steps, then cudaProfilerStart / one step / cudaProfilerStop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/prof_step.py
    ncu --profile-from-start off --set full --clock-control none -o gpurun_out/full python tools/prof_step.py
"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spleeterrt_b200 as srt
from spleeterrt_b200 import workload as W

T, F, N = 512, 1024, 441000
ns = int(os.environ.get("SRT_PROF_STREAMS", "32"))
nets, _ = W.four_stem_nets(); S = len(nets)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sep = srt.Separator(nets, T, F, max_images=ns, max_batch_images=ns, device=0, cuda_stream=stream.cuda_stream)
pcm = [W.synth_pcm(i, n=N) for i in range(4)]
hin = torch.empty((ns, 2, N))
for i in range(ns):
    hin[i, 0] = torch.from_numpy(pcm[i % 4][0]); hin[i, 1] = torch.from_numpy(pcm[i % 4][1])
din = hin.cuda()
dout = torch.empty((ns, S, 2, N), device="cuda")
n_arr = (C.c_size_t * ns)(*([N] * ns))
pl = (C.c_void_p * ns)(*[din[i, 0].data_ptr() for i in range(ns)]); pr = (C.c_void_p * ns)(*[din[i, 1].data_ptr() for i in range(ns)])
po = (C.c_void_p * (ns * S * 2))(*[dout[i, s, c].data_ptr() for i in range(ns) for s in range(S) for c in range(2)])
for _ in range(3):
    sep.separate_raw(pl, pr, n_arr, ns, None, po, device=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
sep.separate_raw(pl, pr, n_arr, ns, None, po, device=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step,", ns, "streams")
sep.close()
