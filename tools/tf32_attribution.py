"""Where the TF32 path's error comes from, without a GPU: the float64 arbiter (oracle/arbiter.py) with TF32 rounding
switched on at one stored tensor at a time (the tensors the GPU epilogues round for a tensor-core consumer), on the
SURVEY §8d test signal.  Prints mask RMS error and the stem RMS error it causes.   python tools/tf32_attribution.py"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O, arbiter as A

T, F = 128, 1024
L, R = O.synth_pcm(0, n=T * 1024 - 4096)
n = L.size
padded = O.FFT * ((n + O.FFT - 1) // O.FFT) + 2 * O.FFT
pl, pr = np.zeros(padded, np.float32), np.zeros(padded, np.float32)
pl[O.FFT:O.FFT + n], pr[O.FFT:O.FFT + n] = L, R
planes = O.stft(pl, pr)
mag = np.zeros((2, T, F), np.float32)
fr = min(T, planes[0].shape[0])
mag[0, :fr] = np.hypot(planes[0][:fr, :F], planes[1][:fr, :F]) * 4096
mag[1, :fr] = np.hypot(planes[2][:fr, :F], planes[3][:fr, :F]) * 4096
nets = O.four_stem_weights()
res = {}
for name, (coeff, _), mode in (("drum net (ELU)", nets[0], 1), ("vocal net (LeakyReLU/ReLU)", nets[3], 0)):
    ref = A.unet(coeff, mag, mode).numpy()
    rows = {}
    for pts in [[p] for p in A.ROUND_POINTS] + ["all"]:
        emu = A.unet(coeff, mag, mode, round_at=pts).numpy()
        key = pts if pts == "all" else pts[0]
        rows[key] = {"mask_rms": float(np.sqrt(np.mean((emu - ref) ** 2))), "stem_rms": A.stem_error_from_mask_error(ref, emu, L, R, T, F)}
        print(name, key, rows[key], flush=True)
    res[name] = rows
print("RESULT", json.dumps(res))
