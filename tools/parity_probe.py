"""Stem / mask error of the GPU path against the oracle port on a small case, printed as JSON (used to A/B numerics
switches such as SRT_UP6_DBG=32).  python tools/parity_probe.py"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spleeterrt_b200 as srt
from oracle import oracle as O
T, F = 128, 1024
nets = (O.four_stem_weights()[:2] + [(np.ascontiguousarray(O.half_to_float(O.real_weights_fp16())[1]), 0)]) if O.have_real_weights() else [(O.synthetic_weights(1), 1)]
L, R = O.synth_pcm(0, n=120000)
ref, rmask = O.separate(nets, L, R, T, F, want_masks=True)
sep = srt.Separator(nets, T, F, max_images=1, device=0)
got = sep.separate([(L, R)])[0]
sep.close()
err = [float(np.sqrt(np.mean((got[s].astype(np.float64) - ref[s]) ** 2))) for s in range(len(nets))]
print("PARITY", os.environ.get("SRT_UP6_DBG", "0"), json.dumps({"stem_rms_err": err, "stem_rms": [float(np.sqrt(np.mean(ref[s] ** 2))) for s in range(len(nets))]}))
