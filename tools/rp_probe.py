#!/usr/bin/env python3
"""GPU diagnostic: per-layer error of the row-patch tensor-core kernel against the oracle for both
descriptor base-offset conventions (SRT_RP_BO=0/1) and for the generic kernel.  Not a test."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spleeterrt_b200 as srt  # noqa: E402
from oracle import oracle as O  # noqa: E402

T, F = 64, 1152
nets = [(O.half_to_float(O.real_weights_fp16())[0], 1)] if O.have_real_weights() else [(O.synthetic_weights(3), 1)]
rng = np.random.default_rng(1)
x = (np.abs(rng.standard_normal((1, 2, T, F))) * 3).astype(np.float32)
mask, tp = O.unet(nets[0][0], x[0], 1, taps=True)
taps = O.split_taps(tp, T, F)
for label, env in (("generic", {"SRT_CONV_RP": "0"}), ("rowpatch bo=1", {"SRT_CONV_RP": "1", "SRT_RP_BO": "1"}),
                   ("rowpatch bo=0", {"SRT_CONV_RP": "1", "SRT_RP_BO": "0"})):
    os.environ.update(env)
    try:
        sep = srt.Separator(nets, T, F, max_images=1)
        y = sep.process_spleeter(x)
        errs = {}
        for name in ("skip2", "skip3", "up4", "up5"):
            got = sep.debug_tensor(name, 1)[0, 0]
            ref = taps[name]
            errs[name] = float(np.sqrt(np.mean((got - ref) ** 2)) / max(np.sqrt(np.mean(ref ** 2)), 1e-12))
        errs["mask_rms"] = float(np.sqrt(np.mean((y[0, 0] - mask) ** 2)))
        sep.close()
        print(label, {k: f"{v:.2e}" for k, v in errs.items()}, flush=True)
    except Exception as e:  # a trap poisons the context; report and stop
        print(label, "FAILED:", e, flush=True)
        break
