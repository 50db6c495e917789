#!/usr/bin/env python3
"""Generate the committed golden fixtures FROM THE REFERENCE'S OWN CODE (oracle/_ref, built by
oracle/build_ref.py from /root/reference).  Run in the build container:

    python oracle/build_ref.py && python tests/golden/make_golden.py

Fixtures are small on purpose (they are committed): a T=64 x F=64 U-Net call with seeded
synthetic fp16-representable weights (both activation modes), a 20-frame STFT/iSTFT, a T=64 x F=64 run of the CLI's 3-output cascade and four short sample-rate conversions."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

r = O.ref_exec()
seed = 4242
coeff = O.synthetic_weights(seed)
rng = np.random.default_rng(99)
x = (np.abs(rng.standard_normal((2, 64, 64))) * 3).astype(np.float32)
out = {"seed": seed, "x": x}
for mode in (0, 1):
    out[f"mask_mode{mode}"] = r.unet(coeff, x, mode)
np.savez_compressed(os.path.join(HERE, "unet_T64_F64.npz"), **out)

n = 4096 * 3 + 8192 + 300
L = (rng.standard_normal(n) * 0.3).astype(np.float32)
R = (rng.standard_normal(n) * 0.3).astype(np.float32)
planes = r.stft(L, R)
oL, oR = r.istft(*planes)
np.savez_compressed(os.path.join(HERE, "stft_small.npz"), L=L, R=R, reL=planes[0][:, :2049], imL=planes[1][:, :2049],
                    reR=planes[2][:, :2049], imR=planes[3][:, :2049], outL=oL, outR=oR)
# 3-output cascade of the CLI (main.c:845-936): drum net (ELU) -> residual spectrum -> vocal net (LeakyReLU/ReLU)
n = 9000
L = (rng.standard_normal(n) * 0.2).astype(np.float32)
R = (rng.standard_normal(n) * 0.2).astype(np.float32)
sd, sv = 5151, 5252
nets = [(O.synthetic_weights(sd), 1), (O.synthetic_weights(sv), 0)]
stems = r.separate_cli(nets, L, R, 64, 64, 3)
np.savez_compressed(os.path.join(HERE, "cascade_T64_F64.npz"), seed_drum=sd, seed_vocal=sv, L=L, R=R, stems=stems)
# resampler front end (main.c:209-224 -> libsamplerate sinc): the reference's code on a SYNTHETIC coefficient table of
# the reference's geometry (the real table is reference data and stays out of the repository)
table = O.synthetic_resampler_table()
rs = {}
for tag, fs, n, ch in (("48k_stereo", 48000, 1500, 2), ("48k_mono", 48000, 1600, 1), ("22k05_stereo", 22050, 700, 2), ("96k_mono", 96000, 2000, 1)):
    x = (rng.standard_normal((n, ch)) * 0.3).astype(np.float32)
    x = x[:, 0] if ch == 1 else x
    rs[tag + "_in"] = x
    rs[tag + "_rate"] = fs
    rs[tag + "_out"] = O.ref_resample(x, 44100.0 / fs, table)
np.savez_compressed(os.path.join(HERE, "resample_small.npz"), **rs)
print("golden fixtures written")
