#!/usr/bin/env python3
"""BASELINE.json config 5: VST block-size sweep — per-block latency p50/p99 of
Spleeter4StemsProcessSamples-equivalent calls on one B200 (4-stem stereo, T=256, F=1536 as the
plugin sets them, PluginProcessor.cpp:124).  Blocks of 2048 are split into 2 x 1024 by the host as
PluginProcessor.cpp:173-181 does.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spleeterrt_b200 as srt  # noqa: E402
from spleeterrt_b200 import workload as W  # noqa: E402

T, F = 256, 1536
nets, desc = W.four_stem_nets()
res = {"config": {"time_step": T, "bin_limit": F, "stems": 4, "weights": desc}, "blocks": {}}
n_blocks = int(os.environ.get("SRT_LAT_BLOCKS", "2400"))
L, R = W.synth_pcm(0, n=1024 * 64)
for block in (256, 512, 1024, 2048):
    st = srt.Streamer([c for c, _ in nets], T, F)
    lat = []
    pos = 0
    for i in range(n_blocks):
        if pos + block > L.size:
            pos = 0
        l, r = L[pos:pos + block], R[pos:pos + block]
        pos += block
        t0 = time.perf_counter()
        for o in range(0, block, 1024):
            st.process(l[o:o + 1024], r[o:o + 1024])
        lat.append((time.perf_counter() - t0) * 1e3)
    st.close()
    lat = np.array(lat[50:])
    res["blocks"][str(block)] = {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)),
                                 "max_ms": float(lat.max()), "mean_ms": float(lat.mean()),
                                 "realtime_budget_ms": block / 44.1, "calls": int(lat.size)}
print(json.dumps(res))
