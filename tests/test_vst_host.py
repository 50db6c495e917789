"""Boundary proof for the VST flavour (SURVEY 8b tier A `vst`): examples/vst_host.c drives the streaming API the way the JUCE
plugin shell does (VST/Source/PluginProcessor.cpp:46-97, 114-125, 130-182) - malloc(getCoeffSize()) + fread of the four .dat
files, malloc(sizeof(Spleeter4Stems)), Spleeter4StemsInit(msr, 1536, 256, coeff), host blocks cut into <= OVPSIZE slices,
Spleeter4StemsFree.  The SAME source is built twice: against include/Spleeter4Stems.h + libspleeterrt_b200.so
(examples/_build/vst_host) and against the reference's header + the reference's own streamer (oracle/_ref/vst_host_ref).
Both binaries process the same file with the same host block size; the eight component channels must agree."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "examples", "_build", "vst_host")
REF = os.path.join(ROOT, "oracle", "_ref", "vst_host_ref")


def rms(a):
    return float(np.sqrt(np.mean(np.square(a, dtype=np.float64))))


def _write_nets(tmp_path):
    from spleeterrt_b200 import workload as W
    nets, _ = W.stem_nets(4)
    paths = []
    for name, (coeff, _) in zip(("drum", "bass", "accompaniment", "vocal"), nets):
        p = str(tmp_path / f"{name}4stems.dat")
        np.ascontiguousarray(coeff, np.float32).tofile(p)
        assert os.path.getsize(p) == 39290900          # the size the plugin freads (PluginProcessor.cpp:60)
        paths.append(p)
    return paths


def test_vst_host_builds_against_both_headers():
    """not gpu: the stand-in host compiles and links against the drop-in header/library (and, in the build container, against
    the reference's own header and objects)."""
    assert os.path.exists(OURS), "examples/_build/vst_host missing: python __graft_entry__.py"
    out = subprocess.run(["nm", "-D", "--undefined-only", OURS], capture_output=True, text=True).stdout
    for sym in ("Spleeter4StemsInit", "Spleeter4StemsProcessSamples", "Spleeter4StemsFree", "getCoeffSize"):
        assert sym in out, sym


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(OURS) and os.path.exists(REF)), reason="vst_host binaries not built")
@pytest.mark.parametrize("block,bin_limit,time_step,hops", [(1536, 1536, 256, 2 * 256 + 48), (300, 512, 64, 2 * 64 + 20)])
def test_plugin_style_host_matches_reference_streamer(tmp_path, block, bin_limit, time_step, hops):
    from spleeterrt_b200 import workload as W
    nets = _write_nets(tmp_path)
    n = hops * 1024 + 517                                        # not a multiple of anything
    L, R = W.synth_pcm_fullscale(5, n=n)
    src = str(tmp_path / "in.f32")
    np.stack([L, R], 1).astype(np.float32).tofile(src)
    outs = {}
    for tag, exe in (("ref", REF), ("b200", OURS)):
        dst = str(tmp_path / f"out_{tag}.f32")
        env = dict(os.environ)
        env.setdefault("OMP_NUM_THREADS", str(min(os.cpu_count() or 1, 16)))
        r = subprocess.run([exe, *nets, src, dst, str(block), str(bin_limit), str(time_step)], capture_output=True, text=True,
                           timeout=900, env=env)
        assert r.returncode == 0, (tag, r.returncode, r.stdout[-300:], r.stderr[-500:])
        info = json.loads(r.stdout.strip().splitlines()[-1])
        assert info["block"] == block and info["blocks"] == (n + block - 1) // block
        outs[tag] = np.fromfile(dst, np.float32).reshape(n, 8)
    a, b = outs["ref"], outs["b200"]
    delay = 2 * time_step * 1024 + 1024
    assert rms(a[delay:]) > 1e-2, "reference streamer produced no signal"
    errs = [rms(a[:, c] - b[:, c]) for c in range(8)]
    print(f"\n[parity] plugin-style host, block {block}, T={time_step} F={bin_limit}: component rms err {errs}")
    assert max(errs) < 2e-5, errs                              # tolerance 1e-4 (north star), 5x margin asserted
