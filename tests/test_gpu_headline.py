"""GPU parity at the configurations the metric is quoted on (-m gpu), through the C ABI, against the reference itself
(oracle/_ref build of the reference's C sources; the pinned C port where that build is absent).

  * bench.py's configuration: four ELU nets (workload.four_stem_nets), stereo, T = 512, F = 1024, one 10 s stream
  * the CLI's real vocal net in mode 0 and the drum net in mode 1 at the same shape
  * both on the -12 dBFS SURVEY 8d signal AND on a full-scale clip (peak 0.999, RMS 0.29)
  * the streamer at the plugin's shape, T = 256 / F = 1536 (PluginProcessor.cpp:124), against libref_vst.so

Tolerance: 1e-4 RMS per stem (BASELINE.json north_star).  The default precision (compensated: TF32 main term + a residual
term in e5m2 / bf16, include/srt_b200.h srt_config.precision) and its all-bf16 variant have to hold it on every input with a
5x margin (2e-5 asserted); the single-pass TF32 mode is checked at 1e-4 on the -12 dBFS signal only - at full scale it
sits AT 1e-4, which the last test records instead of hiding.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

T, F, N10 = 512, 1024, 441000


def rms(a):
    return float(np.sqrt(np.mean(np.square(a, dtype=np.float64))))


@pytest.fixture(scope="module")
def srt():
    import spleeterrt_b200 as m
    m.load_library()
    return m


@pytest.fixture(scope="module")
def W():
    from spleeterrt_b200 import workload
    return workload


def _reference(oracle, nets, L, R, unaffected=0.1):
    if oracle.have_ref():
        return oracle.ref_exec().separate(nets, L, R, T, F, unaffected=unaffected), "reference build"
    return oracle.separate(nets, L, R, T, F, unaffected=unaffected), "port"


@pytest.fixture(scope="module")
def bench_case(oracle, W):
    """bench.py's configuration on stream 0 and its reference stems (computed once: ~2-8 s of host time per net)."""
    nets, _ = W.stem_nets(4)
    L, R = W.synth_pcm(0, n=N10)
    ref, kind = _reference(oracle, nets, L, R)
    return nets, L, R, ref, kind


@pytest.fixture(scope="module")
def fullscale_case(oracle, W):
    """Real drum net (ELU) and real vocal net (LeakyReLU/ReLU, as the CLI runs it) on the full-scale clip."""
    four, _ = W.stem_nets(4)
    nets = [(four[0][0], 1), (four[3][0], 0)]
    L, R = W.synth_pcm_fullscale(0, n=N10)
    ref, kind = _reference(oracle, nets, L, R)
    return nets, L, R, ref, kind


def _separate(srt, nets, L, R, precision, unaffected=0.1):
    sep = srt.Separator(nets, T, F, max_images=1, precision=precision)
    got = sep.separate([(L, R)], unaffected=[unaffected] * len(nets))[0]
    sep.close()
    return got


@pytest.mark.parametrize("precision", [None, "compensated_bf16"])
def test_bench_config_default_precision(srt, bench_case, precision):
    """4 nets, stereo, T=512/F=1024, the §8d signal: every stem within 1e-4 RMS of the reference - with margin."""
    nets, L, R, ref, kind = bench_case
    got = _separate(srt, nets, L, R, precision)
    errs = [rms(got[s] - ref[s]) for s in range(4)]
    lvls = [rms(ref[s]) for s in range(4)]
    print(f"\n[parity] bench config ({precision or 'default'}) vs {kind}: stem rms err {errs}, stem rms {lvls}")
    assert all(l > 1e-3 for l in lvls)
    assert max(errs) < 2e-5, errs                                   # tolerance 1e-4, 5x margin asserted
    # relative to each stem's own level as well, so that the two quiet stems (-43 dBFS: masks near 0, where an absolute mask
    # error of 1e-5 is 2e-4 of the stem) do not hide behind an absolute tolerance
    assert max(e / l for e, l in zip(errs, lvls)) < 1e-3, (errs, lvls)


def test_bench_config_tf32_precision(srt, bench_case):
    nets, L, R, ref, kind = bench_case
    got = _separate(srt, nets, L, R, "tf32")
    errs = [rms(got[s] - ref[s]) for s in range(4)]
    print(f"\n[parity] bench config, single-pass TF32 vs {kind}: stem rms err {errs}")
    assert max(errs) < 1e-4, errs


@pytest.mark.parametrize("precision", [None, "compensated_bf16"])
def test_full_scale_input_default_precision(srt, fullscale_case, precision):
    """Full-scale input (peak 0.999, RMS 0.29), the real drum (ELU) and vocal (mode 0) nets at T=512/F=1024."""
    nets, L, R, ref, kind = fullscale_case
    got = _separate(srt, nets, L, R, precision)
    errs = [rms(got[s] - ref[s]) for s in range(2)]
    lvls = [rms(ref[s]) for s in range(2)]
    print(f"\n[parity] full-scale clip ({precision or 'default'}) vs {kind}: stem rms err {errs}, stem rms {lvls}")
    assert max(errs) < 2e-5, errs
    assert np.abs(got - ref).max() < 1e-3


def test_full_scale_input_tf32_precision_is_recorded(srt, fullscale_case):
    """Single-pass TF32 at full scale: the error is relative to the level (~1e-4 .. 5e-4 of the stem), so 1e-4 absolute is
    not guaranteed.  Asserted: it stays within 5e-4 and is at least 5x worse than the compensated default - i.e. the
    default is what buys the parity, and this mode is a documented trade."""
    nets, L, R, ref, kind = fullscale_case
    fast = _separate(srt, nets, L, R, "tf32")
    comp = _separate(srt, nets, L, R, "compensated")
    ef = [rms(fast[s] - ref[s]) for s in range(2)]
    ec = [rms(comp[s] - ref[s]) for s in range(2)]
    print(f"\n[parity] full-scale clip: tf32 {ef} vs compensated {ec}")
    assert max(ef) < 5e-4
    assert max(ef) > 5 * max(ec)


def test_layer_tensors_are_fp32_grade_in_default_precision(srt, oracle, small_nets):
    """The unrounded tensors at the end of the tensor-core chain (up5: fp32 input of up6; up6; skip1) and the masks, on a
    white-noise magnitude image at a multi-tile shape that exercises both kernel forms (row-patch: down2, down3, up4, up5;
    generic: the rest).  Measured on B200: up5 5e-5 relative RMS compensated against 9e-4 single-pass TF32.  What is left is
    not operand rounding (2^-19) but the accumulation inside the tensor core: ~10^3 chained MMAs per output add their
    products into the fp32 TMEM accumulator with truncation, ~2^-24 each, which sums to a few 1e-5 (the CPU reference's
    round-to-nearest running sum over the same K is ~5e-6)."""
    Ts, Fs = 128, 1024
    rng = np.random.default_rng(77)
    x = (np.abs(rng.standard_normal((1, 2, Ts, Fs))) * 3).astype(np.float32)
    res = {}
    for prec in ("compensated", "compensated_bf16", "tf32"):
        sep = srt.Separator(small_nets, Ts, Fs, max_images=1, precision=prec)
        y = sep.process_spleeter(x)
        worst = {}
        for s, (coeff, mode) in enumerate(small_nets):
            mask, tp = oracle.unet(coeff, x[0], mode, taps=True)
            taps = oracle.split_taps(tp, Ts, Fs)
            for name in ("skip1", "up5", "up6"):
                got = sep.debug_tensor(name, 1)[s, 0]
                worst[name] = max(worst.get(name, 0.0), rms(got - taps[name]) / rms(taps[name]))
            worst["mask"] = max(worst.get("mask", 0.0), rms(y[s, 0] - mask))
        sep.close()
        res[prec] = worst
    print(f"\n[parity] layer tensors: {res}")
    c, b, f = res["compensated"], res["compensated_bf16"], res["tf32"]
    assert b["up5"] < 1.5e-4 and b["up6"] < 1.5e-4 and b["mask"] < 2e-5, res
    assert c["up5"] < 2.5e-4 and c["up6"] < 2.5e-4 and c["mask"] < 4e-5, res            # 8-bit residuals: a little above the bf16 form
    assert f["up5"] > 4 * c["up5"] and f["mask"] > 4 * c["mask"] and f["up5"] > 8 * b["up5"], res


def test_streamer_at_plugin_shape(srt, oracle, W):
    """Spleeter4Stems at the plugin's own shape, T = 256 / F = 1536 (PluginProcessor.cpp:124), 1024-sample blocks, against
    the reference's streamer (libref_vst.so): the first real output tile, 1e-4 RMS per component."""
    Ts, Fs = 256, 1536
    nets, _ = W.stem_nets(4)
    n_total = (2 * Ts + 40) * 1024
    L, R = W.synth_pcm_fullscale(3, n=n_total)
    have_ref = os.path.exists(os.path.join(oracle.REF_DIR, "libref_vst.so"))
    ref = oracle.RefVst(nets, Ts, Fs) if have_ref else oracle.PortVst(nets, Ts, Fs)
    got = srt.Streamer([c for c, _ in nets], Ts, Fs)
    outs_r, outs_g = [], []
    for o in range(0, n_total, 1024):
        outs_r.append(ref.process(L[o:o + 1024], R[o:o + 1024]))
        outs_g.append(got.process(L[o:o + 1024], R[o:o + 1024]))
    ref.close()
    got.close()
    a, b = np.concatenate(outs_r, axis=1), np.concatenate(outs_g, axis=1)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "output availability pattern differs"
    ok = ~np.isnan(a[0])
    tail = slice(2 * Ts * 1024 + 1024, None)
    assert rms(a[:, tail][:, ok[tail]]) > 1e-2
    errs = [rms(a[j][ok] - b[j][ok]) for j in range(8)]
    print(f"\n[parity] streamer T=256 F=1536 vs {'libref_vst.so' if have_ref else 'port'}: component rms err {errs}")
    assert max(errs) < 2e-5, errs
