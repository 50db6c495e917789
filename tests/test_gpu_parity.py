"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle.

Tolerances
  * transforms (fp32 FFT vs the reference's fp32 Hartley): 2e-6 absolute on spectra of O(1e-2)
  * U-Net layer tensors: 5e-3 relative RMS (TF32 operands, fp32 accumulate)
  * soft masks: 5e-4 RMS / 3e-2 max absolute (1e-3 for the exact-sigmoid flavour on white-noise input, see the test)
  * separated stems: 1e-4 RMS per stem (the tolerance BASELINE.json's north_star states)
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rms(a):
    return float(np.sqrt(np.mean(np.square(a, dtype=np.float64))))


@pytest.fixture(scope="module")
def srt():
    import spleeterrt_b200 as m
    m.load_library()
    return m


@pytest.fixture(scope="module")
def xform(srt):
    s = srt.Separator([], 64, 64, max_images=4)
    yield s
    s.close()


# ----------------------------------------------------------------------------- transforms
def test_stft_vs_oracle(srt, oracle, xform):
    rng = np.random.default_rng(5)
    n = 4096 * 3 + 8192 + 300
    L = (rng.standard_normal(n) * 0.3).astype(np.float32)
    R = (rng.standard_normal(n) * 0.3).astype(np.float32)
    got, ref = xform.stft(L, R), oracle.stft(L, R)
    for a, b in zip(got, ref):
        assert a.shape == b.shape
        assert np.abs(a - b).max() < 2e-6
    assert not got[0][:, 2049:].any() and not got[0][-3:].any()      # untouched rows stay zero (stftFix.c:367-371)


def test_stft_golden(srt, xform):
    g = np.load(os.path.join(GOLD, "stft_small.npz"))
    planes = xform.stft(g["L"], g["R"])
    for q, name in enumerate(("reL", "imL", "reR", "imR")):
        assert np.abs(planes[q][:, :2049] - g[name]).max() < 2e-6
    oL, oR = xform.istft(*planes)
    assert np.abs(oL - g["outL"]).max() < 5e-6 and np.abs(oR - g["outR"]).max() < 5e-6


def test_istft_vs_oracle_and_reconstruction(srt, oracle, xform):
    rng = np.random.default_rng(8)
    n = 4096 * 6
    L = (rng.standard_normal(n) * 0.3).astype(np.float32)
    R = (rng.standard_normal(n) * 0.3).astype(np.float32)
    planes = oracle.stft(L, R)
    gL, gR = xform.istft(*planes)
    oL, oR = oracle.istft(*planes)
    assert np.abs(gL - oL).max() < 5e-6 and np.abs(gR - oR).max() < 5e-6
    # interior samples covered by 4 computed frames reconstruct the input (SURVEY §4 pin)
    lo, hi = 3072, (planes[0].shape[0] - 3 - 3) * 1024
    assert np.abs(gL[lo:hi] - L[lo:hi]).max() < 5e-6


def test_stft_linearity_full_size(srt, xform):
    """Size-independent property at the benchmark length (10 s): STFT is linear."""
    rng = np.random.default_rng(9)
    n = 450560
    a = (rng.standard_normal(n) * 0.2).astype(np.float32)
    b = (rng.standard_normal(n) * 0.2).astype(np.float32)
    big = srt.Separator([], 64, 64, max_images=8)
    sa, sb, sab = big.stft(a, b), big.stft(b, a), big.stft(a + b, a + b)
    assert np.abs(sa[0] + sb[0] - sab[0]).max() < 5e-6
    assert np.abs(sa[0] - sb[2]).max() < 1e-7          # channel symmetry of the packed complex FFT
    oL, _ = big.istft(*sa)
    assert np.abs(oL[4096:n - 8192] - a[4096:n - 8192]).max() < 5e-6
    big.close()


# ----------------------------------------------------------------------------- U-Net
def _layer_check(srt, oracle, nets, T, F, impl, n_img=2):
    rng = np.random.default_rng(T * 7 + F)
    x = (np.abs(rng.standard_normal((n_img, 2, T, F))) * 3).astype(np.float32)
    sep = srt.Separator(nets, T, F, max_images=n_img, conv_impl=impl)
    y = sep.process_spleeter(x)
    worst = {}
    for s, (coeff, mode) in enumerate(nets):
        for b in range(n_img):
            mask, tp = oracle.unet(coeff, x[b], mode, taps=True)
            taps = oracle.split_taps(tp, T, F)
            for name, ref in taps.items():
                got = sep.debug_tensor(name, n_img)[s, b]
                rel = rms(got - ref) / max(rms(ref), 1e-12)
                worst[name] = max(worst.get(name, 0.0), rel)
            worst["mask_rms"] = max(worst.get("mask_rms", 0.0), rms(y[s, b] - mask))
            worst["mask_max"] = max(worst.get("mask_max", 0.0), float(np.abs(y[s, b] - mask).max()))
    sep.close()
    return worst


@pytest.mark.parametrize("T,F", [(64, 128), (128, 192)])
def test_unet_layers_simt_path(srt, oracle, small_nets, T, F):
    """Same tables / packed weights / layouts evaluated by plain SIMT loads: isolates host logic."""
    w = _layer_check(srt, oracle, small_nets, T, F, impl=1)
    bad = {k: v for k, v in w.items() if not k.startswith("mask") and v > 5e-3}
    assert not bad, f"layer mismatch {bad} (all: {w})"
    assert w["mask_rms"] < 5e-4 and w["mask_max"] < 3e-2, w


@pytest.mark.parametrize("T,F", [(64, 128), (128, 192), (64, 512)])
def test_unet_layers_tcgen05_path(srt, oracle, small_nets, T, F):
    w = _layer_check(srt, oracle, small_nets, T, F, impl=0)
    bad = {k: v for k, v in w.items() if not k.startswith("mask") and v > 5e-3}
    assert not bad, f"layer mismatch {bad} (all: {w})"
    assert w["mask_rms"] < 5e-4 and w["mask_max"] < 3e-2, w


@pytest.mark.parametrize("T,F", [(64, 1152), (128, 1024)])
def test_unet_layers_rowpatch_kernel(srt, oracle, small_nets, T, F, monkeypatch):
    """Row-patch tensor-core kernel forced on for down2/down3/up4/up5 (multi-tile rows, halos, partial tiles)."""
    monkeypatch.setenv("SRT_CONV_RP", "1")
    w = _layer_check(srt, oracle, small_nets[:1], T, F, impl=0, n_img=2)
    bad = {k: v for k, v in w.items() if not k.startswith("mask") and v > 5e-3}
    assert not bad, f"layer mismatch {bad} (all: {w})"
    assert w["mask_rms"] < 5e-4 and w["mask_max"] < 3e-2, w


def test_unet_golden(srt, oracle):
    g = np.load(os.path.join(GOLD, "unet_T64_F64.npz"))
    coeff = oracle.synthetic_weights(int(g["seed"]))
    sep = srt.Separator([(coeff, 0), (coeff, 1)], 64, 64, max_images=1)
    y = sep.process_spleeter(g["x"])
    for mode in (0, 1):
        d = y[mode, 0] - g[f"mask_mode{mode}"]
        assert rms(d) < 5e-4 and np.abs(d).max() < 3e-2
    sep.close()


def test_tier_a_process_spleeter(srt, oracle, small_nets):
    """The reference's own entry points (spleeter.h) through the shared library.  The sizes are passed with garbage in their upper
    32 bits, as a caller compiled against the VST flavour's `int width, int height` prototype may leave them."""
    lib = srt.load_library()
    lib.allocateSpleeterStr.restype = C.c_void_p
    lib.initSpleeter.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    lib.getMaskPtr.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_float))]
    lib.processSpleeter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.freeSpleeter.argtypes = [C.c_void_p]
    T, F = 64, 128
    coeff, mode = small_nets[1]
    x = (np.abs(np.random.default_rng(1).standard_normal((2, T, F))) * 3).astype(np.float32)
    nn = lib.allocateSpleeterStr()
    lib.initSpleeter(nn, F | (0xdeadbeef << 32), T | (0x7fff1234 << 32), mode, coeff.ctypes.data)
    mp = C.POINTER(C.c_float)()
    lib.getMaskPtr(nn, C.byref(mp))
    lib.processSpleeter(nn, x.ctypes.data, C.cast(mp, C.c_void_p))
    got = np.ctypeslib.as_array(mp, shape=(2, T, F)).copy()
    lib.freeSpleeter(nn)
    ref = oracle.unet(coeff, x, mode)
    assert rms(got - ref) < 5e-4


# ----------------------------------------------------------------------------- full path
def test_separate_vs_oracle_small(srt, oracle, small_nets):
    L, R = oracle.synth_pcm(0, n=30000)
    sep = srt.Separator(small_nets, 64, 512, max_images=1)
    got = sep.separate([(L, R)])[0]
    ref = oracle.separate(small_nets, L, R, 64, 512)
    for s in range(len(small_nets)):
        assert rms(got[s] - ref[s]) < 1e-4, f"stem {s}: rms {rms(got[s] - ref[s])}"
        assert rms(ref[s]) > 1e-4
    assert sep.launch_count() >= 16
    sep.close()


def test_separate_ragged_batch_chunked(srt, oracle, small_nets):
    """Streams of different length, one spanning two tiles, U-Net batch smaller than the batch."""
    T, F = 64, 256
    lens = [20000, 70000, 4097]
    streams = [tuple(x[:n] for x in oracle.synth_pcm(i, n=max(lens))) for i, n in enumerate(lens)]
    sep = srt.Separator(small_nets[:1], T, F, max_images=2, max_batch_images=4)
    got = sep.separate(streams, unaffected=[0.25])
    for (L, R), g in zip(streams, got):
        ref = oracle.separate(small_nets[:1], L, R, T, F, unaffected=0.25)
        assert g.shape == ref.shape
        assert rms(g - ref) < 1e-4
    sep.close()


def test_async_batches_in_flight_match_blocking_call(srt, oracle, small_nets):
    """srt_separate_batch_async: four different batches submitted back to back (three staging slots, the
    fourth submit drains the first) give bit-identical stems to the blocking call, in any wait order."""
    T, F = 64, 256
    sep = srt.Separator(small_nets[:2], T, F, max_images=2, max_batch_images=4)
    batches = [[tuple(x[:n] for x in oracle.synth_pcm(10 * b + i, n=50000)) for i, n in enumerate(lens)]
               for b, lens in enumerate([(30000, 50000), (8192,), (41000, 4100, 12345), (20480, 20480)])]
    want = [sep.separate(b, unaffected=[0.25, 0.25]) for b in batches]
    pend = [sep.separate_async(b, unaffected=[0.25, 0.25]) for b in batches]
    for k in (2, 0, 3, 1):
        got = sep.result(pend[k])
        for g, w in zip(got, want[k]):
            assert np.array_equal(g, w)
    ref = oracle.separate(small_nets[:2], *batches[1][0], T, F, unaffected=0.25)
    assert rms(want[1][0] - ref) < 1e-4
    with pytest.raises(srt.SrtError):
        sep.wait(10 ** 6)      # never issued
    sep.close()


@pytest.mark.parametrize("n_out", [2, 3])
def test_cli_output_modes_vs_oracle(srt, oracle, small_nets, n_out):
    """srt_create_cli: the CLI's 2-output (vocal, input - vocal) and 3-output cascade (drum net -> residual spectrum
    -> vocal net -> time-domain subtraction), main.c:776-970, as one device call; ragged batch, one stream over
    two tiles, U-Net batch smaller than the batch."""
    T, F = 64, 256
    nets = small_nets if n_out == 3 else small_nets[1:]      # (drum: ELU, vocal: LeakyReLU/ReLU) as main.c
    lens = [21000, 70000, 4097]
    streams = [tuple(x[:n] for x in oracle.synth_pcm(20 + i, n=max(lens))) for i, n in enumerate(lens)]
    sep = srt.CliSeparator([c for c, _ in nets], n_out, T, F, max_images=2, max_batch_images=4)
    got = sep.separate(streams, unaffected=[0.1])
    for (L, R), g in zip(streams, got):
        ref = oracle.separate_cli(nets, L, R, T, F, n_out, unaffected=0.1)
        assert g.shape == ref.shape == (n_out, 2, L.size)
        for k in range(n_out):
            assert rms(g[k] - ref[k]) < 1e-4, f"output {k}: rms {rms(g[k] - ref[k])}"
            assert rms(ref[k]) > 1e-4
        if n_out == 2:      # vocal + accompaniment = input, exactly one rounding apart
            assert np.abs(g[0] + g[1] - np.stack([L, R])).max() < 1e-6
    # the asynchronous flavour shares the staging path: bit-identical
    again = sep.result(sep.separate_async(streams, unaffected=[0.1]))
    for a, b in zip(again, got):
        assert np.array_equal(a, b)
    sep.close()


def test_cli_cascade_golden(srt, oracle):
    """3-output cascade against the committed fixture generated from the reference build (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "cascade_T64_F64.npz"))
    coeffs = [oracle.synthetic_weights(int(g["seed_drum"])), oracle.synthetic_weights(int(g["seed_vocal"]))]
    sep = srt.CliSeparator(coeffs, 3, 64, 64, max_images=2)
    y = sep.separate([(g["L"], g["R"])])[0]
    for k in range(3):
        assert rms(y[k] - g["stems"][k]) < 1e-4
    sep.close()


def test_cli_mode_argument_errors(srt, oracle):
    coeff = oracle.synthetic_weights(5)
    with pytest.raises(srt.SrtError):
        srt.CliSeparator([coeff], 4, 64, 64)
    with pytest.raises(srt.SrtError):
        srt.CliSeparator([coeff], 3, 64, 64)


def test_interleaved_frames_in_and_out(srt, oracle, small_nets):
    """srt_separate_batch_interleaved: the decoder's interleaved frames in (stereo and mono, main.c:767-769), the WAV
    writer's interleaved frames out (main.c:806): bit-identical to the planar entry point, and the mono stream equals
    the oracle run with R = L."""
    T, F = 64, 256
    (L0, R0), (L1, _) = oracle.synth_pcm(30, n=33000), oracle.synth_pcm(31, n=9001)
    sep = srt.Separator(small_nets, T, F, max_images=2, max_batch_images=3)
    planar = sep.separate([(L0, R0), (L1, L1)], unaffected=[0.1, 0.1])
    inter = sep.separate_interleaved([np.stack([L0, R0], 1), L1], unaffected=[0.1, 0.1])
    for pl, it in zip(planar, inter):
        assert it.shape == (2, pl.shape[2], 2)
        assert np.array_equal(it.transpose(0, 2, 1), pl)
    ref = oracle.separate(small_nets, L1, L1, T, F)
    assert rms(inter[1].transpose(0, 2, 1) - ref) < 1e-4
    sep.close()
    # CLI 2-output mode through the same formats: accompaniment = input - vocal reads the interleaved input
    cli = srt.CliSeparator([small_nets[1][0]], 2, T, F, max_images=2)
    a = cli.separate([(L0, R0)])[0]
    b = cli.separate_interleaved([np.stack([L0, R0], 1)])[0]
    assert np.array_equal(b.transpose(0, 2, 1), a)
    m = cli.separate_interleaved([L1])[0]
    assert np.abs(m[0] + m[1] - np.stack([L1, L1], 1)).max() < 1e-6
    cli.close()
    with pytest.raises(srt.SrtError):
        sep2 = srt.Separator(small_nets[:1], T, F)
        try:
            sep2.separate_interleaved([np.zeros((5000, 3), np.float32)])
        finally:
            sep2.close()


def test_unity_mask_is_identity_full_size(srt, oracle):
    """Size-independent property at benchmark shape (T=512, F=1024, 10 s): all-zero weights with
    a +100 head bias give mask == 1, so every stem reproduces the input (SURVEY §4 'VST stream' pin)."""
    coeff = np.zeros(oracle.COEFF_FLOATS, np.float32)
    oracle.coeff_views(coeff)["up7.b"][:] = 100.0
    L, R = oracle.synth_pcm(3, n=441000)
    sep = srt.Separator([(coeff, 1)], 512, 1024, max_images=1)
    out = sep.separate([(L, R)], unaffected=[1.0])[0]
    assert np.abs(out[0, 0] - L).max() < 5e-6 and np.abs(out[0, 1] - R).max() < 5e-6
    sep.close()


def test_batch_invariance_full_size(srt, oracle, small_nets):
    """Size-independent property at the benchmark shape (T=512, F=1024): a stream's stems do not depend on what else
    is in the batch, on its position in it, or on how the tiles are split into U-Net passes (units = (stream, tile,
    stem) share no state) - bit for bit.  One stream spans two tiles (15 s)."""
    a = tuple(oracle.synth_pcm(40, n=441000))
    b = tuple(x[:300000] for x in oracle.synth_pcm(41, n=300000))
    c = tuple(oracle.synth_pcm(42, n=15 * 44100))
    sep = srt.Separator(small_nets, 512, 1024, max_images=4, max_batch_images=4)
    alone = sep.separate([a])[0]
    mixed = sep.separate([b, c, a])
    assert np.array_equal(mixed[2], alone)
    sep.close()
    sep1 = srt.Separator(small_nets, 512, 1024, max_images=2, max_batch_images=3)     # two U-Net passes
    again = sep1.separate([c, a])
    assert np.array_equal(again[1], alone) and np.array_equal(again[0], mixed[1])
    sep1.close()
    assert rms(alone) > 1e-3
    # and the masks are soft masks: the two stems of net (drum, vocal) never exceed the input's energy by much
    assert rms(alone[0]) < 2 * rms(np.stack(a)) and rms(alone[1]) < 2 * rms(np.stack(a))


@pytest.mark.parametrize("n", [1, 1023, 4096, 4097, 12288])
def test_boundary_lengths_vs_oracle(srt, oracle, small_nets, n):
    """Framing edge cases of main.c:762-767: a single sample, less than a hop, exactly one / just over one FFT block,
    a whole number of blocks."""
    L, R = (x[:n] for x in oracle.synth_pcm(50, n=16384))
    sep = srt.Separator(small_nets[:1], 64, 128, max_images=1)
    got = sep.separate([(L, R)])[0]
    sep.close()
    ref = oracle.separate(small_nets[:1], L, R, 64, 128)
    assert got.shape == ref.shape == (1, 2, n)
    assert rms(got - ref) < 1e-4 and np.abs(got - ref).max() < 1e-3


def test_long_stream_runs_in_several_unet_passes(srt, oracle, small_nets):
    """A stream with more T-frame tiles than the U-Net batch (max_images): the tiles go through in several passes
    (processMT walks them one by one, main.c:545-577) and the result equals the all-at-once run bit for bit."""
    L, R = oracle.synth_pcm(80, n=200000)                      # 4 tiles at T = 64
    one = srt.Separator(small_nets[:1], 64, 128, max_images=1, max_batch_images=4)
    a = one.separate([(L, R)])[0]
    one.close()
    allb = srt.Separator(small_nets[:1], 64, 128, max_images=4)
    b = allb.separate([(L, R)])[0]
    allb.close()
    assert np.array_equal(a, b)
    assert rms(a - oracle.separate(small_nets[:1], L, R, 64, 128)) < 1e-4


def test_maximum_bin_limit(srt, oracle, small_nets):
    """F = 2048, the largest analyseBinLimit the CLI accepts (main.c:744-748): only the Nyquist bin is left to
    unaffectedWeight."""
    L, R = oracle.synth_pcm(90, n=30000)
    sep = srt.Separator(small_nets[:1], 64, 2048, max_images=1)
    got = sep.separate([(L, R)], unaffected=[0.25])[0]
    sep.close()
    ref = oracle.separate(small_nets[:1], L, R, 64, 2048, unaffected=0.25)
    assert rms(got - ref) < 1e-4 and rms(ref) > 1e-3


def test_empty_inputs_are_errors(srt, oracle, small_nets):
    sep = srt.Separator(small_nets[:1], 64, 128, max_images=1)
    with pytest.raises(srt.SrtError):
        sep.separate([(np.zeros(0, np.float32), np.zeros(0, np.float32))])      # a stream without samples
    import ctypes as C
    with pytest.raises(srt.SrtError):
        sep.separate_raw(None, None, None, 0, None, None)                          # a batch without streams
    sep.close()


def test_two_devices_in_one_process(srt, oracle, small_nets):
    """srt_config.device: contexts on two GPUs of one process give identical stems (the opt-in shared-memory sizes of
    the kernels are per-device attributes and must be set on each)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    L, R = oracle.synth_pcm(60, n=40000)
    outs = []
    for dev in (0, 1):
        sep = srt.Separator(small_nets, 64, 1024, max_images=1, device=dev)
        outs.append(sep.separate([(L, R)])[0])
        sep.close()
    assert np.array_equal(outs[0], outs[1])
    assert rms(outs[0] - oracle.separate(small_nets, L, R, 64, 1024)) < 1e-4


@pytest.mark.parametrize("fs,n,ch", [(48000, 48000, 2), (48000, 4800, 1), (22050, 9000, 2), (96000, 20000, 1), (8000, 999, 2)])
def test_resampler_bit_exact_vs_oracle(srt, oracle, xform, fs, n, ch):
    """srt_resample_host against the oracle converter (itself bit-identical to the reference's libsamplerate build):
    integer bookkeeping on the host, taps in double on the GPU without FMA contraction -> array_equal."""
    rng = np.random.default_rng(fs + n)
    x = (rng.standard_normal((n, ch)) * 0.3).astype(np.float32)
    x = x[:, 0] if ch == 1 else x
    table = oracle.resampler_table() if oracle.have_resampler_table() else oracle.synthetic_resampler_table()
    ref, gen_ref = oracle.resample(x, 44100.0 / fs, table)
    got, gen = xform.resample(x, 44100.0 / fs, table)
    assert gen == gen_ref and got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_resampler_golden_and_chain(srt, oracle, small_nets):
    """The reference-generated fixture, and a 48 kHz stereo clip through resample -> separate on the GPU against the
    oracle's resample -> separate."""
    g = np.load(os.path.join(GOLD, "resample_small.npz"))
    table = oracle.synthetic_resampler_table()
    sep = srt.Separator(small_nets[:1], 64, 256, max_images=1)
    for tag in ("48k_stereo", "48k_mono", "22k05_stereo", "96k_mono"):
        got, _ = sep.resample(g[tag + "_in"], 44100.0 / float(g[tag + "_rate"]), table)
        assert np.array_equal(got, g[tag + "_out"].reshape(got.shape)), tag
    x = np.stack(oracle.synth_pcm(70, n=24000), 1)         # the §8d test signal, read as a 48 kHz clip
    y, gen = sep.resample(x, 44100.0 / 48000.0, table)
    assert gen == y.shape[0] == 22050
    stems = sep.separate_interleaved([y])[0]
    yo, _ = oracle.resample(x, 44100.0 / 48000.0, table)
    ref = oracle.separate(small_nets[:1], yo[:, 0].copy(), yo[:, 1].copy(), 64, 256)
    assert rms(stems.transpose(0, 2, 1) - ref) < 1e-4
    with pytest.raises(srt.SrtError):
        sep.resample(x, 1e-4, table)
    sep.close()


def test_capacity_errors_are_loud(srt, oracle):
    coeff = oracle.synthetic_weights(5)
    sep = srt.Separator([(coeff, 1)], 64, 64, max_images=1)
    L, R = oracle.synth_pcm(0, n=200000)          # needs 4 tiles > max_images
    with pytest.raises(srt.SrtError):
        sep.separate([(L, R)])
    sep.close()
    with pytest.raises(srt.SrtError):
        srt.Separator([(coeff, 1)], 60, 64)       # T must be a multiple of 64


# ----------------------------------------------------------------------------- streaming (VST) flavour
def test_unet_vst_flavour(srt, oracle, small_nets):
    """exact sigmoid + unclamped ELU (VST/Source/spleeter.c:56-77)"""
    T, F = 64, 128
    coeff = small_nets[0][0]
    x = (np.abs(np.random.default_rng(2).standard_normal((2, T, F))) * 3).astype(np.float32)
    sep = srt.Separator([(coeff, 1)], T, F, max_images=1, flavour=1)
    y = sep.process_spleeter(x)[0, 0]
    sep.close()
    ref = oracle.unet(coeff, x, 1, flavour=1)
    # white-noise input, no LUT clipping of the tails: single-pass TF32 operands gave 6.0e-4 mask RMS here (float64 emulation and
    # round 1's GPU path); the default compensated precision has to stay an order of magnitude under that
    assert rms(y - ref) < 1e-4 and np.abs(y - ref).max() < 5e-3
    fast = srt.Separator([(coeff, 1)], T, F, max_images=1, flavour=1, precision="tf32")
    yf = fast.process_spleeter(x)[0, 0]
    fast.close()
    assert rms(yf - ref) < 1e-3 and np.abs(yf - ref).max() < 3e-2


@pytest.mark.parametrize("block", [1024, 512, 300])
def test_streamer_vs_reference_vst(srt, oracle, block):
    """Spleeter4Stems flavour against the reference's own streamer (oracle/_ref/libref_vst.so): same
    delayed output (2*T*1024 + 1024 samples), same untouched-output pattern, 1e-4 RMS per component."""
    T, F = 64, 512
    nets = [(oracle.synthetic_weights(40 + k), 1) for k in range(4)]
    n_total = (2 * T + 6) * 1024
    L, R = oracle.synth_pcm(7, n=n_total)
    have_ref = os.path.exists(os.path.join(oracle.REF_DIR, "libref_vst.so"))
    ref = oracle.RefVst(nets, T, F) if have_ref else oracle.PortVst(nets, T, F)   # the port is pinned to the reference on CPU
    got = srt.Streamer([c for c, _ in nets], T, F)
    outs_r, outs_g = [], []
    for o in range(0, n_total, block):
        l, r = L[o:o + block], R[o:o + block]
        outs_r.append(ref.process(l, r))
        outs_g.append(got.process(l, r))
    launches = got.launch_count()
    ref.close()
    got.close()
    a, b = np.concatenate(outs_r, axis=1), np.concatenate(outs_g, axis=1)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "output availability pattern differs"
    ok = ~np.isnan(a[0])
    tail = slice(2 * T * 1024 + 1024, None)                   # where real (non-zero) output starts
    assert rms(a[:, tail][:, ok[tail]]) > 1e-3
    for j in range(8):
        assert rms(a[j][ok] - b[j][ok]) < 1e-4, f"component {j}: {rms(a[j][ok] - b[j][ok])}"
    assert launches >= n_total // 1024


def test_five_stem_batch(srt, oracle, small_nets):
    """BASELINE.json configs[3] shape in small: 5 nets per stream (n_stems is a context parameter)."""
    T, F = 64, 256
    nets = [(small_nets[0][0], 1), (small_nets[1][0], 0)] + [(oracle.synthetic_weights(60 + k), 1) for k in range(3)]
    streams = [tuple(x[:n] for x in oracle.synth_pcm(20 + i, n=40000)) for i, n in enumerate((40000, 25000))]
    sep = srt.Separator(nets, T, F, max_images=2, max_batch_images=2)
    got = sep.separate(streams)
    sep.close()
    for (L, R), g in zip(streams, got):
        ref = oracle.separate(nets, L, R, T, F)
        assert g.shape == (5, 2, L.size)
        for s in range(5):
            assert rms(g[s] - ref[s]) < 1e-4


def test_fp32_weights_use_two_term_split(srt, oracle, small_nets):
    """fp32 weights that are not fp16/TF32-representable (the VST's `.dat` dumps): the library detects them and
    contracts tf32(w) + tf32(w - tf32(w)); stems stay within 1e-4 RMS of the oracle."""
    rng = np.random.default_rng(31)
    nets = [((c * (1 + 1e-3 * rng.standard_normal(c.shape))).astype(np.float32), m) for c, m in small_nets]
    L, R = oracle.synth_pcm(2, n=30000)
    sep = srt.Separator(nets, 64, 512, max_images=1)
    got = sep.separate([(L, R)])[0]
    sep.close()
    ref = oracle.separate(nets, L, R, 64, 512)
    for s in range(len(nets)):
        assert rms(got[s] - ref[s]) < 1e-4, f"stem {s}: {rms(got[s] - ref[s])}"


def test_unet_pass_replays_as_a_cuda_graph(srt, oracle, small_nets, monkeypatch):
    """The U-Net pass is captured into a CUDA graph on its second use per (first image, images, mask destination) and replayed
    afterwards: results are bit-identical to plain launches (SRT_GRAPHS=0), call after call, with a changing batch in between,
    and the launch counter keeps counting kernels."""
    T, F = 64, 512
    a, b = oracle.synth_pcm(81, n=60000), oracle.synth_pcm(82, n=30000)
    monkeypatch.setenv("SRT_GRAPHS", "0")
    plain = srt.Separator(small_nets, T, F, max_images=2, max_batch_images=2)
    want_a, want_ab = plain.separate([a])[0], plain.separate([a, b])
    n_plain = plain.launch_count()
    plain.close()
    monkeypatch.delenv("SRT_GRAPHS")
    sep = srt.Separator(small_nets, T, F, max_images=2, max_batch_images=2)
    for _ in range(4):                                         # direct, capture + replay, replay, replay
        assert np.array_equal(sep.separate([a])[0], want_a)
    got_ab = sep.separate([a, b])                              # another key
    assert np.array_equal(got_ab[0], want_ab[0]) and np.array_equal(got_ab[1], want_ab[1])
    assert np.array_equal(sep.separate([a])[0], want_a)
    assert sep.launch_count() > 2 * n_plain                    # replays are counted per kernel, not per graph
    sep.close()


def test_shared_weight_sets(srt, oracle, small_nets):
    """srt_config.share_weights (what the tier-A initSpleeter uses: processMT creates its instances from ONE coefficient
    pointer, main.c:557): contexts with the same blobs and configuration share one device copy of the packed weights;
    results equal a private context's bit for bit, in any destruction order, and a different configuration gets its own set."""
    T, F = 64, 256
    L, R = oracle.synth_pcm(83, n=40000)
    nets = [(np.ascontiguousarray(c), m) for c, m in small_nets]
    private = srt.Separator(nets, T, F, max_images=1)
    want = private.separate([(L, R)])[0]
    private.close()
    a = srt.Separator(nets, T, F, max_images=1, share_weights=True)
    b = srt.Separator(nets, T, F, max_images=1, share_weights=True)          # hit: no packing, no upload
    c = srt.Separator(nets, T, 2 * F, max_images=1, share_weights=True)      # other geometry: own set
    assert np.array_equal(a.separate([(L, R)])[0], want)
    a.close()                                                                # b keeps the set alive
    assert np.array_equal(b.separate([(L, R)])[0], want)
    ref_c = oracle.separate(nets, L, R, T, 2 * F)
    assert rms(c.separate([(L, R)])[0] - ref_c) < 1e-4
    b.close()
    c.close()
    d = srt.Separator(nets, T, F, max_images=1, share_weights=True)          # the set died with b: rebuilt
    assert np.array_equal(d.separate([(L, R)])[0], want)
    d.close()
