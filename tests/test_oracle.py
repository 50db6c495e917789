"""CPU tests: the oracle restatement (oracle/srt_oracle.c) is pinned against the reference's own
C code compiled into oracle/_ref (when present) and against the committed golden fixtures."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return oracle.ref_exec()


def test_coeff_layout(oracle):
    assert oracle.port().srt_oracle_coeff_floats() == 9822725          # sizeof(spleeterCoeff)/4
    v = oracle.coeff_views(np.zeros(oracle.COEFF_FLOATS, np.float32))
    assert v["down6.w"].shape == (512, 256, 5, 5) and v["up1.w"].shape == (512, 256, 5, 5)
    assert "down6.bn" not in v


def test_sigmoid_lut_vs_reference(oracle):
    r = _ref(oracle)
    xs = np.concatenate([np.linspace(-8, 8, 4001), np.float32(-7.0) + np.float32(0.01367188) * np.arange(1025, dtype=np.float32),
                         np.random.default_rng(0).uniform(-7.5, 7.5, 20000)]).astype(np.float32)
    a = np.array([r.lib.fastSigmoid(float(x)) for x in xs], np.float32)
    b = np.array([oracle.port().srt_oracle_sigmoid_lut(float(x)) for x in xs], np.float32)
    assert np.array_equal(a, b)                                          # the table is the reference's, bit for bit
    assert b[0] == 0.0 and b[4000] == 1.0                                # hard clip outside +-7 (spleeter.c:32-35)


def test_weights_sha(oracle):
    if not oracle.have_real_weights():
        pytest.skip("model blob not decoded")
    import hashlib
    h = oracle.real_weights_fp16()
    assert hashlib.sha256(h.tobytes()).hexdigest() == "b9837a8b6379c71b6442fbf4b0fb3c7eb668c0c76cf2702eda6ce7b7aa93e459"
    f = oracle.half_to_float(h)
    assert np.array_equal(f[np.abs(f) > 0], h.view(np.float16).astype(np.float32)[np.abs(f) > 0])


@pytest.mark.parametrize("mode", [0, 1])
def test_unet_port_vs_reference(oracle, small_nets, mode):
    r = _ref(oracle)
    rng = np.random.default_rng(3 + mode)
    x = (np.abs(rng.standard_normal((2, 64, 128))) * 4).astype(np.float32)
    coeff = small_nets[0][0] if mode else small_nets[1][0]
    yr = r.unet(coeff, x, mode)
    yp = oracle.unet(coeff, x, mode)
    assert np.abs(yr - yp).max() < 2e-6
    assert yr.std() > 1e-4


def test_stft_istft_port_vs_reference(oracle):
    r = _ref(oracle)
    rng = np.random.default_rng(5)
    n = 4096 * 3 + 8192 + 300                      # deliberately not a multiple of the hop
    L = (rng.standard_normal(n) * 0.3).astype(np.float32)
    R = (rng.standard_normal(n) * 0.3).astype(np.float32)
    sr, sp = r.stft(L, R), oracle.stft(L, R)
    for a, b in zip(sr, sp):
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-8
    orr, op = r.istft(*sr), oracle.istft(*sp)
    for a, b in zip(orr, op):
        assert np.abs(a - b).max() < 1e-6


def test_stft_matches_numpy_fft(oracle):
    """SURVEY §4 pin: re = Re rfft(x*hann(i+1/2))/4096, im = -Im(...)."""
    rng = np.random.default_rng(6)
    n = 4096 * 4
    L = rng.standard_normal(n).astype(np.float32)
    R = rng.standard_normal(n).astype(np.float32)
    re, im, _, _ = oracle.stft(L, R)
    i = np.arange(4096)
    hann = 0.5 * (1 - np.cos(2 * np.pi * (i + 0.5) / 4096))
    for f in range(3):
        X = np.fft.rfft(L[f * 1024:f * 1024 + 4096].astype(np.float64) * hann) / 4096
        assert np.abs(re[f, :2049] - X.real).max() < 2e-6
        assert np.abs(im[f, :2049] + X.imag).max() < 2e-6
    assert not re[:, 2049:].any()


def test_separate_port_vs_reference(oracle, small_nets):
    r = _ref(oracle)
    L, R = oracle.synth_pcm(0, n=30000)
    a = r.separate(small_nets, L, R, 64, 512)
    b = oracle.separate(small_nets, L, R, 64, 512)
    assert np.sqrt(((a - b) ** 2).mean()) < 1e-6
    assert np.abs(a).max() > 1e-3


@pytest.mark.parametrize("n_out", [2, 3])
def test_cli_output_modes_port_vs_reference(oracle, small_nets, n_out):
    """The CLI's 2-output (vocal, accompaniment = input - vocal) and 3-output cascade (drum net -> residual
    spectrum -> vocal net -> time-domain subtraction), main.c:776-970."""
    r = _ref(oracle)
    L, R = oracle.synth_pcm(1, n=21000)
    nets = small_nets if n_out == 3 else small_nets[1:]
    a = r.separate_cli(nets, L, R, 64, 256, n_out)
    b = oracle.separate_cli(nets, L, R, 64, 256, n_out)
    assert np.sqrt(((a - b) ** 2).mean()) < 1e-6
    assert np.abs(a).max() > 1e-3
    # the outputs of either mode add up to the inverse transform of the analysed input (linearity of the iSTFT)
    if n_out == 2:
        assert np.abs(a[0] + a[1] - np.stack([L, R])).max() < 1e-6


def test_golden_cli_cascade(oracle):
    """Committed fixture of the 3-output cascade, generated from the reference build (tests/golden/make_golden.py)."""
    p = os.path.join(GOLD, "cascade_T64_F64.npz")
    if not os.path.exists(p):
        pytest.skip("golden not generated")
    g = np.load(p)
    nets = [(oracle.synthetic_weights(int(g["seed_drum"])), 1), (oracle.synthetic_weights(int(g["seed_vocal"])), 0)]
    y = oracle.separate_cli(nets, g["L"], g["R"], 64, 64, 3)
    assert np.sqrt(((y - g["stems"]) ** 2).mean()) < 1e-6


def test_golden_unet(oracle):
    """Committed fixture generated from the reference build (tests/golden/make_golden.py)."""
    p = os.path.join(GOLD, "unet_T64_F64.npz")
    if not os.path.exists(p):
        pytest.skip("golden not generated")
    g = np.load(p)
    coeff = oracle.synthetic_weights(int(g["seed"]))
    for mode in (0, 1):
        y = oracle.unet(coeff, g["x"], mode)
        assert np.abs(y - g[f"mask_mode{mode}"]).max() < 2e-6


def test_golden_stft(oracle):
    p = os.path.join(GOLD, "stft_small.npz")
    if not os.path.exists(p):
        pytest.skip("golden not generated")
    g = np.load(p)
    planes = oracle.stft(g["L"], g["R"])
    for q, name in enumerate(("reL", "imL", "reR", "imR")):
        assert np.abs(planes[q][:, :2049] - g[name]).max() < 1e-8
    oL, oR = oracle.istft(*planes)
    assert np.abs(oL - g["outL"]).max() < 1e-6 and np.abs(oR - g["outR"]).max() < 1e-6


def test_vst_streamer_port_vs_reference(oracle):
    """oracle port of the real-time streamer vs the reference's own Spleeter4Stems.c build."""
    if not os.path.exists(os.path.join(oracle.REF_DIR, "libref_vst.so")):
        pytest.skip("reference VST build absent")
    T, F = 64, 512
    nets = [(oracle.synthetic_weights(40 + k), 1) for k in range(4)]
    n_total = (2 * T + 4) * 1024
    L, R = oracle.synth_pcm(7, n=n_total)
    ref, prt = oracle.RefVst(nets, T, F), oracle.PortVst(nets, T, F)
    a, b = [], []
    for o in range(0, n_total, 700):                      # ragged host blocks
        a.append(ref.process(L[o:o + 700], R[o:o + 700]))
        b.append(prt.process(L[o:o + 700], R[o:o + 700]))
    ref.close()
    prt.close()
    a, b = np.concatenate(a, axis=1), np.concatenate(b, axis=1)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a[0])
    assert np.sqrt(np.mean(a[:, ok] ** 2)) > 1e-3
    assert np.abs(a[:, ok] - b[:, ok]).max() < 5e-6


# ----------------------------------------------------------------------------- resampler front end
RS_CASES = [(48000, 5000, 2), (48000, 5000, 1), (22050, 3000, 2), (96000, 7000, 1), (88200, 4000, 2), (32000, 4801, 1),
            (48000, 4800, 1), (48000, 4800, 2), (8000, 999, 2), (44100 * 3, 3000, 1), (48000, 60, 2)]


@pytest.mark.parametrize("fs,n,ch", RS_CASES)
def test_resampler_port_vs_reference(oracle, fs, n, ch):
    """JamesDSPOfflineResampling (main.c:209-224) = libsamplerate sinc converter with the host's table: the port is
    bit-identical to the reference build, including where the converter stops (mono may end one frame short)."""
    if not oracle.have_ref_resampler():
        pytest.skip("oracle/_ref/libref_resample.so not built")
    rng = np.random.default_rng(fs + n + ch)
    x = (rng.standard_normal((n, ch)) * 0.3).astype(np.float32)
    x = x[:, 0] if ch == 1 else x
    for table in (None, oracle.synthetic_resampler_table()):
        ref = oracle.ref_resample(x, 44100.0 / fs, table)
        got, gen = oracle.resample(x, 44100.0 / fs, table if table is not None else oracle.resampler_table())
        assert np.array_equal(ref, got)
        assert ref.shape[0] - 1 <= gen <= ref.shape[0]
        assert np.abs(ref).max() > 1e-2


def test_golden_resampler(oracle):
    """Committed fixture generated by the reference's own converter on a synthetic table (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "resample_small.npz"))
    table = oracle.synthetic_resampler_table()
    for tag in ("48k_stereo", "48k_mono", "22k05_stereo", "96k_mono"):
        got, _ = oracle.resample(g[tag + "_in"], 44100.0 / float(g[tag + "_rate"]), table)
        assert np.array_equal(got, g[tag + "_out"].reshape(got.shape)), tag


# ----------------------------------------------------------------------------- independent arbiter
@pytest.mark.parametrize("mode", [0, 1])
def test_arbiter_float64_functional_unet_matches_port(oracle, small_nets, mode):
    """oracle/arbiter.py: textbook conv2d / conv_transpose2d semantics in float64 (torch-CPU) against the C restatement
    (itself pinned to the reference build): layouts, paddings, crops, BN/activation order and the sigmoid LUT agree."""
    from oracle import arbiter
    coeff = small_nets[0][0] if mode else small_nets[1][0]
    x = (np.abs(np.random.default_rng(17 + mode).standard_normal((2, 64, 128))) * 3).astype(np.float32)
    ref = oracle.unet(coeff, x, mode)
    got = arbiter.unet(coeff, x, mode).numpy()
    assert np.abs(got - ref).max() < 2e-5
    vst = arbiter.unet(coeff, x, mode, flavour=1).numpy()
    assert np.abs(vst - oracle.unet(coeff, x, mode, flavour=1)).max() < 2e-5


def test_tf32_emulation_error_budget(oracle, small_nets):
    """Rounding the tensors the GPU epilogues store for tensor-core consumers to TF32 (arbiter, float64 otherwise) moves
    the mask by ~1e-4 RMS: the budget the GPU parity tolerances (5e-4 mask, 1e-4 stem) are set against.  No single
    tensor dominates (tools/tf32_attribution.py)."""
    from oracle import arbiter
    coeff = small_nets[1][0]
    x = (np.abs(np.random.default_rng(23).standard_normal((2, 64, 256))) * 3).astype(np.float32)
    ref = arbiter.unet(coeff, x, 0).numpy()
    emu = arbiter.unet(coeff, x, 0, round_at="all").numpy()
    err = float(np.sqrt(np.mean((emu - ref) ** 2)))
    assert 1e-7 < err < 5e-4
    one = arbiter.unet(coeff, x, 0, round_at=["U2"]).numpy()
    assert float(np.sqrt(np.mean((one - ref) ** 2))) < err
    t = arbiter.round_tf32(__import__("torch").tensor([1.0 + 2.0 ** -11, 1.0 + 2.0 ** -12, -3.0], dtype=__import__("torch").float64))
    assert t.tolist() == [1.0 + 2.0 ** -10, 1.0, -3.0]
