"""CPU tests of the host logic: k-block tables, weight packing, layouts (via the CPU model of the
GPU kernels in tests/host_model.cpp) and the C-ABI surface of the built library."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENC = [2, 16, 32, 64, 128, 256, 512]
ACT = {"leaky": 1, "relu": 2, "elu_clamp": 3, "elu": 4}


def _act(kind, x):
    if kind == 1:
        return np.where(x >= 0, x, 0.2 * x)
    if kind == 2:
        return np.maximum(x, 0)
    return np.where(x >= 0, x, np.where(x < -15, -1.0, np.expm1(np.minimum(x, 0)))).astype(np.float32)


def _run_layer(hm, T, F, idx, coeff, act, s0, s1, out_shape, want_act=0, row=False):
    out = np.zeros(out_shape, np.float32)
    s0 = np.ascontiguousarray(s0, np.float32)
    p1 = None
    if s1 is not None:
        s1 = np.ascontiguousarray(s1, np.float32)
        p1 = s1.ctypes.data
    fn = hm.srt_host_model_row_layer if row else hm.srt_host_model_layer
    rc = fn(T, F, idx, np.ascontiguousarray(coeff).ctypes.data_as(C.c_void_p), act,
                                 s0.ctypes.data_as(C.c_void_p), C.c_void_p(p1), out.ctypes.data_as(C.c_void_p), want_act)
    assert rc == 0
    return out


@pytest.mark.parametrize("T,F,mode", [(64, 128, 1), (64, 192, 0), (128, 64, 1)])
def test_gather_gemm_model_matches_oracle(oracle, host_model, small_nets, T, F, mode):
    coeff = small_nets[0][0] if mode else small_nets[1][0]
    rng = np.random.default_rng(T + F)
    x = (np.abs(rng.standard_normal((2, T, F))) * 3).astype(np.float32)
    mask, tp = oracle.unet(coeff, x, mode, taps=True)
    taps = oracle.split_taps(tp, T, F)
    v = oracle.coeff_views(coeff)
    a_enc, a_dec = (3, 3) if mode else (1, 2)
    # encoder down2..down6: input = act(scale*skip+offset) of the previous layer
    for i in range(1, 6):
        skip = taps[f"skip{i}"]
        bn = v[f"down{i}.bn"]
        act_in = _act(a_enc, bn[1][:, None, None] * skip + bn[0][:, None, None]).astype(np.float32)
        ref = taps[f"skip{i+1}"]
        for row in ([False, True] if i - 1 in (0, 1) else [False]):
            got = _run_layer(host_model, T, F, i - 1, coeff, a_enc, act_in, None, ref.shape, row=row)
            err = np.abs(got - ref).max() / max(1e-6, np.abs(ref).max())
            assert err < 2e-5, f"down{i+1} (row-patch={row}): rel err {err}"
    # decoder up1..up5: the generic form with fused parities where 4 * cout <= 256 (up3..up5) and phase-separated, and the row-patch form
    for d in range(5):
        s0 = taps["skip6"] if d == 0 else taps[f"skip{6-d}"]
        s1 = None if d == 0 else taps[f"up{d}"]
        ref = taps[f"up{d+1}"]
        for row, fuse in ([(False, 1), (False, 0), (True, 1)] if 5 + d in (8, 9) else [(False, 1), (False, 0)]):
            host_model.srt_host_model_set_fuse(fuse)
            try:
                got = _run_layer(host_model, T, F, 5 + d, coeff, a_dec, s0, s1, ref.shape, row=row)
            finally:
                host_model.srt_host_model_set_fuse(1)
            err = np.abs(got - ref).max() / max(1e-6, np.abs(ref).max())
            assert err < 2e-5, f"up{d+1} (row-patch={row}, fused={fuse}): rel err {err}"


def test_down1_tensor_core_plan(oracle, host_model, small_nets):
    """space-to-depth magnitude + 9 taps x 8 channels + stems fused into N reproduces down1 (skip1)."""
    T, F = 64, 128
    rng = np.random.default_rng(11)
    x = (np.abs(rng.standard_normal((2, T, F))) * 3).astype(np.float32)
    coeffs = [np.ascontiguousarray(c) for c, _ in small_nets]
    cp = (C.c_void_p * 2)(*[c.ctypes.data for c in coeffs])
    out = np.zeros((2, 16, T // 2, F // 2), np.float32)
    assert host_model.srt_host_model_down1(T, F, cp, 2, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
    for s, (coeff, mode) in enumerate(small_nets):
        _, tp = oracle.unet(coeff, x, mode, taps=True)
        ref = oracle.split_taps(tp, T, F)["skip1"]
        assert np.abs(out[s] - ref).max() / np.abs(ref).max() < 2e-5      # hi + lo split keeps the first layer fp32-accurate


def test_two_term_weights_for_fp32_blobs(oracle, host_model, small_nets):
    """Weights that are not TF32-exact (fp32 `.dat` dumps): the k-block twins tf32(w) + tf32(w - tf32(w)) recover them."""
    T, F = 64, 128
    rng = np.random.default_rng(21)
    coeff = (small_nets[0][0] * (1 + 1e-3 * rng.standard_normal(small_nets[0][0].shape))).astype(np.float32)
    x = (np.abs(rng.standard_normal((2, T, F))) * 3).astype(np.float32)
    _, tp = oracle.unet(coeff, x, 1, taps=True)
    taps = oracle.split_taps(tp, T, F)
    v = oracle.coeff_views(coeff)
    bn = v["down2.bn"]
    act_in = _act(3, bn[1][:, None, None] * taps["skip2"] + bn[0][:, None, None]).astype(np.float32)
    errs = {}
    for split in (0, 1):
        host_model.srt_host_model_set_split(split)
        for name, idx, row in (("down3", 1, False), ("down3-row", 1, True)):
            got = _run_layer(host_model, T, F, idx, coeff, 3, act_in, None, taps["skip3"].shape, row=row)
            errs[(name, split)] = np.abs(got - taps["skip3"]).max() / np.abs(taps["skip3"]).max()
        got = _run_layer(host_model, T, F, 9, coeff, 3, taps["skip2"], taps["up4"], taps["up5"].shape, row=True)
        errs[("up5-row", split)] = np.abs(got - taps["up5"]).max() / np.abs(taps["up5"]).max()
    host_model.srt_host_model_set_split(0)
    for name in ("down3", "down3-row", "up5-row"):
        assert errs[(name, 1)] < 2e-5, errs
        assert errs[(name, 0)] > 3 * errs[(name, 1)], errs      # single-term weights are visibly worse


@pytest.mark.parametrize("mode", [1, 0])
def test_compensated_precision_tables(oracle, host_model, small_nets, mode):
    """precision 0 (the default): every tensor-core layer contracts tf32(a) with w and the bf16 residual a - tf32(a) with
    bf16(w) (extra 64-channel k-blocks; decoder layers read ONE residual tensor [skip | up]).  The CPU model of both
    kernel forms, fed TF32-rounded sources + bf16 residuals, must agree with the oracle as well as the unrounded model."""
    T, F = 64, 128
    coeff = small_nets[0][0] if mode else small_nets[1][0]
    rng = np.random.default_rng(5 + mode)
    x = (np.abs(rng.standard_normal((2, T, F))) * 3).astype(np.float32)
    _, tp = oracle.unet(coeff, x, mode, taps=True)
    taps = oracle.split_taps(tp, T, F)
    v = oracle.coeff_views(coeff)
    a_enc, a_dec = (3, 3) if mode else (1, 2)
    host_model.srt_host_model_set_comp(1)
    try:
        for i in range(1, 6):
            bn = v[f"down{i}.bn"]
            act_in = _act(a_enc, bn[1][:, None, None] * taps[f"skip{i}"] + bn[0][:, None, None]).astype(np.float32)
            ref = taps[f"skip{i+1}"]
            for row in ([False, True] if i - 1 in (0, 1) else [False]):
                got = _run_layer(host_model, T, F, i - 1, coeff, a_enc, act_in, None, ref.shape, row=row)
                err = np.abs(got - ref).max() / max(1e-6, np.abs(ref).max())
                assert err < 2e-5, f"down{i+1} compensated (row-patch={row}): rel err {err}"
        for d in range(5):
            s0 = taps["skip6"] if d == 0 else taps[f"skip{6-d}"]
            s1 = None if d == 0 else taps[f"up{d}"]
            ref = taps[f"up{d+1}"]
            for row in ([False, True] if 5 + d in (8, 9) else [False]):
                got = _run_layer(host_model, T, F, 5 + d, coeff, a_dec, s0, s1, ref.shape, row=row)
                err = np.abs(got - ref).max() / max(1e-6, np.abs(ref).max())
                assert err < 2e-5, f"up{d+1} compensated (row-patch={row}): rel err {err}"
        # control: the same rounded sources WITHOUT the compensation blocks are visibly worse (so the blocks above did the work)
        host_model.srt_host_model_set_comp(2)
        for row in (False, True):
            got = _run_layer(host_model, T, F, 9, coeff, a_dec, taps["skip2"], taps["up4"], taps["up5"].shape, row=row)
            err = np.abs(got - taps["up5"]).max() / np.abs(taps["up5"]).max()
            assert err > 5e-5, f"control (row-patch={row}): {err}"
    finally:
        host_model.srt_host_model_set_comp(0)


def test_e5m2_conversion(host_model):
    """e5m2_rn (the host's statement of cvt.rn.satfinite.e5m2x2.f32, used to pack the 8-bit compensation weights): every one of the
    256 codes round-trips, values round to the nearest code (ties to even), and out-of-range values saturate."""
    host_model.srt_host_model_e5m2_value.restype = C.c_float
    host_model.srt_host_model_e5m2.argtypes = [C.c_float]
    codes = [b for b in range(256) if (b >> 2) & 31 != 31]
    vals = {b: host_model.srt_host_model_e5m2_value(b) for b in codes}
    for b, v in vals.items():
        if v == 0.0:
            continue
        assert host_model.srt_host_model_e5m2(v) == b, (b, v)
    pos = sorted(v for v in vals.values() if v > 0)
    rng = np.random.default_rng(0)
    for x in np.exp(rng.uniform(np.log(1e-6), np.log(7e4), 4000)).astype(np.float32):
        got = host_model.srt_host_model_e5m2_value(host_model.srt_host_model_e5m2(float(x)))
        best = min([0.0] + pos, key=lambda v: abs(v - float(x)))
        assert abs(got - float(x)) <= abs(best - float(x)) * (1 + 1e-6), (x, got, best)
    assert host_model.srt_host_model_e5m2_value(host_model.srt_host_model_e5m2(1e9)) == 57344.0
    assert host_model.srt_host_model_e5m2_value(host_model.srt_host_model_e5m2(-1e9)) == -57344.0
    assert host_model.srt_host_model_e5m2(1.25 + 0.125) == host_model.srt_host_model_e5m2(1.5)       # tie 1.375 -> even mantissa (1.5 = 0b10)


@pytest.mark.parametrize("mode", [1, 0])
def test_fp8_compensation_tables(oracle, host_model, small_nets, mode):
    """The 8-bit residual format (e5m2(4 lo) x e5m2(w / 4), 128 channels per block; down2 and up5 keep bf16 because their
    residual tensors have 64 channels): tables, packing and scaling reproduce the oracle to ~1e-5, between the bf16 form
    (~1e-6) and uncompensated TF32 (~2e-4)."""
    T, F = 64, 128
    coeff = small_nets[0][0] if mode else small_nets[1][0]
    rng = np.random.default_rng(9 + mode)
    x = (np.abs(rng.standard_normal((2, T, F))) * 3).astype(np.float32)
    _, tp = oracle.unet(coeff, x, mode, taps=True)
    taps = oracle.split_taps(tp, T, F)
    v = oracle.coeff_views(coeff)
    a_enc, a_dec = (3, 3) if mode else (1, 2)
    errs = {}
    for fmt in (3, 1, 2):                     # fp8, bf16, rounded without compensation
        host_model.srt_host_model_set_comp(fmt)
        try:
            for i in (1, 2, 3):               # down2 (row-patch: 64-channel e5m2 blocks in 64-byte rows), down3 (both forms), down4
                bn = v[f"down{i}.bn"]
                act_in = _act(a_enc, bn[1][:, None, None] * taps[f"skip{i}"] + bn[0][:, None, None]).astype(np.float32)
                ref = taps[f"skip{i+1}"]
                for row in ([True] if i == 1 else [False, True] if i == 2 else [False]):
                    got = _run_layer(host_model, T, F, i - 1, coeff, a_enc, act_in, None, ref.shape, row=row)
                    errs[(f"down{i+1}", row, fmt)] = rms_rel(got, ref)
            for d, rows in ((1, [False]), (2, [False]), (3, [False, True]), (4, [True])):      # up2, up3 (fused parities), up4 (both forms), up5 (row-patch, narrow blocks)
                ref = taps[f"up{d+1}"]
                for row in rows:
                    got = _run_layer(host_model, T, F, 5 + d, coeff, a_dec, taps[f"skip{6-d}"], taps[f"up{d}"], ref.shape, row=row)
                    errs[(f"up{d+1}", row, fmt)] = rms_rel(got, ref)
        finally:
            host_model.srt_host_model_set_comp(0)
    for (name, row, fmt), e in errs.items():
        if fmt == 3:
            assert e < 4e-5, (name, row, e)
            assert e < errs[(name, row, 2)] / 4, (name, row, e, errs[(name, row, 2)])      # clearly better than no compensation
            assert e > errs[(name, row, 1)], (name, row)                                      # and coarser than bf16


def rms_rel(a, b):
    return float(np.sqrt(np.mean((a.astype(np.float64) - b) ** 2)) / np.sqrt(np.mean(b.astype(np.float64) ** 2)))


def test_plan_shapes(host_model):
    info = (C.c_int * 10)()
    # shape A (T=512, F=1024), batch 32: tiles are full and the k-block counts match the design
    # up3 / up5 (4 * cout <= 256) fuse their four output parities: one list over the 3 x 3 input offsets x sources x 32-channel slabs
    expect_nkb = {0: [15], 1: [25], 2: [50], 3: [100], 4: [200], 5: [64, 96, 96, 144], 6: [64, 96, 96, 144], 7: [72], 9: [18]}
    for idx, nkb in expect_nkb.items():
        assert host_model.srt_host_model_plan_info(512, 1024, 32, idx, info) == 0
        tw, th, nb, n_tile, n_tiles, phases = info[0:6]
        assert tw * th * nb == 128 and n_tile * n_tiles in (32, 64, 128, 256, 512, 16)
        assert list(info[6:6 + phases]) == nkb
    assert host_model.srt_host_model_plan_info(512, 1024, 32, 7, info) == 0 and info[3] == 256 and info[5] == 1     # up3: N = 4 x 64
    host_model.srt_host_model_set_fuse(0)
    try:
        for idx, nkb in {7: [32, 48, 48, 72], 9: [8, 12, 12, 18]}.items():
            assert host_model.srt_host_model_plan_info(512, 1024, 32, idx, info) == 0
            assert info[5] == 4 and list(info[6:10]) == nkb
    finally:
        host_model.srt_host_model_set_fuse(1)
    # tiny deep layers batch images into the tile
    host_model.srt_host_model_plan_info(64, 64, 32, 4, info)
    assert info[0] * info[1] == 1 and info[2] == 32 or info[0] * info[1] * info[2] == 128


def test_round_tf32(host_model):
    f = host_model.srt_host_model_round_tf32
    assert f(1.0) == 1.0
    x = np.float32(1.0 + 2 ** -11)                   # exactly half an ulp of TF32: ties away from zero
    assert f(float(x)) == np.float32(1.0 + 2 ** -10)
    h = np.float16(0.333).astype(np.float32)          # fp16 values are exact in TF32
    assert f(float(h)) == h


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", txt)
    return {n for n in names if n not in ("defined",)}


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads (no GPU needed for dlopen) and exports every include/*.h symbol."""
    import spleeterrt_b200 as srt
    if not os.path.exists(srt.lib_path()):
        from spleeterrt_b200.build import build
        build()
    lib = C.CDLL(srt.lib_path())
    exported = srt.exported_symbols()
    for header, listed in srt.HEADER_SYMBOLS.items():
        declared = _declared(header)
        assert declared == set(listed), f"{header}: header and HEADER_SYMBOLS disagree: {declared ^ set(listed)}"
        for name in listed:
            assert name in exported, f"{name} (declared in {header}) is not exported"
            getattr(lib, name)
    assert lib.getCoeffSize() == 9822725 * 4          # pure host query, no device touched


def test_dispatch_library_exports_and_schedule_pairs_up():
    """include/srt_dispatch.h: every declared symbol is exported by libspleeterrt_dispatch.so, and the point-to-point schedule
    (the part of the NCCL dispatcher that can be wrong without a GPU) is consistent: for every ordered pair of ranks the sender's
    list of sends and the receiver's list of recvs agree element by element (group, stream, slot, count) - NCCL's matching
    rule - every stream of a non-root rank is scattered exactly once per channel and gathered once per output, root moves
    nothing to itself, and ragged lengths / more ranks than streams / one chunk per stream all hold."""
    import spleeterrt_b200 as srt
    declared = _declared("srt_dispatch.h")
    assert declared == set(srt.DISPATCH_SYMBOLS), declared ^ set(srt.DISPATCH_SYMBOLS)
    exported = srt.exported_symbols(srt.dispatch_lib_path())
    assert declared <= exported, declared - exported
    rng = np.random.default_rng(3)
    for world, root, n_streams, pairs, chunks in [(2, 0, 5, 4, 2), (8, 0, 37, 5, 4), (4, 2, 3, 2, 3), (8, 0, 1024, 5, 4), (3, 1, 7, 1, 8)]:
        n = rng.integers(1, 500000, n_streams)
        sched = [srt.dispatch_schedule(world, r, root, n, pairs, chunks) for r in range(world)]
        for a in range(world):
            for b in range(world):
                if a == b:
                    continue
                sends = [tuple(x[[0, 3, 4, 5]]) for x in sched[a] if x[1] == 0 and x[2] == b]
                recvs = [tuple(x[[0, 3, 4, 5]]) for x in sched[b] if x[1] == 1 and x[2] == a]
                assert sends == recvs, (world, root, a, b)
                if a != root and b != root:
                    assert not sends
        assert all(x[2] != root for x in sched[root])                                   # nothing to itself
        for r in range(world):
            if r == root:
                continue
            ids = list(range(r, n_streams, world))
            got_in = sorted((int(x[3]), int(x[4])) for x in sched[r] if x[1] == 1)
            got_out = sorted((int(x[3]), int(x[4])) for x in sched[r] if x[1] == 0)
            assert got_in == sorted((i, s) for i in ids for s in (0, 1))
            assert got_out == sorted((i, 2 + s) for i in ids for s in range(2 * pairs))
            assert all(int(x[5]) == int(n[x[3]]) for x in sched[r])
            groups = [int(x[0]) for x in sched[r]]
            assert groups == sorted(groups) and (not groups or max(groups) < 2 * chunks)  # scatter chunks, then gather chunks


def test_dispatch_peer_layout():
    """Peer-memory mode: the layout of root's buffers (host only): 16-byte aligned, disjoint, in stream order."""
    import spleeterrt_b200 as srt
    lib = srt.load_dispatch_library()
    n = [441000, 1, 4097, 300001]
    arr = (C.c_size_t * 4)(*n)
    io, oo = (C.c_size_t * 4)(), (C.c_size_t * 4)()
    fi, fo = C.c_size_t(), C.c_size_t()
    assert lib.srt_dispatch_peer_layout(arr, 4, 5, io, oo, C.byref(fi), C.byref(fo)) == 0
    pad = [(x + 3) // 4 * 4 for x in n]
    assert list(io) == [0, 2 * pad[0], 2 * (pad[0] + pad[1]), 2 * (pad[0] + pad[1] + pad[2])] and fi.value == 2 * sum(pad)
    assert list(oo) == [10 * sum(pad[:k]) for k in range(4)] and fo.value == 10 * sum(pad)
    assert all(v % 4 == 0 for v in list(io) + list(oo))


def test_no_cpu_fallback_without_gpu(oracle):
    """Without a CUDA device the product must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import spleeterrt_b200 as srt
    with pytest.raises(srt.SrtError):
        srt.Separator([(oracle.synthetic_weights(1), 1)], 64, 64)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "spleeterrt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "ORACLE_DOC/") or "import" not in txt.split("oracle")[0][-40:], f
                assert "from oracle" not in txt and "import oracle" not in txt and "srt_oracle" not in txt, f


# ----------------------------------------------------------------------------- weight file formats
def test_coeff_dat_round_trip_and_errors(oracle, tmp_path):
    """fp32 `.dat` dumps (PluginProcessor.cpp:48-80): exact round trip; a short or missing file is an error, not
    uninitialised weights."""
    import spleeterrt_b200 as srt
    w = oracle.synthetic_weights(21) * np.float32(1.0001)       # not fp16-representable any more
    p = tmp_path / "drum4stems.dat"
    srt.save_coeff_dat(p, w)
    assert os.path.getsize(p) == 39290900
    assert np.array_equal(np.fromfile(p, np.float32), w)         # the format is the raw struct
    assert np.array_equal(srt.load_coeff_dat(p), w)
    short = tmp_path / "short.dat"
    short.write_bytes(b"\0" * 1000)
    with pytest.raises(srt.SrtError):
        srt.load_coeff_dat(short)
    with pytest.raises(srt.SrtError):
        srt.load_coeff_dat(tmp_path / "missing.dat")


def test_model_fp16_blob_matches_reference_expansion(oracle, tmp_path):
    """spleeterQuantized blob -> nets, expanded like f32Decompress (main.c:423-443): denormal halves flush to zero."""
    import spleeterrt_b200 as srt
    rng = np.random.default_rng(5)
    halves = rng.integers(0, 1 << 16, size=2 * srt.COEFF_FLOATS, dtype=np.uint16)
    halves[(halves & 0x7c00) == 0x7c00] &= 0xbfff                # no inf/nan: the model has none
    halves[:8] = [0x0001, 0x03ff, 0x8001, 0x83ff, 0x0400, 0x8400, 0x0000, 0x8000]   # denormals, smallest normals, zeros
    p = tmp_path / "model_fp16.bin"
    halves.tofile(p)
    nets = srt.load_model_fp16(p)
    assert len(nets) == 2
    want = oracle.half_to_float(halves).reshape(2, -1)           # the oracle's restatement of f32Decompress
    assert np.array_equal(nets[0].view(np.uint32), want[0].view(np.uint32))
    assert np.array_equal(nets[1].view(np.uint32), want[1].view(np.uint32))
    assert np.all(nets[0][:4] == 0) and nets[0][4] == np.float32(2.0 ** -14)
    bad = tmp_path / "odd.bin"
    halves[:1000].tofile(bad)
    with pytest.raises(srt.SrtError):
        srt.load_model_fp16(bad)


@pytest.mark.parametrize("layer,name,form", [(0, "down2", 0), (0, "down2", 1), (3, "down5", 0), (5, "up1", 0), (9, "up5", 0), (9, "up5", 1)])
def test_pack_layer_is_a_zero_padded_permutation(oracle, layer, name, form):
    """srt_pack_layer: every weight of the layer lands exactly once in the K-major swizzled blob (the rest is the
    zero padding of unused taps / parities); non-TF32-exact weights double into hi + lo terms that sum back."""
    import spleeterrt_b200 as srt
    coeff = oracle.synthetic_weights(33)
    w = np.sort(oracle.coeff_views(coeff)[name + ".w"].ravel())
    blob = srt.pack_layer(coeff, layer, 64, 128, form)
    nz = np.sort(blob[blob != 0])
    assert np.array_equal(nz, w[w != 0])
    # fp32 weights that TF32 cannot hold: two terms per weight, hi + lo == w to fp32 rounding of the lo term
    coeff32 = coeff * np.float32(1.00013)
    blob2 = srt.pack_layer(coeff32, layer, 64, 128, form)
    assert blob2.size == 2 * blob.size
    w32 = oracle.coeff_views(coeff32)[name + ".w"].ravel().astype(np.float64)
    assert abs(blob2.astype(np.float64).sum() - w32.sum()) < 1e-6 * np.abs(w32).sum()
    with pytest.raises(srt.SrtError):
        srt.pack_layer(coeff, 11, 64, 128)


# ----------------------------------------------------------------------------- resampler bookkeeping
@pytest.mark.parametrize("ch", [1, 2])
def test_resample_plan_matches_the_oracle_converter(oracle, ch):
    """srt_resample_plan (host only): the product's restatement of the converter's sequential bookkeeping produces as
    many frames as the oracle (which is pinned bit-exactly to the reference), also where n*ratio is an integer and the
    stop condition is a knife edge; positions are monotone and table offsets stay inside one table step."""
    import spleeterrt_b200 as srt
    table = oracle.synthetic_resampler_table()
    rng = np.random.default_rng(1)
    for fs, n in [(48000, 4800), (48000, 160), (48000, 4801), (22050, 1000), (88200, 2000), (96000, 320 * 7), (32000, 3200), (11025, 500),
                  (44100 * 4, 4000), (47999, 5000), (8000, 80)]:
        ratio = 44100.0 / fs
        x = (rng.standard_normal((n, ch)) * 0.1).astype(np.float32)
        _, gen = oracle.resample(x[:, 0] if ch == 1 else x, ratio, table)
        frames, starts, n_out = srt.resample_plan(n, ch, ratio)
        assert n_out == int(np.ceil(n * ratio))
        assert frames.size == gen, (fs, n, ch, frames.size, gen)
        assert np.all(np.diff(frames) >= 0) and frames[0] == 0 and frames[-1] < n + 1
        inc = int(round(491 * min(ratio, 1.0) * 4096))
        assert starts.min() >= 0 and starts.max() <= inc
    with pytest.raises(srt.SrtError):
        srt.resample_plan(100, 3, 1.0)
    with pytest.raises(srt.SrtError):
        srt.resample_plan(100, 2, 1000.0)


def test_narrow_n_tiles_for_small_grids(oracle, host_model, small_nets):
    """build_plans(min_ctas): a one-tile batch gets N tiles of 64-128 columns for the deep layers (more, shorter CTAs);
    the gather-GEMM model over those plans (several N tiles per layer, also on decoder layers) still reproduces the
    oracle's layer tensors, and big batches keep the wide tiles."""
    info = (C.c_int * 10)()
    host_model.srt_host_model_set_min_ctas(148)
    try:
        host_model.srt_host_model_plan_info(512, 1024, 1, 4, info)       # down6, one image: 512 couts -> 8 tiles of 64
        assert (info[3], info[4]) == (64, 8)
        host_model.srt_host_model_plan_info(512, 1024, 1, 5, info)       # up1: 4 phases x 1 tile -> 256 couts in 4 tiles
        assert (info[3], info[4]) == (64, 4)
        host_model.srt_host_model_plan_info(512, 1024, 32, 4, info)      # 32 images: 32 x 2 = 64 CTAs < 148 -> 128-wide tiles
        assert info[3] * info[4] == 512 and info[3] >= 64
        host_model.srt_host_model_plan_info(512, 1024, 32, 9, info)      # up5 at full batch: untouched (N = 4 x 16: parities fused)
        assert (info[3], info[4], info[5]) == (64, 1, 1)
        host_model.srt_host_model_plan_info(512, 1024, 1, 7, info)       # up3, one image: 16 fused CTAs would not fill the SMs -> 4 phases
        assert info[5] == 4 and info[3] == 64
        T, F = 64, 128
        coeff = small_nets[0][0]
        x = (np.abs(np.random.default_rng(5).standard_normal((2, T, F))) * 3).astype(np.float32)
        _, tp = oracle.unet(coeff, x, 1, taps=True)
        taps = oracle.split_taps(tp, T, F)
        v = oracle.coeff_views(coeff)
        for i in (4, 5):                                                   # down5, down6
            bn = v[f"down{i}.bn"]
            act_in = _act(3, bn[1][:, None, None] * taps[f"skip{i}"] + bn[0][:, None, None]).astype(np.float32)
            ref = taps[f"skip{i+1}"]
            got = _run_layer(host_model, T, F, i - 1, coeff, 3, act_in, None, ref.shape)
            assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-5
        for d in (0, 1, 2):                                                # up1, up2, up3
            s0 = taps["skip6"] if d == 0 else taps[f"skip{6-d}"]
            s1 = None if d == 0 else taps[f"up{d}"]
            ref = taps[f"up{d+1}"]
            got = _run_layer(host_model, T, F, 5 + d, coeff, 3, s0, s1, ref.shape)
            assert np.abs(got - ref).max() / max(1e-6, np.abs(ref).max()) < 2e-5
    finally:
        host_model.srt_host_model_set_min_ctas(0)


@pytest.mark.parametrize("T,F,mode", [(64, 128, 1), (64, 192, 0)])
def test_up6_packed_weights_and_8bit_residual_term(oracle, host_model, small_nets, T, F, mode):
    """up6 as up6_tc_kernel contracts it - the packed weight block of srt_plan.cpp read back through the hardware's SWIZZLE_32B
    definition, operands truncated to TF32, residual term through e5m2 (default) or TF32 operands - then the 25-value gather,
    against the oracle's up6 tensor (spleeter.c:289-295)."""
    coeff = small_nets[0][0] if mode else small_nets[1][0]
    rng = np.random.default_rng(7 * T + F)
    x = (np.abs(rng.standard_normal((2, T, F))) * 3).astype(np.float32)
    _, tp = oracle.unet(coeff, x, mode, taps=True)
    taps = oracle.split_taps(tp, T, F)
    ref = taps["up6"][0]
    e1 = np.ascontiguousarray(taps["skip1"], np.float32)
    u5 = np.ascontiguousarray(taps["up5"], np.float32)
    a_dec = 3 if mode else 2
    errs = {}
    for lo8 in (1, 0):
        out = np.zeros((T, F), np.float32)
        rc = host_model.srt_host_model_up6(T, F, np.ascontiguousarray(coeff).ctypes.data_as(C.c_void_p), a_dec, e1.ctypes.data_as(C.c_void_p),
                                           u5.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), lo8)
        assert rc == 0
        errs[lo8] = np.abs(out - ref).max() / max(1e-6, np.abs(ref).max())
    # single-pass TF32 operands (truncated) would sit near 1e-3.  The fp32 residual term brings that to fp32 rounding; the 8-bit term
    # (residual and weight each good to ~3 bits, and truncation makes the residual one-signed, so the weight error does not average
    # out) to ~6e-5 of the tensor's range - the level of the tensor-core layers before it, 4e-6 on the stems
    assert errs[0] < 2e-6, errs
    assert errs[1] < 1.5e-4, errs
