"""ctypes front-end for the parity oracle.

TEST INFRASTRUCTURE ONLY — may be imported from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never from spleeterrt_b200/.

Two back ends:
  * ``port``  — oracle/srt_oracle.c, our own restatement (always available; built on demand
                with gcc into oracle/libsrt_oracle.so)
  * ``ref``   — the reference's own C code compiled by oracle/build_ref.py into
                oracle/_ref/libref_exec.so / libref_vst.so (present when built in the
                container; the files travel to the GPU box)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FFT, HOP, BINS = 4096, 1024, 2049
COEFF_FLOATS = 9822725

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build_port(force=False):
    so = os.path.join(HERE, "libsrt_oracle.so")
    src = os.path.join(HERE, "srt_oracle.c")
    tbl = os.path.join(HERE, "sigmoid_table.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(tbl)):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=gnu11",
                               "-ffp-contract=off", "-o", so, src, "-lm"])
    return so


_port = None


def port():
    global _port
    if _port is None:
        lib = C.CDLL(build_port())
        lib.srt_oracle_coeff_floats.restype = C.c_size_t
        lib.srt_oracle_taps_floats.restype = C.c_size_t
        lib.srt_oracle_taps_floats.argtypes = [C.c_int, C.c_int]
        lib.srt_oracle_unet.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_void_p]
        lib.srt_oracle_sigmoid_lut.restype = C.c_float
        lib.srt_oracle_sigmoid_lut.argtypes = [C.c_float]
        lib.srt_oracle_sigmoid_table.restype = C.POINTER(C.c_float)
        lib.srt_oracle_prewindow.restype = C.POINTER(C.c_float)
        lib.srt_oracle_postwindow.restype = C.POINTER(C.c_float)
        lib.srt_oracle_stft.restype = C.c_size_t
        lib.srt_oracle_stft.argtypes = [_f32p, _f32p, C.c_size_t, _f32p, _f32p, _f32p, _f32p]
        lib.srt_oracle_istft.restype = C.c_size_t
        lib.srt_oracle_istft.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_size_t, _f32p, _f32p]
        lib.srt_oracle_padded_len.restype = C.c_size_t
        lib.srt_oracle_padded_len.argtypes = [C.c_size_t]
        lib.srt_oracle_separate.restype = C.c_int
        lib.srt_oracle_half_to_float.argtypes = [C.c_void_p, _f32p, C.c_size_t]
        assert lib.srt_oracle_coeff_floats() == COEFF_FLOATS
        _port = lib
    return _port


# ----------------------------------------------------------------------------- weights
WEIGHTS_BLOB = os.path.join(os.path.dirname(HERE), "spleeterrt_b200", "weights", "model_fp16.bin")


def have_real_weights():
    return os.path.exists(WEIGHTS_BLOB)


def real_weights_fp16():
    """uint16[2][9822725]: net 0 = ELU 'drum' net, net 1 = LeakyReLU/ReLU 'vocal' net
    (Executable/main.c:759-760)."""
    h = np.fromfile(WEIGHTS_BLOB, dtype=np.uint16)
    return h.reshape(2, COEFF_FLOATS)


def half_to_float(h):
    """f32Decompress (main.c:423-434): IEEE half -> float, denormals as zero."""
    h = np.ascontiguousarray(h, dtype=np.uint16)
    out = np.empty(h.shape, np.float32)
    port().srt_oracle_half_to_float(h.ctypes.data, out.reshape(-1), h.size)
    return out


# Per-layer (name, weight shape, has bias, bn channels) in spleeterCoeff order.
ENC_CH = [2, 16, 32, 64, 128, 256, 512]
DEC_IN = [512, 512, 256, 128, 64, 32]
DEC_OUT = [256, 128, 64, 32, 16, 1]


def coeff_views(coeff):
    """Split one flat float32[9822725] net into named views (no copies)."""
    coeff = np.asarray(coeff)
    out, p = {}, 0

    def take(name, shape):
        nonlocal p
        n = int(np.prod(shape))
        out[name] = coeff[p:p + n].reshape(shape)
        p += n
    for i in range(6):
        take(f"down{i+1}.w", (ENC_CH[i + 1], ENC_CH[i], 5, 5))
        take(f"down{i+1}.b", (ENC_CH[i + 1],))
        if i < 5:
            take(f"down{i+1}.bn", (2, ENC_CH[i + 1]))
    for i in range(6):
        take(f"up{i+1}.w", (DEC_IN[i], DEC_OUT[i], 5, 5))
        take(f"up{i+1}.b", (DEC_OUT[i],))
        take(f"up{i+1}.bn", (2, DEC_OUT[i]))
    take("up7.w", (2, 1, 4, 4))
    take("up7.b", (2,))
    assert p == COEFF_FLOATS
    return out


def synthetic_weights(seed, scale=1.0):
    """Random net with fp16-representable values and activation-preserving scales
    (He-style), used when the reference's model blob is unavailable or for extra stems."""
    rng = np.random.default_rng(seed)
    c = np.zeros(COEFF_FLOATS, np.float32)
    v = coeff_views(c)
    for k, a in v.items():
        if k.endswith(".w"):
            if k.startswith("down") or k.startswith("up7"):
                fan = a.shape[1] * a.shape[2] * a.shape[3]
            else:
                fan = a.shape[0] * a.shape[2] * a.shape[3] / 4.0
            a[...] = rng.normal(0, scale * np.sqrt(1.5 / fan), a.shape)
        elif k.endswith(".b"):
            a[...] = rng.normal(0, 0.05, a.shape)
        else:
            a[0] = rng.normal(0, 0.05, a.shape[1:])
            a[1] = 1.0 + rng.normal(0, 0.05, a.shape[1:])
    return c.astype(np.float16).astype(np.float32)


def jitter_weights(base, seed):
    """SURVEY §8d recipe for the stems the reference does not ship: multiply every conv /
    tconv weight of a real net by (1 + 0.05 N(0,1)), round to fp16 and back."""
    rng = np.random.default_rng(seed)
    c = np.array(base, dtype=np.float32, copy=True)
    for k, a in coeff_views(c).items():
        if k.endswith(".w"):
            a *= (1.0 + 0.05 * rng.standard_normal(a.shape)).astype(np.float32)
    f16 = c.astype(np.float16)
    f16[np.abs(f16) < np.float16(6.104e-05)] = 0  # denormals-as-zero like f32Decompress
    return f16.astype(np.float32)


def four_stem_weights():
    """[(coeff fp32, stemMode)] x4 in the VST stem order drum, bass, accompaniment, vocal
    (PluginProcessor.cpp:50-53); all ELU as the VST does (Spleeter4Stems.c:444-447).
    Real nets where the reference ships them, seeded jitter of a real net otherwise."""
    if have_real_weights():
        w = half_to_float(real_weights_fp16())
        nets = [w[0], jitter_weights(w[0], 777 + 2), jitter_weights(w[1], 777 + 3), w[1]]
    else:
        nets = [synthetic_weights(100 + k) for k in range(4)]
    return [(np.ascontiguousarray(n), 1) for n in nets]


def synth_pcm(stream, n=441000):
    """SURVEY §8d synthetic input: two tones + noise, float32 in [-1, 1]."""
    t = np.arange(n) / 44100.0
    out = []
    for seed in (1234 + stream, 1234 + stream + 10000):
        rng = np.random.default_rng(seed)
        x = (0.25 * np.sin(2 * np.pi * 220 * t)
             + 0.15 * np.sin(2 * np.pi * 3300 * t * (1 + 0.1 * np.sin(2 * np.pi * 0.5 * t)))
             + 0.05 * rng.standard_normal(n))
        out.append(np.clip(x, -1, 1).astype(np.float32))
    return out


# ----------------------------------------------------------------------------- port API
def unet(coeff, x, stem_mode, flavour=0, taps=False):
    """x: float32[2][T][F] -> mask float32[2][T][F] (and the layer taps if asked)."""
    x = np.ascontiguousarray(x, np.float32)
    _, T, F = x.shape
    y = np.empty_like(x)
    tp = None
    if taps:
        tp = np.empty(port().srt_oracle_taps_floats(F, T), np.float32)
    port().srt_oracle_unet(np.ascontiguousarray(coeff, np.float32), F, T, stem_mode, flavour,
                           x.reshape(-1), y.reshape(-1), tp.ctypes.data if taps else None)
    return (y, tp) if taps else y


def split_taps(tp, T, F):
    """taps blob -> dict name -> [C][H][W]"""
    out, p = {}, 0
    for i in range(6):
        h, w, c = T >> (i + 1), F >> (i + 1), ENC_CH[i + 1]
        out[f"skip{i+1}"] = tp[p:p + c * h * w].reshape(c, h, w)
        p += c * h * w
    for i in range(6):
        h, w, c = T >> (5 - i), F >> (5 - i), DEC_OUT[i]
        out[f"up{i+1}"] = tp[p:p + c * h * w].reshape(c, h, w)
        p += c * h * w
    return out


def stft(L, R):
    L = np.ascontiguousarray(L, np.float32)
    R = np.ascontiguousarray(R, np.float32)
    rows = (L.size + HOP - 1) // HOP
    planes = [np.zeros((rows, FFT), np.float32) for _ in range(4)]
    port().srt_oracle_stft(L, R, L.size, *[p.reshape(-1) for p in planes])
    return planes


def istft(reL, imL, reR, imR):
    frames = reL.shape[0]
    n = frames * HOP + FFT - HOP
    oL, oR = np.zeros(n, np.float32), np.zeros(n, np.float32)
    port().srt_oracle_istft(*[np.ascontiguousarray(a, np.float32).reshape(-1) for a in (reL, imL, reR, imR)],
                            frames, oL, oR)
    return oL, oR


def separate(nets, pcmL, pcmR, T, F, unaffected=0.1, flavour=0, want_masks=False):
    """nets: [(coeff, stemMode)].  Returns float32[nStems][2][n] (and masks)."""
    lib = port()
    pcmL = np.ascontiguousarray(pcmL, np.float32)
    pcmR = np.ascontiguousarray(pcmR, np.float32)
    n, ns = pcmL.size, len(nets)
    coeffs = [np.ascontiguousarray(c, np.float32) for c, _ in nets]
    cp = (C.c_void_p * ns)(*[c.ctypes.data for c in coeffs])
    modes = (C.c_int * ns)(*[m for _, m in nets])
    out = np.zeros((ns, 2, n), np.float32)
    op = (C.c_void_p * (2 * ns))(*[out[s, c].ctypes.data for s in range(ns) for c in range(2)])
    frames = lib.srt_oracle_padded_len(n) // HOP
    tiles = (frames + T - 1) // T
    masks = np.zeros((ns, tiles, 2, T, F), np.float32) if want_masks else None
    lib.srt_oracle_separate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _f32p, _f32p, C.c_size_t,
                                        C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    lib.srt_oracle_separate(cp, modes, ns, flavour, pcmL, pcmR, n, T, F, unaffected, op,
                            masks.ctypes.data if want_masks else None)
    return (out, masks) if want_masks else out


def separate_cli(nets, pcmL, pcmR, T, F, n_out, unaffected=0.1):
    """The CLI's output modes (main.c:776-970).  n_out = 2: nets = [vocal net] -> float32[2][2][n]
    (vocal, accompaniment); n_out = 3: nets = [drum net, vocal net] -> float32[3][2][n] (drum, vocal, accompaniment)."""
    lib = port()
    pcmL = np.ascontiguousarray(pcmL, np.float32)
    pcmR = np.ascontiguousarray(pcmR, np.float32)
    n, ns = pcmL.size, len(nets)
    assert ns == n_out - 1
    coeffs = [np.ascontiguousarray(c, np.float32) for c, _ in nets]
    cp = (C.c_void_p * ns)(*[c.ctypes.data for c in coeffs])
    modes = (C.c_int * ns)(*[m for _, m in nets])
    out = np.zeros((n_out, 2, n), np.float32)
    op = (C.c_void_p * (2 * n_out))(*[out[s, c].ctypes.data for s in range(n_out) for c in range(2)])
    lib.srt_oracle_separate_cli.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _f32p, _f32p, C.c_size_t,
                                            C.c_int, C.c_int, C.c_float, C.c_void_p]
    rc = lib.srt_oracle_separate_cli(cp, modes, n_out, pcmL, pcmR, n, T, F, unaffected, op)
    assert rc == 0
    return out


# ----------------------------------------------------------------------------- resampler
RESAMPLER_TABLE = os.path.join(os.path.dirname(HERE), "spleeterrt_b200", "weights", "resampler_mq.f32")
RS_HALF_LEN, RS_INDEX_INC = 22436, 491     # src_sinc.c:141-143


def resampler_table():
    """The sinc coefficient table the reference host decompresses at start-up (main.c:693-694), dumped from the
    reference build by oracle/build_ref.py (reference data, not committed)."""
    return np.fromfile(RESAMPLER_TABLE, np.float32)


def have_resampler_table():
    return os.path.exists(RESAMPLER_TABLE)


def synthetic_resampler_table(seed=3):
    """A windowed-sinc table of the same geometry for runs without the reference's data."""
    i = np.arange(RS_HALF_LEN + 2, dtype=np.float64)
    t = i / RS_INDEX_INC
    w = 0.5 * (1 + np.cos(np.pi * i / (RS_HALF_LEN + 2)))
    return (0.92 * np.sinc(0.92 * t) * w).astype(np.float32)


def resample(x, ratio, table=None):
    """x: float32[n] (mono) or [n][2] interleaved frames -> (float32[ceil(n*ratio)][ch] zero-filled, frames generated),
    JamesDSPOfflineResampling (main.c:209-224, 264-270)."""
    lib = port()
    x = np.ascontiguousarray(x, np.float32)
    ch = 1 if x.ndim == 1 else x.shape[1]
    n = x.shape[0]
    table = np.ascontiguousarray(resampler_table() if table is None else table, np.float32)
    n_out = int(np.ceil(n * ratio))
    out = np.zeros((n_out, ch), np.float32)
    lib.srt_oracle_resample.restype = C.c_long
    lib.srt_oracle_resample.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_long]
    gen = lib.srt_oracle_resample(x.ctypes.data, n, ch, float(ratio), table.ctypes.data, RS_HALF_LEN, RS_INDEX_INC,
                                  out.ctypes.data, n_out)
    return out, int(gen)


def ref_resample(x, ratio, table=None):
    """The reference's own resampler (oracle/_ref/libref_resample.so)."""
    lib = C.CDLL(os.path.join(REF_DIR, "libref_resample.so"))
    x = np.ascontiguousarray(x, np.float32)
    ch = 1 if x.ndim == 1 else x.shape[1]
    n = x.shape[0]
    if table is None:
        table = np.zeros(RS_HALF_LEN + 2, np.float32)
        lib.ref_resampler_table(table.ctypes.data_as(C.c_void_p))
    table = np.ascontiguousarray(table, np.float32)
    n_out = int(np.ceil(n * ratio))
    out = np.zeros((n_out, ch), np.float32)
    lib.ref_resample.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_double, C.c_void_p]
    lib.ref_resample(x.ctypes.data, out.ctypes.data, n, n_out, ch, float(ratio), table.ctypes.data)
    return out


def have_ref_resampler():
    return os.path.exists(os.path.join(REF_DIR, "libref_resample.so"))


# ----------------------------------------------------------------------------- reference API
class RefExec:
    """The reference's Executable flavour, compiled as is (oracle/_ref/libref_exec.so)."""

    class _STFT(C.Structure):
        # Executable/stftFix.h:19-31
        _fields_ = [("mBitRev", C.c_uint * FFT), ("mPreWindow", C.c_float * FFT),
                    ("mPostWindow", C.c_float * FFT), ("mSineTab", C.c_float * FFT),
                    ("threads", C.c_void_p), ("stftThreadData", C.c_void_p), ("istftThreadData", C.c_void_p),
                    ("targetCore", C.c_size_t), ("_data", C.c_void_p * 2), ("shared_info", C.c_void_p)]

    def __init__(self, blas=False):
        """blas=True: oracle/_ref/libref_exec_blas.so, the same sources with gemm.c on its real backend (cblas_sgemm, here
        OpenBLAS; oracle/build_ref.py build_blas) instead of the naive -DCPU_GEMM loops."""
        path = os.path.join(REF_DIR, "libref_exec_blas.so" if blas else "libref_exec.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        lib = C.CDLL(path)
        self.blas = blas
        lib.allocateSpleeterStr.restype = C.c_void_p
        lib.initSpleeter.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
        lib.processSpleeter.argtypes = [C.c_void_p, _f32p, _f32p]
        lib.freeSpleeter.argtypes = [C.c_void_p]
        lib.getCoeffSize.restype = C.c_size_t
        lib.fastSigmoid.restype = C.c_float
        lib.fastSigmoid.argtypes = [C.c_float]
        lib.stft.restype = C.c_size_t
        lib.stft.argtypes = [C.c_void_p, _f32p, _f32p, C.c_size_t] + [C.POINTER(C.POINTER(C.c_float))] * 4
        lib.istft.restype = C.c_size_t
        lib.istft.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t] + \
                             [C.POINTER(C.POINTER(C.c_float))] * 2
        lib.InitSTFT.argtypes = [C.c_void_p, C.c_size_t]
        lib.FreeSTFT.argtypes = [C.c_void_p]
        self.lib = lib
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        assert lib.getCoeffSize() == COEFF_FLOATS * 4

    def unet(self, coeff, x, stem_mode):
        x = np.ascontiguousarray(x, np.float32)
        _, T, F = x.shape
        coeff = np.ascontiguousarray(coeff, np.float32)
        nn = self.lib.allocateSpleeterStr()
        self.lib.initSpleeter(nn, F, T, stem_mode, coeff.ctypes.data)
        y = np.empty_like(x)
        self.lib.processSpleeter(nn, x.reshape(-1), y.reshape(-1))
        self.lib.freeSpleeter(nn)
        self.libc.free(nn)
        return y

    def unet_instance(self, coeff, T, F, stem_mode):
        """persistent instance for timing: returns callable(x)->y"""
        coeff = np.ascontiguousarray(coeff, np.float32)
        nn = self.lib.allocateSpleeterStr()
        self.lib.initSpleeter(nn, F, T, stem_mode, coeff.ctypes.data)
        y = np.empty((2, T, F), np.float32)

        def run(x):
            self.lib.processSpleeter(nn, x.reshape(-1), y.reshape(-1))
            return y
        run._keep = (coeff, nn)
        return run

    def stft(self, L, R, threads=1):
        L = np.ascontiguousarray(L, np.float32)
        R = np.ascontiguousarray(R, np.float32)
        st = self._STFT()
        self.lib.InitSTFT(C.byref(st), threads)
        ptrs = [C.POINTER(C.c_float)() for _ in range(4)]
        rows = self.lib.stft(C.byref(st), L, R, L.size, *[C.byref(p) for p in ptrs])
        planes = [np.ctypeslib.as_array(p, shape=(rows, FFT)).copy() for p in ptrs]
        for p in ptrs:
            self.libc.free(C.cast(p, C.c_void_p))
        self.lib.FreeSTFT(C.byref(st))
        return planes

    def istft(self, reL, imL, reR, imR, threads=1):
        planes = [np.array(a, dtype=np.float32, order="C", copy=True) for a in (reL, imL, reR, imR)]
        frames = planes[0].shape[0]
        st = self._STFT()
        self.lib.InitSTFT(C.byref(st), threads)
        oL, oR = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
        n = self.lib.istft(C.byref(st), *[p.ctypes.data for p in planes], frames, C.byref(oL), C.byref(oR))
        outs = [np.ctypeslib.as_array(p, shape=(n,)).copy() for p in (oL, oR)]
        for p in (oL, oR):
            self.libc.free(C.cast(p, C.c_void_p))
        self.lib.FreeSTFT(C.byref(st))
        return outs

    def separate(self, nets, pcmL, pcmR, T, F, unaffected=0.1):
        """The CLI's flow (main.c:762-806, processMT single-thread branch :447-541) driven
        from Python on the reference's own stft / processSpleeter / istft."""
        n = len(pcmL)
        padded = FFT * ((n + FFT - 1) // FFT) + 2 * FFT
        pl, pr = np.zeros(padded, np.float32), np.zeros(padded, np.float32)
        pl[FFT:FFT + n] = pcmL
        pr[FFT:FFT + n] = pcmR
        spec = self.stft(pl, pr)
        frames = spec[0].shape[0]
        out = np.zeros((len(nets), 2, n), np.float32)
        for s, (coeff, mode) in enumerate(nets):
            run = self.unet_instance(coeff, T, F, mode)
            ms = [p.copy() for p in spec]
            for f0 in range(0, frames, T):
                nt = min(T, frames - f0)
                mag = np.zeros((2, T, F), np.float32)
                for c in range(2):
                    mag[c, :nt] = np.hypot(spec[2 * c][f0:f0 + nt, :F], spec[2 * c + 1][f0:f0 + nt, :F]) * np.float32(FFT)
                mask = run(mag)
                for c in range(2):
                    for q in (2 * c, 2 * c + 1):
                        ms[q][f0:f0 + nt, :F] *= mask[c, :nt]
                        ms[q][f0:f0 + nt, F:BINS] *= np.float32(unaffected)
            oL, oR = self.istft(*ms)
            out[s, 0], out[s, 1] = oL[FFT:FFT + n], oR[FFT:FFT + n]
        return out


    def _net_over_spectrum(self, coeff, mode, planes, T, F, unaffected):
        """processMT's single-thread branch (main.c:447-541) on the reference's own processSpleeter, in place."""
        run = self.unet_instance(coeff, T, F, mode)
        frames = planes[0].shape[0]
        for f0 in range(0, frames, T):
            nt = min(T, frames - f0)
            mag = np.zeros((2, T, F), np.float32)
            for c in range(2):
                mag[c, :nt] = np.hypot(planes[2 * c][f0:f0 + nt, :F], planes[2 * c + 1][f0:f0 + nt, :F]) * np.float32(FFT)
            mask = run(mag)
            for c in range(2):
                for q in (2 * c, 2 * c + 1):
                    planes[q][f0:f0 + nt, :F] *= mask[c, :nt]
                    planes[q][f0:f0 + nt, F:BINS] *= np.float32(unaffected)

    def separate_cli(self, nets, pcmL, pcmR, T, F, n_out, unaffected=0.1):
        """The CLI's 2- and 3-output flows (main.c:776-970) on the reference's own stft / processSpleeter / istft."""
        n = len(pcmL)
        padded = FFT * ((n + FFT - 1) // FFT) + 2 * FFT
        pl, pr = np.zeros(padded, np.float32), np.zeros(padded, np.float32)
        pl[FFT:FFT + n] = pcmL
        pr[FFT:FFT + n] = pcmR
        spec = self.stft(pl, pr)
        cut = slice(FFT, FFT + n)
        out = np.zeros((n_out, 2, n), np.float32)
        if n_out == 2:
            self._net_over_spectrum(nets[0][0], nets[0][1], spec, T, F, unaffected)
            oL, oR = self.istft(*spec)
            out[0, 0], out[0, 1] = oL[cut], oR[cut]
            out[1, 0], out[1, 1] = pl[cut] - oL[cut], pr[cut] - oR[cut]
            return out
        orig = [p.copy() for p in spec]
        self._net_over_spectrum(nets[0][0], nets[0][1], spec, T, F, unaffected)
        resid = [o - d for o, d in zip(orig, spec)]
        dL, dR = self.istft(*spec)
        avL, avR = self.istft(*resid)
        self._net_over_spectrum(nets[1][0], nets[1][1], resid, T, F, unaffected)
        vL, vR = self.istft(*resid)
        out[0, 0], out[0, 1] = dL[cut], dR[cut]
        out[1, 0], out[1, 1] = vL[cut], vR[cut]
        out[2, 0], out[2, 1] = (avL - vL)[cut], (avR - vR)[cut]
        return out


class PortVst:
    """oracle/srt_oracle.c restatement of the streamer (srt_oracle_vst_*)."""

    def __init__(self, nets, T, F, unaffected=None):
        lib = port()
        lib.srt_oracle_vst_create.restype = C.c_void_p
        lib.srt_oracle_vst_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.srt_oracle_vst_process.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_void_p]
        lib.srt_oracle_vst_destroy.argtypes = [C.c_void_p]
        self.lib, self.S = lib, len(nets)
        self.coeffs = [np.ascontiguousarray(c, np.float32) for c, _ in nets]
        cp = (C.c_void_p * self.S)(*[c.ctypes.data for c in self.coeffs])
        uw = (C.c_float * self.S)(*unaffected) if unaffected is not None else None
        self.h = lib.srt_oracle_vst_create(cp, self.S, T, F, uw)

    def process(self, L, R):
        L = np.ascontiguousarray(L, np.float32)
        R = np.ascontiguousarray(R, np.float32)
        out = np.full((2 * self.S, L.size), np.nan, np.float32)
        ptrs = (C.c_void_p * (2 * self.S))(*[out[j].ctypes.data for j in range(2 * self.S)])
        self.lib.srt_oracle_vst_process(self.h, L, R, L.size, ptrs)
        return out

    def close(self):
        if self.h:
            self.lib.srt_oracle_vst_destroy(self.h)
            self.h = None


class RefVst:
    """The reference's real-time streamer (VST/Source/Spleeter4Stems.c) compiled as is into
    oracle/_ref/libref_vst.so (naive CPU_GEMM backend).  One instance = one Spleeter4Stems object."""

    def __init__(self, nets, T, F):
        path = os.path.join(REF_DIR, "libref_vst.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.Spleeter4StemsInit.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self.lib.Spleeter4StemsProcessSamples.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_void_p]
        self.lib.Spleeter4StemsFree.argtypes = [C.c_void_p]
        assert len(nets) == 4
        self.coeffs = [np.ascontiguousarray(c, np.float32) for c, _ in nets]
        self.cp = (C.c_void_p * 4)(*[c.ctypes.data for c in self.coeffs])
        self.obj = C.create_string_buffer(4 << 20)          # >= sizeof(Spleeter4Stems) (~0.4 MB)
        self.lib.Spleeter4StemsInit(self.obj, F, T, self.cp)

    def process(self, L, R):
        """feed one block (<= 1024 samples); returns float32[8][n] with NaN where nothing was written"""
        L = np.ascontiguousarray(L, np.float32)
        R = np.ascontiguousarray(R, np.float32)
        out = np.full((8, L.size), np.nan, np.float32)
        ptrs = (C.c_void_p * 8)(*[out[j].ctypes.data for j in range(8)])
        self.lib.Spleeter4StemsProcessSamples(self.obj, L, R, L.size, ptrs)
        return out

    def close(self):
        if self.obj is not None:
            self.lib.Spleeter4StemsFree(self.obj)
            self.obj = None


_ref_exec = {}


def ref_exec(blas=False):
    if blas not in _ref_exec:
        _ref_exec[blas] = RefExec(blas)
    return _ref_exec[blas]


def have_ref_blas():
    return os.path.exists(os.path.join(REF_DIR, "libref_exec_blas.so"))


def set_blas_threads(n):
    """thread count of the OpenBLAS behind libref_exec_blas.so (what main.c:689 sets on Linux)"""
    lib = ref_exec(True).lib          # the OpenBLAS it is linked to is now in the process
    try:
        lib.openblas_set_num_threads(int(n))
        return True
    except AttributeError:
        return False


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref_exec.so"))
