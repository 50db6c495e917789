"""Benchmark workload definition: deterministic synthetic PCM and the 4-stem
weight set.  Product-side code: does not touch oracle/."""
import os

import numpy as np

from .api import COEFF_FLOATS, half_to_float

PKG = os.path.dirname(os.path.abspath(__file__))
ENC_CH = [2, 16, 32, 64, 128, 256, 512]
DEC_IN = [512, 512, 256, 128, 64, 32]
DEC_OUT = [256, 128, 64, 32, 16, 1]


def weight_shapes():
    """(name, shape) in spleeterCoeff order (Executable/spleeter.h:5-31)."""
    out = []
    for i in range(6):
        out.append((f"down{i+1}.w", (ENC_CH[i + 1], ENC_CH[i], 5, 5)))
        out.append((f"down{i+1}.b", (ENC_CH[i + 1],)))
        if i < 5:
            out.append((f"down{i+1}.bn", (2, ENC_CH[i + 1])))
    for i in range(6):
        out.append((f"up{i+1}.w", (DEC_IN[i], DEC_OUT[i], 5, 5)))
        out.append((f"up{i+1}.b", (DEC_OUT[i],)))
        out.append((f"up{i+1}.bn", (2, DEC_OUT[i])))
    out += [("up7.w", (2, 1, 4, 4)), ("up7.b", (2,))]
    return out


def views(coeff):
    out, p = {}, 0
    for name, shape in weight_shapes():
        n = int(np.prod(shape))
        out[name] = coeff[p:p + n].reshape(shape)
        p += n
    assert p == COEFF_FLOATS
    return out


def model_blob_path():
    for p in (os.environ.get("SRT_MODEL_BLOB", ""), os.path.join(PKG, "weights", "model_fp16.bin")):
        if p and os.path.exists(p):
            return p
    return None


def synthetic_net(seed):
    rng = np.random.default_rng(seed)
    c = np.zeros(COEFF_FLOATS, np.float32)
    for k, a in views(c).items():
        if k.endswith(".w"):
            fan = a.shape[1] * 25 if (k.startswith("down")) else (a.shape[0] * 25 / 4.0 if k.startswith("up") and k != "up7.w" else 16)
            a[...] = rng.normal(0, np.sqrt(1.5 / fan), a.shape)
        elif k.endswith(".b"):
            a[...] = rng.normal(0, 0.05, a.shape)
        else:
            a[0] = rng.normal(0, 0.05, a.shape[1:])
            a[1] = 1.0 + rng.normal(0, 0.05, a.shape[1:])
    return c.astype(np.float16).astype(np.float32)


def jitter(base, seed):
    """Stems the reference does not ship: real net x (1 + 0.05 N(0,1)) per conv weight, re-rounded to fp16."""
    rng = np.random.default_rng(seed)
    c = np.array(base, dtype=np.float32, copy=True)
    for k, a in views(c).items():
        if k.endswith(".w"):
            a *= (1.0 + 0.05 * rng.standard_normal(a.shape)).astype(np.float32)
    f16 = c.astype(np.float16)
    f16[np.abs(f16) < np.float16(6.104e-05)] = 0
    return f16.astype(np.float32)


def four_stem_nets():
    """[(coeff, stemMode)] x 4: drum, bass, accompaniment, vocal (PluginProcessor.cpp:50-53), all ELU
    (Spleeter4Stems.c:444-447).  Returns (nets, description)."""
    blob = model_blob_path()
    if blob:
        w = half_to_float(np.fromfile(blob, dtype=np.uint16).reshape(2, COEFF_FLOATS))
        nets = [w[0], jitter(w[0], 777 + 2), jitter(w[1], 777 + 3), w[1]]
        desc = "reference model.7z nets (drum, vocal) + 2 seeded +-5% jitter nets, fp16-representable"
    else:
        nets = [synthetic_net(100 + k) for k in range(4)]
        desc = "seeded random fp16-representable nets (reference blob unavailable)"
    return [(np.ascontiguousarray(n), 1) for n in nets], desc


def stem_nets(n_stems):
    """four_stem_nets() extended for BASELINE.json config 4 (5 stems: vocals / drums / bass / piano / other): the reference
    has no 5-stem model, so stem k >= 4 is one more seeded jitter of a real net (SURVEY §8d recipe), ELU like the rest."""
    nets, desc = four_stem_nets()
    if n_stems <= 4:
        return nets[:n_stems], desc
    base = [c for c, _ in nets]
    for k in range(4, n_stems):
        src = base[k % 2 * 3]                      # alternate between the two real nets (index 0 and 3)
        nets.append((np.ascontiguousarray(jitter(src, 777 + k) if model_blob_path() else synthetic_net(100 + k)), 1))
    return nets, desc + f"; stems 5..{n_stems}: further seeded jitter nets"


def synth_pcm(stream, n=441000):
    t = np.arange(n) / 44100.0
    out = []
    for seed in (1234 + stream, 1234 + stream + 10000):
        rng = np.random.default_rng(seed)
        x = (0.25 * np.sin(2 * np.pi * 220 * t)
             + 0.15 * np.sin(2 * np.pi * 3300 * t * (1 + 0.1 * np.sin(2 * np.pi * 0.5 * t)))
             + 0.05 * rng.standard_normal(n))
        out.append(np.clip(x, -1, 1).astype(np.float32))
    return out


def synth_pcm_fullscale(stream, n=441000):
    """A full-scale programme-like clip for the precision tests: three tones (one gliding), an amplitude-modulated
    partial and noise, normalised to peak 0.999 (RMS ~0.3).  Errors of the TF32 path scale with the signal level, so
    this - not the -12 dBFS §8d signal - is the input that decides whether 1e-4 RMS per stem holds."""
    t = np.arange(n) / 44100.0
    out = []
    for seed in (4321 + stream, 4321 + stream + 10000):
        rng = np.random.default_rng(seed)
        x = (0.35 * np.sin(2 * np.pi * 220 * t + rng.uniform(0, 6.28))
             + 0.25 * np.sin(2 * np.pi * 3300 * t * (1 + 0.1 * np.sin(2 * np.pi * 0.5 * t)))
             + 0.10 * np.sin(2 * np.pi * 82.4 * t)
             + 0.15 * np.sin(2 * np.pi * 987 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 2.0 * t))
             + 0.12 * rng.standard_normal(n))
        out.append((x * (0.999 / np.abs(x).max())).astype(np.float32))
    return out


def padded_frames(n):
    return (4096 * ((n + 4095) // 4096) + 8192) // 1024


# algorithmic work
FLOP_PER_PIXEL = 23264          # whole U-Net, per mask pixel, per stem
FLOP_PER_PIXEL_TC = 22400       # the 10 tensor-core layers
LAYER_FLOP_PER_PIXEL = {"down2": 1600, "down3": 1600, "down4": 1600, "down5": 1600, "down6": 1600,
                        "up1": 1600, "up2": 3200, "up3": 3200, "up4": 3200, "up5": 3200}
