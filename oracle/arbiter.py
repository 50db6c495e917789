"""Independent arbiter for the U-Net (TEST INFRASTRUCTURE ONLY — nothing in the product imports oracle/).

A second, structurally different statement of processSpleeter (Executable/spleeter.c:177-301) in float64 on torch-CPU
functional ops, used for two things:

  * to check the C restatement (oracle/srt_oracle.c) and, through it, the reference against textbook convolution
    semantics:  encoder  conv2d(pad(h, (1,2,1,2)), W[O,I,5,5], b, stride 2);  skip = that output;
    next input = act(scale * skip + offset)  (none after down6);
    decoder  conv_transpose2d(h, W[I,O,5,5], b, stride 2)[..., 1:2H+1, 1:2W+1] -> act -> scale * . + offset ->
    cat([skip, up]);  head  conv2d(pad(h, 3), W[2,1,4,4], b, dilation 2) -> sigmoid (LUT or exact);
  * to emulate, in float64, WHERE the GPU path rounds to TF32 (the tensors its epilogues store for a tensor-core
    consumer: raw skips E2..E6, activated A1..A5, decoder outputs U1..U4 — DESIGN.md §3), one layer at a time or
    all together, and so attribute the GPU's mask / stem error to layers without a GPU.
"""
import numpy as np
import torch
import torch.nn.functional as Fn

from . import oracle as O

torch.set_grad_enabled(False)


def round_tf32(t):
    """cvt.rna.tf32.f32 on a float64 tensor holding float32-representable magnitudes: keep 10 mantissa bits,
    round to nearest, ties away from zero."""
    a = t.to(torch.float32).contiguous().numpy().view(np.uint32).astype(np.uint64)
    r = ((a + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return torch.from_numpy(r.view(np.float32).astype(np.float64)).reshape(t.shape)


def _act(kind, x):
    if kind == "leaky":
        return torch.where(x >= 0, x, 0.2 * x)
    if kind == "relu":
        return torch.clamp(x, min=0)
    if kind == "elu_clamp":        # Executable/spleeter.c:51-56
        return torch.where(x >= 0, x, torch.where(x < -15, torch.full_like(x, -1.0), torch.expm1(torch.clamp(x, max=0))))
    return torch.where(x >= 0, x, torch.expm1(torch.clamp(x, max=0)))      # VST flavour: plain ELU


def sigmoid_lut(x):
    """fastSigmoid (Executable/spleeter.c:29-42): 1026-entry table of sigma(-7 + 14 i / 1024) printed to 8 decimals, linear
    interpolation with step 0.01367188, saturating outside [-7, 7]."""
    tbl = np.ctypeslib.as_array(O.port().srt_oracle_sigmoid_table(), shape=(1026,)).astype(np.float64)   # the reference's table (data)
    x = x.numpy()
    step = np.float64(np.float32(0.01367188))
    idx = np.clip(((x + 7.0) / step).astype(np.int64), 0, 1024)
    x1 = -7.0 + step * idx
    y = tbl[idx] + (tbl[idx + 1] - tbl[idx]) / step * (x - x1)
    y = np.where(x > 7.0, 1.0, np.where(x < -7.0, 0.0, y))
    return torch.from_numpy(y)


ROUND_POINTS = ["A1", "E2", "A2", "E3", "A3", "E4", "A4", "E5", "A5", "E6", "U1", "U2", "U3", "U4"]


def unet(coeff, x, stem_mode, flavour=0, round_at=(), return_logits=False):
    """x: float32[2][T][F] magnitudes -> mask float64[2][T][F].
    round_at: subset of ROUND_POINTS (or "all"): tensors rounded to TF32 where the GPU path stores them rounded."""
    pts = set(ROUND_POINTS if round_at == "all" else round_at)
    v = {k: torch.from_numpy(np.asarray(a, np.float64)) for k, a in O.coeff_views(np.asarray(coeff, np.float32)).items()}
    elu = "elu" if flavour else "elu_clamp"
    a_enc, a_dec = (elu, elu) if stem_mode else ("leaky", "relu")
    h = torch.from_numpy(np.asarray(x, np.float64))[None]

    def rnd(name, t):
        return round_tf32(t) if name in pts else t
    skips = []
    for i in range(1, 7):
        raw = Fn.conv2d(Fn.pad(h, (1, 2, 1, 2)), v[f"down{i}.w"], v[f"down{i}.b"], stride=2)
        if i < 6:
            bn = v[f"down{i}.bn"]
            h = rnd(f"A{i}", _act(a_enc, bn[1][None, :, None, None] * raw + bn[0][None, :, None, None]))
        skips.append(rnd(f"E{i}", raw))          # E1 stays fp32 (it feeds the hi + lo up6)
    h = skips[5]
    for d in range(1, 7):
        H, W = h.shape[2], h.shape[3]
        up = Fn.conv_transpose2d(h, v[f"up{d}.w"], v[f"up{d}.b"], stride=2)[:, :, 1:2 * H + 1, 1:2 * W + 1]
        bn = v[f"up{d}.bn"]
        up = bn[1][None, :, None, None] * _act(a_dec, up) + bn[0][None, :, None, None]
        if d < 6:
            h = torch.cat([skips[5 - d], rnd(f"U{d}", up)], 1)
        else:
            h = up
    logits = Fn.conv2d(Fn.pad(h, (3, 3, 3, 3)), v["up7.w"], v["up7.b"], dilation=2)[0]
    if return_logits:
        return logits
    return torch.sigmoid(logits) if flavour else sigmoid_lut(logits)


def stem_error_from_mask_error(mask_ref, mask_emu, L, R, T, F):
    """RMS difference of the separated stem when only the mask differs: istft is linear, so push the mask difference
    through the oracle's STFT -> multiply -> iSTFT on the first tile."""
    n = L.size
    padded = O.FFT * ((n + O.FFT - 1) // O.FFT) + 2 * O.FFT
    pl, pr = np.zeros(padded, np.float32), np.zeros(padded, np.float32)
    pl[O.FFT:O.FFT + n], pr[O.FFT:O.FFT + n] = L, R
    planes = O.stft(pl, pr)
    frames = min(planes[0].shape[0], T)
    dm = (np.asarray(mask_emu, np.float64) - np.asarray(mask_ref, np.float64)).astype(np.float32)
    out = [np.zeros_like(p) for p in planes]
    for c, (re, im) in enumerate(((0, 1), (2, 3))):
        out[re][:frames, :F] = planes[re][:frames, :F] * dm[c][:frames]
        out[im][:frames, :F] = planes[im][:frames, :F] * dm[c][:frames]
    oL, oR = O.istft(*out)
    seg = slice(O.FFT, O.FFT + min(n, frames * O.HOP))
    return float(np.sqrt(np.mean(np.concatenate([oL[seg], oR[seg]]).astype(np.float64) ** 2)))
