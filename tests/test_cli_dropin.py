"""BASELINE.json configs[0] through the reference's own host program.

oracle/build_ref.py compiles the reference's UNMODIFIED Executable/main.c once and links that object
twice: against the reference's own C sources (oracle/_ref/spleeter_cli_ref) and against
libspleeterrt_b200.so (oracle/_ref/spleeter_cli_b200).  Both binaries decode the same WAV, run
stft -> processMT -> istft (main.c:762-841; 3-output cascade main.c:845-970) and write float32 WAVs.

  * not gpu: the reference CLI's vocal stem pins the oracle port's tile driver (oracle/srt_oracle.c).
  * gpu    : the B200-linked CLI's stems match the reference CLI's within the 1e-4 RMS the north star states,
             for the 2-output and the 3-output mode (tier-A drop-in: nothing in main.c changed).
Binaries are prebuilt in the build container (the GPU box has no /root/reference)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "spleeter_cli_ref")
B200_CLI = os.path.join(ROOT, "oracle", "_ref", "spleeter_cli_b200")
TOL_RMS = 1e-4   # BASELINE.json north_star: "within 1e-4 RMS per stem (float32)"


def write_wav_f32(path, x, rate=44100):
    """x: float32 [n] (mono) or [n][ch]; IEEE-float WAV."""
    x = np.ascontiguousarray(x, np.float32)
    ch = 1 if x.ndim == 1 else x.shape[1]
    data = x.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 3, ch, rate, rate * ch * 4, ch * 4, 32))
        f.write(b"data" + struct.pack("<I", len(data)) + data)


def read_wav_f32(path):
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    p, fmt = 12, None
    while p + 8 <= len(b):
        cid, sz = b[p:p + 4], struct.unpack("<I", b[p + 4:p + 8])[0]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", b[p + 8:p + 24])
        if cid == b"data":
            assert fmt and fmt[0] == 3 and fmt[5] == 32, fmt
            return np.frombuffer(b[p + 8:p + 8 + sz], np.float32).reshape(-1, fmt[1])
        p += 8 + sz + (sz & 1)
    raise AssertionError("no data chunk")


def run_cli(exe, workdir, wav, T, F, stems, threads=1):
    os.makedirs(workdir, exist_ok=True)
    env = dict(os.environ)
    env.setdefault("OMP_NUM_THREADS", str(min(os.cpu_count() or 1, 16)))
    r = subprocess.run([exe, str(threads), str(T), str(F), str(stems), wav], cwd=workdir, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-500:])
    base = os.path.basename(wav)
    names = ["Vocal", "Accompaniment"] if stems <= 2 else None
    if names is None:   # 3-output mode: whatever "<input>_*.wav" files main.c:845-970 wrote
        names = sorted(fn[len(base) + 1:-4] for fn in os.listdir(workdir) if fn.startswith(base + "_") and fn.endswith(".wav"))
    return {nm: read_wav_f32(os.path.join(workdir, f"{base}_{nm}.wav")) for nm in names}


needs_cli = pytest.mark.skipif(not (os.path.exists(REF_CLI) and os.path.exists(B200_CLI)),
                               reason="oracle/_ref CLI binaries not built (python oracle/build_ref.py in the build container)")


@needs_cli
def test_reference_cli_pins_oracle_tile_driver(oracle, tmp_path):
    """3 s mono WAV, `1 64 512 2`: Vocal = net[1] in mode 0 (main.c:760,782); mono is duplicated to both
    channels (main.c:768-769)."""
    L, _ = oracle.synth_pcm(5, n=3 * 44100)
    wav = str(tmp_path / "in3s.wav")
    write_wav_f32(wav, L)
    out = run_cli(REF_CLI, str(tmp_path / "ref"), wav, 64, 512, 2)
    w = oracle.half_to_float(oracle.real_weights_fp16())
    stems = oracle.separate([(np.ascontiguousarray(w[1]), 0)], L, L, 64, 512, unaffected=0.1)
    voc = out["Vocal"]
    assert voc.shape == (L.size, 2)
    for c in range(2):
        err = float(np.sqrt(np.mean((voc[:, c] - stems[0, c]) ** 2)))
        assert err < 2e-6, err
    acc = out["Accompaniment"]          # input - vocal in the time domain (main.c:794-798)
    assert float(np.abs(acc[:, 0] - (L - voc[:, 0])).max()) < 1e-6


@pytest.mark.gpu
@needs_cli
@pytest.mark.parametrize("T,F,stems,seconds", [(512, 1024, 2, 10.0), (128, 512, 3, 6.0), (64, 1536, 2, 4.0)])
def test_unmodified_cli_on_b200_matches_reference_cli(oracle, tmp_path, T, F, stems, seconds):
    """configs[0] (10 s mono, `1 512 1024 2`) plus the 3-output cascade and a non-power-of-two F, stereo."""
    n = int(seconds * 44100)
    L, R = oracle.synth_pcm(7, n=n)
    wav = str(tmp_path / "in.wav")
    write_wav_f32(wav, L if stems == 2 and T == 512 else np.stack([L, R], axis=1))
    ref = run_cli(REF_CLI, str(tmp_path / "ref"), wav, T, F, stems)
    got = run_cli(B200_CLI, str(tmp_path / "b200"), wav, T, F, stems)
    assert sorted(ref) == sorted(got) and len(ref) == (2 if stems <= 2 else 3)
    for nm in ref:
        assert ref[nm].shape == got[nm].shape == (n, 2)
        for c in range(2):
            err = float(np.sqrt(np.mean((ref[nm][:, c].astype(np.float64) - got[nm][:, c]) ** 2)))
            assert err < TOL_RMS, (nm, c, err)
        assert float(np.sqrt(np.mean(ref[nm].astype(np.float64) ** 2))) > 1e-3, f"{nm}: silent reference output"


@pytest.mark.gpu
@needs_cli
def test_unmodified_cli_with_four_host_threads(oracle, tmp_path):
    """`spawnNthreads = 4` (main.c:544-577): processMT creates FOUR _spleeter instances from ONE coefficient pointer and runs
    processSpleeter concurrently from four host threads (task_Separation, main.c:296-365), plus the tail tile on thread 0.
    A 14 s stereo file at T = 64 is 9 full tiles + a tail.  Per-instance CUDA streams, no shared mutable state: the stems
    equal the reference CLI's (same thread count) and the single-thread run of the same binary bit for bit."""
    n = 14 * 44100
    L, R = oracle.synth_pcm(13, n=n)
    wav = str(tmp_path / "in.wav")
    write_wav_f32(wav, np.stack([L, R], axis=1))
    ref = run_cli(REF_CLI, str(tmp_path / "ref"), wav, 64, 512, 2, threads=4)
    got = run_cli(B200_CLI, str(tmp_path / "b200"), wav, 64, 512, 2, threads=4)
    one = run_cli(B200_CLI, str(tmp_path / "b200_1"), wav, 64, 512, 2, threads=1)
    for nm in ("Vocal", "Accompaniment"):
        assert got[nm].shape == ref[nm].shape == (n, 2)
        for c in range(2):
            err = float(np.sqrt(np.mean((ref[nm][:, c].astype(np.float64) - got[nm][:, c]) ** 2)))
            assert err < TOL_RMS, (nm, c, err)
        assert np.array_equal(got[nm], one[nm]), f"{nm}: the 4-thread run differs from the 1-thread run"


EXAMPLE_CLI = os.path.join(ROOT, "examples", "_build", "spleeter_cli_b200")
MODEL_FP16 = os.path.join(ROOT, "spleeterrt_b200", "weights", "model_fp16.bin")


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF_CLI) and os.path.exists(EXAMPLE_CLI) and os.path.exists(MODEL_FP16)),
                    reason="reference CLI / examples/_build / model blob not built (python __graft_entry__.py in the build container)")
@pytest.mark.parametrize("T,F,stems,seconds,mono", [(128, 512, 3, 6.0, False), (64, 1024, 2, 3.0, True), (256, 1024, 2, 8.0, False)])
def test_device_level_cli_host_matches_reference_cli(oracle, tmp_path, T, F, stems, seconds, mono):
    """examples/spleeter_cli_b200.c: same command line and output files as main.c, but decode -> ONE device call
    (srt_create_cli + srt_separate_batch_interleaved: split, cascade, subtraction and join on the GPU) -> WAV writer.
    Its WAVs match the reference CLI's within the north star's 1e-4 RMS (stereo and mono inputs)."""
    n = int(seconds * 44100)
    L, R = oracle.synth_pcm(9, n=n)
    wav = str(tmp_path / "in.wav")
    write_wav_f32(wav, L if mono else np.stack([L, R], axis=1))
    ref = run_cli(REF_CLI, str(tmp_path / "ref"), wav, T, F, stems)
    os.makedirs(tmp_path / "dev", exist_ok=True)
    r = subprocess.run([EXAMPLE_CLI, "1", str(T), str(F), str(stems), wav, MODEL_FP16], cwd=str(tmp_path / "dev"),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-500:])
    names = ["Vocal", "Accompaniment"] if stems <= 2 else ["Accompaniment", "Drum", "Vocal"]
    assert sorted(ref) == sorted(names)
    for nm in names:
        got = read_wav_f32(str(tmp_path / "dev" / f"in.wav_{nm}.wav"))
        assert got.shape == ref[nm].shape == (n, 2)
        for c in range(2):
            err = float(np.sqrt(np.mean((ref[nm][:, c].astype(np.float64) - got[:, c]) ** 2)))
            assert err < TOL_RMS, (nm, c, err)


RESAMPLER_TABLE = os.path.join(ROOT, "spleeterrt_b200", "weights", "resampler_mq.f32")


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF_CLI) and os.path.exists(EXAMPLE_CLI) and os.path.exists(MODEL_FP16) and os.path.exists(RESAMPLER_TABLE)),
                    reason="reference CLI / examples/_build / model blob / resampler table not built")
def test_device_level_cli_host_resamples_like_the_reference(oracle, tmp_path):
    """A 48 kHz stereo WAV: main.c:264-270 converts it to 44.1 kHz with libsamplerate before the path; the device-level
    host does the same conversion with srt_resample_host (bit-identical) and then matches the reference CLI's stems."""
    n48 = 4 * 48000
    L, R = oracle.synth_pcm(11, n=n48)
    wav = str(tmp_path / "in48.wav")
    write_wav_f32(wav, np.stack([L, R], axis=1), rate=48000)
    ref = run_cli(REF_CLI, str(tmp_path / "ref"), wav, 64, 512, 2)
    os.makedirs(tmp_path / "dev", exist_ok=True)
    env = dict(os.environ, SRT_RESAMPLER_TABLE=RESAMPLER_TABLE)
    r = subprocess.run([EXAMPLE_CLI, "1", "64", "512", "2", wav, MODEL_FP16], cwd=str(tmp_path / "dev"), env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-500:])
    n = int(np.ceil(n48 * 44100.0 / 48000.0))
    for nm in ("Vocal", "Accompaniment"):
        got = read_wav_f32(str(tmp_path / "dev" / f"in48.wav_{nm}.wav"))
        assert got.shape == ref[nm].shape == (n, 2)
        for c in range(2):
            err = float(np.sqrt(np.mean((ref[nm][:, c].astype(np.float64) - got[:, c]) ** 2)))
            assert err < TOL_RMS, (nm, c, err)
    # vocal + accompaniment = the converted input, so the conversions agree to rounding
    a = read_wav_f32(str(tmp_path / "dev" / "in48.wav_Vocal.wav")) + read_wav_f32(str(tmp_path / "dev" / "in48.wav_Accompaniment.wav"))
    b = ref["Vocal"] + ref["Accompaniment"]
    assert float(np.abs(a - b).max()) < 1e-6
