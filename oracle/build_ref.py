#!/usr/bin/env python3
"""Build the *reference itself* as test infrastructure (oracle/_ref/).

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed by the
product path (spleeterrt_b200/, include/, the C-ABI library).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

What this does (all outputs go to oracle/_ref/, which is git-ignored but travels to the
GPU box with gpurun):

  1. decodes /root/reference/Executable/model.7z (one raw LZMA2 stream at byte 32,
     42 853 703 bytes, dict 128 MiB -> model.c, CRC32 0x80032de0) with the stdlib `lzma`
     and stores the 19 645 450 IEEE-half weights (2 nets, spleeterQuantized layout,
     Executable/spleeter.h:32-62) as  spleeterrt_b200/weights/model_fp16.bin (git-ignored data)
  2. compiles the reference's own C sources *where they lie* (never copied) into
        oracle/_ref/libref_exec.so   (Executable flavour: LUT sigmoid, ELU clamp)
        oracle/_ref/libref_vst.so    (VST flavour: exact sigmoid, streaming Spleeter4Stems)
     with -DCPU_GEMM=1 (in-repo naive sgemm, Executable/gemm.c:64-80) and the
     zeroing-malloc shim (gemm_cpu multiplies uninitialised C by BETA=0,
     Executable/gemm.c:71-72 -> NaN on a dirty heap).

It is a no-op when /root/reference is absent (GPU box): the prebuilt files are used.
"""
import hashlib
import lzma
import os
import subprocess
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SRT_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
WEIGHTS_SHA256 = "b9837a8b6379c71b6442fbf4b0fb3c7eb668c0c76cf2702eda6ce7b7aa93e459"


def decode_weights(force=False):
    import numpy as np
    # one copy only (it travels with every gpurun push): the decoded blob lives in the product's
    # git-ignored data directory and the oracle reads the same file
    wdir = os.path.join(os.path.dirname(HERE), "spleeterrt_b200", "weights")
    os.makedirs(wdir, exist_ok=True)
    dst = os.path.join(wdir, "model_fp16.bin")
    if os.path.exists(dst) and not force:
        return dst
    raw = open(os.path.join(REF, "Executable", "model.7z"), "rb").read()
    packed = raw[32:32 + 42853703]
    dec = lzma.LZMADecompressor(format=lzma.FORMAT_RAW,
                                filters=[{"id": lzma.FILTER_LZMA2, "dict_size": 1 << 27}])
    src = dec.decompress(packed)
    assert zlib.crc32(src) == 0x80032DE0, "model.c CRC mismatch"
    body = src[src.index(b"{") + 1: src.rindex(b"}")]
    vals = np.array(body.replace(b"\n", b"").split(b","), dtype=np.int64).astype(np.int32)
    assert vals.size == 9822725
    halves = vals.view(np.uint16)
    assert halves.size == 19645450
    assert hashlib.sha256(halves.tobytes()).hexdigest() == WEIGHTS_SHA256
    halves.tofile(dst)
    return dst


def cc(out, srcs, extra=()):
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-DCPU_GEMM=1", "-w",
           "-include", os.path.join(HERE, "zero_malloc.h"), *extra, "-o", out, *srcs,
           "-lm", "-lpthread"]
    subprocess.check_call(cmd)


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    if not os.path.isdir(REF):
        return False
    ex = os.path.join(REF, "Executable")
    vst = os.path.join(REF, "VST", "Source")
    so = os.path.join(OUT, "libref_exec.so")
    if force or not os.path.exists(so):
        cc(so, [os.path.join(ex, f) for f in
                ("spleeter.c", "gemm.c", "im2col_dilated.c", "stftFix.c", "codelet.c", "cpthread.c")],
           extra=["-I" + ex])
    so = os.path.join(OUT, "libref_vst.so")
    if force or not os.path.exists(so):
        # VST/Source/gemm.c is MKL-only (VST/Source/gemm.c:60-63); the Executable gemm.c with
        # CPU_GEMM=1 implements the same row-major sgemm contract.
        cc(so, [os.path.join(vst, f) for f in
                ("Spleeter4Stems.c", "spleeter.c", "im2col_dilated.c", "codelet.c", "cpthread.c")]
           + [os.path.join(ex, "gemm.c")], extra=["-I" + vst])
    decode_weights(force)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built" if ok else "reference tree absent: using prebuilt oracle/_ref")
