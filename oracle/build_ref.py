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
  3. compiles the reference's unmodified CLI host Executable/main.c once and links it against
     (a) libref_exec.so and (b) libspleeterrt_b200.so  ->  oracle/_ref/spleeter_cli_{ref,b200}

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
    wbin = decode_weights(force)
    build_cli(wbin, force)
    build_resampler(force)
    build_blas(force)
    build_vst_host(force)
    return True


OPENBLAS = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs/libopenblasp-r0-59ffcd50.3.15.so"


def build_blas(force=False):
    """oracle/_ref/libref_exec_blas.so: the reference's Executable flavour on its REAL backend - gemm.c WITHOUT -DCPU_GEMM calls
    cblas_sgemm (Executable/gemm.c:82-89; MKL on Windows, OpenBLAS on Linux where main.c:688-689 calls
    openblas_set_num_threads).  MKL is not in this image; an OpenBLAS 0.3.15 that exports the plain CBLAS symbols is (a wheel's
    private copy), so `#include <mkl.h>` is satisfied by the six-line oracle/blas_shim/mkl.h and the library is linked against
    that file by path.  No zero-malloc shim here: cblas_sgemm with beta = 0 never reads C.  This is the CPU baseline the
    README's "MKL is ~40x faster than the naive loops" refers to (README.MD:155), as far as it can be had here."""
    if not os.path.exists(OPENBLAS):
        return False
    ex = os.path.join(REF, "Executable")
    so = os.path.join(OUT, "libref_exec_blas.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-w", "-I", os.path.join(HERE, "blas_shim"), "-I", ex,
                               "-o", so] + [os.path.join(ex, f) for f in ("spleeter.c", "gemm.c", "im2col_dilated.c", "stftFix.c", "codelet.c", "cpthread.c")]
                              + [OPENBLAS, "-Wl,--disable-new-dtags,-rpath," + os.path.dirname(OPENBLAS), "-lm", "-lpthread"])   # DT_RPATH: also finds OpenBLAS's own libgfortran
    return True


def build_vst_host(force=False):
    """oracle/_ref/vst_host_ref: examples/vst_host.c (our stand-in for the JUCE plugin shell) compiled against the REFERENCE's
    VST/Source/Spleeter4Stems.h and linked to the reference's own streamer (libref_vst.so), to be run beside
    examples/_build/vst_host (same source, include/Spleeter4Stems.h, libspleeterrt_b200.so)."""
    vst = os.path.join(REF, "VST", "Source")
    exe = os.path.join(OUT, "vst_host_ref")
    src = os.path.join(os.path.dirname(HERE), "examples", "vst_host.c")
    if force or not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-w", "-I", vst, src, "-L", OUT, "-lref_vst", "-Wl,-rpath,$ORIGIN", "-lm", "-lpthread", "-o", exe])


def build_resampler(force=False):
    """oracle/_ref/libref_resample.so: the reference's resampling front end (main.c:132-224 + libsamplerate) behind the
    two entry points of oracle/ref_resample_shim.c, and the sinc coefficient table it decompresses, dumped next to the
    model blob (spleeterrt_b200/weights/resampler_mq.f32: reference DATA, git-ignored like the weights)."""
    ex = os.path.join(REF, "Executable")
    so = os.path.join(OUT, "libref_resample.so")
    shim = os.path.join(HERE, "ref_resample_shim.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(shim):
        sdir = os.path.join(OUT, "model_stub")
        os.makedirs(sdir, exist_ok=True)
        with open(os.path.join(sdir, "model.c"), "w") as f:      # main.c only takes the blob's address (main.c:759)
            f.write("#include <stdint.h>\nstatic const int32_t coeffQuantized[1] = {0};\n")
        with open(os.path.join(sdir, "stub.c"), "w") as f:
            f.write("void openblas_set_num_threads(int n) { (void)n; }\n")
        subprocess.check_call(["gcc", "-O2", "-w", "-fPIC", "-shared", "-I", sdir, "-I", ex, shim, os.path.join(sdir, "stub.c"),
                               os.path.join(ex, "libsamplerate", "samplerate.c"), os.path.join(ex, "libsamplerate", "src_sinc.c"),
                               "-L", OUT, "-lref_exec", "-Wl,-rpath,$ORIGIN", "-lm", "-lpthread", "-o", so])
    table = os.path.join(os.path.dirname(HERE), "spleeterrt_b200", "weights", "resampler_mq.f32")
    if force or not os.path.exists(table):
        import ctypes
        import numpy as np
        lib = ctypes.CDLL(so)
        t = np.zeros(22438, np.float32)
        lib.ref_resampler_table(t.ctypes.data_as(ctypes.c_void_p))
        os.makedirs(os.path.dirname(table), exist_ok=True)
        t.tofile(table)


def build_cli(wbin, force=False):
    """The reference's UNMODIFIED CLI host (Executable/main.c), compiled once and linked twice:
         oracle/_ref/spleeter_cli_ref   against the reference's own objects (libref_exec.so)
         oracle/_ref/spleeter_cli_b200  against libspleeterrt_b200.so (the drop-in claim, SURVEY 8b tier A)
    main.c does `#include "model.c"`, 108 MB of C source that the reference ships as model.7z.  Its
    content is data (`static const int32_t coeffQuantized[9822725]`), so instead of compiling the
    decoded text for 69 s the same array is placed in .rodata with `.incbin` from the decoded blob.
    tests/test_cli_dropin.py runs both binaries on the same WAV (BASELINE.json configs[0])."""
    ex = os.path.join(REF, "Executable")
    mdir = os.path.join(OUT, "model_incbin")
    os.makedirs(mdir, exist_ok=True)
    main_o = os.path.join(OUT, "main_cli.o")
    ref_cli = os.path.join(OUT, "spleeter_cli_ref")
    b200_cli = os.path.join(OUT, "spleeter_cli_b200")
    pkg = os.path.join(os.path.dirname(HERE), "spleeterrt_b200")
    lib = os.path.join(pkg, "libspleeterrt_b200.so")
    if force or not os.path.exists(main_o):
        with open(os.path.join(mdir, "model.c"), "w") as f:
            f.write('#include <stdint.h>\n'
                    '__asm__(".section .rodata\\n.balign 64\\n.globl coeffQuantized\\ncoeffQuantized:\\n'
                    '.incbin \\"%s\\"\\n.previous\\n");\n'
                    'extern const int32_t coeffQuantized[9822725];\n' % wbin)
        subprocess.check_call(["gcc", "-O2", "-w", "-c", "-I", mdir, "-I", ex, os.path.join(ex, "main.c"), "-o", main_o])
    host = [os.path.join(ex, "libsamplerate", "samplerate.c"), os.path.join(ex, "libsamplerate", "src_sinc.c")]
    if force or not os.path.exists(ref_cli):
        stub = os.path.join(mdir, "stub.c")
        with open(stub, "w") as f:
            f.write("void openblas_set_num_threads(int n) { (void)n; }\n")   # main.c:689 expects OpenBLAS on Linux
        subprocess.check_call(["gcc", "-O2", "-w", "-fopenmp", "-I", ex, main_o, *host, stub, "-L", OUT, "-lref_exec",
                               "-Wl,-rpath,$ORIGIN", "-lm", "-lpthread", "-o", ref_cli])
    if os.path.exists(lib) and (force or not os.path.exists(b200_cli) or os.path.getmtime(b200_cli) < os.path.getmtime(main_o)):
        subprocess.check_call(["gcc", "-O2", "-w", "-I", ex, main_o, *host, os.path.join(ex, "cpthread.c"),
                               "-L", pkg, "-lspleeterrt_b200", "-Wl,-rpath,$ORIGIN/../../spleeterrt_b200",
                               "-lm", "-lpthread", "-o", b200_cli])


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built" if ok else "reference tree absent: using prebuilt oracle/_ref")
