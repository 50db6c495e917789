"""Build libspleeterrt_b200.so (sm_100a only) in-tree with nvcc.

    python -m spleeterrt_b200.build [--force] [--verbose]

The shared library is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libspleeterrt_b200.so")
SOURCES = ["srt_plan.cpp", "srt_weights.cpp", "srt_ctx.cu", "srt_conv_tc.cu", "srt_conv_rp.cu", "srt_up6_tc.cu", "srt_unet_simt.cu", "srt_stft.cu", "srt_stream.cu", "srt_tier_a.cu", "srt_resample.cu", "srt_probe.cu"]
HEADERS = ["srt_plan.h", "srt_kernels.cuh", "srt_ptx.cuh", "srt_epilogue.cuh", "srt_fft.cuh", "srt_internal.h",
           "../../include/srt_b200.h", "../../include/spleeter.h", "../../include/stftFix.h", "../../include/Spleeter4Stems.h"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-fno-strict-aliasing,-ffp-contract=off"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    objdir = os.path.join(PKG, "_build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC, *FLAGS, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- {src}\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        subprocess.check_call(cmd)
    build_dispatch(force, verbose)
    return LIB


DISPATCH_LIB = os.path.join(PKG, "libspleeterrt_dispatch.so")


def build_dispatch(force=False, verbose=False):
    """libspleeterrt_dispatch.so: the NCCL stream dispatcher (include/srt_dispatch.h), a separate library so that the product
    library itself depends on nothing but libc/libstdc++.  Links the system libnccl (SONAME libnccl.so.2: inside a torch process the
    loader resolves it to the copy torch has already loaded) and libspleeterrt_b200.so."""
    src = os.path.join(CSRC, "srt_dispatch.cu")
    obj = os.path.join(PKG, "_build", "srt_dispatch.o")
    hdrs = [os.path.join(PKG, "..", "include", h) for h in ("srt_dispatch.h", "srt_b200.h")]
    if force or _stale(obj, [src] + hdrs):
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    if force or _stale(DISPATCH_LIB, [obj, LIB]):
        subprocess.check_call([NVCC, "-shared", "-o", DISPATCH_LIB, obj, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
                               "-L", PKG, "-lspleeterrt_b200", "-lnccl", "-Xlinker", "-rpath=$ORIGIN"])
    return DISPATCH_LIB


def build_examples():
    """The C hosts under examples/ (plain gcc against include/ and the in-tree library) -> examples/_build/."""
    root = os.path.dirname(PKG)
    out = os.path.join(root, "examples", "_build")
    os.makedirs(out, exist_ok=True)
    built = []
    for name in ("main_b200", "spleeter_cli_b200", "vst_host"):
        src = os.path.join(root, "examples", name + ".c")
        exe = os.path.join(out, name)
        if _stale(exe, [src, LIB, os.path.join(root, "include", "srt_b200.h")]):
            subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(root, "include"), src, "-L", PKG, "-lspleeterrt_b200",
                                   "-Wl,-rpath," + PKG, "-Wl,-rpath,$ORIGIN/../../spleeterrt_b200", "-lm", "-o", exe])
        built.append(exe)
    return built


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_examples())
