"""The reference's *unmodified* CLI host (Executable/main.c) must compile against include/*.h and
link against libspleeterrt_b200.so (tier-A drop-in).  Runs only where the reference tree
exists (the build container); nothing is copied from it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SRT_REFERENCE_DIR", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "Executable")), reason="reference tree absent")
def test_unmodified_main_c_links_against_the_library():
    import spleeterrt_b200 as srt
    if not os.path.exists(srt.lib_path()):
        from spleeterrt_b200.build import build
        build()
    bdir = os.path.join(ROOT, "tests", "_build", "linkhost")
    os.makedirs(bdir, exist_ok=True)
    # main.c does `#include "model.c"` (108 MB of fp16 weights); a zero blob of the same type is
    # enough to prove the link, and keeps the compile to seconds.
    with open(os.path.join(bdir, "model.c"), "w") as f:
        f.write("#include <stdint.h>\nstatic const int32_t coeffQuantized[9822725] = {0};\n")
    ex = os.path.join(REF, "Executable")
    exe = os.path.join(bdir, "spleeter_b200")
    cmd = ["gcc", "-O1", "-w", "-I", os.path.join(ROOT, "include"), "-I", bdir, "-I", ex,
           os.path.join(ex, "main.c"), os.path.join(ex, "cpthread.c"),
           os.path.join(ex, "libsamplerate", "samplerate.c"), os.path.join(ex, "libsamplerate", "src_sinc.c"),
           "-L", os.path.dirname(srt.lib_path()), "-lspleeterrt_b200", "-Wl,-rpath," + os.path.dirname(srt.lib_path()),
           "-lm", "-lpthread", "-o", exe]
    subprocess.check_call(cmd)
    # every hot-path symbol must be resolved from our library, none from reference objects
    und = subprocess.check_output(["nm", "-u", exe], text=True)
    for sym in ("initSpleeter", "processSpleeter", "getMaskPtr", "freeSpleeter", "allocateSpleeterStr", "getCoeffSize",
                "InitSTFT", "FreeSTFT", "stft", "istft", "openblas_set_num_threads"):
        assert f" U {sym}" in und, sym
    # and the program starts (prints usage, exit code -2 -> 254) without touching the GPU
    rc = subprocess.run([exe], capture_output=True, text=True)
    assert "Invalid program arguments" in rc.stdout
