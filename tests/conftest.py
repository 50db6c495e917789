import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.port()
    return O


@pytest.fixture(scope="session")
def small_nets(oracle):
    """Two nets for small-shape tests: real weights when oracle/_ref has them, else synthetic."""
    if oracle.have_real_weights():
        w = oracle.half_to_float(oracle.real_weights_fp16())
        return [(np.ascontiguousarray(w[0]), 1), (np.ascontiguousarray(w[1]), 0)]
    return [(oracle.synthetic_weights(11), 1), (oracle.synthetic_weights(12), 0)]


@pytest.fixture(scope="session")
def host_model():
    """CPU model of the GPU gather-GEMM built from the product's srt_plan.cpp (tests/host_model.cpp)."""
    import ctypes as C
    bdir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "libhostmodel.so")
    srcs = [os.path.join(ROOT, "tests", "host_model.cpp"), os.path.join(ROOT, "spleeterrt_b200", "csrc", "srt_plan.cpp")]
    deps = srcs + [os.path.join(ROOT, "spleeterrt_b200", "csrc", "srt_plan.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so] + srcs)
    lib = C.CDLL(so)
    lib.srt_host_model_round_tf32.restype = C.c_float
    lib.srt_host_model_round_tf32.argtypes = [C.c_float]
    return lib
