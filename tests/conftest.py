import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def hostcheck_lib():
    """g++ build of the product's device math (tests/_host/host_check.cpp) for the CPU tier."""
    import ctypes
    src = os.path.join(ROOT, "tests", "_host", "host_check.cpp")
    out = os.path.join(ROOT, "tests", "_host", "libhostcheck.so")
    csrc = os.path.join(ROOT, "deepcubea_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-std=c++17", "-O1", "-fPIC", "-shared", "-I" + csrc, src, "-o", out])
    return ctypes.CDLL(out)


@pytest.fixture(scope="session")
def oracle_clib():
    """oracle/_build/liboracle_env.so (C restatement of the reference's native env step)."""
    import ctypes
    out = os.path.join(ROOT, "oracle", "_build", "liboracle_env.so")
    if not os.path.exists(out) or os.path.getmtime(os.path.join(ROOT, "oracle", "oracle_env.c")) > os.path.getmtime(out):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(out)
