import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


def golden(name: str) -> bytes:
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def cref():
    """C restatement of the reference's CPU algorithms (oracle/cpu_ref.c), built on demand."""
    so = os.path.join(ROOT, "oracle", "liboracle_cpu_ref.so")
    src = os.path.join(ROOT, "oracle", "cpu_ref.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    import cpu_ref

    return cpu_ref


@pytest.fixture(scope="session")
def hostemul():
    """Host build of the product's per-point limb code (tests/host_emul)."""
    import ctypes

    if os.environ.get("PTAU_HOSTEMUL_SO"):  # a host build with other feature macros (A/B of a kernel variant)
        return ctypes.CDLL(os.environ["PTAU_HOSTEMUL_SO"])
    d = os.path.join(ROOT, "tests", "host_emul")
    so = os.path.join(d, "libptau_hostemul.so")
    csrc = os.path.join(ROOT, "kzg_setup_powersoftau_b200", "csrc")
    deps = [os.path.join(d, "hostemul.cpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc)
                                                 if f.endswith((".cuh", ".inc"))]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so,
                               os.path.join(d, "hostemul.cpp")])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def ctx():
    import kzg_setup_powersoftau_b200 as kz

    c = kz.Context(1)
    yield c
    c.close()
