"""pytest wiring: registers the hyphenated package as `icicle_snark_b200`, the `gpu` marker, and the
two libraries under test (product: icicle-snark_b200/lib/libicicle_b200.so; oracle: oracle/_ref)."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_package():
    if "icicle_snark_b200" in sys.modules:
        return sys.modules["icicle_snark_b200"]
    pkg_dir = os.path.join(ROOT, "icicle-snark_b200")
    spec = importlib.util.spec_from_file_location(
        "icicle_snark_b200", os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["icicle_snark_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cudart.cuInit(0) != 0:
            return False
        cudart.cuDeviceGetCount(ctypes.byref(n))
        return n.value > 0
    except OSError:
        return False


HAS_GPU = _has_gpu()
if HAS_GPU:
    # torch before any other CUDA library is loaded into the process: the reference's CUDA backend (second oracle,
    # oracle/_ref_cuda) exports its runtime globally, and importing torch after it has been loaded segfaults inside
    # torch's own initialisation.  Whole-suite runs import torch at collection anyway; this makes any -k selection safe.
    import torch  # noqa: F401


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """The product library. Missing .so is a hard failure: there is no fallback."""
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    return pkg.lib()


@pytest.fixture(scope="session")
def gpu(lib):
    lib.set_device("CUDA", 0)
    return lib


@pytest.fixture(scope="session")
def ref():
    from oracle import ref_cpu
    if not ref_cpu.available():
        ref_cpu.build_ref()
    return ref_cpu.ref()


@pytest.fixture()
def rng():
    return np.random.default_rng(20261017)
