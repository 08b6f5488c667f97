"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` tests run on a CPU-only box: the oracle against the reference and the
golden vectors, host-side C code, and that the product libraries load and export every
declared symbol.  `-m gpu` tests are the parity tests proper: they call the CUDA path
through the C ABI and compare with the oracle.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def api(pkg):
    pkg.load()
    return pkg.api


@pytest.fixture(scope="session")
def abi(pkg):
    return pkg.abi


@pytest.fixture(scope="session")
def ol():
    import oracle_lib
    return oracle_lib


@pytest.fixture(scope="session")
def ref_or_skip(ol):
    if not ol.have_ref():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference at build time)")
    return ol


@pytest.fixture(scope="session")
def gpu_api(api):
    """the product API on a box that must have a GPU: a missing device is a FAILURE here,
    never a skip or a fallback"""
    n = api.device_count()
    assert n > 0, "gpu-marked test on a box without a usable CUDA device"
    return api


def random_rays_in_room(rng, n, half_w=30.0):
    """origins inside the room of the default scene, unit directions"""
    o = np.stack([rng.uniform(-half_w, half_w, n), rng.uniform(-18, 18, n), rng.uniform(-25, 45, n)], axis=1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], axis=1)


def psnr_u8(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return 99.0
    return 10.0 * np.log10(255.0 ** 2 / mse)
