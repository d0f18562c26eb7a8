"""pytest configuration: markers + import paths.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI symbol check.
`-m gpu`:       parity tests proper; every one calls the CUDA path through the C ABI.
"""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "orbital-physics_b200")
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def orc():
    from oracle import load_c_oracle
    return load_c_oracle()


def _have_cuda():
    try:
        from core import _native
        return _native.device_count() > 0
    except Exception:
        return False


@pytest.fixture(params=["fake", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    """'fake': engine host logic over the oracle-backed stand-in (CPU); 'cuda': the real C ABI on a B200."""
    from core import _native
    if request.param == "fake":
        from tests.fake_device import FakeDeviceSystem
        monkeypatch.setattr(_native, "DeviceSystem", FakeDeviceSystem)
    else:
        assert _have_cuda(), "gpu-marked test needs a CUDA device and liborbital_b200.so"
    return request.param


def make_objects(g, prefix="in_"):
    """Objects exactly as tests/golden/make_golden.py builds them for the reference."""
    import numpy as np
    from core.physics import Coordinates, Object
    n = len(g[prefix + "x"])
    f64 = np.broadcast_to(np.asarray(g["f64_velocity"], dtype=bool), (n,)) if "f64_velocity" in g else np.zeros(n, bool)
    objs = []
    for i in range(n):
        v = np.array([g[prefix + "vx"][i], g[prefix + "vy"][i], g[prefix + "vz"][i]], dtype=np.float64)
        o = Object(mass=float(g[prefix + "m"][i]), radius=float(g[prefix + "radius"][i]), velocity=v,
                   coordinates=Coordinates(float(g[prefix + "x"][i]), float(g[prefix + "y"][i]), float(g[prefix + "z"][i])),
                   angular_velocity=np.zeros(3), name=f"b{i}")
        if f64[i]:
            o.velocity = v.copy()
        objs.append(o)
    return objs
