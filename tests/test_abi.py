"""The C-ABI library loads and exports every symbol include/orbital_b200.h declares (no compute, no GPU)."""
import ctypes
import os
import re

import pytest

from tests.conftest import REPO


def declared_symbols():
    text = open(os.path.join(REPO, "include", "orbital_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(orb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from core import _native
    syms = declared_symbols()
    assert len(syms) >= 40
    lib = ctypes.CDLL(_native.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in orbital_b200.h but not exported by liborbital_b200.so"
        assert s in _native.SIGNATURES, f"{s} has no ctypes signature in core/_native.py"
    assert set(_native.SIGNATURES) == set(syms)
    assert _native.lib().orb_abi_version() == 1


def test_no_cpu_fallback_without_device():
    """Product path must fail loudly when there is no CUDA device."""
    from core import _native
    if _native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_native.NativeError) as ei:
        _native.DeviceSystem(8)
    assert ei.value.code == 3 and "no CPU fallback" in str(ei.value)
    with pytest.raises(_native.NativeError):
        _native.DeviceEnsemble(4, 16)
    import numpy as np
    from core.physics import Coordinates, Object, ObjectCollection, pairwise_accelerations
    from core.engine import SimulationEngine
    objs = [Object(1e20, 1.0, np.zeros(3), Coordinates(float(i), 0, 0)) for i in range(3)]
    with pytest.raises(_native.NativeError):
        pairwise_accelerations(objs)
    with pytest.raises(_native.NativeError):
        SimulationEngine(ObjectCollection(objs), cache=False)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(REPO, "orbital-physics_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "nbody_oracle" not in src and "liborbital_oracle" not in src, f
