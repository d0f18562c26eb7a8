"""The two sin / cos routines behind the device initial-condition pipeline (csrc/kepler.cu), held on the CPU.

The reference computes states from orbital elements with Python's math.sin / math.cos (core/body.py:184-249,
core/physics.py:43-71) -- the host libm.  csrc/sincos_libm.h restates glibc 2.39's routine operation by operation
(trig mode "libm", the default) and csrc/sincos_cr.h is a correctly rounded pair (mode "cr").  Both headers are
written as sequences of single IEEE operations behind macros, so tests/native/sincos_host.c compiles the SAME
sequences for the host (gcc -ffp-contract=off) and this file pins them:
  * libm restatement == math.sin / math.cos, bit for bit, on every branch of the algorithm;
  * correctly rounded pair == mpmath at 200 bits rounded to nearest;
  * the oracle's element pipeline run with each routine vs the golden states the unmodified reference produced
    (tests/golden/kepler_batch.npz): 100 % bit-identical with the restatement, > 99 % with correct rounding.
The device build of the same headers is compared with these host builds in tests/test_device.py (-m gpu).
"""
import ctypes
import math
import os
import platform
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_f64 = np.ctypeslib.ndpointer(np.float64, flags="C")


def build_host_trig():
    """Compile tests/native/sincos_host.c (test infrastructure) and return the ctypes handle."""
    src = os.path.join(HERE, "native", "sincos_host.c")
    out_dir = os.path.join(HERE, "native", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libsincos_host.so")
    deps = [src] + [os.path.join(HERE, "..", "orbital-physics_b200", "csrc", h)
                    for h in ("sincos_cr.h", "sincos_libm.h", "sincos_libm_tab.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src, "-lm"])
    lib = ctypes.CDLL(out)
    for name in ("sc_host_sincos", "sl_host_sincos"):
        getattr(lib, name).argtypes = [_f64, _f64, _f64, ctypes.c_long]
        getattr(lib, name).restype = ctypes.c_int
    return lib


def host_sincos(lib, which, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s, c = np.empty_like(x), np.empty_like(x)
    ok = getattr(lib, which)(x, s, c, x.size)
    return ok, s, c


def scalar_trig(lib, which):
    """(sin, cos) callables for oracle.ref_numpy.kepler_states."""
    def pair(v):
        _, s, c = host_sincos(lib, which, np.array([v]))
        return s[0], c[0]
    return (lambda v: pair(v)[0]), (lambda v: pair(v)[1])


def branch_arguments(n, seed=11):
    """Arguments for every branch of glibc's sin/cos below |x| = 105414350."""
    rng = np.random.default_rng(seed)
    sign = lambda k: rng.choice([-1.0, 1.0], k)
    return {
        "taylor |x|<0.126": rng.uniform(-0.126, 0.126, n),
        "table |x|<0.855": rng.uniform(-0.855469, 0.855469, n),
        "pi/2-|x| <2.426": rng.uniform(0.855469, 2.426265, n) * sign(n),
        "reduce <2pi": rng.uniform(2.426265, 7.0, n) * sign(n),
        "reduce <1e3": rng.uniform(-1e3, 1e3, n),
        "reduce <1e8": rng.uniform(-1.05e8, 1.05e8, n),
        "tiny": rng.normal(0.0, 1e-7, n),
        "edges": np.array([0.0, -0.0, 5e-324, 1e-300, 2.0 ** -27, 2.0 ** -26, 0.126, -0.126, 0.85546875, 2.4262657165527344,
                           math.pi, -math.pi, math.pi / 2, 2 * math.pi, 1.0, -1.0, 105414000.0]),
    }


@pytest.fixture(scope="module")
def trig():
    return build_host_trig()


def _host_libm_is_the_restated_one():
    return platform.machine() == "x86_64" and platform.libc_ver()[0] == "glibc" and "fma" in open("/proc/cpuinfo").read()


def test_libm_restatement_equals_host_libm(trig):
    if not _host_libm_is_the_restated_one():
        pytest.skip("host libm is not glibc x86-64 with FMA: nothing to compare the restatement with")
    total = 0
    for name, x in branch_arguments(400_000).items():
        ok, s, c = host_sincos(trig, "sl_host_sincos", x)
        assert ok == 1, name
        gs, gc = np.sin(x), np.cos(x)           # numpy calls the same libm as math.sin (checked below)
        assert np.array_equal(s.view(np.int64), gs.view(np.int64)), name
        assert np.array_equal(c.view(np.int64), gc.view(np.int64)), name
        total += x.size
    x = branch_arguments(2000, seed=3)["reduce <2pi"]
    assert all(math.sin(v) == np.sin(v) and math.cos(v) == np.cos(v) for v in x)
    print(f"\nlibm restatement: {total} arguments x (sin, cos) bit-identical to glibc {platform.libc_ver()[1]}")
    ok, _, _ = host_sincos(trig, "sl_host_sincos", np.array([1.1e8]))
    assert ok == 0                                # beyond the restated domain the caller falls back


def test_correctly_rounded_pair_vs_mpmath(trig):
    mpmath = pytest.importorskip("mpmath")
    mpmath.mp.prec = 200
    x = np.concatenate([v[:600] for v in branch_arguments(600, seed=5).values()])
    x = x[np.abs(x) < 2.0 ** 20]
    ok, s, c = host_sincos(trig, "sc_host_sincos", x)
    assert ok == 1
    for v, sv, cv in zip(x, s, c):
        assert sv == float(mpmath.sin(mpmath.mpf(float(v)))), v
        assert cv == float(mpmath.cos(mpmath.mpf(float(v)))), v
    ok, _, _ = host_sincos(trig, "sc_host_sincos", np.array([2.0 ** 20]))
    assert ok == 0


def test_element_pipeline_bits_with_each_routine(trig, golden):
    """oracle element pipeline + each routine vs the reference's golden states (Body.get_state, body.py:184-249)."""
    from oracle import ref_numpy
    g = golden("kepler_batch")
    cols = [g[k] for k in ("M", "e", "a", "b", "n", "inc", "Omega", "omega")]
    sin, cos = scalar_trig(trig, "sl_host_sincos")
    r, v, E = ref_numpy.kepler_states(*cols, sin=sin, cos=cos)
    if _host_libm_is_the_restated_one():
        assert np.array_equal(r, g["r"]) and np.array_equal(v, g["v"]) and np.array_equal(E, g["E"])
    sin, cos = scalar_trig(trig, "sc_host_sincos")
    r, v, E = ref_numpy.kepler_states(*cols, sin=sin, cos=cos)
    same = np.mean(np.all(r == g["r"], axis=1) & np.all(v == g["v"], axis=1))
    err = np.max(np.linalg.norm(r - g["r"], axis=1) / np.linalg.norm(g["r"], axis=1))
    print(f"\ncorrectly rounded trig: {100 * same:.2f} % of the golden states bit-identical, max rel {err:.1e}")
    assert same > 0.98 and err < 1e-14
