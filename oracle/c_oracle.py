"""ctypes binding of oracle/nbody_oracle.c (test infrastructure)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborbital_oracle.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nbody_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


class COracle:
    def __init__(self, path: str = _SO):
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.orc_max_threads.restype = C.c_int
        L.orc_pairwise_half.argtypes = [C.c_int64, _dp, _dp, _dp, _dp, C.c_double, C.c_double, _dp, _dp, _dp,
                                        C.POINTER(C.c_double)]
        L.orc_pairwise_rows.argtypes = [C.c_int64, _dp, _dp, _dp, _dp, C.c_double, C.c_double, _dp, _dp, _dp, C.c_int]
        L.orc_pairwise_sample.argtypes = [C.c_int64, _dp, _dp, _dp, _dp, C.c_double, C.c_double, _i64p, C.c_int64,
                                          _dp, _dp, _dp, C.c_int]
        L.orc_pairwise_sample_ld.argtypes = [C.c_int64, _dp, _dp, _dp, _dp, C.c_double, C.c_double, _i64p, C.c_int64,
                                             _dp, _dp, _dp, _dp, C.c_int]
        L.orc_potential.argtypes = [C.c_int64, _dp, _dp, _dp, _dp, C.c_double, C.c_double]
        L.orc_potential.restype = C.c_double
        L.orc_collisions.argtypes = [C.c_int64] + [_dp] * 8 + [_u8p, C.c_double]
        L.orc_collisions.restype = C.c_int64
        L.orc_step.argtypes = ([C.c_int64] + [_dp] * 8 + [_u8p] + [_dp] * 3 +
                               [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int64, C.c_int,
                                C.POINTER(C.c_double)])
        L.orc_step.restype = C.c_int64
        L.orc_ensemble_step.argtypes = ([C.c_int64, C.c_int64] + [_dp] * 7 +
                                        [C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int, C.c_int])
        L.orc_kinetic.argtypes = [C.c_int64, _dp, _dp, _dp, _dp, _u8p]
        L.orc_kinetic.restype = C.c_double
        L.orc_angmom.argtypes = [C.c_int64] + [_dp] * 8

    @property
    def max_threads(self) -> int:
        return int(self.lib.orc_max_threads())

    @staticmethod
    def _c(a):
        return np.ascontiguousarray(a, dtype=np.float64)

    def pairwise(self, x, y, z, m, eps, G, nthreads: int = 1):
        """-> (acc[n,3], U).  nthreads=1: literal half-matrix loop incl. U."""
        x, y, z, m = map(self._c, (x, y, z, m))
        n = x.shape[0]
        ax, ay, az = np.empty(n), np.empty(n), np.empty(n)
        if nthreads == 1:
            U = C.c_double(0.0)
            self.lib.orc_pairwise_half(n, x, y, z, m, eps, G, ax, ay, az, C.byref(U))
            return np.stack([ax, ay, az], 1), U.value
        self.lib.orc_pairwise_rows(n, x, y, z, m, eps, G, ax, ay, az, nthreads)
        return np.stack([ax, ay, az], 1), None

    def pairwise_sample(self, x, y, z, m, eps, G, rows, long_double: bool = False, nthreads: int = 0):
        x, y, z, m = map(self._c, (x, y, z, m))
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        k = rows.shape[0]
        ax, ay, az = np.empty(k), np.empty(k), np.empty(k)
        if long_double:
            sa = np.empty(k)
            self.lib.orc_pairwise_sample_ld(x.shape[0], x, y, z, m, eps, G, rows, k, ax, ay, az, sa, nthreads)
            return np.stack([ax, ay, az], 1), sa
        self.lib.orc_pairwise_sample(x.shape[0], x, y, z, m, eps, G, rows, k, ax, ay, az, nthreads)
        return np.stack([ax, ay, az], 1)

    def potential(self, x, y, z, m, eps, G) -> float:
        x, y, z, m = map(self._c, (x, y, z, m))
        return float(self.lib.orc_potential(x.shape[0], x, y, z, m, eps, G))


class State:
    """Mutable SoA state the C oracle steps in place (engine.py:19-46 semantics)."""

    def __init__(self, orc: COracle, x, y, z, vx, vy, vz, m, radius, vf32, dt, eps, G=6.67430e-11,
                 restitution=1.0):
        self.orc = orc
        f = lambda a: np.array(a, dtype=np.float64, copy=True)
        self.x, self.y, self.z = f(x), f(y), f(z)
        self.vx, self.vy, self.vz = f(vx), f(vy), f(vz)
        self.m, self.radius = f(m), f(radius)
        self.n = self.x.shape[0]
        self.vf32 = np.ascontiguousarray(np.broadcast_to(np.asarray(vf32, dtype=np.uint8), (self.n,))).copy()
        # Object.__init__ casts constructor velocities to float32 (physics.py:184)
        for v in (self.vx, self.vy, self.vz):
            sel = self.vf32.astype(bool)
            v[sel] = v[sel].astype(np.float32).astype(np.float64)
        self.dt, self.eps, self.G, self.restitution = float(dt), float(eps), float(G), float(restitution)
        acc, self.U = orc.pairwise(self.x, self.y, self.z, self.m, self.eps, self.G, 1)   # engine.py:41
        self.ax, self.ay, self.az = (np.ascontiguousarray(acc[:, k]) for k in range(3))
        self.hits = 0

    def step(self, nsteps: int = 1, collisions: bool = True, nthreads: int = 1):
        U = C.c_double(0.0)
        self.hits += self.orc.lib.orc_step(
            self.n, self.x, self.y, self.z, self.vx, self.vy, self.vz, self.m, self.radius, self.vf32,
            self.ax, self.ay, self.az, self.dt, self.eps, self.G, self.restitution,
            1 if collisions else 0, nsteps, nthreads, C.byref(U))
        if nthreads == 1:
            self.U = U.value
        return self

    @property
    def pos(self):
        return np.stack([self.x, self.y, self.z], 1)

    @property
    def vel(self):
        return np.stack([self.vx, self.vy, self.vz], 1)

    @property
    def acc(self):
        return np.stack([self.ax, self.ay, self.az], 1)

    def kinetic(self):
        return float(self.orc.lib.orc_kinetic(self.n, self.vx, self.vy, self.vz, self.m, self.vf32))

    def angmom(self):
        L = np.empty(3)
        self.orc.lib.orc_angmom(self.n, self.x, self.y, self.z, self.vx, self.vy, self.vz, self.m, L)
        return L


_singleton = None


def load() -> COracle:
    global _singleton
    if _singleton is None:
        build()
        _singleton = COracle()
    return _singleton
