"""Batched ensemble of independent small systems (BASELINE config C3), one CTA per system.

Semantically `nsys` separate reference `SimulationEngine`s (core/engine.py:19-46,
65-97 of the reference) advanced in lockstep, without contact handling.  Systems
are independent, so multi-GPU runs block-partition them with no collectives
(`partition`).
"""
from __future__ import annotations

import numpy as np

from core import _native
from core.physics import default_device

_KEYS = ("x", "y", "z", "vx", "vy", "vz")


def partition(nsys: int, world_size: int, rank: int):
    """Contiguous block [lo, hi) of systems owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(int(nsys), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class EnsembleEngine:
    """State arrays are [nsys, nbody] fp64.

    mode="fast": rsqrt-seed kernel with all 32 lanes busy (<=1e-12 relative);
    mode="faithful": bit-exact with the reference engine on each system.
    `vel_f32=True` reproduces bodies built through `Object(...)` (float32 velocity storage).
    """

    @classmethod
    def from_elements(cls, M, e, a, inc, Omega, omega, m, dt: float, softening: float = 0.0, *,
                      G: float = 6.67430e-11, mode: str = "fast", vel_f32: bool = False, device: int | None = None):
        """Build the ensemble from orbital elements ON THE DEVICE (orb_ens_upload_elements, csrc/kepler.cu).

        Element arrays are [nsys, nbody-1] (radians / metres) for the bodies orbiting body 0, `m` is
        [nsys, nbody].  Replaces nsys*(nbody-1) host evaluations of the reference's `Body.get_state`
        (core/body.py:184-249) + `solve_kepler` (core/physics.py:43-71); body 0 rests at the origin.
        """
        M = np.asarray(M, dtype=np.float64)
        if M.ndim != 2:
            raise ValueError("element arrays must be [nsys, nbody-1]")
        self = cls.__new__(cls)
        self.nsys, self.nbody = M.shape[0], M.shape[1] + 1
        self.dt, self.softening, self.G = float(dt), float(softening), float(G)
        nat_mode = {"fast": _native.MODE_FAST, "faithful": _native.MODE_FAITHFUL}[mode]
        self._dev = _native.DeviceEnsemble(self.nsys, self.nbody, default_device() if device is None else device,
                                           nat_mode, vel_f32)
        self._dev.set_params(self.dt, self.softening, self.G)
        self._dev.upload_elements(M, e, a, inc, Omega, omega, m)
        self.steps_done = 0
        return self

    def __init__(self, x, y, z, vx, vy, vz, m, dt: float, softening: float = 0.0, *, G: float = 6.67430e-11,
                 mode: str = "fast", vel_f32: bool = False, device: int | None = None):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim != 2:
            raise ValueError("ensemble arrays must be [nsys, nbody]")
        self.nsys, self.nbody = x.shape
        self.dt, self.softening, self.G = float(dt), float(softening), float(G)
        nat_mode = {"fast": _native.MODE_FAST, "faithful": _native.MODE_FAITHFUL}[mode]
        vel = [np.asarray(a, dtype=np.float64) for a in (vx, vy, vz)]
        if vel_f32:     # Object.__init__ rounds constructor velocities to float32 (reference physics.py:184)
            vel = [a.astype(np.float32).astype(np.float64) for a in vel]
        self._dev = _native.DeviceEnsemble(self.nsys, self.nbody, default_device() if device is None else device,
                                           nat_mode, vel_f32)
        self._dev.set_params(self.dt, self.softening, self.G)
        self._dev.upload(x, y, z, *vel, m)
        self.steps_done = 0

    def step(self, nsteps: int = 1, fused: bool = True):
        """fused: all steps inside one launch (FP64-bound); else one launch per step (HBM-bound)."""
        self._dev.step(nsteps, fused)
        self.steps_done += int(nsteps)

    def state(self) -> dict:
        return self._dev.download()

    def energy(self) -> np.ndarray:
        return self._dev.energy()

    def synchronize(self):
        self._dev.synchronize()

    def close(self):
        self._dev.close()

    # algorithmic HBM bytes of one un-fused step: x,y,z,v,a read+write (9*16 B) + m read (8 B) per body
    BYTES_PER_BODY_STEP = 152
