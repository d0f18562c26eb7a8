"""Batched ensemble of independent small systems (BASELINE config C3).

Semantically `nsys` separate reference `SimulationEngine`s (core/engine.py:19-46,
65-97 of the reference) advanced in lockstep -- including, when radii are given,
the contact sweep every engine runs after its step (engine.py:85) and per-body
velocity dtypes (physics.py:184 vs :448-449).  Bit-exact mode: one warp per
system; fast mode: nbody/2 lanes per system, 64/nbody systems per warp (small batches: nbody lanes per system).
Systems are independent, so multi-GPU runs block-partition them with no
collectives (`partition`).
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
                 mode: str = "fast", vel_f32=False, device: int | None = None, radius=None,
                 restitution: float = 1.0):
        """`vel_f32`: bool for all bodies, or a [nsys, nbody] array of per-body flags.
        `radius`: [nsys, nbody] (or broadcastable); any radius > 0 turns the per-step contact sweep on."""
        x = np.asarray(x, dtype=np.float64)
        if x.ndim != 2:
            raise ValueError("ensemble arrays must be [nsys, nbody]")
        self.nsys, self.nbody = x.shape
        self.dt, self.softening, self.G = float(dt), float(softening), float(G)
        nat_mode = {"fast": _native.MODE_FAST, "faithful": _native.MODE_FAITHFUL}[mode]
        vel = [np.array(a, dtype=np.float64, copy=True) for a in (vx, vy, vz)]
        flags = None if np.isscalar(vel_f32) else np.broadcast_to(np.asarray(vel_f32, dtype=bool), x.shape)
        sel = flags if flags is not None else (np.ones(x.shape, bool) if vel_f32 else np.zeros(x.shape, bool))
        for a in vel:   # Object.__init__ rounds constructor velocities to float32 (reference physics.py:184)
            a[sel] = a[sel].astype(np.float32).astype(np.float64)
        self._dev = _native.DeviceEnsemble(self.nsys, self.nbody, default_device() if device is None else device,
                                           nat_mode, bool(vel_f32) if flags is None else False)
        self._dev.set_params(self.dt, self.softening, self.G)
        if radius is not None or flags is not None:
            self._dev.set_bodies(radius, flags)
            self._dev.set_contacts(restitution)
        self._dev.upload(x, y, z, *vel, m)
        self.steps_done = 0

    def step(self, nsteps: int = 1, fused: bool = True):
        """fused: all steps inside one launch (FP64-bound); else one launch per step (HBM-bound)."""
        self._dev.step(nsteps, fused)
        self.steps_done += int(nsteps)

    def state(self) -> dict:
        return self._dev.download()

    def acc(self) -> np.ndarray:
        """[3, nsys, nbody]: every system's accelerations of its last force build (engine.acc)."""
        return self._dev.download_acc()

    def contacts_resolved(self) -> int:
        return self._dev.contact_count()

    def energy(self) -> np.ndarray:
        return self._dev.energy()

    def synchronize(self):
        self._dev.synchronize()

    def close(self):
        self._dev.close()

    # HBM bytes of one un-fused step per body: x, u (half-kicked velocity), m read + x, u written
    BYTES_PER_BODY_STEP = 104
