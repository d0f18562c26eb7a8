"""Synthetic initial conditions for the benchmark / parity workloads.

These are the concrete inputs SURVEY.md section 8(d) fixes for BASELINE.json's
configs C1..C4.  They return plain structure-of-arrays fp64 NumPy arrays
(``x y z vx vy vz m radius``) so they can be fed to the device engine, to the
oracle, or (small N) turned into reference-style ``Object`` lists.

The module is deliberately dependency-free (NumPy only, no package-relative
imports) so the golden-vector generator can load it by file path next to the
*reference's* ``core`` package without a name clash.
"""
from __future__ import annotations

import numpy as np

G_SI = 6.67430e-11            # reference: core/constants.py:49-58 (STANDARD.G)
AU_M = 1.495978707e11         # reference: core/constants.py:7
M_SUN = 1.98847e30            # reference: core/units.py:8


class Cloud(dict):
    """SoA body set: keys x y z vx vy vz m radius (+ eps, dt, G hints)."""

    __getattr__ = dict.__getitem__

    @property
    def n(self) -> int:
        return int(self["x"].shape[0])

    def arrays(self):
        return tuple(self[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m", "radius"))


def _cloud(pos, vel, m, radius, **hints) -> Cloud:
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    vel = np.ascontiguousarray(vel, dtype=np.float64)
    c = Cloud(
        x=np.ascontiguousarray(pos[:, 0]), y=np.ascontiguousarray(pos[:, 1]), z=np.ascontiguousarray(pos[:, 2]),
        vx=np.ascontiguousarray(vel[:, 0]), vy=np.ascontiguousarray(vel[:, 1]), vz=np.ascontiguousarray(vel[:, 2]),
        m=np.ascontiguousarray(m, dtype=np.float64),
        radius=np.ascontiguousarray(np.broadcast_to(np.asarray(radius, dtype=np.float64), (pos.shape[0],))),
    )
    c.update(hints)
    return c


def uniform_disk(n: int = 4096, seed: int | None = None) -> Cloud:
    """C1: thin uniform-surface-density disk around a central solar mass.

    Body 0 is the central mass at rest at the origin; bodies 1..n-1 are on
    circular Keplerian orbits between 0.5 and 5 AU with total disk mass
    1e-3 M_c.  eps = 1e8 m, dt = 3600 s.
    """
    rng = np.random.default_rng(n if seed is None else seed)
    k = n - 1
    r_in, r_out = 0.5 * AU_M, 5.0 * AU_M
    u = rng.random(k)
    r = np.sqrt(u * (r_out**2 - r_in**2) + r_in**2)
    th = rng.random(k) * (2.0 * np.pi)
    z = rng.standard_normal(k) * (1e-3 * r)
    pos = np.zeros((n, 3))
    pos[1:, 0] = r * np.cos(th)
    pos[1:, 1] = r * np.sin(th)
    pos[1:, 2] = z
    speed = np.sqrt(G_SI * M_SUN / r)
    vel = np.zeros((n, 3))
    vel[1:, 0] = -speed * np.sin(th)      # z_hat x r_hat
    vel[1:, 1] = speed * np.cos(th)
    m = np.full(n, 1e-3 * M_SUN / k)
    m[0] = M_SUN
    return _cloud(pos, vel, m, 1.0, eps=1e8, dt=3600.0, G=G_SI)


def _isotropic(rng, k):
    cz = rng.uniform(-1.0, 1.0, k)
    ph = rng.uniform(0.0, 2.0 * np.pi, k)
    s = np.sqrt(np.maximum(0.0, 1.0 - cz * cz))
    return np.stack([s * np.cos(ph), s * np.sin(ph), cz], axis=1)


def plummer(n: int = 262144, seed: int | None = None, a: float = 1e12, m_each: float = 1e24) -> Cloud:
    """C2/C4: Plummer sphere, Aarseth-Henon-Wielen sampling.

    r = a / sqrt(u^(-2/3) - 1); speed = q * v_esc with q drawn by rejection on
    g(q) = q^2 (1 - q^2)^(7/2); centre of mass and net momentum removed.
    radius = 0, eps = 1e9 m, dt = 2^-10 sqrt(a^3 / (G M)).
    """
    rng = np.random.default_rng(n if seed is None else seed)
    M = n * m_each
    u = np.clip(rng.random(n), 1e-10, 1.0 - 1e-16)
    r = a / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    pos = _isotropic(rng, n) * r[:, None]
    q = np.empty(n)
    todo = np.arange(n)
    while todo.size:
        cand = rng.random(todo.size)
        y = rng.random(todo.size) * 0.1
        ok = y < cand * cand * (1.0 - cand * cand) ** 3.5
        q[todo[ok]] = cand[ok]
        todo = todo[~ok]
    v_esc = np.sqrt(2.0 * G_SI * M / np.sqrt(r * r + a * a))
    vel = _isotropic(rng, n) * (q * v_esc)[:, None]
    m = np.full(n, m_each)
    pos -= pos.mean(axis=0)
    vel -= vel.mean(axis=0)
    dt = 2.0 ** -10 * np.sqrt(a**3 / (G_SI * M))
    return _cloud(pos, vel, m, 0.0, eps=1e9, dt=float(dt), G=G_SI)


def planetary_system(system_index: int, nbody: int = 16) -> Cloud:
    """C3: one member of the ensemble (seed 10_000 + index): star + planets."""
    rng = np.random.default_rng(10_000 + int(system_index))
    k = nbody - 1
    a = np.exp(rng.uniform(np.log(0.3), np.log(30.0), k)) * AU_M
    mp = np.exp(rng.uniform(np.log(1e23), np.log(1e27), k))
    ph = rng.uniform(0.0, 2.0 * np.pi, k)
    inc = rng.standard_normal(k) * np.deg2rad(2.0)
    pos = np.zeros((nbody, 3))
    vel = np.zeros((nbody, 3))
    pos[1:, 0] = a * np.cos(ph)
    pos[1:, 1] = a * np.sin(ph) * np.cos(inc)
    pos[1:, 2] = a * np.sin(ph) * np.sin(inc)
    speed = np.sqrt(G_SI * M_SUN / a)
    vel[1:, 0] = -speed * np.sin(ph)
    vel[1:, 1] = speed * np.cos(ph) * np.cos(inc)
    vel[1:, 2] = speed * np.cos(ph) * np.sin(inc)
    m = np.concatenate([[M_SUN], mp])
    radius = np.concatenate([[6.9634e8], 6.371e6 * (mp / 5.9722e24) ** (1.0 / 3.0)])
    return _cloud(pos, vel, m, radius, eps=1e6, dt=86400.0, G=G_SI)


def ensemble(nsys: int, nbody: int = 16, first: int = 0):
    """C3 batch: arrays shaped [nsys, nbody] for x y z vx vy vz m radius."""
    out = {k: np.empty((nsys, nbody)) for k in ("x", "y", "z", "vx", "vy", "vz", "m", "radius")}
    for s in range(nsys):
        c = planetary_system(first + s, nbody)
        for k in out:
            out[k][s] = c[k]
    out.update(eps=1e6, dt=86400.0, G=G_SI)
    return out


def ensemble_fast(nsys: int, nbody: int = 16, seed: int = 10_000):
    """Vectorised variant of :func:`ensemble` for 65,536-system benchmark runs.

    Same distributions, one generator for the whole batch (so it is *not*
    member-for-member identical to ``planetary_system``; parity tests use the
    per-system form, throughput runs use this one).
    """
    rng = np.random.default_rng(seed)
    k = nbody - 1
    a = np.exp(rng.uniform(np.log(0.3), np.log(30.0), (nsys, k))) * AU_M
    mp = np.exp(rng.uniform(np.log(1e23), np.log(1e27), (nsys, k)))
    ph = rng.uniform(0.0, 2.0 * np.pi, (nsys, k))
    inc = rng.standard_normal((nsys, k)) * np.deg2rad(2.0)
    z0 = np.zeros((nsys, 1))
    speed = np.sqrt(G_SI * M_SUN / a)
    out = dict(
        x=np.concatenate([z0, a * np.cos(ph)], 1),
        y=np.concatenate([z0, a * np.sin(ph) * np.cos(inc)], 1),
        z=np.concatenate([z0, a * np.sin(ph) * np.sin(inc)], 1),
        vx=np.concatenate([z0, -speed * np.sin(ph)], 1),
        vy=np.concatenate([z0, speed * np.cos(ph) * np.cos(inc)], 1),
        vz=np.concatenate([z0, speed * np.cos(ph) * np.sin(inc)], 1),
        m=np.concatenate([np.full((nsys, 1), M_SUN), mp], 1),
        radius=np.concatenate([np.full((nsys, 1), 6.9634e8), 6.371e6 * (mp / 5.9722e24) ** (1.0 / 3.0)], 1),
    )
    out = {k_: np.ascontiguousarray(v) for k_, v in out.items()}
    out.update(eps=1e6, dt=86400.0, G=G_SI)
    return out


def random_cloud(n: int, seed: int = 0, scale: float = 1e11, mass_lo: float = 1e20, mass_hi: float = 1e28,
                 vscale: float = 3e4, radius: float = 0.0) -> Cloud:
    """Generic random cloud used by the force-parity tests."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-scale, scale, (n, 3))
    vel = rng.standard_normal((n, 3)) * vscale
    m = np.exp(rng.uniform(np.log(mass_lo), np.log(mass_hi), n))
    return _cloud(pos, vel, m, radius, eps=1e-3 * scale, dt=3600.0, G=G_SI)
