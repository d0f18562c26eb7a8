"""NumPy restatement of the reference hot path (test infrastructure).

Uses the *same NumPy primitives* the reference uses at each call site so the
host-BLAS-dependent rounding of ``rij @ rij`` is inherited rather than assumed:
  pairwise()        <- core/physics.py:125-159
  kick()/drift()    <- core/engine.py:69-75,81-82
  rows_vectorised() <- row-parallel form for sampled rows at large N
Pure-Python loops: use only at small N.
"""
from __future__ import annotations

import numpy as np

G_SI = 6.67430e-11


def pairwise(pos: np.ndarray, m: np.ndarray, eps: float = 0.0, G: float = G_SI):
    """physics.py:125-159 on [n,3] positions. Returns (acc[n,3], U)."""
    n = pos.shape[0]
    acc = [np.zeros(3) for _ in range(n)]
    U = 0.0
    eps2 = eps * eps
    P = [np.array([float(p[0]), float(p[1]), float(p[2])]) for p in pos]
    M = [float(v) for v in m]
    for i in range(n):
        for j in range(i + 1, n):
            rij = P[j] - P[i]
            r2 = float(rij @ rij) + eps2
            inv_r = 1.0 / np.sqrt(r2)
            inv_r3 = inv_r / r2
            acc[i] += G * M[j] * inv_r3 * rij
            acc[j] += -G * M[i] * inv_r3 * rij
            U += -G * M[i] * M[j] * inv_r
    return np.array(acc), float(U)


def kick(v: np.ndarray, a: np.ndarray, dt: float) -> np.ndarray:
    """engine.py:70 `obj.velocity += 0.5 * dt * acc` (dtype of v preserved)."""
    v = v.copy()
    v += 0.5 * dt * a
    return v


def drift(r: np.ndarray, v: np.ndarray, dt: float) -> np.ndarray:
    """engine.py:74 `obj.position() + obj.velocity * dt` (f32 v: product in f32)."""
    return r + v * dt


def rows_vectorised(pos: np.ndarray, m: np.ndarray, rows, eps: float, G: float = G_SI, dtype=np.float64):
    """Row-vectorised accelerations for selected targets (not rounding-faithful:
    NumPy's pairwise summation order; validated to ~1e-15 against the faithful form)."""
    pos = pos.astype(dtype)
    m = m.astype(dtype)
    out = np.zeros((len(rows), 3), dtype=dtype)
    eps2 = dtype(eps) * dtype(eps)
    for k, i in enumerate(rows):
        d = pos - pos[i]
        r2 = (d * d).sum(1) + eps2
        r2[i] = 1.0
        w = dtype(G) * m / (r2 * np.sqrt(r2))
        w[i] = 0.0
        out[k] = (w[:, None] * d).sum(0)
    return out
