"""NumPy restatement of the reference hot path (test infrastructure).

Uses the *same NumPy primitives* the reference uses at each call site so the
host-BLAS-dependent rounding of ``rij @ rij`` is inherited rather than assumed:
  pairwise()        <- core/physics.py:125-159
  kick()/drift()    <- core/engine.py:69-75,81-82
  rows_vectorised() <- row-parallel form for sampled rows at large N
  solve_kepler() / kepler_states() <- core/physics.py:43-71, core/body.py:184-249 (IC pipeline)
Pure-Python loops: use only at small N.
"""
from __future__ import annotations

import math

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


def solve_kepler(M: float, e: float, tol: float = 1e-12, max_iter: int = 50, sin=math.sin, cos=math.cos) -> float:
    """physics.py:43-71: Newton on E - e sin E = M, start at M (e < 0.8) or pi; test |dE| after the update.
    `sin` / `cos`: the reference calls the host libm (the default); tests pass the correctly rounded pair
    (csrc/sincos_cr.h built for the host) to get the exact bits the device pipeline must produce."""
    E = M if e < 0.8 else math.pi
    for _ in range(max_iter):
        dE = -(E - e * sin(E) - M) / (1.0 - e * cos(E))
        E += dE
        if abs(dE) < tol:
            break
    return E


def kepler_states(M, e, a, b, n, inc, Omega, omega, tol: float = 1e-12, max_iter: int = 50, sin=math.sin,
                  cos=math.cos):
    """body.py:184-249 over arrays: parent-relative (r[count,3], v[count,3], E[count]); trig = host libm unless
    `sin` / `cos` are given (see solve_kepler)."""
    cnt = len(M)
    r, v, Eo = np.empty((cnt, 3)), np.empty((cnt, 3)), np.empty(cnt)
    for k in range(cnt):
        ek, ak, bk, nk = float(e[k]), float(a[k]), float(b[k]), float(n[k])
        E = solve_kepler(float(M[k]), ek, tol, max_iter, sin, cos)
        cE, sE = cos(E), sin(E)
        x_op = ak * (cE - ek)
        y_op = bk * sE
        vx_op = -ak * nk * sE / (1 - ek * cE)
        vy_op = ak * nk * math.sqrt(1 - ek ** 2) * cE / (1 - ek * cE)
        cw, sw = cos(float(omega[k])), sin(float(omega[k]))
        ci, si = cos(float(inc[k])), sin(float(inc[k]))
        cO, sO = cos(float(Omega[k])), sin(float(Omega[k]))
        R = ((cO * cw - sO * sw * ci, -cO * sw - sO * cw * ci, sO * si),
             (sO * cw + cO * sw * ci, -sO * sw + cO * cw * ci, -cO * si),
             (sw * si, cw * si, ci))
        for c in range(3):
            r[k, c] = R[c][0] * x_op + R[c][1] * y_op + R[c][2] * 0.0
            v[k, c] = R[c][0] * vx_op + R[c][1] * vy_op + R[c][2] * 0.0
        Eo[k] = E
    return r, v, Eo


def ensemble_from_elements(M, e, a, inc, Omega, omega, m, G: float = G_SI, sin=math.sin, cos=math.cos):
    """What orb_ens_upload_elements builds: central body at rest at the origin, the others parent-relative with
    n = sqrt(G m_0 / a**3) (body.py:159-169) and b = a sqrt(1 - e**2) (body.py:120-124). Arrays [nsys, nbody-1]."""
    nsys, k = np.shape(M)
    out = {key: np.zeros((nsys, k + 1)) for key in ("x", "y", "z", "vx", "vy", "vz")}
    for s in range(nsys):
        mu = G * float(m[s, 0])
        nn = np.array([math.sqrt(mu / float(a[s, j]) ** 3) for j in range(k)])
        bb = np.array([float(a[s, j]) * math.sqrt(1 - float(e[s, j]) ** 2) for j in range(k)])
        r, v, _ = kepler_states(M[s], e[s], a[s], bb, nn, inc[s], Omega[s], omega[s], sin=sin, cos=cos)
        for c, key in enumerate(("x", "y", "z")):
            out[key][s, 1:] = r[:, c]
        for c, key in enumerate(("vx", "vy", "vz")):
            out[key][s, 1:] = v[:, c]
    return out
