"""CPU stand-in for core._native.DeviceSystem -- TEST INFRASTRUCTURE ONLY.

Lets the `-m "not gpu"` suite exercise the *host* logic of the engine (lazy
mirrors, dirty tracking, history ring bookkeeping, JSONL frames, run() chunking,
contact resolution) without a GPU by emulating the C ABI's observable behaviour
with the oracle (oracle/nbody_oracle.c).  The product never imports this module;
tests install it with `monkeypatch.setattr(core._native, "DeviceSystem", ...)`.
The `-m gpu` tests run the same scenarios against the real library.
"""
from __future__ import annotations

import numpy as np

from oracle import load_c_oracle
from oracle.c_oracle import State


class FakeDeviceSystem:
    instances = 0

    def __init__(self, n, device=0, mode=0, tgt_lo=None, tgt_hi=None):
        self.n, self.device, self.mode = int(n), device, mode
        self.orc = load_c_oracle()
        self.dt, self.eps, self.G = 1.0, 0.0, 6.67430e-11
        self.st = None
        self.acc = np.zeros((3, self.n))
        self.hist_cap = 0
        self.hist = []
        self.hist_total = 0
        self.pairs = np.empty((0, 2), dtype=np.int64)
        self.launches = 0
        self.closed = False
        FakeDeviceSystem.instances += 1

    def close(self):
        self.closed = True

    def set_params(self, dt, eps, G=6.67430e-11):
        self.dt, self.eps, self.G = float(dt), float(eps), float(G)
        if self.st is not None:
            self.st.dt, self.st.eps, self.st.G = self.dt, self.eps, self.G

    def set_mode(self, mode):
        self.mode = mode

    def set_history(self, capacity):
        self.hist_cap = int(capacity)
        self.hist, self.hist_total = [], 0

    def set_stream(self, s):
        pass

    def upload(self, x, y, z, vx, vy, vz, m, radius, vel_is_f32=None):
        f = np.zeros(self.n, np.uint8) if vel_is_f32 is None else np.asarray(vel_is_f32, np.uint8)
        keep = self.acc
        st = State.__new__(State)
        st.orc = self.orc
        c = lambda a: np.array(a, dtype=np.float64, copy=True)
        st.x, st.y, st.z, st.vx, st.vy, st.vz, st.m, st.radius = map(c, (x, y, z, vx, vy, vz, m, radius))
        st.n = self.n
        st.vf32 = np.ascontiguousarray(np.broadcast_to(f, (self.n,))).copy()
        st.dt, st.eps, st.G, st.restitution = self.dt, self.eps, self.G, 1.0
        st.ax, st.ay, st.az = keep[0].copy(), keep[1].copy(), keep[2].copy()
        st.U, st.hits = 0.0, 0
        self.st = st

    def download_state(self, out=None):
        s = self.st
        return {k: getattr(s, k).copy() for k in ("x", "y", "z", "vx", "vy", "vz")}

    def download_acc(self):
        return np.stack([self.st.ax, self.st.ay, self.st.az]).copy()

    def upload_acc(self, a):
        self.st.ax, self.st.ay, self.st.az = (np.array(a[k], dtype=np.float64, copy=True) for k in range(3))
        self.acc = np.stack([self.st.ax, self.st.ay, self.st.az])

    def accel(self):
        s = self.st
        a, _ = self.orc.pairwise(s.x, s.y, s.z, s.m, self.eps, self.G, 1)
        s.ax, s.ay, s.az = (np.ascontiguousarray(a[:, k]) for k in range(3))
        self.acc = np.stack([s.ax, s.ay, s.az])
        self.launches += 1

    def _overlaps(self):
        s = self.st
        P = np.stack([s.x, s.y, s.z], 1)
        out = []
        for i in range(self.n):
            d = P[i] - P[i + 1:]
            # same rounding as np.linalg.norm on 3-vectors: sqrt(fma chain)
            for k in range(d.shape[0]):
                if np.linalg.norm(d[k]) <= s.radius[i] + s.radius[i + 1 + k]:
                    out.append((i, i + 1 + k))
        return np.array(out, dtype=np.int64).reshape(-1, 2)

    def step(self, nsteps=1):
        done = 0
        self.pairs = np.empty((0, 2), dtype=np.int64)
        detect = bool((self.st.radius > 0).any())
        for _ in range(int(nsteps)):
            self.st.step(1, collisions=False)
            self.acc = np.stack([self.st.ax, self.st.ay, self.st.az])
            done += 1
            self.launches += 4
            if detect:
                ov = self._overlaps()
                if len(ov):
                    self.pairs = ov[::-1].copy()          # unsorted on purpose
                    return done, len(ov)
            self._append()
        return done, 0

    def overlap_pairs(self, cap=1 << 16):
        return self.pairs[:cap].copy(), len(self.pairs)

    def synchronize(self):
        pass

    def force_kernel_info(self):
        return {"name": "fake", "grid": 1, "block": 32, "smem": 0, "launches_per_step": 4}

    def launch_count(self):
        return self.launches

    def potential(self):
        s = self.st
        return self.orc.potential(s.x, s.y, s.z, s.m, self.eps, self.G)

    def energy_angmom(self):
        s = self.st
        return s.kinetic(), s.angmom()

    def _append(self):
        if self.hist_cap <= 0:
            return
        s = self.st
        self.hist.append(np.stack([s.x, s.y, s.z], 1).copy())
        self.hist = self.hist[-self.hist_cap:]
        self.hist_total += 1

    def history_count(self):
        return self.hist_total

    def history_append(self):
        self._append()

    def history_download(self, last_k):
        k = min(int(last_k), len(self.hist))
        if k <= 0:
            return np.empty((0, self.n, 3))
        return np.stack(self.hist[-k:])
