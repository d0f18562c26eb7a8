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
        self.restitution, self.device_contacts = 1.0, False
        self.u_stash = None
        FakeDeviceSystem.instances += 1

    def close(self):
        self.closed = True

    def set_params(self, dt, eps, G=6.67430e-11):
        self.dt, self.eps, self.G = float(dt), float(eps), float(G)
        if self.st is not None:
            self.st.dt, self.st.eps, self.st.G = self.dt, self.eps, self.G

    def set_mode(self, mode):
        self.mode = mode

    def set_contacts(self, restitution, on_device):
        self.restitution, self.device_contacts = float(restitution), bool(on_device)

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
        resolved = 0
        for _ in range(int(nsteps)):
            self.u_stash = None
            if self.device_contacts:
                # emulate orb_set_contacts(on_device=1): the oracle's own sweep, U stashed before the push-out
                self.st.restitution = self.restitution
                before = self.st.hits
                self.st.step(1, collisions=True)
                self.acc = np.stack([self.st.ax, self.st.ay, self.st.az])
                if self.st.hits > before:
                    self.u_stash = self.st.U
                    resolved += self.st.hits - before
                done += 1
                self.launches += 8
                self._append()
                continue
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
        return done, resolved

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
        if self.u_stash is not None:
            return self.u_stash
        return self.orc.potential(s.x, s.y, s.z, s.m, self.eps, self.G)

    def body_potential(self, body, G):
        """orb_body_potential: the potential loop of the reference's Object.lagrangian (physics.py:275-279)."""
        s = self.st
        here = np.array([s.x[body], s.y[body], s.z[body]])
        pe = 0
        for j in range(len(s.m)):
            if j != body:
                r = np.linalg.norm(here - np.array([s.x[j], s.y[j], s.z[j]]))
                pe += -G * float(s.m[body]) * float(s.m[j]) / r
        return float(pe)

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


class FakeShardedDevice:
    """CPU stand-in for one rank's *sharded* DeviceSystem (orb_create_ranked and the split-step entry points):
    per-rank force rows, overlap flags and kicks from the oracle, the replicated contact sweep from the
    oracle's own handle_collisions restatement."""

    def __init__(self, n, device=0, mode=0, tgt_lo=0, tgt_hi=None, rank=None, world=None):
        self.n, self.lo, self.hi = int(n), int(tgt_lo), int(n if tgt_hi is None else tgt_hi)
        self.device, self.mode = device, mode
        self.world = int(world or 1)
        self.orc = load_c_oracle()
        cap = self.world * (-(-self.n // self.world))
        self.pos4 = np.zeros((cap, 4))
        self.vel = np.zeros((3, self.n))
        self.acc = np.zeros((3, self.n))
        self.radius = np.zeros(self.n)
        self.f32 = np.zeros(self.n, bool)
        self.partial = False       # True: emulate the pair-symmetric sharding (partial acc on every rank)
        self.restitution = 1.0
        self.local_pairs = np.empty((0, 2), dtype=np.int64)
        self.merged = np.empty((0, 2), dtype=np.int64)
        self.contacts_total = 0
        self.hist_cap, self.hist, self.hist_total = 0, [], 0
        self.u_stash = None
        self.launches = 0

    def close(self):
        pass

    def set_params(self, dt, eps, G=6.67430e-11):
        self.dt, self.eps, self.G = float(dt), float(eps), float(G)

    def set_contacts(self, restitution, on_device):
        assert on_device
        self.restitution = float(restitution)

    def set_history(self, capacity):
        self.hist_cap, self.hist, self.hist_total = int(capacity), [], 0

    def set_stream(self, s):
        pass

    def set_mode(self, mode):
        self.mode = mode

    def upload(self, x, y, z, vx, vy, vz, m, radius, vel_is_f32=None):
        n = self.n
        self.pos4[:n, 0], self.pos4[:n, 1], self.pos4[:n, 2], self.pos4[:n, 3] = x, y, z, m
        self.vel[0], self.vel[1], self.vel[2] = vx, vy, vz
        self.radius = np.array(radius, dtype=np.float64, copy=True)
        self.f32 = np.zeros(n, bool) if vel_is_f32 is None else np.asarray(vel_is_f32, bool).copy()

    def _cols(self):
        p = self.pos4[: self.n]
        return [np.ascontiguousarray(p[:, k]) for k in range(4)]

    def accel(self):
        rows = np.arange(self.lo, self.hi, dtype=np.int64)
        if self.partial:
            # partial accelerations of ALL bodies whose sum over ranks is the full field: this rank
            # contributes the rows of its own slab and zeros elsewhere (bit-exact after the all-reduce)
            self.acc[:] = 0.0
        x, y, z, m = self._cols()
        a = self.orc.pairwise_sample(x, y, z, m, self.eps, self.G, rows)
        self.acc[:, self.lo:self.hi] = a.T
        self.u_stash = None
        self.launches += 1

    def step_force(self):
        self.accel()
        out = []
        if (self.radius > 0).any():
            P = self.pos4[: self.n, :3]
            for i in range(self.lo, self.hi):
                d = P[i] - P[i + 1:]
                for k in range(d.shape[0]):
                    if np.linalg.norm(d[k]) <= self.radius[i] + self.radius[i + 1 + k]:
                        out.append((i, i + 1 + k))
        self.local_pairs = np.array(out, dtype=np.int64).reshape(-1, 2)
        self.merged = self.local_pairs

    def _kick(self):
        s = slice(self.lo, self.hi)
        h = 0.5 * self.dt
        v = self.vel[:, s] + h * self.acc[:, s]
        f = self.f32[s]
        v[:, f] = v[:, f].astype(np.float32).astype(np.float64)
        self.vel[:, s] = v

    def step_begin(self):
        self._kick()
        s = slice(self.lo, self.hi)
        f = self.f32[s]
        v = self.vel[:, s]
        step = v * self.dt
        step[:, f] = (v[:, f].astype(np.float32) * np.float32(self.dt)).astype(np.float64)
        self.pos4[s, :3] = self.pos4[s, :3] + step.T
        self.launches += 1

    def step_finish(self):
        self.step_force()
        self._kick()

    def step_kick(self):
        self._kick()
        self.launches += 1

    def overlap_count(self):
        return len(self.local_pairs), False

    def overlap_pairs(self, cap=1 << 16):
        return self.local_pairs[:cap][::-1].copy(), len(self.local_pairs)

    def set_overlap_pairs(self, pairs, overflowed=False):
        self.merged = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)

    def step_end(self):
        if len(self.merged):
            x, y, z, m = self._cols()
            if self.n <= 4096:
                self.u_stash = self.orc.potential(x, y, z, m, self.eps, self.G)
            vx, vy, vz = (np.ascontiguousarray(self.vel[k]) for k in range(3))
            hits = self.orc.lib.orc_collisions(self.n, x, y, z, vx, vy, vz, m, np.ascontiguousarray(self.radius),
                                               self.f32.astype(np.uint8), self.restitution)
            self.contacts_total += int(hits)
            self.pos4[: self.n, 0], self.pos4[: self.n, 1], self.pos4[: self.n, 2] = x, y, z
            self.vel[0], self.vel[1], self.vel[2] = vx, vy, vz
        self.local_pairs = self.merged = np.empty((0, 2), dtype=np.int64)
        self.history_append()

    def contact_stats(self):
        return {"contacts_total": self.contacts_total, "full_sweeps": 0}

    def acc_needs_allreduce(self):
        return self.partial

    def synchronize(self):
        pass

    def download_state(self, out=None):
        n = self.n
        return {"x": self.pos4[:n, 0].copy(), "y": self.pos4[:n, 1].copy(), "z": self.pos4[:n, 2].copy(),
                "vx": self.vel[0].copy(), "vy": self.vel[1].copy(), "vz": self.vel[2].copy()}

    def download_acc(self):
        return self.acc.copy()

    def upload_acc(self, a):
        self.acc[:] = np.asarray(a, dtype=np.float64)

    def potential(self):
        if self.u_stash is not None:
            return self.u_stash
        x, y, z, m = self._cols()
        return self.orc.potential(x, y, z, m, self.eps, self.G)

    def energy_angmom(self):
        s = slice(self.lo, self.hi)
        m = self.pos4[s, 3]
        K = float(np.sum(0.5 * m * (self.vel[:, s] ** 2).sum(0)))
        L = np.cross(self.pos4[s, :3], (m * self.vel[:, s]).T).sum(0)
        return K, L

    def history_count(self):
        return self.hist_total

    def history_append(self):
        if self.hist_cap <= 0:
            return
        self.hist.append(self.pos4[: self.n, :3].copy())
        self.hist = self.hist[-self.hist_cap:]
        self.hist_total += 1

    def history_download(self, last_k):
        k = min(int(last_k), len(self.hist))
        if k <= 0:
            return np.empty((0, self.n, 3))
        return np.stack(self.hist[-k:])

    def force_kernel_info(self):
        return {"name": "fake-sharded", "grid": 1, "block": 32, "smem": 0, "launches_per_step": 4}

    def launch_count(self):
        return self.launches
