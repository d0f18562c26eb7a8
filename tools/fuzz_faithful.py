#!/usr/bin/env python
"""Differential fuzz of the bit-exact device path against the C oracle (development aid; GPU box).

Random sizes across the micro / kernel-sequence (two-pass) regimes, random radii (contacts or none), mixed
float32/float64 velocity bodies, contacts resolved on the device, random step counts; every state bit is compared.

    python tools/fuzz_faithful.py [seconds] [seed]
"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "orbital-physics_b200"), REPO):
    sys.path.insert(0, p)
from core import _native as nat  # noqa: E402
from oracle import load_c_oracle  # noqa: E402
from oracle.c_oracle import State  # noqa: E402

G = 6.67430e-11


def same(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    return bool(((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))).all())


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    orc = load_c_oracle()
    t0, cases, hits_total = time.time(), 0, 0
    while time.time() - t0 < budget:
        n = int(rng.choice([2, 3, 7, 15, 33, 64, 65, 66, 97, 130, 257, 400, 513, 700, 1100]))
        contacts = bool(rng.integers(0, 2))
        box = (2.5e3 * n ** (1 / 3) * 6) if contacts else 1e9
        x, y, z = (rng.uniform(-box, box, n) for _ in range(3))
        v = rng.standard_normal((3, n)) * 300
        m = np.exp(rng.uniform(np.log(1e13), np.log(1e17), n))
        radius = rng.uniform(1e3, 5e3, n) if contacts else (np.zeros(n) if rng.integers(0, 2) else np.full(n, 1.0))
        f32 = rng.integers(0, 2, n).astype(np.uint8)
        vel = [np.where(f32 == 1, a.astype(np.float32).astype(np.float64), a) for a in v]
        dt, eps, rest = float(rng.choice([0.5, 2.0, 7.0])), float(rng.choice([0.0, 10.0, 1e3])), float(rng.choice([1.0, 0.6]))
        st = State(orc, x, y, z, *vel, m, radius, f32, dt, eps, G, restitution=rest)
        dev = nat.DeviceSystem(n, 0, nat.MODE_FAITHFUL)
        dev.set_params(dt, eps, G)
        dev.set_contacts(rest, True)
        dev.set_history(5)
        dev.upload(x, y, z, *vel, m, radius, f32)
        dev.accel()
        for k in (1, int(rng.integers(1, 20))):
            done, _ = dev.step(k)
            for _ in range(k):
                st.step(1, collisions=True)
            s = dev.download_state()
            ok = (done == k and same(np.stack([s["x"], s["y"], s["z"]], 1), st.pos)
                  and same(np.stack([s["vx"], s["vy"], s["vz"]], 1), st.vel) and same(dev.download_acc().T, st.acc)
                  and same(dev.history_download(1)[0], st.pos))
            if not ok:
                print(f"MISMATCH n={n} contacts={contacts} dt={dt} eps={eps} rest={rest} k={k} "
                      f"kernel={dev.force_kernel_info()['name']}")
                sys.exit(1)
        hits_total += st.hits
        cases += 1
        dev.close()
    print(f"fuzz ok: {cases} cases, {hits_total} contacts resolved, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
