#!/usr/bin/env python
"""Side-by-side parity / accuracy report (GPU box): GPU engine vs the reference's own outputs.

    python tools/parity_report.py > profiles/rN_parity_report.txt

Reference side: tests/golden/*.npz (outputs of the unmodified reference, see tests/golden/make_golden.py).
Reports, per BASELINE.json north_star: relative acceleration error per step, trajectory divergence and
total-energy drift over K steps, for the solar-system configs (configs[0]) in both velocity-dtype modes and
both engine modes, the N=4096 disk (configs[1]) and sampled rows of the N=262,144 Plummer sphere (configs[2]).
"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "orbital-physics_b200"), REPO):
    sys.path.insert(0, p)

from core import _native, synthetic  # noqa: E402
from core.engine import SimulationEngine  # noqa: E402
from core.physics import ObjectCollection  # noqa: E402
from oracle import load_c_oracle  # noqa: E402
from tests.conftest import make_objects  # noqa: E402

G = 6.67430e-11


def golden(name):
    return np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))


def state(eng):
    objs = eng.objects.objects
    pos = np.array([[o.coordinates.x, o.coordinates.y, o.coordinates.z] for o in objs], dtype=np.float64)
    acc = np.array([eng.acc[o.uuid] for o in objs])
    return pos, acc


def solar(name):
    g = golden(name)
    print(f"\n## {name}: N={len(g['in_x'])}, dt={float(g['dt'])}, eps={float(g['eps'])}  (reference = golden fixture)")
    print(f"{'mode':9s} {'K':>6s} {'max|dr|/|r|':>12s} {'max|da|/|a|':>12s} {'dE_ref':>11s} {'dE_gpu':>11s} {'bit-exact':>9s}")
    for mode in ("faithful", "fast"):
        eng = SimulationEngine(ObjectCollection(make_objects(g)), dt=float(g["dt"]), softening=float(g["eps"]),
                               cache=False, max_hist=None, mode=mode)
        E0 = eng.total_energy()
        done = 0
        for K in g["steps"]:
            K = int(K)
            eng.run(K - done)
            done = K
            p, a = state(eng)
            rp, ra = g[f"pos_{K}"], g[f"acc_{K}"]
            dr = (np.linalg.norm(p - rp, axis=1)[1:] / np.linalg.norm(rp, axis=1)[1:]).max()
            da = (np.linalg.norm(a - ra, axis=1) / np.linalg.norm(ra, axis=1)).max()
            dE_ref = (float(g[f"E_{K}"]) - float(g["E_0"])) / abs(float(g["E_0"]))
            dE_gpu = (eng.total_energy() - E0) / abs(E0)
            exact = np.array_equal(p, rp) and np.array_equal(a, ra)
            print(f"{mode:9s} {K:6d} {dr:12.3e} {da:12.3e} {dE_ref:11.3e} {dE_gpu:11.3e} {str(exact):>9s}")
        eng.close()


def disk():
    g = golden("disk4096_f32")
    c = synthetic.uniform_disk(4096)
    vel = [v.astype(np.float32).astype(np.float64) for v in (c["vx"], c["vy"], c["vz"])]
    print("\n## uniform disk N=4096 (configs[1]): reference engine ctor + 2 steps (fixture) vs GPU")
    print(f"{'mode':9s} {'step':>5s} {'max|da|/|a|':>12s} {'max|dr|/|r|':>12s} {'bit-exact':>9s}")
    for mode, m in (("faithful", _native.MODE_FAITHFUL), ("fast", _native.MODE_FAST)):
        dev = _native.DeviceSystem(4096, 0, m)
        dev.set_params(c["dt"], c["eps"], G)
        dev.upload(c["x"], c["y"], c["z"], *vel, c["m"], c["radius"], np.ones(4096, np.uint8))
        dev.accel()
        for s in (0, 1, 2):
            if s:
                dev.step(1)
            a = dev.download_acc().T
            st = dev.download_state()
            p = np.stack([st["x"], st["y"], st["z"]], 1)
            ra, rp = g[f"acc_{s}"], g[f"pos_{s}"]
            da = (np.linalg.norm(a - ra, axis=1) / np.linalg.norm(ra, axis=1)).max()
            dr = (np.linalg.norm(p - rp, axis=1)[1:] / np.linalg.norm(rp, axis=1)[1:]).max()
            print(f"{mode:9s} {s:5d} {da:12.3e} {dr:12.3e} {str(np.array_equal(a, ra) and np.array_equal(p, rp)):>9s}")
        t0 = time.perf_counter()
        dev.step(1000)
        dt_ms = (time.perf_counter() - t0)
        print(f"{mode:9s} 1000 further steps: {dt_ms:.3f} s  ({4096 * 4096 * 1000 / dt_ms:.3e} interactions/s; "
              f"the reference needs ~105 s per step, BASELINE.md)")
        dev.close()


def plummer():
    orc = load_c_oracle()
    c = synthetic.plummer(262144)
    rng = np.random.default_rng(0)
    rows = np.sort(rng.choice(c.n, 256, replace=False)).astype(np.int64)
    ref = orc.pairwise_sample(c["x"], c["y"], c["z"], c["m"], c["eps"], G, rows)
    ld, sum_abs = orc.pairwise_sample(c["x"], c["y"], c["z"], c["m"], c["eps"], G, rows, long_double=True)
    print("\n## Plummer N=262,144 (configs[2]): 256 sampled target rows; reference-order fp64 oracle and 80-bit yardstick")
    print(f"{'kernel':28s} {'max vs ref':>11s} {'median':>10s} {'max vs ld':>11s} {'max |da|/sum|a_ij|':>19s}")
    for sym in ("1", "0"):
        os.environ["ORBITAL_B200_SYM"] = sym
        dev = _native.DeviceSystem(c.n, 0, _native.MODE_FAST)
        dev.set_params(c["dt"], c["eps"], G)
        dev.upload(*c.arrays())
        dev.accel()
        a = dev.download_acc().T[rows]
        e_ref = np.linalg.norm(a - ref, axis=1) / np.linalg.norm(ref, axis=1)
        e_ld = np.linalg.norm(a - ld, axis=1) / np.linalg.norm(ld, axis=1)
        cond = np.linalg.norm(a - ld, axis=1) / sum_abs
        print(f"{dev.force_kernel_info()['name']:28s} {e_ref.max():11.2e} {np.median(e_ref):10.2e} {e_ld.max():11.2e} {cond.max():19.2e}")
        dev.close()
    e_o = np.linalg.norm(ref - ld, axis=1) / np.linalg.norm(ld, axis=1)
    print(f"{'oracle (reference order) itself':28s} {'-':>11s} {'-':>10s} {e_o.max():11.2e}")


def timing_small():
    print("\n## configs[0] timing: solar system, 10,000 steps (reference: 7.8 s at N=9, 22.3 s at N=15 on one core, BASELINE.md)")
    orc = load_c_oracle()
    from oracle.c_oracle import State
    for name in ("solar9_f32", "solar15_f32"):
        g = golden(name)
        eng = SimulationEngine(ObjectCollection(make_objects(g)), dt=float(g["dt"]), softening=float(g["eps"]),
                               cache=False, max_hist=None)
        eng.run(10); eng.synchronize()
        t0 = time.perf_counter(); eng.run(10000); eng.synchronize(); t_run = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(2000):
            eng.step()
        eng.synchronize(); t_step = (time.perf_counter() - t0) / 2000
        st = State(orc, g["in_x"], g["in_y"], g["in_z"], g["in_vx"], g["in_vy"], g["in_vz"], g["in_m"], g["in_radius"],
                   1, float(g["dt"]), float(g["eps"]))
        t0 = time.perf_counter(); st.step(10000); t_c = time.perf_counter() - t0
        n = len(g["in_x"])
        print(f"{name}: GPU run(10000) {t_run * 1e3:.1f} ms ({t_run / 10000 * 1e6:.2f} us/step, one launch); "
              f"GPU step() from Python {t_step * 1e6:.1f} us/step; C oracle (1 thread) {t_c * 1e3:.1f} ms; "
              f"N(N-1)={n * (n - 1)} ordered interactions/step")
        eng.close()


if __name__ == "__main__":
    info = _native.device_info(0)
    print(f"# parity report on {info['name']} ({info['sm_count']} SMs)")
    for nm in ("solar15_f32", "solar15_f64", "solar9_f32", "solar26_f64"):
        solar(nm)
    disk()
    plummer()
    timing_small()
