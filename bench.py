#!/usr/bin/env python
"""Benchmark of the hot path: all-pairs gravity + leapfrog step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n-bodies N]

metric  : pairwise interactions/s (fp64), N^2 ordered interactions per force evaluation,
          one force evaluation per leapfrog step.
workload: N=1 GPU  -> BASELINE configs[2]: Plummer sphere N=262,144, fast (roofline) kernel.
          N>1 GPUs -> BASELINE configs[4]: Plummer sphere N=2,097,152 over the ranks (strong scaling): every
                      unordered pair is evaluated once, pair blocks dealt to the ranks in snake order; per step an
                      all-gather of the packed positions (32 B x N) and a reduce-scatter of the partial
                      accelerations (24 B x N in, every rank keeps its slab) over NCCL.
A "step" is one pass of the hot path: half-kick+drift -> force -> half-kick, through the C ABI.

Extra records in the same JSON line (none of them changes `value`):
  parity_check       64 sampled rows of the final accelerations against the oracle at the final positions
  scale_denominator  (N=1) a few timed steps at N=2,097,152 on one GPU: the denominator of the 2/4/8-GPU efficiency
  strong_262144      (N>1) the same N=262,144 workload as the 1-GPU line, spread over the ranks
  comm_ms            (N>1) mean CUDA-event time of the two collectives per step
  ic_pipeline        (N=1) orbital elements -> states on the device (262,144 bodies) and a live bit-for-bit check against
                     the unmodified reference's Body.get_state
  configs            (N=1) BASELINE configs[0] (solar system, 10,000 steps in one launch) and configs[1] (disk N=4,096)
  ensemble           BASELINE configs[3]
  cpu_baseline       the oracle port on all host threads + the UNMODIFIED Python reference (baseline/_ref) on 1 core

--impl reference times the reference's CPU implementation of the path: the oracle port (oracle/nbody_oracle.c,
the reference's rounding sequence, all host threads) on a bounded sample of the same workload per step, and the
unmodified Python reference itself (tools/ref_timing.py in a subprocess) on BASELINE configs[0].
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(REPO, "orbital-physics_b200"), REPO):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "pairwise interactions/sec (fp64)"
UNIT = "interactions/s"
FLOP_PER_INTERACTION = 20.0            # BASELINE.json convention
FP64_NOMINAL_TFLOPS = 37.2             # 148 SMs x 64 lanes x 2 flop x 1.965 GHz
REF_DIR = os.path.join(REPO, "baseline", "_ref")
PORT_NOTE = ("row form: one evaluation per ORDERED interaction, i.e. up to 2x the arithmetic of the reference's "
             "half-matrix loop (core/physics.py:136-155) per ordered interaction -- the price of using every host thread")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-bodies", type=int, default=0, help="override the workload size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ensemble", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip scale_denominator / strong series / configs")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU-baseline sample budget (oracle port)")
    return ap.parse_args()


def workload(args):
    n = args.n_bodies or (262144 if args.gpus == 1 else 2097152)
    name = (f"Plummer sphere N={n} all-pairs force + leapfrog step"
            + ("" if args.gpus == 1 else f", every unordered pair once, pair blocks dealt in snake order to {args.gpus} ranks, "
                                         "all-gather 32 B x N + reduce-scatter (all-reduce below 1M bodies) 24 B x N per step (NCCL)"))
    return n, name


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.device)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(pw)))
        return out


# --------------------------------------------------------------------------- CPU arm
def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_sample(n, cloud, budget_s, nthreads=0):
    """Time the oracle port (reference algorithm, C, all host threads) on a bounded row sample."""
    from oracle import load_c_oracle
    orc = load_c_oracle()
    threads = host_threads() if nthreads <= 0 else nthreads
    rng = np.random.default_rng(1)
    probe = rng.choice(n, size=min(n, 64 * threads), replace=False).astype(np.int64)
    t0 = time.perf_counter()
    orc.pairwise_sample(cloud["x"], cloud["y"], cloud["z"], cloud["m"], cloud["eps"], cloud["G"], probe, nthreads=threads)
    rate = len(probe) * (n - 1) / (time.perf_counter() - t0)
    rows_n = int(max(threads, min(n, rate * budget_s / (n - 1))))
    rows = rng.choice(n, size=rows_n, replace=False).astype(np.int64)
    t0 = time.perf_counter()
    orc.pairwise_sample(cloud["x"], cloud["y"], cloud["z"], cloud["m"], cloud["eps"], cloud["G"], rows, nthreads=threads)
    dt = time.perf_counter() - t0
    return {"value": rows_n * (n - 1) / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{rows_n} of {n} target rows x {n - 1} sources of the same Plummer IC, "
                      f"oracle/nbody_oracle.c (reference rounding sequence), {dt:.1f} s wall; {PORT_NOTE}",
            "host_cpus": os.cpu_count(), "cpu_model": cpu_model()}


def ref_timing(*cli, timeout=180):
    """tools/ref_timing.py in a subprocess: imports ONLY the unmodified reference from baseline/_ref."""
    if not os.path.isdir(os.path.join(REF_DIR, "core")):
        return {"unavailable": "baseline/_ref/core missing (made by __graft_entry__.build() where /root/reference exists)"}
    env = {k: v for k, v in os.environ.items() if k not in ("PYTHONPATH",)}
    env.update(OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    try:
        out = subprocess.run([sys.executable, os.path.join(REPO, "tools", "ref_timing.py"), "--ref", REF_DIR, *cli],
                             capture_output=True, text=True, timeout=timeout, env=env, cwd=REPO)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as exc:
        return {"unavailable": f"reference run failed: {exc}"}


def reference_python_records(steps_small, steps_large, want_state=False):
    """The UNMODIFIED reference's SimulationEngine.step on BASELINE configs[0] (N=9, N=15; both velocity dtypes),
    1 core.  Bounded: `steps_*` steps each, not the full 10,000 (us/step is flat in the step count)."""
    recs = []
    for n, steps in ((9, steps_small), (15, steps_large)):
        for vel in ("f32", "f64"):
            r = ref_timing("--case", "solar", "--n", str(n), "--steps", str(steps), "--vel", vel)
            if not want_state:
                r.pop("pos", None); r.pop("vel", None)
            r["kind"] = "reference"
            r["workload"] = (f"core/examples.py:198-217 solar system, bodies[:{n}], dt=86400 s, eps=1e6 m, "
                             f"SimulationEngine.step x {steps} (BASELINE configs[0] asks 10,000; us/step is flat)")
            recs.append(r)
    return recs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from core import synthetic
    n, name = workload(args)
    cloud = synthetic.plummer(n)
    from oracle import load_c_oracle
    orc = load_c_oracle()
    threads = host_threads()          # torchrun exports OMP_NUM_THREADS=1: ask for the cores explicitly
    rng = np.random.default_rng(1)
    # calibrate ~3 s of CPU work per step
    probe = rng.choice(n, size=min(n, 32 * threads), replace=False).astype(np.int64)
    t0 = time.perf_counter()
    orc.pairwise_sample(cloud["x"], cloud["y"], cloud["z"], cloud["m"], cloud["eps"], cloud["G"], probe, nthreads=threads)
    rate = len(probe) * (n - 1) / (time.perf_counter() - t0)
    per_step = max(1.0, min(3.0, 120.0 / max(1, args.steps + args.warmup)))
    rows_n = int(max(threads, min(n, rate * per_step / (n - 1))))
    rows = rng.choice(n, size=rows_n, replace=False).astype(np.int64)

    def one_step():
        orc.pairwise_sample(cloud["x"], cloud["y"], cloud["z"], cloud["m"], cloud["eps"], cloud["G"], rows,
                            nthreads=threads)

    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = time.perf_counter() - t0
    value = args.steps * rows_n * (n - 1) / dt
    sample = (f"each step = {rows_n} of {n} target rows x {n - 1} sources (bounded sample of the force pass), "
              f"oracle port of core/physics.py:125-159, {threads} threads; {PORT_NOTE}")
    # the real thing: the unmodified Python reference on the one config it can run (BASELINE configs[0])
    pyref = reference_python_records(1500, 1500)
    pairs = ref_timing("--case", "pairs", "--n", "192")
    extrap = None
    if "force_us_per_pair" in pairs:
        us = pairs["force_us_per_pair"] + pairs["collisions_us_per_pair"]
        extrap = {"us_per_pair_force_plus_sweep": us, "measured_at_n": 192,
                  "seconds_per_step_at_this_n_extrapolated": us * 1e-6 * n * (n - 1) / 2,
                  "ordered_interactions_per_s": 2.0 / (us * 1e-6), "cores": 1,
                  "note": "core/physics.py:125-159 + :510-535 per unordered pair; the reference cannot run this N"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "n_bodies": n, "mode": "reference rounding sequence (CPU port, row form)",
                   "ic": "Plummer (Aarseth-Henon-Wielen), seed=N",
                   "interactions_per_step": "N^2 (extrapolated from the bounded row sample)",
                   "l2": "n/a (CPU)", "parallelism": f"{threads} host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "cpu_model": cpu_model(),
                         "reference_python": pyref, "reference_python_extrapolated": extrap},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------- ensemble side measurement
def ensemble_measure(device, torch, nsys=65536, nbody=16, steps=200, world=1, rank=0, dist=None):
    """BASELINE configs[3]: 65,536 independent 16-body systems, block-partitioned over the ranks with no
    collective on the data path.  Reports achieved algorithmic GB/s with one step per launch (state round-trips
    HBM) on both byte conventions and interactions/s with all steps fused in one launch; times are max over ranks."""
    from core import _native, synthetic
    from core.ensemble import partition
    lo, hi = partition(nsys, world, rank)
    e = synthetic.ensemble_fast(hi - lo, nbody, seed=10_000 + rank)
    ens = _native.DeviceEnsemble(hi - lo, nbody, device, _native.MODE_FAST)
    ens.set_stream(torch.cuda.current_stream().cuda_stream)
    ens.set_params(e["dt"], e["eps"], e["G"])
    ens.upload(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fused):
        ens.step(steps, fused=fused)          # warm-up on the same path (small batches take the time-sliced kernel)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record(); ens.step(steps, fused=fused); ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    ms_unfused = timed(False)
    ms_fused = timed(True)
    info = ens.info() if hasattr(ens, "info") else {}
    ens.close()
    moved = float(info.get("bytes_per_body_step", 152))
    return {"workload": f"{nsys} independent {nbody}-body systems, block-partitioned over {world} GPU(s), no collective",
            "unfused_gbs": moved * nsys * nbody * steps / (ms_unfused * 1e-3) / 1e9,
            "unfused_gbs_104B_convention": 104.0 * nsys * nbody * steps / (ms_unfused * 1e-3) / 1e9,
            "unfused_system_steps_per_s": nsys * steps / (ms_unfused * 1e-3),
            "fused_interactions_per_s": nsys * nbody * nbody * steps / (ms_fused * 1e-3),
            "bytes_per_body_step_moved": moved, "bytes_per_body_step_survey": 104, "steps": steps, "n_gpus": world,
            "kernel": info.get("kernel"),
            "note": "un-fused = one step per launch (16 launches replayed per CUDA graph); GB/s = bytes the kernel "
                    "actually reads+writes per body-step (bytes_per_body_step_moved) x bodies / time, and the same "
                    "time on SURVEY 8(d)'s 104 B convention; the state of 65,536 systems fits the 126 MB L2, so "
                    "this is not a DRAM figure -- see profiles/ for the 524,288-system (HBM-resident) capture"}


# --------------------------------------------------------------------------- parity inside the bench
def parity_rows(orc, state, acc3n, m, eps, G, rows=64, seed=7):
    """max relative acceleration error of `rows` sampled target rows vs the oracle at the SAME positions."""
    n = len(m)
    idx = np.sort(np.random.default_rng(seed).choice(n, size=min(rows, n), replace=False)).astype(np.int64)
    ref = orc.pairwise_sample(state["x"], state["y"], state["z"], m, eps, G, idx, nthreads=host_threads())
    got = np.ascontiguousarray(acc3n[:, idx].T)
    rel = np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)
    return {"rows": int(len(idx)), "max_rel": float(rel.max()), "tolerance": 1e-12,
            "oracle": "oracle/nbody_oracle.c row sums in the reference's order (core/physics.py:125-159)"}


# --------------------------------------------------------------------------- initial-condition pipeline (N=1 only)
def ic_pipeline(device, count=262144, live=512):
    """orbital elements -> Cartesian states on the device (orb_kepler_states, csrc/kepler.cu) with the
    host libm's sin / cos restated for the device, and the UNMODIFIED reference's Body.get_state (core/body.py:184-249)
    on `live` random bodies in a subprocess: the device must reproduce its states bit for bit."""
    from core import _native
    rng = np.random.default_rng(99)
    a = np.exp(rng.uniform(np.log(0.1), np.log(40.0), count)) * 1.495978707e11
    e = rng.uniform(0.0, 0.95, count)
    el = [rng.uniform(0.0, 2 * np.pi, count), e, a, a * np.sqrt(1 - e * e), np.sqrt(1.32712440018e20 / a ** 3),
          np.abs(rng.normal(0.0, 0.3, count)), rng.uniform(0.0, 2 * np.pi, count), rng.uniform(0.0, 2 * np.pi, count)]
    _native.kepler_states(*[x[:1024] for x in el], device=device)            # warm-up
    t0 = time.perf_counter()
    _native.kepler_states(*el, device=device)
    dt = time.perf_counter() - t0
    out = {"what": "orb_kepler_states: host arrays in, states out (copies and the host pow() pass included)",
           "bodies": count, "ms": 1e3 * dt, "us_per_body": 1e6 * dt / count, "trig_mode": "libm (glibc restated)"}
    ref = ref_timing("--case", "kepler", "--n", str(live))
    if "r" in ref:
        cols = [np.array(ref["elements"][k]) for k in ("M", "e", "a", "b", "n", "inc", "Omega", "omega")]
        r, v = _native.kepler_states(*cols, device=device)
        out["live_reference_check"] = {
            "bodies": live, "bit_exact_positions": bool(np.array_equal(r, np.array(ref["r"]))),
            "bit_exact_velocities": bool(np.array_equal(v, np.array(ref["v"]))),
            "reference_us_per_body": ref["us_per_body"], "speedup_vs_reference": ref["us_per_body"] / out["us_per_body"],
            "what": "unmodified reference Body.get_state (baseline/_ref, subprocess, 1 core) vs the device pipeline"}
    else:
        out["live_reference_check"] = ref
    return out


# --------------------------------------------------------------------------- small configs (N=1 only)
def small_configs(torch, device):
    """BASELINE configs[0] and configs[1] through the public engine API, with their parity figure."""
    from core import _native, synthetic
    from core.engine import SimulationEngine
    from core.examples import solar_system_objects
    from core.physics import ObjectCollection
    from oracle import load_c_oracle
    from oracle.c_oracle import State
    orc = load_c_oracle()
    out = {}
    c0 = []
    for n in (9, 15):
        for vel in ("f32", "f64"):
            bodies, _ = solar_system_objects(moons=False)
            bodies = bodies[:n]
            if vel == "f64":
                for o in bodies:
                    o.velocity = np.asarray(o.velocity, dtype=np.float64).copy()
            arr = lambda f: np.array([f(o) for o in bodies], dtype=np.float64)
            f32 = np.array([o.velocity.dtype == np.float32 for o in bodies], dtype=np.uint8)
            st = State(orc, arr(lambda o: o.coordinates.x), arr(lambda o: o.coordinates.y),
                       arr(lambda o: o.coordinates.z), arr(lambda o: o.velocity[0]), arr(lambda o: o.velocity[1]),
                       arr(lambda o: o.velocity[2]), arr(lambda o: o.mass), arr(lambda o: o.radius), f32, 86400.0, 1e6)
            eng = SimulationEngine(ObjectCollection(bodies), dt=86400.0, softening=1e6, cache=False, max_hist=None,
                                   device=device)
            eng.run(100); eng.synchronize()                       # warm-up (module load, first launch)
            t0 = time.perf_counter()
            eng.run(10_000); eng.synchronize()
            dt = time.perf_counter() - t0
            t0 = time.perf_counter()
            st.step(10_100)
            dt_port = time.perf_counter() - t0
            loop_us = None
            if n == 15 and vel == "f32":
                # the reference's own driver pattern (core/examples.py:198-217): one engine.step() call per step;
                # the engine defers the calls and runs the stretch when the state is observed (the read below)
                t0 = time.perf_counter()
                for _ in range(10_000):
                    eng.step()
                _ = eng.objects.objects[1].coordinates.x
                loop_us = 1e6 * (time.perf_counter() - t0) / 10_000
                st.step(10_000)
            pos = np.array([o.position() for o in eng.objects])
            v = np.array([np.asarray(o.velocity, dtype=np.float64) for o in eng.objects])
            c0.append({"n_bodies": n, "velocity_dtype": "float32" if vel == "f32" else "float64", "steps": 10_000,
                       "kernel": eng.kernel_info()["name"], "ms_total": 1e3 * dt, "us_per_step": 1e2 * dt,
                       "interactions_per_s": n * (n - 1) * 1e4 / dt,
                       "bit_exact_vs_oracle_after_10100_steps": bool(np.array_equal(pos, st.pos) and np.array_equal(v, st.vel)),
                       "oracle_port_1_core_us_per_step": 1e6 * dt_port / 10_100})
            if loop_us is not None:
                c0[-1]["step_call_loop_us_per_step"] = loop_us
                c0[-1]["steps_compared_with_oracle"] = 20_100
                c0[-1]["note"] = ("this run: 10,100 run() steps + 10,000 single step() calls (deferred, one device "
                                  "stretch at the read), then compared with the oracle")
            eng.close()
    out["C0_solar_system"] = {"what": "core/examples.py solar system via SimulationEngine.run(10000): one launch, "
                                      "bit-exact mode (the default below 4,096 bodies)", "runs": c0}
    # live check against the real reference, where its checkout travelled with the snapshot
    live = ref_timing("--case", "solar", "--n", "15", "--steps", "1000", "--vel", "f32")
    if "pos" in live:
        bodies, _ = solar_system_objects(moons=False)
        eng = SimulationEngine(ObjectCollection(bodies), dt=86400.0, softening=1e6, cache=False, max_hist=None,
                               device=device)
        eng.run(1000)
        pos = np.array([o.position() for o in eng.objects])
        v = np.array([np.asarray(o.velocity, dtype=np.float64) for o in eng.objects])
        out["C0_solar_system"]["live_reference_check"] = {
            "steps": 1000, "n_bodies": 15, "bit_exact_positions": bool(np.array_equal(pos, np.array(live["pos"]))),
            "bit_exact_velocities": bool(np.array_equal(v, np.array(live["vel"]))),
            "reference_us_per_step": live["us_per_step"], "speedup_vs_reference": live["us_per_step"] / c0[2]["us_per_step"],
            "what": "unmodified reference (baseline/_ref, subprocess) vs this engine, same 1000 steps"}
        eng.close()
    else:
        out["C0_solar_system"]["live_reference_check"] = live

    c = synthetic.uniform_disk(4096)
    vel = [v.astype(np.float32).astype(np.float64) for v in (c["vx"], c["vy"], c["vz"])]
    g = np.load(os.path.join(REPO, "tests", "golden", "disk4096_f32.npz"))
    c1 = []
    for mode, label in ((_native.MODE_FAITHFUL, "faithful"), (_native.MODE_FAST, "fast")):
        dev = _native.DeviceSystem(4096, device, mode)
        dev.set_params(c["dt"], c["eps"], c["G"])
        dev.upload(c["x"], c["y"], c["z"], *vel, c["m"], c["radius"], np.ones(4096, np.uint8))
        dev.accel()
        dev.step(2)
        a2 = dev.download_acc().T
        if label == "faithful":
            s2 = dev.download_state()
            parity = {"bit_exact_vs_reference_fixture_step_2": bool(
                np.array_equal(a2, g["acc_2"]) and np.array_equal(np.stack([s2["x"], s2["y"], s2["z"]], 1), g["pos_2"]))}
        else:
            s2 = dev.download_state()
            ref, _ = orc.pairwise(s2["x"], s2["y"], s2["z"], c["m"], c["eps"], c["G"], nthreads=host_threads())
            rel = np.linalg.norm(a2 - ref, axis=1) / np.linalg.norm(ref, axis=1)
            parity = {"max_rel_acc_error_all_rows_vs_oracle": float(rel.max()), "tolerance": 1e-12}
        dev.step(50); dev.synchronize()
        t0 = time.perf_counter()
        dev.step(1000); dev.synchronize()
        dt = time.perf_counter() - t0
        c1.append({"mode": label, "kernel": dev.force_kernel_info()["name"], "steps": 1000, "ms_per_step": dt,
                   "interactions_per_s": 4096.0 * 4096.0 * 1000 / dt,
                   "frac_of_fp64_nominal_20flop": FLOP_PER_INTERACTION * 4096.0 * 4096.0 * 1000 / dt / 1e12 / FP64_NOMINAL_TFLOPS,
                   **parity})
        dev.close()
    out["C1_disk_4096"] = {"what": "uniform disk N=4,096 around a central mass (SURVEY 8d C1), orb_step x 1000 under a "
                                   "CUDA graph; the reference's own ctor + 2 steps are the fixture "
                                   "tests/golden/disk4096_f32.npz (~105 s per step on one core)", "runs": c1}
    return out


# --------------------------------------------------------------------------- GPU arm
def timed_sharded(torch, dist, sh, steps, warmup, flush, world):
    """`steps` timed ShardedSystem.step(1) calls after `warmup`; -> (ms_total, force_ms, comm dict), max over ranks."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        flush.zero_()
        sh.step(1)
    barrier()
    sh.events = {}
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        flush.zero_()
        sh.step(1)
    b.record()
    barrier()
    ev, sh.events = sh.events, None
    mean = lambda k: float(np.mean([x.elapsed_time(y) for x, y in ev[k]])) if ev.get(k) else 0.0
    vals = [a.elapsed_time(b), mean("force"), mean("gather"), mean("reduce")]
    if world > 1:
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = [float(v) for v in t]
    return vals[0], vals[1], {"all_gather_pos4": vals[2], "reduce_acc": vals[3]}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from core import _native, synthetic
    from core.distributed import DistComm, ShardedSystem, slab

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or _native.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, name = workload(args)
    cloud = synthetic.plummer(n)
    stream = torch.cuda.current_stream().cuda_stream
    lo, hi = slab(n, world, rank)
    warmup = max(3, args.warmup)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    comm_ms = None
    if world == 1:
        dev = _native.DeviceSystem(n, local, _native.MODE_FAST)
        dev.set_stream(stream)
        dev.set_params(cloud["dt"], cloud["eps"], cloud["G"])
        dev.upload(*cloud.arrays())
        dev.accel()
        force_ev = []

        def one_step(timed):
            flush.zero_()                                   # L2 flush between iterations
            dev.step_begin()
            if timed:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); dev.accel(); b.record()
                force_ev.append((a, b))
            else:
                dev.accel()
            dev.step_kick()

        for _ in range(warmup):
            one_step(False)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        launches0 = dev.launch_count()
        t_ev0, t_ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_ev0.record()
        for _ in range(args.steps):
            one_step(True)
        t_ev1.record()
        barrier()
        clocks = sampler.stop()
        launches = dev.launch_count() - launches0        # our kernels only (torch's L2-flush fill is not counted)
        ms_total = t_ev0.elapsed_time(t_ev1)
        force_ms = float(np.mean([a.elapsed_time(b) for a, b in force_ev]))
        sysm = dev
        e2e_body = lambda: (dev.step_begin(), dev.accel(), dev.step_kick())
    else:
        comm = DistComm(device=local)
        sh = ShardedSystem.from_arrays(*cloud.arrays(), cloud["dt"], cloud["eps"], cloud["G"], mode=_native.MODE_FAST,
                                       comm=comm)
        dev = sh.dev
        sampler = ClockSampler(local)
        for _ in range(warmup):
            flush.zero_(); sh.step(1)
        barrier()
        if rank == 0:
            sampler.start()
        launches0 = sh.launch_count()
        ms_total, force_ms, comm_ms = timed_sharded(torch, dist, sh, args.steps, 0, flush, world)
        clocks = sampler.stop() if rank == 0 else {}
        launches = sh.launch_count() - launches0
        sysm = sh
        e2e_body = lambda: sh.step(1)
    info = sysm.force_kernel_info()
    value = float(n) * n * args.steps / (ms_total * 1e-3)

    # ---- parity of what was just timed: sampled rows of the final accelerations vs the oracle (all ranks take part
    # in the gathers; rank 0 evaluates the oracle)
    state = sysm.download_state()
    acc = sysm.download_acc()
    parity = None
    if rank == 0:
        from oracle import load_c_oracle
        parity = parity_rows(load_c_oracle(), state, acc, cloud["m"], cloud["eps"], cloud["G"])

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    pin = _native.PinnedBuffer(8 * n + 6 * n)
    hin = pin.array[: 8 * n].reshape(8, n)
    hout = pin.array[8 * n:].reshape(6, n)
    for k, a in enumerate(cloud.arrays()):
        hin[k] = a
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        sysm.upload(*[hin[k] for k in range(8)])          # H2D of this step's inputs
        e2e_body()
        sysm.download_state(hout)                          # D2H of the step's result (synchronises)

    sysm.accel()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e = {"value": float(n) * n * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * 8 * n + n,
           "d2h_bytes_per_step": 6 * 8 * n, "steps": e2e_steps,
           "path": "orb_upload (pinned host SoA) -> one leapfrog step -> orb_download_state"
                   + ("" if world == 1 else " (per rank; the step includes both collectives and the velocity gather)")}

    # ---- N>1: the N=262,144 workload of the 1-GPU line spread over the ranks (strong scaling where comm shows)
    strong = None
    if world > 1 and not args.no_extras:
        sysm.close()
        n2 = 262144
        c2 = synthetic.plummer(n2)
        sh2 = ShardedSystem.from_arrays(*c2.arrays(), c2["dt"], c2["eps"], c2["G"], mode=_native.MODE_FAST,
                                        comm=DistComm(device=local))
        ms2, f2, comm2 = timed_sharded(torch, dist, sh2, 20, 3, flush, world)
        st2, a2 = sh2.download_state(), sh2.download_acc()
        strong = {"n_bodies": n2, "steps": 20, "value": float(n2) * n2 * 20 / (ms2 * 1e-3), "unit": UNIT,
                  "ms_per_step": ms2 / 20, "force_ms": f2, "comm_ms": comm2,
                  "note": "per step: L2 flush (256 MiB memset, ~0.1 ms) + Python-driven launch sequence + 2 NCCL "
                          "collectives; compare with the --gpus 1 line (same N, same kernel)"}
        if rank == 0:
            from oracle import load_c_oracle
            strong["parity_check"] = parity_rows(load_c_oracle(), st2, a2, c2["m"], c2["eps"], c2["G"])
        sh2.close()

    ens_result = None
    if not args.no_ensemble:
        try:
            ens_result = ensemble_measure(local, torch, world=world, rank=rank, dist=dist)
        except Exception as exc:      # never lose the headline line over the side measurement
            if world > 1:
                raise
            ens_result = {"error": str(exc)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (force pass: this rank's share of the pair blocks)
    peak = _native.fp64_peak(local, 1.0)
    local_interactions = float(n) * n / world      # every rank evaluates 1/world of the N^2 ordered interactions
    achieved_tf = FLOP_PER_INTERACTION * local_interactions / (force_ms * 1e-3) / 1e12
    # pair-symmetric kernel: 20 FP64 instructions per unordered pair = 10 per ordered interaction; one-sided: 16
    # (9 when all masses are equal: the two per-pair mass multiplies are factored out of the sum)
    uniform = info["name"].endswith(",true>")
    fp64_per_int = (9 if uniform else 10) if "force_sym" in info["name"] else 16
    general = None
    if uniform and world == 1:
        # the same workload through the general-mass variant of the kernel, for comparison
        os.environ["ORBITAL_B200_SYM_UNI"] = "0"
        try:
            dev.accel(); torch.cuda.synchronize()
            evs = []
            for _ in range(3):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); dev.accel(); b.record(); evs.append((a, b))
            torch.cuda.synchronize()
            gms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
            gtf = FLOP_PER_INTERACTION * local_interactions / (gms * 1e-3) / 1e12
            from oracle import load_c_oracle
            gpar = parity_rows(load_c_oracle(), dev.download_state(), dev.download_acc(), cloud["m"], cloud["eps"],
                               cloud["G"])
            general = {"kernel": dev.force_kernel_info()["name"], "kernel_ms": gms, "achieved": gtf,
                       "frac": gtf / peak["tflops_mean"], "frac_of_nominal": gtf / FP64_NOMINAL_TFLOPS,
                       "fp64_instr_per_interaction": 10, "interactions_per_s": local_interactions / (gms * 1e-3),
                       "parity_check": gpar}
        finally:
            del os.environ["ORBITAL_B200_SYM_UNI"]
    # DRAM traffic per launch of the dominant kernel: NOT measured in this run -- taken from the committed
    # `ncu --set full` capture of this exact workload (dram__bytes_read + dram__bytes_write); null otherwise
    traffic, traffic_src = None, None
    if n == 262144 and world == 1:
        traffic, traffic_src = TRAFFIC_FROM_CAPTURE.get(info["name"], (None, None))
    roofline = {
        "bound": "fp64", "achieved": achieved_tf, "peak": peak["tflops_mean"], "unit": "TFLOP/s",
        "frac": achieved_tf / peak["tflops_mean"], "traffic": traffic, "traffic_source": traffic_src,
        "kernel": info["name"], "grid": info["grid"], "block": info["block"],
        "kernel_ms": force_ms, "interactions_per_launch": local_interactions,
        "flop_per_interaction": FLOP_PER_INTERACTION,
        "peak_source": "measured: orb_fp64_peak, best DFMA chain shape on this GPU (two-register form r=fma(r,a,r); "
                       "the pipe is register-read limited, the usual fma(r,a,b) chain stops at ~91 %), mean over 1 s "
                       f"(burst {peak['tflops_best']:.2f} TF at {peak['sm_clock_mhz']:.0f} MHz); "
                       "MEASURED_PEAKS.json has no FP64 figure",
        "peak_nominal": FP64_NOMINAL_TFLOPS, "frac_of_nominal": achieved_tf / FP64_NOMINAL_TFLOPS,
        "algorithmic_hbm_bytes_per_launch": 56 * n,
        "fp64_instr_per_interaction": fp64_per_int,
        "fp64_pipe_util_est": achieved_tf / FLOP_PER_INTERACTION * fp64_per_int * 2 / peak["tflops_mean"],
        "note": "frac uses the 20-flop convention; the kernel issues fp64_instr_per_interaction x 2 flop per ordered "
                "interaction (each unordered pair once; uniform masses factor the mass multiply out), so "
                "fp64_pipe_util_est is the share of FP64 issue slots actually used"
                + (" -- general_mass_variant is the same run without the equal-mass specialisation" if general else ""),
    }
    if "force_sym" in info["name"] and world == 1:
        # what actually bounds this instruction mix on B200: the FP64 pipe takes max(2, distinct 64-bit register
        # operands) cycles per warp instruction (tools/dfma_probe*.cu, profiles/r1_dfma_probe.txt); per unordered
        # pair the kernel issues 12 (14) two-operand instructions and 6 three-operand accumulations
        mhz = clocks.get("sm_mhz") or peak["sm_clock_mhz"]
        model = (12 if uniform else 14) * 2 + 6 * 3
        sm_count = _native.device_info(local)["sm_count"]
        measured = force_ms * 1e-3 * mhz * 1e6 * sm_count * 4 / (local_interactions / 2 / 32)
        roofline["register_file_bound"] = {"model_cycles_per_warp_pair": model, "measured_cycles_per_warp_pair": measured,
                                           "frac": model / measured, "sm_mhz": mhz,
                                           "note": "no operand reuse assumed; with perfect .reuse on the six "
                                                   "accumulations the floor would be 37 (41) cycles"}
    if general:
        roofline["general_mass_variant"] = general
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "n_bodies": n, "mode": "fast", "ic": "Plummer (Aarseth-Henon-Wielen), seed=N",
                   "interactions_per_step": "N^2", "l2": "flushed between steps (256 MiB memset)",
                   "parallelism": "1 GPU" if world == 1 else
                                  f"snake-order pair-block ownership x{world}, all-gather 32 B x N (NCCL) + reduction of the partial accelerations out of peer memory (CUDA IPC over NVLink; NCCL fallback)"},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "parity_check": parity,
    }
    if comm_ms is not None:
        line["comm_ms"] = comm_ms
        line["force_ms"] = force_ms
    if strong is not None:
        line["strong_262144"] = strong
    if world == 1 and not args.no_extras:
        try:
            line["scale_denominator"] = scale_denominator(torch, _native, synthetic, local, flush)
        except Exception as exc:
            line["scale_denominator"] = {"error": str(exc)}
        try:
            line["configs"] = small_configs(torch, local)
        except Exception as exc:
            line["configs"] = {"error": repr(exc)}
        try:
            line["ic_pipeline"] = ic_pipeline(local)
        except Exception as exc:
            line["ic_pipeline"] = {"error": repr(exc)}
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_sample(n, cloud, args.cpu_seconds)
        cb["note"] = ("C port of the reference algorithm on all host threads (kind: port). The reference itself is "
                      "single-threaded Python: reference_python holds the UNMODIFIED reference timed in this run on "
                      "BASELINE configs[0] (the only config it can run in bounded time)")
        cb["reference_python"] = reference_python_records(600, 400)
        line["cpu_baseline"] = cb
    if ens_result is not None:
        line["ensemble"] = ens_result
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu captures of N=262,144 on one GPU
TRAFFIC_FROM_CAPTURE = {
    "force_sym_kernel<8,false,true>": (1010531072.0, "profiles/r2_force_sym_ncu.txt (23.97 MB read + 986.56 MB written: the "
                                                     "P_j partial planes, 69x the 14.7 MB algorithmic; 0.3 % of DRAM bandwidth)"),
    "force_sym_kernel<8,false,false>": (1011329024.0, "profiles/r1_force_sym_ti8_ncu.txt (23.45 MB read + 987.88 MB written)"),
}


def scale_denominator(torch, _native, synthetic, device, flush, n=2097152, steps=2):
    """One GPU on the multi-GPU workload (BASELINE configs[4]): the denominator of the 2/4/8-GPU efficiency."""
    c = synthetic.plummer(n)
    dev = _native.DeviceSystem(n, device, _native.MODE_FAST)
    dev.set_stream(torch.cuda.current_stream().cuda_stream)
    dev.set_params(c["dt"], c["eps"], c["G"])
    dev.upload(*c.arrays())
    dev.accel()
    step = lambda: (flush.zero_(), dev.step_begin(), dev.accel(), dev.step_kick())
    step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    from oracle import load_c_oracle
    par = parity_rows(load_c_oracle(), dev.download_state(), dev.download_acc(), c["m"], c["eps"], c["G"])
    name = dev.force_kernel_info()["name"]
    dev.close()
    return {"n_bodies": n, "n_gpus": 1, "steps": steps, "warmup": 2, "value": float(n) * n * steps / (ms * 1e-3),
            "unit": UNIT, "ms_per_step": ms / steps, "kernel": name, "parity_check": par,
            "what": "Plummer N=2,097,152 on ONE GPU (constructor pass + 1 step as warm-up, then timed steps): divide "
                    "the --gpus 2/4/8 `value` by (n_gpus x this value) for strong-scaling efficiency"}


_RESULT_FD = None


def emit(line: dict):
    """The one JSON line goes to the real stdout; everything else any library prints (e.g. NCCL's version banner,
    which goes to fd 1) was diverted to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    args = parse_args()
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)                      # stray prints of native libraries -> stderr
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
