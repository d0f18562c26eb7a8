#!/usr/bin/env python
"""Benchmark of the hot path: all-pairs gravity + leapfrog step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n-bodies N]

metric  : pairwise interactions/s (fp64), N^2 ordered interactions per force evaluation,
          one force evaluation per leapfrog step.
workload: N=1 GPU  -> BASELINE configs[2]: Plummer sphere N=262,144, fast (roofline) kernel.
          N>1 GPUs -> BASELINE configs[4]: Plummer sphere N=2,097,152, targets partitioned by rank,
                      NCCL all-gather of the packed positions every step (strong scaling).
A "step" is one pass of the hot path: half-kick+drift -> force -> half-kick, through the C ABI
(orb_step_begin / orb_accel / orb_step_kick -- the same kernels orb_step launches, split so the
force pass can be bracketed by CUDA events on the launching stream).

--impl reference times the reference's own CPU algorithm (the oracle port, oracle/nbody_oracle.c,
all host threads) on the same workload/metric with a bounded sample per step.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-bodies", type=int, default=0, help="override the workload size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ensemble", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample budget")
    return ap.parse_args()


def workload(args):
    n = args.n_bodies or (262144 if args.gpus == 1 else 2097152)
    name = (f"Plummer sphere N={n} all-pairs force + leapfrog step"
            + ("" if args.gpus == 1 else f", targets partitioned over {args.gpus} GPUs, NCCL all-gather of positions"))
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
                      f"oracle/nbody_oracle.c row form (reference rounding sequence), {dt:.1f} s wall",
            "host_cpus": os.cpu_count()}, rows_n, dt


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
    per_step = max(1.0, min(3.0, 150.0 / max(1, args.steps + args.warmup)))
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
              f"oracle port of core/physics.py:125-159, {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "n_bodies": n, "mode": "reference rounding sequence (CPU port, row form)",
                   "ic": "Plummer (Aarseth-Henon-Wielen), seed=N",
                   "interactions_per_step": "N^2 (extrapolated from the bounded row sample)",
                   "l2": "n/a (CPU)", "parallelism": f"{threads} host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------- ensemble side measurement
def ensemble_measure(device, torch, nsys=65536, nbody=16, steps=200, world=1, rank=0, dist=None):
    """BASELINE configs[3]: 65,536 independent 16-body systems, block-partitioned over the ranks with no
    collective on the data path.  Reports achieved algorithmic GB/s with one step per launch (state round-trips
    HBM: 152 B per body-step) and interactions/s with all steps fused in one launch; times are max over ranks."""
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
        ens.step(20, fused=fused)
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
    ens.close()
    bytes_step = 152.0 * nsys * nbody
    return {"workload": f"{nsys} independent {nbody}-body systems, {64 // nbody} systems per warp (8 lanes each), "
                        f"block-partitioned over {world} GPU(s), no collective",
            "unfused_gbs": bytes_step * steps / (ms_unfused * 1e-3) / 1e9,
            "unfused_system_steps_per_s": nsys * steps / (ms_unfused * 1e-3),
            "fused_interactions_per_s": nsys * nbody * nbody * steps / (ms_fused * 1e-3),
            "bytes_per_body_step": 152, "steps": steps, "n_gpus": world,
            "note": "un-fused = one step per launch (16 launches replayed per CUDA graph), algorithmic bytes / time; "
                    "the 84 MB state of 65,536 systems fits the 126 MB L2, so this can exceed the DRAM peak -- at "
                    "524,288 systems (671 MB) the same kernel sustains 6,332 GB/s = 98 % of the measured HBM copy "
                    "peak (profiles/r1_ens_sizes.txt, r1_ensemble_fast_hbm_ncu.txt)"}


# --------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from core import _native, synthetic
    from core.distributed import ShardedSystem, slab

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
    if n % world:
        raise SystemExit(f"n={n} not divisible by world size {world}")
    cloud = synthetic.plummer(n)
    stream = torch.cuda.current_stream().cuda_stream
    lo, hi = slab(n, world, rank)

    if world == 1:
        dev = _native.DeviceSystem(n, local, _native.MODE_FAST)
        dev.set_stream(stream)
        dev.set_params(cloud["dt"], cloud["eps"], cloud["G"])
        dev.upload(*cloud.arrays())
        dev.accel()
        gather = lambda: None
        reduce_acc = lambda: None
    else:
        reduce_acc = lambda: None
        sh = ShardedSystem(*cloud.arrays(), cloud["dt"], cloud["eps"], cloud["G"], mode=_native.MODE_FAST, device=local)
        dev = sh.dev
        gather = sh._all_gather_positions
        if sh._partial:
            reduce_acc = lambda: dist.all_reduce(sh._acc)
    info = dev.force_kernel_info()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    force_ev = []

    def one_step(timed):
        flush.zero_()                                   # L2 flush between iterations
        dev.step_begin()
        gather()
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); dev.accel(); b.record()
            force_ev.append((a, b))
        else:
            dev.accel()
        reduce_acc()
        dev.step_kick()

    for _ in range(max(3, args.warmup)):
        one_step(False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = dev.launch_count()
    t_ev0, t_ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_ev0.record()
    for _ in range(args.steps):
        one_step(True)
    t_ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    launches = dev.launch_count() - launches0        # our kernels only (torch's L2-flush fill is not counted)
    ms_total = t_ev0.elapsed_time(t_ev1)
    force_ms = float(np.mean([a.elapsed_time(b) for a, b in force_ev]))
    if world > 1:
        t = torch.tensor([ms_total, force_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, force_ms = float(t[0]), float(t[1])
    value = float(n) * n * args.steps / (ms_total * 1e-3)

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    pin = _native.PinnedBuffer(8 * n + 6 * n)
    hin = pin.array[: 8 * n].reshape(8, n)
    hout = pin.array[8 * n:].reshape(6, n)
    for k, a in enumerate(cloud.arrays()):
        hin[k] = a
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        dev.upload(*[hin[k] for k in range(8)])          # H2D of this step's inputs
        dev.step_begin(); gather(); dev.accel(); reduce_acc(); dev.step_kick()
        dev.download_state(hout)                           # D2H of the step's result (synchronises)

    dev.accel()
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
           "path": "orb_upload (pinned host SoA) -> orb_step_begin/orb_accel/orb_step_kick -> orb_download_state"}

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

    # ---- roofline of the dominant kernel (force pass of the local targets)
    peak = _native.fp64_peak(local, 1.0)
    local_interactions = float(hi - lo) * n        # this rank's share of the N^2 ordered interactions
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
            general = {"kernel": dev.force_kernel_info()["name"], "kernel_ms": gms, "achieved": gtf,
                       "frac": gtf / peak["tflops_mean"], "fp64_instr_per_interaction": 10,
                       "interactions_per_s": local_interactions / (gms * 1e-3)}
        finally:
            del os.environ["ORBITAL_B200_SYM_UNI"]
    # DRAM traffic per launch of the dominant kernel from the committed `ncu --set full` captures of this exact
    # workload (dram__bytes_read + dram__bytes_write, almost all of it the partial planes P_j):
    #   profiles/r1_force_sym_ti8_uniform_ncu.txt  21.89 MB + 985.64 MB   (uniform-mass variant)
    #   profiles/r1_force_sym_ti8_ncu.txt          23.45 MB + 987.88 MB   (general variant)
    # null for any other size / kernel
    traffic = None
    if n == 262144 and world == 1:
        traffic = {"force_sym_kernel<8,false,true>": 1007530752.0,
                   "force_sym_kernel<8,false,false>": 1011329024.0}.get(info["name"])
    roofline = {
        "bound": "fp64", "achieved": achieved_tf, "peak": peak["tflops_mean"], "unit": "TFLOP/s",
        "frac": achieved_tf / peak["tflops_mean"], "traffic": traffic,
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
        "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "n_bodies": n, "mode": "fast", "ic": "Plummer (Aarseth-Henon-Wielen), seed=N",
                   "interactions_per_step": "N^2", "l2": "flushed between steps (256 MiB memset)",
                   "parallelism": "1 GPU" if world == 1 else f"target-partition x{world} + all-gather"},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        cb, _, _ = cpu_sample(n, cloud, args.cpu_seconds)
        cb["note"] = ("C port of the reference algorithm on all host threads; the reference itself is single-threaded "
                      "Python at ~5-7 us per pair (BASELINE.md section 2)")
        line["cpu_baseline"] = cb
    if ens_result is not None:
        line["ensemble"] = ens_result
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
