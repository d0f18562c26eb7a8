"""The multi-GPU (sharded) CUDA path, held to the oracle on ONE device  (`-m gpu`).

`LocalComm([0] * world)` creates `world` handles with orb_create_ranked on device 0 and runs the real per-rank
launch sequence of a `world`-GPU job -- orb_step_begin / orb_step_force (snake-order pair-block ownership in
force_sym_kernel + reduce_sym_kernel, or the target-slab kernels) / orb_step_kick / orb_step_end -- with the
all-gather and all-reduce replaced by device-to-device copies and a fixed-order sum (core/distributed.py).
Only the transport differs from an NCCL run, so these tests are the parity evidence for BASELINE configs[4]
on a one-GPU box; tests/test_distributed.py::test_multi_rank_nccl_matches_single_gpu repeats them over NCCL
where >= 2 GPUs exist.

Bars: bit-exact mode == oracle bit for bit (reference core/physics.py:125-159, core/engine.py:65-97 incl. the
contact sweep of :85); fast mode <= 1e-12 relative acceleration error on ALL rows, every step.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = 6.67430e-11
TOL_FAST = 1e-12          # BASELINE.json: relative acceleration error <= 1e-12 per step


def bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def assert_bits(a, b, what=""):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
    if not same.all():
        rel = np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
        raise AssertionError(f"{what}: {np.count_nonzero(~same)}/{same.size} differ, max rel {rel:.3e}")


def relerr(a, ref):
    return np.linalg.norm(a - ref, axis=1) / np.linalg.norm(ref, axis=1)


@pytest.fixture(scope="module")
def nat():
    from core import _native
    assert _native.device_count() > 0
    return _native


def cloud(n, masses):
    """Plummer sphere; masses='uniform' keeps the equal masses (-> the mass-free variant of the pair-symmetric
    kernel when n is whole I-blocks), 'general' draws them log-uniformly over two decades."""
    from core import synthetic
    c = synthetic.plummer(n, seed=n)
    m = c["m"].copy()
    if masses == "general":
        m *= np.exp(np.random.default_rng(n).uniform(np.log(0.1), np.log(10.0), n))
    f32 = (np.arange(n) % 3 == 0).astype(np.uint8)
    vel = [np.where(f32 == 1, v.astype(np.float32).astype(np.float64), v) for v in (c["vx"], c["vy"], c["vz"])]
    return c, m, f32, vel


def make_sharded(nat, world, c, m, f32, vel, mode, radius=None, restitution=1.0):
    from core.distributed import LocalComm, ShardedSystem
    return ShardedSystem.from_arrays(c["x"], c["y"], c["z"], *vel, m, c["radius"] if radius is None else radius,
                                     c["dt"], c["eps"], G, mode=mode, comm=LocalComm([0] * world), vel_is_f32=f32,
                                     restitution=restitution)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("n", [4096, 6000])
def test_sharded_faithful_bit_exact_vs_oracle(nat, orc, world, n):
    """Target-slab force_faithful_kernel + kick_drift / kick on [lo, hi): bit-exact with the oracle over 3 steps,
    mixed velocity dtypes (6000 bodies: slabs that are not whole warps of targets or whole tiles)."""
    from oracle.c_oracle import State
    c, m, f32, vel = cloud(n, "general")
    st = State(orc, c["x"], c["y"], c["z"], *vel, m, c["radius"], f32, c["dt"], c["eps"], G)
    sh = make_sharded(nat, world, c, m, f32, vel, nat.MODE_FAITHFUL)
    assert sh.world == world and len(sh.devs) == world and not sh.acc_needs_allreduce()
    assert_bits(sh.download_acc().T, st.acc, "constructor force pass")
    for step in range(1, 4):
        sh.step(1)
        st.step(1, collisions=False, nthreads=8)
        s = sh.download_state()
        assert_bits(np.stack([s["x"], s["y"], s["z"]], 1), st.pos, f"pos @ {step}")
        assert_bits(np.stack([s["vx"], s["vy"], s["vz"]], 1), st.vel, f"vel @ {step}")
        assert_bits(sh.download_acc().T, st.acc, f"acc @ {step}")
    sh.close()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("n,masses", [(4096, "general"), (6000, "general"), (12288, "uniform"), (4099, "general")])
def test_sharded_fast_all_rows_within_tolerance(nat, orc, world, n, masses):
    """Pair-symmetric kernel with snake-order I-block ownership per rank (plan_sym rank/world, reduce_sym_kernel) and
    the all-reduce of the partial accelerations: every row <= 1e-12 of the oracle at the SAME positions, each of
    3 steps; n=6000 / 4099 have ragged tiles, I-blocks and slabs, n=12288 takes the uniform-mass variant."""
    c, m, f32, vel = cloud(n, masses)
    sh = make_sharded(nat, world, c, m, f32, vel, nat.MODE_FAST)
    assert sh.acc_needs_allreduce(), "fast sharded engines must run the pair-symmetric kernel"
    name = sh.force_kernel_info()["name"]
    assert name.startswith("force_sym_kernel") and name.endswith(",true>" if masses == "uniform" else ",false>"), name
    worst = 0.0
    for step in range(0, 4):
        if step:
            sh.step(1)
        s = sh.download_state()
        ref, _ = orc.pairwise(s["x"], s["y"], s["z"], m, c["eps"], G, nthreads=16)
        err = relerr(sh.download_acc().T, ref)
        worst = max(worst, float(err.max()))
        assert err.max() <= TOL_FAST, f"step {step}: max rel acc error {err.max():.3e} on row {err.argmax()}"
    # the trajectory follows the oracle's (same integrator arithmetic, accelerations within 1e-12)
    from oracle.c_oracle import State
    st = State(orc, c["x"], c["y"], c["z"], *vel, m, c["radius"], f32, c["dt"], c["eps"], G)
    st.step(3, collisions=False, nthreads=16)
    s = sh.download_state()
    P = np.stack([s["x"], s["y"], s["z"]], 1)
    assert (np.linalg.norm(P - st.pos, axis=1) / np.linalg.norm(st.pos, axis=1)).max() <= 1e-12
    sh.close()


def test_sharded_fast_equals_single_gpu_run(nat):
    """The same pair blocks summed on 1 or 4 ranks differ by rounding only; the one-sided slab kernel
    (ORBITAL_B200_SYM=0: force_fast_kernel on [lo, hi)) agrees too."""
    c, m, f32, vel = cloud(8192, "general")
    one = nat.DeviceSystem(8192, 0, nat.MODE_FAST)
    one.set_params(c["dt"], c["eps"], G)
    one.upload(c["x"], c["y"], c["z"], *vel, m, c["radius"], f32)
    one.accel()
    for _ in range(2):
        one.step_begin(); one.accel(); one.step_kick()
    a1 = one.download_acc().T
    sh = make_sharded(nat, 4, c, m, f32, vel, nat.MODE_FAST)
    sh.step(2)
    assert relerr(sh.download_acc().T, a1).max() <= 1e-12
    sh.close()
    os.environ["ORBITAL_B200_SYM"] = "0"
    try:
        sl = make_sharded(nat, 4, c, m, f32, vel, nat.MODE_FAST)
        assert not sl.acc_needs_allreduce() and sl.force_kernel_info()["name"].startswith("force_fast_kernel")
        sl.step(2)
        assert relerr(sl.download_acc().T, a1).max() <= 1e-12
        sl.close()
    finally:
        del os.environ["ORBITAL_B200_SYM"]
    one.close()


def contact_scene(n, seed):
    rng = np.random.default_rng(seed)
    box = 3e5 * (n / 700.0) ** (1 / 3)
    x, y, z = (rng.uniform(-box, box, n) for _ in range(3))
    v = rng.standard_normal((3, n)) * 400
    m = np.exp(rng.uniform(np.log(1e14), np.log(1e16), n))
    radius = rng.uniform(2e3, 8e3, n)
    f32 = (np.arange(n) % 2).astype(np.uint8)
    vel = [np.where(f32 == 1, a.astype(np.float32).astype(np.float64), a) for a in v]
    return x, y, z, vel, m, radius, f32


@pytest.mark.parametrize("world,n", [(2, 700), (3, 700), (8, 1500)])
def test_sharded_contacts_bit_exact_vs_oracle_sweep(nat, orc, world, n):
    """engine.py:85 on a sharded engine: per-rank overlap flags from the force pass, merged pair list, replicated
    sequential sweep (physics.py:510-535, 391-422) -- bit-exact with the oracle incl. U and the history ring."""
    from core.distributed import LocalComm, ShardedSystem
    from oracle.c_oracle import State
    x, y, z, vel, m, radius, f32 = contact_scene(n, n)
    st = State(orc, x, y, z, *vel, m, radius, f32, 2.0, 10.0, G, restitution=0.8)
    sh = ShardedSystem(n, nat.MODE_FAITHFUL, LocalComm([0] * world))
    sh.set_params(2.0, 10.0, G)
    sh.set_contacts(0.8, True)
    sh.set_history(16)
    sh.upload(x, y, z, *vel, m, radius, f32)
    sh.accel()
    sh.history_append()
    total = 0
    for k in (1, 2, 5):
        done, resolved = sh.step(k)
        assert done == k
        total += resolved
        for _ in range(k):
            st.step(1, collisions=True)
        s = sh.download_state()
        assert_bits(np.stack([s["x"], s["y"], s["z"]], 1), st.pos, f"pos after +{k}")
        assert_bits(np.stack([s["vx"], s["vy"], s["vz"]], 1), st.vel, f"vel after +{k}")
        assert_bits(sh.download_acc().T, st.acc, f"acc after +{k}")
        assert_bits(sh.potential(), st.U, f"U after +{k} (stashed before the push-out)")
    assert total == st.hits and total > 0
    assert sh.history_count() == 9
    assert_bits(sh.history_download(1)[0], st.pos, "history holds post-contact positions")
    sh.close()


def test_sharded_fast_contacts_same_sequence_as_single_gpu(nat, orc):
    """Fast mode with radii on 4 ranks: the pair-symmetric kernel's seed prefilter + exact re-test flags the same
    pairs per rank share; same number of contacts and the same state (to rounding) as the single-GPU fast engine."""
    n = 4096
    x, y, z, vel, m, radius, f32 = contact_scene(n, 5)
    one = nat.DeviceSystem(n, 0, nat.MODE_FAST)
    one.set_params(2.0, 10.0, G); one.set_contacts(0.8, True)
    one.upload(x, y, z, *vel, m, radius, f32); one.accel()
    done, r1 = one.step(6)
    from core.distributed import LocalComm, ShardedSystem
    sh = ShardedSystem(n, nat.MODE_FAST, LocalComm([0] * 4))
    sh.set_params(2.0, 10.0, G); sh.set_contacts(0.8, True)
    sh.upload(x, y, z, *vel, m, radius, f32); sh.accel()
    _, r4 = sh.step(6)
    assert r1 == r4 and r1 > 0
    s1, s4 = one.download_state(), sh.download_state()
    for k in ("x", "y", "z", "vx", "vy", "vz"):
        assert np.allclose(s4[k], s1[k], rtol=1e-10, atol=1e-6 if k[0] == "v" else 1e-3), k
    one.close(); sh.close()


@pytest.mark.parametrize("name", ["coll_dense_f32", "coll_dense_mixed", "coll_hit_f64_e05", "mixed12", "solar26_f32"])
@pytest.mark.parametrize("use_run", [True, False])
def test_engine_on_two_ranks_matches_reference_golden(golden, name, use_run):
    """SimulationEngine(devices=[0, 0]): the reference's own engine outputs, bit for bit, through a 2-rank
    ShardedSystem -- positions, velocities, accelerations, U, E, L, incl. the contact steps."""
    from core import distributed
    from tests.test_engine import build_engine, check_against_golden
    g = golden(name)
    if name.startswith("solar"):
        g = {k: g[k] for k in g.files}
        g["steps"] = np.array([s for s in g["steps"] if s <= 100])
    eng = build_engine(g, devices=[0, 0])
    assert isinstance(eng._dev, distributed.ShardedSystem) and eng._dev.world == 2
    check_against_golden(g, eng, use_run=use_run)
    eng.close()


def test_engine_history_and_frames_on_three_ranks(golden, tmp_path):
    from core.engine import SimulationEngine, load_frames
    from core.physics import ObjectCollection
    from tests.conftest import make_objects
    g = golden("coll_dense_mixed")
    kw = dict(dt=float(g["dt"]), softening=float(g["eps"]), restitution=float(g["restitution"]), max_hist=None,
              cache_every_n=3)
    one = SimulationEngine(ObjectCollection(make_objects(g)), cache_fp=str(tmp_path / "one.jsonl"), **kw)
    many = SimulationEngine(ObjectCollection(make_objects(g)), cache_fp=str(tmp_path / "many.jsonl"),
                            devices=[0, 0, 0], **kw)
    one.run(12); many.run(12)
    for a, b in zip(one.objects, many.objects):
        assert np.array_equal(a.position(), b.position()) and np.array_equal(a.velocity, b.velocity)
    u = many.objects[7].uuid
    assert many.history[u] == one.history[one.objects[7].uuid] and len(many.history[u]) == 13
    assert many.total_energy() == one.total_energy()
    f1, f2 = load_frames(str(tmp_path / "one.jsonl")), load_frames(str(tmp_path / "many.jsonl"))
    assert len(f1) == len(f2) == 4
    assert [o["coordinates"] for o in f1[-1]["objects"]] == [o["coordinates"] for o in f2[-1]["objects"]]
    one.close(); many.close()


# ---------------------------------------------------------------------------------------------------------
# pair-list overflow: more flagged pairs than the list holds -> exact list-free sweep, no loss
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,world", [(48, 1), (200, 1), (700, 1), (700, 2)])
def test_pair_list_overflow_falls_back_to_exact_full_sweep(nat, orc, n, world, monkeypatch):
    """ORBITAL_B200_OVERLAP_CAP=8 forces the overflow on the fused single-CTA kernel (48), the graph path (200,
    700) and a 2-rank engine: the sweep abandons the list and scans all pairs -- still bit-exact."""
    from core.distributed import LocalComm, ShardedSystem
    from oracle.c_oracle import State
    monkeypatch.setenv("ORBITAL_B200_OVERLAP_CAP", "8")
    x, y, z, vel, m, radius, f32 = contact_scene(n, 99 + n)
    radius = radius * {48: 6.0, 200: 3.0, 700: 2.5}[n]        # crowded: 20-80 touching pairs per step (oracle)
    st = State(orc, x, y, z, *vel, m, radius, f32, 2.0, 10.0, G, restitution=0.9)
    if world == 1:
        dev = nat.DeviceSystem(n, 0, nat.MODE_FAITHFUL)
    else:
        dev = ShardedSystem(n, nat.MODE_FAITHFUL, LocalComm([0] * world))
    dev.set_params(2.0, 10.0, G)
    dev.set_contacts(0.9, True)
    dev.upload(x, y, z, *vel, m, radius, f32)
    dev.accel()
    total = 0
    for k in (1, 3):
        done, resolved = dev.step(k)
        assert done == k
        total += resolved
        for _ in range(k):
            st.step(1, collisions=True)
        s = dev.download_state()
        assert_bits(np.stack([s["x"], s["y"], s["z"]], 1), st.pos, f"pos after +{k}")
        assert_bits(np.stack([s["vx"], s["vy"], s["vz"]], 1), st.vel, f"vel after +{k}")
    assert total == st.hits and total > 8
    assert dev.contact_stats()["full_sweeps"] > 0, "the scenario must actually overflow the 8-pair list"
    dev.close()
