"""Multi-rank path of core.distributed (host-side protocol).

CPU: world_size-2 gloo runs over the oracle-backed stand-in (slabs incl. ragged ones, in-place all-gather of
packed positions, partial-acceleration all-reduce, contact lists merged across ranks + replicated sweep,
velocity gather, energy reduction), and SimulationEngine(devices=[0, 0]) over the in-process communicator.
GPU (`-m gpu`): tests/test_sharded_gpu.py runs the real per-rank kernels on one device; the NCCL test below needs
>= 2 GPUs.
"""
import os
import socket
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scenario(n, kind):
    """Plummer cloud with every third velocity stored as float32; kind 'contacts' inflates the radii so that
    several pairs touch (and keep touching) during the run."""
    from core import synthetic
    c = synthetic.plummer(n, seed=11)
    f32 = (np.arange(n) % 3 == 0).astype(np.uint8)
    vel = [np.where(f32 == 1, v.astype(np.float32).astype(np.float64), v) for v in (c["vx"], c["vy"], c["vz"])]
    radius = c["radius"].copy()
    if kind == "contacts":
        P = np.stack([c["x"], c["y"], c["z"]], 1)
        d = np.linalg.norm(P[:, None, :] - P[None, :, :], axis=2) + np.eye(n) * 1e300
        radius[:] = 0.75 * np.sort(d.min(axis=1))[n // 4]      # about a quarter of the bodies start in contact
    return c, f32, vel, radius


def _fake_sharded_class(torch, partial):
    from core.distributed import ShardedSystem
    from tests.fake_device import FakeShardedDevice

    class Sys(ShardedSystem):
        def _make_device(self, rank, lo, hi):
            d = FakeShardedDevice(self.n, 0, self.mode, lo, hi, rank=rank, world=self.world)
            d.partial = partial
            return d

        def _view(self, dev, which, shape):
            return torch.from_numpy({"pos4": dev.pos4, "vel": dev.vel, "acc": dev.acc}[which])

        def _bind_stream(self, dev):
            pass
    return Sys


def _worker(rank, world, port, backend, n, steps, out_dir):
    for p in (os.path.join(REPO, "orbital-physics_b200"), REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from core import _native
    from core.distributed import DistComm, ShardedSystem
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    kind = os.path.basename(out_dir)
    if kind == "fast_nccl":                     # the same run with the partial accelerations summed by NCCL
        os.environ["ORBITAL_B200_PEER_REDUCE"] = "0"
    c, f32, vel, radius = _scenario(n, kind)
    if backend == "gloo":
        Sys = _fake_sharded_class(torch, partial=(kind == "partial"))
        mode = _native.MODE_FAITHFUL
        comm = DistComm()
    else:
        torch.cuda.set_device(rank)
        Sys = ShardedSystem
        mode = _native.MODE_FAITHFUL if kind in ("faithful", "contacts") else _native.MODE_FAST
        comm = DistComm(device=rank)
    sysm = Sys.from_arrays(c["x"], c["y"], c["z"], *vel, c["m"], radius, c["dt"], c["eps"], mode=mode, comm=comm,
                           vel_is_f32=f32)
    _, resolved = sysm.step(steps)
    st = sysm.gather_state()
    K, L = sysm.energy_angmom()
    if rank == 0:
        np.savez(os.path.join(out_dir, "result.npz"), K=K, L=L, resolved=resolved,
                 peer=int(bool(getattr(sysm, "_peer", False))), **st)
    dist.barrier()
    sysm.close()
    dist.destroy_process_group()


def _reference(orc, n, steps, kind="gathered"):
    from oracle.c_oracle import State
    c, f32, vel, radius = _scenario(n, kind)
    st = State(orc, c["x"], c["y"], c["z"], *vel, c["m"], radius, f32, c["dt"], c["eps"])
    st.step(steps, collisions=(kind == "contacts"), nthreads=1 if kind == "contacts" else 4)
    return st


def test_slab_partition():
    from core.distributed import slab
    from core.ensemble import partition
    assert [slab(16, 4, r) for r in range(4)] == [(0, 4), (4, 8), (8, 12), (12, 16)]
    assert [slab(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]      # ragged: ceil(n / world)
    with pytest.raises(ValueError):
        slab(9, 8, 5)                                                                   # a rank without bodies
    parts = [partition(10, 4, r) for r in range(4)]
    assert parts == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sum(b - a for a, b in parts) == 10


@pytest.mark.parametrize("kind,n", [("gathered", 256), ("partial", 256), ("gathered", 251), ("contacts", 96)])
def test_two_rank_gloo_matches_single_process(orc, tmp_path, kind, n):
    """kind=partial exercises the all-reduce of partial accelerations (pair-symmetric sharding); n=251 a ragged
    last slab; kind=contacts the cross-rank pair-list merge and the replicated sweep (engine.py:85)."""
    import torch.multiprocessing as mp
    steps = 4
    out = tmp_path / kind
    out.mkdir()
    mp.spawn(_worker, args=(2, _free_port(), "gloo", n, steps, str(out)), nprocs=2, join=True)
    got = np.load(out / "result.npz")
    st = _reference(orc, n, steps, kind)
    for k, ref in (("x", st.x), ("y", st.y), ("z", st.z), ("vx", st.vx), ("vy", st.vy), ("vz", st.vz)):
        assert np.array_equal(got[k], ref), k
    if kind == "contacts":
        assert st.hits > 0 and int(got["resolved"]) == st.hits
    # the device reduction is plain fp64; the oracle's K restates the reference's float32 dot for float32-velocity
    # bodies (engine.py:107), hence the float32-level tolerance
    assert abs(float(got["K"]) - st.kinetic()) <= 1e-7 * st.kinetic()
    assert np.linalg.norm(got["L"] - st.angmom()) <= 1e-7 * np.linalg.norm(st.angmom()) + 1e-300


def _patch_fake_sharded(monkeypatch):
    import torch
    from core import _native, distributed
    from tests.fake_device import FakeDeviceSystem
    monkeypatch.setattr(_native, "DeviceSystem", FakeDeviceSystem)
    monkeypatch.setattr(distributed, "ShardedSystem", _fake_sharded_class(torch, partial=False))


@pytest.mark.parametrize("name", ["coll_dense_mixed", "coll_hit_f32_e05", "mixed12"])
@pytest.mark.parametrize("use_run", [True, False])
def test_engine_over_in_process_ranks_matches_reference(golden, monkeypatch, name, use_run):
    """SimulationEngine(devices=[0, 0, 0]) over a ShardedSystem == the reference's own engine outputs, bit for bit,
    incl. contacts, U, E, L (host logic over the stand-in; the GPU twin is in tests/test_sharded_gpu.py)."""
    from core import distributed
    from tests.test_engine import build_engine, check_against_golden
    _patch_fake_sharded(monkeypatch)
    g = golden(name)
    eng = build_engine(g, devices=[0, 0, 0])
    assert isinstance(eng._dev, distributed.ShardedSystem) and eng._dev.world == 3
    check_against_golden(g, eng, use_run=use_run)


def test_engine_over_in_process_ranks_history_and_frames(golden, monkeypatch, tmp_path):
    """run / history / JSONL frames over a ShardedSystem equal the single-device engine."""
    from core.engine import SimulationEngine, load_frames
    from core.physics import ObjectCollection
    from tests.conftest import make_objects
    _patch_fake_sharded(monkeypatch)
    g = golden("coll_dense_mixed")
    kw = dict(dt=float(g["dt"]), softening=float(g["eps"]), restitution=float(g["restitution"]), max_hist=None,
              cache_every_n=3)
    steps = 10
    one = SimulationEngine(ObjectCollection(make_objects(g)), cache_fp=str(tmp_path / "one.jsonl"), **kw)
    many = SimulationEngine(ObjectCollection(make_objects(g)), cache_fp=str(tmp_path / "many.jsonl"),
                            devices="0,0", **kw)
    assert many._dev.world == 2
    one.run(steps)
    many.run(steps)
    for a, b in zip(one.objects, many.objects):
        assert np.array_equal(a.position(), b.position()) and np.array_equal(a.velocity, b.velocity)
        assert a.velocity.dtype == b.velocity.dtype
    u = many.objects[5].uuid
    assert many.history[u] == one.history[one.objects[5].uuid] and len(many.history[u]) == steps + 1
    f1, f2 = load_frames(str(tmp_path / "one.jsonl")), load_frames(str(tmp_path / "many.jsonl"))
    assert len(f1) == len(f2) > 0
    assert [o["coordinates"] for o in f1[-1]["objects"]] == [o["coordinates"] for o in f2[-1]["objects"]]
    with pytest.raises(ValueError):
        SimulationEngine(ObjectCollection(make_objects(g)), cache=False, devices=2, contacts="host")


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("kind", ["faithful", "fast", "fast_nccl", "faithful_contacts"])
def test_multi_rank_nccl_matches_single_gpu(orc, tmp_path, kind, world):
    """One process per GPU over NCCL.  kind=fast: the partial accelerations of the pair-symmetric kernel are summed
    straight out of peer memory (CUDA IPC, orb_peer_reduce); fast_nccl: the same run with NCCL's all-reduce."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (the same kernels run on one GPU in tests/test_sharded_gpu.py)")
    import torch.multiprocessing as mp
    n, steps = (4096, 3) if kind != "faithful_contacts" else (600, 4)
    out = tmp_path / ("contacts" if kind == "faithful_contacts" else kind)
    out.mkdir()
    mp.spawn(_worker, args=(world, _free_port(), "nccl", n, steps, str(out)), nprocs=world, join=True)
    got = np.load(out / "result.npz")
    st = _reference(orc, n, steps, "contacts" if kind == "faithful_contacts" else "gathered")
    for k, ref in (("x", st.x), ("y", st.y), ("z", st.z), ("vx", st.vx), ("vy", st.vy), ("vz", st.vz)):
        if kind.startswith("faithful"):
            assert np.array_equal(got[k], ref), k
        else:
            assert np.allclose(got[k], ref, rtol=1e-11, atol=0), k
    if kind.startswith("fast"):
        assert int(got["peer"]) == (1 if kind == "fast" else 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["coll_dense_mixed", "solar26_f32"])
def test_engine_over_two_real_gpus_in_one_process(golden, name):
    """SimulationEngine(devices=2): LocalComm with one handle per GPU, peer copies instead of NCCL -- bit-exact with
    the reference's own outputs (needs 2 GPUs; tests/test_sharded_gpu.py runs the same path with both ranks on GPU 0)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from core import distributed
    from tests.test_engine import build_engine, check_against_golden
    g = golden(name)
    if name.startswith("solar"):
        g = {k: g[k] for k in g.files}
        g["steps"] = np.array([s for s in g["steps"] if s <= 100])
    eng = build_engine(g, devices=2)
    assert isinstance(eng._dev, distributed.ShardedSystem) and sorted(eng._dev.comm.devices.values()) == [0, 1]
    check_against_golden(g, eng)
    eng.close()


@pytest.mark.gpu
def test_fast_sharded_over_two_real_gpus_in_one_process(orc):
    """Pair-symmetric kernel over 2 GPUs driven by one process: all rows <= 1e-12 of the oracle, 2 steps."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from core import _native
    from core.distributed import LocalComm, ShardedSystem
    c, f32, vel, radius = _scenario(8192, "gathered")
    sh = ShardedSystem.from_arrays(c["x"], c["y"], c["z"], *vel, c["m"], radius, c["dt"], c["eps"],
                                   mode=_native.MODE_FAST, comm=LocalComm([0, 1]), vel_is_f32=f32)
    sh.step(2)
    st = sh.download_state()
    ref, _ = orc.pairwise(st["x"], st["y"], st["z"], c["m"], c["eps"], 6.67430e-11, nthreads=8)
    got = sh.download_acc().T
    rel = np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert rel.max() <= 1e-12
    sh.close()


def _engine_dist_worker(rank, world, port, name):
    for p in (os.path.join(REPO, "orbital-physics_b200"), REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from core import _native, distributed
    from tests.fake_device import FakeDeviceSystem
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _native.DeviceSystem = FakeDeviceSystem
    distributed.ShardedSystem = _fake_sharded_class(torch, partial=False)
    from tests.test_engine import build_engine, check_against_golden
    g = np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))
    eng = build_engine(g, devices="dist")
    assert eng._dev.world == world and eng._dev.eager_state
    check_against_golden(g, eng)                     # every rank holds, and checks, the complete state
    eng.run(3)
    if rank == 0:
        # reads are rank-local in this mode: one rank (or a reader thread) may look at the state on its own
        _ = eng.objects[1].position(), eng.objects[2].velocity, eng.total_energy(), eng.history[eng.objects[0].uuid]
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["coll_dense_mixed", "mixed12"])
def test_engine_one_process_per_rank_gloo(name):
    """SimulationEngine(devices="dist") under torch.distributed (gloo, 2 ranks, oracle stand-in): the reference's own
    outputs bit for bit on every rank, and state reads that involve no collective."""
    import torch.multiprocessing as mp
    mp.spawn(_engine_dist_worker, args=(2, _free_port(), name), nprocs=2, join=True)
