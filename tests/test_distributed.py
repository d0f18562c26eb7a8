"""Multi-rank path of core.distributed.ShardedSystem.

CPU: world_size-2 gloo run over the oracle-backed stand-in (host-side logic: slabs, in-place
all-gather of packed positions, velocity gather, energy all-reduce).  GPU (`-m gpu`, needs >= 2 GPUs):
the same scenario on real devices over NCCL, bit-identical to the single-GPU run.
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


def _worker(rank, world, port, backend, n, steps, out_dir):
    for p in (os.path.join(REPO, "orbital-physics_b200"), REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from core import _native, synthetic
    from core.distributed import ShardedSystem
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    c = synthetic.plummer(n, seed=11)
    f32 = (np.arange(n) % 3 == 0).astype(np.uint8)
    vel = [np.where(f32 == 1, v.astype(np.float32).astype(np.float64), v) for v in (c["vx"], c["vy"], c["vz"])]
    if backend == "gloo":
        from tests.fake_device import FakeShardedDevice

        class Sys(ShardedSystem):
            def _make_device(self, mode):
                d = FakeShardedDevice(self.n, 0, mode, self.lo, self.hi)
                d.partial = out_dir.endswith("partial")
                return d

            def _view(self, which, shape):
                return torch.from_numpy({"pos4": self.dev.pos4, "vel": self.dev.vel, "acc": self.dev.acc}[which])

            def _bind_stream(self):
                pass
        mode = _native.MODE_FAITHFUL
        sysm = Sys(c["x"], c["y"], c["z"], *vel, c["m"], c["radius"], c["dt"], c["eps"], mode=mode, vel_is_f32=f32)
    else:
        torch.cuda.set_device(rank)
        mode = _native.MODE_FAITHFUL if out_dir.endswith("faithful") else _native.MODE_FAST
        sysm = ShardedSystem(c["x"], c["y"], c["z"], *vel, c["m"], c["radius"], c["dt"], c["eps"], mode=mode,
                             device=rank, vel_is_f32=f32)
    sysm.step(steps)
    st = sysm.gather_state()
    K, L = sysm.energy_angmom()
    if rank == 0:
        np.savez(os.path.join(out_dir, "result.npz"), K=K, L=L, **st)
    dist.barrier()
    sysm.close()
    dist.destroy_process_group()


def _reference(orc, n, steps):
    from core import synthetic
    from oracle.c_oracle import State
    c = synthetic.plummer(n, seed=11)
    f32 = (np.arange(n) % 3 == 0).astype(np.uint8)
    st = State(orc, *c.arrays(), f32, c["dt"], c["eps"])
    st.step(steps, collisions=False, nthreads=4)
    return st


def test_slab_partition():
    from core.distributed import slab
    from core.ensemble import partition
    assert [slab(16, 4, r) for r in range(4)] == [(0, 4), (4, 8), (8, 12), (12, 16)]
    with pytest.raises(ValueError):
        slab(10, 4, 0)
    parts = [partition(10, 4, r) for r in range(4)]
    assert parts == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sum(b - a for a, b in parts) == 10


@pytest.mark.parametrize("kind", ["gathered", "partial"])
def test_two_rank_gloo_matches_single_process(orc, tmp_path, kind):
    """kind=partial exercises the all-reduce of partial accelerations (pair-symmetric sharding)."""
    import torch.multiprocessing as mp
    n, steps = 256, 4
    out = tmp_path / kind
    out.mkdir()
    mp.spawn(_worker, args=(2, _free_port(), "gloo", n, steps, str(out)), nprocs=2, join=True)
    got = np.load(out / "result.npz")
    st = _reference(orc, n, steps)
    for k, ref in (("x", st.x), ("y", st.y), ("z", st.z), ("vx", st.vx), ("vy", st.vy), ("vz", st.vz)):
        assert np.array_equal(got[k], ref), k
    assert abs(float(got["K"]) - st.kinetic()) <= 1e-9 * st.kinetic()
    assert np.linalg.norm(got["L"] - st.angmom()) <= 1e-9 * np.linalg.norm(st.angmom()) + 1e-300


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("mode", ["faithful", "fast"])
def test_multi_rank_nccl_matches_single_gpu(orc, tmp_path, mode, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from core import _native, synthetic
    n, steps = 4096, 3
    out = tmp_path / mode
    out.mkdir()
    mp.spawn(_worker, args=(world, _free_port(), "nccl", n, steps, str(out)), nprocs=world, join=True)
    got = np.load(out / "result.npz")
    # single-GPU run of the same kernels through the split-step entry points
    c = synthetic.plummer(n, seed=11)
    f32 = (np.arange(n) % 3 == 0).astype(np.uint8)
    vel = [np.where(f32 == 1, v.astype(np.float32).astype(np.float64), v) for v in (c["vx"], c["vy"], c["vz"])]
    dev = _native.DeviceSystem(n, 0, _native.MODE_FAITHFUL if mode == "faithful" else _native.MODE_FAST)
    dev.set_params(c["dt"], c["eps"], c["G"])
    dev.upload(c["x"], c["y"], c["z"], *vel, c["m"], c["radius"], f32)
    dev.accel()
    for _ in range(steps):
        dev.step_begin(); dev.accel(); dev.step_kick()
    one = dev.download_state()
    if mode == "faithful":
        st = _reference(orc, n, steps)
        assert np.array_equal(got["x"], st.x) and np.array_equal(got["vx"], st.vx)
    for k in ("x", "y", "z", "vx", "vy", "vz"):
        if mode == "faithful":
            assert np.array_equal(got[k], one[k]), k
        else:   # the pair blocks are summed in a different order on 2 ranks -> last-bit differences only
            assert np.allclose(got[k], one[k], rtol=1e-12, atol=0), k
    dev.close()
