"""Multi-GPU large-N run: target bodies partitioned by rank, positions all-gathered each step.

One process per GPU (torch.distributed, NCCL over NVLink); SURVEY.md 8(e).
Every rank holds all N sources {x,y,z,m} (32 B per body) and integrates the
contiguous target slab [lo, hi).  Per step:

    orb_step_begin   half-kick + drift of the local slab           (rank-local)
    all_gather       the packed {x,y,z,m} slabs, in place           (NCCL, 32*N bytes total)
    orb_accel        force pass                                     (rank-local)
    [all_reduce      3 x N accelerations, fast mode only]           (NCCL, 24*N bytes)
    orb_step_kick    half-kick of the local slab                    (rank-local)

Faithful mode: every rank evaluates its own targets against all sources; each
target's source order is unchanged by the partition, so the result is
bit-identical to the single-GPU run.  Fast mode: the pair-symmetric kernel
evaluates every unordered pair once, so each rank takes a cyclic share of the
pair blocks and produces a partial acceleration of all N bodies, summed by one
all-reduce.  The reference has no distributed path; the
per-step semantics are the reference's core/engine.py:65-97.
"""
from __future__ import annotations

import numpy as np

from core import _native


def slab(n: int, world_size: int, rank: int):
    """Equal contiguous target slabs (n must be divisible by world_size: pad with massless bodies otherwise)."""
    if n % world_size:
        raise ValueError(f"n={n} must be divisible by world_size={world_size}")
    per = n // world_size
    return rank * per, (rank + 1) * per


class _CudaView:
    """Zero-copy torch view of a device buffer owned by liborbital_b200 (CUDA array interface v3)."""

    def __init__(self, ptr: int, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class ShardedSystem:
    """N bodies over `world_size` ranks. All ranks pass the same full initial condition."""

    def __init__(self, x, y, z, vx, vy, vz, m, radius, dt, eps, G=6.67430e-11, mode=_native.MODE_FAST,
                 group=None, device=None, vel_is_f32=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n = int(np.asarray(x).shape[0])
        self.lo, self.hi = slab(self.n, self.world, self.rank)
        self.device = self.rank % max(1, _native.device_count()) if device is None else int(device)
        self.dev = self._make_device(mode)
        self.dev.set_params(dt, eps, G)
        self.dev.upload(x, y, z, vx, vy, vz, m, radius, vel_is_f32)
        self._pos4 = self._view("pos4", (self.n, 4))
        self._vel = self._view("vel", (3, self.n))
        self._acc = self._view("acc", (3, self.n))
        self._partial = bool(self.dev.acc_needs_allreduce()) and self.world > 1
        self._bind_stream()
        self._force()                         # engine.py:41
        self.steps_done = 0

    # -- backend seams (overridden by the CPU test double) ------------------
    def _make_device(self, mode):
        return _native.DeviceSystem(self.n, self.device, mode, self.lo, self.hi)

    def _view(self, which, shape):
        ptr = {"pos4": self.dev.pos4_ptr, "vel": self.dev.vel_ptr, "acc": self.dev.acc_ptr}[which]()
        return self.torch.as_tensor(_CudaView(ptr, shape), device=f"cuda:{self.device}")

    def _bind_stream(self):
        self.torch.cuda.set_device(self.device)
        self.dev.set_stream(self.torch.cuda.current_stream().cuda_stream)

    # -- stepping -----------------------------------------------------------
    def _all_gather_positions(self):
        if self.world == 1:
            return
        flat = self._pos4.view(-1)
        per = (self.hi - self.lo) * 4
        self.dist.all_gather_into_tensor(flat, flat[self.lo * 4: self.lo * 4 + per], group=self.group)

    def _force(self):
        self.dev.accel()
        if self._partial:
            self.dist.all_reduce(self._acc, group=self.group)

    def step(self, nsteps: int = 1):
        for _ in range(int(nsteps)):
            self.dev.step_begin()
            self._all_gather_positions()
            self._force()
            self.dev.step_kick()
        self.steps_done += int(nsteps)

    def synchronize(self):
        self.dev.synchronize()

    # -- results ------------------------------------------------------------
    def gather_state(self) -> dict:
        """Full x y z vx vy vz on every rank (velocities are all-gathered: each rank owns its slab)."""
        if self.world > 1:
            per = self.hi - self.lo
            for c in range(3):
                row = self._vel[c]
                self.dist.all_gather_into_tensor(row, row[self.lo: self.lo + per].clone(), group=self.group)
        return self.dev.download_state()

    def energy_angmom(self):
        """(K, L[3]) summed over ranks (engine.py:104-121)."""
        K, L = self.dev.energy_angmom()
        if self.world > 1:
            t = self.torch.tensor([K, L[0], L[1], L[2]], dtype=self.torch.float64, device=self._pos4.device)
            self.dist.all_reduce(t, group=self.group)
            K, L = float(t[0]), t[1:].cpu().numpy()
        return K, L

    def close(self):
        self.dev.close()
