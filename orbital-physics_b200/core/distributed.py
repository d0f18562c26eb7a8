"""Multi-GPU large-N run: bodies partitioned over ranks, positions all-gathered each step.

SURVEY.md 8(e).  Every rank holds all N sources {x,y,z,m} (32 B per body) and integrates a contiguous
slab of ceil(N / world) bodies.  One leapfrog step (reference core/engine.py:65-97) is

    orb_step_begin   half-kick + drift of the local slab                       (rank-local)
    all_gather       the packed {x,y,z,m} slabs, in place                      (32 B x N in total)
    orb_step_force   force pass with the overlap test of engine.py:85 fused in (rank-local)
    [reduce_scatter  3 x N partial accelerations, pair-symmetric kernel only]  (24 B x N in, each rank keeps its slab)
    orb_step_kick    second half-kick of the local slab                        (rank-local)
    [contacts        only in a step where some rank flagged a pair: all-gather the velocities and the pair
                     lists; every rank then replays the reference's sequential sweep (physics.py:510-535)
                     over the identical full state -- bit-identical everywhere, each rank keeps its slab]
    orb_step_end     history append (all N bodies) + bookkeeping               (rank-local)

Bit-exact mode: every rank evaluates its own targets against all sources; a target's source order does
not depend on the partition, so the result is bit-identical to one GPU.  Fast mode: the pair-symmetric
kernel evaluates every unordered pair once, so each rank takes every world-th pair block (snake order, which balances the triangle) and
produces a partial acceleration of all N bodies; a reduce-scatter gives every rank the total for its slab.

Two communicators carry the same protocol:

* :class:`DistComm`  -- production: one process per GPU, ``torch.distributed`` (NCCL over NVLink; gloo in
  the CPU tests).  Every process constructs the same :class:`ShardedSystem` / ``SimulationEngine`` and makes
  the same calls in the same order (SPMD).
* :class:`LocalComm` -- all ranks live in this process, one handle per listed device; peers exchange slabs
  with device-to-device copies.  Listing one device several times (``devices=[0, 0]``) runs the real
  per-rank kernels of a multi-GPU job on a single GPU, which is how the sharded kernels are held to the
  oracle on a one-GPU box (tests/test_sharded_gpu.py).

:class:`ShardedSystem` has the interface of ``core._native.DeviceSystem``, so ``SimulationEngine`` drives
either (``SimulationEngine(..., devices=...)``): run(), history, diagnostics and JSONL frames are unchanged.
The reference has no distributed path; the per-step semantics are the reference's.
"""
from __future__ import annotations

import os

import numpy as np

from core import _native


def slab(n: int, world_size: int, rank: int):
    """Contiguous slab of ceil(n / world_size) bodies; the last ranks' slabs may be shorter (never empty)."""
    per = -(-n // world_size)
    lo, hi = rank * per, min(n, (rank + 1) * per)
    if lo >= hi:
        raise ValueError(f"n={n} bodies leave rank {rank} of {world_size} without a slab: use fewer ranks")
    return lo, hi


class _CudaView:
    """Zero-copy torch view of a device buffer owned by liborbital_b200 (CUDA array interface v3)."""

    def __init__(self, ptr: int, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


# --------------------------------------------------------------------------- communicators
class DistComm:
    """One process per GPU over torch.distributed: this process holds exactly one rank."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("DistComm needs torch.distributed.init_process_group() first")
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.local_ranks = [self.rank]
        if device is None:
            device = self.rank % max(1, _native.device_count())
        self.devices = {self.rank: int(device)}

    def all_gather_rows(self, bufs, per):
        """bufs[0]: [world * per, ...] tensor whose rows [rank*per, (rank+1)*per) are this rank's; in place."""
        flat = bufs[0].view(-1)
        k = flat.numel() // self.world
        self.dist.all_gather_into_tensor(flat, flat[self.rank * k:(self.rank + 1) * k], group=self.group)

    def all_reduce_sum(self, bufs):
        self.dist.all_reduce(bufs[0], group=self.group)

    def reduce_scatter_rows(self, bufs, per):
        """bufs[0]: [rows, world * per] partial sums; every rank ends up with the total of ITS columns
        [rank * per, (rank + 1) * per) in place (NCCL's in-place reduce-scatter) -- half the traffic of an all-reduce,
        and all the second half-kick needs."""
        t = bufs[0]
        if self.dist.get_backend(self.group) == "gloo":      # the CPU test backend has no reduce-scatter
            self.dist.all_reduce(t, group=self.group)
            return
        for c in range(t.shape[0]):
            row = t[c]
            self.dist.reduce_scatter_tensor(row[self.rank * per:(self.rank + 1) * per], row, group=self.group)

    def all_gather_host(self, items):
        """items[0]: any picklable object of this rank -> list over ranks."""
        out = [None] * self.world
        self.dist.all_gather_object(out, items[0], group=self.group)
        return out


class LocalComm:
    """All ranks in this process, rank r on devices[r] (a device may carry several ranks)."""

    def __init__(self, devices):
        import torch
        self.torch = torch
        devices = [int(d) for d in devices]
        self.world = len(devices)
        self.local_ranks = list(range(self.world))
        self.devices = dict(enumerate(devices))
        self.rank = 0

    def all_gather_rows(self, bufs, per):
        n_rows = bufs[0].shape[0]
        for r, src in enumerate(bufs):
            lo, hi = r * per, min(n_rows, (r + 1) * per)
            for q, dst in enumerate(bufs):
                if q != r:
                    dst[lo:hi].copy_(src[lo:hi], non_blocking=True)

    def all_reduce_sum(self, bufs):
        total = bufs[0].clone()
        for b in bufs[1:]:                      # fixed rank order: deterministic
            total += b.to(total.device, non_blocking=True)
        for b in bufs:
            b.copy_(total, non_blocking=True)

    def reduce_scatter_rows(self, bufs, per):
        total = bufs[0].clone()
        for b in bufs[1:]:                      # fixed rank order: deterministic
            total += b.to(total.device, non_blocking=True)
        for r, b in enumerate(bufs):
            b[:, r * per:(r + 1) * per].copy_(total[:, r * per:(r + 1) * per], non_blocking=True)

    def all_gather_host(self, items):
        return list(items)


# --------------------------------------------------------------------------- the sharded system
class ShardedSystem:
    """N bodies over `comm.world` ranks behind the interface of ``core._native.DeviceSystem``.

    All ranks pass the same full arrays to :meth:`upload`.  `step(k)` returns `(k, contacts resolved)` like a
    handle with device-side contact resolution -- sharded engines always resolve contacts on the device.
    """

    def __init__(self, n: int, mode: int = _native.MODE_FAST, comm=None, eager_state: bool = False):
        """eager_state: complete the velocities and accelerations on every rank at the end of each `step()` / `accel()`
        call, so that every later READ (download_state, download_acc, energy_angmom) is rank-local.  What
        SimulationEngine uses under a DistComm: its reader threads (the app's contract) and ranks that inspect state
        at different times must never start a collective on their own.  Costs two 24 B x N all-gathers per call."""
        self.eager_state = bool(eager_state)
        self.comm = comm if comm is not None else DistComm()
        self.torch = self.comm.torch
        self.n, self.mode = int(n), int(mode)
        self.world, self.rank = self.comm.world, self.comm.rank
        self.per = -(-self.n // self.world)
        self.slabs = [slab(self.n, self.world, r) for r in range(self.world)]     # raises if a rank would be empty
        self.lo, self.hi = self.slabs[self.rank]
        self.devs, self._pos4, self._vel, self._acc = [], [], [], []
        for r in self.comm.local_ranks:
            lo, hi = self.slabs[r]
            dev = self._make_device(r, lo, hi)
            self.devs.append(dev)
            self._pos4.append(self._view(dev, "pos4", (self.world * self.per, 4)))
            self._vel.append(self._view(dev, "vel", (3, self.n)))
            self._acc.append(self._view(dev, "acc", (3, self.n)))
            self._bind_stream(dev)
        self.dev = self.devs[0]
        self._partial = bool(self.dev.acc_needs_allreduce()) and self.world > 1
        self._peer = self._partial and self._open_peers()
        self._detect = False
        self._acc_full = True
        self._vel_full = True
        self.steps_done = 0
        self.events = None             # set to {} to collect (start, end) CUDA events per phase: "gather", "reduce",
                                       # "force" (bench.py); rank-local handles only

    @classmethod
    def from_arrays(cls, x, y, z, vx, vy, vz, m, radius, dt, eps, G=6.67430e-11, mode=_native.MODE_FAST,
                    comm=None, vel_is_f32=None, restitution=1.0):
        """Construct, configure, upload and run the constructor force pass (engine.py:41)."""
        s = cls(int(np.asarray(x).shape[0]), mode, comm)
        s.set_params(dt, eps, G)
        s.set_contacts(restitution, True)
        s.upload(x, y, z, vx, vy, vz, m, radius, vel_is_f32)
        s.accel()
        return s

    def _open_peers(self) -> bool:
        """One process per GPU over NCCL on one node: map the other ranks' acc buffers (CUDA IPC) so that the partial
        accelerations are summed straight out of peer memory (orb_peer_reduce) instead of by an NCCL collective.
        Every rank takes part in the two exchanges; the reduction is used only if it could be set up on ALL ranks
        (ORBITAL_B200_PEER_REDUCE=0: never)."""
        comm = self.comm
        if not isinstance(comm, DistComm) or self.n != self.per * self.world or not hasattr(self.dev, "peer_export"):
            return False
        if os.environ.get("ORBITAL_B200_PEER_REDUCE", "1") == "0" or comm.dist.get_backend(comm.group) != "nccl":
            return False
        try:
            mine = self.dev.peer_export()
        except _native.NativeError:
            mine = None
        every = comm.all_gather_host([mine])
        ok = all(h is not None for h in every)
        if ok:
            try:
                for r, (handle, offset) in enumerate(every):
                    if r != self.rank:
                        self.dev.peer_open(r, handle, offset)
            except _native.NativeError:
                ok = False
        ok = all(comm.all_gather_host([ok]))
        if ok:
            self._peer_flag = self.torch.zeros(1, device=f"cuda:{self.dev.device}")
        return ok

    # -- backend seams (overridden by the CPU test double) ------------------
    def _make_device(self, rank, lo, hi):
        return _native.DeviceSystem(self.n, self.comm.devices[rank], self.mode, lo, hi, rank=rank, world=self.world)

    def _view(self, dev, which, shape):
        ptr = {"pos4": dev.pos4_ptr, "vel": dev.vel_ptr, "acc": dev.acc_ptr}[which]()
        return self.torch.as_tensor(_CudaView(ptr, shape), device=f"cuda:{dev.device}")

    def _bind_stream(self, dev):
        # the library runs on torch's current stream of the handle's device, so torch's stream semantics order
        # its kernels with the collectives / peer copies
        dev.set_stream(self.torch.cuda.current_stream(dev.device).cuda_stream)

    # -- configuration ------------------------------------------------------
    def set_params(self, dt, eps, G=6.67430e-11):
        for d in self.devs:
            d.set_params(dt, eps, G)

    def set_mode(self, mode):
        for d in self.devs:
            d.set_mode(mode)
        self.mode = int(mode)
        self._partial = bool(self.dev.acc_needs_allreduce()) and self.world > 1

    def set_contacts(self, restitution, on_device=True):
        if not on_device:
            raise ValueError("sharded engines resolve contacts on the device (contacts='device')")
        for d in self.devs:
            d.set_contacts(restitution, True)

    def set_history(self, capacity):
        # every process keeps the full ring (any rank's engine may be asked for it); in-process ranks share one
        for k, d in enumerate(self.devs):
            d.set_history(capacity if k == 0 else 0)

    def set_stream(self, cuda_stream):
        for d in self.devs:
            d.set_stream(cuda_stream)

    # -- transfers ----------------------------------------------------------
    def upload(self, x, y, z, vx, vy, vz, m, radius, vel_is_f32=None):
        for d in self.devs:
            d.upload(x, y, z, vx, vy, vz, m, radius, vel_is_f32)
        self._detect = bool(np.any(np.asarray(radius) > 0.0))
        self._m = np.array(m, dtype=np.float64, copy=True)
        self._vel_full = True

    def _masses(self):
        return self._m

    def _gather_rows3(self, views):
        """Make rows [3, n] complete on every rank from the per-rank slabs (velocities, accelerations)."""
        if self.world == 1:
            return
        if isinstance(self.comm, LocalComm):
            for r, src in enumerate(views):
                lo, hi = self.slabs[r]
                for q, dst in enumerate(views):
                    if q != r:
                        dst[:, lo:hi].copy_(src[:, lo:hi], non_blocking=True)
            return
        t = self.torch
        v = views[0]
        tmp = t.zeros((self.world * self.per,), dtype=v.dtype, device=v.device)
        for c in range(3):
            piece = tmp[self.rank * self.per:(self.rank + 1) * self.per]
            piece[: self.hi - self.lo].copy_(v[c, self.lo:self.hi])
            self.comm.all_gather_rows([tmp], self.per)
            v[c].copy_(tmp[: self.n])

    def _complete_state(self):
        """Velocities and accelerations complete on every rank (collective)."""
        self._gather_rows3(self._vel)
        if not self._acc_full:
            self._gather_rows3(self._acc)
            self._acc_full = True
        self._vel_full = True

    def download_state(self, out=None):
        """Full x y z vx vy vz (positions are complete on every rank; velocities are gathered first unless the
        last step already did)."""
        if not (self.eager_state and self._vel_full):
            self._gather_rows3(self._vel)
        return self.dev.download_state(out)

    def download_acc(self):
        if not self._acc_full:
            self._gather_rows3(self._acc)
            self._acc_full = True
        return self.dev.download_acc()

    def upload_acc(self, acc3n):
        for d in self.devs:
            d.upload_acc(acc3n)
        self._acc_full = True

    # -- stepping -----------------------------------------------------------
    def _timed(self, phase, fn):
        if self.events is None:
            return fn()
        ev = self.torch.cuda.Event
        a, b = ev(enable_timing=True), ev(enable_timing=True)
        a.record()
        fn()
        b.record()
        self.events.setdefault(phase, []).append((a, b))

    def _all_gather_positions(self):
        if self.world > 1:
            self._timed("gather", lambda: self.comm.all_gather_rows(self._pos4, self.per))

    def _reduce_acc(self):
        """Sum the partial accelerations of the pair-symmetric kernel.  Equal slabs: reduce-scatter, every rank gets
        the total for its own slab only (what its half-kick reads); ragged slabs or small systems: all-reduce."""
        # measured on 2 x B200: N = 2,097,152 reduce-scatter (3 rows) 0.96 ms vs all-reduce 1.16 ms; N = 262,144
        # 0.41 vs 0.26 ms -- three latency-bound calls lose to one below ~1M bodies
        if self._peer:
            def peer():
                # every rank's force pass is complete before anyone reads peer memory (stream-ordered, no host sync);
                # the next step's position all-gather keeps the next force pass behind everybody's reduction
                self.comm.dist.all_reduce(self._peer_flag, group=self.comm.group)
                self.dev.peer_reduce()
            self._timed("reduce", peer)
            self._acc_full = False
        elif self._partial and self.n == self.per * self.world and self.n >= (1 << 20):
            self._timed("reduce", lambda: self.comm.reduce_scatter_rows(self._acc, self.per))
            self._acc_full = False
        elif self._partial:
            self._timed("reduce", lambda: self.comm.all_reduce_sum(self._acc))
            self._acc_full = True
        else:
            self._acc_full = self.world == 1

    def _force_pass(self):
        for d in self.devs:
            d.step_force()

    def accel(self):
        """Constructor force pass (engine.py:41): no overlap test."""
        if self._peer:
            # no position all-gather precedes this force pass: keep it behind everybody's previous peer reduction
            self.comm.dist.all_reduce(self._peer_flag, group=self.comm.group)
        for d in self.devs:
            d.accel()
        self._reduce_acc()
        if self.eager_state:
            self._complete_state()

    def _contacts(self) -> int:
        """engine.py:85 across ranks. Returns the number of candidate pairs handed to the sweep (0: none)."""
        local = [d.overlap_count() for d in self.devs]              # synchronises each local handle
        every = self.comm.all_gather_host(local)
        total = sum(c for c, _ in every)
        overflow = any(o for _, o in every)
        if total == 0 and not overflow:
            return 0
        self._gather_rows3(self._vel)                               # the sweep needs every body's velocity
        mine = [d.overlap_pairs(cap=max(1, c))[0] for d, (c, _) in zip(self.devs, local)]
        lists = self.comm.all_gather_host(mine)
        merged = np.concatenate([np.asarray(p, dtype=np.int64).reshape(-1, 2) for p in lists], axis=0)
        for d in self.devs:
            d.set_overlap_pairs(merged, overflow)
        return max(1, len(merged))

    def step(self, nsteps: int = 1):
        before = self.dev.contact_stats()["contacts_total"] if self._detect else 0
        for _ in range(int(nsteps)):
            for d in self.devs:
                d.step_begin()
            self._all_gather_positions()
            self._timed("force", self._force_pass)
            self._reduce_acc()
            for d in self.devs:
                d.step_kick()
            if self._detect:
                self._contacts()
            for d in self.devs:
                d.step_end()
        self.steps_done += int(nsteps)
        self._vel_full = False
        if self.eager_state and nsteps:
            self._complete_state()
        resolved = self.dev.contact_stats()["contacts_total"] - before if self._detect else 0
        return int(nsteps), int(resolved)

    def synchronize(self):
        for d in self.devs:
            d.synchronize()

    # -- results ------------------------------------------------------------
    def gather_state(self) -> dict:
        return self.download_state()

    def overlap_pairs(self, cap: int = 1 << 16):
        raise RuntimeError("sharded engines resolve contacts on the device; there is no halted pair list")

    def potential(self) -> float:
        """U of the resident positions: they are complete on every rank, so each evaluates it alone."""
        return self.dev.potential()

    def energy_angmom(self):
        """(K, L[3]) (engine.py:104-121): per-rank device reductions summed in rank order -- or, with eager_state,
        evaluated from the complete local state without any collective."""
        if self.eager_state and self._vel_full:
            st = self.dev.download_state()
            m = self._masses()
            K = float(0.5 * np.sum(m * (st["vx"] ** 2 + st["vy"] ** 2 + st["vz"] ** 2)))
            r = np.stack([st["x"], st["y"], st["z"]], 1)
            p = np.stack([st["vx"], st["vy"], st["vz"]], 1) * m[:, None]
            return K, np.cross(r, p).sum(axis=0)
        parts = [d.energy_angmom() for d in self.devs]
        local = [np.concatenate([[k], L]) for k, L in parts]
        every = self.comm.all_gather_host(local)
        tot = np.zeros(4)
        for p in every:
            tot += p
        return float(tot[0]), tot[1:].copy()

    def history_count(self):
        return self.dev.history_count()

    def history_append(self):
        for d in self.devs:
            d.history_append()

    def history_download(self, last_k):
        return self.dev.history_download(last_k)

    def force_kernel_info(self) -> dict:
        info = dict(self.dev.force_kernel_info())
        info["world"] = self.world
        return info

    def launch_count(self) -> int:
        return sum(d.launch_count() for d in self.devs)

    def contact_stats(self):
        return self.dev.contact_stats()

    def acc_needs_allreduce(self):
        return self._partial

    def close(self):
        """Collective when the ranks read each other's buffers (peer reduction): every importer unmaps, then a
        barrier, and only then does any rank free the buffer it exported."""
        if self._peer and self.devs:
            self.dev.peer_close()
            self.comm.dist.all_reduce(self._peer_flag, group=self.comm.group)
            self.torch.cuda.synchronize()
            self._peer = False
        for d in self.devs:
            d.close()
        self.devs = []


def make_comm(devices):
    """`devices` of SimulationEngine(..., devices=...): "dist" -> DistComm (torch.distributed must be initialised);
    an int N -> LocalComm over GPUs 0..N-1; a list -> LocalComm with one rank per entry."""
    if isinstance(devices, str):
        if devices.lower() != "dist":
            if devices.isdigit():
                return make_comm(int(devices))
            return make_comm([int(t) for t in devices.split(",") if t.strip()])
        return DistComm()
    if isinstance(devices, int):
        return LocalComm(range(devices))
    return LocalComm(devices)
