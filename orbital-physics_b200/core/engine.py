"""SimulationEngine: the kick-drift-kick leapfrog loop, device resident.

Drop-in for the reference's core/engine.py (same constructor, attributes and
methods; `run_simulation` prints the same diagnostics).  The state of the bound
`ObjectCollection` lives on the GPU as structure-of-arrays fp64 buffers; one
`step()` is one launch sequence through the C ABI (core/_native.py):

    half-kick + drift  ->  all-pairs force (+ fused overlap detection)
                       ->  half-kick + history-ring append

and `run(k)` keeps all k steps on the device (CUDA graph / fused single-CTA
kernel).  The Python `Object`s are lazy mirrors: the first read after a step
pulls the state once; a write makes the host copy authoritative and it is
re-uploaded before the next step.  Contacts are detected in the force pass and
resolved with the reference's sequential semantics -- on the device by default
(`contacts="device"`), or by the host replay (`contacts="host"`).

Single `step()` calls are DEFERRED: a step that nothing can observe only bumps a counter, and the device runs the
whole stretch in one launch when something looks (an attribute read or write of a bound object, `history`, `acc`,
the diagnostics, a JSONL frame, `run()`, a changed parameter or membership, `close()`).  The reference's own
driver loops -- `for _ in range(k): engine.step()` (core/examples.py:198-217) -- therefore cost one launch per
stretch instead of one launch + one device round trip per step (25 us -> ~1 us per step at N = 15); the results are
the same bits because the same steps run in the same order.  `ORBITAL_B200_DEFER=0` runs every step at once.

`devices=` spreads one system over several GPUs (core/distributed.py): the
same engine object, `run()`, `history`, diagnostics and JSONL frames on top of
a `ShardedSystem` instead of a single `DeviceSystem`.

Reference behaviour kept on purpose: `max_hist=-1` keeps only the
latest history point; `cache=True` by default appends a JSONL frame every
`cache_every_n` steps (frame time lags its positions by one dt); accelerations
are *not* recomputed after a collision or an external mutation.
"""
from __future__ import annotations

import json
import operator
import os
import sys
import threading
from collections.abc import Mapping

import numpy as np

from core import _native
from core.constants import STANDARD
from core.physics import Coordinates, ObjectCollection, default_device, select_mode

_RING_BYTES = int(os.environ.get("ORBITAL_B200_HISTORY_BYTES", str(256 << 20)))
_HOST_DIAG_MAX = 4096      # up to this many bodies diagnostics use the reference's NumPy expressions
_DEFER = os.environ.get("ORBITAL_B200_DEFER", "1") != "0"
_DEFER_MAX = 4096          # deferred steps are flushed at the latest when this many have piled up


class _HistoryView(Mapping):
    """`engine.history`: uuid -> list of [x, y, z], materialised lazily from the device ring."""

    def __init__(self, engine: "SimulationEngine"):
        self._e = engine
        self._override = {}

    def __getitem__(self, uuid):
        if uuid in self._override:
            return self._override[uuid]
        return self._e._history_list(uuid, 0)

    def __setitem__(self, uuid, value):
        self._override[uuid] = value

    def __iter__(self):
        return iter(self._e._uuids)

    def __len__(self):
        return len(self._e._uuids)

    def __repr__(self):
        return f"<history of {len(self)} bodies x {self._e._hist_len()} points>"


class SimulationEngine:
    """
    Engine to advance the orbital simulation forward in time.

    Attributes:
        objects (ObjectCollection): The collection of objects in the simulation.
        dt (float): Time step (in seconds).
        restitution (float): Coefficient of restitution for collisions (0 to 1).
        softening (float): Softening length to avoid singularities (in meters).
        history (mapping): uuid -> positions of that object at each recorded step.

    B200-specific keyword-only options (defaults reproduce the reference):
        mode:     "auto" | "faithful" | "fast"  (env ORBITAL_B200_MODE)
        device:   CUDA device index             (env ORBITAL_B200_DEVICE / LOCAL_RANK)
        contacts: "device" | "host"             (env ORBITAL_B200_CONTACTS) -- where overlapping pairs are
                  resolved; both replay the reference's sequential sweep exactly. "device" never leaves the GPU.
        devices:  None | N | [d0, d1, ...] | "dist"  (env ORBITAL_B200_DEVICES) -- multi-GPU: N GPUs driven by this
                  process, an explicit device per rank, or "dist" = one process per GPU under torchrun
                  (torch.distributed already initialised; every rank builds the same engine and makes the
                  same calls).  Bit-exact mode stays bit-identical to one GPU.  Needs contacts="device".
    """

    def __init__(self, objects: ObjectCollection, dt: float = 1.0, softening: float = 0.0,
                 restitution: float = 1.0, max_hist: int = -1, cache: bool = True,
                 cache_fp: str = "history.jsonl", cache_every_n: int = 300, *, mode=None, device=None,
                 contacts=None, devices=None):
        self._lock = threading.RLock()
        self.objects = objects
        self.dt = float(dt)
        self.softening = float(softening)
        self.restitution = float(restitution)
        self.max_hist = max_hist
        self.cache = cache
        if cache_fp and not cache_fp.endswith(".jsonl"):
            raise ValueError("cache_fp must end with .jsonl")
        self.cache_fp = cache_fp
        self._mode_req = mode
        self._contacts = (contacts or os.environ.get("ORBITAL_B200_CONTACTS", "device")).lower()
        if self._contacts not in ("device", "host"):
            raise ValueError("contacts must be 'device' or 'host'")
        self._device = default_device() if device is None else int(device)
        self._devices = devices if devices is not None else (os.environ.get("ORBITAL_B200_DEVICES") or None)
        self._comm = None
        if self._devices is not None and self._contacts != "device":
            raise ValueError("a multi-GPU engine resolves contacts on the device: use contacts='device'")
        self._G = STANDARD.G       # the reference never forwards unit_profile (engine.py:41,78)

        self._dev = None
        self._bound = []
        self._uuids = []
        self._index = {}
        self._snap = None               # host copy of what the device holds (SoA), refreshed by _pull
        self._pull_epoch = 0            # bumped by every _pull; an Object is current iff its _stamp equals it
        self._device_ahead = False
        self._pending = 0               # step() calls not yet run on the device (see _flush)
        self._touched = {}              # slot -> Object materialised / written since the last host->device sync
        self._watch = {}                # slot -> Object whose velocity ndarray was handed out (may be edited in place)
        self._host_dirty = False
        self._force_version = 0
        self._acc_cache = (None, None)
        self._U_cache = (None, None)
        self._params = None
        self._hist_chunks = []
        self._hist_drained = 0
        self._hist_total = 0
        self._hist_cache = (None, None)
        self._hist_legacy = {}
        self._history_view = _HistoryView(self)

        with self._lock:
            self._bind(initial=True)           # upload + initial accelerations (engine.py:41)
        self.time_elapsed = 0.
        self.step_idx = 0
        self.cache_every_n = cache_every_n if self.cache else 0

    # ------------------------------------------------------------------ binding
    def _gather(self):
        objs = self.objects.objects
        n = len(objs)
        pos = np.empty((3, n))
        vel = np.empty((3, n))
        m = np.empty(n)
        rad = np.empty(n)
        f32 = np.zeros(n, dtype=np.uint8)
        for i, o in enumerate(objs):
            c = o._coordinates
            pos[0, i], pos[1, i], pos[2, i] = c.x, c.y, c.z
            v = o._velocity
            if not isinstance(v, np.ndarray):
                v = np.asarray(v, dtype=np.float64)
            vel[0, i], vel[1, i], vel[2, i] = v[0], v[1], v[2]
            f32[i] = v.dtype == np.float32
            m[i] = o._mass
            rad[i] = o._radius
        return {"pos": pos, "vel": vel, "m": m, "radius": rad, "f32": f32}

    def _upload(self, g):
        self._dev.upload(g["pos"][0], g["pos"][1], g["pos"][2], g["vel"][0], g["vel"][1], g["vel"][2],
                         g["m"], g["radius"], g["f32"])
        self._snap = g

    def _sync_params(self):
        p = (float(self.dt), float(self.softening), float(self._G), float(self.restitution))
        if p != self._params:
            self._dev.set_params(*p[:3])
            self._dev.set_contacts(p[3], self._contacts == "device")
            self._params = p

    def _hist_limit(self):
        """None = unbounded, else the number of points the reference's pop(0) logic retains."""
        return None if self.max_hist is None else max(1, int(self.max_hist))

    def _bind(self, initial: bool):
        objs = list(self.objects.objects)
        n = len(objs)
        old_acc = None
        if not initial:
            old_acc = self.acc                              # uuid -> array, from the old binding
            for u in self._uuids:                           # keep what was recorded so far
                self._hist_legacy[u] = self._history_list(u, 0)
        if self._dev is not None:
            self._dev.close()
            self._dev = None
        for o in self._bound:
            if o._engine is self:
                o._engine = None
        self._touched, self._watch = {}, {}
        self._bound = objs
        self._uuids = [o.uuid for o in objs]
        self._index = {u: i for i, u in enumerate(self._uuids)}
        self._hist_chunks, self._hist_drained, self._hist_total = [], 0, 0
        self._hist_cache = (None, None)
        self._params = None
        if n == 0:
            self._snap = None
            self._device_ahead = self._host_dirty = False
            self._acc_cache = (self._force_version, {})
            self._U_cache = (self._force_version, 0.0)
            return
        self._dev = self._new_device(n, select_mode(n, self._mode_req))
        self._sync_params()
        limit = self._hist_limit()
        per_snapshot = 24 * n
        cap_max = max(1, _RING_BYTES // per_snapshot)
        self._ring_cap = limit if (limit is not None and limit <= cap_max) else cap_max
        self._drain_mode = limit is None or limit > cap_max
        self._dev.set_history(self._ring_cap)
        self._upload(self._gather())
        for i, o in enumerate(objs):
            o._engine = self
            o._slot = i
            o._stamp = self._pull_epoch                 # the objects ARE the state that was just uploaded
        self._device_ahead = self._host_dirty = False
        if initial:
            self._dev.accel()                               # engine.py:41
            self._force_version += 1
            self._dev.history_append()                      # engine.py:34 seed point
            self._hist_total = 1
        else:
            a = np.array([old_acc[u] for u in self._uuids]).T   # KeyError for late-added bodies, as the reference
            self._dev.upload_acc(np.ascontiguousarray(a))
            self._acc_cache = (self._force_version, {u: old_acc[u] for u in self._uuids})

    def _new_device(self, n, mode):
        if self._devices is None:
            return _native.DeviceSystem(n, self._device, mode)
        from core import distributed
        if self._comm is None:
            self._comm = distributed.make_comm(self._devices)
        if self._comm.world == 1:
            return _native.DeviceSystem(n, next(iter(self._comm.devices.values())), mode)
        # one process per GPU: reads must never start a collective (reader threads, ranks out of step) -> eager state
        return distributed.ShardedSystem(n, mode, self._comm, eager_state=isinstance(self._comm, distributed.DistComm))

    # ------------------------------------------------------------ lazy mirroring
    # After a step the device is ahead of the Python objects.  The first attribute read of ANY bound object pulls
    # the state once (one transfer into self._snap); each object then materialises its own Coordinates / velocity
    # from that snapshot when it is touched, so reading one body of a 262,144-body engine costs one download, not
    # 262,144 Python objects.  Objects whose velocity ndarray was handed out are refreshed at every pull and at
    # the end of every advance, because the reference updates that very array in place (engine.py:70) and callers
    # may hold on to it -- and may edit it in place, which _push_if_needed picks up.
    def _host_read(self, obj=None, velocity=False):
        with self._lock:
            if self._device_ahead:
                self._pull()
            if obj is not None and obj._engine is self:
                self._materialize(obj)
                self._touched[obj._slot] = obj
                if velocity and obj._slot not in self._watch:
                    self._watch[obj._slot] = obj
                    obj._vel_seen = self._vel_key(obj)

    def _host_write(self, obj=None):
        with self._lock:
            if self._device_ahead:
                self._pull()
            if obj is not None and obj._engine is self:
                self._materialize(obj)
                self._touched[obj._slot] = obj
            self._host_dirty = True

    @staticmethod
    def _vel_key(obj):
        v = obj._velocity
        return (float(v[0]), float(v[1]), float(v[2]), str(getattr(v, "dtype", "")))

    def _materialize(self, o):
        """Bring one bound object up to the last pulled snapshot."""
        if o._stamp == self._pull_epoch:
            return
        i, snap = o._slot, self._snap
        p, w = snap["pos"], snap["vel"]
        o._coordinates = Coordinates(x=p[0, i], y=p[1, i], z=p[2, i])     # np.float64 members, like from_iterable
        v = o._velocity
        if isinstance(v, np.ndarray) and v.shape == (3,) and v.dtype in (np.float32, np.float64):
            v[0], v[1], v[2] = w[0, i], w[1, i], w[2, i]                  # in place: dtype and identity kept
        else:
            o._velocity = np.array([w[0, i], w[1, i], w[2, i]])
        o._stamp = self._pull_epoch
        if i in self._watch:
            o._vel_seen = self._vel_key(o)

    def _materialize_all(self):
        for o in self._bound:
            self._materialize(o)

    def _pull(self):
        """Device -> host snapshot (one transfer); objects follow lazily."""
        self._flush()
        st = self._dev.download_state()
        s = self._snap
        s["pos"] = np.stack([st["x"], st["y"], st["z"]])
        s["vel"] = np.stack([st["vx"], st["vy"], st["vz"]])
        self._pull_epoch += 1
        self._device_ahead = False
        for o in self._watch.values():
            self._materialize(o)

    def _push_if_needed(self):
        """Host -> device if a touched object differs from the last synchronised snapshot."""
        self._flush()                        # deferred steps were asked for before whatever changed
        objs = self.objects.objects
        if len(objs) != len(self._bound) or any(a is not b for a, b in zip(objs, self._bound)):
            if self._device_ahead:
                self._pull()
            self._materialize_all()
            self._bind(initial=False)
            return
        if self._dev is None:
            return
        self._sync_params()
        if not (self._touched or self._watch or self._host_dirty):
            return
        cands = dict(self._watch)
        cands.update(self._touched)
        s = self._snap
        changed = False
        for i, o in cands.items():
            if o._stamp != self._pull_epoch:
                continue                     # never exposed since the last pull: cannot have been edited
            c = o._coordinates
            v = o._velocity
            if not isinstance(v, np.ndarray):
                v = np.asarray(v, dtype=np.float64)
            now = (c.x, c.y, c.z, v[0], v[1], v[2], o._mass, o._radius, v.dtype == np.float32)
            was = (s["pos"][0, i], s["pos"][1, i], s["pos"][2, i], s["vel"][0, i], s["vel"][1, i], s["vel"][2, i],
                   s["m"][i], s["radius"][i], bool(s["f32"][i]))
            if not all(a == b for a, b in zip(now, was)):
                if not changed:
                    if len(self._bound) <= _HOST_DIAG_MAX:
                        self._potential()    # U belongs to the last force build: evaluate before positions change
                    s = {k: np.array(a, copy=True) for k, a in s.items()}
                    changed = True
                s["pos"][:, i] = now[0:3]
                s["vel"][:, i] = now[3:6]
                s["m"][i], s["radius"][i], s["f32"][i] = now[6], now[7], now[8]
        if changed:
            self._upload(s)
            for o in self._watch.values():
                o._vel_seen = self._vel_key(o)
        self._touched = {}
        self._host_dirty = False

    # ------------------------------------------------------------------ history
    def _hist_len(self):
        self._flush()
        limit = self._hist_limit()
        return self._hist_total if limit is None else min(limit, self._hist_total)

    def _drain(self):
        new = self._hist_total - self._hist_drained
        if new > 0:
            self._hist_chunks.append(self._dev.history_download(new))
            self._hist_drained = self._hist_total
            limit = self._hist_limit()
            if limit is not None:
                tot = sum(c.shape[0] for c in self._hist_chunks)
                while tot - self._hist_chunks[0].shape[0] >= limit:
                    tot -= self._hist_chunks.pop(0).shape[0]

    def _history_array(self):
        """[T, n, 3] of the retained points, oldest first (cached until the next append)."""
        with self._lock:
            self._flush()
            key, arr = self._hist_cache
            if key == self._hist_total and arr is not None:
                return arr
            if self._dev is None:
                arr = np.empty((0, 0, 3))
            elif self._drain_mode:
                self._drain()
                arr = np.concatenate(self._hist_chunks, axis=0) if self._hist_chunks else np.empty((0, len(self._bound), 3))
                limit = self._hist_limit()
                if limit is not None:
                    arr = arr[-limit:]
            else:
                arr = self._dev.history_download(self._hist_len())
            self._hist_cache = (self._hist_total, arr)
            return arr

    def _history_list(self, uuid, limit):
        legacy = self._hist_legacy.get(uuid, [])
        if uuid not in self._index:
            if legacy:
                return legacy[-limit:] if limit > 0 else legacy
            raise KeyError(uuid)
        arr = self._history_array()
        col = arr[:, self._index[uuid], :]
        if legacy:
            pts = legacy + col.tolist()
            cap = self._hist_limit()
            if cap is not None:
                pts = pts[-cap:]
            return pts[-limit:] if limit > 0 else pts
        if limit > 0:
            col = col[-limit:]
        return col.tolist()

    @property
    def history(self):
        return self._history_view

    @history.setter
    def history(self, value):
        self._history_view._override = dict(value)

    def named_history(self, limit: int = 0):
        """Return history with object names as keys instead of UUIDs."""
        with self._lock:
            out = {}
            for obj in self.objects:
                if obj.uuid in self._history_view._override:
                    pts = self._history_view._override[obj.uuid]
                    out[obj.name] = pts[-limit:] if limit > 0 else pts
                else:
                    out[obj.name] = self._history_list(obj.uuid, limit)
            return out

    # --------------------------------------------------------------- force state
    @property
    def acc(self):
        """uuid -> acceleration (np.ndarray(3)) from the last force build."""
        with self._lock:
            self._flush()
            ver, cached = self._acc_cache
            if ver == self._force_version and cached is not None:
                return cached
            a = self._dev.download_acc() if self._dev is not None else np.empty((3, 0))
            out = {u: np.array([a[0, i], a[1, i], a[2, i]]) for i, u in enumerate(self._uuids)}
            self._acc_cache = (self._force_version, out)
            return out

    @acc.setter
    def acc(self, value):
        with self._lock:
            self._flush()
            a = np.array([np.asarray(value[u], dtype=np.float64) for u in self._uuids]).T
            self._dev.upload_acc(np.ascontiguousarray(a))
            self._acc_cache = (self._force_version, dict(value))

    def _potential(self):
        self._flush()
        ver, U = self._U_cache
        if ver == self._force_version and U is not None:
            return U
        U = np.float64(self._dev.potential()) if (self._dev is not None and len(self._bound) > 1) else 0.0
        self._U_cache = (self._force_version, U)
        return U

    @property
    def last_potential(self):
        """Total (softened) potential energy at the last force build (computed on first use)."""
        with self._lock:
            return self._potential()

    @last_potential.setter
    def last_potential(self, value):
        with self._lock:
            self._flush()
            self._U_cache = (self._force_version, value)

    # ------------------------------------------------------------------- stepping
    def _resolve_contacts(self):
        """Device halted on overlapping pairs: apply the reference's sweep to them on the host."""
        n = len(self._bound)
        if n <= _HOST_DIAG_MAX:
            self._potential()                       # U of this force build, before push-out moves bodies
        pairs, count = self._dev.overlap_pairs()
        self._pull()
        self._materialize_all()                     # the host sweep reads every body
        if count > len(pairs):                      # list overflowed: fall back to the full sweep
            self.objects.handle_collisions(restitution=self.restitution)
        else:
            self.objects.resolve_contacts(pairs, restitution=self.restitution)
        g = self._gather()
        self._upload(g)
        self._touched, self._host_dirty = {}, False
        for o in self._watch.values():
            o._vel_seen = self._vel_key(o)
        self._dev.history_append()                  # engine.py:88-92 runs after the collision sweep

    def _run_device(self, nsteps: int):
        """nsteps steps on the device as it is bound and parametrised right now (no host -> device sync)."""
        left = int(nsteps)
        while left > 0:
            chunk = left
            if self._drain_mode:
                room = self._ring_cap - (self._hist_total - self._hist_drained)
                if room <= 0:
                    self._drain()
                    room = self._ring_cap
                chunk = min(chunk, room)
            done, overlaps = self._dev.step(chunk)
            self._device_ahead = True
            self._force_version += done
            self._hist_total += done
            left -= done
            if self._contacts == "device":
                if done < chunk:
                    raise RuntimeError("device stopped early although contacts are resolved on the device")
            elif overlaps > 0:
                self._resolve_contacts()
            elif done < chunk:
                raise RuntimeError("device stopped early without reporting a contact")

    def _flush(self):
        """Run the step() calls that were deferred.  While steps are pending nothing on the host is touched or
        dirty (any access flushes first), so they run on the binding and the parameters they were asked for."""
        k = self._pending
        if k:
            self._pending = 0
            self._run_device(k)

    def _can_defer(self):
        """A step() that nothing can observe yet: same members, same parameters, no host-side edits to upload, no
        velocity array handed out (the reference updates those in place every step), one process, one GPU."""
        if not _DEFER or self._dev is None or self._devices is not None:
            return False
        if self._watch or self._touched or self._host_dirty:
            return False
        objs = self.objects.objects
        if len(objs) != len(self._bound) or not all(map(operator.is_, objs, self._bound)):
            return False
        p = self._params                     # (dt, softening, G, restitution) as last sent to the device
        return (p is not None and self.dt == p[0] and self.softening == p[1] and self._G == p[2]
                and self.restitution == p[3])

    def _prune_watch(self):
        """Stop tracking velocity arrays nobody holds any more: once the caller has dropped the ndarray it got from
        `obj.velocity` (reference count back to the attribute's own), no in-place edit can arrive through it and no
        one can see it lag, so the object returns to the lazy path (edits made while it was held are in _touched)."""
        for slot in [s for s, o in self._watch.items() if sys.getrefcount(o._velocity) <= 2]:
            del self._watch[slot]

    def _advance(self, nsteps: int):
        """nsteps complete steps (engine.py:69-92), all on the device unless a contact halts it."""
        self._push_if_needed()
        if self._dev is None:
            return
        self._run_device(nsteps)
        if self._watch:
            self._prune_watch()
        if self._watch:
            self._pull()                # velocity arrays that were handed out track the state, as in the reference

    def _tick(self):
        """Book-keeping after one step (engine.py:94-97)."""
        if self.cache and ((self.step_idx % self.cache_every_n) == 0):
            self.save_frame()
        self.step_idx += 1
        self.time_elapsed += self.dt

    def step(self):
        with self._lock:
            if self._can_defer():
                if self._pending >= _DEFER_MAX:
                    self._flush()
                self._pending += 1
                self._device_ahead = True       # the mirrors are behind: the next read pulls, and the pull flushes
            else:
                self._advance(1)
            self._tick()

    def run(self, steps: int):
        """`steps` leapfrog steps; the device runs whole stretches between JSONL frames."""
        with self._lock:
            left = int(steps)
            while left > 0:
                k = left
                if self.cache and self.cache_every_n > 0:
                    k = min(k, (-self.step_idx) % self.cache_every_n + 1)   # stop right after a frame step
                elif self.cache:
                    k = 1                                                     # modulo by zero, as the reference
                self._advance(k)
                for _ in range(k - 1):                                        # no frame due inside the stretch
                    self.step_idx += 1
                    self.time_elapsed += self.dt
                self._tick()
                left -= k

    def save_frame(self):
        """Append the current state to the JSONL cache (same schema as the reference, engine.py:48-57)."""
        state = {
            "time_elapsed": self.time_elapsed,
            "objects": self.objects.to_dict(),
            "history": self.named_history(limit=1),
        }
        with open(self.cache_fp, "a") as f:
            json.dump(state, f)
            f.write('\n')

    def _peek_velocity(self, obj):
        """obj.velocity for the engine's own read-only use: does not count as handing the array out."""
        if obj._engine is not self:
            return obj.velocity
        if self._device_ahead:
            self._pull()
        self._materialize(obj)
        return obj._velocity

    # ---------------------------------------------------------------- diagnostics
    def total_energy(self):
        with self._lock:
            if len(self._bound) > _HOST_DIAG_MAX:
                self._push_if_needed()
                K, _ = self._dev.energy_angmom()
                return K + self.last_potential
            K = 0.0
            for obj in self.objects:
                vel = self._peek_velocity(obj)
                v2 = float(vel @ vel)
                K += 0.5 * obj.mass * v2
            return K + self.last_potential

    def _body_potential(self, obj, system, G):
        """Potential term of `obj.lagrangian(system)` on the device (orb_body_potential: the reference's loop order,
        core/physics.py:275-279), or None when `system` is not exactly this engine's bodies in their order (the
        caller then runs the host loop, which reads the lazy mirrors)."""
        with self._lock:
            if self._dev is None or not hasattr(self._dev, "body_potential") or len(self._bound) < 16:
                return None
            if system is not self.objects and system is not self.objects.objects:
                return None
            self._push_if_needed()               # deferred steps, host-side edits, membership
            if obj._engine is not self:
                return None
            return self._dev.body_potential(obj._slot, float(G))

    def angular_momentum(self):
        with self._lock:
            if len(self._bound) > _HOST_DIAG_MAX:
                self._push_if_needed()
                return self._dev.energy_angmom()[1]
            L = np.zeros(3)
            for obj in self.objects:
                L += np.cross(obj.position(), obj.mass * self._peek_velocity(obj))
            return L

    # -------------------------------------------------------------------- resume
    @classmethod
    def resume(cls, cache_fp: str, index: int = -1, dt: float = 1.0, velocity_dtype: str = "auto", **kwargs):
        """Rebuild an engine from a JSONL frame written by :meth:`save_frame` and continue the run.

        The reference writes frames (core/engine.py:48-57) but has no loader; this is the missing half.
        `index` selects the frame (default: the last one).  A frame stores positions *after* its step but the
        time *before* it (engine.py:94-97), so the resumed clock is `frame time + dt`.  Velocities are stored as
        JSON numbers: with `velocity_dtype="auto"` a body whose three components are exactly representable in
        float32 resumes as a float32-velocity body (how `Object(...)` stores them), anything else as float64, so
        a resumed run continues bit-identically in both of the reference's velocity modes.  Names come from the
        frame's `history` keys (`Object.to_dict` omits them).
        """
        from core.physics import Object
        frames = load_frames(cache_fp)
        frame = frames[index]
        names = list(frame.get("history", {}).keys())
        objs = []
        for k, d in enumerate(frame["objects"]):
            d = dict(d)
            if k < len(names):
                d["name"] = names[k]
            o = Object.from_dict(d)
            v64 = np.array(d["velocity"], dtype=np.float64)
            as32 = v64.astype(np.float32)
            if velocity_dtype == "float64" or (velocity_dtype == "auto" and not np.array_equal(as32.astype(np.float64), v64)):
                o.velocity = v64
            objs.append(o)
        kwargs.setdefault("cache_fp", cache_fp)
        eng = cls(ObjectCollection(objs), dt=dt, **kwargs)
        eng.time_elapsed = float(frame["time_elapsed"]) + eng.dt
        eng.step_idx = int(round(eng.time_elapsed / eng.dt)) if eng.dt else 0
        return eng

    # ---------------------------------------------------------------------- misc
    def synchronize(self):
        with self._lock:
            self._flush()
            if self._dev is not None:
                self._dev.synchronize()

    def kernel_info(self) -> dict:
        with self._lock:
            self._flush()
            return self._dev.force_kernel_info()

    def close(self):
        with self._lock:
            if self._device_ahead:
                self._pull()
            if self._dev is not None:
                self._materialize_all()
            for o in self._bound:
                if o._engine is self:
                    o._engine = None
            if self._dev is not None:
                self._dev.close()
                self._dev = None


def load_frames(cache_fp: str) -> list:
    """All frames of a JSONL cache file (one dict per line: time_elapsed, objects, history)."""
    with open(cache_fp) as f:
        return [json.loads(line) for line in f if line.strip()]


def run_simulation(engine: SimulationEngine, steps: int, print_every: int = 100):
    """Step `steps` times, printing relative energy / angular-momentum drift every `print_every` steps
    (at s = 0, print_every, 2*print_every, ... counted before the increment, as the reference)."""
    E0 = engine.total_energy()
    L0 = engine.angular_momentum()
    s = 0
    while s < steps:
        # the reference evaluates E, L after the step whose index is a multiple of print_every
        nxt = s if s % print_every == 0 else s + (print_every - s % print_every)
        k = min(steps, nxt + 1) - s
        if k > 1:
            engine.run(k - 1)
        engine.step()
        s += k
        if (s - 1) % print_every == 0:
            E = engine.total_energy()
            L = engine.angular_momentum()
            dE = (E - E0) / abs(E0)
            dL = np.linalg.norm(L - L0) / (np.linalg.norm(L0) + 1e-30)
            print(f"step {s - 1}: ΔE={dE:.3e}, ΔL={dL:.3e}")
