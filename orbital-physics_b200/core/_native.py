"""ctypes binding of liborbital_b200.so (include/orbital_b200.h).

This is the only place the Python host layer touches native code.  There is no
CPU fallback: if the shared library is missing, or no CUDA device is present,
the first compute call raises :class:`NativeError` -- loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

MODE_FAITHFUL = 0
MODE_FAST = 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get(
    "ORBITAL_B200_LIB", os.path.join(os.path.dirname(_HERE), "csrc", "liborbital_b200.so"))

_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_i64p = C.POINTER(C.c_int64)
_dblp = C.POINTER(C.c_double)
_intp = C.POINTER(C.c_int)


class NativeError(RuntimeError):
    """A liborbital_b200 call failed (status code + orb_last_error())."""

    def __init__(self, code: int, message: str):
        super().__init__(f"liborbital_b200 error {code}: {message}")
        self.code = code


# every exported symbol: name -> (restype, argtypes).  tests/test_abi.py checks this
# table against include/orbital_b200.h.
SIGNATURES = {
    "orb_abi_version": (C.c_int, []),
    "orb_last_error": (C.c_char_p, []),
    "orb_device_count": (C.c_int, [_intp]),
    "orb_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_int, _intp, _intp, _intp, _i64p]),
    "orb_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_int64]),
    "orb_host_free": (C.c_int, [_vp]),
    "orb_fp64_peak": (C.c_int, [C.c_int, C.c_double, _dblp, _dblp, _dblp]),
    "orb_create": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int, C.c_int]),
    "orb_create_sharded": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "orb_create_ranked": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]),
    "orb_destroy": (C.c_int, [_vp]),
    "orb_set_params": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double]),
    "orb_set_mode": (C.c_int, [_vp, C.c_int]),
    "orb_set_contacts": (C.c_int, [_vp, C.c_double, C.c_int]),
    "orb_set_history": (C.c_int, [_vp, C.c_int64]),
    "orb_set_stream": (C.c_int, [_vp, _vp]),
    "orb_upload": (C.c_int, [_vp] + [_f64] * 8 + [_vp]),
    "orb_download_state": (C.c_int, [_vp] + [_vp] * 6),
    "orb_download_acc": (C.c_int, [_vp, _f64, _f64, _f64]),
    "orb_upload_acc": (C.c_int, [_vp, _f64, _f64, _f64]),
    "orb_accel": (C.c_int, [_vp]),
    "orb_step": (C.c_int, [_vp, C.c_int64, _i64p, _i64p]),
    "orb_overlap_pairs": (C.c_int, [_vp, _i64, C.c_int64, _i64p]),
    "orb_step_begin": (C.c_int, [_vp]),
    "orb_step_force": (C.c_int, [_vp]),
    "orb_step_finish": (C.c_int, [_vp]),
    "orb_step_end": (C.c_int, [_vp]),
    "orb_overlap_count": (C.c_int, [_vp, _i64p, _intp]),
    "orb_set_overlap_pairs": (C.c_int, [_vp, _vp, C.c_int64, C.c_int]),
    "orb_contact_stats": (C.c_int, [_vp, _i64p, _i64p]),
    "orb_step_kick": (C.c_int, [_vp]),
    "orb_acc_needs_allreduce": (C.c_int, [_vp, _intp]),
    "orb_synchronize": (C.c_int, [_vp]),
    "orb_pos4_ptr": (C.c_int, [_vp, C.POINTER(_vp), _i64p]),
    "orb_vel_ptr": (C.c_int, [_vp, C.POINTER(_vp)]),
    "orb_acc_ptr": (C.c_int, [_vp, C.POINTER(_vp)]),
    "orb_peer_export": (C.c_int, [_vp, C.c_char_p, _i64p]),
    "orb_peer_open": (C.c_int, [_vp, C.c_int, C.c_char_p, C.c_int64]),
    "orb_peer_reduce": (C.c_int, [_vp]),
    "orb_peer_close": (C.c_int, [_vp]),
    "orb_force_kernel_info": (C.c_int, [_vp, C.c_char_p, C.c_int, _intp, _intp, _intp, _intp]),
    "orb_launch_count": (C.c_int, [_vp, _i64p]),
    "orb_potential": (C.c_int, [_vp, _dblp]),
    "orb_body_potential": (C.c_int, [_vp, C.c_int64, C.c_double, _dblp]),
    "orb_energy_angmom": (C.c_int, [_vp, _dblp, _f64]),
    "orb_history_count": (C.c_int, [_vp, _i64p]),
    "orb_history_append": (C.c_int, [_vp]),
    "orb_history_download": (C.c_int, [_vp, C.c_int64, _vp, _i64p]),
    "orb_ens_create": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]),
    "orb_ens_destroy": (C.c_int, [_vp]),
    "orb_ens_set_params": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double]),
    "orb_ens_set_stream": (C.c_int, [_vp, _vp]),
    "orb_ens_set_bodies": (C.c_int, [_vp, _vp, _vp]),
    "orb_ens_set_contacts": (C.c_int, [_vp, C.c_double]),
    "orb_ens_contact_count": (C.c_int, [_vp, _i64p]),
    "orb_ens_download_acc": (C.c_int, [_vp, _vp, _vp, _vp]),
    "orb_ens_upload": (C.c_int, [_vp] + [_f64] * 7),
    "orb_ens_upload_elements": (C.c_int, [_vp] + [_f64] * 7),
    "orb_ens_step": (C.c_int, [_vp, C.c_int64, C.c_int]),
    "orb_ens_download": (C.c_int, [_vp] + [_vp] * 6),
    "orb_ens_energy": (C.c_int, [_vp, _f64]),
    "orb_ens_synchronize": (C.c_int, [_vp]),
    "orb_ens_launch_count": (C.c_int, [_vp, _i64p]),
    "orb_set_trig_mode": (C.c_int, [C.c_int]),
    "orb_get_trig_mode": (C.c_int, []),
    "orb_kepler_states": (C.c_int, [C.c_int, C.c_int64] + [_f64] * 8 + [C.c_double, C.c_int, _f64, _f64, _vp]),
}

_lib = None
_lib_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load the shared library once; raise NativeError if it is not built."""
    global _lib
    if _lib is None:
        with _lib_lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise NativeError(
                        -1, f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                            "g.build()'` or `make -C orbital-physics_b200/csrc` (there is no CPU fallback)")
                L = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                if L.orb_abi_version() != 1:
                    raise NativeError(-1, "ABI version mismatch between core/_native.py and liborbital_b200.so")
                _lib = L
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise NativeError(code, (lib().orb_last_error() or b"").decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int(0)
    check(lib().orb_device_count(C.byref(n)))
    return n.value


def device_info(device: int = 0) -> dict:
    name = C.create_string_buffer(256)
    sm, maj, mnr, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
    check(lib().orb_device_info(device, name, 256, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem)))
    return {"name": name.value.decode(), "sm_count": sm.value, "cc": (maj.value, mnr.value),
            "total_mem_bytes": mem.value}


def fp64_peak(device: int = 0, seconds: float = 1.0) -> dict:
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    check(lib().orb_fp64_peak(device, seconds, C.byref(a), C.byref(b), C.byref(c)))
    return {"tflops_best": a.value, "tflops_mean": b.value, "sm_clock_mhz": c.value}


def _ptr(a):
    """void* of a writable contiguous fp64 array, or NULL."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_vp)


_CUDA_STREAM_LEGACY = 0x1     # cudaStreamLegacy: the explicit handle of the NULL stream


def _stream_arg(cuda_stream):
    if cuda_stream is None:
        return None                     # NULL -> the library's own stream
    return _vp(int(cuda_stream) or _CUDA_STREAM_LEGACY)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class PinnedBuffer:
    """Page-locked host memory exposed as a NumPy fp64 array (orb_host_alloc)."""

    def __init__(self, n_doubles: int):
        self._p = _vp()
        check(lib().orb_host_alloc(C.byref(self._p), int(n_doubles) * 8))
        buf = (C.c_double * int(n_doubles)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.float64)

    def close(self):
        if self._p:
            self.array = None
            lib().orb_host_free(self._p)
            self._p = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceSystem:
    """One N-body system resident on a B200: thin object wrapper over the orb_* C ABI.

    State arrays are SoA fp64 NumPy arrays on the host side; nothing is computed
    on the host.  Used by core.engine.SimulationEngine, by bench.py and by the
    multi-GPU driver (core.distributed).
    """

    def __init__(self, n: int, device: int = 0, mode: int = MODE_FAITHFUL, tgt_lo: int | None = None,
                 tgt_hi: int | None = None, rank: int | None = None, world: int | None = None):
        self._h = _vp()
        self.n = int(n)
        self.device = int(device)
        self.mode = int(mode)
        self.tgt_lo = 0 if tgt_lo is None else int(tgt_lo)
        self.tgt_hi = self.n if tgt_hi is None else int(tgt_hi)
        self.pos4_capacity = self.n
        L = lib()
        if world is not None:
            check(L.orb_create_ranked(C.byref(self._h), self.n, self.tgt_lo, self.tgt_hi, int(rank or 0), int(world),
                                      self.device, self.mode))
            self.pos4_capacity = max(self.n, int(world) * -(-self.n // int(world)))
        elif self.tgt_lo == 0 and self.tgt_hi == self.n:
            check(L.orb_create(C.byref(self._h), self.n, self.device, self.mode))
        else:
            check(L.orb_create_sharded(C.byref(self._h), self.n, self.tgt_lo, self.tgt_hi, self.device, self.mode))

    # -- lifecycle ---------------------------------------------------------
    def close(self):
        if self._h:
            lib().orb_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- configuration -----------------------------------------------------
    def set_params(self, dt: float, eps: float, G: float = 6.67430e-11):
        check(lib().orb_set_params(self._h, float(dt), float(eps), float(G)))

    def set_mode(self, mode: int):
        check(lib().orb_set_mode(self._h, int(mode)))
        self.mode = int(mode)

    def set_contacts(self, restitution: float, on_device: bool):
        check(lib().orb_set_contacts(self._h, float(restitution), int(bool(on_device))))

    def set_history(self, capacity: int):
        check(lib().orb_set_history(self._h, int(capacity)))

    def set_stream(self, cuda_stream: int | None):
        """None: the handle's own stream. An integer cudaStream_t; 0 means the legacy default stream."""
        check(lib().orb_set_stream(self._h, _stream_arg(cuda_stream)))

    # -- transfers ---------------------------------------------------------
    def upload(self, x, y, z, vx, vy, vz, m, radius, vel_is_f32=None):
        arrs = [_c64(a) for a in (x, y, z, vx, vy, vz, m, radius)]
        for a in arrs:
            if a.shape != (self.n,):
                raise ValueError(f"expected arrays of shape ({self.n},), got {a.shape}")
        flags = None
        if vel_is_f32 is not None:
            flags = np.ascontiguousarray(np.broadcast_to(np.asarray(vel_is_f32, dtype=np.uint8), (self.n,)))
        check(lib().orb_upload(self._h, *arrs, flags.ctypes.data_as(_vp) if flags is not None else None))

    def download_state(self, out=None):
        """-> dict of x y z vx vy vz (fresh arrays, or views of `out` [6,n])."""
        if out is None:
            out = np.empty((6, self.n))
        check(lib().orb_download_state(self._h, *[_ptr(out[k]) for k in range(6)]))
        return dict(zip(("x", "y", "z", "vx", "vy", "vz"), out))

    def download_acc(self):
        a = np.empty((3, self.n))
        check(lib().orb_download_acc(self._h, a[0], a[1], a[2]))
        return a

    def upload_acc(self, acc3n):
        a = _c64(acc3n)
        check(lib().orb_upload_acc(self._h, a[0], a[1], a[2]))

    # -- hot path ----------------------------------------------------------
    def accel(self):
        check(lib().orb_accel(self._h))

    def step(self, nsteps: int = 1):
        """-> (steps_done, n_overlaps).  steps_done < nsteps iff the device halted on a contact."""
        done, nov = C.c_int64(0), C.c_int64(0)
        check(lib().orb_step(self._h, int(nsteps), C.byref(done), C.byref(nov)))
        return done.value, nov.value

    def overlap_pairs(self, cap: int = 1 << 16):
        buf = np.empty((cap, 2), dtype=np.int64)
        cnt = C.c_int64(0)
        check(lib().orb_overlap_pairs(self._h, buf.reshape(-1), cap, C.byref(cnt)))
        return buf[: min(cnt.value, cap)].copy(), cnt.value

    def step_begin(self):
        check(lib().orb_step_begin(self._h))

    def step_force(self):
        check(lib().orb_step_force(self._h))

    def step_finish(self):
        check(lib().orb_step_finish(self._h))

    def step_end(self):
        check(lib().orb_step_end(self._h))

    def overlap_count(self):
        """-> (pairs flagged by the last force pass, list overflowed?)  Synchronises."""
        c, o = C.c_int64(0), C.c_int(0)
        check(lib().orb_overlap_count(self._h, C.byref(c), C.byref(o)))
        return c.value, bool(o.value)

    def set_overlap_pairs(self, pairs, overflowed: bool = False):
        p = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
        check(lib().orb_set_overlap_pairs(self._h, p.ctypes.data_as(_vp) if len(p) else None, len(p),
                                          int(bool(overflowed))))

    def contact_stats(self):
        a, b = C.c_int64(0), C.c_int64(0)
        check(lib().orb_contact_stats(self._h, C.byref(a), C.byref(b)))
        return {"contacts_total": a.value, "full_sweeps": b.value}

    def step_kick(self):
        check(lib().orb_step_kick(self._h))

    def acc_needs_allreduce(self) -> bool:
        f = C.c_int(0)
        check(lib().orb_acc_needs_allreduce(self._h, C.byref(f)))
        return bool(f.value)

    def synchronize(self):
        check(lib().orb_synchronize(self._h))

    # -- device views ------------------------------------------------------
    def pos4_ptr(self) -> int:
        p, n = _vp(), C.c_int64()
        check(lib().orb_pos4_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value

    def vel_ptr(self) -> int:
        p = _vp()
        check(lib().orb_vel_ptr(self._h, C.byref(p)))
        return p.value

    def acc_ptr(self) -> int:
        p = _vp()
        check(lib().orb_acc_ptr(self._h, C.byref(p)))
        return p.value

    # -- peer-memory reduction of the partial accelerations (one process per GPU) --
    def peer_export(self):
        """(CUDA IPC handle: 64 bytes, offset of the acc buffer inside the exported allocation)."""
        h = C.create_string_buffer(64)
        off = C.c_int64()
        check(lib().orb_peer_export(self._h, h, C.byref(off)))
        return bytes(h.raw), int(off.value)

    def peer_open(self, rank: int, handle: bytes, offset: int):
        check(lib().orb_peer_open(self._h, int(rank), C.create_string_buffer(handle, 64), int(offset)))

    def peer_reduce(self):
        check(lib().orb_peer_reduce(self._h))

    def peer_close(self):
        check(lib().orb_peer_close(self._h))

    def force_kernel_info(self) -> dict:
        name = C.create_string_buffer(128)
        g, b, s, l = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib().orb_force_kernel_info(self._h, name, 128, C.byref(g), C.byref(b), C.byref(s), C.byref(l)))
        return {"name": name.value.decode(), "grid": g.value, "block": b.value, "smem": s.value,
                "launches_per_step": l.value}

    def launch_count(self) -> int:
        v = C.c_int64()
        check(lib().orb_launch_count(self._h, C.byref(v)))
        return v.value

    # -- diagnostics -------------------------------------------------------
    def potential(self) -> float:
        u = C.c_double()
        check(lib().orb_potential(self._h, C.byref(u)))
        return u.value

    def body_potential(self, body: int, G: float) -> float:
        """Potential term of Object.lagrangian for one body, reference order (orb_body_potential)."""
        u = C.c_double()
        check(lib().orb_body_potential(self._h, int(body), float(G), C.byref(u)))
        return u.value

    def energy_angmom(self):
        k = C.c_double()
        L3 = np.empty(3)
        check(lib().orb_energy_angmom(self._h, C.byref(k), L3))
        return k.value, L3

    # -- history -----------------------------------------------------------
    def history_count(self) -> int:
        v = C.c_int64()
        check(lib().orb_history_count(self._h, C.byref(v)))
        return v.value

    def history_append(self):
        check(lib().orb_history_append(self._h))

    def history_download(self, last_k: int) -> np.ndarray:
        """-> [k, n, 3] oldest first."""
        last_k = int(last_k)
        out = np.empty((max(last_k, 0), self.n, 3))
        got = C.c_int64(0)
        check(lib().orb_history_download(self._h, last_k, _ptr(out.reshape(-1)) if last_k > 0 else None, C.byref(got)))
        return out[: got.value]


TRIG_LIBM, TRIG_CR, TRIG_FAST = 0, 1, 2          # ORB_TRIG_* (include/orbital_b200.h)
_TRIG_NAMES = {"libm": TRIG_LIBM, "cr": TRIG_CR, "fast": TRIG_FAST}


def set_trig_mode(mode) -> None:
    """What stands in for the reference's math.sin / math.cos in the device IC pipeline: "libm" (default, glibc
    2.39 restated -> bit-identical states on an x86-64 glibc host), "cr" (correctly rounded), "fast" (CUDA sincos)."""
    check(lib().orb_set_trig_mode(int(_TRIG_NAMES.get(mode, mode))))


def get_trig_mode() -> int:
    return int(lib().orb_get_trig_mode())


def kepler_states(M, e, a, b, n, inc, Omega, omega, tol: float = 1e-12, max_iter: int = 50, device: int = 0,
                  return_E: bool = False):
    """Batched elements -> parent-relative (r[count,3], v[count,3]) on the device (orb_kepler_states)."""
    arrs = [_c64(np.atleast_1d(v)).reshape(-1) for v in (M, e, a, b, n, inc, Omega, omega)]
    count = arrs[0].shape[0]
    if any(v.shape[0] != count for v in arrs):
        raise ValueError("element arrays must have the same length")
    r3, v3 = np.empty((3, count)), np.empty((3, count))
    E = np.empty(count) if return_E else None
    check(lib().orb_kepler_states(int(device), count, *arrs, float(tol), int(max_iter), r3.reshape(-1),
                                  v3.reshape(-1), _ptr(E) if return_E else None))
    out = (np.ascontiguousarray(r3.T), np.ascontiguousarray(v3.T))
    return out + (E,) if return_E else out


class DeviceEnsemble:
    """nsys independent systems of nbody bodies (orb_ens_*): one warp per system in bit-exact mode, nbody/2 lanes per
    system (64/nbody systems per warp) in fast mode -- nbody lanes per system for small batches."""

    def __init__(self, nsys: int, nbody: int, device: int = 0, mode: int = MODE_FAST, vel_f32: bool = False):
        self._h = _vp()
        self.nsys, self.nbody, self.device, self.mode = int(nsys), int(nbody), int(device), int(mode)
        check(lib().orb_ens_create(C.byref(self._h), self.nsys, self.nbody, self.device, self.mode, int(bool(vel_f32))))

    def close(self):
        if self._h:
            lib().orb_ens_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, dt, eps, G=6.67430e-11):
        check(lib().orb_ens_set_params(self._h, float(dt), float(eps), float(G)))

    def set_stream(self, cuda_stream):
        check(lib().orb_ens_set_stream(self._h, _stream_arg(cuda_stream)))

    def upload(self, x, y, z, vx, vy, vz, m):
        arrs = [_c64(a) for a in (x, y, z, vx, vy, vz, m)]
        for a in arrs:
            if a.shape != (self.nsys, self.nbody):
                raise ValueError(f"expected shape ({self.nsys},{self.nbody}), got {a.shape}")
        check(lib().orb_ens_upload(self._h, *[a.reshape(-1) for a in arrs]))

    def set_bodies(self, radius=None, vel_is_f32=None):
        """Per-body radii (-> contact sweep every step) and float32-velocity flags, [nsys, nbody] each."""
        shape = (self.nsys, self.nbody)
        r = f = None
        if radius is not None:
            r = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, dtype=np.float64), shape))
        if vel_is_f32 is not None:
            f = np.ascontiguousarray(np.broadcast_to(np.asarray(vel_is_f32, dtype=np.uint8), shape))
        check(lib().orb_ens_set_bodies(self._h, r.ctypes.data_as(_vp) if r is not None else None,
                                       f.ctypes.data_as(_vp) if f is not None else None))

    def set_contacts(self, restitution: float):
        check(lib().orb_ens_set_contacts(self._h, float(restitution)))

    def contact_count(self) -> int:
        v = C.c_int64()
        check(lib().orb_ens_contact_count(self._h, C.byref(v)))
        return v.value

    def download_acc(self):
        a = np.empty((3, self.nsys, self.nbody))
        check(lib().orb_ens_download_acc(self._h, *[_ptr(a[k].reshape(-1)) for k in range(3)]))
        return a

    def info(self) -> dict:
        # one step per launch, 16 launches per graph: x,u,m in + x,u out (104 B) per body-step, plus the
        # accelerations read by the first and written by the last launch of a graph (2 x 24 B / 16)
        return {"bytes_per_body_step": 104.0 + 48.0 / 16.0,
                "kernel": "ens_step_fast_kernel" if self.mode == MODE_FAST else "ens_step_kernel"}

    def upload_elements(self, M, e, a, inc, Omega, omega, m):
        """Initial condition from orbital elements, generated on the device (orb_ens_upload_elements)."""
        els = [_c64(v) for v in (M, e, a, inc, Omega, omega)]
        for v in els:
            if v.shape != (self.nsys, self.nbody - 1):
                raise ValueError(f"expected element shape ({self.nsys},{self.nbody - 1}), got {v.shape}")
        m = _c64(m)
        if m.shape != (self.nsys, self.nbody):
            raise ValueError(f"expected mass shape ({self.nsys},{self.nbody}), got {m.shape}")
        check(lib().orb_ens_upload_elements(self._h, *[v.reshape(-1) for v in els], m.reshape(-1)))

    def step(self, nsteps: int = 1, fused: bool = True):
        check(lib().orb_ens_step(self._h, int(nsteps), int(bool(fused))))

    def download(self, out=None):
        if out is None:
            out = np.empty((6, self.nsys, self.nbody))
        check(lib().orb_ens_download(self._h, *[_ptr(out[k].reshape(-1)) for k in range(6)]))
        return dict(zip(("x", "y", "z", "vx", "vy", "vz"), out))

    def energy(self):
        E = np.empty(self.nsys)
        check(lib().orb_ens_energy(self._h, E))
        return E

    def synchronize(self):
        check(lib().orb_ens_synchronize(self._h))

    def launch_count(self) -> int:
        v = C.c_int64()
        check(lib().orb_ens_launch_count(self._h, C.byref(v)))
        return v.value
