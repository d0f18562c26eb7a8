"""State containers and the force operator of the B200-native engine.

Drop-in for the reference's core/physics.py (same public names and behaviour:
`Coordinates`, `Object`, `ObjectCollection`, `pairwise_accelerations`,
`collide_spheres`, `set_circular_orbit`, `solve_kepler`, `moment_of_inertia`, ...).

What is different underneath:
  * `pairwise_accelerations` (reference core/physics.py:125-159) runs on the GPU
    through the C ABI (core/_native.py); there is no Python pair loop and no CPU
    fallback.
  * an `Object` that belongs to a `SimulationEngine` is a *lazy mirror* of
    device-resident state: reading `coordinates` / `velocity` pulls the state
    from the device once if the device is ahead, writing marks the host copy as
    authoritative so the engine re-uploads before the next step.
  * contact *resolution* (`collide_spheres`, reference :391-422) stays on the
    host with the reference's sequential in-place semantics; contact *detection*
    (:517-518) is fused into the device force pass.
"""
from __future__ import annotations

import heapq
import math
import os
import threading
from dataclasses import dataclass
from typing import Iterable, Literal
from uuid import uuid4

import numpy as np

from core import _native
from core.constants import ASTRO, STANDARD, UnitProfile

_FAITHFUL_MAX = int(os.environ.get("ORBITAL_B200_FAITHFUL_MAX", "4096"))


def select_mode(n: int, mode: "str | int | None" = None) -> int:
    """'faithful' (bit-exact reference rounding) or 'fast' (roofline kernel, <=1e-12).

    Default ('auto', or env ORBITAL_B200_MODE): faithful up to 4096 bodies, fast above.
    """
    if mode is None:
        mode = os.environ.get("ORBITAL_B200_MODE", "auto")
    if isinstance(mode, int):
        return mode
    mode = mode.lower()
    if mode == "faithful":
        return _native.MODE_FAITHFUL
    if mode == "fast":
        return _native.MODE_FAST
    if mode != "auto":
        raise ValueError(f"unknown mode {mode!r} (faithful | fast | auto)")
    return _native.MODE_FAITHFUL if n <= _FAITHFUL_MAX else _native.MODE_FAST


def default_device() -> int:
    return int(os.environ.get("ORBITAL_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


# ---------------------------------------------------------------------------
# Coordinates
# ---------------------------------------------------------------------------
@dataclass
class Coordinates:
    """A point in 3-D space (arbitrary origin). Always truthy, even at the origin."""

    x: float
    y: float
    z: float

    def to_array(self) -> np.ndarray:
        return np.array([self.x, self.y, self.z])

    @classmethod
    def from_iterable(cls, lst: Iterable[float]) -> "Coordinates":
        return cls(x=lst[0], y=lst[1], z=lst[2])

    @classmethod
    def random(cls) -> "Coordinates":
        """Uniform in [-1, 1]^3."""
        u = np.random.uniform
        return cls(x=u(-1, 1), y=u(-1, 1), z=u(-1, 1))


# ---------------------------------------------------------------------------
# Small host-only helpers (unchanged semantics; not on the hot path)
# ---------------------------------------------------------------------------
def solve_kepler(M: float, e: float, tol: float = 1e-12, max_iter: int = 50) -> float:
    """Eccentric anomaly E with M = E - e sin E (Newton; start at M, or pi for e >= 0.8)."""
    E = M if e < 0.8 else math.pi
    for _ in range(max_iter):
        f = E - e * math.sin(E) - M
        fp = 1.0 - e * math.cos(E)
        dE = -f / fp
        E += dE
        if abs(dE) < tol:
            break
    return E


def moment_of_inertia(mass: float, radius: float, length: float = None,
                      shape: Literal["sphere", "cylinder", "rod"] = "sphere") -> float:
    """Moment of inertia of a solid sphere / solid cylinder (about its axis) / thin rod (about its centre)."""
    if shape == "sphere":
        return (2 / 5) * mass * radius**2
    if shape == "cylinder":
        return 0.5 * mass * radius**2
    if shape == "rod":
        if length is None:
            raise ValueError("Length must be provided for rod shape.")
        return (1 / 12) * mass * length**2
    raise ValueError(f"Unknown shape: {shape}")


def random_angular_velocity(max_rotation_rps: float = 1.0, dim: int = 3) -> np.ndarray:
    """Random rotation axis times a rate uniform in [0, max_rotation_rps)."""
    axis = np.random.randn(dim)
    axis /= np.linalg.norm(axis)
    return np.random.uniform(0, max_rotation_rps) * axis


# ---------------------------------------------------------------------------
# Object
# ---------------------------------------------------------------------------
class Object:
    """A massive body: mass, radius, position, velocity (+ spin state, id, name).

    As in the reference (core/physics.py:169-191) a velocity passed to the
    constructor is stored as **float32**; assigning `obj.velocity = array` later
    keeps the assigned dtype (fp64 in the reference's own examples). The engine
    reproduces both behaviours on the device.
    """

    def __init__(self, mass: float, radius: float, velocity: np.ndarray, coordinates: Coordinates = None,
                 moi: float = None, angular_velocity: np.ndarray = None, uuid: str = None,
                 unit_profile: UnitProfile = STANDARD, name: str = None):
        self._engine = None            # set by SimulationEngine when the object is bound
        self._slot = -1                # index in the engine's device arrays
        self._stamp = -1               # engine pull epoch this object's host copy corresponds to
        self._mass = mass
        self._radius = radius
        self._coordinates = coordinates if coordinates else Coordinates.random()
        self._velocity = (velocity.astype(np.float32) if velocity is not None
                          else np.zeros(3).astype(np.float32))
        self.moi = moi if moi is not None else moment_of_inertia(mass, radius, shape="sphere")
        self.angular_velocity = (angular_velocity.astype(np.float32) if angular_velocity is not None
                                 else random_angular_velocity().astype(np.float32))
        self.uuid = uuid if uuid else uuid4().hex
        self.name = name if name is not None else self.uuid[:6]
        self.unit_profile = unit_profile

    # -- lazy-mirror plumbing -------------------------------------------------
    def _before_read(self, velocity=False):
        eng = self._engine
        if eng is not None:
            eng._host_read(self, velocity)

    def _before_write(self):
        eng = self._engine
        if eng is not None:
            eng._host_write(self)

    @property
    def coordinates(self) -> Coordinates:
        self._before_read()
        return self._coordinates

    @coordinates.setter
    def coordinates(self, value: Coordinates):
        self._before_write()
        self._coordinates = value

    @property
    def velocity(self) -> np.ndarray:
        self._before_read(velocity=True)
        return self._velocity

    @velocity.setter
    def velocity(self, value):
        self._before_write()
        self._velocity = value

    @property
    def mass(self):
        return self._mass

    @mass.setter
    def mass(self, value):
        self._before_write()
        self._mass = value

    @property
    def radius(self):
        return self._radius

    @radius.setter
    def radius(self, value):
        self._before_write()
        self._radius = value

    # -- reference API ----------------------------------------------------------
    def position(self) -> np.ndarray:
        return self.coordinates.to_array()

    def to_dict(self):
        c = self.coordinates
        return {
            "mass": self.mass,
            "radius": self.radius,
            "coordinates": {"x": c.x, "y": c.y, "z": c.z},
            "velocity": self.velocity.tolist(),
            "moi": self.moi,
            "angular_velocity": self.angular_velocity.tolist(),
            "uuid": self.uuid,
            "unit_profile": self.unit_profile.name.value,
        }

    @classmethod
    def from_dict(cls, data: dict) -> "Object":
        tag = data.get("unit_profile", "si")
        profile = {"si": STANDARD, "astro": ASTRO}.get(tag)
        c = data["coordinates"]
        return cls(
            mass=data["mass"],
            radius=data["radius"],
            coordinates=Coordinates.from_iterable([c["x"], c["y"], c["z"]]),
            velocity=np.array(data["velocity"]),
            moi=data.get("moi"),
            angular_velocity=np.array(data.get("angular_velocity", [0.0, 0.0, 0.0])),
            uuid=data.get("uuid"),
            unit_profile=profile,
            name=data.get("name"),
        )

    def set_unit_profile(self, unit_profile: UnitProfile):
        self.unit_profile = unit_profile

    def __eq__(self, other):
        return self.uuid == other.uuid

    __hash__ = None

    def __repr__(self):
        return f"Object({self.to_dict()})"

    def lagrangian(self, system: Iterable["Object"]) -> float:
        """Kinetic (translation + spin) minus potential energy of this body in `system`.

        For a body bound to an engine and `system` = that engine's collection the O(N) potential loop runs on the
        device in the same order (orb_body_potential, bit-identical); anything else takes the loop below."""
        T = 0.5 * self.mass * np.linalg.norm(self.velocity) ** 2
        T += 0.5 * self.moi * np.linalg.norm(self.angular_velocity) ** 2
        eng = self._engine
        if eng is not None:
            pe = eng._body_potential(self, system, self.unit_profile.G)
            if pe is not None:
                return T - np.float64(pe)        # the host loop's pe is an np.float64 (T may be float32: NEP 50)
        here = self.coordinates.to_array()
        pe = 0
        for other in system:
            if other is not self:
                r = np.linalg.norm(here - other.coordinates.to_array())
                pe += -self.unit_profile.G * self.mass * other.mass / r
        return T - pe

    def force_vector(self, other: "Object") -> np.ndarray:
        """Newtonian force this body feels from `other` (zero if coincident)."""
        sep = other.coordinates.to_array() - self.coordinates.to_array()
        dist = np.linalg.norm(sep)
        if dist == 0:
            return np.zeros(3)
        magnitude = self.unit_profile.G * self.mass * other.mass / dist**2
        return magnitude * (sep / dist)

    def update(self, acceleration: np.ndarray, dt: float) -> None:
        """Semi-implicit Euler: v += a dt, then x += v dt."""
        self.velocity += acceleration * dt
        self.coordinates = Coordinates.from_iterable(self.coordinates.to_array() + self.velocity * dt)


# ---------------------------------------------------------------------------
# The force operator (narrow, operator-level seam -- SURVEY.md 8b)
# ---------------------------------------------------------------------------
_op_lock = threading.Lock()
_op_cache: dict = {}


def _operator_system(n: int, mode: int) -> "_native.DeviceSystem":
    key = (n, mode, default_device())
    sysm = _op_cache.get(key)
    if sysm is None:
        if len(_op_cache) > 8:
            for old in list(_op_cache.values()):
                old.close()
            _op_cache.clear()
        sysm = _op_cache[key] = _native.DeviceSystem(n, default_device(), mode)
    return sysm


def pairwise_accelerations(objects: "list[Object]", eps: float = 0.0, unit_profile: UnitProfile = STANDARD,
                           mode: "str | None" = None):
    """All-pairs softened Newtonian accelerations and total potential, on the GPU.

    Same contract as the reference (core/physics.py:125-159):
        returns (dict uuid -> np.ndarray(3) float64, U)
    In 'faithful' mode (default for n <= 4096) the result is bit-identical to the
    reference's Python loop.
    """
    n = len(objects)
    if n == 0:
        return {}, 0.0
    pos = np.array([o.position() for o in objects], dtype=np.float64).reshape(n, 3)
    m = np.array([float(o.mass) for o in objects], dtype=np.float64)
    zeros = np.zeros(n)
    with _op_lock:
        dev = _operator_system(n, select_mode(n, mode))
        dev.set_params(1.0, float(eps), float(unit_profile.G))
        dev.upload(pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy(), zeros, zeros, zeros, m, zeros)
        dev.accel()
        a = dev.download_acc()
        U = dev.potential() if n > 1 else 0.0
    acc = {o.uuid: np.array([a[0, i], a[1, i], a[2, i]]) for i, o in enumerate(objects)}
    return acc, np.float64(U) if n > 1 else 0.0


# ---------------------------------------------------------------------------
# Collisions (host-side resolution, reference semantics)
# ---------------------------------------------------------------------------
def fragmentation_probability(obj1: Object, obj2: Object) -> float:
    """Logistic in (collision energy / threshold energy); threshold ~ 1e3 J per kg of combined mass."""
    v_rel = np.linalg.norm(obj1.velocity - obj2.velocity)
    mu = (obj1.mass * obj2.mass) / (obj1.mass + obj2.mass)
    E_coll = 0.5 * mu * v_rel**2
    E_thresh = 0.5 * (obj1.mass + obj2.mass) * 1e3
    k = 5
    return 1 / (1 + np.exp(-k * (E_coll / E_thresh - 1)))


def resolve_collision(obj1: Object, obj2: Object, collection: "ObjectCollection"):
    """Absorb the lighter body if the mass ratio exceeds 10, else possibly fragment (remove) both."""
    heavy, light = (obj1, obj2) if obj1.mass > obj2.mass else (obj2, obj1)
    if max(obj1.mass, obj2.mass) / min(obj1.mass, obj2.mass) > 10:
        heavy.mass += light.mass
        heavy.radius = (heavy.radius**3 + light.radius**3) ** (1 / 3)
        collection.remove(light)
    elif np.random.rand() < fragmentation_probability(obj1, obj2):
        collection.remove(obj1)
        collection.remove(obj2)


def collide_spheres(obj1: Object, obj2: Object, restitution: float = 1.0):
    """Impulse along the line of centres with restitution, then push the pair out of overlap.

    Behaviour of the reference's collide_spheres (core/physics.py:391-422): no-op
    when coincident or separating; velocities are updated in place (dtype kept),
    positions are replaced by new Coordinates.
    """
    r1, r2 = obj1.position(), obj2.position()
    normal = r1 - r2
    dist = np.linalg.norm(normal)
    if dist == 0:
        return
    normal /= dist
    m1, m2 = obj1.mass, obj2.mass
    closing = np.dot(obj1.velocity - obj2.velocity, normal)
    if closing >= 0:
        return
    inv1, inv2 = 1.0 / m1, 1.0 / m2
    e = float(np.clip(restitution, 0.0, 1.0))
    j = -(1 + e) * closing / (inv1 + inv2)
    impulse = j * normal
    obj1.velocity += impulse / m1
    obj2.velocity -= impulse / m2
    overlap = obj1.radius + obj2.radius - dist
    if overlap > 0:
        corr = overlap / (inv1 + inv2)
        obj1.coordinates = Coordinates.from_iterable(r1 + normal * (corr / m1))
        obj2.coordinates = Coordinates.from_iterable(r2 - normal * (corr / m2))


def set_circular_orbit(primary: Object, secondary: Object, plane_normal=np.array([0.0, 0.0, 1.0]),
                       unit_profile: UnitProfile = STANDARD):
    """Give `secondary` the circular two-body speed about `primary`; zero total momentum.

    Assigns fp64 velocity arrays (so both bodies leave fp32-velocity mode).
    """
    sep = secondary.position() - primary.position()
    R = np.linalg.norm(sep)
    if R == 0:
        raise ValueError("Bodies at same position.")
    t = np.cross(plane_normal / np.linalg.norm(plane_normal), sep / R)
    if np.linalg.norm(t) < 1e-12:
        t = np.cross(np.array([0.0, 1.0, 0.0]), sep / R)
    t /= np.linalg.norm(t)
    v_mag = np.sqrt(unit_profile.G * (primary.mass + secondary.mass) / R)
    v2 = v_mag * t
    v1 = -(secondary.mass / primary.mass) * v2
    primary.velocity = v1
    secondary.velocity = v2


# ---------------------------------------------------------------------------
# ObjectCollection
# ---------------------------------------------------------------------------
class ObjectCollection(object):
    """An ordered list of Objects."""

    def __init__(self, objects: "list[Object]"):
        self.objects = objects

    def to_dict(self):
        return [o.to_dict() for o in self.objects]

    @classmethod
    def from_dict(cls, data: "list[dict]") -> "ObjectCollection":
        return cls([Object.from_dict(d) for d in data])

    def __len__(self):
        return len(self.objects)

    def __getitem__(self, index):
        return self.objects[index]

    def __iter__(self):
        return iter(self.objects)

    def force_vector_map(self):
        """uuid -> net acceleration from `Object.force_vector` over all other bodies (host loop, legacy API)."""
        out = {o.uuid: np.zeros(3) for o in self.objects}
        for i, o in enumerate(self.objects):
            for j, other in enumerate(self.objects):
                if i != j:
                    out[o.uuid] += o.force_vector(other) / o.mass
        return out

    def extend(self, new_objects: Iterable[Object]) -> None:
        self.objects.extend(new_objects)

    def append(self, new_object: Object) -> None:
        self.objects.append(new_object)

    def pop(self, index: int = -1) -> Object:
        return self.objects.pop(index)

    def remove(self, obj: Object) -> None:
        self.objects.remove(obj)

    # -- contacts ---------------------------------------------------------------
    @staticmethod
    def _touching(oi: Object, oj: Object) -> bool:
        return np.linalg.norm(oi.position() - oj.position()) <= (oi.radius + oj.radius)

    def _merge(self, oi: Object, oj: Object):
        m_new = oi.mass + oj.mass
        v_new = (oi.mass * oi.velocity + oj.mass * oj.velocity) / m_new
        r_new = (oi.mass * oi.position() + oj.mass * oj.position()) / m_new
        R_new = (oi.radius**3 + oj.radius**3) ** (1 / 3)
        oi.mass = m_new
        oi.velocity = v_new
        oi.coordinates = Coordinates.from_iterable(r_new)
        oi.radius = R_new

    def handle_collisions(self, restitution: float = 1.0, merge_on_capture: bool = False):
        """Sequential in-place sweep over pairs i<j (reference core/physics.py:510-535).

        Host implementation of the full sweep. The engine does not call this per
        step: it detects overlaps on the device and resolves only the affected
        pairs through :meth:`resolve_contacts`, which visits them in the same order.
        """
        objs = self.objects
        n = len(objs)
        absorbed = []
        for i in range(n):
            for j in range(i + 1, n):
                if self._touching(objs[i], objs[j]):
                    if merge_on_capture:
                        self._merge(objs[i], objs[j])
                        absorbed.append(objs[j])
                    else:
                        collide_spheres(objs[i], objs[j], restitution=restitution)
        for o in absorbed:
            self.remove(o)

    def resolve_contacts(self, flagged_pairs, restitution: float = 1.0) -> int:
        """Resolve device-flagged contacts exactly as the full lexicographic sweep would.

        `flagged_pairs`: pairs (i<j) that overlapped at the positions the device saw.
        Pairs are visited in lexicographic order and re-tested against the *current*
        host positions (an earlier contact in the sweep may have moved a body);
        whenever a contact moves a body, every later pair involving it is queued
        for testing too.  Equivalent to `handle_collisions(restitution)` because a
        pair that was not flagged and whose bodies did not move cannot be touching.
        Returns the number of touching pairs processed.
        """
        objs = self.objects
        n = len(objs)
        heap = [(int(i), int(j)) for i, j in flagged_pairs]
        heapq.heapify(heap)
        queued = set(heap)
        radius = np.array([float(o.radius) for o in objs])
        hits = 0
        while heap:
            i, j = heapq.heappop(heap)
            oi, oj = objs[i], objs[j]
            if not self._touching(oi, oj):
                continue
            hits += 1
            before = (oi._coordinates, oj._coordinates)
            collide_spheres(oi, oj, restitution=restitution)
            moved = [k for k, o, b in ((i, oi, before[0]), (j, oj, before[1])) if o._coordinates is not b]
            if not moved:
                continue
            pos = np.array([[o._coordinates.x, o._coordinates.y, o._coordinates.z] for o in objs], dtype=np.float64)
            for k in moved:
                d = np.linalg.norm(pos - pos[k], axis=1)
                near = np.nonzero(d <= (radius + radius[k]) * (1 + 1e-9) + 1e-300)[0]
                for c in near:
                    c = int(c)
                    if c == k:
                        continue
                    pair = (min(k, c), max(k, c))
                    if pair > (i, j) and pair not in queued:
                        queued.add(pair)
                        heapq.heappush(heap, pair)
        return hits
