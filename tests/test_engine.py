"""SimulationEngine parity: the reference's own engine outputs (tests/golden/*.npz) vs this engine.

Every test runs twice: `backend=fake` (CPU: host logic over the oracle stand-in,
part of `-m "not gpu"`) and `backend=cuda` (`-m gpu`: the real sm_100a kernels
through the C ABI).  Bar: bit-exact positions / velocities / accelerations in
faithful mode (reference core/engine.py:65-97, core/physics.py:125-159,391-422).
"""
import json
import threading

import numpy as np
import pytest

from tests.conftest import make_objects


def bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def assert_bits(a, b, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
    if not same.all():
        rel = np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
        raise AssertionError(f"{what}: {np.count_nonzero(~same)} of {same.size} values differ (max rel {rel:.3e})")


def state_of(engine):
    objs = engine.objects.objects
    pos = np.array([[o.coordinates.x, o.coordinates.y, o.coordinates.z] for o in objs], dtype=np.float64)
    vel = np.array([np.asarray(o.velocity, dtype=np.float64) for o in objs])
    acc = np.array([engine.acc[o.uuid] for o in objs])
    return pos, vel, acc


def build_engine(g, **kw):
    from core.engine import SimulationEngine
    from core.physics import ObjectCollection
    objs = make_objects(g)
    kw.setdefault("cache", False)
    kw.setdefault("max_hist", None)
    return SimulationEngine(ObjectCollection(objs), dt=float(g["dt"]), softening=float(g["eps"]),
                            restitution=float(g["restitution"]), **kw)


def check_against_golden(g, eng, use_run=True, energy_tol=1e-12):
    p, v, a = state_of(eng)
    assert_bits(v, g["vel_0"], "initial velocity")
    assert_bits(a, g["acc_0"], "initial acc")
    assert_bits(eng.last_potential, g["U_0"], "initial U")
    assert abs(eng.total_energy() - float(g["E_0"])) <= energy_tol * abs(float(g["E_0"]))
    done = 0
    for s in g["steps"]:
        s = int(s)
        if use_run:
            eng.run(s - done)
        else:
            for _ in range(s - done):
                eng.step()
        done = s
        p, v, a = state_of(eng)
        assert_bits(p, g[f"pos_{s}"], f"pos @ {s}")
        assert_bits(v, g[f"vel_{s}"], f"vel @ {s}")
        assert_bits(a, g[f"acc_{s}"], f"acc @ {s}")
        assert_bits(eng.last_potential, g[f"U_{s}"], f"U @ {s}")
        E, L = eng.total_energy(), eng.angular_momentum()
        assert abs(E - float(g[f"E_{s}"])) <= energy_tol * abs(float(g[f"E_{s}"])), f"E @ {s}"
        assert np.linalg.norm(L - g[f"L_{s}"]) <= energy_tol * np.linalg.norm(g[f"L_{s}"]) + 1e-300, f"L @ {s}"
    assert eng.step_idx == done
    assert eng.time_elapsed == pytest.approx(done * float(g["dt"]), rel=1e-12)


@pytest.mark.parametrize("name", ["solar15_f32", "solar15_f64", "solar9_f32", "solar9_f64", "solar26_f32", "solar26_f64"])
def test_solar_system_matches_reference(golden, backend, name):
    """BASELINE configs[0]: Sun + planets via the engine step loop, both velocity-dtype modes."""
    g = golden(name)
    if backend == "fake":        # keep the CPU suite short: the oracle itself is pinned to 10k steps elsewhere
        g = {k: g[k] for k in g.files}
        g["steps"] = np.array([s for s in g["steps"] if s <= 100])
    check_against_golden(g, build_engine(g))


@pytest.mark.parametrize("name", ["mixed12"])
def test_mixed_velocity_dtypes(golden, backend, name):
    check_against_golden(golden(name), build_engine(golden(name)), use_run=False)


@pytest.mark.parametrize("name", ["coll_hit_f32_e1", "coll_hit_f64_e1", "coll_hit_f32_e05", "coll_hit_f64_e05",
                                  "coll_dense_f32", "coll_dense_mixed"])
@pytest.mark.parametrize("use_run", [True, False])
@pytest.mark.parametrize("contacts", ["device", "host"])
def test_collisions_match_reference(golden, backend, name, use_run, contacts):
    """Contacts detected in the force pass and resolved on the device (default) or on the host
    == the reference's sequential in-place sweep, bit for bit (incl. U, E, L of the contact steps)."""
    g = golden(name)
    check_against_golden(g, build_engine(g, contacts=contacts), use_run=use_run)


def test_two_body_example(golden, backend):
    from core.engine import SimulationEngine
    from core.physics import Coordinates, Object, ObjectCollection, set_circular_orbit
    g = golden("two_body")
    b1 = Object(5.972e24, 6.371e6, velocity=np.zeros(3), coordinates=Coordinates(0, 0, 0))
    b2 = Object(7.348e22, 1.737e6, velocity=np.zeros(3), coordinates=Coordinates(384400e3, 0, 0))
    set_circular_orbit(primary=b1, secondary=b2)
    assert b1.velocity.dtype == np.float64
    assert_bits(np.array([b1.velocity, b2.velocity]), g["v0"])
    eng = SimulationEngine(ObjectCollection([b1, b2]), dt=3600.0, softening=1e3, cache=False, max_hist=None)
    for _ in range(500):
        eng.step()
    p, v, a = state_of(eng)
    assert_bits(p, g["pos_500"]); assert_bits(v, g["vel_500"]); assert_bits(a, g["acc_500"])
    assert_bits(np.array(eng.history[b2.uuid]), g["hist_b2"], "history trail")
    assert eng.total_energy() == pytest.approx(float(g["E_500"]), rel=1e-13)


def test_three_body_example(golden, backend, capsys):
    from core import examples
    g = golden("three_body")
    import core.engine
    real = core.engine.SimulationEngine

    class NoCache(real):
        def __init__(self, *a, **k):
            k.setdefault("cache", False)
            k.setdefault("max_hist", None)
            super().__init__(*a, **k)

    examples.SimulationEngine = NoCache
    try:
        eng = examples.three_body_equilateral(steps=1000)
    finally:
        examples.SimulationEngine = real
    p, v, a = state_of(eng)
    assert eng.objects[0].velocity.dtype == np.float32
    assert_bits(p, g["pos_1000"]); assert_bits(v, g["vel_1000"]); assert_bits(a, g["acc_1000"])
    out = capsys.readouterr().out
    assert "step 0: ΔE=" in out and "step 500: ΔE=" in out and "step 1000" not in out


def test_history_length_quirks(golden, backend):
    """max_hist=-1 keeps one point; N keeps N; None is unbounded."""
    g = golden("solar9_f64")
    for max_hist, want in ((-1, 1), (0, 1), (1, 1), (5, 5), (None, 13)):
        eng = build_engine(g, max_hist=max_hist)
        u = eng.objects[3].uuid
        assert len(eng.history[u]) == 1
        first = eng.history[u][0]
        assert first == [float(g["in_x"][3]), float(g["in_y"][3]), float(g["in_z"][3])]
        for _ in range(7):
            eng.step()
        eng.run(5)
        h = eng.history[u]
        assert len(h) == want, (max_hist, len(h))
        assert h[-1] == eng.objects[3].position().tolist()
        assert isinstance(h[-1][0], float)
        if max_hist is None:
            assert h[0] == first
        assert set(eng.history.keys()) == {o.uuid for o in eng.objects}
        nh = eng.named_history(limit=2)
        assert list(nh) == [o.name for o in eng.objects]
        assert nh[eng.objects[3].name] == h[-2:]


def test_history_drains_when_ring_is_small(golden, backend, monkeypatch):
    import core.engine
    g = golden("solar9_f64")
    ref = build_engine(g, max_hist=None)
    ref.run(40)
    monkeypatch.setattr(core.engine, "_RING_BYTES", 24 * 9 * 7)      # ring of 7 snapshots
    eng = build_engine(g, max_hist=None)
    eng.run(25)
    for _ in range(15):
        eng.step()
    u = eng.objects[2].uuid
    assert eng._drain_mode and eng._ring_cap == 7
    assert_bits(np.array(eng.history[u]), np.array(ref.history[ref.objects[2].uuid]))
    eng2 = build_engine(g, max_hist=20)                               # bounded but larger than the ring
    eng2.run(40)
    assert_bits(np.array(eng2.history[eng2.objects[2].uuid]), np.array(ref.history[ref.objects[2].uuid])[-20:])


def test_jsonl_cache_frames(golden, backend, tmp_path):
    """engine.py:48-57,94-97: frames at step_idx % n == 0; frame time lags positions by one dt."""
    from core.engine import SimulationEngine
    g = golden("solar9_f32")
    with pytest.raises(ValueError):
        build_engine(g, cache=True, cache_fp=str(tmp_path / "bad.json"))
    fp = str(tmp_path / "frames.jsonl")
    eng = build_engine(g, cache=True, cache_fp=fp, cache_every_n=4, max_hist=-1)
    assert eng.cache_every_n == 4
    eng.step()
    eng.run(9)
    lines = [json.loads(l) for l in open(fp)]
    assert len(lines) == 3                                         # steps 0, 4, 8
    assert [l["time_elapsed"] for l in lines] == [0.0, 4 * eng.dt, 8 * eng.dt]
    assert set(lines[0]) == {"time_elapsed", "objects", "history"}
    o0 = lines[0]["objects"][1]
    assert set(o0) == {"mass", "radius", "coordinates", "velocity", "moi", "angular_velocity", "uuid", "unit_profile"}
    assert o0["unit_profile"] == "si"
    assert_bits([o0["coordinates"][k] for k in "xyz"], golden("solar9_f32")["pos_1"][1])
    assert lines[0]["history"]["b1"] == [[o0["coordinates"][k] for k in "xyz"]]
    assert build_engine(g, cache=False).cache_every_n == 0
    assert SimulationEngine.__init__.__defaults__ == (1.0, 0.0, 1.0, -1, True, "history.jsonl", 300)


def test_host_mutation_between_steps(golden, backend, orc):
    """Writes to a bound Object reach the device; accelerations stay stale, as in the reference."""
    from oracle.c_oracle import State
    g = golden("solar9_f64")
    eng = build_engine(g)
    st = State(orc, g["in_x"], g["in_y"], g["in_z"], g["in_vx"], g["in_vy"], g["in_vz"], g["in_m"],
               g["in_radius"], 0, float(g["dt"]), float(g["eps"]))
    eng.run(3); st.step(3)
    # (a) rebinding velocity (fp64), (b) in-place edit of an exposed array, (c) new coordinates, (d) mass change
    eng.objects[4].velocity = eng.objects[4].velocity * 1.01
    st.vx[4], st.vy[4], st.vz[4] = (np.array([st.vx[4], st.vy[4], st.vz[4]]) * 1.01)
    eng.objects[5].velocity[1] += 12.5
    st.vy[5] += 12.5
    from core.physics import Coordinates
    c = eng.objects[6].coordinates
    eng.objects[6].coordinates = Coordinates(c.x + 1e6, c.y, c.z)
    st.x[6] += 1e6
    eng.objects[2].mass = eng.objects[2].mass * 2
    st.m[2] *= 2
    eng.run(2); st.step(2)
    p, v, a = state_of(eng)
    assert_bits(p, st.pos); assert_bits(v, st.vel); assert_bits(a, st.acc)
    # reading without writing must not perturb anything
    _ = [o.position() for o in eng.objects]
    eng.step(); st.step(1)
    assert_bits(state_of(eng)[0], st.pos)
    # fp64 -> fp32 switch by assigning a float32 array
    eng.objects[1].velocity = eng.objects[1].velocity.astype(np.float32)
    st.vf32[1] = 1
    for arr in (st.vx, st.vy, st.vz):
        arr[1] = np.float64(np.float32(arr[1]))
    eng.run(4); st.step(4)
    assert_bits(state_of(eng)[0], st.pos); assert_bits(state_of(eng)[1], st.vel)
    assert eng.objects[1].velocity.dtype == np.float32


def test_velocity_array_kept_across_steps_is_live_like_the_reference(golden, backend, orc):
    """`v = obj.velocity; engine.step(); v += dv; engine.step()`: the reference kicks that very array in place
    (engine.py:70,82), so `v` tracks the state and the in-place edit is honoured."""
    from oracle.c_oracle import State
    g = golden("solar9_f64")
    eng = build_engine(g)
    st = State(orc, g["in_x"], g["in_y"], g["in_z"], g["in_vx"], g["in_vy"], g["in_vz"], g["in_m"],
               g["in_radius"], 0, float(g["dt"]), float(g["eps"]))
    v = eng.objects[3].velocity                  # handed out once, before any step
    eng.step(); st.step(1)
    assert_bits(v, [st.vx[3], st.vy[3], st.vz[3]], "alias follows the step")
    v += np.array([3.0, -2.0, 0.5])              # never touches the property again
    st.vx[3] += 3.0; st.vy[3] += -2.0; st.vz[3] += 0.5
    eng.run(2); st.step(2)
    assert_bits(v, [st.vx[3], st.vy[3], st.vz[3]], "in-place edit through the alias reached the device")
    p, vel, a = state_of(eng)
    assert_bits(p, st.pos); assert_bits(vel, st.vel); assert_bits(a, st.acc)
    assert eng.objects[3].velocity is v


def test_objects_materialise_lazily(golden, backend):
    """Reading one body after a step pulls the state once and builds that body only."""
    g = golden("solar15_f64")
    eng = build_engine(g)
    eng.run(10)
    epoch = eng._pull_epoch
    p7 = eng.objects[7].position()
    assert eng._pull_epoch == epoch + 1
    stale = [o for o in eng._bound if o._stamp != eng._pull_epoch]
    assert len(stale) == len(eng._bound) - 1, "only the body that was read is rebuilt"
    assert_bits(p7, g["pos_10"][7])
    assert eng._pull_epoch == epoch + 1, "further reads reuse the snapshot"
    assert_bits(state_of(eng)[0], g["pos_10"])
    eng.run(90)
    assert_bits(state_of(eng)[0], g["pos_100"])


def test_late_added_object_raises_keyerror_like_reference(golden, backend):
    from core.physics import Coordinates, Object
    g = golden("solar9_f64")
    eng = build_engine(g)
    eng.step()
    eng.objects.append(Object(1e20, 1e3, velocity=np.zeros(3), coordinates=Coordinates(1e12, 0, 0)))
    with pytest.raises(KeyError):
        eng.step()


def test_removed_object_continues(golden, backend, orc):
    from oracle.c_oracle import State
    g = golden("solar9_f64")
    eng = build_engine(g)
    eng.run(2)
    gone = eng.objects.pop(8)
    p, v, a = state_of(eng)
    eng.run(3)
    st = State(orc, p[:, 0], p[:, 1], p[:, 2], v[:, 0], v[:, 1], v[:, 2], g["in_m"][:8], g["in_radius"][:8], 0,
               float(g["dt"]), float(g["eps"]))
    st.ax, st.ay, st.az = (np.ascontiguousarray(a[:, k]) for k in range(3))   # stale acc carried over
    st.step(3)
    assert_bits(state_of(eng)[0], st.pos)
    assert len(eng.history[eng.objects[0].uuid]) == 6
    assert len(eng.history[gone.uuid]) == 3          # the reference's dict keeps the stale trail too


def test_reader_thread_never_sees_torn_state(golden, backend):
    """app/app.py:104-115: one stepping thread, readers without a lock."""
    g = golden("solar15_f64")
    eng = build_engine(g, max_hist=50)
    stop = threading.Event()
    errors = []

    def reader():
        while not stop.is_set():
            try:
                for o in eng.objects:
                    p = o.position()
                    assert p.shape == (3,) and np.isfinite(p).all()
                eng.named_history(limit=5)
                _ = eng.time_elapsed
            except Exception as exc:      # pragma: no cover
                errors.append(exc)
                return

    threads = [threading.Thread(target=reader) for _ in range(3)]
    for t in threads:
        t.start()
    for _ in range(60):
        eng.step()
    stop.set()
    for t in threads:
        t.join()
    assert not errors
    assert_bits(state_of(eng)[0], _replay(golden, "solar15_f64", 60))


def _replay(golden, name, steps):
    from oracle import load_c_oracle
    from oracle.c_oracle import State
    g = golden(name)
    st = State(load_c_oracle(), g["in_x"], g["in_y"], g["in_z"], g["in_vx"], g["in_vy"], g["in_vz"], g["in_m"],
               g["in_radius"], (~g["f64_velocity"]).astype(np.uint8), float(g["dt"]), float(g["eps"]))
    st.step(steps)
    return st.pos


def test_engine_accepts_ad_hoc_attributes(golden, backend):
    eng = build_engine(golden("solar9_f32"))
    eng.body_map = {"x": 1}
    eng.sim_epoch_jd = 2451545.0
    assert eng.body_map["x"] == 1
    assert list(eng.objects)[0].name == "b0"


def test_run_simulation_prints_like_reference(golden, backend, capsys):
    from core.engine import run_simulation
    eng = build_engine(golden("solar9_f64"))
    run_simulation(eng, steps=25, print_every=10)
    out = capsys.readouterr().out.strip().splitlines()
    assert [l.split(":")[0] for l in out] == ["step 0", "step 10", "step 20"]
    assert eng.step_idx == 25
    assert_bits(state_of(eng)[0], _replay(golden, "solar9_f64", 25))


@pytest.mark.parametrize("name", ["solar9_f32", "solar9_f64", "mixed12"])
def test_resume_from_jsonl_frame_is_bit_identical(golden, backend, tmp_path, name):
    """loader for the reference's write-only JSONL cache; resumed run == uninterrupted run."""
    from core.engine import SimulationEngine, load_frames
    g = golden(name)
    fp = str(tmp_path / "run.jsonl")
    full = build_engine(g, cache=True, cache_fp=fp, cache_every_n=5)
    full.run(23)                                             # frames after steps 0, 5, 10, 15, 20
    frames = load_frames(fp)
    assert len(frames) == 5 and frames[2]["time_elapsed"] == 10 * full.dt
    res = SimulationEngine.resume(fp, index=2, dt=float(g["dt"]), softening=float(g["eps"]),
                                  restitution=float(g["restitution"]), cache=False, max_hist=None)
    assert res.step_idx == 11 and res.time_elapsed == 11 * res.dt
    assert [o.name for o in res.objects] == [o.name for o in full.objects]
    assert [o.velocity.dtype for o in res.objects] == [o.velocity.dtype for o in full.objects]
    res.run(12)                                              # steps 11..22 -> same point as `full`
    assert res.step_idx == full.step_idx
    pf, vf, af = state_of(full)
    pr, vr, ar = state_of(res)
    assert_bits(pr, pf); assert_bits(vr, vf); assert_bits(ar, af)


def test_degenerate_collections(backend):
    """Empty and single-body collections behave like the reference: no pairs, zero acceleration, U = 0."""
    from core.engine import SimulationEngine, run_simulation
    from core.physics import Coordinates, Object, ObjectCollection
    empty = SimulationEngine(ObjectCollection([]), dt=10.0, cache=False)
    empty.step(); empty.run(3)
    assert empty.step_idx == 4 and empty.total_energy() == 0.0 and len(empty.history) == 0
    assert empty.angular_momentum().tolist() == [0.0, 0.0, 0.0] and empty.acc == {}
    one = Object(5.0, 1.0, np.array([1.0, 2.0, 3.0]), Coordinates(10.0, 20.0, 30.0))
    eng = SimulationEngine(ObjectCollection([one]), dt=2.0, cache=False, max_hist=None)
    assert eng.acc[one.uuid].tolist() == [0.0, 0.0, 0.0] and eng.last_potential == 0.0
    eng.run(5)
    assert one.position().tolist() == [20.0, 40.0, 60.0]              # free drift with the fp32 velocity
    assert eng.total_energy() == 0.5 * 5.0 * float(one.velocity @ one.velocity)
    assert len(eng.history[one.uuid]) == 6
    run_simulation(eng, steps=3, print_every=1)
    assert eng.step_idx == 8


def test_massless_and_coincident_bodies(backend, orc):
    """Test particles (m = 0) and softened coincident bodies follow the reference arithmetic."""
    from oracle.c_oracle import State
    from core.engine import SimulationEngine
    from core.physics import Coordinates, Object, ObjectCollection
    x = np.array([0.0, 1.0e9, 1.0e9, -3.0e9, 2.0e9]); y = np.array([0.0, 0.0, 0.0, 1.0e9, 5.0e8]); z = np.zeros(5)
    v = np.array([[0, 0, 0], [0, 3e3, 0], [0, 3e3, 10.0], [1e3, 0, 0], [0, -2e3, 5.0]], dtype=np.float64)
    m = np.array([1e27, 0.0, 0.0, 5e24, 1e20])
    objs = [Object(float(m[i]), 0.0, v[i], Coordinates(float(x[i]), float(y[i]), float(z[i])), angular_velocity=np.zeros(3))
            for i in range(5)]
    for o, vi in zip(objs, v):
        o.velocity = vi.copy()
    eng = SimulationEngine(ObjectCollection(objs), dt=100.0, softening=1e6, cache=False, max_hist=None)
    st = State(orc, x, y, z, v[:, 0], v[:, 1], v[:, 2], m, np.zeros(5), 0, 100.0, 1e6)
    eng.run(50); st.step(50)
    pos = np.array([o.position() for o in objs]); vel = np.array([o.velocity for o in objs])
    assert np.array_equal(pos, st.pos) and np.array_equal(vel, st.vel)
    assert np.isfinite(pos).all()


@pytest.mark.parametrize("name", ["solar15_f32", "coll_dense_mixed"])
@pytest.mark.parametrize("contacts", ["device", "host"])
def test_deferred_steps_equal_immediate_steps(backend, golden, monkeypatch, name, contacts):
    """step() calls that nothing observes are deferred and run as ONE device stretch when something looks
    (the reference's driver loops call engine.step() once per step, core/examples.py:198-217).  Same steps in the
    same order: every observable -- positions, velocities, accelerations, history, clock, JSONL-free run -- has the
    same bits as with ORBITAL_B200_DEFER=0, across a parameter change, a host-side write and (coll_dense_mixed)
    contacts resolved on the device or by the host replay; and the device is entered a handful of times, not 75."""
    import core.engine as ce
    from core.physics import Coordinates
    g = golden(name)

    def drive(defer):
        monkeypatch.setattr(ce, "_DEFER", defer)
        eng = build_engine(g, contacts=contacts)
        objs = eng.objects.objects
        calls = []
        real = eng._dev.step
        def counted(k):
            r = real(k)
            calls.append(int(r[0]))                         # steps completed by this entry into the device
            return r
        eng._dev.step = counted
        seen = []
        look = lambda: seen.append(np.array([[o.coordinates.x, o.coordinates.y, o.coordinates.z] for o in objs]))
        for _ in range(40):
            eng.step()
        look()                                              # a read: runs the 40 steps
        eng.dt = eng.dt * 0.5                               # parameter change: later steps only
        for _ in range(25):
            eng.step()
        c = objs[2].coordinates                             # read + host-side write of one body
        objs[2].coordinates = Coordinates(c.x * (1 + 1e-9), c.y, c.z)
        for _ in range(10):
            eng.step()
        seen.append(np.array([eng.acc[o.uuid] for o in objs]))
        look()
        seen.append(np.array([eng._peek_velocity(o) for o in objs], dtype=np.float64))
        seen.append(np.array([eng.history[o.uuid] for o in objs]))
        seen.append(np.array([eng.last_potential, eng.total_energy(), eng.time_elapsed, eng.step_idx], dtype=np.float64))
        eng.close()
        return seen, calls

    want, calls_now = drive(False)
    got, calls_deferred = drive(True)
    for k, (a, b) in enumerate(zip(got, want)):
        assert_bits(a, b, f"observable {k}")
    assert sum(calls_now) == sum(calls_deferred) == 75
    assert len(calls_now) >= 75
    if contacts == "device" or name == "solar15_f32":
        assert calls_deferred == [40, 1, 24, 1, 9], calls_deferred
    else:
        assert len(calls_deferred) < len(calls_now)


def test_deferred_steps_cap_and_membership(backend, golden, monkeypatch):
    """Deferred steps pile up to _DEFER_MAX at most; a body added between steps still raises KeyError in step()
    like the reference (core/engine.py:70), after the steps that were asked for before it have run."""
    import core.engine as ce
    from core.physics import Coordinates, Object
    g = golden("solar9_f32")
    monkeypatch.setattr(ce, "_DEFER_MAX", 16)
    eng = build_engine(g)
    calls = []
    real = eng._dev.step
    eng._dev.step = lambda k: (calls.append(int(k)), real(k))[1]
    for _ in range(40):
        eng.step()
    assert calls == [16, 16] and eng._pending == 8
    eng.objects.append(Object(mass=1.0, radius=0.0, velocity=np.zeros(3), coordinates=Coordinates(1e13, 0.0, 0.0),
                              angular_velocity=np.zeros(3), name="late"))
    with pytest.raises(KeyError):
        eng.step()
    assert calls == [16, 16, 8]
    eng.close()


def test_velocity_array_is_tracked_only_while_it_is_held(backend, golden):
    """An ndarray handed out by `obj.velocity` is the object's live array (the reference kicks it in place,
    core/engine.py:70): it follows every step and an in-place edit reaches the device.  Once the caller drops it the
    object goes back to the lazy / deferred path -- reading a velocity once does not cost a download per step for
    the rest of the run."""
    g = golden("solar9_f64")
    ref = build_engine(g)
    eng = build_engine(g)
    calls = []
    real = eng._dev.step
    eng._dev.step = lambda k: (calls.append(int(k)), real(k))[1]
    v = eng.objects.objects[3].velocity
    rv = ref.objects.objects[3].velocity
    for _ in range(3):
        eng.step(); ref.step()
        assert_bits(v, rv, "held array follows the state")
    v += 1.0e-3                                            # in-place edit through the kept reference
    rv += 1.0e-3
    del v
    eng.step(); ref.step()                                 # uploads the edit, then notices nobody holds the array
    assert not eng._watch and calls == [1, 1, 1, 1]
    for _ in range(20):
        eng.step(); ref.step()
    assert calls == [1, 1, 1, 1]                           # deferred again ...
    assert_bits(eng.objects.objects[3].coordinates.x, ref.objects.objects[3].coordinates.x, "after release")
    assert calls == [1, 1, 1, 1, 20]                       # ... and run as one stretch by the read
    p1, v1, a1 = state_of(eng)
    p2, v2, a2 = state_of(ref)
    assert_bits(p1, p2, "pos"); assert_bits(v1, v2, "vel"); assert_bits(a1, a2, "acc")
    eng.close(); ref.close()


def test_lagrangian_potential_on_the_device(backend, golden):
    """Object.lagrangian(engine.objects) (reference core/physics.py:243-283): the O(N) potential loop of an
    engine-bound body runs on the device in the reference's order -- same bits as the host loop, before and after
    steps, after a host-side edit; any other `system` argument still takes the host loop."""
    from core.physics import Coordinates, Object
    g = golden("solar26_f32")
    eng = build_engine(g)
    objs = eng.objects.objects

    def host_value(o):                                   # the reference expression, on plain unbound copies
        clones = [Object(mass=b.mass, radius=b.radius, velocity=np.array(b.velocity, copy=True),
                         coordinates=Coordinates(b.coordinates.x, b.coordinates.y, b.coordinates.z),
                         angular_velocity=np.zeros(3), name=b.name) for b in objs]
        return clones[objs.index(o)].lagrangian(clones)

    calls = []
    real = eng._dev.body_potential
    eng._dev.body_potential = lambda i, G: (calls.append(i), real(i, G))[1]
    for k in (0, 3, 25):                 # (body 0: the reference's float32 spin term overflows to nan -- kept)
        assert_bits(objs[k].lagrangian(eng.objects), host_value(objs[k]), f"L[{k}] at t=0")
    assert np.isfinite(objs[3].lagrangian(eng.objects)) and np.isfinite(objs[25].lagrangian(eng.objects))
    for _ in range(7):
        eng.step()
    assert_bits(objs[9].lagrangian(eng.objects.objects), host_value(objs[9]), "L[9] after 7 steps")
    c = objs[4].coordinates
    objs[4].coordinates = Coordinates(c.x * 1.25, c.y, c.z)           # host-side edit reaches the device first
    assert_bits(objs[9].lagrangian(eng.objects), host_value(objs[9]), "L[9] after an edit")
    assert np.isfinite(objs[9].lagrangian(eng.objects))
    assert calls == [0, 3, 25, 3, 25, 9, 9, 9]
    assert_bits(objs[9].lagrangian(list(objs)), host_value(objs[9]), "host loop for a foreign iterable")
    assert_bits(objs[9].lagrangian(objs[:12]), objs[9].lagrangian(iter(objs[:12])), "subset: host loop")
    assert calls == [0, 3, 25, 3, 25, 9, 9, 9]
    eng.close()
