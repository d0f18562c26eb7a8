"""Drop-in proof: the reference's OWN `app/app.py` and `core/examples.py` run unchanged on this `core` package.

The reference's files are loaded from its checkout: /root/reference in the build container, or the git-ignored mirror
baseline/_ref that __graft_entry__.build() makes and that travels to the GPU box.  `backend=fake` runs the host logic
over the oracle stand-in (CPU), `backend=cuda` (`-m gpu`) the real sm_100a kernels.  Flask / matplotlib are not
installed here, so they are stubbed exactly as far as the reference touches them.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = next((p for p in (os.environ.get("ORBITAL_REFERENCE"), "/root/reference", os.path.join(_HERE, "baseline", "_ref"))
            if p and os.path.isfile(os.path.join(p, "core", "examples.py"))), "/root/reference")
REF = os.path.abspath(REF)
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "core", "examples.py")),
                                reason="reference checkout not present")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_examples_run_unchanged(backend, monkeypatch, tmp_path, golden, capsys):
    import core.plot
    calls = []
    monkeypatch.setattr(core.plot, "plot_orbits", lambda engine, **kw: calls.append(("plot", kw)))
    monkeypatch.setattr(core.plot, "render_orbital_mp4", lambda engine, **kw: calls.append(("mp4", kw)))
    monkeypatch.chdir(tmp_path)                       # the examples write history.jsonl into the CWD
    ex = _load(os.path.join(REF, "core", "examples.py"), "reference_examples")
    import core.engine
    assert ex.SimulationEngine is core.engine.SimulationEngine      # bound to OUR engine by name
    eng = ex.sun_earth_moon(steps=201)
    assert eng.step_idx == 201 and [o.velocity.dtype for o in eng.objects] == [np.float64] * 3
    eng3 = ex.three_body_equilateral(steps=1000)
    g = golden("three_body")
    pos = np.array([o.position() for o in eng3.objects])
    assert np.array_equal(pos, g["pos_1000"])          # bit-identical to the reference engine's run
    eng15 = ex.sol_from_kepler_dataset(days=100)
    g15 = golden("solar15_f32")                        # the reference's own run of the same scenario
    assert np.array_equal(np.array([o.position() for o in eng15.objects]), g15["pos_100"])
    assert len(eng15.objects) == 15 and eng15.objects[3].name == "Earth"
    ex.two_body_problem(steps=50)
    assert [c[0] for c in calls] == ["plot", "mp4", "mp4", "plot"]
    assert os.path.exists(tmp_path / "history.jsonl")  # cache=True default, as in the reference
    out = capsys.readouterr().out
    assert "step 0: ΔE=" in out


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "app", "app.py")), reason="reference app not present")
def test_reference_flask_app_runs_unchanged(backend, monkeypatch):
    # --- minimal Flask stand-in: only what app/app.py uses (Flask, jsonify, render_template, route/get) ---
    flask = types.ModuleType("flask")

    class Flask:
        def __init__(self, name):
            self.routes = {}

        def route(self, path, **kw):
            def deco(fn):
                self.routes[path] = fn
                return fn
            return deco

        get = route

    flask.Flask = Flask
    flask.jsonify = lambda *a, **k: (a[0] if a else k)
    flask.render_template = lambda tpl, **ctx: ctx
    monkeypatch.setitem(sys.modules, "flask", flask)
    monkeypatch.setenv("SIM_INITIAL_STEPS", "40")
    monkeypatch.setenv("SIM_MAX_HISTORY", "25")
    monkeypatch.chdir(REF)                             # app.py reads ./config.json
    app_mod = _load(os.path.join(REF, "app", "app.py"), "reference_app")
    try:
        eng = app_mod.engine
        import core.engine
        assert isinstance(eng, core.engine.SimulationEngine) and len(eng.objects) == 26
        assert eng.step_idx >= 40 and hasattr(eng, "body_map") and eng.sim_epoch_jd == 2451545.0
        state = app_mod.routes_state() if hasattr(app_mod, "routes_state") else app_mod.get_bodies()
        assert len(state["bodies"]) == 26
        earth = [b for b in state["bodies"] if b["name"] == "Earth"][0]
        assert set(earth) == {"id", "name", "mass_kg", "radius_km", "T_seconds", "fg_ms2", "position"}
        assert abs(np.hypot(earth["position"]["x"], earth["position"]["y"]) / 1.496e11 - 1.0) < 0.05
        assert state["time_elapsed"] == eng.time_elapsed or state["time_elapsed"] <= eng.time_elapsed
        page = app_mod.index()                         # '/' route: named_history(limit=5000)
        assert set(page["initial_state"]) == {o.name for o in eng.objects}
        assert 1 <= len(page["initial_state"]["Earth"]) <= 25
        assert app_mod.health()[1] == 200
    finally:
        app_mod.STOP_SIMULATION = True                 # let the daemon stepping thread exit
        app_mod.thread.join(timeout=5)
