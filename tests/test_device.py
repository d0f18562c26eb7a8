"""Parity tests proper: the sm_100a kernels, called through the C ABI, against the oracle.

All tests here need a B200 (`-m gpu`).  Oracle: oracle/nbody_oracle.c (pinned
bit-exact to the reference in tests/test_oracle.py).  Bars:
  faithful mode: bit-exact accelerations / state (reference core/physics.py:125-159, engine.py:65-97)
  fast mode:     relative acceleration error <= 1e-12 per step (BASELINE.json north_star)
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


def device_accel(nat, c, mode, eps=None):
    dev = nat.DeviceSystem(c.n, 0, mode)
    dev.set_params(c["dt"], c["eps"] if eps is None else eps, G)
    dev.upload(*c.arrays())
    dev.accel()
    a = dev.download_acc().T.copy()
    return dev, a


def test_device_info(nat):
    info = nat.device_info(0)
    assert info["cc"][0] >= 10 and info["sm_count"] >= 100


@pytest.mark.parametrize("n", [1, 2, 3, 9, 15, 26, 33, 64, 257, 513, 1024, 4096])
def test_force_faithful_bit_exact(nat, orc, n):
    """pairwise_accelerations vs GPU on random clouds."""
    from core import synthetic
    c = synthetic.random_cloud(n, seed=1000 + n)
    for eps in (c["eps"], 0.0):
        dev, a = device_accel(nat, c, nat.MODE_FAITHFUL, eps)
        ref, U = orc.pairwise(c["x"], c["y"], c["z"], c["m"], eps, G, nthreads=1)
        assert_bits(a, ref, f"n={n} eps={eps}")
        if n > 1:
            assert_bits(dev.potential(), U, f"U n={n}")
        dev.close()


def test_force_golden_vectors_direct(nat, golden):
    """The reference's own outputs (no oracle in between)."""
    from core import synthetic
    g = golden("force_random")
    for k in range(int(g["ncases"])):
        c = synthetic._cloud(np.stack([g[f"c{k}_x"], g[f"c{k}_y"], g[f"c{k}_z"]], 1), np.zeros((len(g[f"c{k}_x"]), 3)),
                             g[f"c{k}_m"], 0.0, dt=1.0, eps=float(g[f"c{k}_eps"]))
        dev, a = device_accel(nat, c, nat.MODE_FAITHFUL)
        assert_bits(a, g[f"c{k}_acc"], f"golden case {k}")
        assert_bits(dev.potential(), g[f"c{k}_U"], f"golden U {k}")
        dev.close()
    # coincident bodies: finite with softening, inf/nan (no exception) without
    c = synthetic._cloud(np.stack([g["coinc_x"], g["coinc_y"], g["coinc_z"]], 1), np.zeros((5, 3)), g["coinc_m"], 0.0,
                         dt=1.0, eps=float(g["coinc_soft_eps"]))
    dev, a = device_accel(nat, c, nat.MODE_FAITHFUL)
    assert_bits(a, g["coinc_soft_acc"]); dev.close()
    dev, a = device_accel(nat, c, nat.MODE_FAITHFUL, eps=0.0)
    assert_bits(a, g["coinc_hard_acc"]); dev.close()
    dev, a = device_accel(nat, c, nat.MODE_FAST, eps=0.0)
    assert np.isfinite(a).all() == np.isfinite(g["coinc_hard_acc"]).all() or not np.isfinite(a).all()
    dev.close()


@pytest.mark.parametrize("n", [2, 31, 64, 1000, 4096, 20000])
@pytest.mark.parametrize("ti", ["1", "2", "4", "6", "8", None])
def test_force_fast_within_tolerance(nat, orc, n, ti, monkeypatch):
    from core import synthetic
    monkeypatch.setenv("ORBITAL_B200_SYM", "0")            # the one-sided kernel (also used by sharded engines)
    if ti is None:
        monkeypatch.delenv("ORBITAL_B200_TI", raising=False)
    else:
        monkeypatch.setenv("ORBITAL_B200_TI", ti)
    c = synthetic.random_cloud(n, seed=2000 + n)
    dev, a = device_accel(nat, c, nat.MODE_FAST)
    rows = np.arange(n, dtype=np.int64) if n <= 4096 else np.arange(0, n, 37, dtype=np.int64)
    ref = orc.pairwise_sample(c["x"], c["y"], c["z"], c["m"], c["eps"], G, rows)
    err = relerr(a[rows], ref)
    assert err.max() <= TOL_FAST, f"n={n} ti={ti}: max rel err {err.max():.3e}"
    dev.close()


@pytest.mark.parametrize("n", [2, 31, 64, 257, 1000, 4096, 20000, 33333])
@pytest.mark.parametrize("ti", ["1", "2", "4", "6", "8", None])
def test_force_symmetric_kernel_within_tolerance(nat, orc, n, ti, monkeypatch):
    """Pair-symmetric kernel (default fast path on one GPU): each pair once, applied to both bodies."""
    from core import synthetic
    monkeypatch.setenv("ORBITAL_B200_SYM", "1")
    if ti is None:
        monkeypatch.delenv("ORBITAL_B200_SYM_TI", raising=False)
    else:
        monkeypatch.setenv("ORBITAL_B200_SYM_TI", ti)
    c = synthetic.random_cloud(n, seed=3000 + n)
    dev, a = device_accel(nat, c, nat.MODE_FAST)
    assert "force_sym_kernel" in dev.force_kernel_info()["name"]
    # every row up to 4096; beyond that a stride plus the rows around every I-block / tile boundary
    rows = np.arange(n, dtype=np.int64) if n <= 4096 else np.unique(np.concatenate(
        [np.arange(0, n, 41), np.arange(0, n, 128), np.arange(127, n, 128), [n - 1]])).astype(np.int64)
    ref = orc.pairwise_sample(c["x"], c["y"], c["z"], c["m"], c["eps"], G, rows)
    err = relerr(a[rows], ref)
    assert err.max() <= TOL_FAST, f"n={n} ti={ti}: max rel err {err.max():.3e}"
    dev.accel()
    assert_bits(dev.download_acc().T, a, "run-to-run determinism")
    dev.close()


@pytest.mark.parametrize("n", [6144, 12288, 6000])
@pytest.mark.parametrize("ti", ["1", "2", "4", "6", "8", None])
def test_force_symmetric_uniform_mass_variant(nat, orc, n, ti, monkeypatch):
    """Equal masses + whole I-blocks/tiles: the kernel drops the per-pair mass multiplies (G*m applied once)."""
    from core import synthetic
    monkeypatch.setenv("ORBITAL_B200_SYM", "1")
    if ti is None:
        monkeypatch.delenv("ORBITAL_B200_SYM_TI", raising=False)
    else:
        monkeypatch.setenv("ORBITAL_B200_SYM_TI", ti)
    c = synthetic.plummer(n, seed=n)
    assert np.all(c["m"] == c["m"][0])
    dev, a = device_accel(nat, c, nat.MODE_FAST)
    name = dev.force_kernel_info()["name"]
    aligned = n % 256 == 0 and n % (128 * int(name.split("<")[1].split(",")[0])) == 0
    assert name.endswith(",true>") == aligned, name
    rows = np.arange(0, n, 3, dtype=np.int64)
    ref = orc.pairwise_sample(c["x"], c["y"], c["z"], c["m"], c["eps"], G, rows)
    assert relerr(a[rows], ref).max() <= TOL_FAST
    monkeypatch.setenv("ORBITAL_B200_SYM_UNI", "0")          # general-mass variant on the same input
    dev.accel()
    assert dev.force_kernel_info()["name"].endswith(",false>")
    b = dev.download_acc().T
    assert relerr(a, b).max() <= 1e-13
    # one body with a different mass must switch the variant off
    monkeypatch.delenv("ORBITAL_B200_SYM_UNI")
    arrs = list(c.arrays())
    m2 = arrs[6].copy(); m2[n // 2] *= 3.0; arrs[6] = m2
    dev.upload(*arrs)
    dev.accel()
    assert dev.force_kernel_info()["name"].endswith(",false>")
    ref2 = orc.pairwise_sample(c["x"], c["y"], c["z"], m2, c["eps"], G, rows)
    assert relerr(dev.download_acc().T[rows], ref2).max() <= TOL_FAST
    dev.close()


@pytest.mark.parametrize("chunks,pj_bytes", [("1", None), ("7", None), (None, str(3 * 8 * 9000 * 5)), ("3", str(3 * 8 * 9000))])
def test_force_symmetric_chunks_and_panels(nat, orc, chunks, pj_bytes, monkeypatch):
    """Chunked tile ranges and multi-panel P_j processing give the same answer."""
    from core import synthetic
    monkeypatch.setenv("ORBITAL_B200_SYM", "1")
    monkeypatch.setenv("ORBITAL_B200_SYM_TI", "2")
    if chunks:
        monkeypatch.setenv("ORBITAL_B200_SYM_CHUNKS", chunks)
    if pj_bytes:
        monkeypatch.setenv("ORBITAL_B200_SYM_PJ_BYTES", pj_bytes)
    c = synthetic.plummer(9000, seed=9)
    dev, a = device_accel(nat, c, nat.MODE_FAST)
    rows = np.arange(0, 9000, 7, dtype=np.int64)
    ref = orc.pairwise_sample(c["x"], c["y"], c["z"], c["m"], c["eps"], G, rows)
    assert relerr(a[rows], ref).max() <= TOL_FAST
    F = (c["m"][:, None] * a).sum(0)
    assert np.linalg.norm(F) <= 1e-12 * (c["m"] * np.linalg.norm(a, axis=1)).sum()
    dev.close()


@pytest.mark.parametrize("slabs", ["1", "3", "7"])
def test_force_fast_slab_decomposition_is_deterministic(nat, slabs, monkeypatch):
    from core import synthetic
    monkeypatch.setenv("ORBITAL_B200_SYM", "0")
    monkeypatch.setenv("ORBITAL_B200_SLABS", slabs)
    c = synthetic.plummer(6000, seed=5)
    dev, a1 = device_accel(nat, c, nat.MODE_FAST)
    dev.accel()
    a2 = dev.download_acc().T
    assert_bits(a1, a2, "run-to-run")
    dev.close()


def test_force_fast_plummer_262144_sampled_rows(nat, orc):
    """BASELINE config C2 at full size: sampled-row oracle + long-double yardstick + Newton's third law."""
    from core import synthetic
    c = synthetic.plummer(262144)
    dev, a = device_accel(nat, c, nat.MODE_FAST)
    rng = np.random.default_rng(0)
    r = np.sqrt(c["x"] ** 2 + c["y"] ** 2 + c["z"] ** 2)
    rows = np.unique(np.concatenate([rng.choice(c.n, 240, replace=False), np.argsort(r)[:16]])).astype(np.int64)
    ref = orc.pairwise_sample(c["x"], c["y"], c["z"], c["m"], c["eps"], G, rows)
    ld, sum_abs = orc.pairwise_sample(c["x"], c["y"], c["z"], c["m"], c["eps"], G, rows, long_double=True)
    err_ref = relerr(a[rows], ref)
    err_ld = relerr(a[rows], ld)
    cond = np.linalg.norm(a[rows] - ld, axis=1) / sum_abs
    print(f"\nN=262144 fast kernel: vs reference-order fp64 max {err_ref.max():.2e} median {np.median(err_ref):.2e}; "
          f"vs long double max {err_ld.max():.2e}; |da|/sum|a_ij| max {cond.max():.2e}")
    assert err_ref.max() <= TOL_FAST and err_ld.max() <= TOL_FAST
    # size-independent property: total force vanishes (sum m_i a_i = 0) to rounding
    F = (c["m"][:, None] * a).sum(0)
    scale = (c["m"] * np.linalg.norm(a, axis=1)).sum()
    assert np.linalg.norm(F) <= 1e-11 * scale
    dev.close()


@pytest.mark.parametrize("n", [65, 100, 300, 512])
def test_fused_single_cta_kernel_equals_kernel_sequence(nat, orc, n, monkeypatch):
    """64 < n <= 512: the multi-CTA kernel sequence is the default (10-19 us/step against 25-1256 for the fused
    tiny_steps_kernel); both must produce the oracle's bits, history ring included."""
    from core import synthetic
    from oracle.c_oracle import State
    c = synthetic.random_cloud(n, seed=500 + n)
    K = 12
    st = State(orc, c["x"], c["y"], c["z"], c["vx"], c["vy"], c["vz"], c["m"], c["radius"], 0, c["dt"], c["eps"])
    st.step(K)
    for limit, kernel in (("512", "tiny_steps_kernel"), (None, "faithful_pairs_kernel")):
        if limit:
            monkeypatch.setenv("ORBITAL_B200_TINY_MAX", limit)
        else:
            monkeypatch.delenv("ORBITAL_B200_TINY_MAX", raising=False)
        dev = nat.DeviceSystem(c.n, 0, nat.MODE_FAITHFUL)
        dev.set_params(c["dt"], c["eps"], G)
        dev.set_history(4)
        dev.upload(*c.arrays())
        dev.accel()
        assert kernel in dev.force_kernel_info()["name"]
        assert dev.step(K)[0] == K
        s = dev.download_state()
        assert_bits(np.stack([s["x"], s["y"], s["z"]], 1), st.pos, f"{kernel} positions")
        assert_bits(np.stack([s["vx"], s["vy"], s["vz"]], 1), st.vel, f"{kernel} velocities")
        h = dev.history_download(4)
        assert h.shape[0] == 4
        assert_bits(h[-1], st.pos, f"{kernel} last history point")
        dev.close()


@pytest.mark.parametrize("n", [513, 1500, 4099])
def test_force_faithful_two_pass_equals_one_pass(nat, n, monkeypatch):
    """512 < n <= 32768: pair matrix of 1/r^3 (each pair's sqrt/div once) + ordered row sums; must be bit-identical
    to the one-pass kernel, including the overlap list."""
    from core import synthetic
    c = synthetic.random_cloud(n, seed=77 + n, radius=2.5e9)       # radii large enough for a few hundred overlaps
    out = []
    for pairs in ("1", "0"):
        monkeypatch.setenv("ORBITAL_B200_FAITHFUL_PAIRS", pairs)
        dev = nat.DeviceSystem(c.n, 0, nat.MODE_FAITHFUL)
        dev.set_params(c["dt"], 0.0, G)
        dev.upload(*c.arrays())
        dev.accel()
        name = dev.force_kernel_info()["name"]
        assert ("faithful_pairs_kernel" in name) == (pairs == "1"), name
        a = dev.download_acc()
        done, nov = dev.step(1)
        pl, cnt = dev.overlap_pairs()
        out.append((a, dev.download_state(), sorted(map(tuple, pl.tolist())), cnt))
        dev.close()
    assert_bits(out[0][0], out[1][0], "acc")
    for k in out[0][1]:
        assert_bits(out[0][1][k], out[1][1][k], f"state {k}")
    assert out[0][3] == out[1][3] and out[0][2] == out[1][2] and out[0][3] > 0


@pytest.mark.parametrize("n", [600, 1500])
def test_step_faithful_kernel_sequence_bit_exact(nat, orc, n):
    """n > 512 takes the multi-kernel CUDA-graph path; n = 600 also has radii -> overlap detection on."""
    from core import synthetic
    from oracle.c_oracle import State
    c = synthetic.random_cloud(n, seed=n, radius=(1e3 if n == 600 else 0.0))
    f32 = (np.arange(n) % 2).astype(np.uint8)
    dev = nat.DeviceSystem(n, 0, nat.MODE_FAITHFUL)
    dev.set_params(c["dt"], c["eps"], G)
    dev.set_history(8)
    vel = [np.where(f32 == 1, v.astype(np.float32).astype(np.float64), v) for v in (c["vx"], c["vy"], c["vz"])]
    dev.upload(c["x"], c["y"], c["z"], *vel, c["m"], c["radius"], f32)
    dev.accel()
    dev.history_append()
    st = State(orc, *c.arrays(), f32, c["dt"], c["eps"])
    for k in (1, 3, 20):
        done, nov = dev.step(k)
        assert (done, nov) == (k, 0)
        st.step(k, collisions=False)
        s = dev.download_state()
        assert_bits(np.stack([s["x"], s["y"], s["z"]], 1), st.pos, f"pos after +{k}")
        assert_bits(np.stack([s["vx"], s["vy"], s["vz"]], 1), st.vel, f"vel after +{k}")
        assert_bits(dev.download_acc().T, st.acc, f"acc after +{k}")
    assert dev.history_count() == 25
    h = dev.history_download(8)
    assert h.shape == (8, n, 3)
    assert_bits(h[-1], st.pos, "last ring entry")
    assert dev.launch_count() >= 24 * 3          # kick_drift + pair matrix + (ordered rows, kick, history, bookkeeping)
    dev.close()


@pytest.mark.parametrize("rows", ["1", "2", "4"])
def test_faithful_pass2_variants_are_bit_identical(nat, orc, monkeypatch, rows):
    """ORBITAL_B200_ROWS selects the pass-2 kernel of the two-pass bit-exact force (one warp per 32 targets,
    producer / consumer warps, four lanes per target): all three equal the oracle bit for bit, ragged sizes too."""
    from core import synthetic
    monkeypatch.setenv("ORBITAL_B200_ROWS", rows)
    for n in (513, 1000, 2048):
        c = synthetic.random_cloud(n, seed=7 * n)
        ref, _ = orc.pairwise(c["x"], c["y"], c["z"], c["m"], c["eps"], G, 1)
        dev, a = device_accel(nat, c, nat.MODE_FAITHFUL)
        assert_bits(a, ref, f"rows={rows} n={n}")
        assert dev.step(3) == (3, 0)
        dev.close()


def test_disk4096_matches_reference_engine(nat, golden):
    """BASELINE config C1: the real reference's ctor + 2 steps at N=4096 (fixture) vs the faithful GPU path."""
    from core import synthetic
    g = golden("disk4096_f32")
    c = synthetic.uniform_disk(4096)
    dev = nat.DeviceSystem(4096, 0, nat.MODE_FAITHFUL)
    dev.set_params(c["dt"], c["eps"], G)
    vel = [v.astype(np.float32).astype(np.float64) for v in (c["vx"], c["vy"], c["vz"])]
    dev.upload(c["x"], c["y"], c["z"], *vel, c["m"], c["radius"], np.ones(4096, np.uint8))
    dev.accel()
    assert_bits(dev.download_acc().T, g["acc_0"], "ctor acc")
    assert_bits(dev.potential(), g["U_0"], "ctor U")
    for s in (1, 2):
        assert dev.step(1) == (1, 0)
        st = dev.download_state()
        assert_bits(np.stack([st["x"], st["y"], st["z"]], 1), g[f"pos_{s}"], f"pos {s}")
        assert_bits(np.stack([st["vx"], st["vy"], st["vz"]], 1), g[f"vel_{s}"], f"vel {s}")
        assert_bits(dev.download_acc().T, g[f"acc_{s}"], f"acc {s}")
        assert_bits(dev.potential(), g[f"U_{s}"], f"U {s}")
    # fast mode on the same steps: <= 1e-12 relative acceleration error per step
    fast = nat.DeviceSystem(4096, 0, nat.MODE_FAST)
    fast.set_params(c["dt"], c["eps"], G)
    fast.upload(c["x"], c["y"], c["z"], *vel, c["m"], c["radius"], np.ones(4096, np.uint8))
    fast.accel()
    assert relerr(fast.download_acc().T, g["acc_0"]).max() <= TOL_FAST
    assert fast.step(2) == (2, 0)
    assert relerr(fast.download_acc().T, g["acc_2"]).max() <= 1e-9     # trajectories have diverged by ~1 ulp of state
    sf = fast.download_state()
    dpos = np.linalg.norm(np.stack([sf["x"], sf["y"], sf["z"]], 1) - g["pos_2"], axis=1)
    assert (dpos[1:] / np.linalg.norm(g["pos_2"][1:], axis=1)).max() <= 1e-12
    assert abs(fast.potential() - float(g["U_2"])) <= 1e-12 * abs(float(g["U_2"]))
    dev.close(); fast.close()


@pytest.mark.parametrize("mode_name", ["faithful", "fast"])
@pytest.mark.parametrize("n", [24, 700])
def test_overlap_detection_matches_exact_test(nat, mode_name, n):
    """Fused detection (physics.py:517-518): the flagged pair set equals the exact host test; device halts."""
    rng = np.random.default_rng(n)
    box = 4e4 if n == 24 else 3e5
    x, y, z = (rng.uniform(-box, box, n) for _ in range(3))
    v = rng.standard_normal((3, n)) * 50
    m = np.exp(rng.uniform(np.log(1e14), np.log(1e16), n))
    radius = rng.uniform(2e3, 8e3, n)
    mode = nat.MODE_FAITHFUL if mode_name == "faithful" else nat.MODE_FAST
    dev = nat.DeviceSystem(n, 0, mode)
    dev.set_params(2.0, 10.0, G)
    dev.set_history(4)
    dev.upload(x, y, z, v[0], v[1], v[2], m, radius)
    dev.accel()
    dev.history_append()
    done, nov = dev.step(5)
    assert done == 1 and nov > 0, "device must halt after the first overlapping step"
    assert dev.history_count() == 1, "the halting step's snapshot is appended by the host after resolution"
    s = dev.download_state()
    P = np.stack([s["x"], s["y"], s["z"]], 1)
    want = {(i, j) for i in range(n) for j in range(i + 1, n)
            if np.linalg.norm(P[i] - P[j]) <= radius[i] + radius[j]}
    pairs, count = dev.overlap_pairs()
    got = {tuple(p) for p in pairs.tolist()}
    assert count == len(pairs) == nov
    assert got == want, f"{len(got ^ want)} pairs differ"
    # after 'resolution' (here: shrink the radii) the run resumes
    dev.upload(s["x"], s["y"], s["z"], s["vx"], s["vy"], s["vz"], m, np.full(n, 1e-3))
    dev.history_append()
    assert dev.step(4) == (4, 0)
    assert dev.history_count() == 6
    dev.close()


@pytest.mark.parametrize("n", [100, 700])
def test_device_side_contact_resolution_matches_oracle_sweep(nat, orc, n):
    """orb_set_contacts(on_device=1) on the fused single-CTA path (n=100) and the multi-kernel path (n=700):
    bit-exact with the reference's sequential sweep (oracle orc_collisions), incl. history and U."""
    from oracle.c_oracle import State
    rng = np.random.default_rng(n)
    box = 8e4 if n == 100 else 3e5
    x, y, z = (rng.uniform(-box, box, n) for _ in range(3))
    v = rng.standard_normal((3, n)) * 400
    m = np.exp(rng.uniform(np.log(1e14), np.log(1e16), n))
    radius = rng.uniform(2e3, 8e3, n)
    f32 = (np.arange(n) % 2).astype(np.uint8)
    vel = [np.where(f32 == 1, a.astype(np.float32).astype(np.float64), a) for a in v]
    st = State(orc, x, y, z, *vel, m, radius, f32, 2.0, 10.0, G, restitution=0.8)
    dev = nat.DeviceSystem(n, 0, nat.MODE_FAITHFUL)
    dev.set_params(2.0, 10.0, G)
    dev.set_contacts(0.8, True)
    dev.set_history(16)
    dev.upload(x, y, z, *vel, m, radius, f32)
    dev.accel()
    dev.history_append()
    total = 0
    for k in (1, 2, 5):
        done, resolved = dev.step(k)
        assert done == k, "the device never halts when it resolves contacts itself"
        total += resolved
        for _ in range(k):
            st.step(1, collisions=True)
        s = dev.download_state()
        assert_bits(np.stack([s["x"], s["y"], s["z"]], 1), st.pos, f"pos after +{k}")
        assert_bits(np.stack([s["vx"], s["vy"], s["vz"]], 1), st.vel, f"vel after +{k}")
        assert_bits(dev.download_acc().T, st.acc, f"acc after +{k}")
        assert_bits(dev.potential(), st.U, f"U after +{k} (stashed before the push-out)")
    assert total == st.hits and total > 0
    assert dev.history_count() == 9
    assert_bits(dev.history_download(1)[0], st.pos, "history holds post-contact positions")
    dev.close()


def test_engine_fast_mode_collisions_close_to_reference(golden):
    """Fast-mode engine with contacts: same contact sequence, state within 1e-9 of the reference run."""
    from tests.test_engine import build_engine, state_of
    g = golden("coll_dense_f32")
    eng = build_engine(g, mode="fast")
    eng.run(30)
    p, v, _ = state_of(eng)
    assert np.abs(p - g["pos_30"]).max() <= 1e-9 * np.abs(g["pos_30"]).max()
    assert np.abs(v - g["vel_30"]).max() <= 1e-6 * np.abs(g["vel_30"]).max()


def test_ensemble_faithful_bit_exact_and_fast_close(nat, orc):
    """BASELINE config C3 (small batch): every system == a standalone reference engine on its 16 bodies."""
    from core import synthetic
    from core.ensemble import EnsembleEngine
    e = synthetic.ensemble(48, 16)
    K = 200
    for vel_f32 in (False, True):
        want = {k: np.array(e[k], dtype=np.float64, copy=True) for k in ("x", "y", "z", "vx", "vy", "vz")}
        if vel_f32:
            for k in ("vx", "vy", "vz"):
                want[k] = want[k].astype(np.float32).astype(np.float64)
        orc.lib.orc_ensemble_step(48, 16, want["x"].reshape(-1), want["y"].reshape(-1), want["z"].reshape(-1),
                                  want["vx"].reshape(-1), want["vy"].reshape(-1), want["vz"].reshape(-1),
                                  np.ascontiguousarray(e["m"]).reshape(-1), e["dt"], e["eps"], G, K, int(vel_f32), 0)
        for fused in (True, False):
            ens = EnsembleEngine(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")), dt=e["dt"],
                                 softening=e["eps"], mode="faithful", vel_f32=vel_f32)
            if fused:
                ens.step(K, fused=True)
            else:
                ens.step(K // 2, fused=False); ens.step(K - K // 2, fused=True)
            got = ens.state()
            for k in want:
                assert_bits(got[k], want[k], f"ensemble {k} vel_f32={vel_f32} fused={fused}")
            ens.close()
    fast = EnsembleEngine(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")), dt=e["dt"], softening=e["eps"],
                          mode="fast")
    E0 = fast.energy()
    fast.step(K, fused=True)
    got = fast.state()
    ref64 = {k: np.array(e[k], dtype=np.float64, copy=True) for k in ("x", "y", "z", "vx", "vy", "vz")}
    orc.lib.orc_ensemble_step(48, 16, *(ref64[k].reshape(-1) for k in ("x", "y", "z", "vx", "vy", "vz")),
                              np.ascontiguousarray(e["m"]).reshape(-1), e["dt"], e["eps"], G, K, 0, 0)
    num = np.sqrt(sum((got[k] - ref64[k]) ** 2 for k in ("x", "y", "z")))
    den = np.sqrt(sum(ref64[k] ** 2 for k in ("x", "y", "z")))
    assert (num[:, 1:] / den[:, 1:]).max() <= 1e-9, "fast ensemble trajectory divergence after 200 steps"
    E1 = fast.energy()
    assert np.abs((E1 - E0) / E0).max() < 1e-3
    fast.close()


@pytest.mark.parametrize("narrow", ["0", "1"])
@pytest.mark.parametrize("nbody", [2, 5, 13, 16, 17, 32])
def test_ensemble_fast_acceleration_within_tolerance_every_step(nat, orc, nbody, narrow, monkeypatch):
    """ens_step_fast_kernel (two bodies per lane) and ens_step_fast1_kernel (one body per lane, the layout small
    batches take): relative acceleration error <= 1e-12 per step (BASELINE north_star), checked on every
    body of 48 systems against the oracle at the SAME positions, after un-fused and fused steps, both velocity
    dtypes -- the ensemble keeps every system's engine.acc (orb_ens_download_acc)."""
    from core import synthetic
    from core.ensemble import EnsembleEngine
    monkeypatch.setenv("ORBITAL_B200_ENS_NARROW", narrow)
    e = synthetic.ensemble(48, nbody)
    args = [e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")]
    worst = 0.0
    for vel_f32 in (False, True):
        ens = EnsembleEngine(*args, dt=e["dt"], softening=e["eps"], mode="fast", vel_f32=vel_f32)
        for fused, k in ((False, 0), (False, 1), (True, 1), (False, 3), (True, 20), (False, 17)):
            if k:
                ens.step(k, fused=fused)
            st, a = ens.state(), ens.acc()
            for s in range(48):
                ref, _ = orc.pairwise(st["x"][s], st["y"][s], st["z"][s], e["m"][s], e["eps"], G, 1)
                err = relerr(np.ascontiguousarray(a[:, s, :].T), ref)
                worst = max(worst, float(err.max()))
                assert err.max() <= TOL_FAST, f"nbody={nbody} f32={vel_f32} system {s}: {err.max():.3e}"
        ens.close()
    print(f"\nensemble fast kernel nbody={nbody} narrow={narrow}: max rel acceleration error {worst:.2e}")


def _ensemble_contact_scene(nsys, nb, seed):
    """Crowded small systems: bodies in a box a few radii wide so that several pairs touch in most steps."""
    rng = np.random.default_rng(seed)
    sh = (nsys, nb)
    out = dict(x=rng.uniform(-4e4, 4e4, sh), y=rng.uniform(-4e4, 4e4, sh), z=rng.uniform(-4e4, 4e4, sh),
               vx=rng.standard_normal(sh) * 300, vy=rng.standard_normal(sh) * 300, vz=rng.standard_normal(sh) * 300,
               m=np.exp(rng.uniform(np.log(1e14), np.log(1e16), sh)), radius=rng.uniform(2e3, 9e3, sh))
    out["radius"][::3] = 0.0                                   # every third system has no radii at all
    out["f32"] = rng.integers(0, 2, sh).astype(np.uint8)       # mixed velocity dtypes inside a system
    return out


@pytest.mark.parametrize("nb", [6, 16, 21])
def test_ensemble_contacts_and_mixed_dtypes_equal_standalone_engines(nat, orc, nb):
    """each system == a standalone reference SimulationEngine on its bodies, INCLUDING the
    contact sweep of engine.py:85 and per-body velocity dtypes.  Bit-exact mode: bit for bit, fused and un-fused;
    fast mode: same contacts, state within 1e-9."""
    from core.ensemble import EnsembleEngine
    from oracle.c_oracle import State
    nsys, K, dt, eps, rest = 36, 23, 2.0, 10.0, 0.8
    e = _ensemble_contact_scene(nsys, nb, 100 + nb)
    vel = {k: np.where(e["f32"] == 1, e[k].astype(np.float32).astype(np.float64), e[k]) for k in ("vx", "vy", "vz")}
    want = {k: np.empty((nsys, nb)) for k in ("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az")}
    hits = 0
    for s in range(nsys):
        st = State(orc, e["x"][s], e["y"][s], e["z"][s], vel["vx"][s], vel["vy"][s], vel["vz"][s], e["m"][s],
                   e["radius"][s], e["f32"][s], dt, eps, G, restitution=rest)
        st.step(K, collisions=True)
        hits += st.hits
        for k in want:
            want[k][s] = getattr(st, k)
    assert hits > 10, "the scene must actually produce contacts"
    args = (e["x"], e["y"], e["z"], vel["vx"], vel["vy"], vel["vz"], e["m"])
    for fused in (True, False):
        ens = EnsembleEngine(*args, dt=dt, softening=eps, mode="faithful", vel_f32=e["f32"], radius=e["radius"],
                             restitution=rest)
        if fused:
            ens.step(K, fused=True)
        else:
            ens.step(K - 4, fused=False); ens.step(4, fused=True)
        got, acc = ens.state(), ens.acc()
        for k in ("x", "y", "z", "vx", "vy", "vz"):
            assert_bits(got[k], want[k], f"ensemble nb={nb} fused={fused} {k}")
        for c, k in enumerate(("ax", "ay", "az")):
            assert_bits(acc[c], want[k], f"ensemble nb={nb} fused={fused} {k}")
        assert ens.contacts_resolved() == hits
        ens.close()
    fast = EnsembleEngine(*args, dt=dt, softening=eps, mode="fast", vel_f32=e["f32"], radius=e["radius"],
                          restitution=rest)
    fast.step(K, fused=False)
    got = fast.state()
    assert fast.contacts_resolved() == hits
    for k in ("x", "y", "z"):
        assert np.abs(got[k] - want[k]).max() <= 1e-9 * np.abs(want[k]).max(), k
    fast.close()


@pytest.mark.parametrize("nb", [5, 16, 32])
def test_ensemble_time_sliced_fused_kernel_is_bit_identical(nat, monkeypatch, nb):
    """Small batches in fused mode run ens_fast_sliced_kernel (work queue of (system group, 16-step slice) items,
    state handed over in its synchronised form): same bits as the single pass and as one launch per step."""
    from core import synthetic
    monkeypatch.setenv("ORBITAL_B200_ENS_NARROW", "0")  # (small batches default to one body per lane, never sliced)
    e = synthetic.ensemble(301, nb)
    args = [e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")]
    K = 117                                              # 7 slices of 16 + one of 5
    outs = []
    for slice_env, fused in (("16", True), ("0", True), ("16", False), ("7", True)):
        monkeypatch.setenv("ORBITAL_B200_ENS_SLICE", slice_env)
        ens = nat.DeviceEnsemble(301, nb, 0, nat.MODE_FAST, vel_f32=True)
        ens.set_params(e["dt"], e["eps"], e["G"])
        ens.upload(*[np.asarray(a, dtype=np.float32).astype(np.float64) if i in (3, 4, 5) else a
                     for i, a in enumerate(args)])
        ens.step(K, fused=fused)
        outs.append((ens.download(), ens.download_acc()))
        ens.close()
    for st, acc in outs[1:]:
        for k in st:
            assert_bits(st[k], outs[0][0][k], f"sliced vs single-pass {k}")
        assert_bits(acc, outs[0][1], "accelerations")


@pytest.mark.parametrize("nb", [2, 3, 7, 16, 21, 32])
def test_ensemble_one_body_per_lane_layout(nat, monkeypatch, nb):
    """The narrow layout (ens_step_fast1_kernel: one body per lane, what a batch with fewer than 6 warps of work per
    SM sub-partition takes by default): fused, one launch per step (incl. the 16-step graphs) and mixed calls give
    the same bits; per-body velocity dtypes; padded slots and the ragged last warp contribute nothing; the
    trajectories stay within 1e-10 of the two-body layout (which other tests hold to the oracle) over 40 steps."""
    from core import synthetic
    nsys = 203
    e = synthetic.ensemble(nsys, nb)
    args = [e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")]
    flags = (np.arange(nsys * nb).reshape(nsys, nb) % 3 == 0).astype(np.uint8)
    vel = [np.where(flags, np.asarray(a, dtype=np.float32).astype(np.float64), a) for a in args[3:6]]
    outs = []
    for narrow, plan in (("1", ((40, True),)), ("1", ((40, False),)), ("1", ((7, False), (16, True), (17, False))),
                         ("0", ((40, True),))):
        monkeypatch.setenv("ORBITAL_B200_ENS_NARROW", narrow)
        ens = nat.DeviceEnsemble(nsys, nb, 0, nat.MODE_FAST)
        ens.set_params(e["dt"], e["eps"], e["G"])
        ens.set_bodies(np.zeros((nsys, nb)), flags)
        ens.upload(*args[:3], *vel, args[6])
        for k, fused in plan:
            ens.step(k, fused=fused)
        outs.append((ens.download(), ens.download_acc()))
        ens.close()
    for st, acc in outs[1:3]:
        for k in st:
            assert_bits(st[k], outs[0][0][k], f"narrow layout, fused vs un-fused {k}")
        assert_bits(acc, outs[0][1], "accelerations")
    wide = outs[3][0]
    for k in ("x", "y", "z"):
        scale = np.abs(wide[k]).max()
        assert np.all(np.isfinite(outs[0][0][k]))
        assert np.abs(outs[0][0][k] - wide[k]).max() <= 1e-10 * scale, f"narrow vs two-body layout {k}"


def test_ensemble_odd_sizes(nat, orc):
    from core import synthetic
    from core.ensemble import EnsembleEngine
    for nb in (2, 3, 4, 5, 8, 9, 16, 17, 31, 32):
        e = synthetic.ensemble(7, nb)
        args = [e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")]
        ref = {k: np.array(e[k], dtype=np.float64, copy=True) for k in ("x", "y", "z", "vx", "vy", "vz")}
        orc.lib.orc_ensemble_step(7, nb, *(ref[k].reshape(-1) for k in ("x", "y", "z", "vx", "vy", "vz")),
                                  np.ascontiguousarray(e["m"]).reshape(-1), e["dt"], e["eps"], G, 10, 0, 0)
        f = EnsembleEngine(*args, dt=e["dt"], softening=e["eps"], mode="faithful"); f.step(10)
        for k in ref:
            assert_bits(f.state()[k], ref[k], f"nb={nb} {k}")
        q = EnsembleEngine(*args, dt=e["dt"], softening=e["eps"], mode="fast"); q.step(10)
        assert np.abs(q.state()["x"] - ref["x"]).max() <= 1e-10 * np.abs(ref["x"]).max()
        f.close(); q.close()
    # no softening, central body exactly at the origin, padded slots (nb < next power of two), more systems than
    # one CTA holds: padded / out-of-range slots must contribute exactly nothing (no 0 * inf)
    for nb, nsys in ((5, 67), (13, 33)):
        e = synthetic.ensemble(nsys, nb)
        args = [e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")]
        ref = {k: np.array(e[k], dtype=np.float64, copy=True) for k in ("x", "y", "z", "vx", "vy", "vz")}
        orc.lib.orc_ensemble_step(nsys, nb, *(ref[k].reshape(-1) for k in ("x", "y", "z", "vx", "vy", "vz")),
                                  np.ascontiguousarray(e["m"]).reshape(-1), e["dt"], 0.0, G, 5, 0, 0)
        for fused in (True, False):
            q = EnsembleEngine(*args, dt=e["dt"], softening=0.0, mode="fast"); q.step(5, fused=fused)
            got = q.state()
            for k in ref:
                assert np.all(np.isfinite(got[k]))
                assert np.abs(got[k] - ref[k]).max() <= 1e-10 * np.abs(ref[k]).max(), (nb, k, fused)
            q.close()


def test_graph_replay_on_the_legacy_default_stream(nat):
    """Handles bound to the NULL stream (what torch's default stream is) still replay their CUDA graphs: capture
    happens on the library's own stream, the launch goes to the bound stream."""
    from core import synthetic
    c = synthetic.random_cloud(600, seed=3)          # > 512 bodies: the kernel-sequence path that replays graphs
    outs = []
    for legacy in (False, True):
        dev = nat.DeviceSystem(c.n, 0, nat.MODE_FAITHFUL)
        if legacy:
            dev.set_stream(0)
        dev.set_params(c["dt"], c["eps"], G)
        dev.upload(*c.arrays())
        dev.accel()
        assert dev.step(40)[0] == 40
        outs.append(dev.download_state())
        dev.close()
    for k in outs[0]:
        assert_bits(outs[0][k], outs[1][k], f"legacy-stream engine {k}")
    e = synthetic.ensemble(40, 16)
    res = []
    for legacy in (False, True):
        ens = nat.DeviceEnsemble(40, 16, 0, nat.MODE_FAST)
        if legacy:
            ens.set_stream(0)
        ens.set_params(e["dt"], e["eps"], e["G"])
        ens.upload(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")))
        ens.step(37, fused=False)                       # 2 graph replays of 16 + 5 single launches
        res.append(ens.download())
        ens.close()
    ens = nat.DeviceEnsemble(40, 16, 0, nat.MODE_FAST)
    ens.set_params(e["dt"], e["eps"], e["G"])
    ens.upload(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")))
    ens.step(37, fused=True)
    res.append(ens.download())
    ens.close()
    for k in res[0]:
        assert_bits(res[0][k], res[1][k], f"legacy-stream ensemble {k}")
        assert_bits(res[0][k], res[2][k], f"fused vs un-fused ensemble {k}")


def test_kepler_states_device_vs_reference(nat, golden):
    """Batched elements -> state on the device vs the reference's Body.get_state outputs (kepler_batch.npz).

    Same operation order as body.py:184-249; the trig mode decides the rest:
      libm (default)  glibc's sin/cos restated (csrc/sincos_libm.h): EVERY state and eccentric anomaly bit-identical
                      to what the unmodified reference produced on this image's host;
      cr              correctly rounded sin/cos (csrc/sincos_cr.h): bit-identical to the oracle pipeline run with the
                      host build of the same routine, > 99 % of the golden states;
      fast            CUDA sincos: |dE| <= 4e-15/(1-e), relative state error <= 1e-13."""
    from oracle import ref_numpy
    from tests.test_sincos import build_host_trig, scalar_trig
    g = golden("kepler_batch")
    cols = [g[k] for k in ("M", "e", "a", "b", "n", "inc", "Omega", "omega")]
    assert nat.get_trig_mode() == nat.TRIG_LIBM
    r, v, E = nat.kepler_states(*cols, return_E=True)
    assert_bits(E, g["E"], "eccentric anomaly, libm mode")
    assert_bits(r, g["r"], "positions, libm mode")
    assert_bits(v, g["v"], "velocities, libm mode")
    try:
        nat.set_trig_mode("cr")
        r, v, E = nat.kepler_states(*cols, return_E=True)
        sin, cos = scalar_trig(build_host_trig(), "sc_host_sincos")
        rr, vv, EE = ref_numpy.kepler_states(*cols, sin=sin, cos=cos)
        assert_bits(E, EE, "eccentric anomaly, cr mode"); assert_bits(r, rr, "r, cr mode"); assert_bits(v, vv, "v, cr mode")
        exact_cr = np.mean(np.all(r == g["r"], axis=1) & np.all(v == g["v"], axis=1))
        assert exact_cr > 0.98
        nat.set_trig_mode("fast")
        r, v, E = nat.kepler_states(*cols, return_E=True)
    finally:
        nat.set_trig_mode("libm")
    dE = np.abs(E - g["E"]) * (1.0 - g["e"])
    er = np.linalg.norm(r - g["r"], axis=1) / np.linalg.norm(g["r"], axis=1)
    ev = np.linalg.norm(v - g["v"], axis=1) / np.linalg.norm(g["v"], axis=1)
    exact = np.mean(np.all(r == g["r"], axis=1) & np.all(v == g["v"], axis=1))
    print(f"\nkepler batch: libm mode 100 % bit-identical; cr mode {100 * exact_cr:.2f} %; fast mode max |dE|(1-e) "
          f"{dE.max():.2e}, max rel r {er.max():.2e}, v {ev.max():.2e}, bit-identical states {100 * exact:.1f} %")
    assert dE.max() <= 4e-15 and er.max() <= 1e-13 and ev.max() <= 1e-13
    with pytest.raises(nat.NativeError):
        nat.set_trig_mode(7)
    # the solve_kepler grid of the reference (kepler.npz), through the same kernel
    k = golden("kepler")
    MM, ee = np.meshgrid(k["kep_M"], k["kep_e"])
    one = np.ones(MM.size)
    _, _, E2 = nat.kepler_states(MM.ravel(), ee.ravel(), one, one, one, 0 * one, 0 * one, 0 * one, return_E=True)
    assert_bits(E2, k["kep_E"].ravel(), "solve_kepler grid, libm mode")
    # host API
    from core.datasets import solar_system_v2
    system = solar_system_v2(moons=True)
    system.standardize_units(mass_unit="kilograms", distance_unit="meters", angle_unit="radians", time_unit="seconds")
    rs, vs = system.get_states()
    assert np.allclose(rs, k["r"], rtol=1e-13, atol=0) and np.allclose(vs, k["v"], rtol=1e-13, atol=1e-20)


def test_device_trig_equals_host_builds(nat):
    """The device build of csrc/sincos_libm.h / sincos_cr.h vs the host builds the CPU suite pins to glibc / mpmath
    (tests/test_sincos.py), on every branch: with e = 0, a = b = 1 and no rotation the state is (cos M, sin M, 0)."""
    from tests.test_sincos import branch_arguments, build_host_trig, host_sincos
    trig = build_host_trig()
    try:
        for mode, fn, lim in (("libm", "sl_host_sincos", 1.05e8), ("cr", "sc_host_sincos", 2.0 ** 20)):
            nat.set_trig_mode(mode)
            for name, x in branch_arguments(200_000, seed=23).items():
                x = x[np.abs(x) < lim]
                one, zero = np.ones_like(x), np.zeros_like(x)
                r, _ = nat.kepler_states(x, zero, one, one, one, zero, zero, zero, max_iter=1)
                ok, s, c = host_sincos(trig, fn, x)
                assert ok == 1
                assert_bits(r[:, 0], c, f"device cos, {mode}, {name}")
                assert_bits(r[:, 1], s + 0.0, f"device sin, {mode}, {name}")     # ry = 0*cos + sin: -0 becomes +0
    finally:
        nat.set_trig_mode("libm")


def test_ensemble_generated_from_elements_on_device(nat):
    """orb_ens_upload_elements vs the oracle restatement of the host pipeline, then a short energy-conserving run."""
    from core.ensemble import EnsembleEngine
    from oracle import ref_numpy
    rng = np.random.default_rng(5)
    nsys, nb = 48, 16
    k = nb - 1
    a = np.exp(rng.uniform(np.log(0.3), np.log(30.0), (nsys, k))) * 1.495978707e11
    e = rng.uniform(0.0, 0.3, (nsys, k))
    M, Om, om = (rng.uniform(0.0, 2 * np.pi, (nsys, k)) for _ in range(3))
    inc = np.abs(rng.normal(0.0, np.deg2rad(2.0), (nsys, k)))
    m = np.concatenate([np.full((nsys, 1), 1.98847e30), np.exp(rng.uniform(np.log(1e23), np.log(1e27), (nsys, k)))], 1)
    ref = ref_numpy.ensemble_from_elements(M, e, a, inc, Om, om, m)
    for f32 in (False, True):
        eng = EnsembleEngine.from_elements(M, e, a, inc, Om, om, m, dt=8640.0, softening=1e6, vel_f32=f32)
        st = eng.state()
        for key in ("x", "y", "z"):                     # default trig mode: the host libm's bits (csrc/sincos_libm.h)
            assert_bits(st[key], ref[key], f"ensemble from elements {key}")
        for key in ("vx", "vy", "vz"):
            want = ref[key].astype(np.float32).astype(np.float64) if f32 else ref[key]
            assert_bits(st[key], want, f"ensemble from elements {key} f32={f32}")
        assert np.all(st["x"][:, 0] == 0.0) and np.all(st["vx"][:, 0] == 0.0)
        E0 = eng.energy()
        eng.step(200)
        drift = np.abs(eng.energy() - E0) / np.abs(E0)
        assert drift.max() < (1e-3 if f32 else 1e-4), drift.max()       # leapfrog, >= 600 steps per orbit
        eng.close()
    # same state uploaded from the host -> same trajectories (generation is the only difference)
    eng_a = EnsembleEngine.from_elements(M, e, a, inc, Om, om, m, dt=86400.0, softening=1e6)
    s0 = eng_a.state()
    eng_b = EnsembleEngine(s0["x"], s0["y"], s0["z"], s0["vx"], s0["vy"], s0["vz"], m, dt=86400.0, softening=1e6)
    eng_a.step(50); eng_b.step(50)
    sa, sb = eng_a.state(), eng_b.state()
    assert all(np.array_equal(sa[key], sb[key]) for key in sa)
    eng_a.close(); eng_b.close()


def test_pairwise_accelerations_operator(golden):
    """The narrow operator seam: same signature / return types as the reference function."""
    from core.physics import Coordinates, Object, pairwise_accelerations
    g = golden("force_random")
    k = 5
    objs = [Object(float(g[f"c{k}_m"][i]), 0.0, None,
                   Coordinates(float(g[f"c{k}_x"][i]), float(g[f"c{k}_y"][i]), float(g[f"c{k}_z"][i])))
            for i in range(len(g[f"c{k}_m"]))]
    acc, U = pairwise_accelerations(objs, eps=float(g[f"c{k}_eps"]))
    assert set(acc) == {o.uuid for o in objs} and acc[objs[0].uuid].shape == (3,)
    assert_bits(np.array([acc[o.uuid] for o in objs]), g[f"c{k}_acc"])
    assert_bits(U, g[f"c{k}_U"])
    acc_f, U_f = pairwise_accelerations(objs, eps=float(g[f"c{k}_eps"]), mode="fast")
    assert relerr(np.array([acc_f[o.uuid] for o in objs]), g[f"c{k}_acc"]).max() <= TOL_FAST
    assert abs(U_f - U) <= 1e-13 * abs(U)


def test_diagnostics_device_reductions(nat, orc):
    from core import synthetic
    c = synthetic.plummer(30000, seed=3)
    dev = nat.DeviceSystem(c.n, 0, nat.MODE_FAST)
    dev.set_params(c["dt"], c["eps"], G)
    dev.upload(*c.arrays())
    K, L = dev.energy_angmom()
    K_ref = float(np.sum(0.5 * c["m"] * (c["vx"] ** 2 + c["vy"] ** 2 + c["vz"] ** 2)))
    assert abs(K - K_ref) <= 1e-12 * K_ref
    pos = np.stack([c["x"], c["y"], c["z"]], 1); mom = c["m"][:, None] * np.stack([c["vx"], c["vy"], c["vz"]], 1)
    L_ref = np.cross(pos, mom).sum(0)
    assert np.linalg.norm(L - L_ref) <= 1e-10 * np.abs(np.cross(pos, mom)).sum()
    U = dev.potential()
    U_ref = orc.potential(c["x"], c["y"], c["z"], c["m"], c["eps"], G)
    assert abs(U - U_ref) <= 1e-12 * abs(U_ref)
    dev.close()


@pytest.mark.parametrize("n", [4097, 4608, 9001])
def test_potential_large_n_kernel(nat, n):
    """n > 4096: pair-once potential kernel (seed + polynomial 1/sqrt), ragged sizes and masses spread over eight
    decades.  Yardstick: rows in long double (the reference-order running sum itself is only good to ~1e-11 here:
    it drops terms below half an ulp of the partial sum)."""
    from core import synthetic
    c = synthetic.random_cloud(n, seed=n)
    pos = np.stack([c["x"], c["y"], c["z"]], 1).astype(np.longdouble)
    m = c["m"].astype(np.longdouble)
    dev = nat.DeviceSystem(c.n, 0, nat.MODE_FAST)
    for eps in (c["eps"], 0.0):
        dev.set_params(c["dt"], eps, G)
        dev.upload(*c.arrays())
        U_ref = np.longdouble(0)
        for i in range(n - 1):
            d = pos[i + 1:] - pos[i]
            U_ref -= m[i] * np.sum(m[i + 1:] / np.sqrt((d * d).sum(1) + np.longdouble(eps) ** 2))
        U_ref = float(np.longdouble(G) * U_ref)
        assert abs(dev.potential() - U_ref) <= 1e-13 * abs(U_ref), (n, eps)
    dev.close()


def test_fp64_peak_microbenchmark(nat):
    p = nat.fp64_peak(0, 0.3)
    print(f"\nFP64 DFMA peak: best {p['tflops_best']:.2f} TF, mean {p['tflops_mean']:.2f} TF at {p['sm_clock_mhz']:.0f} MHz")
    assert 10.0 < p["tflops_best"] < 60.0


def test_engine_large_n_uses_device_diagnostics(nat, orc):
    """n > 4096: auto mode selects the fast kernel; E and L come from device reductions (engine.py:104-121)."""
    from core import synthetic
    from core.engine import SimulationEngine
    from core.physics import Coordinates, Object, ObjectCollection
    c = synthetic.plummer(5000, seed=21)
    objs = []
    for i in range(c.n):
        o = Object(float(c["m"][i]), 0.0, None, Coordinates(float(c["x"][i]), float(c["y"][i]), float(c["z"][i])),
                   angular_velocity=np.zeros(3))
        o.velocity = np.array([c["vx"][i], c["vy"][i], c["vz"][i]])
        objs.append(o)
    eng = SimulationEngine(ObjectCollection(objs), dt=c["dt"], softening=c["eps"], cache=False, max_hist=-1)
    assert "force_sym_kernel" in eng.kernel_info()["name"]
    E0 = eng.total_energy()
    eng.run(20)
    E1, L1 = eng.total_energy(), eng.angular_momentum()
    pos = np.array([o.position() for o in objs]); vel = np.array([o.velocity for o in objs])
    K = float(np.sum(0.5 * c["m"] * (vel ** 2).sum(1)))
    U = orc.potential(pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy(), c["m"], c["eps"], G)
    assert abs(E1 - (K + U)) <= 1e-12 * abs(K + U)
    L_ref = np.cross(pos, c["m"][:, None] * vel).sum(0)
    assert np.linalg.norm(L1 - L_ref) <= 1e-10 * np.abs(np.cross(pos, c["m"][:, None] * vel)).sum()
    assert abs((E1 - E0) / E0) < 1e-6                       # leapfrog energy conservation over 20 steps
    eng.close()
