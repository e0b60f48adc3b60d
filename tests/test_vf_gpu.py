"""GPU parity tests of the fused visual-flocking step (through the C ABI) against
  * fixtures produced by the unmodified reference (tests/golden/vf_golden.npz), and
  * the CPU oracle (oracle/restate.py) on seeded random scenes.

Bar (BASELINE.json north_star): binary fields bit-exact except at exact bin-boundary ties
(<= 1 bin per edge, counted); dV integrals / velocities / headings within 1e-5 relative.
With the fp64 re-evaluation of near-boundary pairs (ABM_VF_EXACT_FIXUP, default) the
fields are expected to match bit for bit, and the tests assert exactly that.
"""
import numpy as np
import pytest

from golden_io import load_vf_lines_cases, load_pf_cases, load_vf_cases
from oracle import restate as rs

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # tolerance stated by north_star for integrals, velocities and headings (fp32 state)


def _engine(c_or_cfg, B, N, **kw):
    from abm_b200 import VFEngine
    return VFEngine(B, N, keep_fields=True, keep_terms=True, **kw)


# Every step kernel meets the fixtures and the oracle ITSELF (not only through kernel-vs-kernel equality): "auto" is
# the engine's own choice for the scene, the others are forced with ABM_VF_KERNEL (abm_api.cu).
KERNELS = ["auto", "symmetric", "onesided", "warp"]
KERNEL_NAMES = {"symmetric": "abm::vf_step_sym_kernel", "onesided": "abm::vf_step_kernel",
                "warp": "abm::vf_step_warp_kernel"}


def _force_kernel(monkeypatch, kernel):
    if kernel == "auto":
        monkeypatch.delenv("ABM_VF_KERNEL", raising=False)
    else:
        monkeypatch.setenv("ABM_VF_KERNEL", kernel)


def _ran_forced_kernel(eng, kernel):
    """False when the forced kernel does not apply to the scene (the symmetric kernel needs N <= 1024) -- the caller
    skips; any other mismatch is a failure."""
    if kernel == "auto":
        return True
    if eng.last_kernel() == KERNEL_NAMES[kernel]:
        return True
    assert kernel == "symmetric", f"{kernel} was forced but {eng.last_kernel()} ran"
    return False


def _check_state(st, ref, b=None):
    for k in ("x", "y", "theta", "vel"):
        got = st[k] if b is None else st[k][b]
        np.testing.assert_allclose(got.reshape(-1), ref[k], rtol=RTOL, atol=1e-5, err_msg=k)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("case", load_vf_cases(), ids=lambda c: f"N{c['N']}_R{c['R']}_{c['boundary']}")
def test_step_matches_reference_fixture(built_lib, monkeypatch, case, kernel):
    c = case
    _force_kernel(monkeypatch, kernel)
    fov = (-c["fov_ratio"] * np.pi, c["fov_ratio"] * np.pi)
    eng = _engine(None, 1, c["N"], resolution=c["R"], fov=fov, boundary=c["boundary"], width=c["W"],
                  height=c["W"], limit_movement=c["limit"])
    eng.set_params()
    if c["alp0"] is not None:
        eng.set_agent_overrides(c["alp0"], c["bet0"], c["v0"])
    eng.set_state(c["x"], c["y"], c["theta"], c["vel"], c["radius"])
    eng.step(1)
    if not _ran_forced_kernel(eng, kernel):
        eng.close()
        pytest.skip("the symmetric kernel does not apply to this scene (heterogeneous radii)")
    assert np.array_equal(eng.fields_packed()[0], c["fields"])          # bit-exact stored fields
    np.testing.assert_allclose(eng.terms()[0], c["terms"], rtol=RTOL, atol=1e-9)
    st = eng.get_state()
    got = np.stack([st["x"][0], st["y"][0], st["theta"][0], st["vel"][0]], axis=1)
    np.testing.assert_allclose(got, c["new"], rtol=RTOL, atol=1e-5)
    eng.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("case", load_vf_lines_cases(), ids=lambda c: f"N{c['N']}_R{c['R']}_{c['boundary']}")
def test_line_following_matches_reference_fixture(built_lib, monkeypatch, case, kernel):
    """agents with lines to follow (vf_agent.py:273-276 -> vf_supcalc.follow_lines_local, :293-328; turn-back
    wall rule vf_agent.py:80-129) through abm_vf_set_line_map, under every step kernel, against fixtures of the
    unmodified reference; switching the map off again gives the flocking step."""
    c = case
    _force_kernel(monkeypatch, kernel)
    fov = (-c["fov_ratio"] * np.pi, c["fov_ratio"] * np.pi)
    eng = _engine(None, 1, c["N"], resolution=c["R"], fov=fov, boundary=c["boundary"], width=c["W"],
                  height=c["W"], limit_movement=c["limit"])
    eng.set_params()
    eng.set_line_map(c["line_map"], sensor_radius=9, sensor_distance=20)
    eng.set_state(c["x"], c["y"], c["theta"], c["vel"], c["radius"])
    eng.step(1)
    assert _ran_forced_kernel(eng, kernel)
    assert np.array_equal(eng.fields_packed()[0], c["fields"])          # bit-exact stored fields
    st = eng.get_state()
    got = np.stack([st["x"][0], st["y"][0], st["theta"][0], st["vel"][0]], axis=1)
    np.testing.assert_allclose(got, c["new"], rtol=RTOL, atol=1e-5)
    eng.set_line_map(None)
    eng.set_state(c["x"], c["y"], c["theta"], c["vel"], c["radius"])
    eng.step(1)
    cfg = rs.VFConfig(R=c["R"], fov=fov, boundary=c["boundary"], width=c["W"], height=c["W"], limit_movement=c["limit"])
    plain = rs.vf_step_frozen(c["x"], c["y"], c["theta"], c["vel"], c["radius"], cfg)
    np.testing.assert_allclose(eng.get_state()["theta"][0], plain["theta"], rtol=RTOL, atol=1e-5)
    eng.close()


@pytest.mark.parametrize("pc", load_pf_cases(), ids=lambda p: f"R{p['R']}_{p['boundary']}")
def test_projection_field_function_matches_reference_fixture(built_lib, pc):
    from abm_b200 import vf_supcalc
    rows = vf_supcalc.projection_field(pc["fov"], pc["R"], np.array(pc["pos"]), pc["r"], pc["th"],
                                       [np.array(o) for o in pc["objs"]], object_sizes=pc["sizes"],
                                       boundary_cond=pc["boundary"], arena_width=pc["W"], arena_height=pc["W"],
                                       vision_range=pc["vr"])
    assert rows.shape == (len(pc["objs"]), pc["R"])
    assert np.array_equal(rs.pack_bits(rows > 0), pc["rows"])


def test_reference_golden_vector_through_plugin(built_lib):
    """test_cs_supcalc.py:143-158, verbatim through the drop-in function."""
    from abm_b200 import vf_supcalc
    out = vf_supcalc.projection_field((-np.pi, np.pi), 8, np.array([-1, -1]), 1, 0, [np.array([0, -1])])
    assert out.tolist() == [[0, 0, 0, 0, 1, 1, 0, 0]]


def test_flocking_terms_function_kat(built_lib):
    """KAT-V1 of SURVEY App. B through VSWRM_flocking_state_variables."""
    from types import SimpleNamespace
    from abm_b200 import vf_supcalc
    field = np.zeros(1200)
    for a, b in [(130, 160), (473, 511), (754, 828), (1060, 1090)]:
        field[a:b] = 1
    prm = SimpleNamespace(GAM=0.1, V0=1, ALP0=1, ALP1=0.09, BET0=1, BET1=0.09, ALP2=0, BET2=0)
    Phi = np.arange(-np.pi, np.pi, 2 * np.pi / 1200)
    got = vf_supcalc.VSWRM_flocking_state_variables(0.5, Phi, np.flip(field), prm, verbose=True)
    expect = (-0.11295710879884255, 0.16720418209113855, -0.13687988331415218, -0.02607722548469037,
              0.20545193082925006, -0.038247748738111514)
    np.testing.assert_allclose(got, expect, rtol=1e-11)
    dv, dpsi = vf_supcalc.VSWRM_flocking_state_variables(0.5, Phi, np.flip(field), prm)
    assert np.isclose(dv, expect[0]) and np.isclose(dpsi, expect[1])


def _random_scene(rng, B, N, W, spread=None):
    pad = 30.0
    lo, hi = (pad - 10, pad + W - 10) if spread is None else spread
    x = rng.uniform(lo, hi, (B, N)).astype(np.float32)
    y = rng.uniform(lo, hi, (B, N)).astype(np.float32)
    th = rng.uniform(0, 2 * np.pi, (B, N)).astype(np.float32)
    v = rng.uniform(0, 2.5, (B, N)).astype(np.float32)
    return x, y, th, v


SCENES = [
    # B, N, R, W, boundary, fov_ratio, radius
    (3, 50, 1200, 500.0, "walls", 1.0, 10.0),
    (2, 100, 1200, 900.0, "walls", 1.0, 10.0),       # config 2 shape
    (2, 100, 1200, 900.0, "infinite", 1.0, 5.0),
    (2, 70, 2400, 600.0, "walls", 0.5, 10.0),
    (2, 33, 1201, 300.0, "infinite", 1.0, 10.0),     # odd R
    (1, 300, 1200, 1559.0, "walls", 1.0, 10.0),      # more than one CTA per replicate
    (4, 10, 320, 900.0, "walls", 1.0, 10.0),
    (1, 600, 1200, 400.0, "walls", 1.0, 10.0),       # crowded, > one record stage
    (2, 90, 1200, 500.0, "walls", 1.0, 0.0),         # unequal radii (0: drawn per agent): two half widths per pair
    (2, 130, 1200, 600.0, "infinite", 1.0, 0.0),     # ... on the torus, crowded
    (1, 75, 2400, 450.0, "walls", 0.5, 0.0),         # ... limited FOV
    (2, 50, 8, 300.0, "walls", 1.0, 10.0),           # the smallest ring the engine accepts (the reference's golden vector)
    (2, 40, 33, 300.0, "infinite", 1.0, 10.0),       # one word and a bit, odd
    (1, 60, 4800, 500.0, "walls", 1.0, 10.0),        # 150 words per row
    (1, 1100, 1200, 1300.0, "walls", 1.0, 10.0),     # more agents than the symmetric kernel holds (1024), dense
    (1, 1030, 1200, 1300.0, "infinite", 1.0, 0.0),   # ... unequal radii, torus
]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("B,N,R,W,boundary,fovr,radius", SCENES)
def test_step_matches_oracle_random(built_lib, monkeypatch, B, N, R, W, boundary, fovr, radius, kernel):
    _force_kernel(monkeypatch, kernel)
    rng = np.random.default_rng(1234 + N + R)
    x, y, th, v = _random_scene(rng, B, N, W)
    if radius == 0.0:                                                    # every agent its own radius
        radius = rng.choice([4.0, 7.5, 10.0, 16.0], (B, N)).astype(np.float32)
    fov = (-fovr * np.pi, fovr * np.pi)
    eng = _engine(None, B, N, resolution=R, fov=fov, boundary=boundary, width=W, height=W)
    eng.set_params()
    eng.set_state(x, y, th, v, radius)
    eng.step(1)
    if not _ran_forced_kernel(eng, kernel):                              # N <= 1024: every kernel applies ...
        assert R < 64 or N > 1024                                        # ... but on a tiny ring the zero-width distance is
        eng.close()                                                      # a few radii (distance culling: no symmetric kernel),
        pytest.skip("the symmetric kernel does not apply to this scene")  # and it holds at most 1024 agents
    fields, terms, st = eng.fields(), eng.terms(), eng.get_state()
    cfg = rs.VFConfig(R=R, fov=fov, boundary=boundary, width=W, height=W)
    sample = range(N) if N <= 130 else sorted(rng.choice(N, 60, replace=False).tolist())
    for b in range(B):
        ref = rs.vf_step_frozen(x[b], y[b], th[b], v[b], radius if np.ndim(radius) == 0 else radius[b], cfg, agents=sample)
        idx = np.array(list(sample))
        assert np.array_equal(fields[b][idx], ref["rows"][idx][:, ::-1]), "stored field mismatch"
        np.testing.assert_allclose(terms[b][idx], ref["terms"][idx], rtol=RTOL, atol=1e-9)
        for k in ("x", "y", "theta", "vel"):
            np.testing.assert_allclose(st[k][b][idx], ref[k][idx], rtol=RTOL, atol=1e-5, err_msg=k)
    cnt = eng.counters()
    assert cnt["launches"] == 1
    eng.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("boundary", ["walls", "infinite"])
def test_degenerate_geometry_matches_oracle(built_lib, monkeypatch, boundary, kernel):
    """The corners of the pair geometry in one scene, against the oracle under every kernel: coincident agents (skipped,
    vf_supcalc.py:57), overlapping discs (distance < radius: atan(r / d) near pi / 2, intervals hundreds of bins wide that
    wrap around the ring), touching discs, agents dead ahead / dead astern / exactly abeam (bin ties and the +-pi seam),
    agents on and beyond the walls, standing agents, headings of exactly 0 and 2 pi."""
    _force_kernel(monkeypatch, kernel)
    W, R, r = 300.0, 1200, 10.0
    base = np.array([[150, 150], [150, 150], [150, 150],          # three coincident agents
                     [150.5, 150], [153, 150], [150, 159.99], [160, 150], [170, 150],   # overlapping / touching / near
                     [150, 100], [150, 200], [100, 150], [200, 150],                    # abeam / ahead / astern of agent 0
                     [20, 20], [20, 150], [329, 329], [335, 150], [150, 15], [10, 340],  # at and beyond the walls
                     [31, 31], [299, 299], [60, 240], [240, 60]], np.float32)
    N = len(base)
    rng = np.random.default_rng(5)
    x = np.stack([base[:, 0], base[::-1, 0]]); y = np.stack([base[:, 1], base[::-1, 1]])
    th = rng.uniform(0, 2 * np.pi, (2, N)).astype(np.float32)
    th[0, :4] = [0.0, np.float32(2 * np.pi), np.float32(np.pi), np.float32(np.pi / 2)]
    v = rng.uniform(0, 2.5, (2, N)).astype(np.float32)
    v[:, ::5] = 0.0
    eng = _engine(None, 2, N, resolution=R, boundary=boundary, width=W, height=W)
    eng.set_params()
    eng.set_state(x, y, th, v, r)
    eng.step(1)
    assert _ran_forced_kernel(eng, kernel)
    fields, terms, st = eng.fields(), eng.terms(), eng.get_state()
    cfg = rs.VFConfig(R=R, boundary=boundary, width=W, height=W)
    widest = 0
    for b in range(2):
        ref = rs.vf_step_frozen(x[b], y[b], th[b], v[b], r, cfg)
        assert np.array_equal(fields[b], ref["rows"][:, ::-1]), "stored field mismatch"
        np.testing.assert_allclose(terms[b], ref["terms"], rtol=RTOL, atol=1e-9)
        _check_state(st, ref, b)
        widest = max(widest, int(ref["rows"].sum(axis=1).max()))
    assert widest > 600                                  # an overlapping neighbour fills more than half the ring
    eng.close()


def test_invalid_calls_fail_loudly(built_lib):
    """Error behaviour of the boundary: wrong sizes, calls in the wrong order and impossible configurations raise
    (ABM_E_* codes with a message from abm_last_error), nothing is computed on a fallback path."""
    from abm_b200 import VFEngine, _lib
    eng = VFEngine(2, 8, resolution=1200, width=300.0, height=300.0)
    with pytest.raises(_lib.AbmError):
        eng.step(1)                                      # no state yet
    with pytest.raises(_lib.AbmError):
        eng.get_state()
    x = np.zeros((2, 8), np.float32)
    with pytest.raises(Exception):
        eng.set_state(x[:, :4], x, x, x, 10.0)           # wrong shape
    eng.set_params()
    eng.set_state(x + 100, x + 100, x, x, 10.0)
    with pytest.raises(_lib.AbmError):
        eng.step(-1)
    with pytest.raises(Exception):
        eng.set_line_map(np.zeros(5, np.float32))        # not two-dimensional
    with pytest.raises(Exception):
        eng.step_host(np.zeros((2, 8, 3), np.float32), np.zeros((2, 8, 4), np.float32))
    eng.step(1)                                          # (all agents coincident: nothing visible, still a valid step)
    assert np.isfinite(eng.get_state()["x"]).all()
    eng.close()
    with pytest.raises(Exception):
        VFEngine(1, 8, resolution=1, width=300.0, height=300.0)      # a ring of one bin
    with pytest.raises(Exception):
        VFEngine(0, 8, resolution=1200, width=300.0, height=300.0)


@pytest.mark.parametrize("boundary,fovr", [("walls", 1.0), ("infinite", 1.0), ("walls", 0.6)])
def test_three_word_variant_with_unequal_radii_matches_oracle(built_lib, monkeypatch, boundary, fovr):
    """The symmetric kernel's three-word fast path (ABM_VF_SYM_WIDE=1, otherwise chosen in crowded scenes) together with
    its unequal-radii variant, in a crowded scene where most intervals are wider than 32 bins: against the oracle."""
    monkeypatch.setenv("ABM_VF_KERNEL", "symmetric")
    monkeypatch.setenv("ABM_VF_SYM_WIDE", "1")
    rng = np.random.default_rng(77)
    B, N, W = 2, 120, 260.0
    R = int(1200 / fovr)
    x, y, th, v = _random_scene(rng, B, N, W)
    rad = rng.choice([5.0, 9.0, 12.0, 18.0], (B, N)).astype(np.float32)
    fov = (-fovr * np.pi, fovr * np.pi)
    eng = _engine(None, B, N, resolution=R, fov=fov, boundary=boundary, width=W, height=W)
    eng.set_params(); eng.set_state(x, y, th, v, rad); eng.step(1)
    assert eng.last_kernel() == "abm::vf_step_sym_kernel" and eng.kernel_stats()["symmetric_wide"] == 1
    fields, terms, st = eng.fields(), eng.terms(), eng.get_state()
    cfg = rs.VFConfig(R=R, fov=fov, boundary=boundary, width=W, height=W)
    for b in range(B):
        ref = rs.vf_step_frozen(x[b], y[b], th[b], v[b], rad[b], cfg)
        assert np.array_equal(fields[b], ref["rows"][:, ::-1]), "stored field mismatch"
        np.testing.assert_allclose(terms[b], ref["terms"], rtol=RTOL, atol=1e-9)
        _check_state(st, ref, b)
    eng.close()


def test_per_replicate_parameter_sweep(built_lib):
    """One launch, one parameter set per replicate (the MetaProtocol sweep shape)."""
    rng = np.random.default_rng(5)
    B, N, R, W = 6, 24, 1200, 400.0
    x, y, th, v = _random_scene(rng, B, N, W)
    alp0 = np.linspace(0, 5, B); bet0 = np.linspace(5, 0, B)
    eng = _engine(None, B, N, resolution=R, width=W, height=W)
    eng.set_params(ALP0=alp0, BET0=bet0, ALP1=0.0014, BET1=0.0014)
    eng.set_state(x, y, th, v, 10.0)
    eng.step(1)
    terms, st = eng.terms(), eng.get_state()
    for b in range(B):
        cfg = rs.VFConfig(R=R, width=W, height=W, ALP0=alp0[b], BET0=bet0[b], ALP1=0.0014, BET1=0.0014)
        ref = rs.vf_step_frozen(x[b], y[b], th[b], v[b], 10.0, cfg)
        np.testing.assert_allclose(terms[b], ref["terms"], rtol=RTOL, atol=1e-9)
        _check_state(st, ref, b)
    eng.close()


def test_multi_step_lockstep_with_oracle(built_lib):
    """20 steps; after every step the oracle is restarted from the engine's fp32 state, so
    each step is judged on its own from identical state (no drift accumulation)."""
    rng = np.random.default_rng(11)
    B, N, R, W = 1, 40, 1200, 300.0
    x, y, th, v = _random_scene(rng, B, N, W)
    eng = _engine(None, B, N, resolution=R, width=W, height=W)
    eng.set_params()
    eng.set_state(x, y, th, v, 10.0)
    cfg = rs.VFConfig(R=R, width=W, height=W)
    cur = dict(x=x, y=y, theta=th, vel=v)
    for _ in range(20):
        ref = rs.vf_step_frozen(cur["x"][0], cur["y"][0], cur["theta"][0], cur["vel"][0], 10.0, cfg)
        eng.step(1)
        st = eng.get_state()
        assert np.array_equal(eng.fields()[0], ref["rows"][:, ::-1])
        _check_state(st, ref, 0)
        cur = st
    assert eng.counters()["launches"] == 20
    eng.close()


def test_device_pointer_state_roundtrip(built_lib):
    import torch
    rng = np.random.default_rng(3)
    B, N = 2, 64
    x, y, th, v = _random_scene(rng, B, N, 500.0)
    from abm_b200 import VFEngine
    eng = VFEngine(B, N, resolution=1200, width=500.0, height=500.0)
    dev = [torch.from_numpy(a).cuda() for a in (x, y, th, v)]
    rad = torch.full((B, N), 10.0, device="cuda")
    eng.set_state(*dev, rad)
    out = {k: torch.empty(B, N, device="cuda") for k in ("x", "y", "theta", "vel")}
    eng.get_state(out)
    torch.cuda.synchronize()
    for k, a in zip(("x", "y", "theta", "vel"), (x, y, th, v)):
        assert np.array_equal(out[k].cpu().numpy(), a)
    eng.close()


def test_empty_field_and_single_agent(built_lib):
    """N=1: nothing to see -> field empty, integrals zero, agent relaxes towards V0."""
    eng = _engine(None, 2, 1, resolution=1200, width=300.0, height=300.0)
    eng.set_params()
    eng.set_state(np.array([[100.0], [150.0]]), np.array([[100.0], [90.0]]), np.array([[0.5], [4.0]]),
                  np.array([[0.0], [2.0]]), 10.0)
    eng.step(1)
    assert not eng.fields().any()
    t = eng.terms()
    np.testing.assert_allclose(t[:, 0, 0], [0.1 * (1 - 0.0), 0.1 * (1 - 2.0)], rtol=1e-12)
    assert np.all(t[:, 0, 1:] == 0)
    eng.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_full_size_properties(built_lib, monkeypatch, kernel):
    """BASELINE configs[3] shape at reduced replicate count (1024 agents x 8 replicates, R=1200), under every step
    kernel (the benchmark's own, `symmetric`, forced at this batch size):
    (1) replicates are independent: replicate b of a batch == the same replicate alone;
    (2) agent order inside the neighbour table does not matter (union is commutative);
    (3) the oracle on EVERY agent of one replicate of the full-size scene."""
    _force_kernel(monkeypatch, kernel)
    rng = np.random.default_rng(2024)
    B, N, R = 8, 1024, 1200
    W = float(np.ceil(900 * np.sqrt(N / 100)))
    x, y, th, v = _random_scene(rng, B, N, W)
    from abm_b200 import VFEngine
    eng = VFEngine(B, N, resolution=R, width=W, height=W, keep_fields=True, keep_terms=True)
    eng.set_params(); eng.set_state(x, y, th, v, 10.0); eng.step(1)
    assert _ran_forced_kernel(eng, kernel)
    f_all, t_all, st_all = eng.fields_packed(), eng.terms(), eng.get_state()
    eng.close()
    # (1)
    e1 = VFEngine(1, N, resolution=R, width=W, height=W, keep_fields=True, keep_terms=True)
    e1.set_params(); e1.set_state(x[3:4], y[3:4], th[3:4], v[3:4], 10.0); e1.step(1)
    assert np.array_equal(e1.fields_packed()[0], f_all[3])
    assert np.array_equal(e1.terms()[0], t_all[3])
    # (2) permute agents of replicate 3
    perm = rng.permutation(N)
    e1.set_state(x[3:4, perm], y[3:4, perm], th[3:4, perm], v[3:4, perm], 10.0); e1.step(1)
    assert np.array_equal(e1.fields_packed()[0], f_all[3][perm])
    # (3) the oracle on all 1024 agents of replicate 3 of the full-size scene
    cfg = rs.VFConfig(R=R, width=W, height=W)
    ref = rs.vf_step_frozen(x[3], y[3], th[3], v[3], 10.0, cfg)
    assert np.array_equal(rs.unpack_bits(f_all[3], R), ref["rows"][:, ::-1])
    np.testing.assert_allclose(t_all[3], ref["terms"], rtol=RTOL, atol=1e-9)
    _check_state(st_all, ref, 3)
    e1.close()


def test_fixup_counters_and_fp32_only_mode(built_lib):
    """With the fp64 re-evaluation switched off the fp32 path alone may differ from the oracle
    only by <= 1 bin per edge; the number of such agents is counted and reported."""
    rng = np.random.default_rng(77)
    B, N, R, W = 4, 256, 1200, 1440.0
    x, y, th, v = _random_scene(rng, B, N, W)
    from abm_b200 import VFEngine
    exact = VFEngine(B, N, resolution=R, width=W, height=W, keep_fields=True)
    fast = VFEngine(B, N, resolution=R, width=W, height=W, keep_fields=True, exact_fixup=False)
    for e in (exact, fast):
        e.set_params(); e.set_state(x, y, th, v, 10.0); e.step(1)
    fe, ff = exact.fields(), fast.fields()
    cnt = exact.counters()
    pairs = B * N * (N - 1)
    assert 0 < cnt["fp64_pairs"] < 0.02 * pairs          # a small fraction goes through fp64
    assert cnt["fp32_fp64_differ"] <= cnt["fp64_pairs"]
    diff_agents = int((fe != ff).any(axis=-1).sum())
    # an fp32-only field may differ from the exact one only by interval ends moved by ONE bin: every
    # differing bin is isolated (its ring neighbours agree) and sits on an end of one of the focal
    # agent's per-object intervals (oracle), whether or not that end is an edge of the union field
    d = (fe != ff)
    assert not (d & np.roll(d, 1, axis=-1)).any(), "differences wider than one bin"
    cfg = rs.VFConfig(R=R, width=W, height=W)
    for b, i in zip(*np.nonzero(d.any(axis=-1))):
        iv = rs.vf_intervals(x[b].astype(np.float64), y[b].astype(np.float64), 10.0, th[b].astype(np.float64), int(i), cfg)
        ps, pe = iv["ps"][iv["valid"]], iv["pe"][iv["valid"]]
        ends = np.concatenate([ps - 1, ps, pe - 1, pe, (pe + 1)[pe >= R - 1]]) % R
        allowed = np.zeros(R, bool); allowed[ends] = True
        bad = d[b, i][::-1] & ~allowed                      # stored fields are flipped (vf_supcalc.py:134)
        assert not bad.any(), f"replicate {b} agent {i}: difference away from an interval end at bins {np.nonzero(bad)[0]}"
    assert diff_agents <= 0.2 * B * N
    print(f"fp64 pairs {cnt['fp64_pairs']} of {pairs} ({cnt['fp64_pairs'] / pairs:.2e}); fp32/fp64 index "
          f"differences {cnt['fp32_fp64_differ']}; agents whose fp32-only field differs: {diff_agents} of {B * N}")
    exact.close(); fast.close()


def test_spatial_sort_is_invisible(built_lib, monkeypatch):
    """The internal Morton ordering (refreshed every few steps; used by the one-thread-per-focal-agent
    kernel) must not change any result: sorted and unsorted engines agree bit for bit after many
    steps, in the caller's agent order, including fields, terms and per-agent overrides set before
    and after the first sort."""
    from abm_b200 import VFEngine
    monkeypatch.setenv("ABM_VF_KERNEL", "onesided")
    rng = np.random.default_rng(41)
    B, N, R, W = 3, 300, 1200, 900.0
    x, y, th, v = _random_scene(rng, B, N, W)
    alp0 = rng.uniform(0.5, 2.0, (B, N)).astype(np.float32); alp0[:, ::3] = np.nan
    res = {}
    for sort in (False, True):
        eng = VFEngine(B, N, resolution=R, width=W, height=W, keep_fields=True, keep_terms=True,
                       spatial_sort=sort, resort_every=4)
        eng.set_params(); eng.set_state(x, y, th, v, 10.0)
        eng.set_agent_overrides(alp0=alp0)
        eng.step(7)
        eng.set_agent_overrides(alp0=alp0, v0=alp0)          # now in a permuted internal order
        eng.step(6)
        res[sort] = (eng.get_state(), eng.fields_packed(), eng.terms(), eng.permutation())
        eng.close()
    for k in ("x", "y", "theta", "vel"):
        assert np.array_equal(res[False][0][k], res[True][0][k]), k
    assert np.array_equal(res[False][1], res[True][1])
    assert np.array_equal(res[False][2], res[True][2])
    assert np.array_equal(res[False][3], np.tile(np.arange(N, dtype=np.int32), (B, 1)))
    assert not np.array_equal(res[True][3], res[False][3])
    assert np.array_equal(np.sort(res[True][3], axis=1), res[False][3])


@pytest.mark.parametrize("B,N,R,boundary,fov_ratio", [
    (3, 1024, 1200, "walls", 1.0),      # the benchmark shape: no padding, compile-time R
    (2, 300, 1200, "infinite", 1.0),    # padded blocks, torus
    (2, 200, 1201, "walls", 1.0),       # odd R (run-time constants)
    (2, 130, 640, "walls", 0.5),        # limited FOV, small R
    (5, 40, 1200, "walls", 1.0),        # a single block pair
    (1, 700, 2400, "infinite", 0.75),   # large R
])
@pytest.mark.parametrize("radii", ["equal", "unequal"])
def test_symmetric_and_onesided_kernels_agree(built_lib, monkeypatch, B, N, R, boundary, fov_ratio, radii):
    """The three step kernels (every unordered pair once / one thread per focal agent / one warp per focal agent) are
    independent implementations of the same step; with the fp64 re-evaluation on, both must produce the exact
    fields, so they agree bit for bit -- fields, terms and new state -- also over several steps,
    crowded scenes (wide intervals, overlapping agents) included."""
    from abm_b200 import VFEngine
    rng = np.random.default_rng(1000 + N)
    W = 60.0 * np.sqrt(N)                     # crowded: neighbours at a few radii
    x, y, th, v = _random_scene(rng, B, N, W)
    x[0, 1], y[0, 1] = x[0, 0], y[0, 0]       # coincident pair (vf_supcalc.py:57)
    x[0, 3], y[0, 3] = x[0, 2] + 3.0, y[0, 2] # overlapping pair (d < r)
    rad = 10.0 if radii == "equal" else rng.choice([5.0, 10.0, 12.5, 20.0], (B, N)).astype(np.float32)
    fov = (-fov_ratio * np.pi, fov_ratio * np.pi)
    res = {}
    for kern in ("onesided", "symmetric", "warp"):
        monkeypatch.setenv("ABM_VF_KERNEL", kern)
        eng = VFEngine(B, N, resolution=R, width=W, height=W, boundary=boundary, fov=fov, keep_fields=True,
                       keep_terms=True, spatial_sort=False)
        eng.set_params(); eng.set_state(x, y, th, v, rad)
        eng.step(1)
        assert _ran_forced_kernel(eng, kern) or R == 2400     # (700 rows of 2400 bins do not fit the symmetric kernel's shared memory)
        f1, t1 = eng.fields_packed().copy(), eng.terms().copy()
        eng.step(3)
        res[kern] = (f1, t1, eng.fields_packed(), eng.terms(), eng.get_state(), eng.counters())
        eng.close()
    a = res["onesided"]
    b = res["symmetric"]                      # same epilogue code: everything bit for bit, also after further steps
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
    for k in ("x", "y", "theta", "vel"):
        assert np.array_equal(a[4][k], b[4][k]), k
    b = res["warp"]                           # its edge sums are reduced across lanes: same fields, terms to rounding
    assert np.array_equal(a[0], b[0])
    np.testing.assert_allclose(b[1], a[1], rtol=1e-11, atol=1e-12)
    for k in ("x", "y", "theta", "vel"):
        np.testing.assert_allclose(b[4][k], a[4][k], rtol=1e-5, atol=1e-4, err_msg=k)
    print("fp64 pairs onesided / symmetric / warp:", *(res[k][5]["fp64_pairs"] for k in ("onesided", "symmetric", "warp")))


@pytest.mark.parametrize("boundary", ["walls", "infinite"])
def test_summary_metrics_match_numpy(built_lib, boundary):
    """On-device per-replicate summary metrics against the formulas of the reference's offline
    analysis (data_loader.py: polarization :1817-1823, inter-individual distance restricted to the upper
    triangle :1440-1452 with supcalc.distance_torus on the torus, nearest-neighbour distance :1480-1485,
    collision criterion iid < 2 * RADIUS_AGENT :1860-1861), evaluated in float64 numpy on the same state."""
    from abm_b200 import VFEngine
    rng = np.random.default_rng(5)
    B, N, W = 6, 173, 700.0
    x, y, th, v = _random_scene(rng, B, N, W)
    x[1] = rng.uniform(300, 340, N); y[1] = rng.uniform(300, 340, N)        # a crowded replicate: collisions
    spread = np.linspace(60, 640, N).astype(np.float32)
    x[2], y[2] = spread, spread[::-1].copy()                                 # a sparse one: none
    th[3] = 1.0                                                              # fully polarized
    eng = VFEngine(B, N, resolution=1200, width=W, height=W, boundary=boundary)
    eng.set_params(); eng.set_state(x, y, th, v, 10.0)
    for _ in range(2):                                                       # before and after a step
        st = eng.get_state()
        m = eng.metrics()
        for b in range(B):
            X, Y, T = (st[k][b].astype(np.float64) for k in ("x", "y", "theta"))
            pol = np.hypot(np.cos(T).sum(), np.sin(T).sum()) / N
            dx = np.abs(X[:, None] - X[None, :]); dy = np.abs(Y[:, None] - Y[None, :])
            if boundary == "infinite":
                dx = np.where(dx > W / 2, W - dx, dx); dy = np.where(dy > W / 2, W - dy, dy)
            d = np.hypot(dx, dy)
            iu = np.triu_indices(N, k=1)
            dn = d.copy(); np.fill_diagonal(dn, np.inf)
            np.testing.assert_allclose(m["polarization"][b], pol, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(m["mean_iid"][b], d[iu].mean(), rtol=1e-5)
            np.testing.assert_allclose(m["mean_nn_dist"][b], dn.min(axis=1).mean(), rtol=1e-5)
            assert m["collision"][b] == float(((d[iu] > 0) & (d[iu] < 20.0)).any())
        eng.step(1)
    assert m["collision"][1] == 1.0 and m["polarization"].max() <= 1.0 + 1e-6
    eng.close()


def test_kernel_choice_adapts_to_crowding(built_lib, monkeypatch):
    """The step kernel is chosen per step: the symmetric kernel reports how many pairs left its (two-word) fast path and
    a crowded scene (many intervals wider than 32 bins) switches to its three-word fast path -- without changing any
    result (every variant is exact)."""
    import torch
    from abm_b200 import VFEngine
    monkeypatch.delenv("ABM_VF_KERNEL", raising=False)
    rng = np.random.default_rng(11)
    B, N, W = 160, 256, 1500.0         # (a batch large enough for one CTA per replicate to fill the GPU, see below)
    res = {}
    for name, spread in (("crowded", 60.0), ("sparse", 700.0)):
        ang = rng.uniform(0, 2 * np.pi, (B, N)); rr = np.sqrt(rng.uniform(0, 1, (B, N))) * spread
        x = (750 + rr * np.cos(ang)).astype(np.float32); y = (750 + rr * np.sin(ang)).astype(np.float32)
        th = rng.uniform(0, 2 * np.pi, (B, N)).astype(np.float32); v = np.zeros_like(x)
        eng = VFEngine(B, N, resolution=1200, width=W, height=W, keep_fields=True)
        eng.set_params(); eng.set_state(x, y, th, v, 10.0)
        eng.step(1)
        assert eng.last_kernel() == "abm::vf_step_sym_kernel"
        entries, launches = eng.slow_entries()
        frac = entries / launches / (0.5 * B * N * (N - 1))            # of the unordered pairs
        torch.cuda.synchronize()
        for _ in range(4):
            eng.step(1); torch.cuda.synchronize()
        res[name] = (frac, eng.kernel_stats(), eng.fields_packed().copy(), eng.get_state())
        eng.close()
        for wide in ("0", "1"):                                   # the same five steps with ONE variant throughout
            monkeypatch.setenv("ABM_VF_SYM_WIDE", wide)
            ref = VFEngine(B, N, resolution=1200, width=W, height=W, keep_fields=True)
            ref.set_params(); ref.set_state(x, y, th, v, 10.0); ref.step(5)
            assert ref.kernel_stats()["symmetric_wide" if wide == "1" else "symmetric"] == 5
            assert np.array_equal(ref.fields_packed(), res[name][2])
            for k in ("x", "y", "theta", "vel"):
                assert np.array_equal(ref.get_state()[k], res[name][3][k])
            ref.close()
        monkeypatch.delenv("ABM_VF_SYM_WIDE")
    assert res["crowded"][0] > 0.04 and res["crowded"][1]["symmetric_wide"] >= 3
    assert res["sparse"][0] < 0.04 and res["sparse"][1] == dict(symmetric=5, symmetric_wide=0, onesided=0, warp=0)
    # a batch that cannot fill the GPU with one CTA per replicate runs with a warp per focal agent -- same results
    small = VFEngine(2, N, resolution=1200, width=W, height=W, keep_fields=True)
    small.set_params(); small.set_state(x[:2], y[:2], th[:2], v[:2], 10.0); small.step(5)
    assert small.last_kernel() == "abm::vf_step_warp_kernel"
    assert np.array_equal(small.fields_packed(), res["sparse"][2][:2])
    for k in ("x", "y", "theta", "vel"):
        np.testing.assert_allclose(small.get_state()[k], res["sparse"][3][k][:2], rtol=1e-6, atol=1e-6)
    small.close()


@pytest.mark.parametrize("kernel", ["symmetric", "onesided", "warp"])
def test_torus_neighbour_across_the_seam_near_a_tie(built_lib, monkeypatch, kernel):
    """Regression (found by scratch/soak_compare.py, 2 wrong bins in 1e11 directions): on the torus the minimal-image
    difference must be rounded at the magnitude of the RESULT -- fl(x_j - x_i) of coordinates an arena apart carries
    ulp(arena) / 2, which for a close neighbour across the periodic seam exceeded the fp32 guard band of the bin index.
    Second cause, second scene: a neighbour exactly half an arena away in fp32 (|dy| = 1440.00003 in the reference's
    float64) -- the wrap decision cannot be taken from the rounded difference; such pairs go to the fp64 path.
    The scenes (tests/golden/torus_seam_cases.npz) are states of the soak run that exposed them."""
    import os
    from abm_b200 import VFEngine
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "torus_seam_cases.npz"))
    W = float(d["W"])
    monkeypatch.setenv("ABM_VF_KERNEL", kernel)
    for c in (0, 1):
        x, y, th, v = (d[f"{k}{c}"][None, :] for k in ("x", "y", "theta", "vel"))
        i = int(d[f"agent{c}"])
        eng = VFEngine(1, x.shape[1], resolution=1200, width=W, height=W, boundary="infinite", keep_fields=True)
        eng.set_params(); eng.set_state(x, y, th, v, 10.0); eng.step(1)
        cfg = rs.VFConfig(R=1200, width=W, height=W, boundary="infinite")
        idx = sorted({i, 0, 17, 500})
        ref = rs.vf_step_frozen(x[0], y[0], th[0], v[0], 10.0, cfg, agents=idx)
        assert np.array_equal(eng.fields()[0][idx], ref["rows"][idx][:, ::-1])
        eng.close()


def test_pinned_async_state_transfers_two_engines_two_streams(built_lib):
    """ABM_HOST_PINNED_ASYNC (on_device == 2): abm_set_state / abm_get_state with pinned host buffers only enqueue on
    the stream.  Two engines on two streams, each fed from and drained into its own pinned buffers every step (the
    pipelined end-to-end arm of bench.py), end in exactly the state of the blocking calls."""
    import torch
    from abm_b200 import VFEngine
    rng = np.random.default_rng(21)
    B, N, W, T = 4, 300, 1500.0, 6
    def pinned(a):
        return torch.from_numpy(np.ascontiguousarray(a, np.float32)).pin_memory().numpy()
    scenes = [_random_scene(rng, B, N, W) for _ in range(2)]
    ref = []
    for x, y, th, v in scenes:                                   # blocking reference: state through host buffers every step
        eng = VFEngine(B, N, resolution=1200, width=W, height=W)
        eng.set_params(); st = dict(x=x, y=y, theta=th, vel=v)
        for t in range(T):
            eng.set_state(st["x"], st["y"], st["theta"], st["vel"], 10.0 if t == 0 else None)
            eng.step(1); st = eng.get_state()
        ref.append(st); eng.close()
    engs = [VFEngine(B, N, resolution=1200, width=W, height=W) for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    bufs = [[{k: pinned(a) for k, a in zip(("x", "y", "theta", "vel"), sc)} for _ in range(2)] for sc in scenes]
    rad = pinned(np.full((B, N), 10.0))
    cur = [0, 0]
    for k in (0, 1):
        engs[k].set_params()
    for t in range(T):
        for k in (0, 1):
            with torch.cuda.stream(streams[k]):
                h = bufs[k][cur[k]]
                engs[k].set_state(h["x"], h["y"], h["theta"], h["vel"], rad if t == 0 else None, nonblocking=True)
                engs[k].step(1)
                engs[k].get_state(bufs[k][cur[k] ^ 1], nonblocking=True)
                cur[k] ^= 1
    torch.cuda.synchronize()
    for k in (0, 1):
        for key in ("x", "y", "theta", "vel"):
            assert np.array_equal(bufs[k][cur[k]][key], ref[k][key]), key
        engs[k].close()
    with pytest.raises(ValueError):                              # no silent copies in the non-blocking mode
        e = VFEngine(1, 8, resolution=1200, width=W, height=W)
        z = np.zeros((1, 8), np.float64)
        e.set_state(z, z, z, z, 10.0, nonblocking=True)


@pytest.mark.parametrize("boundary", ["walls", "infinite"])
def test_warp_kernel_on_a_sparse_heterogeneous_swarm(built_lib, monkeypatch, boundary):
    """The warp-per-focal-agent kernel in its own territory -- one large sparse swarm with distance culling, record
    tiles skipped by bounding box on the Morton-sorted state, heterogeneous radii -- against the one-thread-per-focal-
    agent kernel (fields bit for bit; several steps incl. a re-sort) and against the oracle on sampled agents."""
    from abm_b200 import VFEngine
    rng = np.random.default_rng(77)
    N, R, W = 6000, 1200, 9000.0
    x, y, th, v = _random_scene(rng, 1, N, W)
    rad = rng.choice([5.0, 10.0, 14.0], (1, N)).astype(np.float32)
    res = {}
    for kern in ("onesided", "warp"):
        monkeypatch.setenv("ABM_VF_KERNEL", kern)
        eng = VFEngine(1, N, resolution=R, width=W, height=W, boundary=boundary, keep_fields=True, keep_terms=True,
                       resort_every=2)
        eng.set_params(); eng.set_state(x, y, th, v, rad)
        eng.step(1)
        assert eng.last_kernel() == ("abm::vf_step_warp_kernel" if kern == "warp" else "abm::vf_step_kernel")
        f1 = eng.fields_packed().copy()
        eng.step(4)
        res[kern] = (f1, eng.fields_packed(), eng.terms(), eng.get_state())
        eng.close()
    a, b = res["onesided"], res["warp"]
    assert np.array_equal(a[0], b[0])                      # fields of the first step: bit for bit
    # later steps: the warp kernel reduces its edge sums across lanes, so the fp64 terms agree to rounding and the fp32
    # state to an ulp; fields are compared where both saw identical state
    np.testing.assert_allclose(b[2], a[2], rtol=1e-9, atol=1e-9)
    for k in ("x", "y", "theta", "vel"):
        np.testing.assert_allclose(b[3][k], a[3][k], rtol=1e-5, atol=1e-3, err_msg=k)
    assert (a[1] != b[1]).any(axis=-1).mean() < 0.01
    cfg = rs.VFConfig(R=R, width=W, height=W, boundary=boundary)
    sample = [0, 1, 777, 3000, N - 1]
    ref = rs.vf_step_frozen(x[0], y[0], th[0], v[0], rad[0], cfg, agents=sample)
    assert np.array_equal(rs.unpack_bits(b[0][0], R)[sample], ref["rows"][sample][:, ::-1])


def test_long_run_kernels_stay_identical(built_lib, monkeypatch):
    """Size-independent property at the benchmark's agent count: 300 steps of 8 replicates x 1024 agents from the
    benchmark's initial condition end in bit-identical state and fields under the symmetric and the one-thread-per-
    focal-agent kernel (both exact, same epilogue) -- any unflagged fp32 bin error in either kernel would show up here
    and then diverge -- and the on-device metrics of the final state agree with numpy."""
    import bench
    from abm_b200 import VFEngine
    B, N = 8, 1024
    W = bench.arena_side(N)
    x, y, th, v = bench.synthetic_state(B, N)
    res = {}
    for kern in ("onesided", "symmetric"):
        monkeypatch.setenv("ABM_VF_KERNEL", kern)
        eng = VFEngine(B, N, resolution=1200, width=W, height=W, keep_fields=True)
        eng.set_params(**bench.PARAMS); eng.set_state(x, y, th, v, 10.0)
        eng.step(300)
        res[kern] = (eng.get_state(), eng.fields_packed().copy(), eng.counters(), eng.metrics())
        eng.close()
    a, b = res["onesided"], res["symmetric"]
    for k in ("x", "y", "theta", "vel"):
        assert np.array_equal(a[0][k], b[0][k]), k
    assert np.array_equal(a[1], b[1])
    assert np.isfinite(a[0]["x"]).all() and (a[0]["vel"] != 0).any()
    pol = np.hypot(np.cos(b[0]["theta"].astype(np.float64)).sum(1), np.sin(b[0]["theta"].astype(np.float64)).sum(1)) / N
    np.testing.assert_allclose(b[3]["polarization"], pol, rtol=1e-5, atol=1e-6)
    print("fp64 re-evaluations per 1e6 directions, onesided / symmetric:",
          *(round(1e6 * r[2]["fp64_pairs"] / (300.0 * B * N * (N - 1)), 1) for r in (a, b)))


def test_set_state_keeps_radii_when_omitted(built_lib):
    """abm_set_state with radius == NULL keeps the radii of the previous call (heterogeneous ones too)."""
    from abm_b200 import VFEngine, _lib
    rng = np.random.default_rng(8)
    B, N, W = 2, 60, 500.0
    x, y, th, v = _random_scene(rng, B, N, W)
    rad = rng.choice([6.0, 10.0, 13.0], (B, N)).astype(np.float32)
    a = VFEngine(B, N, resolution=1200, width=W, height=W, keep_fields=True)
    b = VFEngine(B, N, resolution=1200, width=W, height=W, keep_fields=True)
    with pytest.raises(_lib.AbmError):
        a.set_state(x, y, th, v)                       # nothing to keep yet
    for e in (a, b):
        e.set_params(); e.set_state(x, y, th, v, rad); e.step(2)
    st = a.get_state()
    a.set_state(st["x"], st["y"], st["theta"], st["vel"])          # radii omitted
    b.set_state(st["x"], st["y"], st["theta"], st["vel"], rad)
    a.step(1); b.step(1)
    assert np.array_equal(a.fields_packed(), b.fields_packed())
    a.close(); b.close()


@pytest.mark.parametrize("B,N,chunked", [(450, 1024, True), (40, 128, False), (1, 3000, False)],
                         ids=["chunks_of_whole_waves", "one_chunk", "large_sparse_swarm"])
def test_step_host_matches_the_three_calls(built_lib, B, N, chunked):
    """abm_vf_step_host (upload, steps, download in one non-blocking call; the replicates of a large batch go through in
    chunks whose copies overlap the other chunks' steps) against set_state_packed -> step -> get_state_packed: the same
    bits out, the same fields and terms kept, repeated calls feeding the output back in."""
    import torch
    from abm_b200 import VFEngine
    rng = np.random.default_rng(41)
    W = {3000: 9000.0, 1024: 2880.0}.get(N, 700.0)
    x, y, th, v = _random_scene(rng, B, N, W)
    packed = np.ascontiguousarray(np.stack([x, y, th, v], axis=-1))
    ea = VFEngine(B, N, resolution=1200, width=W, height=W, keep_fields=True, keep_terms=True)
    eb = VFEngine(B, N, resolution=1200, width=W, height=W, keep_fields=True, keep_terms=True)
    for e in (ea, eb):
        e.set_params()
        e.set_state_packed(packed, 10.0)           # the radii (and what the kernel choice needs of them)
        e.step(1)
    bufs = [torch.from_numpy(packed.copy()).pin_memory().numpy() for _ in range(2)]
    ref = packed
    for it in range(3):
        n_steps = 1 + (it % 2)
        ea.set_state_packed(ref)
        ea.step(n_steps)
        ref = ea.get_state_packed()
        launches0 = eb.counters()["launches"]
        eb.step_host(bufs[it % 2], bufs[(it + 1) % 2], n_steps)
        eb.synchronize()
        assert np.array_equal(bufs[(it + 1) % 2], ref), it
        assert np.array_equal(eb.fields(), ea.fields()) and np.array_equal(eb.terms(), ea.terms())
        n_launch = eb.counters()["launches"] - launches0
        if chunked:
            assert eb.last_kernel() == "abm::vf_step_sym_kernel" and n_launch > n_steps     # one launch per chunk and step
        st = eb.get_state()
        assert np.array_equal(st["x"], ref[..., 0]) and np.array_equal(st["vel"], ref[..., 3])
    # (the state the engine keeps is the downloaded one: a plain step continues from it)
    ea.step(1); eb.step(1)
    assert np.array_equal(eb.get_state_packed(), ea.get_state_packed())
    ea.close(); eb.close()


def test_step_host_adapts_to_crowding_like_step(built_lib):
    """The crowded-scene detection (pairs off the fast path, counted by the kernel, read back without blocking) also works
    when the steps come in through abm_vf_step_host's replicate chunks: after a few calls the three-word variant runs,
    and the results stay those of the plain calls."""
    import torch
    from abm_b200 import VFEngine
    rng = np.random.default_rng(43)
    B, N, W = 450, 1024, 2880.0
    x, y, th, v = _random_scene(rng, B, N, W, spread=(1000.0, 1900.0))       # 1024 agents in 900 x 900 px: ~8 % of the pairs
                                                                             # leave the two-word fast path (threshold 4 %)
    packed = np.ascontiguousarray(np.stack([x, y, th, v], axis=-1))
    ea = VFEngine(B, N, resolution=1200, width=W, height=W)
    eb = VFEngine(B, N, resolution=1200, width=W, height=W)
    for e in (ea, eb):
        e.set_params(); e.set_state_packed(packed, 10.0); e.step(1)
    bufs = [torch.from_numpy(packed.copy()).pin_memory().numpy() for _ in range(2)]
    ref = packed
    for it in range(8):
        ea.set_state_packed(ref); ea.step(1); ref = ea.get_state_packed()
        eb.step_host(bufs[it % 2], bufs[(it + 1) % 2], 1); eb.synchronize()
        assert np.array_equal(bufs[(it + 1) % 2], ref), it
    assert ea.kernel_stats()["symmetric_wide"] > 0 and eb.kernel_stats()["symmetric_wide"] > 0
    ea.close(); eb.close()


@pytest.mark.parametrize("sorted_swarm", [False, True])
def test_packed_state_matches_soa_state(built_lib, monkeypatch, sorted_swarm):
    """abm_set_state_packed / abm_get_state_packed (ONE interleaved (x, y, theta, vel) array per direction) against the
    four-array calls: the same bits in, the same bits out -- host arrays, device tensors, pinned non-blocking, and on an
    engine that keeps its agents in an internal (Morton) order (one large sparse swarm)."""
    import torch
    from abm_b200 import VFEngine
    rng = np.random.default_rng(31)
    B, N, W = (1, 5000, 9000.0) if sorted_swarm else (5, 200, 1200.0)
    x, y, th, v = _random_scene(rng, B, N, W)
    packed = np.ascontiguousarray(np.stack([x, y, th, v], axis=-1))
    ea = VFEngine(B, N, resolution=1200, width=W, height=W, resort_every=2)
    eb = VFEngine(B, N, resolution=1200, width=W, height=W, resort_every=2)
    for e in (ea, eb):
        e.set_params()
    ea.set_state(x, y, th, v, 10.0)
    eb.set_state_packed(packed, 10.0)
    got = eb.get_state_packed()
    assert np.array_equal(got, packed)                                   # round trip before any step
    ea.step(3); eb.step(3)
    if sorted_swarm:
        assert eb.last_kernel() == "abm::vf_step_warp_kernel" and not np.array_equal(eb.permutation()[0], np.arange(N))
    sa, pb = ea.get_state(), eb.get_state_packed()
    for i, k in enumerate(("x", "y", "theta", "vel")):
        assert np.array_equal(pb[..., i], sa[k]), k
    # the radii are kept when omitted; device tensors; pinned non-blocking
    dev = torch.from_numpy(pb).cuda()
    eb.set_state_packed(dev)
    out_dev = torch.empty_like(dev)
    eb.get_state_packed(out_dev)
    torch.cuda.synchronize()
    assert np.array_equal(out_dev.cpu().numpy(), pb)
    pin_in = torch.from_numpy(pb.copy()).pin_memory().numpy()
    pin_out = torch.empty(B, N, 4).pin_memory().numpy()
    eb.set_state_packed(pin_in, nonblocking=True)
    eb.step(2)
    eb.get_state_packed(pin_out, nonblocking=True)
    torch.cuda.synchronize()
    ea.set_state(sa["x"], sa["y"], sa["theta"], sa["vel"])
    ea.step(2)
    sa2 = ea.get_state()
    for i, k in enumerate(("x", "y", "theta", "vel")):
        assert np.array_equal(pin_out[..., i], sa2[k]), k
    ea.close(); eb.close()


def test_host_driven_loop_never_reads_back(built_lib):
    """A call of abm_set_state / abm_set_state_packed WITHOUT radii keeps what the engine knows about them: the next
    step must not fetch the batch's min / max radius again (a device -> host copy and a stream synchronisation on every
    step of a host-driven loop).  Observable: with non-blocking transfers a step enqueued behind a
    long-running kernel returns immediately."""
    import time
    import torch
    from abm_b200 import VFEngine
    B, N, W = 64, 1024, 2880.0
    rng = np.random.default_rng(5)
    x, y, th, v = _random_scene(rng, B, N, W)
    pin = torch.from_numpy(np.ascontiguousarray(np.stack([x, y, th, v], axis=-1))).pin_memory().numpy()
    out = torch.empty(B, N, 4).pin_memory().numpy()
    eng = VFEngine(B, N, resolution=1200, width=W, height=W)
    eng.set_params()
    eng.set_state_packed(pin, 10.0)
    eng.step(1)                                   # learns the radii (one read-back, once)
    torch.cuda.synchronize()
    blocker = torch.empty(1 << 28, device="cuda")
    t0 = time.perf_counter()
    for _ in range(40):
        blocker.normal_()                         # ~1 ms of GPU work each, queued ahead of the loop below
    for _ in range(5):
        eng.set_state_packed(pin, nonblocking=True)
        eng.step(1)
        eng.get_state_packed(out, nonblocking=True)
    t_enqueue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_total = time.perf_counter() - t0
    assert t_enqueue < 0.5 * t_total, (t_enqueue, t_total)      # the host ran ahead of the device
    eng.close()


def test_engines_on_two_devices_in_one_process(built_lib):
    """The opt-in to > 48 KB of dynamic shared memory is per DEVICE: a process that creates engines on two devices must
    be able to launch the large-shared-memory kernels on both."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from abm_b200 import VFEngine
    rng = np.random.default_rng(9)
    B, N, W = 160, 1024, 2880.0
    x, y, th, v = _random_scene(rng, B, N, W)
    res = []
    for dev in (0, 1):
        with torch.cuda.device(dev):
            eng = VFEngine(B, N, resolution=1200, width=W, height=W, device=dev)
            eng.set_params(); eng.set_state(x, y, th, v, 10.0); eng.step(2)
            assert eng.last_kernel() == "abm::vf_step_sym_kernel"
            res.append(eng.get_state()); eng.close()
    for k in ("x", "y", "theta", "vel"):
        assert np.array_equal(res[0][k], res[1][k])


@pytest.mark.parametrize("B,N,boundary,fovr,hetero", [(1, 100, "walls", 1.0, False), (3, 40, "infinite", 1.0, True),
                                                      (2, 64, "walls", 0.5, False), (1, 64, "infinite", 1.0, True),
                                                      (1, 23, "walls", 0.5, False), (1, 5, "walls", 1.0, False)])
def test_small_runs_step_inside_one_launch(built_lib, monkeypatch, B, N, boundary, fovr, hetero):
    """A run so small that a step is shorter than a kernel launch (BASELINE configs[1]: one run of 100 agents) takes all
    n_steps in ONE launch with a barrier over all CTAs between steps -- one replicate of at most 64 agents as ONE
    thread-block cluster (hardware cluster barrier), otherwise a cooperative launch with a grid-wide barrier in global
    memory: the same trajectory, bit for bit, as a launch per step, and the last step's fields / terms are the ones kept."""
    from abm_b200 import VFEngine
    rng = np.random.default_rng(40 + N)
    W = 900.0 if N == 100 else 400.0
    x, y, th, v = _random_scene(rng, B, N, W)
    rad = rng.choice([6.0, 10.0], (B, N)).astype(np.float32) if hetero else 10.0
    fov = (-fovr * np.pi, fovr * np.pi)
    R = int(1200 / fovr)
    res = {}
    for per_step in (False, True):
        if per_step:
            monkeypatch.setenv("ABM_VF_ONE_STEP_PER_LAUNCH", "1")
        else:
            monkeypatch.delenv("ABM_VF_ONE_STEP_PER_LAUNCH", raising=False)
        eng = VFEngine(B, N, resolution=R, fov=fov, boundary=boundary, width=W, height=W, keep_fields=True, keep_terms=True)
        eng.set_params(); eng.set_state(x, y, th, v, rad)
        eng.step(37)
        assert eng.last_kernel() == "abm::vf_step_warp_kernel"
        assert eng.counters()["launches"] == (37 if per_step else 1)
        assert eng.cluster_launches() == (1 if (B == 1 and N <= 64 and not per_step) else 0)
        eng.step(2)                                   # an even and an odd count: both table parities
        res[per_step] = (eng.get_state(), eng.fields_packed(), eng.terms())
        eng.close()
    for k in ("x", "y", "theta", "vel"):
        assert np.array_equal(res[False][0][k], res[True][0][k]), k
    assert np.array_equal(res[False][1], res[True][1]) and np.array_equal(res[False][2], res[True][2])
    # and the last step against the oracle, from the state before it
    eng = VFEngine(B, N, resolution=R, fov=fov, boundary=boundary, width=W, height=W)
    eng.set_params(); eng.set_state(x, y, th, v, rad); eng.step(38)
    st = eng.get_state(); eng.close()
    cfg = rs.VFConfig(R=R, fov=fov, boundary=boundary, width=W, height=W)
    for b in range(B):
        ref = rs.vf_step_frozen(st["x"][b], st["y"][b], st["theta"][b], st["vel"][b], rad if not hetero else rad[b], cfg)
        assert np.array_equal(rs.unpack_bits(res[False][1][b], R), ref["rows"][:, ::-1])
        _check_state(res[False][0], ref, b)
