"""CPU tests: the oracle restatement (oracle/restate.py) against
  * the reference's own golden vectors (test_cs_supcalc.py:143-168),
  * the survey's known-answer vectors,
  * fixtures produced by executing the unmodified reference (tests/golden/vf_golden.npz),
  * and, when /root/reference is present, the live reference on fresh random scenes.
"""
import copy

import numpy as np
import pytest

from golden_io import load_pf_cases, load_vf_cases
from oracle import ref_shim
from oracle import restate as rs


def test_reference_golden_projection_vector():
    # fov=(-pi,pi), R=8, position=(-1,-1), radius=1, orientation=0, object=(0,-1) -> [[0,0,0,0,1,1,0,0]]
    cfg = rs.VFConfig(R=8)
    px, py = np.array([-1.0, 0.0]), np.array([-1.0, -1.0])
    rows = rs.vf_rows_per_object(px, py, 1.0, np.array([0.0, 0.0]), 0, cfg)
    assert rows.astype(int).tolist() == [[0, 0, 0, 0, 1, 1, 0, 0]]


@pytest.mark.parametrize("v1,v2,expect", [((1, 0), (1, 0), 0.0), ((1, 0), (0, 1), -np.pi / 2),
                                          ((1, 0), (-1, 0), -np.pi)])
def test_reference_golden_closed_angle(v1, v2, expect):
    got = float(rs.closed_angle_vf(np.float64(v1[0]), np.float64(v1[1]), np.float64(v2[0]), np.float64(v2[1])))
    assert np.isclose(got, expect)


def test_kat_v1_runs_and_terms():
    px = np.array([300.0, 350, 300, 180, 310]); py = np.array([300.0, 300, 200, 330, 420])
    th = np.array([1.0, 0, 0, 0, 0])
    cfg = rs.VFConfig(R=1200)
    stored = rs.vf_stored_field(px, py, 10.0, th, 0, cfg)
    assert rs.runs_of(stored) == [(130, 160), (473, 511), (754, 828), (1060, 1090)]
    t = rs.vswrm_terms(0.5, stored[::-1].astype(float), cfg)
    expect = (-0.11295710879884255, 0.16720418209113855, -0.13687988331415218, -0.02607722548469037,
              0.20545193082925006, -0.038247748738111514)
    np.testing.assert_allclose(t, expect, rtol=1e-12)


def test_kat_v2_torus():
    cfg = rs.VFConfig(R=1200, boundary="infinite", width=900, height=900)
    stored = rs.vf_stored_field(np.array([20.0, 880.0]), np.array([20.0, 30.0]), 10.0, np.zeros(2), 0, cfg)
    assert rs.runs_of(stored) == [(1108, 1198)]


def test_kat_v3_limited_fov():
    px = np.array([300.0, 350, 300, 180, 310]); py = np.array([300.0, 300, 200, 330, 420])
    cfg = rs.VFConfig(R=2400, fov=(-np.pi / 2, np.pi / 2))
    d = rs.vf_intervals(px, py, np.full(5, 10.0), np.array([1.0, 0, 0, 0, 0]), 0, cfg)
    assert d["fov_px"] == (600, 1799)
    rows = rs.vf_rows_per_object(px, py, 10.0, np.array([1.0, 0, 0, 0, 0]), 0, cfg)
    assert rows.sum(axis=1).tolist() == [150, 76, 0, 0]


def test_kat_dphi():
    for v, e in [([1, 1, 1, 0, 0, 0, 0, 0], [1, 0, 0, -1, 0, 0, 0, 0]),
                 ([0, 0, 1, 1, 1, 0, 0, 0], [0, 1, 0, 0, -1, 0, 0, 0]),
                 ([0, 0, 0, 0, 0, 0, 1, 1], [0, 0, 0, 0, 0, 1, 0, -1])]:
        assert rs.dphi_v(np.array(v, float)).tolist() == e


@pytest.mark.parametrize("case", load_vf_cases(), ids=lambda c: f"N{c['N']}_R{c['R']}_{c['boundary']}")
def test_restatement_matches_reference_fixture(case):
    c = case
    cfg = rs.VFConfig(R=c["R"], fov=(-c["fov_ratio"] * np.pi, c["fov_ratio"] * np.pi), boundary=c["boundary"],
                      width=c["W"], height=c["W"], limit_movement=c["limit"])
    out = rs.vf_step_frozen(c["x"], c["y"], c["theta"], c["vel"], c["radius"], cfg, c["alp0"], c["bet0"], c["v0"])
    assert np.array_equal(rs.pack_bits(out["rows"][:, ::-1]), c["fields"])
    np.testing.assert_allclose(out["terms"], c["terms"], rtol=1e-11, atol=1e-13)
    got = np.stack([out["x"], out["y"], out["theta"], out["vel"]], axis=1)
    np.testing.assert_allclose(got, c["new"], rtol=1e-13, atol=1e-11)


@pytest.mark.parametrize("pc", load_pf_cases(), ids=lambda p: f"R{p['R']}_{p['boundary']}")
def test_restatement_rows_match_reference_fixture(pc):
    n = len(pc["objs"])
    px = np.concatenate([[pc["pos"][0]], pc["objs"][:, 0]])
    py = np.concatenate([[pc["pos"][1]], pc["objs"][:, 1]])
    sizes = np.full(n, pc["r"]) if pc["sizes"] is None else pc["sizes"]
    r = np.concatenate([[pc["r"]], sizes])
    cfg = rs.VFConfig(R=pc["R"], fov=pc["fov"], boundary=pc["boundary"], width=pc["W"] or 0, height=pc["W"] or 0)
    rows = rs.vf_rows_per_object(px, py, r, np.concatenate([[pc["th"]], np.zeros(n)]), 0, cfg)
    if pc["vr"] is not None:   # vision_range (vf_supcalc.py:91-93) is not used by VFAgent; apply it here
        d = rs.vf_intervals(px, py, r, np.concatenate([[pc["th"]], np.zeros(n)]), 0, cfg)
        rows[d["dist"][1:] > pc["vr"]] = False
    assert np.array_equal(rs.pack_bits(rows), pc["rows"])


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_restatement_matches_live_reference_random():
    rng = np.random.default_rng(99)
    checked = 0
    for _ in range(6):
        N = int(rng.integers(2, 14))
        R = int(rng.choice([320, 1200, 1201]))
        W = float(rng.choice([200, 900]))
        bnd = str(rng.choice(["walls", "infinite"]))
        x = rng.uniform(10, 50 + W, N).astype(np.float32).astype(np.float64)
        y = rng.uniform(10, 50 + W, N).astype(np.float32).astype(np.float64)
        th = rng.uniform(0, 2 * np.pi, N).astype(np.float32).astype(np.float64)
        v = rng.uniform(0, 2, N).astype(np.float32).astype(np.float64)
        cfg = rs.VFConfig(R=R, boundary=bnd, width=W, height=W)
        agents = ref_shim.make_vf_agents(x, y, th, v, 10, R=R, width=W, height=W, boundary=bnd)
        out = rs.vf_step_frozen(x, y, th, v, 10.0, cfg)
        for i in range(N):
            cp = copy.deepcopy(agents)
            a = cp[i]
            a.verbose_supcalc = True
            a.update(cp)
            assert np.array_equal(a.soc_v_field > 0, out["rows"][i][::-1])
            np.testing.assert_allclose([a.dv, a.dphi], out["terms"][i][:2], rtol=1e-11, atol=1e-13)
            np.testing.assert_allclose([a.position[0], a.position[1], a.orientation, a.velocity],
                                       [out["x"][i], out["y"][i], out["theta"][i], out["vel"][i]], rtol=1e-13)
            checked += 1
    assert checked > 20


@pytest.mark.parametrize("case", load_vf_cases()[:6], ids=lambda c: f"N{c['N']}_R{c['R']}_{c['boundary']}")
def test_literal_port_matches_reference_fixture(case):
    """oracle/literal.py (the per-pair CPU-baseline port) against real reference output."""
    from oracle import literal
    c = case
    cfg = rs.VFConfig(R=c["R"], fov=(-c["fov_ratio"] * np.pi, c["fov_ratio"] * np.pi), boundary=c["boundary"],
                      width=c["W"], height=c["W"], limit_movement=c["limit"])
    for i in range(c["N"]):
        field, nx, ny, th, v = literal.agent_update(i, c["x"], c["y"], c["theta"], c["vel"], c["radius"], cfg,
                                                    c["alp0"], c["bet0"], c["v0"])
        assert np.array_equal(rs.pack_bits(field > 0), c["fields"][i])
        np.testing.assert_allclose([nx, ny, th, v], c["new"][i], rtol=1e-14, atol=1e-12)


def test_restatement_matches_reference_fixture_line_following():
    """follow_lines_local (vf_supcalc.py:293-328) and the turn-back wall rule of an agent with lines (vf_agent.py:80-129,
    273-276) against fixtures of the unmodified reference: every branch occurs (steering, standing agents, equal
    non-zero sensors -> 0.01, both sensors dark -> 0, empty windows at the arena's edge -> 0)."""
    from golden_io import load_vf_lines_cases
    n_const = n_zero = n_steer = 0
    for c in load_vf_lines_cases():
        cfg = rs.VFConfig(R=c["R"], fov=(-c["fov_ratio"] * np.pi, c["fov_ratio"] * np.pi), boundary=c["boundary"],
                          width=c["W"], height=c["W"], limit_movement=c["limit"])
        lm = c["line_map"].astype(np.float64)
        out = rs.vf_step_frozen(c["x"], c["y"], c["theta"], c["vel"], c["radius"], cfg, line_map=lm)
        assert np.array_equal(rs.pack_bits(out["rows"][:, ::-1]), c["fields"])
        got = np.stack([out["x"], out["y"], out["theta"], out["vel"]], axis=1)
        np.testing.assert_allclose(got, c["new"], rtol=1e-12, atol=1e-12)
        plain = rs.vf_step_frozen(c["x"], c["y"], c["theta"], c["vel"], c["radius"], cfg)
        assert (np.abs(plain["theta"] - c["new"][:, 2]) > 1e-9).sum() > c["N"] // 2     # not the flocking heading
        for i in range(c["N"]):
            d = rs.follow_lines_local((c["x"][i], c["y"][i]), c["radius"][i], c["theta"][i], lm, c["vel"][i], 9, 20)
            n_const += d == 0.01; n_zero += d == 0.0; n_steer += d not in (0.0, 0.01)
    assert n_const >= 3 and n_zero >= 20 and n_steer >= 20
