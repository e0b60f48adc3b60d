"""CPU tests pinning the oracle restatement of the cooperative-signaling projection field
(oracle/restate_cs.py) to
  * the reference's own golden vectors for this function (cs_agent/tests/test_cs_supcalc.py:143-168),
  * tests/golden/cs_golden.npz, produced by executing the unmodified reference, and
  * the live reference on random scenes when /root/reference is present (build container only).
"""
import numpy as np
import pytest

from golden_io import load_cs_cases
from oracle import ref_shim
from oracle import restate as rs
from oracle import restate_cs as rc


def test_reference_golden_projection_vector():
    """test_cs_supcalc.py:143-158."""
    out = rc.cs_projection_field((-np.pi, np.pi), 8, np.array([-1, -1]), 1, 0, [np.array([0, -1])], None)
    assert out.shape == (1, 8)
    assert out.tolist() == [[0., 0., 0., 0., 1., 1., 0., 0.]]


@pytest.mark.parametrize("v1,v2,expected", [([1, 0], [1, 0], 0), ([1, 0], [0, 1], -np.pi / 2),
                                            ([1, 0], [-1, 0], -np.pi)])
def test_reference_golden_closed_angles(v1, v2, expected):
    """test_cs_supcalc.py:161-168."""
    assert np.isclose(rs.closed_angle_vf(*map(float, v1), *map(float, v2)), expected)


@pytest.mark.parametrize("case", load_cs_cases(), ids=lambda c: f"R{c['R']}_n{len(c['objs'])}")
def test_restatement_matches_reference_fixture(case):
    c = case
    out = rc.cs_projection_field(c["fov"], c["R"], c["pos"], c["r"], c["th"], c["objs"], c["meters"], c["mps"])
    assert out.shape == c["rows"].shape
    assert np.array_equal(out, c["rows"])


def test_fixture_exercises_every_switch():
    cases = load_cs_cases()
    assert any(c["meters"] is not None for c in cases) and any(c["meters"] is None for c in cases)
    assert any(c["mps"] is not None for c in cases) and any(c["mps"] is None for c in cases)
    assert any(c["fov"][1] < np.pi for c in cases)
    # max_proj_size drops at least one projection that the FOV alone would keep, and keeps others
    dropped = kept = 0
    for c in cases:
        if c["mps"] is None:
            continue
        d_all = rc.cs_intervals(c["fov"], c["R"], c["pos"], c["r"], c["th"], c["objs"], None)["drawn"]
        d_max = rc.cs_intervals(c["fov"], c["R"], c["pos"], c["r"], c["th"], c["objs"], c["mps"])["drawn"]
        dropped += int((d_all & ~d_max).sum())
        kept += int(d_max.sum())
    assert dropped > 0 and kept > 0
    # wrapped projections (ends outside [0, R)) are present
    assert any((lambda d: bool((d["drawn"] & ((d["ps"] < 0) | (d["pe"] >= c["R"]))).any()))(
        rc.cs_intervals(c["fov"], c["R"], c["pos"], c["r"], c["th"], c["objs"], c["mps"])) for c in cases)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree only exists in the build container")
def test_restatement_matches_live_reference_random():
    ref_shim.install()
    from abm.projects.cooperative_signaling.cs_agent import cs_supcalc as cs
    rng = np.random.default_rng(77)
    for _ in range(300):
        R = int(rng.choice([8, 320, 601, 1200, 2400]))
        n = int(rng.integers(1, 12))
        rad = float(rng.choice([1, 5, 10, 15.5]))
        pos = np.round(rng.uniform(0, 300, 2), int(rng.integers(0, 3)))
        objs = [np.round(rng.uniform(0, 300, 2), int(rng.integers(0, 3))) for _ in range(n)]
        if rng.random() < 0.2:
            objs[0] = pos.copy()
        if rng.random() < 0.3:
            objs[-1] = pos + rng.uniform(-3, 3, 2)
        fr = float(rng.choice([1.0, 0.75, 0.5, 0.25]))
        fov = (-fr * np.pi, fr * np.pi)
        th = float(rng.uniform(0, 2 * np.pi)) if rng.random() < 0.8 else float(rng.choice([0, np.pi / 2, np.pi]))
        meters = None if rng.random() < 0.5 else list(rng.uniform(0, 1, n))
        mps = None if rng.random() < 0.5 else float(rng.uniform(1, max(R / 3, 2)))
        a = cs.projection_field(fov, R, pos, rad, th, objs, meters, mps)
        b = rc.cs_projection_field(fov, R, pos, rad, th, objs, meters, mps)
        assert a.shape == b.shape and np.array_equal(a, b)
