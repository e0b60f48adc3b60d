"""GPU parity tests of the cooperative-signaling projection field through the C ABI
(abm_cs_projection_field via abm_b200.cs_supcalc.projection_field) against
  * the reference's own golden vector (cs_agent/tests/test_cs_supcalc.py:143-158),
  * tests/golden/cs_golden.npz, produced by executing the unmodified reference, and
  * the CPU oracle (oracle/restate_cs.py) on seeded random scenes.
Bar: rows identical (0 / 1 / meter values are exact; bins bit for bit).
"""
import numpy as np
import pytest

from golden_io import load_cs_cases
from oracle import restate_cs as rc

pytestmark = pytest.mark.gpu


def test_reference_golden_vector_through_plugin(built_lib):
    from abm_b200 import cs_supcalc
    out = cs_supcalc.projection_field((-np.pi, np.pi), 8, np.array([-1, -1]), 1, 0, [np.array([0, -1])], None)
    assert out.shape == (1, 8)
    assert np.all(out == [[0., 0., 0., 0., 1., 1., 0., 0.]])


@pytest.mark.parametrize("case", load_cs_cases(), ids=lambda c: f"R{c['R']}_n{len(c['objs'])}")
def test_projection_field_matches_reference_fixture(built_lib, case):
    from abm_b200 import cs_supcalc
    c = case
    out = cs_supcalc.projection_field(c["fov"], c["R"], c["pos"], c["r"], c["th"], [o for o in c["objs"]],
                                      c["meters"], c["mps"])
    assert out.shape == c["rows"].shape
    assert np.array_equal(out, c["rows"])


def test_projection_field_matches_oracle_random(built_lib):
    from abm_b200 import cs_supcalc
    rng = np.random.default_rng(4242)
    n_rows = n_drawn = 0
    for _ in range(150):
        R = int(rng.choice([8, 320, 601, 1200, 2400, 4096]))
        n = int(rng.integers(1, 200))
        rad = float(rng.choice([1, 5, 10, 15.5]))
        pos = rng.uniform(0, 900, 2)
        objs = [rng.uniform(0, 900, 2) for _ in range(n)]
        objs[0] = pos.copy()
        objs[-1] = pos + rng.uniform(-3, 3, 2)
        fr = float(rng.choice([1.0, 0.75, 0.5, 0.25]))
        fov = (-fr * np.pi, fr * np.pi)
        th = float(rng.uniform(0, 2 * np.pi))
        meters = None if rng.random() < 0.5 else list(rng.uniform(0, 1, n))
        mps = None if rng.random() < 0.5 else float(rng.uniform(1, max(R / 3, 2)))
        want = rc.cs_projection_field(fov, R, pos, rad, th, objs, meters, mps)
        got = cs_supcalc.projection_field(fov, R, pos, rad, th, objs, meters, mps)
        assert got.shape == want.shape
        assert np.array_equal(got, want)
        n_rows += n
        n_drawn += int((want != 0).any(axis=1).sum())
    assert n_drawn > n_rows // 10


def test_empty_object_list_and_bad_resolution(built_lib):
    from abm_b200 import AbmError, cs_supcalc
    assert cs_supcalc.projection_field((-np.pi, np.pi), 1200, np.zeros(2), 10, 0.0, []).shape == (0, 1200)
    with pytest.raises(AbmError):
        cs_supcalc.projection_field((-np.pi, np.pi), 5000, np.zeros(2), 10, 0.0, [np.ones(2)])
