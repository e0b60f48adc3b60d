"""Generates tests/golden/cs_golden.npz by EXECUTING THE UNMODIFIED REFERENCE function
abm/projects/cooperative_signaling/cs_agent/cs_supcalc.py::projection_field (imported from
/root/reference through oracle/ref_shim.py).  Runs only in the build container; the fixture is
committed so the GPU box can check the oracle restatement and the CUDA entry point against real
reference output.

    python tests/golden/make_golden_cs.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def scenes():
    """(fov, R, position, radius, orientation, objects, meters, max_proj_size) -- the first entry is the
    reference's own golden vector (cs_agent/tests/test_cs_supcalc.py:143-158)."""
    yield ((-np.pi, np.pi), 8, np.array([-1.0, -1.0]), 1.0, 0.0, [np.array([0.0, -1.0])], None, None)
    # objects straight behind the focal agent: projections that wrap around the ends of the retina (:274-279)
    yield ((-np.pi, np.pi), 1200, np.array([100.0, 100.0]), 10.0, 0.0,
           [np.array([60.0, 100.0]), np.array([60.0, 99.0]), np.array([60.0, 101.5]), np.array([103.0, 101.0])],
           None, None)
    yield ((-np.pi, np.pi), 601, np.array([100.0, 100.0]), 5.0, np.pi / 2,
           [np.array([100.0, 130.0]), np.array([100.5, 112.0]), np.array([99.0, 160.0])], [0.25, 0.5, 1.0], 400.0)
    yield ((-0.75 * np.pi, 0.75 * np.pi), 1200, np.array([50.0, 50.0]), 10.0, 1.0,
           [np.array([50.0, 50.0]) + 30 * np.array([-np.cos(1.0), np.sin(1.0)]), np.array([52.0, 49.0])],
           [0.3, 0.6], None)
    rng = np.random.default_rng(20261018)
    for c in range(23):
        R = int(rng.choice([320, 601, 1200, 2400]))
        n = int(rng.integers(1, 14))
        rad = float(rng.choice([1.0, 5.0, 10.0, 15.5]))
        pos = np.round(rng.uniform(0, 300, 2), int(rng.integers(0, 3)))
        objs = [np.round(rng.uniform(0, 300, 2), int(rng.integers(0, 3))) for _ in range(n)]
        if c % 4 == 0:
            objs[0] = pos.copy()                        # coincident: skipped (:233)
        if c % 3 == 0:
            objs[-1] = pos + rng.uniform(-3, 3, 2)      # overlapping neighbour: very wide projection, wraps
        fr = float(rng.choice([1.0, 0.75, 0.5, 0.25]))
        th = float(rng.uniform(0, 2 * np.pi)) if c % 5 else float(rng.choice([0.0, np.pi / 2, np.pi]))
        meters = None if c % 2 else list(rng.uniform(0, 1, n))
        mps = None if c % 3 == 1 else float(rng.uniform(2, R / 3))
        yield ((-fr * np.pi, fr * np.pi), R, pos, rad, th, objs, meters, mps)


def main():
    if not ref_shim.reference_available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    ref_shim.install()
    from abm.projects.cooperative_signaling.cs_agent import cs_supcalc as cs
    out = {}
    n_cases = 0
    for c, (fov, R, pos, rad, th, objs, meters, mps) in enumerate(scenes()):
        rows = cs.projection_field(fov, R, pos, rad, th, objs, meters, mps)
        p = f"c{c}_"
        out[p + "scalars"] = np.array([fov[0], fov[1], R, pos[0], pos[1], rad, th, -1.0 if mps is None else mps],
                                      np.float64)
        out[p + "objs"] = np.array(objs, np.float64)
        out[p + "meters"] = np.array([] if meters is None else meters, np.float64)
        out[p + "rows"] = rows
        n_cases += 1
    out["n_cases"] = np.int64(n_cases)
    np.savez_compressed(os.path.join(OUT, "cs_golden.npz"), **out)
    print("wrote", os.path.join(OUT, "cs_golden.npz"), n_cases, "cases")


if __name__ == "__main__":
    main()
