"""Generates tests/golden/vf_lines_golden.npz: visual-flocking agents that have lines to follow (vf_agent.py:273-276 ->
vf_supcalc.follow_lines_local, vf_supcalc.py:293-328).  The UNMODIFIED reference's VFAgent.update runs from a frozen
snapshot as in make_golden.py, with ``agent.lines`` non-empty and a synthetic ``agent.line_map`` (the reference rasterises
mouse-drawn lines into it; here: blurred random polylines, values in [0, 1], the same map for every agent).  Agents close
to the arena's edges exercise the wrap / clip / empty-window cases of the numpy slices.  Build container only:

    python tests/golden/make_golden_lines.py
"""
import copy
import os
import sys

import numpy as np
from scipy.ndimage import gaussian_filter

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.restate import pack_bits  # noqa: E402


def line_map(rng, d0, d1, n_lines):
    m = np.zeros((d0, d1))
    for _ in range(n_lines):
        p, q = rng.uniform(0, [d0, d1]), rng.uniform(0, [d0, d1])
        for t in np.linspace(0, 1, 4 * max(d0, d1)):
            c = p + t * (q - p)
            i, j = int(c[0]), int(c[1])
            m[max(i - 2, 0):i + 3, max(j - 2, 0):j + 3] = 1.0
    m = gaussian_filter(m, sigma=2.0)
    m[m < 0.05] = 0.0                       # large exactly-zero areas (the s1 == s2 == 0 branch)
    m = np.minimum(m / m.max(), 1.0)
    m[d0 // 3:d0 // 3 + 110, d1 // 3:d1 // 3 + 110] = 1.0    # a saturated block (s1 == s2 != 0: the 0.01 branch)
    return m


def run(sc, lmap):
    agents = ref_shim.make_vf_agents(sc["x"], sc["y"], sc["theta"], sc["vel"], sc["radius"], R=sc["R"],
                                     fov_ratio=sc["fov_ratio"], width=sc["W"], height=sc["W"],
                                     boundary=sc["boundary"], params=mg.VF_PARAMS, limit_movement=sc["limit"])
    N = len(agents)
    fields = np.zeros((N, sc["R"]), bool)
    new = np.zeros((N, 4))
    for ag in agents:
        ag.line_map = lmap                  # shared, read-only
        ag.lines = [[(0, 0), (1, 1)]]       # len(self.lines) != 0 (vf_agent.py:273)
    for i in range(N):
        cp = [copy.copy(a) for a in agents]
        a = cp[i]
        a.position = np.array(a.position, dtype=np.float64)
        a.verbose_supcalc = False           # the line-following branch is the non-verbose one (vf_agent.py:268-276)
        a.update(cp)
        fields[i] = a.soc_v_field > 0
        new[i] = [a.position[0], a.position[1], a.orientation, a.velocity]
    return fields, new


def main():
    if not ref_shim.reference_available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    rng = np.random.default_rng(20261022)
    specs = [
        # N, R, W, boundary, fov, radius, limit, overrides
        (40, 1200, 300, "walls", 1.0, 10, False, False),
        (32, 1200, 400, "infinite", 1.0, 10, False, False),
        (24, 601, 250, "walls", 1.0, 6, True, False),
    ]
    out = {"n_cases": np.int64(len(specs))}
    for c, spec in enumerate(specs):
        sc = mg.vf_scene(rng, *spec)
        sc["vel"][::7] = 0.0                                     # np.sign(agvel) == 0: no steering (:317-320)
        # (WIDTH + window_pad, HEIGHT + window_pad), vf_agent.py:32; fp32 values (what the engine holds) cast to float64
        lmap = line_map(rng, spec[2] + 30, spec[2] + 30, 6).astype(np.float32).astype(np.float64)
        fields, new = run(sc, lmap)
        p = f"c{c}_"
        for k in ("x", "y", "theta", "vel", "radius"):
            out[p + k] = sc[k]
        out[p + "cfg"] = np.array([spec[1], spec[2], 1.0 if spec[3] == "infinite" else 0.0, spec[4], float(spec[6])])
        out[p + "line_map"] = lmap.astype(np.float32)
        out[p + "fields"] = pack_bits(fields)
        out[p + "new"] = new
    np.savez_compressed(os.path.join(HERE, "vf_lines_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "vf_lines_golden.npz"))


if __name__ == "__main__":
    main()
