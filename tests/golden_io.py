"""Loader for tests/golden/vf_golden.npz (written by tests/golden/make_golden.py from real
reference output)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_vf_cases():
    z = np.load(os.path.join(GOLDEN, "vf_golden.npz"))
    cases = []
    for c in range(int(z["n_cases"])):
        p = f"c{c}_"
        N, R, W, inf, lim, ovr = (int(v) for v in z[p + "meta"])
        case = dict(N=N, R=R, W=float(W), boundary="infinite" if inf else "walls", limit=bool(lim),
                    fov_ratio=float(z[p + "fov_ratio"]),
                    x=z[p + "x"], y=z[p + "y"], theta=z[p + "theta"], vel=z[p + "vel"], radius=z[p + "radius"],
                    fields=z[p + "fields"], terms=z[p + "terms"], new=z[p + "new"],
                    alp0=z[p + "alp0"] if ovr else None, bet0=z[p + "bet0"] if ovr else None,
                    v0=z[p + "v0"] if ovr else None)
        cases.append(case)
    return cases


def load_pf_cases():
    z = np.load(os.path.join(GOLDEN, "vf_golden.npz"))
    cases = []
    for c in range(int(z["n_pf"])):
        p = f"pf{c}_"
        s = z[p + "scalars"]
        sizes = z[p + "sizes"]
        cases.append(dict(fov=(s[0], s[1]), R=int(s[2]), pos=(s[3], s[4]), r=s[5], th=s[6],
                          boundary="infinite" if s[7] else "walls", W=s[8] if s[8] else None,
                          vr=None if s[9] < 0 else s[9], objs=z[p + "objs"],
                          sizes=None if sizes.size == 0 else sizes, rows=z[p + "rows"]))
    return cases
