"""Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (scioip34/ABM at
/root/reference, imported through oracle/ref_shim.py).  Runs only in the build container;
the fixtures it writes are committed so that the GPU box (no reference tree) can check
both the oracle restatement and the CUDA path against real reference output.

    python tests/golden/make_golden.py

Frozen-snapshot definition: for every agent i all agents are deep-copied and
the reference's own ``update`` is called on copy i against the un-updated copies.
State fed to the reference = fp32 values cast to float64 (what the engine holds).
"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle.restate import pack_bits  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
VF_PARAMS = dict(GAM=0.1, V0=1.0, ALP0=1.0, ALP1=0.09, BET0=1.0, BET1=0.09)


def f32(a):
    return np.asarray(a, np.float32).astype(np.float64)


def vf_scene(rng, N, R, W, boundary, fov_ratio, radius, limit, overrides):
    pad = 30
    x = f32(rng.uniform(pad - 15, pad + W - 5, N))
    y = f32(rng.uniform(pad - 15, pad + W - 5, N))
    th = f32(rng.uniform(0, 2 * np.pi, N))
    vel = f32(rng.uniform(0, 2.5, N))
    rad = np.full(N, radius, np.float64)
    if radius == 0:   # heterogeneous radii
        rad = rng.choice([4.0, 6.0, 10.0, 15.0], N)
    alp0 = bet0 = v0 = None
    if overrides:
        alp0 = f32(rng.uniform(0, 3, N)); bet0 = f32(rng.uniform(0, 3, N)); v0 = f32(rng.uniform(0.5, 2, N))
    return dict(x=x, y=y, theta=th, vel=vel, radius=rad, R=R, W=W, boundary=boundary, fov_ratio=fov_ratio,
                limit=limit, alp0=alp0, bet0=bet0, v0=v0)


def run_vf_reference(sc):
    agents = ref_shim.make_vf_agents(sc["x"], sc["y"], sc["theta"], sc["vel"], sc["radius"], R=sc["R"],
                                     fov_ratio=sc["fov_ratio"], width=sc["W"], height=sc["W"],
                                     boundary=sc["boundary"], params=VF_PARAMS, alp0=sc["alp0"], bet0=sc["bet0"],
                                     v0=sc["v0"], limit_movement=sc["limit"])
    N = len(agents)
    fields = np.zeros((N, sc["R"]), bool)
    terms = np.zeros((N, 6))
    new = np.zeros((N, 4))
    for i in range(N):
        cp = copy.deepcopy(agents)
        a = cp[i]
        a.verbose_supcalc = True
        a.update(cp)
        fields[i] = a.soc_v_field > 0
        terms[i] = [a.dv, a.dphi, a.ablob, a.aedge, a.bblob, a.bedge]
        new[i] = [a.position[0], a.position[1], a.orientation, a.velocity]
    return fields, terms, new


def main():
    if not ref_shim.reference_available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    rng = np.random.default_rng(20261017)
    specs = [
        # N, R, W, boundary, fov, radius, limit, overrides
        (12, 1200, 300, "walls", 1.0, 10, False, False),
        (24, 1200, 900, "walls", 1.0, 10, False, False),
        (16, 1200, 900, "infinite", 1.0, 10, False, False),
        (10, 1200, 900, "infinite", 1.0, 5, False, True),
        (12, 2400, 400, "walls", 0.5, 10, False, False),     # R already rescaled: 1200 / 0.5
        (12, 1600, 400, "walls", 0.75, 10, True, False),
        (9, 601, 250, "walls", 1.0, 0, False, False),        # odd R, heterogeneous radii
        (8, 320, 200, "infinite", 1.0, 10, True, True),
        (2, 1200, 900, "walls", 1.0, 10, False, False),
        (20, 1200, 120, "walls", 1.0, 10, False, False),     # crowded: overlaps, wrap-around intervals
    ]
    out = {"n_cases": np.int64(len(specs))}
    for c, (N, R, W, bnd, fov, rad, lim, ovr) in enumerate(specs):
        sc = vf_scene(rng, N, R, W, bnd, fov, rad, lim, ovr)
        fields, terms, new = run_vf_reference(sc)
        p = f"c{c}_"
        for k in ("x", "y", "theta", "vel", "radius"):
            out[p + k] = sc[k]
        out[p + "meta"] = np.array([N, R, W, 1 if bnd == "infinite" else 0, int(lim), int(ovr)], np.int64)
        out[p + "fov_ratio"] = np.float64(fov)
        if ovr:
            out[p + "alp0"], out[p + "bet0"], out[p + "v0"] = sc["alp0"], sc["bet0"], sc["v0"]
        out[p + "fields"] = pack_bits(fields)
        out[p + "terms"] = terms
        out[p + "new"] = new
    # function-level: per-object rows of vf_supcalc.projection_field
    vs, _, _ = ref_shim.load_vf()
    pf_cases = [
        # the reference's own golden vector (test_cs_supcalc.py:143-158), identical for VF
        dict(fov=(-np.pi, np.pi), R=8, pos=(-1.0, -1.0), r=1, th=0.0, objs=[(0.0, -1.0)], sizes=None,
             bnd="walls", W=None, vr=None),
        # KAT-V1 / V2 / V3 of SURVEY App. B
        dict(fov=(-np.pi, np.pi), R=1200, pos=(300.0, 300.0), r=10, th=1.0,
             objs=[(350.0, 300.0), (300.0, 200.0), (180.0, 330.0), (310.0, 420.0)], sizes=[10] * 4, bnd="walls",
             W=None, vr=None),
        dict(fov=(-np.pi, np.pi), R=1200, pos=(20.0, 20.0), r=10, th=0.0, objs=[(880.0, 30.0)], sizes=[10],
             bnd="infinite", W=900, vr=None),
        dict(fov=(-np.pi / 2, np.pi / 2), R=2400, pos=(300.0, 300.0), r=10, th=1.0,
             objs=[(350.0, 300.0), (300.0, 200.0), (180.0, 330.0), (310.0, 420.0)], sizes=[10] * 4, bnd="walls",
             W=None, vr=None),
        # vision range + object_sizes None
        dict(fov=(-np.pi, np.pi), R=600, pos=(100.0, 100.0), r=8, th=2.5,
             objs=[(130.0, 100.0), (100.0, 400.0), (90.0, 95.0), (100.0, 100.0)], sizes=None, bnd="walls", W=None,
             vr=150.0),
    ]
    out["n_pf"] = np.int64(len(pf_cases))
    for c, pc in enumerate(pf_cases):
        rows = vs.projection_field(pc["fov"], pc["R"], np.array(pc["pos"]), pc["r"], pc["th"],
                                   [np.array(o) for o in pc["objs"]], object_sizes=pc["sizes"],
                                   boundary_cond=pc["bnd"], arena_width=pc["W"], arena_height=pc["W"],
                                   vision_range=pc["vr"])
        p = f"pf{c}_"
        out[p + "scalars"] = np.array([pc["fov"][0], pc["fov"][1], pc["R"], pc["pos"][0], pc["pos"][1], pc["r"],
                                       pc["th"], 1 if pc["bnd"] == "infinite" else 0, pc["W"] or 0,
                                       -1 if pc["vr"] is None else pc["vr"]], np.float64)
        out[p + "objs"] = np.array(pc["objs"], np.float64)
        out[p + "sizes"] = np.array(pc["sizes"] if pc["sizes"] is not None else [], np.float64)
        out[p + "rows"] = pack_bits(rows > 0)
    np.savez_compressed(os.path.join(OUT, "vf_golden.npz"), **out)
    print("wrote", os.path.join(OUT, "vf_golden.npz"))


if __name__ == "__main__":
    main()
