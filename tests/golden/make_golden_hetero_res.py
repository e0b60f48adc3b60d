"""Generates tests/golden/base_hetero_res_golden.npz: agents with their OWN visual-field resolution (v_field_res entry of
agent_behave_param_list, sims.py:507) on top of per-agent decision parameters, FOV and vision range.  The UNMODIFIED
reference's constructor takes the dictionaries (agent.py:58, 73: each agent's fields have its own length), Agent.update
runs from a frozen snapshot as in make_golden_base.py.  Fields are stored padded with zeros up to the scene's
resolution (the largest).  Build container only:

    python tests/golden/make_golden_hetero_res.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden_base as mgb  # noqa: E402
import make_golden_hetero as mgh  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.restate import pack_bits  # noqa: E402

KEYS = mgh.AGENT_KEYS + mgh.GEO_KEYS + ["v_field_res"]


def main():
    if not ref_shim.reference_available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    rng = np.random.default_rng(20261021)
    specs = [
        # N, R (largest), W, fov, vision_range, visual_exclusion, patchwise_exclusion, Eps_w, intpos
        (24, 1200, 300.0, 1.0, 2000.0, True, True, 2.0, True),
        (20, 1200, 250.0, 0.75, 150.0, False, True, 2.0, False),     # + per-agent FOV / vision range
        (30, 1201, 200.0, 1.0, 2000.0, True, False, 2.0, True),      # crowded, odd resolutions among them
    ]
    choices = [[1200, 800, 601, 320], [1200, 1000, 64], [1201, 777, 400, 33]]
    out = {"n_cases": np.int64(len(specs)), "agent_keys": np.array(KEYS)}
    for c, spec in enumerate(specs):
        cfg, st, dth = mgb.scene(rng, *spec)
        tab, plist = mgh.behave_params(rng, spec[0], cfg, hetero_geometry=(c == 1))
        res = rng.choice(choices[c], spec[0])
        res[0] = choices[c][0]                     # the scene's resolution occurs
        tab["v_field_res"] = res.astype(float)
        for i, d in enumerate(plist):
            d["v_field_res"] = int(res[i])
        fields, outv = mgh.run_reference(cfg, st, dth, plist)
        p = f"c{c}_"
        out[p + "cfg"] = np.array([float(getattr(cfg, k)) for k in mgb.CFG_KEYS])
        out[p + "fov"] = np.array(cfg.fov)
        for k in mgb.STATE_KEYS:
            out[p + "st_" + k] = np.asarray(st[k])
        out[p + "dth"] = dth
        out[p + "agent_params"] = np.stack([tab[k] for k in KEYS], axis=1)
        out[p + "fields"] = pack_bits(fields)
        for k in mgb.OUT_KEYS:
            out[p + "out_" + k] = outv[k]
    np.savez_compressed(os.path.join(HERE, "base_hetero_res_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "base_hetero_res_golden.npz"))


if __name__ == "__main__":
    main()
