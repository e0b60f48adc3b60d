"""Generates tests/golden/base_hetero_radius_golden.npz: agents with their OWN radius (agent_radius entry of
agent_behave_param_list, sims.py:502) on top of per-agent decision parameters; pins the restatement and the CUDA path
(abm_base_set_agent_radii).  The
UNMODIFIED reference's constructor takes the dictionaries, Agent.update runs from a frozen snapshot as in
make_golden_base.py.  Build container only:

    python tests/golden/make_golden_hetero_radius.py
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

KEYS = mgh.AGENT_KEYS + mgh.GEO_KEYS + ["agent_radius"]


def main():
    if not ref_shim.reference_available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    rng = np.random.default_rng(20261020)
    specs = [
        # N, R, W, fov, vision_range, visual_exclusion, patchwise_exclusion, Eps_w, intpos
        (24, 1200, 300.0, 1.0, 2000.0, True, True, 2.0, True),
        (20, 1200, 250.0, 0.75, 150.0, False, True, 2.0, False),
        (30, 601, 200.0, 1.0, 2000.0, True, False, 2.0, True),      # crowded: heavy occlusion, wall contacts
    ]
    out = {"n_cases": np.int64(len(specs)), "agent_keys": np.array(KEYS)}
    for c, spec in enumerate(specs):
        cfg, st, dth = mgb.scene(rng, *spec)
        tab, plist = mgh.behave_params(rng, spec[0], cfg, hetero_geometry=(c == 1))
        tab["agent_radius"] = rng.choice([5.0, 8.0, 10.0, 14.0], spec[0])
        for i, d in enumerate(plist):
            r = float(tab["agent_radius"][i])
            d["agent_radius"] = int(r)
        fields, res = mgh.run_reference(cfg, st, dth, plist)
        p = f"c{c}_"
        out[p + "cfg"] = np.array([float(getattr(cfg, k)) for k in mgb.CFG_KEYS])
        out[p + "fov"] = np.array(cfg.fov)
        for k in mgb.STATE_KEYS:
            out[p + "st_" + k] = np.asarray(st[k])
        out[p + "dth"] = dth
        out[p + "agent_params"] = np.stack([tab[k] for k in KEYS], axis=1)
        out[p + "fields"] = pack_bits(fields)
        for k in mgb.OUT_KEYS:
            out[p + "out_" + k] = res[k]
    np.savez_compressed(os.path.join(HERE, "base_hetero_radius_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "base_hetero_radius_golden.npz"))


if __name__ == "__main__":
    main()
