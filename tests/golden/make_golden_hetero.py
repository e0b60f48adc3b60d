"""Generates tests/golden/base_hetero_golden.npz: heterogeneous agents (agent_behave_param_list, sims.py:499-517).
The UNMODIFIED reference's Agent constructor takes the per-agent dictionaries (agent.py:83-108) and its
Agent.update runs from a frozen snapshot exactly as in make_golden_base.py.  Build container only:

    python tests/golden/make_golden_hetero.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden_base as mgb  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.restate import pack_bits  # noqa: E402

AGENT_KEYS = ["S_wu", "T_w", "Eps_w", "g_w", "B_w", "w_max", "S_uw", "T_u", "Eps_u", "g_u", "B_u", "u_max", "F_N", "F_R",
              "exp_vel_max", "exp_stop_ratio"]
GEO_KEYS = ["agent_fov", "vision_range"]     # constructor arguments FOV / vision_range of every agent (sims.py:506, 511)


def behave_params(rng, N, cfg, hetero_geometry=False):
    f32 = mgb.f32
    tab = dict(S_wu=f32(rng.uniform(0, 0.5, N)), T_w=f32(rng.uniform(0.2, 0.8, N)), Eps_w=f32(rng.uniform(0, 5, N)),
               g_w=f32(rng.uniform(0.05, 0.12, N)), B_w=f32(rng.uniform(0, 0.1, N)), w_max=f32(rng.uniform(0.7, 1.2, N)),
               S_uw=f32(rng.uniform(0, 0.1, N)), T_u=f32(rng.uniform(0.2, 0.8, N)), Eps_u=f32(rng.uniform(0.5, 3, N)),
               g_u=f32(rng.uniform(0.05, 0.12, N)), B_u=f32(rng.uniform(0, 0.1, N)), u_max=f32(rng.uniform(0.7, 1.2, N)),
               F_N=f32(rng.uniform(0.2, 2, N)), F_R=f32(rng.uniform(0.2, 1, N)),
               exp_vel_max=f32(rng.uniform(1, 4, N)), exp_stop_ratio=f32(rng.uniform(0.05, 0.3, N)))
    tab["agent_fov"] = np.full(N, cfg.fov[1] / np.pi)
    tab["vision_range"] = np.full(N, float(cfg.vision_range))
    if hetero_geometry:
        tab["agent_fov"] = rng.choice([1.0, 0.9, 0.75, 0.5, 0.25], N)
        tab["vision_range"] = rng.choice([60.0, 120.0, 200.0, 2000.0], N)
    plist = []
    for i in range(N):
        d = {k: float(tab[k][i]) for k in AGENT_KEYS + GEO_KEYS}
        d.update(Tau=cfg.Tau, agent_radius=10, v_field_res=cfg.R, pooling_time=0, pooling_prob=0, agent_consumption=1)
        plist.append(d)
    return tab, plist


def run_reference(cfg, st, dth, plist):
    orig = ref_shim.make_base_agents
    ref_shim.make_base_agents = lambda st_, cfg_: orig(st_, cfg_, behave_params_list=plist)
    try:
        return mgb.run_reference(cfg, st, dth)
    finally:
        ref_shim.make_base_agents = orig


def main():
    if not ref_shim.reference_available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    rng = np.random.default_rng(20261019)
    specs = [
        # N, R, W, fov, vision_range, visual_exclusion, patchwise_exclusion, Eps_w (module default, unused), intpos
        (24, 1200, 300.0, 1.0, 2000.0, True, True, 2.0, True),
        (16, 1200, 500.0, 0.5, 2000.0, False, True, 2.0, True),
        (30, 601, 250.0, 0.75, 200.0, True, False, 2.0, False),
        # per-agent FOV and vision range as well (the scene's fov / vision_range are then only the engine-wide defaults)
        (28, 1200, 300.0, 1.0, 2000.0, True, True, 2.0, True),
        (20, 1200, 250.0, 1.0, 2000.0, False, True, 2.0, False),
    ]
    out = {"n_cases": np.int64(len(specs)), "agent_keys": np.array(AGENT_KEYS + GEO_KEYS)}
    for c, spec in enumerate(specs):
        cfg, st, dth = mgb.scene(rng, *spec)
        tab, plist = behave_params(rng, spec[0], cfg, hetero_geometry=(c >= 3))
        fields, res = run_reference(cfg, st, dth, plist)
        p = f"c{c}_"
        out[p + "cfg"] = np.array([float(getattr(cfg, k)) for k in mgb.CFG_KEYS])
        out[p + "fov"] = np.array(cfg.fov)
        for k in mgb.STATE_KEYS:
            out[p + "st_" + k] = np.asarray(st[k])
        out[p + "dth"] = dth
        out[p + "agent_params"] = np.stack([tab[k] for k in AGENT_KEYS + GEO_KEYS], axis=1)
        out[p + "fields"] = pack_bits(fields)
        for k in mgb.OUT_KEYS:
            out[p + "out_" + k] = res[k]
    np.savez_compressed(os.path.join(HERE, "base_hetero_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "base_hetero_golden.npz"))


if __name__ == "__main__":
    main()
