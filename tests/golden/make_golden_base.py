"""Generates tests/golden/base_golden.npz by EXECUTING THE UNMODIFIED REFERENCE's
Agent.update (abm/agent/agent.py:212-283, imported through oracle/ref_shim.py) on seeded
scenes, every agent from the same frozen snapshot.  The np.random.uniform draw
of supcalc.random_walk is replaced by a recorded value per agent (it is an INPUT of the
one-step parity contract).  Build container only:  python tests/golden/make_golden_base.py
"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle import restate_base as rb  # noqa: E402
from oracle.restate import pack_bits  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CFG_KEYS = ["R", "width", "height", "vision_range", "visual_exclusion", "patchwise_exclusion", "T_w", "Eps_w", "g_w",
            "B_w", "w_max", "T_u", "Eps_u", "g_u", "B_u", "u_max", "S_wu", "S_uw", "Tau", "F_N", "F_R", "exp_vel_max",
            "exp_theta_min", "exp_theta_max", "reloc_theta_max", "exp_stop_ratio"]
STATE_KEYS = ["x", "y", "theta", "vel", "w", "u", "novelty", "env_status", "override", "mode", "patch_id", "collected",
              "collected_before"]
OUT_KEYS = ["x", "y", "theta", "vel", "w", "u", "I_priv", "override", "mode", "collected_before"]


def f32(a):
    return np.asarray(a, np.float32).astype(np.float64)


def scene(rng, N, R, W, fovr, vision_range, vis_excl, patchwise, eps_w, intpos):
    cfg = rb.BaseConfig(R=R, fov=(-fovr * np.pi, fovr * np.pi), width=W, height=W, vision_range=vision_range,
                        visual_exclusion=vis_excl, patchwise_exclusion=patchwise, Eps_w=eps_w, Eps_u=1.0,
                        S_wu=0.25, S_uw=0.01, F_N=0.5, F_R=0.5, exp_vel_max=3.0, exp_theta_min=-0.5,
                        exp_theta_max=0.5, reloc_theta_max=1.8, exp_stop_ratio=0.175)
    if intpos:   # create_agents uses integer positions (sims.py:530-531)
        x = rng.integers(20, 30 + int(W), N).astype(float)
        y = rng.integers(20, 30 + int(W), N).astype(float)
    else:
        x, y = f32(rng.uniform(20, 30 + W, N)), f32(rng.uniform(20, 30 + W, N))
    override = rng.choice([0, 0, 1, 1, 3], N)
    st = dict(x=x, y=y, theta=f32(rng.uniform(0, 2 * np.pi, N)), vel=f32(rng.uniform(0, 3, N)), radius=10.0,
              w=f32(rng.uniform(-0.2, 1, N)), u=f32(rng.uniform(-0.2, 1, N)),
              novelty=(rng.uniform(0, 1, (N, cfg.Tau)) < 0.15).astype(float),
              env_status=rng.choice([-1, 1], N), override=override,
              mode=np.where(override == 1, 1, np.where(override == 3, 3, 0)),
              patch_id=rng.choice([-1, 0, 1, 2], N), collected=f32(rng.uniform(0, 5, N)))
    st["collected_before"] = f32(st["collected"] - rng.choice([0.0, 0.25, 1.0], N))
    dth = f32(rng.uniform(cfg.exp_theta_min, cfg.exp_theta_max, N))
    return cfg, st, dth


def run_reference(cfg, st, dth):
    agents = ref_shim.make_base_agents(st, cfg)
    N = len(agents)
    fields = np.zeros((N, cfg.R), bool)
    res = {k: np.zeros(N) for k in OUT_KEYS}
    ov = {None: 0, "exploit": 1, "collide": 3}
    md = {"explore": 0, "exploit": 1, "relocate": 2, "collide": 3}
    for i in range(N):
        cp = copy.deepcopy(agents)
        a = cp[i]
        orig = np.random.uniform
        np.random.uniform = lambda lo, hi, _v=dth[i]: _v
        try:
            a.update(cp)
        finally:
            np.random.uniform = orig
        fields[i, :len(a.soc_v_field)] = a.soc_v_field > 0     # (own v_field_res <= cfg.R: heterogeneous agents)
        vals = dict(x=a.position[0], y=a.position[1], theta=a.orientation, vel=a.velocity, w=a.w, u=a.u,
                    I_priv=a.I_priv, override=ov[a.overriding_mode], mode=md[a.mode],
                    collected_before=a.collected_r_before)
        for k in OUT_KEYS:
            res[k][i] = vals[k]
    return fields, res


def main():
    if not ref_shim.reference_available():
        raise SystemExit("reference tree not available")
    rng = np.random.default_rng(20261018)
    specs = [
        # N, R, W, fov, vision_range, visual_exclusion, patchwise, Eps_w, integer positions
        (10, 1200, 500, 1.0, 2000, False, True, 0.0, True),      # config 1 shape (reference .env: no exclusion)
        (50, 1200, 500, 1.0, 2000, True, True, 2.0, True),       # config 3 shape (figExp3BN50PatchyCollOcc)
        (50, 1200, 500, 1.0, 2000, True, False, 5.0, False),
        (30, 1200, 500, 0.5, 2000, True, True, 1.0, True),       # limited FOV (reference .env AGENT_FOV=0.5)
        (24, 320, 300, 0.9, 150, True, True, 0.75, False),       # limited vision range
        (16, 1201, 200, 1.0, 2000, True, True, 3.0, True),       # odd R, crowded
        (40, 600, 250, 0.25, 2000, False, False, 0.25, True),
        (2, 1200, 500, 1.0, 2000, True, True, 2.0, True),
    ]
    out = {"n_cases": np.int64(len(specs))}
    for c, sp in enumerate(specs):
        cfg, st, dth = scene(rng, *sp)
        fields, res = run_reference(cfg, st, dth)
        p = f"c{c}_"
        out[p + "cfg"] = np.array([float(getattr(cfg, k)) for k in CFG_KEYS])
        out[p + "fov"] = np.array(cfg.fov)
        for k in STATE_KEYS:
            out[p + "st_" + k] = np.asarray(st[k])
        out[p + "dth"] = dth
        out[p + "fields"] = pack_bits(fields)
        for k in OUT_KEYS:
            out[p + "out_" + k] = res[k]
    np.savez_compressed(os.path.join(OUT, "base_golden.npz"), **out)
    print("wrote", os.path.join(OUT, "base_golden.npz"))


if __name__ == "__main__":
    main()
