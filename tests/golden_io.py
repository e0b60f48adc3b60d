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


def load_vf_lines_cases():
    """tests/golden/vf_lines_golden.npz (make_golden_lines.py): VF agents with lines to follow -- the unmodified
    reference's VFAgent.update with a non-empty ``lines`` and a synthetic ``line_map`` (vf_agent.py:273-276, 80-129)."""
    z = np.load(os.path.join(GOLDEN, "vf_lines_golden.npz"))
    cases = []
    for c in range(int(z["n_cases"])):
        p = f"c{c}_"
        R, W, torus, fov, limit = z[p + "cfg"]
        cases.append(dict(N=len(z[p + "x"]), R=int(R), W=float(W), boundary="infinite" if torus else "walls",
                          fov_ratio=float(fov), limit=bool(limit), x=z[p + "x"], y=z[p + "y"], theta=z[p + "theta"],
                          vel=z[p + "vel"], radius=z[p + "radius"], line_map=z[p + "line_map"], fields=z[p + "fields"],
                          new=z[p + "new"]))
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


BASE_CFG_KEYS = ["R", "width", "height", "vision_range", "visual_exclusion", "patchwise_exclusion", "T_w", "Eps_w",
                 "g_w", "B_w", "w_max", "T_u", "Eps_u", "g_u", "B_u", "u_max", "S_wu", "S_uw", "Tau", "F_N", "F_R",
                 "exp_vel_max", "exp_theta_min", "exp_theta_max", "reloc_theta_max", "exp_stop_ratio"]
BASE_STATE_KEYS = ["x", "y", "theta", "vel", "w", "u", "novelty", "env_status", "override", "mode", "patch_id",
                   "collected", "collected_before"]
BASE_OUT_KEYS = ["x", "y", "theta", "vel", "w", "u", "I_priv", "override", "mode", "collected_before"]


def load_base_cases():
    """tests/golden/base_golden.npz (written by make_golden_base.py from real reference output)."""
    from oracle import restate_base as rb
    z = np.load(os.path.join(GOLDEN, "base_golden.npz"))
    cases = []
    for c in range(int(z["n_cases"])):
        p = f"c{c}_"
        vals = dict(zip(BASE_CFG_KEYS, z[p + "cfg"]))
        for k in ("R", "Tau"):
            vals[k] = int(vals[k])
        for k in ("visual_exclusion", "patchwise_exclusion"):
            vals[k] = bool(vals[k])
        cfg = rb.BaseConfig(fov=tuple(z[p + "fov"]), **vals)
        st = {k: z[p + "st_" + k] for k in BASE_STATE_KEYS}
        st["radius"] = 10.0
        cases.append(dict(cfg=cfg, st=st, dth=z[p + "dth"], fields=z[p + "fields"],
                          out={k: z[p + "out_" + k] for k in BASE_OUT_KEYS}))
    return cases


def load_cs_cases():
    """tests/golden/cs_golden.npz (written by make_golden_cs.py from real reference output of
    cooperative_signaling/cs_agent/cs_supcalc.py::projection_field)."""
    z = np.load(os.path.join(GOLDEN, "cs_golden.npz"))
    cases = []
    for c in range(int(z["n_cases"])):
        p = f"c{c}_"
        s = z[p + "scalars"]
        meters = z[p + "meters"]
        cases.append(dict(fov=(s[0], s[1]), R=int(s[2]), pos=np.array([s[3], s[4]]), r=float(s[5]), th=float(s[6]),
                          mps=None if s[7] < 0 else float(s[7]), objs=z[p + "objs"],
                          meters=None if meters.size == 0 else meters, rows=z[p + "rows"]))
    return cases


def load_base_hetero_cases(fixture="base_hetero_golden.npz"):
    """tests/golden/base_hetero_golden.npz (make_golden_hetero.py): heterogeneous agents, the per-agent
    parameters went through the reference's own constructor (agent.py:83-108).  ``agent_cfgs``: one
    BaseConfig per agent; ``agent_params``: dict name -> (N,) array.  With
    fixture="base_hetero_radius_golden.npz" (make_golden_hetero_radius.py) the state's radius is
    the (N,) array of the agents' own radii; with "base_hetero_res_golden.npz" (make_golden_hetero_res.py) every
    agent's BaseConfig has its own R and the fields are padded with zeros up to cfg.R."""
    import dataclasses
    from oracle import restate_base as rb
    z = np.load(os.path.join(GOLDEN, fixture))
    keys = [str(k) for k in z["agent_keys"]]
    cases = []
    for c in range(int(z["n_cases"])):
        p = f"c{c}_"
        vals = dict(zip(BASE_CFG_KEYS, z[p + "cfg"]))
        for k in ("R", "Tau"):
            vals[k] = int(vals[k])
        for k in ("visual_exclusion", "patchwise_exclusion"):
            vals[k] = bool(vals[k])
        cfg = rb.BaseConfig(fov=tuple(z[p + "fov"]), **vals)
        tab = z[p + "agent_params"]
        agent_params = {k: tab[:, j] for j, k in enumerate(keys)}
        agent_cfgs = []
        for i in range(tab.shape[0]):
            kw = {k: float(agent_params[k][i]) for k in keys if k not in ("agent_fov", "agent_radius", "v_field_res")}
            if "v_field_res" in agent_params:                                                   # sims.py:507
                kw["R"] = int(agent_params["v_field_res"][i])
            f = float(agent_params["agent_fov"][i])
            agent_cfgs.append(dataclasses.replace(cfg, fov=(-f * np.pi, f * np.pi), **kw))      # sims.py:506
        st = {k: z[p + "st_" + k] for k in BASE_STATE_KEYS}
        st["radius"] = agent_params["agent_radius"].copy() if "agent_radius" in agent_params else 10.0
        cases.append(dict(cfg=cfg, st=st, dth=z[p + "dth"], fields=z[p + "fields"], agent_cfgs=agent_cfgs,
                          agent_params=agent_params, out={k: z[p + "out_" + k] for k in BASE_OUT_KEYS}))
    return cases
