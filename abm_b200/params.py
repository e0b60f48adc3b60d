"""The reference's `.env` parameter surface: parsing rules of abm/app.py:27-63 /
app_visual_flocking.py:70-106 and the module-level constants of abm/contrib/decision_params.py,
movement_params.py and projects/visual_flocking/vf_contrib/vf_params.py -- here as plain
dataclasses built from one env dictionary instead of modules re-imported per agent."""
from __future__ import annotations

import os
import re
from dataclasses import dataclass


def read_env(path: str) -> dict:
    """KEY=VALUE lines ('#' comments, optional quotes) -- the subset of python-dotenv's
    dotenv_values the reference's files use."""
    out = {}
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#") or "=" not in line:
                continue
            k, v = line.split("=", 1)
            v = v.strip()
            if v[:1] in ("\"", "'") and v.find(v[0], 1) > 0:     # quoted: up to the closing quote (a comment may follow)
                v = v[1:v.find(v[0], 1)]
            else:                                                # unquoted: an inline comment starts at whitespace + '#'
                v = re.sub(r"\s+#.*", "", v).rstrip()
            out[k.strip()] = v
    return out


def env_path(root_dir: str | None = None) -> str:
    """{EXPERIMENT_NAME}.env at the repository root (sims.py:22-26)."""
    root = root_dir or os.getcwd()
    return os.path.join(root, f"{os.getenv('EXPERIMENT_NAME', '')}.env")


def _i(v):
    return int(float(v))


def _b(v):
    return bool(int(float(v)))


def simulation_kwargs(envconf: dict) -> dict:
    """.env -> Simulation / VFSimulation constructor kwargs (app.py:27-63)."""
    g = envconf.get
    return dict(
        N=_i(envconf["N"]), T=_i(envconf["T"]), v_field_res=int(envconf["VISUAL_FIELD_RESOLUTION"]),
        agent_fov=float(envconf["AGENT_FOV"]), framerate=_i(g("INIT_FRAMERATE", 25)),
        with_visualization=_b(g("WITH_VISUALIZATION", 0)), width=_i(envconf["ENV_WIDTH"]),
        height=_i(envconf["ENV_HEIGHT"]), show_vis_field=_b(g("SHOW_VISUAL_FIELDS", 0)),
        show_vis_field_return=_b(g("SHOW_VISUAL_FIELDS_RETURN", 0)), pooling_time=_i(g("POOLING_TIME", 0)),
        pooling_prob=float(g("POOLING_PROBABILITY", 0.05)), agent_radius=_i(envconf["RADIUS_AGENT"]),
        N_resc=_i(g("N_RESOURCES", 0)), allow_border_patch_overlap=_b(g("PATCH_BORDER_OVERLAP", 1)),
        min_resc_perpatch=_i(g("MIN_RESOURCE_PER_PATCH", 100)), max_resc_perpatch=_i(g("MAX_RESOURCE_PER_PATCH", -1)),
        min_resc_quality=float(g("MIN_RESOURCE_QUALITY", 0.25)), max_resc_quality=float(g("MAX_RESOURCE_QUALITY", -1)),
        patch_radius=_i(g("RADIUS_RESOURCE", 30)), regenerate_patches=_b(g("REGENERATE_PATCHES", 1)),
        agent_consumption=_i(g("AGENT_CONSUMPTION", 1)), ghost_mode=_b(g("GHOST_WHILE_EXPLOIT", 1)),
        patchwise_exclusion=_b(g("PATCHWISE_SOCIAL_EXCLUSION", 1)), teleport_exploit=_b(g("TELEPORT_TO_MIDDLE", 0)),
        vision_range=_i(g("VISION_RANGE", 2000)), visual_exclusion=_b(g("VISUAL_EXCLUSION", 0)),
        show_vision_range=_b(g("SHOW_VISION_RANGE", 0)), use_ifdb_logging=_b(g("USE_IFDB_LOGGING", 0)),
        use_ram_logging=_b(g("USE_RAM_LOGGING", 0)), save_csv_files=_b(g("SAVE_CSV_FILES", 0)),
        use_zarr=_b(g("USE_ZARR_FORMAT", 1)), window_pad=30, collide_agents=_b(g("AGENT_AGENT_COLLISION", 0)))


@dataclass
class VFParams:
    """vf_contrib/vf_params.py:12-23.  NOTE the reference reads GAM from the key VF_GAMMA while
    every experiment sets VF_GAM, so GAM is always 0.1 there; kept."""
    GAM: float = 0.1
    V0: float = 1.0
    ALP0: float = 1.0
    ALP1: float = 0.09
    ALP2: float = 0.0
    BET0: float = 1.0
    BET1: float = 0.09
    BET2: float = 0.0
    BOUNDARY: str = "walls"
    LIMIT_MOVEMENT: bool = False
    MAX_VEL: float = 3.0
    MAX_TH: float = 0.1

    @classmethod
    def from_env(cls, e: dict) -> "VFParams":
        g = e.get
        return cls(GAM=float(g("VF_GAMMA", 0.1)), V0=float(g("VF_V0", 1)), ALP0=float(g("VF_ALP0", 1)),
                   ALP1=float(g("VF_ALP1", 0.09)), ALP2=float(g("VF_ALP2", 0)), BET0=float(g("VF_BET0", 1)),
                   BET1=float(g("VF_BET1", 0.09)), BET2=float(g("VF_BET2", 0)), BOUNDARY=g("BOUNDARY", "walls"),
                   LIMIT_MOVEMENT=bool(float(g("VF_LIMIT_MOVEMENT", "0"))), MAX_VEL=float(g("VF_MAX_VEL", "3")),
                   MAX_TH=float(g("VF_MAX_TH", "0.1")))


@dataclass
class DecisionParams:
    """contrib/decision_params.py:13-42 and contrib/movement_params.py:13-23."""
    T_w: float = 0.5
    Eps_w: float = 3.0
    g_w: float = 0.085
    B_w: float = 0.0
    w_max: float = 1.0
    T_u: float = 0.5
    Eps_u: float = 3.0
    g_u: float = 0.085
    B_u: float = 0.0
    u_max: float = 1.0
    S_wu: float = 0.25
    S_uw: float = 0.01
    Tau: int = 10
    F_N: float = 2.0
    F_R: float = 1.0
    exp_vel_max: float = 1.0
    exp_theta_min: float = -0.3
    exp_theta_max: float = 0.3
    reloc_theta_max: float = 0.5
    exp_stop_ratio: float = 0.08

    @classmethod
    def from_env(cls, e: dict) -> "DecisionParams":
        g = e.get
        return cls(T_w=float(g("DEC_TW", 0.5)), Eps_w=float(g("DEC_EPSW", 3)), g_w=float(g("DEC_GW", 0.085)),
                   B_w=float(g("DEC_BW", 0)), w_max=float(g("DEC_WMAX", 1)), T_u=float(g("DEC_TU", 0.5)),
                   Eps_u=float(g("DEC_EPSU", 3)), g_u=float(g("DEC_GU", 0.085)), B_u=float(g("DEC_BU", 0)),
                   u_max=float(g("DEC_UMAX", 1)), S_wu=float(g("DEC_SWU", 0.25)), S_uw=float(g("DEC_SUW", 0.01)),
                   Tau=int(float(g("DEC_TAU", 10))), F_N=float(g("DEC_FN", 2)), F_R=float(g("DEC_FR", 1)),
                   exp_vel_max=float(g("MOV_EXP_VEL_MAX", 1)), exp_theta_min=float(g("MOV_EXP_TH_MIN", -0.3)),
                   exp_theta_max=float(g("MOV_EXP_TH_MAX", 0.3)), reloc_theta_max=float(g("MOV_REL_TH_MAX", 0.5)),
                   exp_stop_ratio=float(g("CONS_STOP_RATIO", 0.08)))

    def engine_kwargs(self) -> dict:
        d = {k: getattr(self, k) for k in ("T_w", "Eps_w", "g_w", "B_w", "w_max", "T_u", "Eps_u", "g_u", "B_u",
                                           "u_max", "S_wu", "S_uw", "F_N", "F_R", "exp_vel_max", "exp_theta_min",
                                           "exp_theta_max", "reloc_theta_max", "exp_stop_ratio")}
        return d
