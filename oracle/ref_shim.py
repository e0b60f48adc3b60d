"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (abm_b200/).

Harness glue that lets the UNMODIFIED reference (scioip34/ABM, mounted read-only
at /root/reference in the build container) be imported without pygame /
matplotlib, so that its own functions and agent classes can be executed to

  * validate the restatements in ``oracle/restate.py`` / ``oracle/literal.py``, and
  * generate the golden fixtures under ``tests/golden/`` (see
    ``tests/golden/make_golden.py``).

``/root/reference`` does not exist on the GPU box: there the shim falls back to ``oracle/_ref`` (a byte-for-byte,
git-ignored copy of the hot-path modules made by ``oracle/build_ref.py`` in the build container, shipped with the
snapshot like the built ``.so``); nothing reads ``/root/reference`` at run time on the GPU box, and everything that
uses the shim checks ``reference_available()`` first.

What is stubbed: ``pygame`` (Sprite / Group with a
sequential ``update``, no-op Surface, MagicMock drawing sub-modules),
``matplotlib.cm.get_cmap``, and ``scipy.integrate.trapz`` (removed in
SciPy >= 1.14; the reference calls it at vf_supcalc.py:234-246).
Nothing under /root/reference is edited.
"""
import os
import sys
import types
from unittest.mock import MagicMock

def _reference_root() -> str:
    """The mounted reference tree (build container), else the byte-for-byte copy of its hot-path modules that
    oracle/build_ref.py leaves in the git-ignored oracle/_ref (travels to the GPU box with the snapshot)."""
    env = os.environ.get("ABM_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/abm/agent"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "abm", "agent", "agent.py"))


class _Rect:
    def __init__(self):
        self.x = 0
        self.y = 0
        self.centerx = 0
        self.centery = 0

    def collidepoint(self, *_a, **_k):
        return False


class _Surface:
    def __init__(self, *_a, **_k):
        pass

    def fill(self, *_a, **_k):
        pass

    def set_colorkey(self, *_a, **_k):
        pass

    def set_alpha(self, *_a, **_k):
        pass

    def blit(self, *_a, **_k):
        pass

    def get_rect(self, *_a, **_k):
        return _Rect()


class _Sprite:
    def __init__(self, *_a, **_k):
        self._groups = []

    def kill(self):
        for g in list(self._groups):
            if self in g:
                g.remove(self)
        self._groups = []


class _Group(list):
    """Insertion-ordered group whose update() walks a snapshot sequentially
    (what pygame.sprite.Group.update does; reference call sims.py:861)."""

    def add(self, *sprites):
        for s in sprites:
            if s not in self:
                self.append(s)
                if hasattr(s, "_groups"):
                    s._groups.append(self)

    def sprites(self):
        return list(self)

    def update(self, *args, **kwargs):
        for s in list(self):
            s.update(*args, **kwargs)


def _collide_circle(left, right):
    """pygame.sprite.collide_circle as pygame DOCUMENTS it (pygame itself is not in the image and not vendored by the
    reference, so this arithmetic cannot be pinned -- SURVEY 8c): circles around the rect centres, the sprites' own
    `radius` attributes, touching counts as colliding."""
    dx = left.rect.centerx - right.rect.centerx
    dy = left.rect.centery - right.rect.centery
    rs = left.radius + right.radius
    return dx * dx + dy * dy <= rs * rs


def _groupcollide(groupa, groupb, dokilla, dokillb, collided=None):
    """pygame.sprite.groupcollide (documented semantics, no kills needed by the reference's calls): {sprite of groupa:
    [colliding sprites of groupb]} for the sprites of groupa with at least one collision, in group order."""
    out = {}
    for a in list(groupa):
        hits = [b for b in list(groupb) if collided(a, b)]
        if hits:
            out[a] = hits
    return out


def install():
    """Install the stubs and put the reference on sys.path.  Idempotent."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "pygame" not in sys.modules or not getattr(sys.modules["pygame"], "_abm_stub", False):
        pg = types.ModuleType("pygame")
        pg._abm_stub = True
        sprite = types.ModuleType("pygame.sprite")
        sprite.Sprite = _Sprite
        sprite.Group = _Group
        sprite.collide_circle = _collide_circle
        sprite.groupcollide = _groupcollide
        pg.sprite = sprite
        pg.Surface = _Surface
        for name in ("draw", "mask", "gfxdraw", "transform", "math", "display", "time", "event",
                     "font", "mouse", "key", "surfarray"):
            m = MagicMock()
            setattr(pg, name, m)
            sys.modules[f"pygame.{name}"] = m
        pg.init = lambda *a, **k: None
        pg.quit = lambda *a, **k: None
        sys.modules["pygame"] = pg
        sys.modules["pygame.sprite"] = sprite
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        cm = types.ModuleType("matplotlib.cm")
        cm.get_cmap = lambda _name: (lambda _x: (0.0, 0.0, 0.0, 1.0))
        mpl.cm = cm
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.cm"] = cm
    # logging / plotting back ends the hot path never reaches (abm/monitoring/ifdb.py imports them)
    for name in ("influxdb", "zarr", "fastcluster", "xvfbwrapper", "pygame_widgets"):
        try:
            __import__(name)
        except ImportError:
            sys.modules[name] = MagicMock()
    import scipy.integrate as _si
    if not hasattr(_si, "trapz"):
        _si.trapz = _si.trapezoid
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference's param modules read {EXPERIMENT_NAME}.env relative to the reference root
    os.environ.setdefault("EXPERIMENT_NAME", "")


def load_vf():
    """Returns (vf_supcalc, vf_agent module, vf_params) of the real reference."""
    install()
    from abm.projects.visual_flocking.vf_agent import vf_supcalc, vf_agent
    from abm.projects.visual_flocking.vf_contrib import vf_params
    return vf_supcalc, vf_agent, vf_params


def load_base():
    """Returns (supcalc, agent module, decision_params, movement_params)."""
    install()
    from abm.agent import supcalc, agent
    from abm.contrib import decision_params, movement_params
    return supcalc, agent, decision_params, movement_params


def load_loader():
    """The real reference's offline analysis module (abm/loader/data_loader.py: ExperimentLoader with
    calculate_polarization, calculate_interindividual_distance, calculate_mean_NN_dist, calculate_collision_time,
    calculate_search_efficiency, calculate_relocation_time); plotting / zarr back ends stubbed."""
    install()
    import matplotlib
    if "matplotlib.pyplot" not in sys.modules:
        sys.modules["matplotlib.pyplot"] = MagicMock()
        matplotlib.pyplot = sys.modules["matplotlib.pyplot"]
    from abm.loader import data_loader
    return data_loader


def make_loader(tmp_dir, agent_summary, env):
    """An ExperimentLoader of the real reference around in-memory summary arrays, without touching a data folder: what
    its calculate_* methods read.  ``agent_summary``: arrays of shape (batches, agents, T); the loader's convention is
    (batches, *varying parameter dims, agents, T) and some of its means need at least one parameter dimension, so a
    singleton one is inserted: results come back as (batches, 1, ...)."""
    import numpy as np
    dl = load_loader()
    ld = object.__new__(dl.ExperimentLoader)
    ld.experiment_path = str(tmp_dir)
    os.makedirs(os.path.join(str(tmp_dir), "summary"), exist_ok=True)
    agent_summary = {k: np.asarray(v)[:, None] for k, v in agent_summary.items()}
    ld.varying_params, ld.agent_summary, ld.env = {"SWEPT": [0]}, agent_summary, env
    ld.num_batches = next(iter(agent_summary.values())).shape[0]
    ld.iid_matrix, ld.t_end, ld.undersample = None, None, 1
    return ld


def make_vf_agents(x, y, theta, vel, radius, *, R, fov_ratio=1.0, width, height, window_pad=30,
                   boundary="walls", params=None, alp0=None, bet0=None, v0=None,
                   limit_movement=False, max_vel=3.0, max_th=0.1):
    """Build real reference VFAgent objects from arrays (state cast to float64).

    Constructor kwargs follow vf_sims.py:191-208.  vf_params attributes are
    overridden AFTER construction because every constructor reloads them from
    the reference's root .env (vf_agent.py:15-16)."""
    import numpy as np
    vf_supcalc, vf_agent, vf_params = load_vf()
    agents = []
    n = len(x)
    radius = np.broadcast_to(np.asarray(radius), (n,))
    for i in range(n):
        rad = radius[i]
        rad = int(rad) if float(rad).is_integer() else float(rad)
        a = vf_agent.VFAgent(
            id=i, radius=rad, position=(float(x[i]), float(y[i])), orientation=float(theta[i]),
            env_size=(int(width), int(height)), color=(0, 0, 0), v_field_res=R,
            FOV=(-fov_ratio * np.pi, fov_ratio * np.pi), window_pad=int(window_pad), pooling_time=0,
            pooling_prob=0, consumption=1, vision_range=2000, visual_exclusion=False,
            patchwise_exclusion=True, behave_params=None)
        a.velocity = float(vel[i])
        a.boundary_cond = boundary
        a.limit_movement = limit_movement
        a.max_vel = max_vel
        a.max_th = max_th
        if alp0 is not None:
            a.ALP0 = float(alp0[i])
        if bet0 is not None:
            a.BET0 = float(bet0[i])
        if v0 is not None:
            a.V0 = float(v0[i])
        agents.append(a)
    p = dict(GAM=0.1, V0=1.0, ALP0=1.0, ALP1=0.09, ALP2=0.0, BET0=1.0, BET1=0.09, BET2=0.0)
    if params:
        p.update(params)
    for k, v in p.items():
        setattr(vf_params, k, v)
    return agents


DECISION_KEYS = ("T_w", "Eps_w", "g_w", "B_w", "w_max", "T_u", "Eps_u", "g_u", "B_u", "u_max", "S_wu", "S_uw",
                 "F_N", "F_R")


def make_base_agents(st, cfg, behave_params_list=None):
    """Build real reference Agent objects (abm/agent/agent.py) from an oracle state dict
    (see oracle/restate_base.base_step_frozen) and a BaseConfig.  Constructor kwargs follow
    sims.py:481-498; decision / movement parameters are set on the instances and on the
    modules after construction (constructors reload them from the reference's .env)."""
    import numpy as np
    supcalc, agent_mod, decision_params, movement_params = load_base()
    mode_names = {0: "explore", 1: "exploit", 2: "relocate", 3: "collide"}
    agents = []
    N = len(st["x"])
    rad = st["radius"]
    rad = int(rad) if float(rad).is_integer() else float(rad)
    for i in range(N):
        if behave_params_list is not None:                                             # sims.py:502
            rad = behave_params_list[i]["agent_radius"]
        a = agent_mod.Agent(
            id=i, radius=rad, position=(float(st["x"][i]), float(st["y"][i])), orientation=float(st["theta"][i]),
            env_size=(int(cfg.width), int(cfg.height)), color=(0, 0, 0),
            v_field_res=cfg.R if behave_params_list is None else int(behave_params_list[i].get("v_field_res", cfg.R)),  # sims.py:507
            FOV=tuple(cfg.fov) if behave_params_list is None else                      # sims.py:506
            (-float(behave_params_list[i]["agent_fov"]) * np.pi, float(behave_params_list[i]["agent_fov"]) * np.pi),
            window_pad=int(cfg.window_pad), pooling_time=0, pooling_prob=0, consumption=cfg.agent_consumption,
            vision_range=cfg.vision_range if behave_params_list is None else behave_params_list[i]["vision_range"],
            visual_exclusion=cfg.visual_exclusion,
            patchwise_exclusion=cfg.patchwise_exclusion,
            behave_params=None if behave_params_list is None else behave_params_list[i])
        a.velocity = float(st["vel"][i])
        a.w, a.u = float(st["w"][i]), float(st["u"][i])
        a.novelty = np.array(st["novelty"][i], dtype=np.float64)
        a.env_status = int(st["env_status"][i])
        a.exploited_patch_id = int(st["patch_id"][i])
        a.collected_r = float(st["collected"][i])
        a.collected_r_before = float(st["collected_before"][i])
        ov = int(st["override"][i])
        a.overriding_mode = {0: None, 1: "exploit", 3: "collide"}[ov]
        a.mode = mode_names[int(st["mode"][i])]
        if behave_params_list is None:   # else: the reference's own constructor took them from behave_params (agent.py:83-108)
            for k in DECISION_KEYS:
                setattr(a, k, getattr(cfg, k))
            a.max_exp_vel = cfg.exp_vel_max
            a.exp_stop_ratio = cfg.exp_stop_ratio
        agents.append(a)
    movement_params.exp_vel_max = cfg.exp_vel_max
    movement_params.exp_theta_min = cfg.exp_theta_min
    movement_params.exp_theta_max = cfg.exp_theta_max
    movement_params.reloc_theta_max = cfg.reloc_theta_max
    movement_params.exp_stop_ratio = cfg.exp_stop_ratio
    return agents
