"""Host-side mirror of the reference's class-level interface: `Simulation`
(abm/simulation/sims.py:59-920) and `VFSimulation` (projects/visual_flocking/vf_simulation/
vf_sims.py:15-397) with the same constructor kwargs, `.start()`, `.prepare_start()`,
`.step_sim()`, `.agents` / `.rescources` iterables and per-agent attributes -- backed by the
CUDA engines.  Rendering, event handling and InfluxDB / zarr logging are out of scope.

Extension over the reference: ``n_replicates`` (default 1) runs that many independent
simulations in one batch; `.agents` views replicate 0, `.replicate_agents(b)` any other.
Agents are updated synchronously (Jacobi) from a frozen snapshot, where the reference updates
them sequentially in place (sims.py:861); see DESIGN.md.
"""
from __future__ import annotations

import os

import numpy as np

from .base_engine import BaseEngine
from .engine import VFEngine
from .params import DecisionParams, VFParams

_MODE_NAMES = {0: "explore", 1: "exploit", 2: "relocate", 3: "collide"}


class _RunOutputs:
    """The reference's end-of-run output (sims.py:878-914 -> ifdb.save_agent_data_RAM / save_resource_data_RAM every
    step, ifdb.save_ifdb_as_csv + env_saver.save_env_vars at the end): with ``use_ram_logging`` and ``save_csv_files``
    a run leaves ``<root>/<SAVE_ROOT_DIR>/<timestamp>/{ag_*.zarr, res_*.zarr, env_params.json}`` (ifdb_params.py:34-40,
    ifdb.py:435-535, env_saver.py:15-28) -- here one such folder PER REPLICATE of the batch, written by
    abm_b200.recorder.  InfluxDB logging and the json dump (USE_ZARR_FORMAT=0) are out of scope and refused."""

    def _init_outputs(self, use_ifdb_logging, use_ram_logging, save_csv_files, use_zarr, save_root_dir, root_dir,
                      env_params):
        if use_ifdb_logging and not use_ram_logging:   # (sims.py:206-208: RAM logging switches InfluxDB logging off -- the
            # reference's experiment files set both)
            raise NotImplementedError("InfluxDB logging is out of scope of abm_b200: use USE_RAM_LOGGING=1")
        self.save_in_ram, self.save_csv_files, self.use_zarr = bool(use_ram_logging), bool(save_csv_files), bool(use_zarr)
        if self.save_csv_files and not self.save_in_ram:                           # sims.py:909-912
            raise Exception("Tried to save simulation data as csv file due to env configuration, "
                            "but IFDB/RAM logging was turned off. Nothing to save! Please turn on IFDB/RAM logging"
                            " or turn off CSV saving feature.")
        if self.save_csv_files and not self.use_zarr:
            raise NotImplementedError("USE_ZARR_FORMAT=0 (json dump of the RAM log) is not supported: zarr arrays only")
        self.save_root_dir = save_root_dir if save_root_dir is not None else "abm/data/simulation_data"
        self.root_dir = root_dir or os.getcwd()
        self.env_params = env_params
        self.saved_dirs = []

    def _run_dirs(self):
        """One timestamped folder per replicate (ifdb_params.py:38-40: SAVE_ROOT_DIR/%Y-%m-%d_%H-%M-%S); the reference's
        sequential runs get distinct stamps from the wall clock, the replicates of a batch from consecutive seconds
        (in replicate order, so sorting the folders gives the order of the sweep)."""
        import datetime
        roots = self.save_root_dir if isinstance(self.save_root_dir, (list, tuple)) else [self.save_root_dir] * self.B
        t0 = datetime.datetime.now()
        used, out = set(), []
        for b in range(self.B):
            k = 0
            while True:
                d = os.path.join(self.root_dir, roots[b], (t0 + datetime.timedelta(seconds=b + k)).strftime("%Y-%m-%d_%H-%M-%S"))
                if d not in used and not os.path.exists(d):
                    break
                k += self.B
            used.add(d)
            out.append(d)
        return out

    def _env_list(self):
        e = self.env_params
        if e is None:
            return None
        return list(e) if isinstance(e, (list, tuple)) else [e] * self.B

    def _run_with_outputs(self, n_steps, make_recorder):
        """Advance n_steps; with RAM logging + saving on, record every step and write the folders at the end."""
        if not (self.save_in_ram and self.save_csv_files):
            self.engine.step(n_steps)                     # (a RAM log that is never saved has no observable effect)
            return
        rec = make_recorder(self._run_dirs(), self._env_list())
        for _ in range(n_steps):
            self.engine.step(1)
            rec.record()                                  # the state AFTER the update, like sims.py:861-887
        self.saved_dirs = rec.close()


class _View:
    """Attribute view of one agent / patch of one replicate; reads go through the owner's
    host cache (one bulk download per step), writes mark the owner dirty (one bulk upload)."""

    def __init__(self, owner, b, i):
        object.__setattr__(self, "_o", owner)
        object.__setattr__(self, "_b", b)
        object.__setattr__(self, "id", i)


class VFAgentView(_View):
    """VFAgent attributes read by ifdb.save_agent_data_RAM and attraction_repulsion_map.py."""
    _SCALARS = {"orientation": "theta", "velocity": "vel"}

    @property
    def position(self):
        c = self._o._cache()
        return np.array([c["x"][self._b, self.id], c["y"][self._b, self.id]], dtype=np.float64)

    @position.setter
    def position(self, p):
        c = self._o._cache()
        c["x"][self._b, self.id], c["y"][self._b, self.id] = p[0], p[1]
        self._o._dirty = True

    def __getattr__(self, name):
        o = object.__getattribute__(self, "_o")
        if name in VFAgentView._SCALARS:
            return float(o._cache()[VFAgentView._SCALARS[name]][self._b, self.id])
        if name == "radius":
            return o.agent_radii
        if name == "soc_v_field":
            return o._fields()[self._b, self.id].astype(np.float64)
        if name in ("dv", "dphi", "ablob", "aedge", "bblob", "bedge"):
            k = ("dv", "dphi", "ablob", "aedge", "bblob", "bedge").index(name)
            return float(o._terms()[self._b, self.id, k])
        if name in ("ALP0", "BET0", "V0"):
            v = o._overrides[name][self._b, self.id]
            return None if np.isnan(v) else float(v)
        if name == "verbose_supcalc":
            return True
        raise AttributeError(name)

    def __setattr__(self, name, value):
        o = self._o
        if name in VFAgentView._SCALARS:
            o._cache()[VFAgentView._SCALARS[name]][self._b, self.id] = value
            o._dirty = True
        elif name in ("ALP0", "BET0", "V0"):
            o._overrides[name][self._b, self.id] = np.nan if value is None else value
            o._overrides_dirty = True
        elif name == "position":
            VFAgentView.position.fset(self, value)
        elif name == "verbose_supcalc":
            pass   # the engine always keeps the six terms when keep_terms is on
        else:
            raise AttributeError(f"cannot set {name}")


class VFSimulation(_RunOutputs):
    """Visual-flocking simulation (vf_sims.py:15-397).  Same kwargs as the reference's
    Simulation.__init__ (sims.py:60-68); the ones the hot path does not read are accepted
    and ignored.  ``vf_params``: a params.VFParams (defaults = vf_params.py defaults)."""

    def __init__(self, N, T, v_field_res=800, width=600, height=480, framerate=25, window_pad=30,
                 with_visualization=False, agent_radius=10, agent_fov=1.0, agent_behave_param_list=None,
                 vf_params: VFParams | None = None, n_replicates: int = 1, device: int = 0, seed=None,
                 keep_fields: bool = True, use_ifdb_logging=False, use_ram_logging=False, save_csv_files=False,
                 use_zarr=True, save_root_dir=None, root_dir=None, env_params=None, **ignored):
        if with_visualization:
            raise NotImplementedError("rendering is out of scope of abm_b200 (headless only)")
        self.B = int(n_replicates)
        self._init_outputs(use_ifdb_logging, use_ram_logging, save_csv_files, use_zarr, save_root_dir, root_dir, env_params)
        # vf_sims.py:184-210, 222-228: the visual-flocking simulation accepts the list but constructs every VFAgent
        # with behave_params=None -- the dictionaries change nothing; kept for the evolution summary only (:368-371)
        self.agent_behave_param_list = agent_behave_param_list
        self.heterogen_agents = agent_behave_param_list is not None
        self.N, self.T, self.t = int(N), int(T), 0
        self.WIDTH, self.HEIGHT, self.window_pad = width, height, window_pad
        self.agent_radii = agent_radius
        self.fov_ratio = agent_fov
        self.agent_fov = (-agent_fov * np.pi, agent_fov * np.pi)                  # sims.py:160-161
        self.v_field_res = int(v_field_res * (1 / agent_fov))                     # vf_sims.py:41-44
        self.vf_params = vf_params or VFParams()
        self.B = int(n_replicates)
        self._rng = np.random if seed is None else np.random.RandomState(seed)
        p = self.vf_params
        self.engine = VFEngine(self.B, self.N, resolution=self.v_field_res, fov=self.agent_fov, boundary=p.BOUNDARY,
                               width=width, height=height, window_pad=window_pad, limit_movement=p.LIMIT_MOVEMENT,
                               max_vel=p.MAX_VEL, max_th=p.MAX_TH, keep_fields=keep_fields, keep_terms=True,
                               device=device)
        self.engine.set_params(GAM=p.GAM, V0=p.V0, ALP0=p.ALP0, ALP1=p.ALP1, BET0=p.BET0, BET1=p.BET1)
        self.agents = []
        self._state = None
        self._dirty = False
        self._stale = True
        self._overrides = {k: np.full((self.B, self.N), np.nan, np.float32) for k in ("ALP0", "BET0", "V0")}
        self._overrides_dirty = False
        self._f = self._t = None

    # -- reference API ---------------------------------------------------------------------
    def create_agents(self):
        """vf_sims.py:212-228: heading ~ U(0, 2pi), the SAME angle places the agent on a disc."""
        st = {k: np.zeros((self.B, self.N), np.float32) for k in ("x", "y", "theta", "vel")}
        for b in range(self.B):
            for i in range(self.N):
                orient = self._rng.uniform(0, 2 * np.pi)
                dist = self._rng.rand() * (self.HEIGHT / 2 - 2 * self.window_pad - self.agent_radii)
                st["x"][b, i] = np.cos(orient) * dist + self.WIDTH / 2
                st["y"][b, i] = np.sin(orient) * dist + self.HEIGHT / 2
                st["theta"][b, i] = orient
        self._state, self._dirty, self._stale = st, True, False
        self.agents = [VFAgentView(self, 0, i) for i in range(self.N)]

    def replicate_agents(self, b):
        return [VFAgentView(self, b, i) for i in range(self.N)]

    def prepare_start(self):
        """vf_sims.py:342-352."""
        self.create_agents()

    def set_line_map(self, line_map=None, sensor_radius=9, sensor_distance=20):
        """Lines for the agents to follow: what the reference's GUI leaves in every ``agent.line_map`` after lines were
        drawn (vf_agent.py:246-260; shape (WIDTH + window_pad, HEIGHT + window_pad)); the agents then steer by
        vf_supcalc.follow_lines_local instead of the flocking heading change (vf_agent.py:273-276).  None: no lines."""
        self.engine.set_line_map(line_map, sensor_radius, sensor_distance)

    def step_sim(self):
        """vf_sims.py:291-340: one time step for every agent of every replicate."""
        self._sync_up()
        self.engine.step(1)
        self.t += 1
        self._stale = True
        self._f = self._t = None

    def start(self):
        """vf_sims.py:355-366."""
        self.prepare_start()
        self._sync_up()
        from .recorder import VFRecorder
        self._run_with_outputs(self.T - self.t, lambda dirs, envs: VFRecorder(self.engine, None, dirs=dirs, env_params=envs))
        self.t = self.T
        self._stale = True
        self._f = self._t = None

    # -- plumbing --------------------------------------------------------------------------
    def _cache(self):
        if self._stale:
            self._state = self.engine.get_state()
            self._stale = False
        return self._state

    def _sync_up(self):
        if self._state is None:
            raise RuntimeError("prepare_start() has not been called")
        if self._dirty:
            s = self._state
            self.engine.set_state(s["x"], s["y"], s["theta"], s["vel"], float(self.agent_radii))
            self._dirty = False
        if self._overrides_dirty:
            o = self._overrides
            self.engine.set_agent_overrides(o["ALP0"], o["BET0"], o["V0"])
            self._overrides_dirty = False

    def _fields(self):
        if self._f is None:
            self._f = self.engine.fields()
        return self._f

    def _terms(self):
        if self._t is None:
            self._t = self.engine.terms()
        return self._t


class AgentView(_View):
    """Agent attributes read by ifdb.save_agent_data_RAM (ifdb.py:83-96)."""
    _MAP = {"orientation": "theta", "velocity": "vel", "w": "w", "u": "u", "I_priv": "i_priv",
            "collected_r": "collected", "exploited_patch_id": "patch_id", "env_status": "env_status"}

    @property
    def position(self):
        c = self._o._cache()
        return np.array([c["x"][self._b, self.id], c["y"][self._b, self.id]], dtype=np.float64)

    def __getattr__(self, name):
        o = object.__getattribute__(self, "_o")
        if name in AgentView._MAP:
            v = o._cache()[AgentView._MAP[name]][self._b, self.id]
            return int(v) if name in ("exploited_patch_id", "env_status") else float(v)
        if name == "mode":
            return _MODE_NAMES[int(o._cache()["mode"][self._b, self.id])]
        if name == "radius":
            rl = getattr(o, "agent_radii_list", None)                # sims.py:502: the agent's own radius
            return o.agent_radii if rl is None else float(rl[self.id])
        if name == "v_field_res":
            rl = getattr(o, "agent_res_list", None)                  # sims.py:507: the agent's own resolution
            return o.engine.R if rl is None else int(rl[self.id])
        if name == "soc_v_field":
            return o._fields()[self._b, self.id, :self.v_field_res].astype(np.float64)   # agent.py:73: its own length
        raise AttributeError(name)

    def get_mode(self):
        """Agent.get_mode (agent.py:659-669)."""
        c = self._o._cache()
        ov = int(c["override_mode"][self._b, self.id])
        if ov == 1:
            return "exploit"
        if ov == 3:
            return "collide"
        return "relocate" if c["w"][self._b, self.id] > self._o.decision_params.T_w else "explore"


class ResourceView(_View):
    """Rescource attributes read by ifdb.save_resource_data_RAM (ifdb.py:250-253)."""

    def __getattr__(self, name):
        o = object.__getattribute__(self, "_o")
        p = o._patch_cache()
        if name == "position":
            return np.array([p["x"][self._b, self.id], p["y"][self._b, self.id]], dtype=np.float64)
        m = {"radius": "radius", "resc_left": "left", "unit_per_timestep": "quality"}
        if name in m:
            return float(p[m[name]][self._b, self.id])
        raise AttributeError(name)


class Simulation(_RunOutputs):
    """Collective-foraging simulation (sims.py:59-920), headless.  Same kwargs as the
    reference; ``decision_params``: a params.DecisionParams (defaults = the param modules)."""

    def __init__(self, N, T, v_field_res=800, width=600, height=480, framerate=25, window_pad=30,
                 with_visualization=False, show_vis_field=False, show_vis_field_return=False, pooling_time=0,
                 pooling_prob=0.05, agent_radius=10, N_resc=10, min_resc_perpatch=200, max_resc_perpatch=1000,
                 min_resc_quality=0.1, max_resc_quality=1, patch_radius=30, regenerate_patches=True,
                 agent_consumption=1, teleport_exploit=True, vision_range=150, agent_fov=1.0, visual_exclusion=False,
                 show_vision_range=False, use_ifdb_logging=False, use_ram_logging=False, save_csv_files=False,
                 ghost_mode=True, patchwise_exclusion=True, parallel=False, use_zarr=True,
                 allow_border_patch_overlap=False, agent_behave_param_list=None, collide_agents=False,
                 decision_params: DecisionParams | None = None, n_replicates: int = 1, device: int = 0, seed=None,
                 keep_fields: bool = True, save_root_dir=None, root_dir=None, env_params=None):
        if with_visualization:
            raise NotImplementedError("rendering is out of scope of abm_b200 (headless only)")
        self.B = int(n_replicates)
        self._init_outputs(use_ifdb_logging, use_ram_logging, save_csv_files, use_zarr, save_root_dir, root_dir, env_params)
        if pooling_time != 0:
            raise NotImplementedError("POOLING_TIME != 0 is not supported (every reference experiment uses 0)")
        self.heterogen_agents = agent_behave_param_list is not None                # sims.py:170-173
        agent_radii_list = agent_res_list = None
        if self.heterogen_agents:
            agent_radius, v_field_res, agent_radii_list, agent_res_list = self._check_behave_params(
                agent_behave_param_list, int(N), agent_radius, v_field_res,
                (decision_params or DecisionParams()).Tau)
        self.agent_behave_param_list = agent_behave_param_list
        self.N, self.T, self.t = int(N), int(T), 0
        self.WIDTH, self.HEIGHT, self.window_pad = width, height, window_pad
        self.agent_radii, self.N_resc, self.resc_radius = agent_radius, int(N_resc), patch_radius
        self.allow_border_patch_overlap = allow_border_patch_overlap
        self.min_resc_units, self.max_resc_units = min_resc_perpatch, max_resc_perpatch
        self.min_resc_quality, self.max_resc_quality = min_resc_quality, max_resc_quality
        if self.max_resc_quality < 0:                                             # sims.py:176-179
            self.max_resc_quality = self.min_resc_quality
        if self.max_resc_units < 0:
            self.max_resc_units = self.min_resc_units + 1
        self.decision_params = decision_params or DecisionParams()
        self.B = int(n_replicates)
        self._rng = np.random if seed is None else np.random.RandomState(seed)
        # the engine's counter-based RNG (random-walk turns, patch regeneration) is keyed by a 64-bit seed: an unseeded
        # run draws a fresh one, like the reference's unseeded np.random; kept on the simulation for reproducibility
        self.engine_seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        self.engine = BaseEngine(self.B, self.N, self.N_resc, resolution=int(v_field_res), agent_fov=agent_fov,
                                 width=width, height=height, window_pad=window_pad, vision_range=vision_range,
                                 agent_radius=agent_radius, visual_exclusion=visual_exclusion,
                                 patchwise_exclusion=patchwise_exclusion, teleport_exploit=teleport_exploit,
                                 regenerate_patches=regenerate_patches, patch_border_overlap=allow_border_patch_overlap,
                                 patch_radius=patch_radius, min_resc_quality=min_resc_quality,
                                 max_resc_quality=max_resc_quality, min_resc_perpatch=min_resc_perpatch,
                                 max_resc_perpatch=max_resc_perpatch, tau=self.decision_params.Tau,
                                 keep_fields=keep_fields, collide_agents=collide_agents, ghost_mode=ghost_mode,
                                 seed=self.engine_seed, device=device)
        prm = dict(agent_consumption=agent_consumption, **self.decision_params.engine_kwargs())
        if self.heterogen_agents:
            # agent.py:83-108: these keys of behave_params replace the param modules' values, per agent;
            # sims.py:509 agent_consumption.  The same list is used in every replicate.
            for key in self._BEHAVE_DECISION_KEYS + ("exp_vel_max", "exp_stop_ratio", "agent_consumption"):
                row = np.array([float(bp[key]) for bp in agent_behave_param_list], np.float64)
                prm[key] = np.broadcast_to(row, (self.B, self.N)).copy()
            # sims.py:506, 511: every agent is constructed with its own FOV and vision_range
            self.engine.set_agent_geometry(
                agent_fov=np.array([float(bp["agent_fov"]) for bp in agent_behave_param_list]),
                vision_range=np.array([float(bp["vision_range"]) for bp in agent_behave_param_list]))
            if agent_radii_list is not None:                                       # sims.py:502: the agent's own radius
                self.engine.set_agent_radii(agent_radii_list)
            if agent_res_list is not None:                                         # sims.py:507: its own v_field_res
                self.engine.set_agent_resolution(agent_res_list)
        self.agent_radii_list, self.agent_res_list = agent_radii_list, agent_res_list
        self.engine.set_params(**prm)
        self.agents, self.rescources = [], []
        self._a = self._p = self._f = None

    _BEHAVE_DECISION_KEYS = ("S_wu", "T_w", "Eps_w", "g_w", "B_w", "w_max", "S_uw", "T_u", "Eps_u", "g_u", "B_u",
                             "u_max", "F_N", "F_R")

    @staticmethod
    def _check_behave_params(plist, N, agent_radius, v_field_res, tau):
        """agent_behave_param_list (sims.py:499-517, template: contrib/evolution.py:1-26): the decision / movement
        entries, agent_consumption, agent_fov, vision_range, agent_radius and v_field_res may differ between agents
        (the engine's resolution -- its row stride -- is then the largest); Tau and pooling must be the same for all
        agents (they are engine-wide) and then replace the constructor's values like the reference does.
        Returns (engine-wide radius, engine resolution, per-agent radii or None, per-agent resolutions or None)."""
        if len(plist) != N:
            raise ValueError("agent_behave_param_list must hold one dictionary per agent")
        res = np.array([int(bp.get("v_field_res", v_field_res)) for bp in plist])
        if any(int(bp.get("Tau", tau)) != int(tau) for bp in plist):
            raise NotImplementedError("agent_behave_param_list: 'Tau' must equal decision_params.Tau for all agents")
        if any(float(bp.get("pooling_time", 0)) != 0 for bp in plist):
            raise NotImplementedError("POOLING_TIME != 0 is not supported (every reference experiment uses 0)")
        radii = np.array([float(bp.get("agent_radius", agent_radius)) for bp in plist])
        radius = float(radii[0])
        hetero = None if np.all(radii == radii[0]) else radii
        return ((int(radius) if radius.is_integer() else radius), int(res.max()), hetero,
                None if np.all(res == res[0]) else res)

    def create_agents(self):
        """sims.py:526-537: integer positions, heading ~ U(0, 2pi)."""
        r, pad = self.agent_radii, self.window_pad
        x = self._rng.randint(pad - r, self.WIDTH + pad - r, (self.B, self.N))
        y = self._rng.randint(pad - r, self.HEIGHT + pad - r, (self.B, self.N))
        th = self._rng.uniform(0, 2 * np.pi, (self.B, self.N))
        self.engine.set_agents(x=x, y=y, theta=th)
        self.agents = [AgentView(self, 0, i) for i in range(self.N)]

    def set_replicate_patch_params(self, patch_radius, min_resc_perpatch, max_resc_perpatch, min_resc_quality,
                                   max_resc_quality):
        """One set of patch parameters per replicate (length-B arrays): a sweep over RADIUS_RESOURCE / the resource units /
        the quality as ONE batch.  The rules for negative maxima are the constructor's (sims.py:176-179), per replicate."""
        a = lambda v: np.broadcast_to(np.asarray(v, np.float64), (self.B,)).copy()
        R, u0, u1, q0, q1 = a(patch_radius), a(min_resc_perpatch), a(max_resc_perpatch), a(min_resc_quality), a(max_resc_quality)
        q1 = np.where(q1 < 0, q0, q1)
        u1 = np.where(u1 < 0, u0 + 1, u1)
        self._patch_params = dict(R=R, u0=u0, u1=u1, q0=q0, q1=q1)
        self.engine.set_regeneration_params(patch_radius=R, min_resc_quality=q0, max_resc_quality=q1, min_resc_perpatch=u0,
                                            max_resc_perpatch=u1)

    def create_resources(self):
        """sims.py:539-541 -> add_new_resource_patch (:332-374): rejection-sampled, no patch-patch overlap."""
        P = self.N_resc
        pp = getattr(self, "_patch_params", None)
        pa = {k: np.zeros((self.B, P), np.float32) for k in ("x", "y", "radius", "left", "quality")}
        pa["id"] = np.tile(np.arange(1, P + 1, dtype=np.int32), (self.B, 1))
        for b in range(self.B):
            R = self.resc_radius if pp is None else int(pp["R"][b]) if float(pp["R"][b]).is_integer() else float(pp["R"][b])
            u0, u1 = (self.min_resc_units, self.max_resc_units) if pp is None else (int(pp["u0"][b]), int(pp["u1"][b]))
            q0, q1 = (self.min_resc_quality, self.max_resc_quality) if pp is None else (float(pp["q0"][b]), float(pp["q1"][b]))
            for p in range(P):
                for _retry in range(10000):
                    if self.allow_border_patch_overlap:
                        x = self._rng.randint(self.window_pad - R, self.WIDTH + self.window_pad - R)
                        y = self._rng.randint(self.window_pad - R, self.HEIGHT + self.window_pad - R)
                    else:
                        x = self._rng.randint(self.window_pad, self.WIDTH + self.window_pad - 2 * R)
                        y = self._rng.randint(self.window_pad, self.HEIGHT + self.window_pad - 2 * R)
                    units = self._rng.randint(u0, u1)
                    quality = self._rng.uniform(q0, q1)
                    d2 = (pa["x"][b, :p] - x) ** 2 + (pa["y"][b, :p] - y) ** 2
                    if not (d2 <= (2 * R) ** 2).any():
                        break
                else:
                    raise Exception("Reached timeout while trying to create resources without overlap!")
                pa["x"][b, p], pa["y"][b, p], pa["radius"][b, p] = x, y, R
                pa["left"][b, p], pa["quality"][b, p] = units, quality
        if P:
            self.engine.set_patches(**pa)
        self.rescources = [ResourceView(self, 0, p) for p in range(P)]

    def step_sim(self):
        """One pass of the main loop body (sims.py:733-864)."""
        self.engine.step(1)
        self.t += 1
        self._a = self._p = self._f = None

    def start(self):
        """sims.py:702-920 without rendering / logging."""
        self.create_agents()
        self.create_resources()
        from .recorder import BaseRecorder
        self._run_with_outputs(self.T - self.t, lambda dirs, envs: BaseRecorder(self.engine, None, dirs=dirs, env_params=envs))
        self.t = self.T
        self._a = self._p = self._f = None

    def replicate_agents(self, b):
        return [AgentView(self, b, i) for i in range(self.N)]

    def _cache(self):
        if self._a is None:
            self._a = self.engine.get_agents()
        return self._a

    def _patch_cache(self):
        if self._p is None:
            self._p = self.engine.get_patches()
        return self._p

    def _fields(self):
        if self._f is None:
            self._f = self.engine.fields()
        return self._f
