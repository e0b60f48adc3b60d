"""Batched mirror of the reference's sweep front-end (abm/metarunner/metarunner.py:30-254):
`Tunable`, `Constant`, `TunedPairRestrain` and `MetaProtocol` with the same methods.  Where
the reference writes one temporary .env per parameter combination x batch and runs them ONE
AFTER THE OTHER (metarunner.py:251-254), `run_protocols` groups the generated configurations
by their batch shape (everything except the per-replicate tunable parameters) and runs each
group as ONE replicate batch on the GPU, one parameter set per replicate."""
from __future__ import annotations

import glob
import itertools
import os
import shutil
import warnings

import numpy as np

from . import params
from .simulation import Simulation, VFSimulation

# .env keys that map to per-replicate engine parameters (everything else defines the batch shape)
VF_REPLICATE_KEYS = {"VF_GAMMA": "GAM", "VF_V0": "V0", "VF_ALP0": "ALP0", "VF_ALP1": "ALP1", "VF_BET0": "BET0",
                     "VF_BET1": "BET1"}
BASE_REPLICATE_KEYS = {"DEC_TW": "T_w", "DEC_EPSW": "Eps_w", "DEC_GW": "g_w", "DEC_BW": "B_w", "DEC_WMAX": "w_max",
                       "DEC_TU": "T_u", "DEC_EPSU": "Eps_u", "DEC_GU": "g_u", "DEC_BU": "B_u", "DEC_UMAX": "u_max",
                       "DEC_SWU": "S_wu", "DEC_SUW": "S_uw", "DEC_FN": "F_N", "DEC_FR": "F_R",
                       "MOV_EXP_VEL_MAX": "exp_vel_max", "MOV_EXP_TH_MIN": "exp_theta_min",
                       "MOV_EXP_TH_MAX": "exp_theta_max", "MOV_REL_TH_MAX": "reloc_theta_max",
                       "CONS_STOP_RATIO": "exp_stop_ratio", "AGENT_CONSUMPTION": "agent_consumption"}


class Tunable:
    """One swept `.env` variable (interface of metarunner.py:59-96): either ``num_data_points`` evenly spaced values
    from ``min_v`` to ``max_v`` inclusive, or the explicit list ``values_override`` (which wins when both are given)."""

    def __init__(self, var_name, min_v=None, max_v=None, num_data_points=None, values_override=None):
        self.name = var_name
        explicit = values_override is not None
        if not explicit and min_v is None:
            raise Exception(f"Tunable {var_name!r} needs either (min_v, max_v, num_data_points) or values_override")
        if explicit and min_v is not None:
            warnings.warn(f"Tunable {var_name!r}: values_override given together with a range; the explicit values are used")
        if explicit:
            self.values = values_override
            self.min_val, self.max_val, self.n_data = min(values_override), max(values_override), len(values_override)
        else:
            self.values = np.linspace(min_v, max_v, num=num_data_points, endpoint=True)
            self.min_val, self.max_val, self.n_data = min_v, max_v, num_data_points
        self.generated = not explicit

    def get_values(self):
        return self.values

    def print(self):
        print(f"Tunable {self.name}: {self.n_data} value(s) in [{self.min_val}, {self.max_val}]")


class Constant:
    """A variable fixed for the whole sweep (interface of metarunner.py:30-43): a Tunable with a single value."""

    def __init__(self, var_name, constant):
        self.tunable = Tunable(var_name, values_override=[constant])
        self.name = var_name

    def get_values(self):
        return self.tunable.get_values()

    def print(self):
        print(f"Constant {self.name} = {self.get_values()[0]}")


class TunedPairRestrain:
    """Keeps only the combinations in which the two named variables multiply to ``restrained_product`` (interface of
    metarunner.py:45-57; with `add_quadratic_tuned_pair`: first * second^2)."""

    def __init__(self, var_name1, var_name2, restrained_product):
        self.var1, self.var2 = var_name1, var_name2
        self.product_restrain = restrained_product

    def get_vars(self):
        return [self.var1, self.var2]

    def print(self):
        print(f"Restraint: {self.var1} x {self.var2} == {self.product_restrain}")


class MetaProtocol:
    """Sweep driver with the methods of metarunner.py:98-254; the runs are batched (module docstring).

    ``default_envconf``: the base `.env` dictionary (default: ``{EXPERIMENT_NAME}.env`` under ``root_dir``, like the
    reference's module-level ``envconf``); ``root_dir``: what the reference calls root_abm_dir (default: cwd)."""

    def __init__(self, experiment_name=None, num_batches=1, parallel=False, description=None, headless=True,
                 default_envconf=None, root_dir=None):
        if parallel and experiment_name is None:
            raise Exception("parallel=True needs an experiment_name (it names the temporary env folder of this sweep)")
        self.root_dir = root_dir or os.getcwd()
        if default_envconf is None:
            base = params.env_path(self.root_dir)
            default_envconf = params.read_env(base) if os.path.isfile(base) else {}
        self.default_envconf = dict(default_envconf)
        self.experiment_name, self.num_batches = experiment_name, num_batches
        self.description, self.headless, self.parallel_run = description, headless, parallel
        self.tunables, self.tuned_pairs, self.q_tuned_pairs = [], [], []
        self.temp_dir = os.path.join("abm/data/metaprotocol/temp", experiment_name) if experiment_name \
            else "abm/data/metaprotocol/temp"
        self.results = []

    def add_criterion(self, criterion):
        self.tunables.append(criterion)

    def add_tuned_pair(self, tuned_pair):
        self.tuned_pairs.append(tuned_pair)

    def add_quadratic_tuned_pair(self, tuned_pair):
        self.q_tuned_pairs.append(tuned_pair)

    def consider_tuned_pairs(self, combos):
        """metarunner.py:129-166."""
        names = [t.name for t in self.tunables]
        keep = []
        for combo in combos:
            ok = True
            for tp in self.tuned_pairs:
                prod = 1
                for n, v in zip(names, combo):
                    if n in tp.get_vars():
                        prod *= v
                ok &= (prod == tp.product_restrain)
            for tp in self.q_tuned_pairs:
                prod = 1
                for n, v in zip(names, combo):
                    if n == tp.get_vars()[0]:
                        prod *= v
                    elif n == tp.get_vars()[1]:
                        prod *= v * v
                ok &= bool(np.isclose(prod, tp.product_restrain))
            if ok:
                keep.append(combo)
        return keep

    def combinations(self):
        names = [t.name for t in self.tunables]
        combos = self.consider_tuned_pairs(list(itertools.product(*[t.get_values() for t in self.tunables])))
        return names, combos

    def generate_temp_env_files(self):
        """metarunner.py:168-199: one {i}_b{nb}.env per combination and batch."""
        temp_dir = os.path.join(self.root_dir, self.temp_dir)
        if os.path.isdir(temp_dir):
            warnings.warn(f"{temp_dir} already holds env files of an earlier sweep: they are replaced")
            shutil.rmtree(temp_dir)
        os.makedirs(temp_dir, exist_ok=True)
        names, combos = self.combinations()
        for nb in range(self.num_batches):
            for i, combo in enumerate(combos):
                env = dict(self.default_envconf)
                for n, v in zip(names, combo):
                    env[n] = int(v) if isinstance(v, bool) else v
                env["SAVE_ROOT_DIR"] = os.path.join("abm/data/simulation_data", self.experiment_name or "UnknownExp",
                                                    f"batch_{nb}")
                with open(os.path.join(temp_dir, f"{i}_b{nb}.env"), "w") as f:
                    for k, v in env.items():
                        f.write(f"{k}={v}\n")
        return len(combos) * self.num_batches

    def save_description(self):
        """README.txt with the sweep's description in the experiment folder (metarunner.py:203-214)."""
        if self.description is None:
            return
        folder = os.path.join(self.root_dir, "abm/data/simulation_data", self.experiment_name or "UnknownExp")
        os.makedirs(folder, exist_ok=True)
        with open(os.path.join(folder, "README.txt"), "w") as f:
            f.write(self.description)

    def run_protocols(self, project="Base", seed=None, keep_env_files=False):
        """All remaining protocols of the temp folder (metarunner.py:240-254), grouped into replicate batches: the env
        files that differ only in per-replicate parameters (and in their SAVE_ROOT_DIR) run as ONE batch, one parameter
        set per replicate.  Every replicate writes what the reference's run of that env file would write --
        ``<root>/<SAVE_ROOT_DIR>/<timestamp>/{ag_*.zarr, res_*.zarr, env_params.json}`` with
        SAVE_ROOT_DIR = abm/data/simulation_data/<experiment>/batch_<nb> -- when the env file asks for it
        (USE_RAM_LOGGING=1, SAVE_CSV_FILES=1, USE_ZARR_FORMAT=1).  ``seed``: None = independent runs (fresh seeds);
        an int makes the sweep reproducible, group g using seed + g.
        Returns the list of (env_paths, simulation) per batch; every simulation holds the final state of its replicates
        and ``saved_dirs``."""
        self.save_description()
        temp_dir = os.path.join(self.root_dir, self.temp_dir)
        paths = sorted(glob.iglob(os.path.join(temp_dir, "*.env")))
        rep_keys = VF_REPLICATE_KEYS if project == "VisualFlocking" else BASE_REPLICATE_KEYS
        # the foraging engine takes FOV and vision range per agent (abm_base_set_agent_geometry): per replicate as well
        # (in the flocking project AGENT_FOV rescales the resolution, vf_sims.py:41-44: a different shape)
        geo_keys = ("AGENT_FOV", "VISION_RANGE") if project == "Base" else ()
        # ... and one set of patch parameters per replicate (abm_base_set_regeneration_params); N_RESOURCES stays a shape
        patch_keys = ("RADIUS_RESOURCE", "MIN_RESOURCE_PER_PATCH", "MAX_RESOURCE_PER_PATCH", "MIN_RESOURCE_QUALITY",
                      "MAX_RESOURCE_QUALITY") if project == "Base" else ()
        geo_keys = geo_keys + patch_keys
        groups = {}
        for p in paths:
            env = params.read_env(p)
            shape = tuple(sorted((k, v) for k, v in env.items()
                                 if k not in rep_keys and k not in geo_keys and k != "SAVE_ROOT_DIR"))
            groups.setdefault(shape, []).append((p, env))
        self.results = []
        for g, members in enumerate(groups.values()):
            envs = [e for _, e in members]
            kw = params.simulation_kwargs(envs[0])
            kw.update(n_replicates=len(envs), seed=None if seed is None else int(seed) + g, root_dir=self.root_dir,
                      save_root_dir=[e.get("SAVE_ROOT_DIR", "abm/data/simulation_data") for e in envs],
                      env_params=[dict(e) for e in envs])
            if project == "VisualFlocking":
                sim = VFSimulation(vf_params=params.VFParams.from_env(envs[0]), **kw)
                per = {name: [float(e.get(k, getattr(sim.vf_params, name))) for e in envs]
                       for k, name in rep_keys.items()}
                sim.engine.set_params(**per)
            elif project == "Base":
                dp = params.DecisionParams.from_env(envs[0])
                sim = Simulation(decision_params=dp, **kw)
                base = dict(dp.engine_kwargs(), agent_consumption=kw["agent_consumption"])
                per = {name: [float(e.get(k, base[name])) for e in envs] for k, name in rep_keys.items()}
                sim.engine.set_params(**per)
                fovs = np.array([float(e.get("AGENT_FOV", kw["agent_fov"])) for e in envs])
                ranges = np.array([float(int(float(e.get("VISION_RANGE", kw["vision_range"])))) for e in envs])   # app.py:53
                pk = [params.simulation_kwargs(e) for e in envs]
                patch = {k: np.array([float(q[k]) for q in pk]) for k in ("patch_radius", "min_resc_perpatch", "max_resc_perpatch",
                                                                           "min_resc_quality", "max_resc_quality")}
                if any((v != v[0]).any() for v in patch.values()):
                    sim.set_replicate_patch_params(**patch)
                if (fovs != fovs[0]).any() or (ranges != ranges[0]).any():
                    ones = np.ones((1, sim.N))
                    sim.engine.set_agent_geometry(agent_fov=fovs[:, None] * ones, vision_range=ranges[:, None] * ones)
            else:
                raise NotImplementedError(f"project {project!r} is out of scope of abm_b200")
            sim.start()
            self.results.append(([p for p, _ in members], sim))
            if not keep_env_files:
                for p, _ in members:
                    os.remove(p)           # a finished protocol's env file is consumed (metarunner.py:225)
        return self.results
