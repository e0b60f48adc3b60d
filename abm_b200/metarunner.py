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
    """A parameter range to loop through (metarunner.py:59-96)."""

    def __init__(self, var_name, min_v=None, max_v=None, num_data_points=None, values_override=None):
        if min_v is None and values_override is None:
            raise Exception("Neither value borders nor override values have been given to create Tunable!")
        elif min_v is not None and values_override is not None:
            warnings.warn("Both value borders and override values are defined when creating Tunable, using override"
                          "values as default!")
        self.name = var_name
        if values_override is None:
            self.min_val, self.max_val, self.n_data, self.generated = min_v, max_v, num_data_points, True
            self.values = np.linspace(min_v, max_v, num=num_data_points, endpoint=True)
        else:
            self.min_val, self.max_val = min(values_override), max(values_override)
            self.n_data, self.generated, self.values = len(values_override), False, values_override

    def print(self):
        print(f"Tunable: {self.name} = {self.min_val}  -  -  -n={self.n_data}-  -  -  {self.max_val}")

    def get_values(self):
        return self.values


class Constant:
    """A constant parameter value (metarunner.py:30-43)."""

    def __init__(self, var_name, constant):
        self.tunable = Tunable(var_name, values_override=[constant])
        self.name = self.tunable.name

    def get_values(self):
        return self.tunable.values

    def print(self):
        print(f"Constant {self.tunable.name} = {self.tunable.values[0]}")


class TunedPairRestrain:
    """Parameter pair restrained by its product (metarunner.py:45-57)."""

    def __init__(self, var_name1, var_name2, restrained_product):
        self.var1, self.var2, self.product_restrain = var_name1, var_name2, restrained_product

    def get_vars(self):
        return [self.var1, self.var2]

    def print(self):
        print(f"Product of {self.var1} and {self.var2} should be {self.product_restrain}")


class MetaProtocol:
    """metarunner.py:98-254, batched."""

    def __init__(self, experiment_name=None, num_batches=1, parallel=False, description=None, headless=True,
                 default_envconf=None, root_dir=None):
        self.root_dir = root_dir or os.getcwd()
        if default_envconf is None:
            p = params.env_path(self.root_dir)
            default_envconf = params.read_env(p) if os.path.isfile(p) else {}
        self.default_envconf = dict(default_envconf)
        self.tunables, self.tuned_pairs, self.q_tuned_pairs = [], [], []
        self.experiment_name, self.num_batches, self.description = experiment_name, num_batches, description
        self.headless = headless
        if experiment_name is None and parallel:
            raise Exception("Can't run multiple experiments parallely without experiment name!")
        self.parallel_run = parallel
        sub = f"abm/data/metaprotocol/temp/{experiment_name}" if experiment_name else "abm/data/metaprotocol/temp"
        self.temp_dir = sub
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
            warnings.warn("Temprary directory for env files is not empty and will be overwritten")
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

    def run_protocols(self, project="Base", seed=None, keep_env_files=False):
        """All remaining protocols of the temp folder (metarunner.py:240-254), grouped into
        replicate batches.  Returns the list of (env_paths, simulation) per batch; every
        simulation holds the final state of its replicates."""
        temp_dir = os.path.join(self.root_dir, self.temp_dir)
        paths = sorted(glob.iglob(os.path.join(temp_dir, "*.env")))
        rep_keys = VF_REPLICATE_KEYS if project == "VisualFlocking" else BASE_REPLICATE_KEYS
        groups = {}
        for p in paths:
            env = params.read_env(p)
            shape = tuple(sorted((k, v) for k, v in env.items() if k not in rep_keys and k != "SAVE_ROOT_DIR"))
            groups.setdefault(shape, []).append((p, env))
        self.results = []
        for members in groups.values():
            envs = [e for _, e in members]
            kw = params.simulation_kwargs(envs[0])
            if project == "VisualFlocking":
                sim = VFSimulation(vf_params=params.VFParams.from_env(envs[0]), n_replicates=len(envs), seed=seed, **kw)
                per = {name: [float(e.get(k, getattr(sim.vf_params, name))) for e in envs]
                       for k, name in rep_keys.items()}
                sim.engine.set_params(**per)
            elif project == "Base":
                dp = params.DecisionParams.from_env(envs[0])
                sim = Simulation(decision_params=dp, n_replicates=len(envs), seed=seed, **kw)
                base = dict(dp.engine_kwargs(), agent_consumption=kw["agent_consumption"])
                per = {name: [float(e.get(k, base[name])) for e in envs] for k, name in rep_keys.items()}
                sim.engine.set_params(**per)
            else:
                raise NotImplementedError(f"project {project!r} is out of scope of abm_b200")
            sim.start()
            self.results.append(([p for p, _ in members], sim))
            if not keep_env_files:
                for p, _ in members:
                    os.remove(p)           # a finished protocol's env file is consumed (metarunner.py:225)
        return self.results
