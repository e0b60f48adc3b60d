"""Entry point mirroring abm/app.py:16-70: `.env` -> Simulation(**kwargs).start()."""
from . import params
from .simulation import Simulation


def start(parallel=False, headless=True, agent_behave_param_list=None, env_file=None, **extra):
    envconf = params.read_env(env_file or params.env_path())
    kw = params.simulation_kwargs(envconf)
    kw.update(parallel=parallel, agent_behave_param_list=agent_behave_param_list,
              save_root_dir=envconf.get("SAVE_ROOT_DIR", "abm/data/simulation_data"), env_params=dict(envconf))
    kw.update(extra)
    sim = Simulation(decision_params=params.DecisionParams.from_env(envconf), **kw)
    sim.start()
    return sim


def start_headless():
    return start(headless=True)
