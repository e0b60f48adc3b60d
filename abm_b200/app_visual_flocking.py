"""Entry point mirroring abm/app_visual_flocking.py:40-108."""
from . import params
from .simulation import VFSimulation


def start(parallel=False, headless=True, agent_behave_param_list=None, env_file=None, **extra):
    envconf = params.read_env(env_file or params.env_path())
    if envconf.get("APP_VERSION", "Base") != "VisualFlocking":                    # app_visual_flocking.py:51-54
        raise Exception(".env file was not created for project visual flocking or no APP_VERSION parameter found!")
    kw = params.simulation_kwargs(envconf)
    kw.update(parallel=parallel, agent_behave_param_list=agent_behave_param_list,
              save_root_dir=envconf.get("SAVE_ROOT_DIR", "abm/data/simulation_data"), env_params=dict(envconf))
    kw.update(extra)
    sim = VFSimulation(vf_params=params.VFParams.from_env(envconf), **kw)
    sim.start()
    return sim
