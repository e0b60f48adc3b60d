"""``import abm_b200.compat`` -- the reference's import paths resolve to this package's mirrors, so that a script written
against scioip34/ABM (an experiment file of ``abm/data/metaprotocol/experiments``, a notebook that calls
``vf_supcalc.projection_field``) runs on the B200 engine without editing its imports:

    import abm_b200.compat                                     # once, before the script's own imports
    from abm.metarunner.metarunner import Tunable, Constant, MetaProtocol, TunedPairRestrain
    from abm import app, app_visual_flocking
    from abm.projects.visual_flocking.vf_agent import vf_supcalc

| reference module | resolves to |
|---|---|
| ``abm.metarunner.metarunner`` (metarunner.py:30-254) | ``abm_b200.metarunner`` |
| ``abm.app`` / ``abm.app_visual_flocking`` (app.py:16-70, app_visual_flocking.py:40-108) | ``abm_b200.app`` / ``abm_b200.app_visual_flocking`` |
| ``abm.simulation.sims`` (``Simulation``, sims.py:59) | ``abm_b200.simulation`` |
| ``abm.projects.visual_flocking.vf_simulation.vf_sims`` (``VFSimulation``) | ``abm_b200.simulation`` |
| ``abm.agent.supcalc`` (``projection_field``, ``F_reloc_LR``: the hot-path functions only) | ``abm_b200.supcalc`` |
| ``abm.projects.visual_flocking.vf_agent.vf_supcalc`` | ``abm_b200.vf_supcalc`` |
| ``abm.projects.cooperative_signaling.cs_agent.cs_supcalc`` (``projection_field`` only) | ``abm_b200.cs_supcalc`` |

Refuses to shadow a real ``abm`` package that is already imported (``install(force=True)`` overrides).  Everything outside
the hot path (rendering, replay, InfluxDB, the interactive playground) is NOT provided: importing it raises
ModuleNotFoundError as it would without the reference installed."""
import importlib
import sys
import types

_MAP = {
    "abm.metarunner.metarunner": "abm_b200.metarunner",
    "abm.app": "abm_b200.app",
    "abm.app_visual_flocking": "abm_b200.app_visual_flocking",
    "abm.simulation.sims": "abm_b200.simulation",
    "abm.projects.visual_flocking.vf_simulation.vf_sims": "abm_b200.simulation",
    "abm.agent.supcalc": "abm_b200.supcalc",
    "abm.projects.visual_flocking.vf_agent.vf_supcalc": "abm_b200.vf_supcalc",
    "abm.projects.cooperative_signaling.cs_agent.cs_supcalc": "abm_b200.cs_supcalc",
}


def install(force: bool = False) -> None:
    """Register the aliases in ``sys.modules`` (idempotent)."""
    existing = sys.modules.get("abm")
    if existing is not None and not getattr(existing, "_abm_b200_compat", False) and not force:
        raise ImportError("a real 'abm' package is already imported; abm_b200.compat.install(force=True) shadows it")
    for ref_name, ours in _MAP.items():
        parts = ref_name.split(".")
        for k in range(1, len(parts)):                         # the parent packages, as empty namespace modules
            pkg_name = ".".join(parts[:k])
            pkg = sys.modules.get(pkg_name)
            if pkg is None or (force and not getattr(pkg, "_abm_b200_compat", False)):
                pkg = types.ModuleType(pkg_name)
                pkg.__path__ = []                              # a package with no files: unknown submodules fail loudly
                pkg._abm_b200_compat = True
                sys.modules[pkg_name] = pkg
                if k > 1:
                    setattr(sys.modules[".".join(parts[:k - 1])], parts[k - 1], pkg)
        mod = importlib.import_module(ours)
        sys.modules[ref_name] = mod
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)


install()
