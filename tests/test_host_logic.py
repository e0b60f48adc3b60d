"""CPU tests of the host-side mirror of the reference interface: .env parsing rules
(app.py:27-63), parameter modules, sweep generation (metarunner.py:168-199)."""
import os

import numpy as np
import pytest

from abm_b200 import metarunner as mr
from abm_b200 import params

ENV_TEXT = """
N=10
T=5000
VISUAL_FIELD_RESOLUTION=320
ENV_WIDTH=900
ENV_HEIGHT=900
RADIUS_AGENT=5.5
AGENT_FOV=0.5
N_RESOURCES=3
VISION_RANGE=2000
VISUAL_EXCLUSION=1
AGENT_AGENT_COLLISION=0
DEC_EPSW=2
DEC_TAU=10
MOV_EXP_VEL_MAX=3
APP_VERSION=VisualFlocking
VF_GAM=0.5
VF_ALP0=2
BOUNDARY=infinite
"""


def _write_env(tmp_path):
    p = tmp_path / ".env"
    p.write_text(ENV_TEXT)
    return str(p)


def test_env_parsing_rules(tmp_path):
    e = params.read_env(_write_env(tmp_path))
    kw = params.simulation_kwargs(e)
    assert kw["N"] == 10 and kw["T"] == 5000 and kw["v_field_res"] == 320
    assert kw["agent_radius"] == 5          # int(float("5.5")) (app.py:39)
    assert kw["agent_fov"] == 0.5 and kw["visual_exclusion"] is True and kw["collide_agents"] is False
    assert kw["window_pad"] == 30


def test_vf_gamma_key_quirk(tmp_path):
    """vf_params.py:12 reads VF_GAMMA; experiments set VF_GAM -> GAM stays 0.1."""
    e = params.read_env(_write_env(tmp_path))
    p = params.VFParams.from_env(e)
    assert p.GAM == 0.1 and p.ALP0 == 2.0 and p.BOUNDARY == "infinite"
    d = params.DecisionParams.from_env(e)
    assert d.Eps_w == 2.0 and d.exp_vel_max == 3.0 and d.Tau == 10 and d.S_wu == 0.25


def test_metaprotocol_env_generation(tmp_path):
    e = params.read_env(_write_env(tmp_path))
    mp = mr.MetaProtocol("exp", num_batches=2, default_envconf=e, root_dir=str(tmp_path))
    mp.add_criterion(mr.Tunable("VF_ALP0", 0, 5, 3))
    mp.add_criterion(mr.Tunable("VF_BET0", values_override=[0.5, 1.0]))
    mp.add_criterion(mr.Constant("N", 12))
    assert mp.generate_temp_env_files() == 12
    files = sorted(os.listdir(tmp_path / "abm/data/metaprotocol/temp/exp"))
    assert len(files) == 12 and files[0] == "0_b0.env"
    env0 = params.read_env(str(tmp_path / "abm/data/metaprotocol/temp/exp/5_b1.env"))
    assert env0["N"] == "12" and float(env0["VF_ALP0"]) == 5.0 and float(env0["VF_BET0"]) == 1.0
    assert env0["SAVE_ROOT_DIR"].endswith("exp/batch_1")


def test_tuned_pair_restraint(tmp_path):
    mp = mr.MetaProtocol("e2", default_envconf={}, root_dir=str(tmp_path))
    mp.add_criterion(mr.Tunable("A", values_override=[1, 2, 4]))
    mp.add_criterion(mr.Tunable("B", values_override=[1, 2, 4]))
    mp.add_tuned_pair(mr.TunedPairRestrain("A", "B", 4))
    names, combos = mp.combinations()
    assert sorted(combos) == [(1, 4), (2, 2), (4, 1)]


def test_tunable_requires_values():
    with pytest.raises(Exception):
        mr.Tunable("X")
    assert np.allclose(mr.Tunable("X", 0, 1, 5).get_values(), [0, .25, .5, .75, 1])


def test_simulation_facade_rejects_per_agent_geometry():
    """agent_behave_param_list (sims.py:499-517): per-agent decision parameters, FOV, vision range, radius and
    resolution are supported, a per-agent Tau / pooling is refused loudly, before any engine is created."""
    from abm_b200.simulation import Simulation
    base = dict(S_wu=0, T_w=0.5, Eps_w=0, g_w=0.085, B_w=0, w_max=1, Tau=10, S_uw=0, T_u=0.5, Eps_u=3, g_u=0.085,
                B_u=0, u_max=1, F_N=2, F_R=1, exp_vel_max=3, exp_stop_ratio=0.15, agent_radius=10, v_field_res=1200,
                pooling_time=0, pooling_prob=0, agent_consumption=1, vision_range=2000, agent_fov=0.9)
    plist = [dict(base, v_field_res=1200 if i else 600) for i in range(4)]
    radius, res, radii, res_list = Simulation._check_behave_params(plist, 4, 10, 800, 10)
    assert (radius, res, radii) == (10, 1200, None) and list(res_list) == [600, 1200, 1200, 1200]
    with pytest.raises(NotImplementedError, match="POOLING"):
        Simulation(N=4, T=10, agent_behave_param_list=[dict(base, pooling_time=3)] * 4)
    with pytest.raises(ValueError):
        Simulation(N=5, T=10, agent_behave_param_list=plist)
    with pytest.raises(NotImplementedError, match="Tau"):
        Simulation(N=4, T=10, agent_behave_param_list=[dict(base, Tau=5)] * 4)


@pytest.mark.skipif(not os.path.isdir("/root/reference/abm/metarunner"), reason="reference tree not mounted")
@pytest.mark.parametrize("exp_file", ["figExp3BN50PatchyCollOcc.py", "VFExp4c.py"])
def test_references_own_experiment_files_generate_the_same_env_files(tmp_path, monkeypatch, exp_file):
    """Drop-in check of the sweep language: the reference's OWN experiment files (BASELINE configs[2] and configs[3] are
    these two) are executed unchanged, once against the reference's metarunner (a temporary copy, so that nothing is
    written into the reference tree; `abm.app` stubbed, `run_protocols` a no-op) and once against abm_b200.metarunner
    under the same module name: the generated env files -- names, keys, values -- must be identical."""
    import importlib.util
    import shutil
    import sys
    import types
    src = "/root/reference"
    exp_path = os.path.join(src, "abm/data/metaprotocol/experiments", exp_file)
    exp_name = "dropin_exp"
    monkeypatch.setenv("EXPERIMENT_NAME", exp_name)
    code = compile(open(exp_path).read(), exp_path, "exec")

    def run_with(module, root):
        saved = {k: sys.modules.get(k) for k in ("abm", "abm.app", "abm.metarunner", "abm.metarunner.metarunner")}
        try:
            pkg = types.ModuleType("abm"); pkg.__path__ = []
            sys.modules["abm"] = pkg
            sys.modules["abm.app"] = types.ModuleType("abm.app"); pkg.app = sys.modules["abm.app"]
            sys.modules["abm.metarunner"] = types.ModuleType("abm.metarunner")
            mod = module()
            sys.modules["abm.metarunner.metarunner"] = mod
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                exec(code, {"__name__": "__exp__"})
        finally:
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
        d = os.path.join(root, "abm/data/metaprotocol/temp", exp_name)
        return {f: params.read_env(os.path.join(d, f)) for f in sorted(os.listdir(d))}

    # ---- the reference's metarunner, from a temporary copy (root_abm_dir = three levels above the module file) ----
    ref_root = tmp_path / "ref"
    os.makedirs(ref_root / "abm/metarunner")
    shutil.copyfile(os.path.join(src, "abm/metarunner/metarunner.py"), ref_root / "abm/metarunner/metarunner.py")
    shutil.copyfile(os.path.join(src, ".env"), ref_root / f"{exp_name}.env")

    def load_reference():
        spec = importlib.util.spec_from_file_location("abm.metarunner.metarunner", ref_root / "abm/metarunner/metarunner.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.MetaProtocol.run_protocols = lambda self, *a, **k: None
        return mod

    want = run_with(load_reference, str(ref_root))

    # ---- abm_b200.metarunner under the same name, same default .env, its own root ----
    our_root = tmp_path / "ours"
    os.makedirs(our_root)
    shutil.copyfile(os.path.join(src, ".env"), our_root / f"{exp_name}.env")

    def load_ours():
        mod = types.ModuleType("abm.metarunner.metarunner")
        for name in ("Tunable", "Constant", "TunedPairRestrain"):
            setattr(mod, name, getattr(mr, name))

        class MetaProtocol(mr.MetaProtocol):
            def __init__(self, *a, **k):
                super().__init__(*a, root_dir=str(our_root), **k)

            def run_protocols(self, *a, **k):
                return None
        mod.MetaProtocol = MetaProtocol
        return mod

    got = run_with(load_ours, str(our_root))
    assert list(got) == list(want) and len(want) >= 12
    for f in want:
        assert got[f] == want[f], f


@pytest.mark.skipif(not os.path.isdir("/root/reference/abm/metarunner"), reason="reference tree not mounted")
@pytest.mark.parametrize("app_name", ["app", "app_visual_flocking"])
def test_env_to_kwargs_equals_the_references_own_app(monkeypatch, app_name):
    """`.env` -> constructor kwargs: the reference's own `start()` (abm/app.py:16-70, app_visual_flocking.py:40-108) is
    run on its own root `.env` with the simulation class replaced by a recorder of its keyword arguments; every kwarg it
    passes must equal what abm_b200.params.simulation_kwargs makes of the same file (types included)."""
    import importlib
    from oracle import ref_shim
    import sys
    import types
    ref_shim.install()
    monkeypatch.setenv("EXPERIMENT_NAME", "")
    # the interactive playground classes the apps import at module level pull in GUI widget packages: not on this path
    for name, attrs in (("abm.simulation.isims", ("PlaygroundSimulation",)),
                        ("abm.projects.visual_flocking.vf_simulation.vf_isims", ("VFPlaygroundSimulation",)),
                        ("abm.projects.visual_flocking.vf_contrib.vf_playgroundtool", ("setup_visflock_playground",)),
                        ("abm.contrib.playgroundtool", ())):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a_ in attrs:
                setattr(m, a_, object)
            monkeypatch.setitem(sys.modules, name, m)
    mod = importlib.import_module(f"abm.{app_name}")
    captured = {}

    class Recorder:
        def __init__(self, **kw):
            captured.update(kw)

        def start(self):
            pass
    cls_name = "Simulation" if app_name == "app" else "VFSimulation"
    monkeypatch.setattr(mod, cls_name, Recorder)
    mod.start(parallel=False, headless=False)
    assert len(captured) > 30
    ours = params.simulation_kwargs(params.read_env("/root/reference/.env"))
    for k, v in captured.items():
        if k in ("parallel", "agent_behave_param_list"):
            continue
        assert k in ours, f"the reference passes {k!r}, simulation_kwargs does not"
        assert ours[k] == v and type(ours[k]) is type(v), (k, ours[k], v)


@pytest.mark.skipif(not os.path.isdir("/root/reference/abm/metarunner"), reason="reference tree not mounted")
def test_parameter_sets_equal_the_references_param_modules(monkeypatch):
    """DecisionParams / VFParams .from_env against the reference's own parameter modules (contrib/decision_params.py,
    movement_params.py, vf_contrib/vf_params.py), which read the same root `.env` at import: every value the hot path
    uses must agree -- including the VF_GAMMA key quirk (the file sets VF_GAM, the module reads VF_GAMMA)."""
    import importlib
    from oracle import ref_shim
    ref_shim.install()
    monkeypatch.setenv("EXPERIMENT_NAME", "")
    env = params.read_env("/root/reference/.env")
    dp = params.DecisionParams.from_env(env)
    dec = importlib.reload(importlib.import_module("abm.contrib.decision_params"))
    mov = importlib.reload(importlib.import_module("abm.contrib.movement_params"))
    for k in ("T_w", "Eps_w", "g_w", "B_w", "w_max", "T_u", "Eps_u", "g_u", "B_u", "u_max", "S_wu", "S_uw", "Tau", "F_N", "F_R"):
        assert getattr(dp, k) == getattr(dec, k), k
    for k in ("exp_vel_max", "exp_theta_min", "exp_theta_max", "reloc_theta_max", "exp_stop_ratio"):
        assert getattr(dp, k) == getattr(mov, k), k
    vp = params.VFParams.from_env(env)
    vfp = importlib.reload(importlib.import_module("abm.projects.visual_flocking.vf_contrib.vf_params"))
    for k in ("GAM", "V0", "ALP0", "ALP1", "ALP2", "BET0", "BET1", "BET2", "BOUNDARY", "LIMIT_MOVEMENT", "MAX_VEL", "MAX_TH"):
        assert getattr(vp, k) == getattr(vfp, k), k


def test_compat_module_resolves_the_references_import_paths(tmp_path):
    """`import abm_b200.compat`: the reference's import lines work unchanged and resolve to this package's mirrors; a
    module outside the hot path fails like a missing package.  In a fresh interpreter (sys.modules is process-wide)."""
    import subprocess
    import sys
    import textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {root!r})
        import abm_b200.compat
        from abm.metarunner.metarunner import Tunable, Constant, MetaProtocol, TunedPairRestrain
        from abm import app, app_visual_flocking
        from abm.simulation.sims import Simulation
        from abm.projects.visual_flocking.vf_simulation.vf_sims import VFSimulation
        from abm.projects.visual_flocking.vf_agent import vf_supcalc
        from abm.projects.cooperative_signaling.cs_agent import cs_supcalc
        from abm.agent import supcalc
        import abm_b200.metarunner, abm_b200.simulation, abm_b200.vf_supcalc
        assert MetaProtocol is abm_b200.metarunner.MetaProtocol and Simulation is abm_b200.simulation.Simulation
        assert VFSimulation is abm_b200.simulation.VFSimulation and vf_supcalc is abm_b200.vf_supcalc
        assert callable(supcalc.projection_field) and callable(cs_supcalc.projection_field) and callable(app.start)
        mp = MetaProtocol("e", default_envconf={{"N": "3"}}, root_dir={str(tmp_path)!r})
        mp.add_criterion(Tunable("X", values_override=[1, 2])); mp.add_criterion(Constant("Y", 5))
        assert mp.generate_temp_env_files() == 2
        try:
            import abm.replay.replay
            raise SystemExit("a module outside the hot path must not resolve")
        except ModuleNotFoundError:
            pass
        abm_b200.compat.install()                      # idempotent
        print("COMPAT_OK")
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert "COMPAT_OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.skipif(not os.path.isdir("/root/reference/abm/metarunner"), reason="reference tree not mounted")
@pytest.mark.parametrize("variant", ["batches_and_bools", "tuned_pair", "quadratic_pair", "no_experiment_name", "linspace"])
def test_env_generation_equals_the_references_metarunner(tmp_path, monkeypatch, variant):
    """The sweep language against the reference's own classes (a temporary copy of abm/metarunner/metarunner.py, `abm.app`
    stubbed): the same criteria given to both MetaProtocols -- several batches, boolean values, a product restraint, a
    quadratic restraint, no experiment name, a linspace Tunable -- generate the same env files (names, keys, values)."""
    import importlib.util
    import shutil
    import sys
    import types
    exp_name = None if variant == "no_experiment_name" else "gen_exp"
    monkeypatch.setenv("EXPERIMENT_NAME", exp_name or "")
    ref_root = tmp_path / "ref"
    os.makedirs(ref_root / "abm/metarunner")
    shutil.copyfile("/root/reference/abm/metarunner/metarunner.py", ref_root / "abm/metarunner/metarunner.py")
    shutil.copyfile("/root/reference/.env", ref_root / f"{exp_name or ''}.env")
    our_root = tmp_path / "ours"
    os.makedirs(our_root)
    shutil.copyfile("/root/reference/.env", our_root / f"{exp_name or ''}.env")
    saved = {k: sys.modules.get(k) for k in ("abm", "abm.app")}
    try:
        pkg = types.ModuleType("abm"); pkg.__path__ = []
        sys.modules["abm"] = pkg
        sys.modules["abm.app"] = types.ModuleType("abm.app"); pkg.app = sys.modules["abm.app"]
        spec = importlib.util.spec_from_file_location("ref_metarunner_copy", ref_root / "abm/metarunner/metarunner.py")
        ref = importlib.util.module_from_spec(spec)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    def build(m, **kw):
        mp = m.MetaProtocol(experiment_name=exp_name, num_batches=3 if variant == "batches_and_bools" else 1,
                            parallel=exp_name is not None, description="d", headless=True, **kw)
        if variant == "batches_and_bools":
            mp.add_criterion(m.Tunable("VISUAL_EXCLUSION", values_override=[True, False]))
            mp.add_criterion(m.Constant("N", 7))
        elif variant == "tuned_pair":
            mp.add_criterion(m.Tunable("N_RESOURCES", values_override=[1, 2, 3, 4, 6]))
            mp.add_criterion(m.Tunable("MIN_RESOURCE_PER_PATCH", values_override=[100, 200, 300, 400, 600]))
            mp.add_tuned_pair(m.TunedPairRestrain("N_RESOURCES", "MIN_RESOURCE_PER_PATCH", 1200))
        elif variant == "quadratic_pair":
            mp.add_criterion(m.Tunable("N_RESOURCES", values_override=[1, 4, 16]))
            mp.add_criterion(m.Tunable("RADIUS_RESOURCE", values_override=[10.0, 20.0, 40.0]))
            mp.add_quadratic_tuned_pair(m.TunedPairRestrain("N_RESOURCES", "RADIUS_RESOURCE", 1600))
        elif variant == "linspace":
            mp.add_criterion(m.Tunable("DEC_EPSW", 0, 5, 6))
            mp.add_criterion(m.Tunable("DEC_EPSU", 0.5, 1.5, 3))
        else:
            mp.add_criterion(m.Tunable("DEC_EPSW", values_override=[0, 1]))
        with contextlib.redirect_stdout(io.StringIO()):
            mp.generate_temp_env_files()
        return mp

    build(ref)
    build(mr, root_dir=str(our_root))
    sub = os.path.join("abm/data/metaprotocol/temp", exp_name) if exp_name else "abm/data/metaprotocol/temp"
    want = {f: params.read_env(os.path.join(ref_root, sub, f)) for f in sorted(os.listdir(os.path.join(ref_root, sub)))}
    got = {f: params.read_env(os.path.join(our_root, sub, f)) for f in sorted(os.listdir(os.path.join(our_root, sub)))}
    assert list(got) == list(want) and len(want) >= 2
    for f in want:
        assert got[f] == want[f], (f, {k: (got[f].get(k), want[f].get(k)) for k in set(got[f]) | set(want[f]) if got[f].get(k) != want[f].get(k)})


def test_read_env_equals_dotenv(tmp_path):
    """params.read_env against python-dotenv's dotenv_values (what the reference reads its `.env` files with) on the
    reference's own root `.env` (if at hand) and on a file with comments, blank lines, quotes, inline comments and `=`
    inside values."""
    dotenv = pytest.importorskip("dotenv")
    tricky = tmp_path / "t.env"
    tricky.write_text("\n".join(["# a comment", "", "N=10", 'NAME="quoted value"', "SINGLE='single'", "SPACED = 5", "WITH_EQ=a=b",
                                 "INLINE=3 # trailing comment", "EMPTY=", "SAVE_ROOT_DIR=abm/data/simulation_data/x/batch_0", ""]))
    paths = [str(tricky)]
    if os.path.isfile("/root/reference/.env"):
        paths.append("/root/reference/.env")
    for p in paths:
        assert params.read_env(p) == dict(dotenv.dotenv_values(p)), p
