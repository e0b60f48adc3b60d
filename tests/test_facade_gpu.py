"""GPU tests of the class-level drop-in interface (Simulation / VFSimulation / MetaProtocol)."""
import os

import numpy as np
import pytest

from oracle import restate as rs

pytestmark = pytest.mark.gpu


def test_vfsimulation_programmatic_stepping(built_lib):
    """The stepping API used by vf_utils/attraction_repulsion_map.py:27-66: prepare_start,
    mutate agents, step_sim, read dv / dphi / ablob ... -- checked against the oracle."""
    from abm_b200.simulation import VFSimulation
    from abm_b200.params import VFParams
    # vf_sims.py:184-228: the list is accepted and every VFAgent is still built with behave_params=None
    sim = VFSimulation(N=2, T=10, v_field_res=1200, width=500, height=500, agent_radius=10, agent_fov=1.0,
                       vf_params=VFParams(ALP0=1.0, BET0=1.0, ALP1=0.09, BET1=0.09), seed=3,
                       agent_behave_param_list=[{"Eps_w": 5.0, "agent_fov": 0.25}] * 2)
    sim.prepare_start()
    a0, a1 = sim.agents
    a0.position = [250.0, 250.0]; a0.orientation = 0.0; a0.velocity = 1.0
    a1.position = [330.0, 210.0]; a1.orientation = 2.0; a1.velocity = 1.0
    a0.verbose_supcalc = True
    sim.step_sim()
    cfg = rs.VFConfig(R=1200, width=500, height=500)
    ref = rs.vf_step_frozen([250.0, 330.0], [250.0, 210.0], [0.0, 2.0], [1.0, 1.0], 10.0, cfg)
    np.testing.assert_allclose([a0.dv, a0.dphi, a0.ablob, a0.aedge, a0.bblob, a0.bedge], ref["terms"][0], rtol=1e-9)
    np.testing.assert_allclose(a0.position, [ref["x"][0], ref["y"][0]], rtol=1e-5)
    assert np.isclose(a1.orientation, ref["theta"][1], rtol=1e-5)
    assert np.array_equal(a0.soc_v_field > 0, ref["rows"][0][::-1])
    assert sim.t == 1


def test_vfsimulation_rescales_resolution_with_fov(built_lib):
    from abm_b200.simulation import VFSimulation
    sim = VFSimulation(N=4, T=3, v_field_res=1200, width=400, height=400, agent_fov=0.5, seed=1)
    assert sim.v_field_res == 2400                # vf_sims.py:41-44
    sim.start()
    assert sim.t == 3 and len(sim.agents[0].soc_v_field) == 2400


def test_simulation_facade_runs(built_lib):
    from abm_b200.simulation import Simulation
    from abm_b200.params import DecisionParams
    sim = Simulation(N=10, T=50, v_field_res=1200, width=500, height=500, N_resc=3, patch_radius=30,
                     min_resc_perpatch=100, max_resc_perpatch=-1, min_resc_quality=0.25, max_resc_quality=-1,
                     vision_range=2000, agent_fov=1.0, visual_exclusion=True, teleport_exploit=False,
                     allow_border_patch_overlap=True, collide_agents=False, n_replicates=4, seed=5,
                     decision_params=DecisionParams(Eps_w=2.0, Eps_u=1.0, exp_vel_max=3.0))
    sim.start()
    assert sim.t == 50 and len(sim.agents) == 10 and len(sim.rescources) == 3
    a = sim.agents[0]
    assert a.mode in ("explore", "exploit", "relocate", "collide") and a.get_mode() in ("explore", "exploit", "relocate")
    assert np.isfinite(a.position).all() and len(a.soc_v_field) == 1200
    assert sim.rescources[0].radius == 30.0 and sim.rescources[0].resc_left <= 101


def test_apps_start_from_the_references_own_env_file(built_lib, tmp_path):
    """abm/app.py:16-70 and app_visual_flocking.py:40-108 mirrored: the reference's OWN root `.env` (copied unmodified to
    oracle/_ref) -> kwargs -> Simulation / VFSimulation .start(), with only the run length and the output folder
    overridden; the runs leave the reference's output layout."""
    import glob
    from oracle import ref_shim
    env_file = os.path.join(ref_shim._reference_root(), ".env")
    if not os.path.isfile(env_file):
        pytest.skip("the reference's .env is not at hand")
    from abm_b200 import app, app_visual_flocking, params
    env = params.read_env(env_file)
    over = dict(T=40, save_root_dir=str(tmp_path / "out"), n_replicates=2, seed=3)   # (the file sets USE_IFDB_LOGGING=1 AND
    # USE_RAM_LOGGING=1: RAM logging wins, sims.py:206-208)
    vsim = app_visual_flocking.start(env_file=env_file, **over)                     # the file says APP_VERSION=VisualFlocking
    fov = float(env["AGENT_FOV"])
    assert vsim.t == 40 and vsim.N == int(env["N"]) and vsim.v_field_res == int(int(env["VISUAL_FIELD_RESOLUTION"]) / fov)
    assert vsim.WIDTH == float(env["ENV_WIDTH"]) and np.isfinite(vsim.engine.get_state()["x"]).all()
    bsim = app.start(env_file=env_file, **over)                                     # the foraging app reads the same file
    assert bsim.t == 40 and bsim.N == int(env["N"]) and bsim.N_resc == int(env["N_RESOURCES"])
    assert np.isfinite(bsim.engine.get_agents()["x"]).all()
    if int(env.get("USE_RAM_LOGGING", "0")) and int(env.get("SAVE_CSV_FILES", "0")):
        assert len(glob.glob(str(tmp_path / "out" / "*" / "ag_posx.zarr"))) + \
               len(glob.glob(str(tmp_path / "out" / "*" / "*" / "ag_posx.zarr"))) >= 2


@pytest.mark.parametrize("boundary,limit", [("walls", True), ("infinite", False)])
def test_vfsimulation_kwargs_reach_the_kernels(built_lib, boundary, limit):
    """VFSimulation's constructor kwargs and the vf_params values (vf_sims.py:31-68, vf_params.py:12-23), all away from their
    defaults -- a non-square arena, a limited FOV (which rescales the resolution, vf_sims.py:41-44), boundary condition,
    movement limits, every flocking parameter -- through the façade: four steps of `step_sim` in lockstep with the oracle
    configured from the same kwargs."""
    from abm_b200.simulation import VFSimulation
    from abm_b200.params import VFParams
    vp = VFParams(GAM=0.2, V0=1.5, ALP0=0.7, ALP1=0.05, BET0=1.4, BET1=0.12, BOUNDARY=boundary, LIMIT_MOVEMENT=limit,
                  MAX_VEL=2.0, MAX_TH=0.05)
    sim = VFSimulation(N=40, T=10, v_field_res=900, width=420, height=300, window_pad=30, agent_radius=7, agent_fov=0.75,
                       vf_params=vp, n_replicates=3, seed=12)
    sim.prepare_start()
    R = int(900 * (1 / 0.75))
    assert sim.v_field_res == R
    cfg = rs.VFConfig(R=R, fov=(-0.75 * np.pi, 0.75 * np.pi), boundary=boundary, width=420, height=300, window_pad=30,
                      GAM=0.2, V0=1.5, ALP0=0.7, ALP1=0.05, BET0=1.4, BET1=0.12, limit_movement=limit, max_vel=2.0, max_th=0.05)
    for step in range(4):
        before = [(ag.position.copy(), ag.orientation, ag.velocity) for b in range(3) for ag in sim.replicate_agents(b)]
        sim.step_sim()
        st = sim.engine.get_state()
        fields = sim.engine.fields()
        for b in range(3):
            rows = before[b * 40:(b + 1) * 40]
            x = np.array([r[0][0] for r in rows]); y = np.array([r[0][1] for r in rows])
            th = np.array([r[1] for r in rows]); v = np.array([r[2] for r in rows])
            ref = rs.vf_step_frozen(x, y, th, v, 7.0, cfg)
            assert np.array_equal(fields[b], ref["rows"][:, ::-1]), (step, b)
            for k in ("x", "y", "theta", "vel"):
                np.testing.assert_allclose(st[k][b], ref[k], rtol=1e-5, atol=1e-5, err_msg=f"{k} step {step}")
    assert sim.t == 4


def test_simulation_kwargs_reach_the_kernels(built_lib):
    """The reference's constructor kwargs (sims.py:60-68) and decision / movement parameters, all away from their defaults,
    through `Simulation` into the engine: five steps of the façade's engine in lockstep with the oracle configured from the
    SAME kwargs by the reference's rules (agent_fov as a fraction of pi, vision range, exclusion flags, teleport, ghost
    mode, radius, resolution, consumption) -- a wrong or dropped kwarg shows as a field or state difference."""
    from abm_b200.simulation import Simulation
    from abm_b200.params import DecisionParams
    from oracle import restate_base as rb
    dp = DecisionParams(T_w=0.4, Eps_w=2.5, g_w=0.07, B_w=0.05, w_max=0.9, T_u=0.45, Eps_u=1.5, g_u=0.09, B_u=0.02,
                        u_max=0.95, S_wu=0.2, S_uw=0.03, Tau=7, F_N=1.5, F_R=0.8, exp_vel_max=2.5, exp_theta_min=-0.4,
                        exp_theta_max=0.2, reloc_theta_max=1.1, exp_stop_ratio=0.12)
    kw = dict(N=24, T=100, v_field_res=900, width=260, height=220, window_pad=30, agent_radius=8, N_resc=3, patch_radius=28,
              min_resc_perpatch=300, max_resc_perpatch=-1, min_resc_quality=0.4, max_resc_quality=-1,
              regenerate_patches=False, agent_consumption=0.5, teleport_exploit=True, vision_range=140, agent_fov=0.6,
              visual_exclusion=True, ghost_mode=False, patchwise_exclusion=False, collide_agents=True)
    sim = Simulation(n_replicates=2, seed=9, decision_params=dp, **kw)
    sim.create_agents(); sim.create_resources()
    eng = sim.engine
    cfg = rb.BaseConfig(R=900, fov=(-0.6 * np.pi, 0.6 * np.pi), width=260, height=220, window_pad=30, vision_range=140.0,
                        visual_exclusion=True, patchwise_exclusion=False, teleport_exploit=True, agent_consumption=0.5,
                        **{k: getattr(dp, k) for k in ("T_w", "Eps_w", "g_w", "B_w", "w_max", "T_u", "Eps_u", "g_u", "B_u",
                                                       "u_max", "S_wu", "S_uw", "Tau", "F_N", "F_R", "exp_vel_max",
                                                       "exp_theta_min", "exp_theta_max", "reloc_theta_max", "exp_stop_ratio")})
    rng = np.random.default_rng(4)
    a = eng.get_agents()
    a["u"][:, :8] = 0.9                                   # some agents about to exploit (the run is only five steps long)
    p = eng.get_patches()
    a["x"][:, :8] = p["x"][:, :1] + 28 - 8 + rng.integers(-10, 10, (2, 8)); a["y"][:, :8] = p["y"][:, :1] + 28 - 8
    eng.set_agents(x=a["x"], y=a["y"], theta=a["theta"], u=a["u"])
    for step in range(5):
        a0, p0 = eng.get_agents(), eng.get_patches()
        dth = np.asarray(rng.uniform(-0.4, 0.2, (2, kw["N"])), np.float32)
        eng.step(1, inject_dtheta=dth)
        got, fields = eng.get_agents(), eng.fields()
        for b in range(2):
            st = dict(x=a0["x"][b].astype(float), y=a0["y"][b].astype(float), theta=a0["theta"][b].astype(float),
                      vel=a0["vel"][b].astype(float), w=a0["w"][b].astype(float), u=a0["u"][b].astype(float),
                      collected=a0["collected"][b].astype(float), collected_before=a0["collected_before"][b].astype(float),
                      env_status=a0["env_status"][b].copy(), override=a0["override_mode"][b].copy(), mode=a0["mode"][b].copy(),
                      patch_id=a0["patch_id"][b].copy(), radius=8.0,
                      novelty=((a0["novelty"][b][:, None] >> np.arange(cfg.Tau)) & 1).astype(float))
            pa = {k: np.array(p0[k][b], dtype=float if k != "id" else int) for k in p0}
            collided = rb.base_collision_phase(st, cfg, False)
            rb.base_patch_phase(st, pa, cfg, collided=set(collided))
            ref = rb.base_step_frozen(st, cfg, dth[b].astype(np.float64))
            assert np.array_equal(fields[b], ref["fields"]), (step, b)
            for k, g in dict(x="x", y="y", theta="theta", vel="vel", w="w", u="u").items():
                np.testing.assert_allclose(got[g][b], ref[k], rtol=1e-5, atol=1e-5, err_msg=f"{k} step {step}")
            assert np.array_equal(got["mode"][b], np.asarray(ref["mode"]).astype(int))
            np.testing.assert_allclose(got["collected"][b], st["collected"], rtol=1e-5, atol=1e-6)
    assert got["collected"].sum() > 0                      # somebody exploited a patch at consumption 0.5 / quality 0.4


def _behave_template(**over):
    """contrib/evolution.py:1-26."""
    d = dict(S_wu=0, T_w=0.5, Eps_w=0, g_w=0.085, B_w=0, w_max=1, Tau=10, S_uw=0, T_u=0.5, Eps_u=3, g_u=0.085, B_u=0,
             u_max=1, F_N=2, F_R=1, exp_vel_max=3, exp_stop_ratio=0.15, agent_radius=10, v_field_res=1200,
             pooling_time=0, pooling_prob=0, agent_consumption=1, vision_range=2000, agent_fov=0.9,
             evo_summary_path=None)
    d.update(over)
    return d


def test_simulation_facade_heterogeneous_agents(built_lib):
    """agent_behave_param_list (sims.py:170-173, 499-517): the dictionaries' geometry replaces the constructor's,
    their decision entries become one parameter set per agent.  Agents whose Eps_u is 0 never reach the
    exploitation threshold, so only the others can collect."""
    from abm_b200.simulation import Simulation
    N = 12
    plist = [_behave_template(Eps_u=0.0 if i % 2 else 3.0, exp_vel_max=2.0 + 0.1 * i, agent_fov=0.5 + 0.04 * i,
                              vision_range=100 + 50 * i) for i in range(N)]
    sim = Simulation(N=N, T=400, v_field_res=320, width=300, height=300, N_resc=3, patch_radius=40,
                     min_resc_perpatch=500, max_resc_perpatch=-1, min_resc_quality=0.25, max_resc_quality=-1,
                     vision_range=150, agent_fov=0.5, visual_exclusion=True, teleport_exploit=False,
                     allow_border_patch_overlap=True, collide_agents=False, n_replicates=8, seed=11,
                     agent_behave_param_list=plist)
    assert sim.heterogen_agents and sim.engine.R == 1200            # v_field_res / fov / range from the list
    sim.start()
    a = sim.engine.get_agents()
    assert np.isfinite(a["x"]).all()
    assert (a["collected"][:, 1::2] == 0).all()                     # Eps_u = 0: u stays at its baseline
    assert (a["collected"][:, 0::2] > 0).any()
    explore = a["mode"] == 0
    vmax = np.array([p["exp_vel_max"] for p in plist])
    v = np.abs(a["vel"])
    assert np.allclose(v[explore], np.broadcast_to(vmax, v.shape)[explore], rtol=1e-6)   # own max_exp_vel


def test_simulation_facade_per_agent_resolution_and_radius(built_lib):
    """v_field_res and agent_radius of agent_behave_param_list (sims.py:502, 507): the engine's resolution is the
    largest, every agent view has its own field length and radius."""
    from abm_b200.simulation import Simulation
    N = 9
    plist = [_behave_template(v_field_res=[1200, 640, 333][i % 3], agent_radius=[10, 6, 14][i % 3], Eps_w=2.0,
                              agent_fov=1.0, vision_range=2000) for i in range(N)]
    sim = Simulation(N=N, T=60, v_field_res=320, width=300, height=300, N_resc=2, patch_radius=40,
                     min_resc_perpatch=500, max_resc_perpatch=-1, min_resc_quality=0.25, max_resc_quality=-1,
                     visual_exclusion=True, teleport_exploit=False, allow_border_patch_overlap=True,
                     collide_agents=True, n_replicates=4, seed=5, agent_behave_param_list=plist)
    assert sim.engine.R == 1200 and list(sim.agent_res_list[:3]) == [1200, 640, 333]
    sim.start()
    a = sim.engine.get_agents()
    assert np.isfinite(a["x"]).all() and np.isfinite(a["w"]).all()
    assert [len(ag.soc_v_field) for ag in sim.agents[:3]] == [1200, 640, 333]
    assert [ag.radius for ag in sim.agents[:3]] == [10.0, 6.0, 14.0]
    f = sim.engine.fields()
    for i in range(N):
        assert not f[:, i, plist[i]["v_field_res"]:].any()


def test_metaprotocol_runs_sweep_as_one_batch(built_lib, tmp_path):
    """the generated env files run as ONE replicate batch, and every replicate leaves what the reference's
    sequential run of its env file leaves on disk (metarunner.py:168-254, ifdb_params.py:34-40, ifdb.py:435-535,
    env_saver.py:15-28): <root>/abm/data/simulation_data/<exp>/batch_<nb>/<timestamp>/{ag_*.zarr, env_params.json}."""
    import glob
    import json
    import os
    from abm_b200 import metarunner as mr
    from abm_b200.recorder import read_zarr_v2
    env = dict(N="12", T="20", VISUAL_FIELD_RESOLUTION="1200", ENV_WIDTH="400", ENV_HEIGHT="400", RADIUS_AGENT="10",
               AGENT_FOV="1", APP_VERSION="VisualFlocking", BOUNDARY="walls", VF_ALP1="0.09", VF_BET1="0.09",
               USE_RAM_LOGGING="1", SAVE_CSV_FILES="1", USE_ZARR_FORMAT="1")
    mp = mr.MetaProtocol("sweep", num_batches=2, default_envconf=env, root_dir=str(tmp_path),
                         description="ALP0 x BET0 sweep")
    mp.add_criterion(mr.Tunable("VF_ALP0", values_override=[0, 1, 3]))
    mp.add_criterion(mr.Tunable("VF_BET0", values_override=[0, 2]))
    assert mp.generate_temp_env_files() == 12
    res = mp.run_protocols(project="VisualFlocking", seed=7)
    assert len(res) == 1                              # one replicate batch for the whole sweep
    paths, sim = res[0]
    assert len(paths) == 12 and sim.B == 12 and sim.t == 20
    assert sim.engine.counters()["launches"] == 20
    st = sim.engine.get_state()
    assert np.isfinite(st["x"]).all()
    # ---- on-disk layout ----
    exp_dir = tmp_path / "abm" / "data" / "simulation_data" / "sweep"
    assert (exp_dir / "README.txt").read_text() == "ALP0 x BET0 sweep"
    assert len(sim.saved_dirs) == 12
    zero_gain = []
    for nb in (0, 1):
        runs = sorted(glob.glob(str(exp_dir / f"batch_{nb}" / "*")))
        assert len(runs) == 6                          # 3 x 2 combinations per batch, one timestamped folder each
        for d in runs:
            assert len(os.path.basename(d)) == len("2024-01-01_00-00-00")
            with open(os.path.join(d, "env_params.json")) as f:
                saved = json.load(f)
            assert saved["SAVE_ROOT_DIR"].endswith(os.path.join("sweep", f"batch_{nb}")) and saved["N"] == "12"
            b = sim.saved_dirs.index(d)
            ori = read_zarr_v2(os.path.join(d, "ag_ori.zarr"))
            posx = read_zarr_v2(os.path.join(d, "ag_posx.zarr"))
            assert ori.shape == (12, 20) and posx.shape == (12, 20)
            # the replicate's own trajectory, in the reference's row order (agent id - 1: agent 0 in the last row)
            assert np.array_equal(ori[:, -1], np.roll(st["theta"][b].astype(np.float64), -1))
            assert np.array_equal(posx[:, -1], np.trunc(np.roll(st["x"][b].astype(np.float64), -1)))   # int(position), ifdb.py:83-84
            if float(saved["VF_ALP0"]) == 0 and float(saved["VF_BET0"]) == 0:
                zero_gain.append(ori)
    # replicates with ALP0 = BET0 = 0 only relax their speed towards V0: their headings never change
    # (away from the walls: reflections turn them)
    assert len(zero_gain) == 2
    for ori in zero_gain:
        unchanged = (ori == ori[:, :1]).all(axis=1)
        assert unchanged.sum() >= 8
    assert mp.run_protocols(project="VisualFlocking") == []      # env files are consumed


def test_metaprotocol_foraging_sweep_groups_by_shape(built_lib, tmp_path):
    """The foraging project through MetaProtocol (the reference's figure experiments are such sweeps): env files that
    differ only in decision / movement parameters run as ONE batch with one parameter set per replicate; a parameter
    that changes the batch's shape (N_RESOURCES) starts another batch.  Replicates with DEC_EPSU = 0 never build up the
    exploitation drive, so only the others can collect; MOV_EXP_VEL_MAX is the speed of every exploring agent of its
    replicate."""
    from abm_b200 import metarunner as mr
    env = dict(N="16", T="250", VISUAL_FIELD_RESOLUTION="320", ENV_WIDTH="300", ENV_HEIGHT="300", RADIUS_AGENT="10",
               AGENT_FOV="1", VISION_RANGE="2000", RADIUS_RESOURCE="45", MIN_RESOURCE_PER_PATCH="400",
               MAX_RESOURCE_PER_PATCH="401", MIN_RESOURCE_QUALITY="0.25", MAX_RESOURCE_QUALITY="0.25", VISUAL_EXCLUSION="1",
               TELEPORT_TO_MIDDLE="0", PATCH_BORDER_OVERLAP="1", AGENT_AGENT_COLLISION="0", GHOST_WHILE_EXPLOIT="1",
               AGENT_CONSUMPTION="1", REGENERATE_PATCHES="1", USE_RAM_LOGGING="0", DEC_EPSW="2")
    mp = mr.MetaProtocol("forage", num_batches=1, default_envconf=env, root_dir=str(tmp_path))
    mp.add_criterion(mr.Tunable("DEC_EPSU", values_override=[0, 3]))
    mp.add_criterion(mr.Tunable("MOV_EXP_VEL_MAX", values_override=[1.5, 3]))
    mp.add_criterion(mr.Tunable("N_RESOURCES", values_override=[2, 4]))
    assert mp.generate_temp_env_files() == 8
    res = mp.run_protocols(project="Base", seed=3)
    assert len(res) == 2 and sorted(sim.N_resc for _, sim in res) == [2, 4]         # one batch per shape
    for paths, sim in res:
        assert len(paths) == 4 and sim.B == 4 and sim.t == 250
        a = sim.engine.get_agents()
        assert np.isfinite(a["x"]).all()
        # the env files are consumed; the per-replicate parameters are in env_params
        for b, e in enumerate(sim.env_params):
            eps_u, vmax = float(e["DEC_EPSU"]), float(e["MOV_EXP_VEL_MAX"])
            if eps_u == 0:
                assert (a["collected"][b] == 0).all()
            explore = a["mode"][b] == 0
            assert np.allclose(np.abs(a["vel"][b][explore]), vmax, rtol=1e-6)
        assert any(float(e["DEC_EPSU"]) > 0 and a["collected"][b].sum() > 0 for b, e in enumerate(sim.env_params))


def test_metaprotocol_foraging_sweep_over_fov_is_one_batch(built_lib, tmp_path):
    """A sweep over AGENT_FOV (the reference's figExp2A files sweep it from 0 to 1) and VISION_RANGE runs as ONE batch of
    the foraging engine -- FOV and vision range are per-agent quantities there -- and every replicate sees through ITS
    field of view: nothing at 0, only the front half of the ring at 0.5, everything at 1."""
    from abm_b200 import metarunner as mr
    env = dict(N="20", T="60", VISUAL_FIELD_RESOLUTION="400", ENV_WIDTH="250", ENV_HEIGHT="250", RADIUS_AGENT="10",
               N_RESOURCES="2", RADIUS_RESOURCE="60", MIN_RESOURCE_PER_PATCH="4000", MAX_RESOURCE_PER_PATCH="4001",
               MIN_RESOURCE_QUALITY="0.25", MAX_RESOURCE_QUALITY="0.25", VISUAL_EXCLUSION="0", TELEPORT_TO_MIDDLE="0",
               PATCH_BORDER_OVERLAP="1", AGENT_AGENT_COLLISION="0", GHOST_WHILE_EXPLOIT="1", AGENT_CONSUMPTION="1",
               REGENERATE_PATCHES="1", USE_RAM_LOGGING="0", DEC_EPSW="2", DEC_EPSU="3")
    mp = mr.MetaProtocol("fov", num_batches=1, default_envconf=env, root_dir=str(tmp_path))
    mp.add_criterion(mr.Tunable("AGENT_FOV", values_override=[0, 0.5, 1]))
    mp.add_criterion(mr.Tunable("VISION_RANGE", values_override=[2000, 30]))
    assert mp.generate_temp_env_files() == 6
    res = mp.run_protocols(project="Base", seed=5)
    assert len(res) == 1 and res[0][1].B == 6                     # ONE batch
    sim = res[0][1]
    seen = {}
    for _ in range(40):                                           # the union of the fields over some more steps
        sim.step_sim()
        f = sim.engine.fields()
        for b, e in enumerate(sim.env_params):
            key = (float(e["AGENT_FOV"]), int(e["VISION_RANGE"]))
            seen[key] = seen.get(key, 0) | f[b].any(axis=0)
    R = 400
    assert not seen[(0.0, 2000)].any() and not seen[(0.0, 30)].any()
    half = seen[(0.5, 2000)]
    assert half.any() and not half[:R // 4 - 1].any() and not half[3 * R // 4 + 2:].any()     # only the front half
    full = seen[(1.0, 2000)]
    assert full[:R // 4 - 1].any() and full[3 * R // 4 + 2:].any()
    assert seen[(1.0, 30)].sum() < full.sum()                      # a 30-px vision range sees fewer exploiters


def test_metaprotocol_foraging_sweep_over_patch_parameters_is_one_batch(built_lib, tmp_path):
    """A sweep over the patch parameters (the reference's figExp3 file -- BASELINE configs[2] -- sweeps RADIUS_RESOURCE and
    DEC_EPSW) runs as ONE batch: every replicate creates AND regenerates its patches with its own radius, units and
    quality (add_new_resource_patch, sims.py:332-374, reads them from the simulation)."""
    from abm_b200 import metarunner as mr
    env = dict(N="16", T="400", VISUAL_FIELD_RESOLUTION="320", ENV_WIDTH="300", ENV_HEIGHT="300", RADIUS_AGENT="10",
               AGENT_FOV="1", VISION_RANGE="2000", N_RESOURCES="3", MAX_RESOURCE_PER_PATCH="-1", MAX_RESOURCE_QUALITY="-1",
               VISUAL_EXCLUSION="1", TELEPORT_TO_MIDDLE="0", PATCH_BORDER_OVERLAP="1", AGENT_AGENT_COLLISION="0",
               GHOST_WHILE_EXPLOIT="1", AGENT_CONSUMPTION="1", REGENERATE_PATCHES="1", USE_RAM_LOGGING="0",
               DEC_EPSW="2", DEC_EPSU="3", DEC_SWU="0", DEC_SUW="0")
    mp = mr.MetaProtocol("patches", num_batches=1, default_envconf=env, root_dir=str(tmp_path))
    mp.add_criterion(mr.Tunable("RADIUS_RESOURCE", values_override=[20, 45]))
    mp.add_criterion(mr.Tunable("MIN_RESOURCE_PER_PATCH", values_override=[4, 60]))
    mp.add_criterion(mr.Tunable("MIN_RESOURCE_QUALITY", values_override=[0.25, 1.0]))
    assert mp.generate_temp_env_files() == 8
    res = mp.run_protocols(project="Base", seed=11)
    assert len(res) == 1 and res[0][1].B == 8
    sim = res[0][1]
    p, cnt = sim.engine.get_patches(), sim.engine.counters()
    assert cnt["patches_regenerated"] > 0 and cnt["regeneration_failed"] == 0      # the 4-unit patches are eaten up
    for b, e in enumerate(sim.env_params):
        R, units, q = float(e["RADIUS_RESOURCE"]), int(e["MIN_RESOURCE_PER_PATCH"]), float(e["MIN_RESOURCE_QUALITY"])
        assert (p["radius"][b] == R).all()                                          # created and re-created with ITS radius
        assert (p["quality"][b] == np.float32(q)).all()                             # max < 0: the minimum is the value
        assert (p["left"][b] <= units).all() and (p["left"][b] > 0).all()           # units in [min, min + 1)


def test_simulation_writes_reference_output_folder(built_lib, tmp_path):
    """Simulation(use_ram_logging, save_csv_files, use_zarr) -> <root>/<SAVE_ROOT_DIR>/<timestamp>/ with the agent and
    resource arrays of ifdb.py:435-535 and env_params.json; saving without logging raises like sims.py:909-912."""
    import os
    from abm_b200.recorder import read_zarr_v2
    from abm_b200.simulation import Simulation
    kw = dict(N=10, T=15, v_field_res=1200, width=500, height=500, agent_radius=10, N_resc=3, patch_radius=30,
              min_resc_perpatch=30, max_resc_perpatch=40, min_resc_quality=0.25, max_resc_quality=-1,
              vision_range=2000, visual_exclusion=True, teleport_exploit=False, seed=3)
    with pytest.raises(Exception, match="Nothing to save"):
        Simulation(save_csv_files=True, use_ram_logging=False, **kw)
    sim = Simulation(use_ram_logging=True, save_csv_files=True, use_zarr=True, root_dir=str(tmp_path),
                     save_root_dir="abm/data/simulation_data/single", env_params={"N": "10", "T": "15"}, **kw)
    sim.start()
    assert len(sim.saved_dirs) == 1 and sim.saved_dirs[0].startswith(str(tmp_path / "abm/data/simulation_data/single"))
    d = sim.saved_dirs[0]
    for name in ("ag_posx", "ag_posy", "ag_ori", "ag_vel", "ag_mode", "ag_w", "ag_u", "ag_ipriv", "ag_collr", "ag_explr"):
        assert read_zarr_v2(os.path.join(d, name + ".zarr")).shape == (10, 15), name
    for name in ("res_posx", "res_posy", "res_rad", "res_left", "res_qual"):
        assert read_zarr_v2(os.path.join(d, name + ".zarr")).shape == (3, 15), name
    assert os.path.isfile(os.path.join(d, "env_params.json"))
    a = sim.engine.get_agents()
    from abm_b200.recorder import agent_row
    w_rows = read_zarr_v2(os.path.join(d, "ag_w.zarr"))[:, -1]
    assert np.array_equal(w_rows, np.roll(a["w"][0].astype(np.float64), -1))      # agent i in row i - 1, agent 0 in the last
    assert agent_row(0, 10) == 9 and agent_row(3, 10) == 2 and w_rows[agent_row(3, 10)] == np.float64(a["w"][0][3])


def test_tiled_swarm_matches_single_gpu_if_two_gpus(built_lib):
    """Large-swarm mode (agent tiles + one all-gather per step) over NCCL; needs >= 2 GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29544",
                        os.path.join(root, "tests", "multigpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTIGPU_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_recorder_writes_reference_zarr_layout(built_lib, tmp_path):
    """per-step state appended on the device, chunks copied out asynchronously and written as the
    reference's `ag_*.zarr` arrays (ifdb.py:470-508: shape (num_agents, T), float64; posx / posy int-truncated
    like ifdb.py:83-84) -- checked against per-step get_state() and by parsing the zarr-v2 metadata."""
    import json
    from abm_b200 import VFEngine
    from abm_b200.recorder import VFRecorder, read_zarr_v2
    rng = np.random.default_rng(3)
    B, N, W, T = 4, 40, 500.0, 11
    x = rng.uniform(60, W, (B, N)).astype(np.float32); y = rng.uniform(60, W, (B, N)).astype(np.float32)
    th = rng.uniform(0, 2 * np.pi, (B, N)).astype(np.float32); v = rng.uniform(0, 2, (B, N)).astype(np.float32)
    eng = VFEngine(B, N, resolution=1200, width=W, height=W)
    eng.set_params(); eng.set_state(x, y, th, v, 10.0)
    rec = VFRecorder(eng, str(tmp_path / "run"), replicates=[0, 2], every=1, chunk=4, env_params={"N": N, "T": T})
    expect = []
    for _ in range(T):
        eng.step(1)
        rec.record()
        expect.append(eng.get_state())
    dirs = rec.close()
    assert [os.path.basename(d) for d in dirs] == ["replicate_00000", "replicate_00002"]
    for d, b in zip(dirs, (0, 2)):
        meta = json.load(open(os.path.join(d, "ag_posx.zarr", ".zarray")))
        assert meta["shape"] == [N, T] and meta["chunks"] == [N, 4] and meta["dtype"] == "<f8" and meta["zarr_format"] == 2
        assert json.load(open(os.path.join(d, "env_params.json")))["T"] == T
        for name, key, trunc in (("posx", "x", True), ("posy", "y", True), ("ori", "theta", False), ("vel", "vel", False)):
            got = read_zarr_v2(os.path.join(d, f"ag_{name}.zarr"))
            want = np.roll(np.stack([e[key][b].astype(np.float64) for e in expect], axis=1), -1, axis=0)   # row = agent id - 1
            assert np.array_equal(got, np.trunc(want) if trunc else want), name
        assert not read_zarr_v2(os.path.join(d, "ag_mode.zarr")).any()
    eng.close()


@pytest.mark.parametrize("exp_file,project_dirs", [("VFExp4c.py", 363), ("figExp3BN50PatchyCollOcc.py", 70)])
def test_references_experiment_files_run_unchanged_on_the_engine(built_lib, tmp_path, exp_file, project_dirs):
    """The drop-in claim end to end: the reference's OWN experiment file (BASELINE configs[3] / configs[2]; an unmodified
    copy under oracle/_ref) is executed in a fresh interpreter with `import abm_b200.compat` in front and nothing else
    changed but the run length (T = 25000 -> 120): its imports, its criteria, `generate_temp_env_files()` and
    `run_protocols()` all run on the B200 engine, and every run of the sweep leaves the reference's output folder."""
    import glob
    import shutil
    import subprocess
    import sys
    import textwrap
    from oracle import ref_shim
    ref_root = ref_shim._reference_root()
    exp_path = os.path.join(ref_root, "abm/data/metaprotocol/experiments", exp_file)
    if not os.path.isfile(exp_path):
        pytest.skip("the reference's experiment files are not at hand")
    shutil.copyfile(os.path.join(ref_root, ".env"), tmp_path / "dropin.env")        # the default env, `<root>/<EXP>.env`
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {repo!r})
        import abm_b200.compat
        src = open({exp_path!r}).read()
        assert src.count('Constant("T", 25000)') == 1
        exec(compile(src.replace('Constant("T", 25000)', 'Constant("T", 120)'), {exp_path!r}, "exec"))
        print("EXPERIMENT_DONE")
    """)
    env = dict(os.environ, EXPERIMENT_NAME="dropin")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=str(tmp_path), env=env)
    assert "EXPERIMENT_DONE" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    runs = glob.glob(str(tmp_path / "abm/data/simulation_data/dropin/batch_0/*"))
    assert len(runs) == project_dirs
    from abm_b200.recorder import read_zarr_v2
    ori = read_zarr_v2(os.path.join(sorted(runs)[0], "ag_ori.zarr"))
    assert ori.shape[1] == 120 and np.isfinite(ori).all()
    assert os.path.isfile(os.path.join(sorted(runs)[-1], "env_params.json"))
    assert glob.glob(str(tmp_path / "abm/data/metaprotocol/temp/dropin/*.env")) == []   # every protocol was consumed
