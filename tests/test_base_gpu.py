"""GPU parity tests of the BASE / foraging path (through the C ABI) against fixtures produced
by the unmodified reference's Agent.update and against the CPU oracle (oracle/restate_base.py).
Fields: bit-exact.  Scalars (w, u, velocity, heading, position): 1e-5 relative (fp32 state)."""
import numpy as np
import pytest

from golden_io import BASE_OUT_KEYS, load_base_cases, load_base_hetero_cases
from oracle import restate as rs
from oracle import restate_base as rb

pytestmark = pytest.mark.gpu
RTOL = 1e-5
PHASE_ENV, PHASE_AGENTS = 1, 2


@pytest.fixture(autouse=True, params=["one_cta_per_replicate", "grid_per_phase"])
def step_path(request, monkeypatch):
    """Every test of this module runs on BOTH step paths: the fused kernel (a CTA per replicate runs collisions,
    environment and agent phase) and one grid per phase with a warp per focal agent.  The engine picks between them from
    the batch shape (abm_base_api.cu); ABM_BASE_FUSED forces the choice."""
    monkeypatch.setenv("ABM_BASE_FUSED", "1" if request.param == "one_cta_per_replicate" else "0")
    return request.param


def _engine_for(cfg, B, N, P=0, **kw):
    from abm_b200 import BaseEngine
    fovr = cfg.fov[1] / np.pi
    eng = BaseEngine(B, N, P, resolution=cfg.R, agent_fov=fovr, width=cfg.width, height=cfg.height,
                     vision_range=cfg.vision_range, agent_radius=10.0, visual_exclusion=cfg.visual_exclusion,
                     patchwise_exclusion=cfg.patchwise_exclusion, teleport_exploit=cfg.teleport_exploit,
                     tau=cfg.Tau, keep_fields=True, **kw)
    eng.set_params(T_w=cfg.T_w, Eps_w=cfg.Eps_w, g_w=cfg.g_w, B_w=cfg.B_w, w_max=cfg.w_max, T_u=cfg.T_u,
                   Eps_u=cfg.Eps_u, g_u=cfg.g_u, B_u=cfg.B_u, u_max=cfg.u_max, S_wu=cfg.S_wu, S_uw=cfg.S_uw,
                   F_N=cfg.F_N, F_R=cfg.F_R, exp_vel_max=cfg.exp_vel_max, exp_theta_min=cfg.exp_theta_min,
                   exp_theta_max=cfg.exp_theta_max, reloc_theta_max=cfg.reloc_theta_max,
                   exp_stop_ratio=cfg.exp_stop_ratio, agent_consumption=cfg.agent_consumption)
    return eng


def _upload(eng, st, B=1):
    def r(a):
        return np.asarray(a).reshape((B, -1) + np.asarray(a).shape[2:] if B > 1 else (1,) + np.asarray(a).shape)
    eng.set_agents(x=r(st["x"]), y=r(st["y"]), theta=r(st["theta"]), vel=r(st["vel"]), w=r(st["w"]), u=r(st["u"]),
                   collected=r(st["collected"]), collected_before=r(st["collected_before"]),
                   env_status=r(st["env_status"]), override_mode=r(st["override"]), mode=r(st["mode"]),
                   patch_id=r(st["patch_id"]), novelty=r(st["novelty"]))


def _compare_agents(got, ref, b=0, idx=None):
    names = dict(x="x", y="y", theta="theta", vel="vel", w="w", u="u", I_priv="i_priv",
                 collected_before="collected_before")
    sel = slice(None) if idx is None else idx
    for k, g in names.items():
        np.testing.assert_allclose(got[g][b][sel], np.asarray(ref[k])[sel], rtol=RTOL, atol=1e-5, err_msg=k)
    assert np.array_equal(got["override_mode"][b][sel], np.asarray(ref["override"])[sel].astype(int))
    assert np.array_equal(got["mode"][b][sel], np.asarray(ref["mode"])[sel].astype(int))


@pytest.mark.parametrize("case", load_base_cases(), ids=lambda c: f"N{len(c['dth'])}_R{c['cfg'].R}")
def test_agent_phase_matches_reference_fixture(built_lib, case):
    cfg, st = case["cfg"], case["st"]
    N = len(case["dth"])
    eng = _engine_for(cfg, 1, N)
    _upload(eng, st)
    eng.step(1, inject_dtheta=case["dth"], phases=PHASE_AGENTS)
    assert np.array_equal(rs.pack_bits(eng.fields()[0]), case["fields"])          # bit-exact stored fields
    _compare_agents(eng.get_agents(), case["out"])
    eng.close()


@pytest.mark.parametrize("case", load_base_hetero_cases(), ids=lambda c: f"N{len(c['dth'])}_R{c['cfg'].R}")
def test_agent_phase_matches_reference_fixture_heterogeneous_agents(built_lib, case):
    """One parameter set per agent (agent_behave_param_list, sims.py:499-517; agent.py:83-108) and, in the last two
    cases, the agent's own FOV and vision range; the fixture went through the reference's own constructor.  Two replicates: the second gets the agents' sets in reverse order,
    so that the (replicate, agent) indexing of the table is exercised."""
    cfg, st = case["cfg"], case["st"]
    N = len(case["dth"])
    eng = _engine_for(cfg, 2, N)
    geo = ("agent_fov", "vision_range")
    eng.set_params(exp_theta_min=cfg.exp_theta_min, exp_theta_max=cfg.exp_theta_max,
                   reloc_theta_max=cfg.reloc_theta_max,
                   **{k: np.stack([v, v[::-1]]) for k, v in case["agent_params"].items() if k not in geo})
    eng.set_agent_geometry(**{k: np.stack([case["agent_params"][k], case["agent_params"][k][::-1]]) for k in geo})
    two = {k: np.stack([np.asarray(v), np.asarray(v)]) for k, v in st.items() if k != "radius"}
    eng.set_agents(x=two["x"], y=two["y"], theta=two["theta"], vel=two["vel"], w=two["w"], u=two["u"],
                   collected=two["collected"], collected_before=two["collected_before"],
                   env_status=two["env_status"], override_mode=two["override"], mode=two["mode"],
                   patch_id=two["patch_id"], novelty=two["novelty"])
    eng.step(1, inject_dtheta=np.stack([case["dth"], case["dth"]]), phases=PHASE_AGENTS)
    got = eng.get_agents()
    assert np.array_equal(rs.pack_bits(eng.fields()[0]), case["fields"])          # bit-exact stored fields
    _compare_agents(got, case["out"], 0)
    ref1 = rb.base_step_frozen(st, cfg, case["dth"], agent_cfgs=case["agent_cfgs"][::-1])
    _compare_agents(got, ref1, 1)
    assert not np.allclose(got["w"][0], got["w"][1])
    assert np.array_equal(eng.fields()[1], ref1["fields"])
    if len(set(case["agent_params"]["agent_fov"])) > 1:                   # back to the engine-wide geometry
        eng.set_agent_geometry()
        eng.set_agents(x=two["x"], y=two["y"], theta=two["theta"], vel=two["vel"], w=two["w"], u=two["u"],
                       collected=two["collected"], collected_before=two["collected_before"],
                       env_status=two["env_status"], override_mode=two["override"], mode=two["mode"],
                       patch_id=two["patch_id"], novelty=two["novelty"])
        eng.step(1, inject_dtheta=np.stack([case["dth"], case["dth"]]), phases=PHASE_AGENTS)
        import dataclasses
        same_geo = [dataclasses.replace(c, fov=cfg.fov, vision_range=cfg.vision_range) for c in case["agent_cfgs"]]
        ref2 = rb.base_step_frozen(st, cfg, case["dth"], agent_cfgs=same_geo)
        assert np.array_equal(eng.fields()[0], ref2["fields"])
        assert not np.array_equal(rs.pack_bits(ref2["fields"]), case["fields"])   # the geometry mattered
    eng.close()


def test_env_phase_with_per_agent_consumption(built_lib):
    """agent.consumption in the depletion order (sims.py:824-828, rescource.py:118-133), one parameter set per
    agent, against the oracle."""
    import dataclasses
    from abm_b200 import BaseEngine
    rng = np.random.default_rng(18)
    B, N, P, W = 3, 40, 3, 300.0
    cfg = rb.BaseConfig(R=320, width=W, height=W, vision_range=2000.0, visual_exclusion=True)
    cons = rng.choice([0.25, 0.5, 1.0, 2.0], (B, N))
    vmax = np.asarray(rng.uniform(1, 4, (B, N)), np.float32).astype(np.float64)
    eng = BaseEngine(B, N, P, resolution=320, width=W, height=W, regenerate_patches=False, tau=cfg.Tau,
                     vision_range=2000.0, visual_exclusion=True, collide_agents=True, ghost_mode=False)
    eng.set_params(agent_consumption=cons, exp_vel_max=vmax)
    states, patches = [], []
    for b in range(B):
        st = _random_state(rng, N, W, cfg)
        st["override"] = rng.choice([0, 1], N); st["mode"] = st["override"].copy()
        states.append(st)
        patches.append(dict(x=np.array([40.0, 150.0, 230.0]), y=np.array([50.0, 160.0, 60.0]),
                            radius=np.array([45.0, 40.0, 35.0]), left=np.array([2.5, 400.0, 0.5]),
                            quality=np.array([0.75, 1.5, 1.0]), id=np.array([0, 1, 2])))
    S = {k: np.stack([s[k] for s in states]) for k in states[0] if k != "radius"}
    Pm = {k: np.stack([p[k] for p in patches]) for k in patches[0]}
    eng.set_agents(x=S["x"], y=S["y"], theta=S["theta"], vel=S["vel"], w=S["w"], u=S["u"], collected=S["collected"],
                   collected_before=S["collected_before"], env_status=S["env_status"], override_mode=S["override"],
                   mode=S["mode"], patch_id=S["patch_id"], novelty=S["novelty"])
    eng.set_patches(**Pm)
    eng.step(1, phases=PHASE_ENV)
    got, gp = eng.get_agents(), eng.get_patches()
    takes = set()
    for b in range(B):
        acfg = [dataclasses.replace(cfg, agent_consumption=float(cons[b, i]), exp_vel_max=float(vmax[b, i]))
                for i in range(N)]
        st = {k: (np.array(v, dtype=float) if k in ("theta", "collected", "collected_before") else np.array(v))
              for k, v in states[b].items()}
        before = st["collected"].copy()
        rb.base_patch_phase(st, patches[b], cfg, agent_cfgs=acfg)
        takes |= set(np.round(st["collected"] - before, 6).tolist())
        np.testing.assert_allclose(got["collected"][b], st["collected"], rtol=RTOL)
        np.testing.assert_allclose(got["collected_before"][b], st["collected_before"], rtol=RTOL)
        np.testing.assert_allclose(gp["left"][b], patches[b]["left"], rtol=RTOL, atol=1e-6)
        assert np.array_equal(got["env_status"][b], st["env_status"])
    assert len(takes) >= 4          # several different amounts were taken: the per-agent consumption is in use
    eng.close()


def _random_state(rng, N, W, cfg):
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    override = rng.choice([0, 0, 1, 1, 3], N)
    st = dict(x=rng.integers(20, 30 + int(W), N).astype(float), y=rng.integers(20, 30 + int(W), N).astype(float),
              theta=f32(rng.uniform(0, 2 * np.pi, N)), vel=f32(rng.uniform(0, 3, N)), radius=10.0,
              w=f32(rng.uniform(-0.2, 1, N)), u=f32(rng.uniform(-0.2, 1, N)),
              novelty=(rng.uniform(0, 1, (N, cfg.Tau)) < 0.15).astype(float),
              env_status=rng.choice([-1, 1], N), override=override,
              mode=np.where(override == 1, 1, np.where(override == 3, 3, 0)),
              patch_id=rng.choice([-1, 0, 1, 2], N), collected=f32(rng.uniform(0, 5, N)))
    st["collected_before"] = f32(st["collected"] - rng.choice([0.0, 0.25, 1.0], N))
    return st


@pytest.mark.parametrize("B,N,R,W,vis_excl,fovr", [
    (4, 50, 1200, 500.0, True, 1.0),        # config 3 shape
    (3, 10, 1200, 900.0, False, 0.5),       # config 1 shape
    (2, 100, 1200, 500.0, True, 1.0),       # the reference's figure experiments (N = 100): two mask words in the fused kernel
    (2, 65, 1200, 300.0, True, 1.0),        # one object beyond the first mask word, crowded
    (1, 128, 1200, 350.0, True, 1.0),       # the largest replicate of the fused kernel, crowded
    (2, 37, 601, 300.0, True, 0.75),
    (1, 200, 1200, 400.0, True, 1.0),       # crowded: heavy occlusion
    (1, 20, 8192, 250.0, True, 1.0),        # the largest resolution of the fused kernel (16-bit packed interval ends)
    (1, 20, 10000, 250.0, True, 0.8),       # beyond it: the per-phase kernels whatever the batch shape
    (2, 12, 8, 200.0, True, 1.0),           # the smallest ring the engine accepts
])
def test_agent_phase_matches_oracle_random(built_lib, B, N, R, W, vis_excl, fovr):
    rng = np.random.default_rng(100 + N + R)
    cfg = rb.BaseConfig(R=R, fov=(-fovr * np.pi, fovr * np.pi), width=W, height=W, visual_exclusion=vis_excl,
                        Eps_w=2.0, Eps_u=1.0, exp_vel_max=3.0, exp_theta_min=-0.5, exp_theta_max=0.5,
                        reloc_theta_max=1.8, exp_stop_ratio=0.175, F_N=0.5, F_R=0.5)
    states = [_random_state(rng, N, W, cfg) for _ in range(B)]
    dth = np.asarray(rng.uniform(-0.5, 0.5, (B, N)), np.float32)
    eng = _engine_for(cfg, B, N)
    stacked = {k: np.stack([s[k] for s in states]) for k in states[0] if k != "radius"}
    eng.set_agents(x=stacked["x"], y=stacked["y"], theta=stacked["theta"], vel=stacked["vel"], w=stacked["w"],
                   u=stacked["u"], collected=stacked["collected"], collected_before=stacked["collected_before"],
                   env_status=stacked["env_status"], override_mode=stacked["override"], mode=stacked["mode"],
                   patch_id=stacked["patch_id"], novelty=stacked["novelty"])
    eng.step(1, inject_dtheta=dth, phases=PHASE_AGENTS)
    got, fields = eng.get_agents(), eng.fields()
    sample = None if N <= 100 else np.sort(rng.choice(N, 40, replace=False))
    for b in range(B):
        ref = rb.base_step_frozen(states[b], cfg, dth[b].astype(np.float64),
                                  agents=None if sample is None else sample.tolist())
        sel = slice(None) if sample is None else sample
        assert np.array_equal(fields[b][sel], ref["fields"][sel])
        _compare_agents(got, ref, b, sample)
    eng.close()


def test_env_phase_matches_oracle(built_lib):
    """Agent-patch interaction (sims.py:790-858): membership, heading bias, notify / novelty,
    depletion order, depletion of a patch within the step (regeneration switched off)."""
    rng = np.random.default_rng(8)
    B, N, P, W = 5, 40, 3, 300.0
    cfg = rb.BaseConfig(R=320, width=W, height=W, agent_consumption=1.0)
    from abm_b200 import BaseEngine
    eng = BaseEngine(B, N, P, resolution=320, width=W, height=W, regenerate_patches=False, tau=cfg.Tau)
    eng.set_params(agent_consumption=1.0)
    states, patches = [], []
    for b in range(B):
        st = _random_state(rng, N, W, cfg)
        st["override"] = rng.choice([0, 1], N); st["mode"] = st["override"].copy()
        states.append(st)
        patches.append(dict(x=np.array([40.0, 150.0, 230.0]), y=np.array([50.0, 160.0, 60.0]),
                            radius=np.array([45.0, 40.0, 35.0]), left=np.array([2.5, 400.0, 0.5]),
                            quality=np.array([0.75, 0.25, 1.0]), id=np.array([0, 1, 2])))
    S = {k: np.stack([s[k] for s in states]) for k in states[0] if k != "radius"}
    Pm = {k: np.stack([p[k] for p in patches]) for k in patches[0]}
    eng.set_agents(x=S["x"], y=S["y"], theta=S["theta"], vel=S["vel"], w=S["w"], u=S["u"], collected=S["collected"],
                   collected_before=S["collected_before"], env_status=S["env_status"], override_mode=S["override"],
                   mode=S["mode"], patch_id=S["patch_id"], novelty=S["novelty"])
    eng.set_patches(**Pm)
    eng.step(1, phases=PHASE_ENV)
    got, gp = eng.get_agents(), eng.get_patches()
    any_depleted = False
    for b in range(B):
        st, pa = states[b], patches[b]
        st = {k: (np.array(v, dtype=float) if k in ("theta", "collected", "collected_before") else np.array(v))
              for k, v in st.items()}
        depleted = rb.base_patch_phase(st, pa, cfg)
        any_depleted |= bool(depleted)
        np.testing.assert_allclose(got["theta"][b], st["theta"], rtol=RTOL)
        np.testing.assert_allclose(got["collected"][b], st["collected"], rtol=RTOL)
        np.testing.assert_allclose(got["collected_before"][b], st["collected_before"], rtol=RTOL)
        assert np.array_equal(got["env_status"][b], st["env_status"])
        assert np.array_equal(got["patch_id"][b], st["patch_id"])
        nov = ((got["novelty"][b][:, None] >> np.arange(cfg.Tau)) & 1).astype(float)
        assert np.array_equal(nov, st["novelty"])
        np.testing.assert_allclose(gp["left"][b], pa["left"], rtol=RTOL, atol=1e-6)
        for p in depleted:
            assert gp["radius"][b][p] == 0.0        # killed patch (no regeneration)
    assert any_depleted
    eng.close()


@pytest.mark.parametrize("teleport", [False, True])
def test_env_phase_with_many_patches_matches_oracle(built_lib, teleport):
    """The patch counts of the reference's figure experiments (N_RESOURCES up to 100): most patches have nobody on them
    and are skipped by the kernels (a parallel pre-test, abm_base.cu) -- against the oracle, which visits every patch in
    order like the reference (sims.py:790-858).  Some patches overlap, some contain another patch's centre (an agent
    teleported to a patch centre lands inside the next patch), some are dead (radius 0), one replicate has no agent on
    any patch."""
    rng = np.random.default_rng(21)
    B, N, P, W = 4, 70, 90, 500.0
    cfg = rb.BaseConfig(R=320, width=W, height=W, agent_consumption=1.0, teleport_exploit=teleport)
    from abm_b200 import BaseEngine
    eng = BaseEngine(B, N, P, resolution=320, width=W, height=W, regenerate_patches=False, tau=cfg.Tau,
                     teleport_exploit=teleport)
    eng.set_params(agent_consumption=1.0)
    states, patches = [], []
    for b in range(B):
        st = _random_state(rng, N, W, cfg)
        st["override"] = rng.choice([0, 1], N); st["mode"] = st["override"].copy()
        px, py = rng.uniform(20, 480, P), rng.uniform(20, 480, P)
        pr = rng.choice([8.0, 15.0, 25.0], P)
        px[1], py[1], pr[1] = px[0] + pr[0] - 20.0 + 3.0, py[0] + pr[0] - 20.0, 20.0   # centre of patch 0 inside patch 1
        pr[5] = 0.0                                                                     # a killed patch
        if b == 3:                                                                      # nobody anywhere near a patch
            st["x"] = rng.integers(600, 700, N).astype(float); st["y"] = rng.integers(600, 700, N).astype(float)
        else:                                                                           # a few agents right on patch 0
            st["x"][:4] = px[0] + pr[0] - 10.0 + np.array([0.0, 1.0, -2.0, 3.0]); st["y"][:4] = py[0] + pr[0] - 10.0
        states.append(st)
        patches.append(dict(x=px, y=py, radius=pr, left=rng.choice([0.5, 3.0, 400.0], P),
                            quality=rng.choice([0.25, 0.75, 1.0], P), id=np.arange(P)))
    S = {k: np.stack([s[k] for s in states]) for k in states[0] if k != "radius"}
    Pm = {k: np.stack([p_[k] for p_ in patches]) for k in patches[0]}
    eng.set_agents(x=S["x"], y=S["y"], theta=S["theta"], vel=S["vel"], w=S["w"], u=S["u"], collected=S["collected"],
                   collected_before=S["collected_before"], env_status=S["env_status"], override_mode=S["override"],
                   mode=S["mode"], patch_id=S["patch_id"], novelty=S["novelty"])
    eng.set_patches(**Pm)
    eng.step(1, phases=PHASE_ENV)
    got, gp = eng.get_agents(), eng.get_patches()
    n_on, n_depleted = 0, 0
    for b in range(B):
        st = {k: (np.array(v, dtype=float) if k in ("x", "y", "theta", "collected", "collected_before") else np.array(v))
              for k, v in states[b].items()}
        pa = {k: np.array(v, dtype=float if k != "id" else int) for k, v in patches[b].items()}
        depleted = rb.base_patch_phase(st, pa, cfg)
        n_depleted += len(depleted); n_on += int((st["env_status"] == 1).sum())
        for k in ("x", "y", "theta", "collected", "collected_before"):
            np.testing.assert_allclose(got[k][b], st[k], rtol=RTOL, atol=1e-5, err_msg=k)
        assert np.array_equal(got["env_status"][b], st["env_status"])
        assert np.array_equal(got["patch_id"][b], st["patch_id"])
        np.testing.assert_allclose(gp["left"][b], pa["left"], rtol=RTOL, atol=1e-6)
        for p_ in depleted:
            assert gp["radius"][b][p_] == 0.0
    assert n_on > 20 and n_depleted > 0
    eng.close()


def test_full_loop_runs_and_forages(built_lib, step_path):
    """configs[0] shape (N=10, 3 patches, R=1200) for 300 steps with the engine's own RNG:
    sanity properties -- agents stay inside the arena, some resource gets collected, patches
    are regenerated, state stays finite, replicates with different seeds differ."""
    from abm_b200 import BaseEngine
    B, N, P, W = 8, 10, 3, 500.0
    rng = np.random.default_rng(0)
    eng = BaseEngine(B, N, P, resolution=1200, width=W, height=W, visual_exclusion=True, patch_radius=30.0,
                     min_resc_perpatch=20, max_resc_perpatch=30, min_resc_quality=0.25, seed=42)
    eng.set_params(Eps_w=2.0, Eps_u=1.0, F_N=0.5, F_R=0.5, exp_vel_max=3.0, exp_theta_min=-0.5, exp_theta_max=0.5,
                   reloc_theta_max=1.8, exp_stop_ratio=0.175)
    eng.set_agents(x=rng.integers(20, 520, (B, N)), y=rng.integers(20, 520, (B, N)),
                   theta=rng.uniform(0, 2 * np.pi, (B, N)))
    eng.set_patches(x=rng.integers(40, 440, (B, P)), y=rng.integers(40, 440, (B, P)), radius=np.full((B, P), 30.0),
                    left=np.full((B, P), 25.0), quality=np.full((B, P), 0.25), id=np.tile(np.arange(P), (B, 1)))
    eng.step(300)
    a = eng.get_agents()
    for k in ("x", "y", "theta", "vel", "w", "u", "collected"):
        assert np.isfinite(a[k]).all(), k
    assert (a["x"] + 10 >= 30 - 1e-3).all() and (a["x"] + 10 <= 30 + W + 1e-3).all()
    assert (a["y"] + 10 >= 30 - 1e-3).all() and (a["y"] + 10 <= 30 + W + 1e-3).all()
    assert a["collected"].sum() > 0
    assert not np.array_equal(a["x"][0], a["x"][1])
    c = eng.counters()
    assert c["steps"] == 300 and c["regeneration_failed"] == 0
    assert c["launches"] == (1 if step_path == "one_cta_per_replicate" else 600)         # ONE launch for all 300 steps
    eng.close()


def test_summary_metrics_match_logged_arrays(built_lib):
    """SURVEY f3 for the foraging path: search efficiency and relative relocation time reduced on the device against
    the reference's offline definitions (data_loader.py :1294-1353, :1903-1928) applied with numpy to the per-step
    `mode` / `collresource` arrays the reference would have logged (fetched from the engine after every step)."""
    from abm_b200 import BaseEngine
    B, N, P, W, T = 6, 12, 3, 500.0, 120
    rng = np.random.default_rng(5)
    eng = BaseEngine(B, N, P, resolution=1200, width=W, height=W, visual_exclusion=True, collide_agents=True,
                     patch_radius=30.0, min_resc_perpatch=20, max_resc_perpatch=30, min_resc_quality=0.25, seed=7)
    eng.set_params(Eps_w=2.0, Eps_u=1.0, F_N=0.5, F_R=0.5, exp_vel_max=3.0, exp_theta_min=-0.5, exp_theta_max=0.5,
                   reloc_theta_max=1.8, exp_stop_ratio=0.175)
    eng.set_agents(x=rng.integers(20, 520, (B, N)), y=rng.integers(20, 520, (B, N)),
                   theta=rng.uniform(0, 2 * np.pi, (B, N)))
    eng.set_patches(x=rng.integers(40, 440, (B, P)), y=rng.integers(40, 440, (B, P)), radius=np.full((B, P), 30.0),
                    left=np.full((B, P), 25.0), quality=np.full((B, P), 0.25), id=np.tile(np.arange(P), (B, 1)))
    mode = np.empty((B, N, T), np.int32)
    coll = np.empty((B, N, T), np.float64)
    for t in range(T):
        eng.step(1)
        a = eng.get_agents()
        mode[..., t] = a["mode"]; coll[..., t] = a["collected"]
    m = eng.metrics()
    eff = coll[..., -1] / T                                                       # collres / dT with t_start = 0
    np.testing.assert_allclose(m["search_efficiency"], eff.mean(axis=1), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(m["mean_collected"], coll[..., -1].mean(axis=1), rtol=1e-5, atol=1e-7)
    for code, key in ((2, "relocation_time"), (0, "explore_time"), (1, "exploit_time"), (3, "collide_time")):
        ref = (mode == code).astype(int).mean(axis=2).mean(axis=1)                # mean over time, then over agents
        np.testing.assert_allclose(m[key], ref, rtol=1e-6, atol=1e-7)
    assert m["relocation_time"].max() > 0 and m["exploit_time"].max() > 0
    # a new time window: the mode fractions start again, collected_r keeps counting
    eng.metrics(reset=True)
    eng.step(10)
    m2 = eng.metrics()
    a = eng.get_agents()
    np.testing.assert_allclose(m2["search_efficiency"], a["collected"].mean(axis=1) / 10, rtol=1e-5, atol=1e-7)
    s = m2["relocation_time"] + m2["explore_time"] + m2["exploit_time"] + m2["collide_time"]
    np.testing.assert_allclose(s, 1.0, rtol=1e-6)
    eng.close()


@pytest.mark.parametrize("ghost,teleport,vis_excl", [(True, False, False), (False, False, True), (True, True, True)])
def test_collision_phase_matches_oracle(built_lib, ghost, teleport, vis_excl):
    """Agent-agent collision avoidance (sims.py:736-783, 421-468) against the oracle restatement
    (pair detection restates pygame's documented collide_circle: parity unpinned vs the reference;
    the proximity field is pinned by test_collision_proximity_matches_live_reference)."""
    from abm_b200 import BaseEngine
    rng = np.random.default_rng(31 + ghost + 2 * teleport)
    B, N, W = 4, 40, 200.0
    cfg = rb.BaseConfig(R=1200, width=W, height=W, visual_exclusion=vis_excl, teleport_exploit=teleport,
                        exp_vel_max=3.0)
    eng = BaseEngine(B, N, 0, resolution=1200, width=W, height=W, visual_exclusion=vis_excl,
                     teleport_exploit=teleport, collide_agents=True, ghost_mode=ghost, tau=cfg.Tau)
    eng.set_params(exp_vel_max=3.0)
    states = []
    for b in range(B):
        st = _random_state(rng, N, W, cfg)
        st["override"] = rng.choice([0, 0, 1, 3], N)
        st["mode"] = st["override"].copy()
        states.append(st)
    S = {k: np.stack([s[k] for s in states]) for k in states[0] if k != "radius"}
    eng.set_agents(x=S["x"], y=S["y"], theta=S["theta"], vel=S["vel"], w=S["w"], u=S["u"], collected=S["collected"],
                   collected_before=S["collected_before"], env_status=S["env_status"], override_mode=S["override"],
                   mode=S["mode"], patch_id=S["patch_id"], novelty=S["novelty"])
    eng.step(1, phases=4)     # collision phase only
    got = eng.get_agents()
    n_coll = 0
    for b in range(B):
        st = {k: (np.array(v, dtype=float) if k in ("theta", "vel") else np.array(v)) for k, v in states[b].items()}
        collided = rb.base_collision_phase(st, cfg, ghost_mode=ghost)
        n_coll += len(set(collided))
        np.testing.assert_allclose(got["theta"][b], st["theta"], rtol=RTOL, err_msg="theta")
        np.testing.assert_allclose(got["vel"][b], st["vel"], rtol=RTOL, atol=1e-6, err_msg="vel")
        assert np.array_equal(got["override_mode"][b], st["override"])
        assert np.array_equal(got["mode"][b], st["mode"])
        assert np.array_equal(got["env_status"][b], st["env_status"])
        assert np.array_equal(got["patch_id"][b], st["patch_id"])
    assert n_coll > 10
    eng.close()


def test_collision_phase_with_per_agent_max_exp_vel(built_lib):
    """agent2.velocity = agent2.max_exp_vel (sims.py:465) with one parameter set per agent."""
    import dataclasses
    from abm_b200 import BaseEngine
    rng = np.random.default_rng(77)
    B, N, W = 3, 40, 200.0
    cfg = rb.BaseConfig(R=1200, width=W, height=W, visual_exclusion=True)
    vmax = np.asarray(rng.uniform(1, 4, (B, N)), np.float32).astype(np.float64)
    eng = BaseEngine(B, N, 0, resolution=1200, width=W, height=W, visual_exclusion=True, collide_agents=True,
                     ghost_mode=False, tau=cfg.Tau)
    eng.set_params(exp_vel_max=vmax)
    states = []
    for b in range(B):
        st = _random_state(rng, N, W, cfg)
        st["override"] = rng.choice([0, 0, 1, 3], N)
        st["mode"] = st["override"].copy()
        states.append(st)
    S = {k: np.stack([s[k] for s in states]) for k in states[0] if k != "radius"}
    eng.set_agents(x=S["x"], y=S["y"], theta=S["theta"], vel=S["vel"], w=S["w"], u=S["u"], collected=S["collected"],
                   collected_before=S["collected_before"], env_status=S["env_status"], override_mode=S["override"],
                   mode=S["mode"], patch_id=S["patch_id"], novelty=S["novelty"])
    eng.step(1, phases=4)     # collision phase only
    got = eng.get_agents()
    n_set = 0
    for b in range(B):
        acfg = [dataclasses.replace(cfg, exp_vel_max=float(vmax[b, i])) for i in range(N)]
        st = {k: (np.array(v, dtype=float) if k in ("theta", "vel") else np.array(v)) for k, v in states[b].items()}
        rb.base_collision_phase(st, cfg, ghost_mode=False, agent_cfgs=acfg)
        np.testing.assert_allclose(got["theta"][b], st["theta"], rtol=RTOL, err_msg="theta")
        np.testing.assert_allclose(got["vel"][b], st["vel"], rtol=RTOL, atol=1e-6, err_msg="vel")
        n_set += int(np.isclose(st["vel"], vmax[b]).sum())
    assert n_set > 3          # some agents were sent straight on with their own max_exp_vel
    eng.close()


def test_full_loop_with_collisions(built_lib, step_path):
    """configs[2]-like shape (N=50, 3 patches, occlusion + collisions on) for 200 steps: stays finite,
    inside the arena, and agents end up less overlapped than without collision avoidance."""
    from abm_b200 import BaseEngine
    B, N, P, W = 16, 50, 3, 500.0
    rng = np.random.default_rng(3)
    x0, y0 = rng.integers(20, 520, (B, N)), rng.integers(20, 520, (B, N))
    th0 = rng.uniform(0, 2 * np.pi, (B, N))
    pa = dict(x=rng.integers(60, 400, (B, P)), y=rng.integers(60, 400, (B, P)), radius=np.full((B, P), 30.0),
              left=np.full((B, P), 200.0), quality=np.full((B, P), 0.25), id=np.tile(np.arange(P), (B, 1)))
    overlaps = {}
    for collide in (False, True):
        eng = BaseEngine(B, N, P, resolution=1200, width=W, height=W, visual_exclusion=True, collide_agents=collide,
                         ghost_mode=False, seed=9)
        eng.set_params(Eps_w=2.0, Eps_u=1.0, F_N=0.5, F_R=0.5, exp_vel_max=3.0, exp_theta_min=-0.5,
                       exp_theta_max=0.5, reloc_theta_max=1.8, exp_stop_ratio=0.175)
        eng.set_agents(x=x0, y=y0, theta=th0)
        eng.set_patches(**pa)
        eng.step(200)
        a = eng.get_agents()
        assert np.isfinite(a["x"]).all() and np.isfinite(a["theta"]).all()
        d = np.sqrt((a["x"][:, :, None] - a["x"][:, None, :]) ** 2 + (a["y"][:, :, None] - a["y"][:, None, :]) ** 2)
        overlaps[collide] = int(((d < 12.0).sum() - B * N) // 2)
        if step_path == "one_cta_per_replicate":
            assert eng.counters()["launches"] == 1                          # one launch for the whole run either way
        eng.close()
    print("deeply overlapping pairs without / with collision avoidance:", overlaps[False], overlaps[True])
    assert overlaps[True] <= overlaps[False]


def test_function_level_kat_b1(built_lib):
    """SURVEY App. B KAT-B1 through the function-level drop-ins: Agent.projection_field as
    abm_b200.supcalc.projection_field, supcalc.F_reloc_LR, vf_supcalc.dPhi_V_of."""
    from abm_b200 import supcalc, vf_supcalc
    pos = {0: (250, 250), 1: (400, 200), 2: (330, 230), 3: (250, 60), 4: (250, 150), 5: (60, 301)}
    social = [pos[1], pos[3], pos[5]]                  # exploiting on another patch
    occl = [pos[0], pos[2], pos[4]]                    # everyone else in range (self included: skipped)
    FOV = (-np.pi, np.pi)
    f = supcalc.projection_field(pos[0], 10, 0.3, 1200, FOV, social)
    assert rs.runs_of(f > 0) == [(1, 18), (348, 369), (584, 609), (1199, 1200)]
    np.testing.assert_allclose(supcalc.F_reloc_LR(1.0, f, v_desired=3.0, reloc_theta_max=1.8), (2.0, 0.132), rtol=1e-12)
    f = supcalc.projection_field(pos[0], 10, 0.3, 1200, FOV, social, non_expl_agents=occl, visual_exclusion=True)
    assert rs.runs_of(f > 0) == [(1, 18), (584, 588), (1199, 1200)]
    np.testing.assert_allclose(supcalc.F_reloc_LR(1.0, f, v_desired=3.0, reloc_theta_max=1.8), (2.0, 0.06), rtol=1e-12)
    # keep_distance_info: amplitude 1 - d_last / vision_range with the LAST obstacle of the list (agent.py:590)
    f = supcalc.projection_field(pos[0], 10, 0.3, 1200, FOV, social, keep_distance_info=True, vision_range=2000)
    d_last = np.hypot(60 - 250, 301 - 250)
    assert np.isclose(f.max(), 1 - d_last / 2000)
    for v, e in [([1, 1, 1, 0, 0, 0, 0, 0], [1, 0, 0, -1, 0, 0, 0, 0]), ([0, 0, 1, 1, 1, 0, 0, 0], [0, 1, 0, 0, -1, 0, 0, 0]),
                 ([0, 0, 0, 0, 0, 0, 1, 1], [0, 0, 0, 0, 0, 1, 0, -1])]:
        assert vf_supcalc.dPhi_V_of(np.zeros(8), np.array(v, float)).tolist() == e


def test_base_recorder_writes_reference_zarr_layout(built_lib, tmp_path):
    """SURVEY f1 for the foraging engine: ag_*.zarr / res_*.zarr in the reference's layout (ifdb.py:437-508)."""
    import json
    import os
    from abm_b200 import BaseEngine
    from abm_b200.recorder import BaseRecorder, read_zarr_v2
    B, N, P, W, T = 3, 12, 2, 300.0, 7
    rng = np.random.default_rng(4)
    eng = BaseEngine(B, N, P, resolution=320, width=W, height=W, seed=2)
    eng.set_params()
    eng.set_agents(x=rng.integers(30, 300, (B, N)), y=rng.integers(30, 300, (B, N)), theta=rng.uniform(0, 6.28, (B, N)))
    eng.set_patches(x=rng.integers(60, 200, (B, P)), y=rng.integers(60, 200, (B, P)), radius=np.full((B, P), 30.0),
                    left=np.full((B, P), 50.0), quality=np.full((B, P), 0.25), id=np.tile(np.arange(P), (B, 1)))
    rec = BaseRecorder(eng, str(tmp_path / "run"), replicates=[1], env_params={"N": N})
    want = []
    for _ in range(T):
        eng.step(1); rec.record(); want.append((eng.get_agents(), eng.get_patches()))
    (d,) = rec.close()
    assert json.load(open(os.path.join(d, "env_params.json")))["N"] == N
    for name, key, trunc in (("posx", "x", True), ("ori", "theta", False), ("w", "w", False), ("mode", "mode", False),
                             ("collr", "collected", False), ("explr", "patch_id", False)):
        got = read_zarr_v2(os.path.join(d, f"ag_{name}.zarr"))
        exp = np.stack([w[0][key][1].astype(np.float64) for w in want], axis=1)
        exp = np.roll(exp, -1, axis=0)            # the reference's row order: agent id - 1, agent 0 in the last row (ifdb.py:504-508)
        assert got.shape == (N, T) and np.array_equal(got, np.trunc(exp) if trunc else exp), name
    got = read_zarr_v2(os.path.join(d, "res_left.zarr"))
    assert got.shape == (P, T) and np.array_equal(got, np.stack([w[1]["left"][1].astype(np.float64) for w in want], axis=1))
    eng.close()


@pytest.mark.parametrize("teleport,ghost,hetero", [(True, False, False), (False, True, False), (True, False, True)])
def test_full_step_in_lockstep_with_oracle(built_lib, teleport, ghost, hetero):
    """The whole loop body (sims.py:733-864) as ONE engine step -- collisions, agent-patch interaction, Agent.update of
    every agent from the snapshot the environment phase leaves -- against the oracle's three phases chained in the
    reference's order, for 10 consecutive steps in lockstep (every step starts from the engine's own state, so nothing
    drifts): what one phase hands to the next (turned headings and collide modes, teleports, notifications, the collided
    set, depleted patches) is compared as well as the phases themselves.  hetero: every agent with its own radius, field
    resolution, FOV, vision range, speed and consumption (agent_behave_param_list, sims.py:499-517)."""
    import dataclasses
    from abm_b200 import BaseEngine
    rng = np.random.default_rng(31)
    B, N, P, W = 2, 36, 3, 240.0
    cfg = rb.BaseConfig(R=1200, width=W, height=W, visual_exclusion=True, teleport_exploit=teleport, Eps_w=2.0, Eps_u=1.0,
                        exp_vel_max=3.0, exp_theta_min=-0.5, exp_theta_max=0.5, reloc_theta_max=1.8, exp_stop_ratio=0.175,
                        F_N=0.5, F_R=0.5, agent_consumption=1.0)
    eng = _engine_for(cfg, B, N, P, collide_agents=True, ghost_mode=ghost, regenerate_patches=False)
    radii = np.full((B, N), 10.0)
    acfgs = [None] * B
    if hetero:
        radii = rng.choice([6.0, 10.0, 13.0], (B, N)); res = rng.choice([1200, 800, 333], (B, N))
        fov = rng.choice([1.0, 0.75, 0.5], (B, N)); vr = rng.choice([80.0, 200.0, 2000.0], (B, N))
        vmax = np.asarray(rng.uniform(1, 4, (B, N)), np.float32).astype(np.float64); cons = rng.choice([0.25, 1.0, 2.0], (B, N))
        names = ("T_w", "Eps_w", "g_w", "B_w", "w_max", "T_u", "Eps_u", "g_u", "B_u", "u_max", "S_wu", "S_uw", "F_N", "F_R",
                 "exp_theta_min", "exp_theta_max", "reloc_theta_max", "exp_stop_ratio")
        eng.set_params(exp_vel_max=vmax, agent_consumption=cons, **{k: getattr(cfg, k) for k in names})
        eng.set_agent_radii(radii); eng.set_agent_resolution(res); eng.set_agent_geometry(agent_fov=fov, vision_range=vr)
        acfgs = [[dataclasses.replace(cfg, R=int(res[b, i]), fov=(-fov[b, i] * np.pi, fov[b, i] * np.pi),
                                      vision_range=float(vr[b, i]), exp_vel_max=float(vmax[b, i]),
                                      agent_consumption=float(cons[b, i])) for i in range(N)] for b in range(B)]
    x0, y0 = rng.integers(30, 250, (B, N)).astype(float), rng.integers(30, 250, (B, N)).astype(float)
    x0[:, :6] = 40.0 + 32.0 - 10.0 + rng.integers(-12, 12, (B, 6)); y0[:, :6] = 50.0 + 32.0 - 10.0 + rng.integers(-12, 12, (B, 6))
    x0[:, 6:10] = 150.0 + 22.0 + rng.integers(-10, 10, (B, 4)); y0[:, 6:10] = 150.0 + 22.0 + rng.integers(-10, 10, (B, 4))
    on = np.arange(N) < 10                                              # the agents standing on patches 1 and 2 exploit them
    u0 = np.where(on, 0.9, 0.0) * np.ones((B, 1))
    ov0 = np.where(on, 1, 0) * np.ones((B, 1), int)
    pid0 = np.where(np.arange(N) < 6, 1, np.where(on, 2, -1)) * np.ones((B, 1), int)
    eng.set_agents(x=x0, y=y0, theta=rng.uniform(0, 2 * np.pi, (B, N)), u=u0, override_mode=ov0, mode=ov0,
                   env_status=np.where(on, 1, -1) * np.ones((B, 1), int), patch_id=pid0)
    eng.set_patches(x=np.tile([40.0, 150.0, 60.0], (B, 1)), y=np.tile([50.0, 150.0, 170.0], (B, 1)),
                    radius=np.full((B, P), 32.0), left=np.tile([3.0, 400.0, 400.0], (B, 1)),
                    quality=np.full((B, P), 0.5), id=np.tile(np.arange(1, P + 1), (B, 1)))
    n_collided = n_exploit = n_depleted = 0
    for step in range(10):
        a0, p0 = eng.get_agents(), eng.get_patches()
        dth = np.asarray(rng.uniform(-0.5, 0.5, (B, N)), np.float32)
        eng.step(1, inject_dtheta=dth)
        got, gp, fields = eng.get_agents(), eng.get_patches(), eng.fields()
        for b in range(B):
            st = dict(x=a0["x"][b].astype(float), y=a0["y"][b].astype(float), theta=a0["theta"][b].astype(float),
                      vel=a0["vel"][b].astype(float), w=a0["w"][b].astype(float), u=a0["u"][b].astype(float),
                      collected=a0["collected"][b].astype(float), collected_before=a0["collected_before"][b].astype(float),
                      env_status=a0["env_status"][b].copy(), override=a0["override_mode"][b].copy(), mode=a0["mode"][b].copy(),
                      patch_id=a0["patch_id"][b].copy(), radius=radii[b].copy() if hetero else 10.0,
                      novelty=((a0["novelty"][b][:, None] >> np.arange(cfg.Tau)) & 1).astype(float))
            pa = {k: np.array(p0[k][b], dtype=float if k != "id" else int) for k in p0}
            collided = rb.base_collision_phase(st, cfg, ghost, agent_cfgs=acfgs[b])    # sims.py:736-783
            depleted = rb.base_patch_phase(st, pa, cfg, collided=set(collided), agent_cfgs=acfgs[b])   # :790-858
            ref = rb.base_step_frozen(st, cfg, dth[b].astype(np.float64), agent_cfgs=acfgs[b])         # :861
            n_collided += len(set(collided)); n_depleted += len(depleted); n_exploit += int((st["override"] == 1).sum())
            assert np.array_equal(fields[b], ref["fields"]), (step, b)
            _compare_agents(got, ref, b)
            np.testing.assert_allclose(got["collected"][b], st["collected"], rtol=RTOL, atol=1e-6)
            assert np.array_equal(got["env_status"][b], st["env_status"]) and np.array_equal(got["patch_id"][b], st["patch_id"])
            nov = ((got["novelty"][b][:, None] >> np.arange(cfg.Tau)) & 1).astype(float)
            assert np.array_equal(nov, st["novelty"])
            np.testing.assert_allclose(gp["left"][b], pa["left"], rtol=RTOL, atol=1e-6)
            for p_ in depleted:
                assert gp["radius"][b][p_] == 0.0
    assert n_collided > 20 and n_exploit > 5 and n_depleted >= 1
    eng.close()


def test_step_path_follows_the_batch_shape(built_lib, monkeypatch, step_path):
    """Without ABM_BASE_FUSED the engine picks the step path from the batch shape: small replicates and batches that fill
    the GPU with one CTA per replicate take the fused kernel, few larger replicates one grid per phase."""
    if step_path != "one_cta_per_replicate":
        pytest.skip("independent of the forced path")
    from abm_b200 import BaseEngine
    monkeypatch.delenv("ABM_BASE_FUSED", raising=False)
    rng = np.random.default_rng(2)
    for B, N, fused in [(1, 10, True), (40, 25, True), (8, 50, False), (2, 100, False), (700, 50, True)]:
        eng = BaseEngine(B, N, 2, resolution=1200, width=500.0, height=500.0, visual_exclusion=True, seed=1)
        eng.set_params(Eps_w=2.0, Eps_u=1.0)
        eng.set_agents(x=rng.integers(20, 520, (B, N)), y=rng.integers(20, 520, (B, N)), theta=rng.uniform(0, 6, (B, N)))
        eng.set_patches(x=rng.integers(60, 400, (B, 2)), y=rng.integers(60, 400, (B, 2)), radius=np.full((B, 2), 30.0),
                        left=np.full((B, 2), 50.0), quality=np.full((B, 2), 0.25), id=np.tile(np.arange(2), (B, 1)))
        eng.step(4)
        n = eng.counters()["launches"]
        assert n == ((1 if B <= 148 else 4) if fused else 8), (B, N, n)      # env + agents: two grids per step when not fused
        eng.close()


@pytest.mark.parametrize("N", [50, 100])
def test_fused_step_equals_separate_phases(built_lib, monkeypatch, N):
    """The one-launch step (a CTA per replicate runs collisions, agent-patch interaction and Agent.update in the
    reference's order, sims.py:733-864) against one grid per phase: identical trajectories, bit for bit, over 150 steps
    of a config-3-like sweep (occlusion, collisions, depletion and regeneration, one Eps_w per replicate); N = 100: the
    replicate size of the reference's figure experiments (objects in two mask words)."""
    from abm_b200 import BaseEngine
    B, P, W = 24, 3, 500.0
    rng = np.random.default_rng(17)
    x0, y0 = rng.integers(20, 520, (B, N)), rng.integers(20, 520, (B, N))
    th0 = rng.uniform(0, 2 * np.pi, (B, N))
    pa = dict(x=rng.integers(60, 400, (B, P)), y=rng.integers(60, 400, (B, P)), radius=np.full((B, P), 30.0),
              left=np.full((B, P), 40.0), quality=np.full((B, P), 0.25), id=np.tile(np.arange(P), (B, 1)))
    res = {}
    for separate in (False, True):
        monkeypatch.setenv("ABM_BASE_FUSED", "0" if separate else "1")
        eng = BaseEngine(B, N, P, resolution=1200, width=W, height=W, visual_exclusion=True, collide_agents=True,
                         ghost_mode=False, min_resc_perpatch=30, max_resc_perpatch=50, seed=5, keep_fields=True)
        eng.set_params(Eps_w=np.linspace(0, 5, B), Eps_u=1.0, F_N=0.5, F_R=0.5, exp_vel_max=3.0, exp_theta_min=-0.5,
                       exp_theta_max=0.5, reloc_theta_max=1.8, exp_stop_ratio=0.175)
        eng.set_agents(x=x0, y=y0, theta=th0)
        eng.set_patches(**pa)
        eng.step(150)
        res[separate] = (eng.get_agents(), eng.get_patches(), eng.fields(), eng.counters())
        eng.close()
    assert res[False][3]["launches"] == 1 and res[True][3]["launches"] == 450
    assert res[False][3]["patches_regenerated"] == res[True][3]["patches_regenerated"] > 0
    for k, v in res[False][0].items():
        assert np.array_equal(v, res[True][0][k]), k
    for k, v in res[False][1].items():
        assert np.array_equal(v, res[True][1][k]), k
    assert np.array_equal(res[False][2], res[True][2])


@pytest.mark.parametrize("border_overlap", [False, True])
def test_patch_regeneration_matches_oracle(built_lib, border_overlap):
    """kill_resource / add_new_resource_patch (sims.py:321-374) in the environment phase with INJECTED draws
    (abm_base_inject_regeneration): patches are exhausted by their exploiters in this step, the prepared tries first
    land on other patches (retry loop), and position / units / quality / id of every regenerated patch as well as the
    number of regenerations equal the oracle's (oracle/restate_base.base_regenerate_patch, itself pinned to the live
    reference in tests/test_oracle_base.py)."""
    rng = np.random.default_rng(33)
    B, N, P, W = 6, 24, 4, 400.0
    R, pad = 30.0, 30.0
    cfg = rb.BaseConfig(R=600, width=W, height=W, agent_consumption=1.0)
    states, patches = [], []
    T = 10
    draws = np.zeros((B, P, T, 4))
    for b in range(B):
        pos = []
        while len(pos) < P:
            x, y = int(rng.integers(pad, W + pad - 2 * R)), int(rng.integers(pad, W + pad - 2 * R))
            if all((x - a_) ** 2 + (y - b_) ** 2 > (2 * R) ** 2 for a_, b_ in pos):
                pos.append((x, y))
        pa = dict(x=np.array([q[0] for q in pos], float), y=np.array([q[1] for q in pos], float),
                  radius=np.full(P, R), left=rng.choice([0.5, 1.0, 40.0], P).astype(float), quality=np.full(P, 0.75),
                  id=np.arange(1, P + 1))
        st = _random_state(rng, N, W, cfg)
        for i in range(N):                         # two exploiting agents on every patch: the scarce ones die this step
            if i < 2 * P:
                q = i // 2
                st["x"][i], st["y"][i] = pos[q][0] + R - 10 + (i % 2) * 6, pos[q][1] + R - 10
                st["override"][i], st["mode"][i] = 1, 1
        for p in range(P):
            for t in range(T):
                if t < 2 and rng.uniform() < 0.8:  # lands on another patch: rejected
                    q = int(rng.choice([k for k in range(P) if k != p]))
                    draws[b, p, t, 0] = pos[q][0] + int(rng.integers(-20, 20))
                    draws[b, p, t, 1] = pos[q][1] + int(rng.integers(-20, 20))
                else:
                    lo = pad - R if border_overlap else pad
                    draws[b, p, t, 0] = int(rng.integers(lo, W + pad - 2 * R))
                    draws[b, p, t, 1] = int(rng.integers(lo, W + pad - 2 * R))
                draws[b, p, t, 2] = int(rng.integers(20, 60)); draws[b, p, t, 3] = np.float32(rng.uniform(0.1, 1.0))
        states.append(st); patches.append(pa)
    eng = _engine_for(cfg, B, N, P, regenerate_patches=True, patch_border_overlap=border_overlap, patch_radius=R)
    stacked = {k: np.stack([s[k] for s in states]) for k in states[0] if k != "radius"}
    eng.set_agents(x=stacked["x"], y=stacked["y"], theta=stacked["theta"], vel=stacked["vel"], w=stacked["w"],
                   u=stacked["u"], collected=stacked["collected"], collected_before=stacked["collected_before"],
                   env_status=stacked["env_status"], override_mode=stacked["override"], mode=stacked["mode"],
                   patch_id=stacked["patch_id"], novelty=stacked["novelty"])
    eng.set_patches(**{k: np.stack([p_[k] for p_ in patches]) for k in patches[0]})
    eng.inject_regeneration(draws)
    eng.step(1, phases=PHASE_ENV)
    gp, cnt = eng.get_patches(), eng.counters()
    n_regen = 0
    for b in range(B):
        st = {k: (np.array(v, dtype=float) if k in ("theta", "collected", "collected_before") else np.array(v))
              for k, v in states[b].items()}
        pa = {k: np.array(v, dtype=float if k != "id" else int) for k, v in patches[b].items()}
        # the reference regenerates a patch the moment it is exhausted, inside the patch loop (sims.py:829-836)
        for p in rb.base_patch_phase(st, pa, cfg):
            assert rb.base_regenerate_patch(pa, p, draws[b, p], R) >= 1
            n_regen += 1
        for k in ("x", "y", "radius", "left", "quality"):
            np.testing.assert_allclose(gp[k][b], pa[k], rtol=1e-6, err_msg=f"{k} replicate {b}")
        assert np.array_equal(gp["id"][b], pa["id"])
    assert n_regen >= 4 and cnt["patches_regenerated"] == n_regen and cnt["regeneration_failed"] == 0
    eng.inject_regeneration(None)
    eng.close()


@pytest.mark.parametrize("case", load_base_hetero_cases("base_hetero_radius_golden.npz"),
                         ids=lambda c: f"N{len(c['dth'])}_R{c['cfg'].R}")
def test_agent_phase_matches_reference_fixture_heterogeneous_radii(built_lib, case):
    """agent_radius of agent_behave_param_list (sims.py:502) in the kernels (abm_base_set_agent_radii): every agent's own
    radius in the candidate distance and at the walls, the FOCAL radius for both centres and the projected size
    (agent.py:504-509, 529) -- against the fixture the unmodified reference's constructor path produced."""
    cfg, st = case["cfg"], case["st"]
    N = len(case["dth"])
    eng = _engine_for(cfg, 1, N)
    geo = ("agent_fov", "vision_range", "agent_radius")
    eng.set_params(exp_theta_min=cfg.exp_theta_min, exp_theta_max=cfg.exp_theta_max, reloc_theta_max=cfg.reloc_theta_max,
                   **{k: v[None] for k, v in case["agent_params"].items() if k not in geo})
    eng.set_agent_geometry(agent_fov=case["agent_params"]["agent_fov"][None],
                           vision_range=case["agent_params"]["vision_range"][None])
    eng.set_agent_radii(np.asarray(st["radius"])[None])
    _upload(eng, st)
    eng.step(1, inject_dtheta=case["dth"], phases=PHASE_AGENTS)
    assert np.array_equal(rs.pack_bits(eng.fields()[0]), case["fields"])          # bit-exact stored fields
    _compare_agents(eng.get_agents(), case["out"])
    eng.set_agent_radii(None)                                                     # back to one radius: different fields
    _upload(eng, st)
    eng.step(1, inject_dtheta=case["dth"], phases=PHASE_AGENTS)
    assert not np.array_equal(rs.pack_bits(eng.fields()[0]), case["fields"])
    eng.close()


@pytest.mark.parametrize("case", load_base_hetero_cases("base_hetero_res_golden.npz"),
                         ids=lambda c: f"N{len(c['dth'])}_R{c['cfg'].R}")
def test_agent_phase_matches_reference_fixture_heterogeneous_resolutions(built_lib, case):
    """v_field_res of agent_behave_param_list (sims.py:507) in the kernels (abm_base_set_agent_resolution): every agent's
    own linspace grid, projected size, wrap-around and field means (agent.py:58, 480-481, 543, 577-588;
    supcalc.py:86-91) -- against the fixture the unmodified reference's constructor path produced, in the fused
    kernel and in the one-grid-per-phase kernels."""
    cfg, st = case["cfg"], case["st"]
    N = len(case["dth"])
    eng = _engine_for(cfg, 1, N)
    geo = ("agent_fov", "vision_range", "v_field_res")
    eng.set_params(exp_theta_min=cfg.exp_theta_min, exp_theta_max=cfg.exp_theta_max, reloc_theta_max=cfg.reloc_theta_max,
                   **{k: v[None] for k, v in case["agent_params"].items() if k not in geo})
    eng.set_agent_geometry(agent_fov=case["agent_params"]["agent_fov"][None],
                           vision_range=case["agent_params"]["vision_range"][None])
    res = case["agent_params"]["v_field_res"].astype(int)
    eng.set_agent_resolution(res[None])
    _upload(eng, st)
    eng.step(1, inject_dtheta=case["dth"], phases=PHASE_AGENTS)
    fields = eng.fields()[0]
    assert np.array_equal(rs.pack_bits(fields), case["fields"])                   # bit-exact stored fields
    assert all(not fields[i, r:].any() for i, r in enumerate(res))
    _compare_agents(eng.get_agents(), case["out"])
    eng.set_agent_resolution(None)                                                # back to one resolution: different fields
    _upload(eng, st)
    eng.step(1, inject_dtheta=case["dth"], phases=PHASE_AGENTS)
    assert not np.array_equal(rs.pack_bits(eng.fields()[0]), case["fields"])
    with pytest.raises(Exception):
        eng.set_agent_resolution(np.full((1, N), cfg.R + 1))                      # beyond the row stride
    eng.close()


def test_collision_lidar_uses_the_hit_agents_own_resolution(built_lib):
    """sims.py:449-462: the proximity field of the hit agent is ITS projection_field (own v_field_res): left / right
    means and the frontal +-100 bins -- against the oracle with one BaseConfig per agent."""
    import dataclasses
    rng = np.random.default_rng(77)
    B, N, W = 2, 40, 260.0
    cfg = rb.BaseConfig(R=1200, width=W, height=W, visual_exclusion=True, teleport_exploit=True, exp_vel_max=2.0)
    res = rng.choice([1200, 700, 333, 150], (B, N))
    states = []
    for b in range(B):
        st = _random_state(rng, N, W, cfg)
        st["override"] = rng.choice([0, 0, 0, 1], N); st["mode"] = np.where(st["override"] == 1, 1, 0)
        states.append(st)
    eng = _engine_for(cfg, B, N, 0, collide_agents=True, ghost_mode=False, regenerate_patches=False)
    eng.set_agent_resolution(res)
    stacked = {k: np.stack([s[k] for s in states]) for k in states[0] if k != "radius"}
    eng.set_agents(x=stacked["x"], y=stacked["y"], theta=stacked["theta"], vel=stacked["vel"], w=stacked["w"],
                   u=stacked["u"], collected=stacked["collected"], collected_before=stacked["collected_before"],
                   env_status=stacked["env_status"], override_mode=stacked["override"], mode=stacked["mode"],
                   patch_id=stacked["patch_id"], novelty=stacked["novelty"])
    eng.step(1, phases=4)                                                         # the collision phase alone
    got = eng.get_agents()
    n_coll, n_differs = 0, 0
    for b in range(B):
        def run(agent_cfgs):
            st = {k: (np.array(v, dtype=float) if k in ("theta", "vel", "x", "y") else np.array(v)) for k, v in states[b].items()}
            collided = rb.base_collision_phase(st, cfg, False, agent_cfgs=agent_cfgs)
            return st, collided
        st, collided = run([dataclasses.replace(cfg, R=int(r)) for r in res[b]])
        st_same, _ = run(None)
        n_coll += len(set(collided))
        n_differs += int(np.sum(st["theta"] != st_same["theta"]) + np.sum(st["vel"] != st_same["vel"]))
        np.testing.assert_allclose(got["theta"][b], st["theta"], rtol=RTOL, atol=1e-5)
        np.testing.assert_allclose(got["vel"][b], st["vel"], rtol=RTOL, atol=1e-5)
        assert np.array_equal(got["override_mode"][b], st["override"].astype(int))
    assert n_coll > 6 and n_differs > 0          # ... and one shared resolution gives other turns / stops
    eng.close()


@pytest.mark.parametrize("ghost,vis_excl", [(True, False), (False, True)])
def test_env_and_collision_phases_with_per_agent_radii(built_lib, ghost, vis_excl):
    """The agents' own radii in the patch membership / bias / teleport (sims.py:45-56, 544-552, 820-821) and in the
    collision circles, the vicinity test and the LIDAR field of the hit agent (sims.py:421-468, 739-752) against the
    oracle (the pair detection restates pygame's documented collide_circle: parity unpinned)."""
    rng = np.random.default_rng(55)
    B, N, P, W = 3, 40, 3, 300.0
    cfg = rb.BaseConfig(R=1200, width=W, height=W, visual_exclusion=vis_excl, teleport_exploit=True, exp_vel_max=2.0)
    radii = rng.choice([6.0, 10.0, 14.0], (B, N))
    states, patches = [], []
    for b in range(B):
        st = _random_state(rng, N, W, cfg)
        st["override"] = rng.choice([0, 0, 0, 1], N); st["mode"] = np.where(st["override"] == 1, 1, 0)
        st["radius"] = radii[b].copy()
        states.append(st)
        patches.append(dict(x=rng.integers(40, 250, P).astype(float), y=rng.integers(40, 250, P).astype(float),
                            radius=np.full(P, 35.0), left=np.full(P, 100.0), quality=np.full(P, 0.5),
                            id=np.arange(1, P + 1)))
    eng = _engine_for(cfg, B, N, P, collide_agents=True, ghost_mode=ghost, regenerate_patches=False)
    eng.set_agent_radii(radii)
    stacked = {k: np.stack([s[k] for s in states]) for k in states[0] if k != "radius"}
    eng.set_agents(x=stacked["x"], y=stacked["y"], theta=stacked["theta"], vel=stacked["vel"], w=stacked["w"],
                   u=stacked["u"], collected=stacked["collected"], collected_before=stacked["collected_before"],
                   env_status=stacked["env_status"], override_mode=stacked["override"], mode=stacked["mode"],
                   patch_id=stacked["patch_id"], novelty=stacked["novelty"])
    eng.set_patches(**{k: np.stack([p_[k] for p_ in patches]) for k in patches[0]})
    eng.step(1, phases=4 | PHASE_ENV)                       # collisions, then the environment phase (sims.py order)
    got = eng.get_agents()
    n_coll = 0
    for b in range(B):
        st = {k: (np.array(v, dtype=float) if k in ("theta", "vel", "collected", "collected_before", "x", "y") else np.array(v))
              for k, v in states[b].items()}
        pa = {k: np.array(v, dtype=float if k != "id" else int) for k, v in patches[b].items()}
        collided = rb.base_collision_phase(st, cfg, ghost)
        n_coll += len(set(collided))
        rb.base_patch_phase(st, pa, cfg, collided=set(collided))
        for k, g in dict(x="x", y="y", theta="theta", vel="vel").items():
            np.testing.assert_allclose(got[g][b], st[k], rtol=RTOL, atol=1e-5, err_msg=k)
        assert np.array_equal(got["override_mode"][b], st["override"].astype(int))
        assert np.array_equal(got["env_status"][b], st["env_status"])
        assert np.array_equal(got["patch_id"][b], st["patch_id"])
    assert n_coll > 6
    eng.close()
