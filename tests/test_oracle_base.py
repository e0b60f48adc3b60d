"""CPU tests: the BASE oracle (oracle/restate_base.py) against SURVEY KAT-B1, against fixtures
produced by executing the unmodified reference's Agent.update (tests/golden/base_golden.npz)
and, when /root/reference is mounted, the reference's own notify / deplete / bias functions."""
import dataclasses

import numpy as np
import pytest

from golden_io import BASE_OUT_KEYS, load_base_cases, load_base_hetero_cases
from oracle import ref_shim
from oracle import restate as rs
from oracle import restate_base as rb


@pytest.mark.parametrize("vis_excl,runs,mean,reloc", [
    (False, [(1, 18), (348, 369), (584, 609), (1199, 1200)], 0.053333, (2.0, 0.132)),
    (True, [(1, 18), (584, 588), (1199, 1200)], 0.018333, (2.0, 0.06)),
])
def test_kat_b1(vis_excl, runs, mean, reloc):
    x = np.array([250., 400, 330, 250, 250, 60]); y = np.array([250., 200, 230, 60, 150, 301])
    th = np.array([0.3, 0, 0, 0, 0, 0])
    ex = np.array([0, 1, 0, 1, 0, 1], bool); pid = np.array([-1, 1, -1, 1, -1, 1])
    cfg = rb.BaseConfig(R=1200, vision_range=2000, visual_exclusion=vis_excl, patchwise_exclusion=True,
                        reloc_theta_max=1.8)
    f, src = rb.base_field(0, x, y, 10.0, th, ex, pid, cfg)
    assert rs.runs_of(f) == runs
    assert round(float(f.mean()), 6) == mean
    np.testing.assert_allclose(rb.relocation_force(1.0, f, 3.0, cfg), reloc, rtol=1e-12)
    if vis_excl:   # source data (k, s, e, s_ex, e_ex, d) of SURVEY App. B
        by_j = {o["j"]: o for o in src}
        assert (by_j[1]["k"], by_j[1]["s"], by_j[1]["e"], by_j[1]["sx"], by_j[1]["ex"]) == (604, 591, 616, 612, 616)
        assert (by_j[3]["sx"], by_j[3]["ex"]) == (0, 0)
        assert (by_j[5]["s"], by_j[5]["e"], by_j[5]["sx"], by_j[5]["ex"]) == (1182, 1201, 1182, 1201)
        assert round(by_j[5]["d"], 3) == 196.726


def test_dead_ahead_object_is_invisible():
    """agent.py:520-523: closed angle 0 maps to 2pi."""
    x = np.array([100.0, 200.0]); y = np.array([100.0, 100.0])
    cfg = rb.BaseConfig(R=1200, visual_exclusion=False)
    f, _ = rb.base_field(0, x, y, 10.0, np.zeros(2), np.array([False, True]), np.array([-1, 2]), cfg)
    assert not f.any()
    f, _ = rb.base_field(0, x, np.array([100.0, 100.0001]), 10.0, np.zeros(2), np.array([False, True]),
                         np.array([-1, 2]), cfg)
    assert f.sum() > 30


@pytest.mark.parametrize("case", load_base_cases(), ids=lambda c: f"N{len(c['dth'])}_R{c['cfg'].R}")
def test_restatement_matches_reference_fixture(case):
    out = rb.base_step_frozen(case["st"], case["cfg"], case["dth"])
    assert np.array_equal(rs.pack_bits(out["fields"]), case["fields"])
    for k in BASE_OUT_KEYS:
        np.testing.assert_allclose(out[k], case["out"][k], rtol=1e-12, atol=1e-12, err_msg=k)


@pytest.mark.parametrize("case", load_base_hetero_cases(), ids=lambda c: f"N{len(c['dth'])}_R{c['cfg'].R}")
def test_restatement_matches_reference_fixture_heterogeneous_agents(case):
    """agent_behave_param_list (sims.py:499-517): per-agent decision parameters, max_exp_vel, exp_stop_ratio taken
    by the reference's own constructor (agent.py:83-108)."""
    out = rb.base_step_frozen(case["st"], case["cfg"], case["dth"], agent_cfgs=case["agent_cfgs"])
    assert np.array_equal(rs.pack_bits(out["fields"]), case["fields"])
    for k in BASE_OUT_KEYS:
        np.testing.assert_allclose(out[k], case["out"][k], rtol=1e-12, atol=1e-12, err_msg=k)
    # the fixture must not be reproducible with one shared parameter set
    shared = rb.base_step_frozen(case["st"], case["agent_cfgs"][0], case["dth"])
    assert not np.allclose(shared["w"], case["out"]["w"]) and not np.allclose(shared["vel"], case["out"]["vel"])


@pytest.mark.parametrize("case", load_base_hetero_cases("base_hetero_radius_golden.npz"),
                         ids=lambda c: f"N{len(c['dth'])}_R{c['cfg'].R}")
def test_restatement_matches_reference_fixture_heterogeneous_radii(case):
    """agent_radius of agent_behave_param_list (sims.py:502): every agent's own radius in the candidate distance
    (supcalc.py:73-78) and in its wall reflection, the FOCAL radius for both centres and the projection size
    (agent.py:504-509, 529).  The same fixture pins the CUDA path (abm_base_set_agent_radii, tests/test_base_gpu.py)."""
    assert np.ndim(case["st"]["radius"]) == 1 and len(set(case["st"]["radius"])) > 1
    out = rb.base_step_frozen(case["st"], case["cfg"], case["dth"], agent_cfgs=case["agent_cfgs"])
    assert np.array_equal(rs.pack_bits(out["fields"]), case["fields"])
    for k in BASE_OUT_KEYS:
        np.testing.assert_allclose(out[k], case["out"][k], rtol=1e-12, atol=1e-12, err_msg=k)
    # one shared radius does not reproduce it
    shared = rb.base_step_frozen(dict(case["st"], radius=10.0), case["cfg"], case["dth"], agent_cfgs=case["agent_cfgs"])
    assert not np.array_equal(rs.pack_bits(shared["fields"]), case["fields"])


@pytest.mark.parametrize("case", load_base_hetero_cases("base_hetero_res_golden.npz"),
                         ids=lambda c: f"N{len(c['dth'])}_R{c['cfg'].R}")
def test_restatement_matches_reference_fixture_heterogeneous_resolutions(case):
    """v_field_res of agent_behave_param_list (sims.py:507): every agent's own linspace grid, projection size, wrap and
    field means (agent.py:58, 480-481, 543, 577-588; supcalc.py:86-91), fields padded with zeros up to cfg.R."""
    res = [c.R for c in case["agent_cfgs"]]
    assert len(set(res)) > 2 and max(res) == case["cfg"].R
    out = rb.base_step_frozen(case["st"], case["cfg"], case["dth"], agent_cfgs=case["agent_cfgs"])
    assert np.array_equal(rs.pack_bits(out["fields"]), case["fields"])
    for k in BASE_OUT_KEYS:
        np.testing.assert_allclose(out[k], case["out"][k], rtol=1e-12, atol=1e-12, err_msg=k)
    for i, r in enumerate(res):
        assert not out["fields"][i, r:].any()
    # one shared resolution does not reproduce it
    same = [dataclasses.replace(c, R=case["cfg"].R) for c in case["agent_cfgs"]]
    shared = rb.base_step_frozen(case["st"], case["cfg"], case["dth"], agent_cfgs=same)
    assert not np.array_equal(rs.pack_bits(shared["fields"]), case["fields"])


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_patch_phase_pieces_match_live_reference():
    """notify_agent (sims.py:29-42), Rescource.deplete (rescource.py:118-133) and
    bias_agent_towards_res_center (sims.py:544-552) of the real reference against the oracle's
    patch phase on a one-patch scene."""
    ref_shim.install()
    import types
    from abm.simulation import sims
    from abm.environment.rescource import Rescource
    rng = np.random.default_rng(4)
    N = 12
    cfg = rb.BaseConfig(R=320, width=300, height=300, agent_consumption=1.0)
    st = dict(x=rng.integers(100, 160, N).astype(float), y=rng.integers(100, 160, N).astype(float),
              theta=rng.uniform(0, 2 * np.pi, N), vel=np.zeros(N), radius=10.0, w=np.zeros(N), u=np.zeros(N),
              novelty=np.zeros((N, 10)), env_status=rng.choice([-1, 1], N), override=rng.choice([0, 1], N),
              mode=np.zeros(N, int), patch_id=np.full(N, -1), collected=np.zeros(N), collected_before=np.zeros(N))
    patches = dict(x=np.array([100.0]), y=np.array([100.0]), radius=np.array([30.0]), left=np.array([3.0]),
                   quality=np.array([0.75]), id=np.array([7]))
    agents = ref_shim.make_base_agents(st, cfg)
    res = Rescource(7, 30, (100.0, 100.0), (300, 300), (0, 0, 0), 30, 3.0, 0.75)
    # the loop body of sims.py:805-844 driven with the reference's own functions
    members = [a for a in agents if sims.supcalc.distance(res, a) < res.radius]
    destroy = 0
    dummy = types.SimpleNamespace()
    for a in members:
        sims.Simulation.bias_agent_towards_res_center(dummy, a, res)
        if destroy:
            sims.notify_agent(a, -1)
        else:
            sims.notify_agent(a, 1, res.id)
            if a.get_mode() == "exploit":
                depl, destroy = res.deplete(a.consumption)
                a.collected_r_before = a.collected_r
                a.collected_r += depl
                if destroy:
                    for a2 in members:
                        sims.notify_agent(a2, -1)
    for a in agents:
        if a not in members:
            sims.notify_agent(a, -1)
    depleted = rb.base_patch_phase(st, patches, cfg)
    assert depleted == ([0] if destroy else [])
    assert np.isclose(patches["left"][0], res.resc_left)
    for i, a in enumerate(agents):
        assert np.isclose(st["theta"][i], a.orientation, rtol=1e-14)
        assert st["env_status"][i] == a.env_status and st["patch_id"][i] == a.exploited_patch_id
        assert np.array_equal(st["novelty"][i], a.novelty)
        assert np.isclose(st["collected"][i], a.collected_r)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("ghost,vis_excl", [(True, False), (False, True)])
def test_collision_proximity_matches_live_reference(ghost, vis_excl):
    """The part of the collision phase that IS reference code -- agent_agent_collision_proximity
    (sims.py:421-468), which builds the LIDAR field with Agent.projection_field -- driven with the
    partner lists of the restatement.  (The pair detection itself is pygame's: unpinned.)"""
    ref_shim.install()
    import types
    from abm.simulation import sims
    rng = np.random.default_rng(21 + ghost)
    N = 26
    cfg = rb.BaseConfig(R=1200, width=160, height=160, visual_exclusion=vis_excl, exp_vel_max=3.0, vision_range=2000)
    ov = rng.choice([0, 0, 1], N)
    st = dict(x=rng.integers(30, 160, N).astype(float), y=rng.integers(30, 160, N).astype(float),
              theta=rng.uniform(0, 2 * np.pi, N), vel=rng.uniform(0, 3, N), radius=10.0, w=np.zeros(N), u=np.zeros(N),
              novelty=np.zeros((N, 10)), env_status=np.zeros(N, int), override=ov, mode=ov.copy(),
              patch_id=np.full(N, -1), collected=np.zeros(N), collected_before=np.zeros(N))
    agents = ref_shim.make_base_agents(st, cfg)
    dummy = types.SimpleNamespace(ghost_mode=ghost, agents=agents)
    ix, iy = np.trunc(st["x"]), np.trunc(st["y"])
    n_pairs = 0
    for a1 in range(N):
        partners = [a2 for a2 in range(N) if a2 != a1 and (ix[a1] - ix[a2]) ** 2 + (iy[a1] - iy[a2]) ** 2 <= 24 ** 2]
        if partners:
            sims.Simulation.agent_agent_collision_proximity(dummy, agents[a1], [agents[j] for j in partners])
            n_pairs += len(partners)
    assert n_pairs > 10
    rb.base_collision_phase(st, cfg, ghost_mode=ghost)
    for i, a in enumerate(agents):
        assert np.isclose(st["theta"][i], a.orientation, rtol=1e-13), i
        assert np.isclose(st["vel"][i], a.velocity, rtol=1e-13), i
        # the restatement also ran the post-loop bookkeeping of sims.py:778-783: compare before it
        if a.overriding_mode == "exploit":
            assert st["override"][i] == rb.OV_EXPLOIT


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("border_overlap", [False, True])
def test_patch_regeneration_matches_live_reference(monkeypatch, border_overlap):
    """Simulation.kill_resource / add_new_resource_patch(force_id) (sims.py:321-374) of the real reference, its four
    numpy draws per try replaced by a prepared sequence, against oracle/restate_base.base_regenerate_patch fed the same
    draws: same number of tries, same position / units / quality / id.  (The overlap predicate itself is pygame's
    collide_circle -- restated from pygame's documentation in the harness, see ref_shim._collide_circle.)"""
    ref_shim.install()
    import pygame
    from abm.simulation import sims
    from abm.environment.rescource import Rescource
    rng = np.random.default_rng(12)
    W = H = 400
    pad, R = 30, 30
    for trial in range(40):
        P = int(rng.integers(2, 6))
        # existing patches (non-overlapping, like create_resources leaves them)
        pos = []
        while len(pos) < P:
            x, y = int(rng.integers(pad, W + pad - 2 * R)), int(rng.integers(pad, H + pad - 2 * R))
            if all((x - a) ** 2 + (y - b) ** 2 > (2 * R) ** 2 for a, b in pos):
                pos.append((x, y))
        group = pygame.sprite.Group()
        res = [Rescource(i + 1, R, pos[i], (W, H), (0, 0, 0), pad, 50, 0.5) for i in range(P)]
        for r_ in res:
            group.add(r_)
        p = int(rng.integers(0, P))
        # draws: a few tries that land on an existing patch, then free positions
        T = 12
        draws = np.zeros((T, 4))
        for t in range(T):
            if t < 3 and rng.uniform() < 0.7:
                q = int(rng.choice([k for k in range(P) if k != p]))
                draws[t, 0] = pos[q][0] + int(rng.integers(-R, R)); draws[t, 1] = pos[q][1] + int(rng.integers(-R, R))
            else:
                draws[t, 0] = int(rng.integers(pad - R if border_overlap else pad, W + pad - R))
                draws[t, 1] = int(rng.integers(pad - R if border_overlap else pad, H + pad - R))
            draws[t, 2] = int(rng.integers(20, 60)); draws[t, 3] = rng.uniform(0.1, 1.0)
        # the reference, its RNG calls replaced by the prepared values in the order it makes them (:351-361)
        seq = [v for row in draws for v in (int(row[0]), int(row[1]), int(row[2]), float(row[3]))]
        it = iter(seq)
        monkeypatch.setattr(np.random, "randint", lambda *a, **k: next(it))
        monkeypatch.setattr(np.random, "uniform", lambda *a, **k: next(it))
        sim = object.__new__(sims.Simulation)
        sim.rescources, sim.agents = group, pygame.sprite.Group()
        sim.resc_radius, sim.allow_border_patch_overlap, sim.regenerate_resources = R, border_overlap, True
        sim.window_pad, sim.WIDTH, sim.HEIGHT = pad, W, H
        sim.min_resc_units, sim.max_resc_units, sim.min_resc_quality, sim.max_resc_quality = 20, 60, 0.1, 1.0
        victim = res[p]
        victim.show_stats = False
        try:
            sim.kill_resource(victim)
            failed = False
        except StopIteration:
            failed = True                         # every prepared try overlapped
        monkeypatch.undo()
        patches = dict(x=np.array([q[0] for q in pos], float), y=np.array([q[1] for q in pos], float),
                       radius=np.full(P, float(R)), left=np.full(P, 50.0), quality=np.full(P, 0.5),
                       id=np.arange(1, P + 1))
        used = rb.base_regenerate_patch(patches, p, draws, R)
        if failed:
            assert used == 0
            continue
        new = [r_ for r_ in group if r_.id == p + 1]
        assert len(new) == 1 and len(group) == P                       # same id, one patch per id
        n = new[0]
        assert used >= 1 and (n.position[0], n.position[1]) == (patches["x"][p], patches["y"][p])
        assert n.resc_left == patches["left"][p] and n.unit_per_timestep == patches["quality"][p]
        assert (4 * used) == len(seq) - len(list(it))                  # the reference consumed exactly `used` tries
