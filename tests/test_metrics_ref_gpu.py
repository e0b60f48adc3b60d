"""the on-device summary metrics against the REFERENCE'S OWN analysis functions
(abm/loader/data_loader.py, unmodified: the modules travel in oracle/_ref) applied to the trajectories the engine
produced -- not against formulas re-written in the test.  The loader works on time series of logged states; the engine's
`metrics()` is called after every step, so the two are compared step by step and as time averages."""
import numpy as np
import pytest

from oracle import ref_shim

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shim.reference_available(), reason="reference modules (oracle/_ref) not available")]


@pytest.mark.parametrize("boundary", ["walls"])
def test_vf_metrics_match_reference_loader(built_lib, tmp_path, boundary):
    """polarization (data_loader.py:1761-1836), inter-individual distance (:1367-1460), mean nearest-neighbour distance
    (:1461-1488) and agent-agent collision time (:1838-1869) of the real ExperimentLoader on the engine's trajectory."""
    from abm_b200 import VFEngine
    B, N, T, W = 3, 14, 1000, 160.0                         # calculate_collision_time refuses fewer than 1000 time steps
    rng = np.random.default_rng(8)
    x = rng.uniform(40, 150, (B, N)).astype(np.float32); y = rng.uniform(40, 150, (B, N)).astype(np.float32)
    th = rng.uniform(0, 2 * np.pi, (B, N)).astype(np.float32); v = np.zeros((B, N), np.float32)
    eng = VFEngine(B, N, resolution=1200, width=W, height=W, boundary=boundary)
    eng.set_params(ALP0=np.array([0.5, 1.0, 3.0]), BET0=np.array([0.5, 1.0, 0.2]))
    eng.set_state(x, y, th, v, 10.0)
    posx = np.zeros((B, N, T)); posy = np.zeros((B, N, T)); ori = np.zeros((B, N, T))
    mine = {k: np.zeros((B, T)) for k in ("polarization", "mean_iid", "mean_nn_dist", "collision", "colliding_agents")}
    for t in range(T):
        eng.step(1)
        st = eng.get_state()
        posx[..., t], posy[..., t], ori[..., t] = st["x"], st["y"], st["theta"]
        m = eng.metrics()
        for k in mine:
            mine[k][:, t] = m[k]
    eng.close()
    ld = ref_shim.make_loader(tmp_path, dict(posx=posx, posy=posy, orientation=ori),
                              {"T": T, "RADIUS_AGENT": 10, "BOUNDARY": boundary, "ENV_WIDTH": W, "ENV_HEIGHT": W})
    pol, _ = ld.calculate_polarization()                                          # (B, 1, T)
    np.testing.assert_allclose(mine["polarization"], pol[:, 0], rtol=2e-5, atol=1e-6)
    ld.calculate_interindividual_distance()                                      # iid (B, 1, N, N, T), upper triangle
    iu = np.triu_indices(N, k=1)
    np.testing.assert_allclose(mine["mean_iid"], ld.iid_matrix[:, 0][:, iu[0], iu[1], :].mean(axis=1), rtol=2e-5)
    np.testing.assert_allclose(mine["mean_iid"].mean(axis=0), ld.mean_iid[0], rtol=2e-5)      # the loader's own mean
    aacoll, mean_aacoll = ld.calculate_collision_time()      # (B, 1, N): fraction of the steps agent i collides with a j > i
    np.testing.assert_allclose(mine["colliding_agents"].mean(axis=1), aacoll[:, 0].mean(axis=1), atol=1e-6)
    np.testing.assert_allclose(mine["colliding_agents"].mean(), mean_aacoll[0], atol=1e-6)     # the experiment's "aacoll"
    assert ((mine["collision"] > 0) == (mine["colliding_agents"] > 0)).all() and mine["collision"].mean() > 0
    # nearest-neighbour distance: the loader takes nanmin over a matrix whose lower triangle it never filled (zeros), so
    # its own result is the first agent's value / N; on the symmetrised matrix the same function is the quantity meant
    ld.iid_matrix = ld.iid_matrix + np.swapaxes(ld.iid_matrix, 2, 3)
    ld.calculate_mean_NN_dist()
    np.testing.assert_allclose(mine["mean_nn_dist"].mean(axis=0), ld.mean_nn_dist[0], rtol=2e-5)


def test_base_metrics_match_reference_loader(built_lib, tmp_path):
    """search efficiency (data_loader.py:1294-1353) and relative relocation time (:1903-1928) of the real ExperimentLoader on
    the per-step `collresource` / `mode` arrays of an engine run, against BaseEngine.metrics() accumulated on the
    device over the same run."""
    from abm_b200 import BaseEngine
    B, N, P, W, T = 6, 20, 3, 300.0, 250
    rng = np.random.default_rng(2)
    eng = BaseEngine(B, N, P, resolution=1200, width=W, height=W, visual_exclusion=True, patch_radius=30.0,
                     min_resc_perpatch=40, max_resc_perpatch=60, seed=21)
    eng.set_params(Eps_w=np.linspace(0.5, 3, B), Eps_u=1.0, F_N=0.5, F_R=0.5, exp_vel_max=3.0, exp_theta_min=-0.5,
                   exp_theta_max=0.5, reloc_theta_max=1.8, exp_stop_ratio=0.175)
    eng.set_agents(x=rng.integers(20, 320, (B, N)), y=rng.integers(20, 320, (B, N)), theta=rng.uniform(0, 2 * np.pi, (B, N)))
    eng.set_patches(x=rng.integers(40, 240, (B, P)), y=rng.integers(40, 240, (B, P)), radius=np.full((B, P), 30.0),
                    left=np.full((B, P), 50.0), quality=np.full((B, P), 0.25), id=np.tile(np.arange(P), (B, 1)))
    mode = np.zeros((B, N, T)); coll = np.zeros((B, N, T))
    for t in range(T):
        eng.step(1)
        a = eng.get_agents(["mode", "collected"])
        mode[..., t], coll[..., t] = a["mode"], a["collected"]
    m = eng.metrics()
    eng.close()
    ld = ref_shim.make_loader(tmp_path, dict(mode=mode, collresource=coll, orientation=mode), {"T": T})
    ld.calculate_search_efficiency()                                             # (B, N): (collres[-1] - collres[0]) / T
    np.testing.assert_allclose(m["search_efficiency"], ld.efficiency[:, 0].mean(axis=1) + coll[..., 0].mean(axis=1) / T,
                               rtol=1e-5, atol=1e-7)
    reloc, _ = ld.calculate_relocation_time()                                    # (B, 1, N): mean over time of (mode == 2)
    np.testing.assert_allclose(m["relocation_time"], reloc[:, 0].mean(axis=1), rtol=1e-6, atol=1e-7)
    assert m["relocation_time"].max() > 0 and coll[..., -1].sum() > 0
