"""CPU property tests (hypothesis) of host-side arithmetic and of oracle helpers that restate numpy semantics: the shard /
tile maps partition their index sets for every size, the numpy-slice restatement equals numpy's own slicing, and the
line-following restatement equals a direct numpy evaluation of the reference's expression on random maps."""
import numpy as np
from hypothesis import given, settings, strategies as st

from abm_b200 import multigpu as mg
from oracle import restate as rs
from oracle import restate_base as rb


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 5000), st.integers(1, 16))
def test_replicate_shards_partition_any_batch(B, world):
    spans = [mg.replicate_shard(B, world, r) for r in range(world)]
    assert sum(c for _, c in spans) == B and spans[0][0] == 0
    for (b0, c0), (b1, _) in zip(spans, spans[1:]):
        assert b1 == b0 + c0
    counts = [c for _, c in spans]
    assert max(counts) - min(counts) <= 1                      # balanced


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 8), st.integers(1, 12), st.sampled_from([1, 8, 128]))
def test_cyclic_slots_partition_the_swarm(world, blocks_per_rank, block):
    n = world * blocks_per_rank * block
    slots = [mg.cyclic_slots(n, world, r, block) for r in range(world)]
    allslots = np.concatenate(slots)
    assert np.array_equal(np.sort(allslots), np.arange(n))     # every slot exactly once
    for r, s in enumerate(slots):
        assert len(s) == n // world and np.all(np.diff(s) > 0)
        assert np.all((s // block) % world == r)               # rank r owns every world-th block


@settings(max_examples=300, deadline=None)
@given(st.integers(-40, 40), st.integers(-40, 40), st.integers(1, 25))
def test_slice_bounds_equal_numpy(a, b, n):
    """The oracle's restatement of numpy basic slicing (negative bounds wrap once, out-of-range bounds clip) -- used by
    the BASE fill (agent.py:577-588) and by follow_lines_local (vf_supcalc.py:310-312)."""
    v = np.arange(n)
    lo, hi = rb._np_slice(a, b, n)
    assert np.array_equal(v[a:b], v[lo:hi] if hi > lo else v[0:0])
    lo2, hi2 = rs._slice_bounds(a, b, n)
    assert np.array_equal(v[a:b], v[lo2:hi2] if hi2 > lo2 else v[0:0])


@settings(max_examples=150, deadline=None)
@given(st.floats(-20, 120), st.floats(-20, 120), st.floats(0, 6.3), st.floats(-2, 2), st.integers(0, 2**31 - 1))
def test_follow_lines_restatement_equals_the_reference_expression(x, y, ori, vel, seed):
    """follow_lines_local restated (oracle/restate.py) against the reference's expression evaluated directly with numpy
    (vf_supcalc.py:295-329) on a random map: same windows, same branches."""
    rng = np.random.default_rng(seed)
    lm = rng.choice([0.0, 0.0, 0.25, 1.0], (90, 70))
    r, sr, sd = 10.0, 9, 20
    got = rs.follow_lines_local((x, y), r, ori, lm, vel, sr, sd)
    s1 = [y + r - sd + (1 + np.sin(ori + (3 * np.pi / 4))) * sd, x + r - sd + (1 - np.cos(ori + (3 * np.pi / 4))) * sd]
    s2 = [y + r - sd + (1 + np.sin(ori - (3 * np.pi / 4))) * sd, x + r - sd + (1 - np.cos(ori - (3 * np.pi / 4))) * sd]
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m1 = np.nanmean(lm[int(s1[1] - sr):int(s1[1] + sr), int(s1[0] - sr):int(s1[0] + sr)])
            m2 = np.nanmean(lm[int(s2[1] - sr):int(s2[1] + sr), int(s2[0] - sr):int(s2[0] + sr)])
    if np.isnan(m1) or np.isnan(m2):
        want = 0
    else:
        oc = 0.5 * (m2 - m1) if np.sign(vel) else 0
        want = oc if m1 != m2 else (0.01 if m1 != 0 else 0)
    assert abs(got - want) <= 1e-12


@settings(max_examples=50, deadline=None)
@given(st.integers(1, 40), st.integers(1, 70), st.integers(1, 33), st.integers(0, 2**31 - 1))
def test_zarr_v2_directories_round_trip_and_agent_rows(n_agents, T, chunk, seed):
    """The recorders write zarr-format-2 directories without the zarr package: the minimal reader returns what the writer
    laid out, chunk by chunk (ragged last chunk included); `agent_row` is the reference's row rule (agent id - 1, so agent
    0 lands in the last row, ifdb.py:504-508) -- a bijection of the agents onto the rows."""
    import json
    import os
    import tempfile
    from abm_b200.recorder import _write_zarray, agent_row, read_zarr_v2
    rng = np.random.default_rng(seed)
    data = rng.normal(size=(n_agents, T))
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ag_ori.zarr")
        os.makedirs(path)
        _write_zarray(path, (n_agents, T), (n_agents, chunk))
        for k in range(-(-T // chunk)):
            blk = np.zeros((n_agents, chunk))
            w = min(chunk, T - k * chunk)
            blk[:, :w] = np.roll(data, -1, axis=0)[:, k * chunk:k * chunk + w]      # row = agent id - 1
            blk.tofile(os.path.join(path, f"0.{k}"))
        meta = json.load(open(os.path.join(path, ".zarray")))
        assert meta["zarr_format"] == 2 and meta["shape"] == [n_agents, T] and meta["dtype"] == "<f8"
        got = read_zarr_v2(path)
    assert got.shape == (n_agents, T)
    rows = [agent_row(i, n_agents) for i in range(n_agents)]
    assert sorted(rows) == list(range(n_agents)) and rows[0] == n_agents - 1
    for i in range(n_agents):
        assert np.array_equal(got[rows[i]], data[i])
