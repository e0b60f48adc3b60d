"""TEST INFRASTRUCTURE ONLY -- CPU oracle (float64, numpy) for the BASE / collective-foraging
variant of the hot path: social visual field with distance-ordered occlusion, decision
process, mode machine, kinematics and patch exploitation.  Restated from the reference's
observable behaviour; every function cites the reference file:line it follows.  See
oracle/restate.py for the conventions and the import rules (tests / smoke / bench only).

Pinned by tests/test_oracle_base.py against SURVEY.md KAT-B1 and against fixtures produced
by executing the unmodified reference's Agent.update (tests/golden/base_golden.npz).
Agent-agent collision (sims.py:736-783) depends on pygame's collide_circle on int-truncated
rect centres, which is not in the reference tree: that part is "parity unpinned" (SURVEY 8c).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .restate import bin_grid, closed_angle_base, heading_vector, nearest_bin, reflect_from_walls, \
    wrap_heading_once

# override codes (Agent.overriding_mode, agent.py:671-693) and logged mode codes (ifdb.py:197-206)
OV_NONE, OV_EXPLOIT, OV_COLLIDE = 0, 1, 3
MODE_EXPLORE, MODE_EXPLOIT, MODE_RELOCATE, MODE_COLLIDE = 0, 1, 2, 3


@dataclass
class BaseConfig:
    R: int = 1200
    fov: tuple = (-np.pi, np.pi)
    width: float = 500.0
    height: float = 500.0
    window_pad: float = 30.0
    vision_range: float = 2000.0
    visual_exclusion: bool = True
    patchwise_exclusion: bool = True
    # decision_params.py:13-42
    T_w: float = 0.5; Eps_w: float = 3.0; g_w: float = 0.085; B_w: float = 0.0; w_max: float = 1.0
    T_u: float = 0.5; Eps_u: float = 3.0; g_u: float = 0.085; B_u: float = 0.0; u_max: float = 1.0
    S_wu: float = 0.25; S_uw: float = 0.01
    Tau: int = 10; F_N: float = 2.0; F_R: float = 1.0
    # movement_params.py:13-23
    exp_vel_max: float = 1.0; exp_theta_min: float = -0.3; exp_theta_max: float = 0.3
    reloc_theta_max: float = 0.5; exp_stop_ratio: float = 0.08
    # sims.py kwargs
    agent_consumption: float = 1.0
    teleport_exploit: bool = False


def _np_slice(a: int, b: int, R: int):
    """Effective [lo, hi) of the numpy basic slice v[a:b] on a length-R array (negative
    indices count from the end, everything is clamped) -- the reference's fills rely on it
    (agent.py:579-590)."""
    if a < 0:
        a = max(a + R, 0)
    if b < 0:
        b = max(b + R, 0)
    return min(a, R), min(b, R)


def base_candidates(i, x, y, r, is_exploit, patch_id, cfg: BaseConfig):
    """Agent.calc_social_V_proj (agent.py:396-419): returns (social, occluders) index lists
    in the reference's list order.  ``is_exploit[j]`` = (agents[j].get_mode() == "exploit")."""
    N = len(x)
    ra = np.broadcast_to(np.asarray(r, np.float64), (N,))                         # every agent's OWN radius here
    d = np.sqrt(((x + ra) - (x[i] + ra[i])) ** 2 + ((y + ra) - (y[i] + ra[i])) ** 2)   # supcalc.distance :73-78
    cand = [j for j in range(N) if d[j] <= cfg.vision_range]                      # :400 (includes self)
    expl = [j for j in cand if j != i and is_exploit[j]]                          # :402-403
    non_expl = [j for j in cand if j not in expl]                                 # :405
    if cfg.patchwise_exclusion:                                                   # :406-410
        non_expl.extend([j for j in expl if patch_id[j] == patch_id[i]])
        expl = [j for j in expl if patch_id[j] != patch_id[i]]
    expl = [j for j in expl if patch_id[j] != -1]                                 # :413
    return expl, non_expl


def base_source_data(i, x, y, r, theta, social, occluders, cfg: BaseConfig, fov=None):
    """Per-object interval data of Agent.projection_field (agent.py:484-556) for the object
    list [social..., occluders...]; returns list of dicts in insertion order (visible only)."""
    fov = cfg.fov if fov is None else fov
    R = cfg.R
    phis = bin_grid(R)
    objs = list(social) + list(occluders)
    n_social = len(social)
    v1x, v1y = heading_vector(x[i], y[i], r, theta[i])
    out = []
    for idx, j in enumerate(objs):
        if x[j] == x[i] and y[j] == y[i]:                                         # :502
            continue
        v2x = (x[j] + r) - (x[i] + r)                                             # :504-509 (focal radius for both)
        v2y = (y[j] + r) - (y[i] + r)
        ca = float(closed_angle_base(np.float64(v1x), np.float64(v1y), np.float64(v2x), np.float64(v2y)))
        dist = float(np.sqrt(v2x * v2x + v2y * v2y))                              # :526-528
        vis_angle = 2 * np.arctan(r / (1 * dist))                                 # :529
        k = int(nearest_bin(phis, np.float64(ca)))                                # :532
        if fov[0] < ca < fov[1]:                                                  # :535
            size = (vis_angle / (2 * np.pi)) * R                                  # :543
            s = int(k - size / 2)                                                 # :545 (truncation toward 0)
            e = int(k + size / 2)
            out.append(dict(j=j, k=k, d=dist, s=s, e=e, sx=s, ex=e, social=idx < n_social, order=idx))
    return out


def base_occlude(src):
    """Agent.exlude_V_source_data (agent.py:421-445): stable sort by distance, then for every
    object but the nearest apply the three clipping rules of every strictly closer object,
    sequentially, on the object's current (sx, ex); raw ends of the closer object; no wrap."""
    ranked = sorted(src, key=lambda v: v["d"])                                    # :424 (stable)
    for rank, f in enumerate(ranked):
        if rank == 0:
            continue
        for o in ranked:
            if o["d"] < f["d"]:
                if f["sx"] <= o["s"] <= f["ex"]:
                    f["ex"] = o["s"]
                if f["sx"] <= o["e"] <= f["ex"]:
                    f["sx"] = o["e"]
                if o["s"] <= f["sx"] and o["e"] >= f["ex"]:
                    f["sx"] = 0
                    f["ex"] = 0
    return ranked


def base_fill(src, cfg: BaseConfig, fov=None):
    """Fill + flip + FOV mask of Agent.projection_field (agent.py:569-595), binary form.
    Returns the STORED (flipped, masked) field as bool (R,)."""
    fov = cfg.fov if fov is None else fov
    R = cfg.R
    v = np.zeros(R, bool)
    for o in src:
        s, e = o["sx"], o["ex"]
        if s < 0:                                                                 # :578-580
            lo, hi = _np_slice(R + s, R, R)
            v[lo:hi] = True
            s = 0
        if e >= R:                                                                # :582-584
            lo, hi = _np_slice(0, e - R, R)
            v[lo:hi] = True
            e = R - 1
        lo, hi = _np_slice(s, e, R)                                               # :588
        v[lo:hi] = True
    stored = v[::-1].copy()                                                       # :593
    phis = bin_grid(R)
    stored[phis < fov[0]] = False                                                 # :594-595
    stored[phis > fov[1]] = False
    return stored


def base_field(i, x, y, r, theta, is_exploit, patch_id, cfg: BaseConfig):
    """agent.soc_v_field after Agent.calc_social_V_proj (agent.py:396-419, 457-597)."""
    social, occl = base_candidates(i, x, y, r, is_exploit, patch_id, cfg)
    if np.ndim(r):      # heterogeneous radii: the projection uses the FOCAL radius for both centres (agent.py:504-509)
        r = float(np.asarray(r)[i])
    if cfg.visual_exclusion:
        src = base_source_data(i, x, y, r, theta, social, occl, cfg)
        src = base_occlude(src)                                                   # :559-560
        src = [o for o in src if o["social"]]                                     # :562-564
    else:
        src = base_source_data(i, x, y, r, theta, social, [], cfg)
    return base_fill(src, cfg), src


def relocation_force(vel, field, v_desired, cfg: BaseConfig):
    """supcalc.F_reloc_LR (supcalc.py:81-92) on the stored field."""
    R = len(field)
    left = np.mean(field[0:int(R / 2)])
    right = np.mean(field[int(R / 2):])
    return v_desired - vel, (left - right) * cfg.reloc_theta_max


def base_agent_update(i, st, cfg: BaseConfig, dtheta_random):
    """Agent.update (agent.py:212-283) for agent i from the frozen snapshot ``st`` (dict of
    arrays, see base_step_frozen).  ``dtheta_random`` replaces the np.random.uniform draw of
    supcalc.random_walk (supcalc.py:45).  Returns a dict of agent i's new scalars + field."""
    r = st["radius"]                                 # scalar, or (N,) for heterogeneous radii (sims.py:502)
    is_exploit = st["override"] == OV_EXPLOIT        # get_mode() == "exploit" (agent.py:659-669)
    field, _src = base_field(i, st["x"], st["y"], r, st["theta"], is_exploit, st["patch_id"], cfg)
    if np.ndim(r):
        r = float(np.asarray(r)[i])                  # own radius in reflect_from_walls (agent.py:347-394)
    # calc_I_priv (agent.py:168-175)
    collected_unit = st["collected"][i] - st["collected_before"][i]
    I_priv = cfg.F_N * np.max(st["novelty"][i]) + cfg.F_R * collected_unit
    # update_decision_processes (agent.py:194-210)
    w, u = st["w"][i], st["u"][i]
    w_p = w if w > cfg.T_w else 0
    u_p = u if u > cfg.T_u else 0
    dw = cfg.Eps_w * np.mean(field) - cfg.g_w * (w - cfg.B_w) - u_p * cfg.S_uw
    du = cfg.Eps_u * I_priv - cfg.g_u * (u - cfg.B_u) - w_p * cfg.S_wu
    w = w + dw
    u = u + du
    w = min(max(w, -cfg.w_max), cfg.w_max)
    u = min(max(u, -cfg.u_max), cfg.u_max)
    # mode machine (agent.py:233-265); tr_u compares with T_w (agent.py:652-657)
    W, U = w > cfg.T_w, u > cfg.T_w
    override, mode = int(st["override"][i]), int(st["mode"][i])
    vel0, th0 = st["vel"][i], st["theta"][i]
    env1 = st["env_status"][i] == 1
    if override != OV_COLLIDE:
        if (not W and not U) or (U and not W and not env1):
            dvel, dth = cfg.exp_vel_max, dtheta_random                           # random_walk
            override, mode = OV_NONE, MODE_EXPLORE
        elif U and env1:                                                         # both W&&U and U&&!W
            dvel, dth = -vel0 * cfg.exp_stop_ratio, 0.0
            override, mode = OV_EXPLOIT, MODE_EXPLOIT
        else:                                                                    # W and (not U or not env1)
            dvel, dth = relocation_force(vel0, field, cfg.exp_vel_max, cfg)
            override, mode = OV_NONE, MODE_RELOCATE
    else:
        dvel, dth = 0.0, 0.0
    th = wrap_heading_once(th0 + dth)                                            # :269-270
    vel = vel0 + dvel
    # prove_velocity (agent.py:612-620): only while get_mode() == 'explore'
    if override == OV_NONE and not W:
        if np.abs(vel) > 1:
            vel = cfg.exp_vel_max
    x = st["x"][i] + vel * np.cos(th)                                            # :275-276
    y = st["y"][i] - vel * np.sin(th)
    x, y, th = reflect_from_walls(x, y, th, r, cfg.width, cfg.height, cfg.window_pad)
    return dict(field=field, x=x, y=y, theta=th, vel=vel, w=w, u=u, I_priv=I_priv, override=override,
                mode=mode, collected_before=st["collected"][i])


def base_step_frozen(st, cfg: BaseConfig, dtheta_random, agents=None, agent_cfgs=None):
    """Synchronous agent phase (sims.py:861 with every agent seeing the same snapshot).
    st: dict with x, y, theta, vel, w, u (N,), novelty (N, Tau), env_status, override, mode,
    patch_id (N,) ints, collected, collected_before (N,), radius (scalar).
    ``agent_cfgs``: optional list of one BaseConfig per agent (heterogeneous agents, agent.py:83-108: the
    behave_params entries replace the decision parameters, max_exp_vel and exp_stop_ratio of that agent;
    FOV, vision range and the field resolution R may differ too (sims.py:499-517); ``fields`` rows are padded with
    False up to ``cfg.R``, which must be the largest)."""
    N = len(st["x"])
    out = {k: np.array(st[k], copy=True) for k in ("x", "y", "theta", "vel", "w", "u", "override", "mode",
                                                   "collected_before")}
    out["I_priv"] = np.zeros(N)
    out["fields"] = np.zeros((N, cfg.R), bool)
    for i in (range(N) if agents is None else agents):
        r = base_agent_update(i, st, cfg if agent_cfgs is None else agent_cfgs[i], dtheta_random[i])
        out["fields"][i, :len(r["field"])] = r["field"]
        for k in ("x", "y", "theta", "vel", "w", "u", "override", "mode", "collected_before", "I_priv"):
            out[k][i] = r[k]
    return out


# --------------------------------------------------------------------------------------
# environment phase: agent-patch interaction (sims.py:29-56, 544-552, 790-858; rescource.py:118-133)
# --------------------------------------------------------------------------------------

def notify(st, i, status, res_id=None):
    """sims.notify_agent (sims.py:29-42)."""
    before = st["env_status"][i]
    st["env_status"][i] = status
    st["novelty"][i] = np.roll(st["novelty"][i], 1)
    st["novelty"][i][0] = 1 if status - before > 0 else 0
    st["patch_id"][i] = -1 if res_id is None else res_id


def base_patch_phase(st, patches, cfg: BaseConfig, collided=(), agent_cfgs=None):
    """Agent-patch interaction of one time step, in place.  ``patches``: dict with arrays
    x, y (top-left), radius, left, quality, id (n_patch,).  Patches are visited in slot order,
    agents in index order (group order, sims.py:805-844).  Returns the list of depleted slots
    (to be regenerated: sims.py:321-374, RNG-driven, not part of the parity contract)."""
    N = len(st["x"])
    ra = np.broadcast_to(np.asarray(st["radius"], np.float64), (N,))   # every agent's own radius (supcalc.distance :73-78)
    on_patch = np.zeros(N, bool)
    depleted = []
    for p in range(len(patches["x"])):
        pcx, pcy = patches["x"][p] + patches["radius"][p], patches["y"][p] + patches["radius"][p]
        members = [i for i in range(N)
                   if np.sqrt((st["x"][i] + ra[i] - pcx) ** 2 + (st["y"][i] + ra[i] - pcy) ** 2) < patches["radius"][p]]   # :45-56
        destroy = False
        for i in members:
            r = ra[i]
            # bias_agent_towards_res_center (sims.py:544-552), relative_speed 0.02, no wrap
            dx, dy = pcx - (st["x"][i] + r), pcy - (st["y"][i] + r)
            cl = (np.arctan2(dy, dx) + st["theta"][i]) % (2 * np.pi)
            st["theta"][i] += (cl - np.pi) * 0.02
            if destroy:
                notify(st, i, -1)                                                 # :812-813
            else:
                notify(st, i, 1, patches["id"][p])                               # :816-818 (pooling_time == 0)
                if cfg.teleport_exploit:                                          # :820-821
                    st["x"][i] = patches["x"][p] + patches["radius"][p] - r
                    st["y"][i] = patches["y"][p] + patches["radius"][p] - r
                if st["override"][i] == OV_EXPLOIT:                               # :824
                    consumption = (cfg if agent_cfgs is None else agent_cfgs[i]).agent_consumption   # agent.consumption
                    take = min(consumption, patches["quality"][p])                # rescource.py:121-122
                    if patches["left"][p] >= take:
                        patches["left"][p] -= take
                    else:
                        take = patches["left"][p]
                        patches["left"][p] = 0
                    destroy = not (patches["left"][p] > 0)
                    st["collected_before"][i] = st["collected"][i]               # :827
                    st["collected"][i] += take
                    if destroy:                                                   # :829-836
                        for i2 in members:
                            notify(st, i2, -1)
            on_patch[i] = True
        if destroy:
            depleted.append(p)
    for i in range(N):                                                            # :847-855
        if not on_patch[i] and i not in collided:
            notify(st, i, -1)
    return depleted


def base_regenerate_patch(patches, p, draws, patch_radius):
    """Simulation.kill_resource + add_new_resource_patch(force_id) (sims.py:321-374) for patch slot ``p`` with the four
    random draws of every try GIVEN: ``draws`` (n_tries, 4) = (x, y, units, quality) in the order the reference draws
    them (:351-361).  A try is accepted when the new patch overlaps no other patch (proove_sprite with
    prove_with_res only, :362-372: pygame's collide_circle on the rect centres, touching counts -- pygame's documented
    rule, restated).  The patch keeps its id and slot.  Returns the number of tries used (0: every try overlapped)."""
    R = float(patch_radius)
    for t, (x, y, units, quality) in enumerate(np.asarray(draws, np.float64)):
        ok = True
        for q in range(len(patches["x"])):
            if q == p or patches["radius"][q] <= 0:            # the killed patch itself / dead slots are not in the group
                continue
            r2 = float(patches["radius"][q])
            dx = (x + R) - (patches["x"][q] + r2)
            dy = (y + R) - (patches["y"][q] + r2)
            if dx * dx + dy * dy <= (R + r2) ** 2:
                ok = False
                break
        if ok:
            patches["x"][p], patches["y"][p], patches["radius"][p] = x, y, R
            patches["left"][p], patches["quality"][p] = int(units), quality
            return t + 1
    patches["radius"][p] = 0.0
    return 0


# --------------------------------------------------------------------------------------
# collision phase (sims.py:736-783, 421-468; interactions.py:5-10)  -- PARITY UNPINNED
# --------------------------------------------------------------------------------------

def base_collision_phase(st, cfg: BaseConfig, ghost_mode: bool, agent_cfgs=None):
    """Agent-agent collision avoidance of one time step, in place.  Restated from the
    reference's control flow plus pygame's DOCUMENTED semantics (pygame itself is not in the
    reference tree, so this part cannot be pinned against reference output -- SURVEY 8c):
    groupcollide(agents, agents, within_group_collision) -> for every agent a1 in group order
    the list of a2 != a1 with (rect centre distance)^2 <= (r1 + r2)^2, radii temporarily +2,
    rect.x / rect.y = int-truncated position (agent.py:303-304).  Returns the list
    `collided_agents` (with repetitions, as the reference builds it)."""
    N = len(st["x"])
    ra = np.broadcast_to(np.asarray(st["radius"], np.float64), (N,))              # own radii (heterogeneous agents)
    # rect centres: int-truncated position + half the (integer) rect size (agent.py:303-304; pygame Rect)
    ix, iy = np.trunc(st["x"]) + np.trunc(ra), np.trunc(st["y"]) + np.trunc(ra)
    collided = []
    for a1 in range(N):
        partners = [a2 for a2 in range(N)                                         # radii + 2 each, sims.py:739-752
                    if a2 != a1 and (ix[a1] - ix[a2]) ** 2 + (iy[a1] - iy[a2]) ** 2 <= (ra[a1] + ra[a2] + 4) ** 2]
        if not partners:
            continue
        for a2 in partners:                                                       # agent_agent_collision_proximity :421-468
            do = True
            if ghost_mode:
                do = st["override"][a2] != OV_EXPLOIT and st["override"][a1] != OV_EXPLOIT
            if not do:
                continue
            if st["override"][a2] != OV_EXPLOIT:
                st["override"][a2] = OV_COLLIDE
                st["mode"][a2] = MODE_COLLIDE
            r = ra[a2]                                                            # the hit agent is the focal agent
            c2 = cfg if agent_cfgs is None else agent_cfgs[a2]                    # ... with its own resolution / range
            R = c2.R
            h = int(R / 2)
            d = np.sqrt(((st["x"] + ra) - (st["x"][a2] + r)) ** 2 + ((st["y"] + ra) - (st["y"][a2] + r)) ** 2)
            vicinity = [j for j in range(N) if d[j] < 2 * r + 20 and j != a2]     # :446-447 (own centres)
            full = (-np.pi, np.pi)
            src = base_source_data(a2, st["x"], st["y"], r, st["theta"], vicinity, [], c2, fov=full)
            if cfg.visual_exclusion:
                src = base_occlude(src)
            field = base_fill(src, c2, fov=full)                                  # :449 (binary part)
            last = [j for j in vicinity if not (st["x"][j] == st["x"][a2] and st["y"][j] == st["y"][a2])]
            amp = 1.0
            if last:                      # leaked loop variable of projection_field (agent.py:526-528, :590): the distance
                j = last[-1]              # with BOTH centres formed with the focal radius
                dl = np.sqrt(((st["x"][j] + r) - (st["x"][a2] + r)) ** 2 + ((st["y"][j] + r) - (st["y"][a2] + r)) ** 2)
                amp = 1 - dl / c2.vision_range
            left = amp * field[0:h].sum() / h
            right = amp * field[h:].sum() / (R - h)
            D = np.sign(left - right)
            if D == 0:
                D = -1
            if st["override"][a2] != OV_EXPLOIT:
                st["theta"][a2] -= D * 0.2                                        # :458-459, not wrapped
            lo, hi = _np_slice(h - 100, h + 100, R)
            if amp * field[lo:hi].sum() > 0:                                      # :462-463
                st["vel"][a2] = 0
            elif st["override"][a2] != OV_EXPLOIT:
                st["vel"][a2] = (cfg if agent_cfgs is None else agent_cfgs[a2]).exp_vel_max   # agent2.max_exp_vel (:465)
        for a2 in partners:                                                       # sims.py:759-776
            e1, e2 = st["override"][a1] == OV_EXPLOIT, st["override"][a2] == OV_EXPLOIT
            if cfg.teleport_exploit:
                if not e1:
                    collided.append(a1)
                if not e2:
                    collided.append(a2)
            elif not ghost_mode:
                collided += [a1, a2]
            elif not e1 and not e2:
                collided += [a1, a2]
    cset = set(collided)
    for i in range(N):                                                            # :778-783
        if i not in cset and st["override"][i] == OV_COLLIDE:
            st["override"][i] = OV_NONE
            st["mode"][i] = MODE_EXPLORE
        if i in cset and st["override"][i] == OV_COLLIDE:
            notify(st, i, -1)
    return collided
