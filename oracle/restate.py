"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the ABM hot path (float64, numpy).

An independent restatement of the reference's algorithm (scioip34/ABM), written
from its observable behaviour; every function cites the reference file:line it
follows (paths relative to the reference root).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package -- as the checker, never as the
product.  The product path (``abm_b200/``) never imports it and fails loudly when
its CUDA library is missing.

Pinning: this restatement is checked against (a) the reference's own golden
vectors (abm/projects/cooperative_signaling/cs_agent/tests/test_cs_supcalc.py:143-168),
(b) the survey's known-answer vectors and (c) fixtures in
``tests/golden/`` produced by executing the UNMODIFIED reference in the build
container (``tests/golden/make_golden.py``).  See tests/test_oracle_*.py.

Conventions: screen coordinates (y down); ``position`` is the top-left corner of
the agent sprite, centre = position + radius; R = visual-field resolution;
"row" = un-flipped field as it is built, "stored" = row[::-1] (what the reference
keeps in ``agent.soc_v_field``).  State arrays are float64 copies of the engine's
fp32 state.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

TWO_PI = 2.0 * np.pi


# --------------------------------------------------------------------------------------
# angle / bin helpers
# --------------------------------------------------------------------------------------

def bin_grid(R: int) -> np.ndarray:
    """Bin centres used to PLACE projections: linspace(-pi, pi, R)
    (vf_supcalc.py:39, agent.py:481)."""
    return np.linspace(-np.pi, np.pi, R)


def phi_grid(R: int) -> np.ndarray:
    """Grid used to INTEGRATE the flocking terms: arange(-pi, pi, 2pi/R)
    (vf_agent.py:44).  Differs from bin_grid."""
    return np.arange(-np.pi, np.pi, (2 * np.pi) / R)


def nearest_bin(phis: np.ndarray, value) -> np.ndarray:
    """first-argmin |phis - value| (supcalc.py:8-11) for an array of values.

    Evaluated on the three candidates around the closed-form estimate
    ceil(x - 0.5) using the very same phis values, so the tie rule (lower index
    wins) and linspace rounding are reproduced exactly."""
    value = np.asarray(value, dtype=np.float64)
    R = phis.shape[0]
    step = TWO_PI / (R - 1)
    k0 = np.ceil((value + np.pi) / step - 0.5).astype(np.int64)
    k0 = np.clip(k0, 0, R - 1)
    cand = np.stack([np.clip(k0 - 1, 0, R - 1), k0, np.clip(k0 + 1, 0, R - 1)], axis=-1)
    dist = np.abs(phis[cand] - value[..., None])
    # first minimum among candidates ordered by index; duplicates (after clipping) are harmless
    order = np.argsort(cand, axis=-1, kind="stable")
    cand_s = np.take_along_axis(cand, order, axis=-1)
    dist_s = np.take_along_axis(dist, order, axis=-1)
    pick = np.argmin(dist_s, axis=-1)
    return np.take_along_axis(cand_s, pick[..., None], axis=-1)[..., 0]


def signed_angle(v1x, v1y, v2x, v2y):
    """angle_between (supcalc.py:19-34): arccos of the clipped dot product of the
    unit vectors, negated when the z-component of the cross product is < 0."""
    n1 = np.sqrt(v1x * v1x + v1y * v1y)
    n2 = np.sqrt(v2x * v2x + v2y * v2y)
    with np.errstate(invalid="ignore", divide="ignore"):
        a1x, a1y = v1x / n1, v1y / n1
        a2x, a2y = v2x / n2, v2y / n2
        ang = np.arccos(np.clip(a1x * a2x + a1y * a2y, -1.0, 1.0))
        neg = (a1x * a2y - a1y * a2x) < 0
    return np.where(neg, -ang, ang)


def closed_angle_vf(v1x, v1y, v2x, v2y):
    """calculate_closed_angle (vf_supcalc.py:142-158): [0, pi] -> negated, else 2pi - a."""
    a = np.mod(signed_angle(v1x, v1y, v2x, v2y), TWO_PI)
    return np.where((a >= 0) & (a <= np.pi), -a, TWO_PI - a)


def closed_angle_base(v1x, v1y, v2x, v2y):
    """BASE variant (agent.py:515-523): only the OPEN interval (0, pi) is negated, so a
    dead-ahead object (a == 0) maps to 2pi and becomes invisible."""
    a = np.mod(signed_angle(v1x, v1y, v2x, v2y), TWO_PI)
    return np.where((a > 0) & (a < np.pi), -a, TWO_PI - a)


def heading_vector(px, py, r, theta):
    """v1 = point on the agent's rim in heading direction minus its centre
    (vf_supcalc.py:42-49, agent.py:484-495), with the reference's rounding order."""
    cx, cy = px + r, py + r
    ex = px + (1 + np.cos(theta)) * r
    ey = py + (1 - np.sin(theta)) * r
    return ex - cx, ey - cy


# --------------------------------------------------------------------------------------
# VF: projection field (vf_supcalc.py:20-138) and union (vf_agent.py:216-237)
# --------------------------------------------------------------------------------------

@dataclass
class VFConfig:
    """Scalar configuration of a visual-flocking run (vf_params.py:12-23,
    app_visual_flocking.py:70-106)."""
    R: int = 1200
    fov: tuple = (-np.pi, np.pi)          # agent FOV in radians (sims.py:160-161)
    boundary: str = "walls"                # "walls" | "infinite"
    width: float = 900.0
    height: float = 900.0
    window_pad: float = 30.0
    GAM: float = 0.1
    V0: float = 1.0
    ALP0: float = 1.0
    ALP1: float = 0.09
    BET0: float = 1.0
    BET1: float = 0.09
    limit_movement: bool = False
    max_vel: float = 3.0
    max_th: float = 0.1


def vf_intervals(px, py, r, theta, i: int, cfg: VFConfig):
    """Per-object interval data for focal agent ``i`` against ALL agents j != i
    (vf_supcalc.py:55-117 with the object list of vf_agent.py:219-232).

    Returns dict of arrays over j (length N, entry i masked out by ``valid``):
    k (centre bin), h (half width), ps, pe (raw ends), drawn (passes :57 and :119)."""
    px = np.asarray(px, np.float64); py = np.asarray(py, np.float64)
    r = np.broadcast_to(np.asarray(r, np.float64), px.shape)
    N = px.shape[0]
    R = cfg.R
    phis = bin_grid(R)
    valid = np.arange(N) != i
    # :57 objects whose position coincides exactly with the focal position are skipped
    valid &= ~((px == px[i]) & (py == py[i]))
    cix, ciy = px[i] + r[i], py[i] + r[i]
    cjx, cjy = px + r, py + r                                     # :61-64 (object_sizes given)
    v1x, v1y = heading_vector(px[i], py[i], r[i], theta[i])
    v2x, v2y = cjx - cix, cjy - ciy
    if cfg.boundary == "infinite":                                # :70-83
        W, H = cfg.width, cfg.height
        farx = np.abs(v2x) > W / 2
        cjx = np.where(farx & (cix < cjx), cjx - W, np.where(farx & (cix > cjx), cjx + W, cjx))
        fary = np.abs(v2y) > H / 2
        cjy = np.where(fary & (ciy < cjy), cjy - H, np.where(fary & (ciy > cjy), cjy + H, cjy))
        v2x, v2y = cjx - cix, cjy - ciy
    ca = closed_angle_vf(v1x, v1y, v2x, v2y)                      # :86
    dist = np.sqrt((cjx - cix) ** 2 + (cjy - ciy) ** 2)           # :88
    with np.errstate(divide="ignore", invalid="ignore"):
        vis_angle = 2 * np.arctan(r / (1 * dist))                 # :99
    ca_safe = np.where(np.isfinite(ca), ca, 0.0)
    k = nearest_bin(phis, ca_safe)                                # :102
    f0 = int(nearest_bin(phis, np.float64(cfg.fov[0])))           # :105
    f1 = int(nearest_bin(phis, np.float64(cfg.fov[1])))
    proj_size = (vis_angle / TWO_PI) * R                          # :114
    h = np.floor(np.where(np.isfinite(proj_size), proj_size, 0.0) / 2).astype(np.int64)
    ps = k - h                                                    # :116-117
    pe = k + h
    drawn = valid & (((f0 < ps) & (ps < f1)) | ((f0 < pe) & (pe < f1)))   # :119
    return dict(k=k, h=h, ps=ps, pe=pe, drawn=drawn, valid=valid, ca=ca, dist=dist, fov_px=(f0, f1))


def _fill_segments(R, seg_a, seg_b):
    """Union of half-open segments [a, b) on a length-R row (bool)."""
    delta = np.zeros(R + 1, np.int64)
    keep = seg_b > seg_a
    np.add.at(delta, seg_a[keep], 1)
    np.add.at(delta, seg_b[keep], -1)
    return np.cumsum(delta[:R]) > 0


def vf_row(px, py, r, theta, i: int, cfg: VFConfig) -> np.ndarray:
    """Un-flipped union row V (bool, R) of focal agent i: the OR over objects of the
    rows built at vf_supcalc.py:121-129 (wrap rule: SURVEY A.1)."""
    R = cfg.R
    d = vf_intervals(px, py, r, theta, i, cfg)
    sel = d["drawn"]
    ps, pe = d["ps"][sel], d["pe"][sel]
    a_list, b_list = [], []
    neg = ps < 0                                                  # :122-124
    a_list.append(np.clip(R + ps[neg], 0, R)); b_list.append(np.full(neg.sum(), R))
    ps = np.where(neg, 0, ps)
    big = pe >= R                                                 # :125-127
    a_list.append(np.zeros(big.sum(), np.int64)); b_list.append(np.clip(pe[big] - (R - 1), 0, R))
    pe = np.where(big, R, pe)
    a_list.append(np.clip(ps, 0, R)); b_list.append(np.clip(pe, 0, R))   # :129
    return _fill_segments(R, np.concatenate(a_list).astype(np.int64), np.concatenate(b_list).astype(np.int64))


def vf_rows_per_object(px, py, r, theta, i: int, cfg: VFConfig) -> np.ndarray:
    """(N-1, R) STORED (flipped) rows, one per object j != i, i.e. the return value of
    vf_supcalc.projection_field (:134) for the object list of vf_agent.py:219-232."""
    R = cfg.R
    d = vf_intervals(px, py, r, theta, i, cfg)
    N = len(d["k"])
    out = np.zeros((N, R), bool)
    for j in range(N):
        if not d["drawn"][j]:
            continue
        ps, pe = int(d["ps"][j]), int(d["pe"][j])
        if ps < 0:
            out[j, max(R + ps, 0):R] = True
            ps = 0
        if pe >= R:
            out[j, 0:pe - (R - 1)] = True
            pe = R
        out[j, ps:pe] = True
    keep = np.arange(N) != i
    return out[keep][:, ::-1]


def vf_stored_field(px, py, r, theta, i: int, cfg: VFConfig) -> np.ndarray:
    """agent.soc_v_field of VFAgent.calc_soc_v_proj (vf_agent.py:216-237): flipped union."""
    return vf_row(px, py, r, theta, i, cfg)[::-1]


# --------------------------------------------------------------------------------------
# VF: dV/dphi and the flocking integrals (vf_supcalc.py:161-277)
# --------------------------------------------------------------------------------------

def dphi_v(V: np.ndarray) -> np.ndarray:
    """dPhi_V_of (vf_supcalc.py:257-277): circular first difference; BACKWARD difference
    iff V[0] - V[R-1] > 0, FORWARD otherwise."""
    V = np.asarray(V, np.float64)
    if V[0] - V[-1] > 0:
        return V - np.roll(V, 1)
    return np.roll(V, -1) - V


def vswrm_terms(vel, V_unflipped, cfg: VFConfig, alp0=None, bet0=None, v0=None):
    """VSWRM_flocking_state_variables (vf_supcalc.py:161-254), verbose form.
    ``V_unflipped`` is np.flip(agent.soc_v_field) as passed at vf_agent.py:270.
    ALP2 / BET2 multiply an all-zero dt_V (:199) and are therefore absent.
    Returns (dvel, dpsi, a_blob, a_edge, b_blob, b_edge)."""
    V = np.asarray(V_unflipped, np.float64)
    R = V.shape[0]
    Phi = phi_grid(R)
    A0 = cfg.ALP0 if alp0 is None else alp0                       # :191-196
    B0 = cfg.BET0 if bet0 is None else bet0
    V0 = cfg.V0 if v0 is None else v0
    E = np.square(dphi_v(V))                                      # :205, :210
    G = -V                                                        # :202, :207
    a_blob = A0 * np.trapezoid(np.cos(Phi) * G, Phi)              # :243
    a_edge = A0 * cfg.ALP1 * np.sum(np.cos(Phi) * E)              # :244
    b_blob = B0 * np.trapezoid(np.sin(Phi) * G, Phi)              # :246
    b_edge = B0 * cfg.BET1 * np.sum(np.sin(Phi) * E)              # :247
    dvel = cfg.GAM * (V0 - vel) + a_blob + a_edge                 # :249-251
    dpsi = b_blob + b_edge                                        # :253
    return dvel, dpsi, a_blob, a_edge, b_blob, b_edge


# --------------------------------------------------------------------------------------
# VF: kinematics, walls, torus (vf_agent.py:289-330, :188-204; agent.py:347-394, 605-610)
# --------------------------------------------------------------------------------------

def wrap_heading_once(theta: float) -> float:
    """Agent.prove_orientation (agent.py:605-610): ONE conditional wrap each way."""
    if theta < 0:
        theta = 2 * np.pi + theta
    if theta > np.pi * 2:
        theta = theta - 2 * np.pi
    return theta


def reflect_from_walls(x, y, theta, r, width, height, pad, turn=np.pi / 2):
    """Agent.reflect_from_walls (agent.py:347-394).  The x / y tests use the centre
    BEFORE any fix; heading changes cascade through the four tests.  ``turn``: a VFAgent with lines
    to follow turns back by pi instead of pi / 2 (vf_agent.py:80-129)."""
    bx0, bx1 = pad, pad + width
    by0, by1 = pad, pad + height
    cx, cy = x + r, y + r
    pi = np.pi
    if cx < bx0:
        x = bx0 - r
        if pi / 2 <= theta < pi:
            theta -= turn
        elif pi <= theta <= 3 * pi / 2:
            theta += turn
        theta = wrap_heading_once(theta)
    if cx > bx1:
        x = bx1 - r - 1
        if 3 * pi / 2 <= theta < 2 * pi:
            theta -= turn
        elif 0 <= theta <= pi / 2:
            theta += turn
        theta = wrap_heading_once(theta)
    if cy < by0:
        y = by0 - r
        if pi / 2 <= theta <= pi:
            theta += turn
        elif 0 <= theta < pi / 2:
            theta -= turn
        theta = wrap_heading_once(theta)
    if cy > by1:
        y = by1 - r - 1
        if 3 * pi / 2 <= theta <= 2 * pi:
            theta += turn
        elif pi <= theta < 3 * pi / 2:
            theta -= turn
        theta = wrap_heading_once(theta)
    return x, y, theta


def teleport_torus(x, y, r, width, height, pad):
    """VFAgent.teleport_infinite_arena (vf_agent.py:188-204), asymmetric offsets."""
    bx0, bx1 = pad, pad + width
    by0, by1 = pad, pad + height
    cx, cy = x + r, y + r
    if cx < bx0:
        x = bx1 - r
    elif cx > bx1:
        x = bx0 + r
    if cy < by0:
        y = by1 - r
    elif cy > by1:
        y = by0 + r
    return x, y


def _limit(value, lim):
    """prove_turning / prove_velocity of VFAgent (vf_agent.py:311-330)."""
    s = np.sign(value)
    if s == 0:
        s = +1
    if np.abs(value) > lim:
        value = lim * s
    return value


def vf_move(x, y, theta, vel, r, dvel, dpsi, cfg: VFConfig, lines=False):
    """VFAgent.update_agent_position (vf_agent.py:289-309) followed by the boundary
    rule of VFAgent.update (:71-74)."""
    if cfg.limit_movement:
        dpsi = _limit(dpsi, cfg.max_th)
    theta = wrap_heading_once(theta + dpsi)
    vel = vel + dvel
    if cfg.limit_movement:
        vel = _limit(vel, cfg.max_vel)
    x = x + vel * np.cos(theta)
    y = y - vel * np.sin(theta)
    if cfg.boundary == "walls":
        x, y, theta = reflect_from_walls(x, y, theta, r, cfg.width, cfg.height, cfg.window_pad,
                                         np.pi if lines else np.pi / 2)
    elif cfg.boundary == "infinite":
        x, y = teleport_torus(x, y, r, cfg.width, cfg.height, cfg.window_pad)
    return x, y, theta, vel


def follow_lines_local(pos, radius, orientation, linemap, vel, sensor_radius=10, sensor_distance=5):
    """vf_supcalc.follow_lines_local (vf_supcalc.py:293-328): two sensors ahead-left / ahead-right of the agent read the
    mean of the line map in a square window; the heading change steers towards the brighter one.  The windows are numpy
    basic slices of int()-truncated bounds (negative bounds wrap, out-of-range ones clip; an empty window -> nan -> 0)."""
    a34 = 3 * np.pi / 4
    s1 = [pos[1] + radius - sensor_distance + (1 + np.sin(orientation + a34)) * sensor_distance,       # :295-298
          pos[0] + radius - sensor_distance + (1 - np.cos(orientation + a34)) * sensor_distance]
    s2 = [pos[1] + radius - sensor_distance + (1 + np.sin(orientation - a34)) * sensor_distance,       # :299-302
          pos[0] + radius - sensor_distance + (1 - np.cos(orientation - a34)) * sensor_distance]

    def window_mean(s):
        d0, d1 = linemap.shape
        a, b = _slice_bounds(int(s[1] - sensor_radius), int(s[1] + sensor_radius), d0)                # :310-312
        c, d = _slice_bounds(int(s[0] - sensor_radius), int(s[0] + sensor_radius), d1)
        if b <= a or d <= c:
            return np.nan
        w = linemap[a:b, c:d]
        ok = ~np.isnan(w)
        return w[ok].sum() / ok.sum() if ok.any() else np.nan                                          # np.nanmean

    m1, m2 = window_mean(s1), window_mean(s2)
    if np.isnan(m1) or np.isnan(m2):                                                                   # :314-315
        return 0.0
    ori_change = 0.5 * (m2 - m1) if np.sign(vel) else 0.0                                              # :317-320
    if m1 != m2:                                                                                       # :321-324
        return ori_change
    return 0.01 if m1 != 0 else 0.0                                                                    # :326-329


def _slice_bounds(a: int, b: int, n: int):
    """Effective [lo, hi) of the numpy basic slice [a:b] on an axis of length n."""
    if a < 0:
        a = max(a + n, 0)
    if b < 0:
        b = max(b + n, 0)
    return min(a, n), min(b, n)


def vf_step_frozen(px, py, theta, vel, r, cfg: VFConfig, alp0=None, bet0=None, v0=None,
                   agents=None, line_map=None, sensor_radius=9, sensor_distance=20):
    """One synchronous (Jacobi) step: every agent is updated from the SAME frozen
    snapshot -- the per-agent parity definition of SURVEY 8c (the reference itself
    updates agents sequentially in place, vf_sims.py:302; see vf_step_sequential).

    ``line_map``: an agent with lines to follow (vf_agent.py:273-276; sensor defaults :30-31) takes its heading change
    from follow_lines_local instead of the flocking integral (the speed change stays the flocking one).

    Returns dict(x, y, theta, vel, rows (N,R bool un-flipped), terms (N,6))."""
    px = np.asarray(px, np.float64); py = np.asarray(py, np.float64)
    theta = np.asarray(theta, np.float64); vel = np.asarray(vel, np.float64)
    N = px.shape[0]
    rr = np.broadcast_to(np.asarray(r, np.float64), (N,))
    idx = range(N) if agents is None else agents
    nx, ny, nt, nv = px.copy(), py.copy(), theta.copy(), vel.copy()
    rows = np.zeros((N, cfg.R), bool)
    terms = np.zeros((N, 6))
    for i in idx:
        rows[i] = vf_row(px, py, rr, theta, i, cfg)
        t = vswrm_terms(vel[i], rows[i].astype(np.float64), cfg,
                        None if alp0 is None else alp0[i],
                        None if bet0 is None else bet0[i],
                        None if v0 is None else v0[i])
        terms[i] = t
        dpsi = t[1]
        if line_map is not None:
            dpsi = follow_lines_local((px[i], py[i]), rr[i], theta[i], line_map, vel[i], sensor_radius, sensor_distance)
        nx[i], ny[i], nt[i], nv[i] = vf_move(px[i], py[i], theta[i], vel[i], rr[i], t[0], dpsi, cfg,
                                             lines=line_map is not None)
    return dict(x=nx, y=ny, theta=nt, vel=nv, rows=rows, terms=terms)


def vf_step_sequential(px, py, theta, vel, r, cfg: VFConfig, alp0=None, bet0=None, v0=None):
    """Reference-literal update order (Gauss-Seidel): agent i sees the already-moved
    agents j < i (pygame Group.update, vf_sims.py:302).  Reported, not gated on."""
    px = np.array(px, np.float64); py = np.array(py, np.float64)
    theta = np.array(theta, np.float64); vel = np.array(vel, np.float64)
    N = px.shape[0]
    rr = np.broadcast_to(np.asarray(r, np.float64), (N,))
    for i in range(N):
        row = vf_row(px, py, rr, theta, i, cfg)
        t = vswrm_terms(vel[i], row.astype(np.float64), cfg,
                        None if alp0 is None else alp0[i],
                        None if bet0 is None else bet0[i],
                        None if v0 is None else v0[i])
        px[i], py[i], theta[i], vel[i] = vf_move(px[i], py[i], theta[i], vel[i], rr[i], t[0], t[1], cfg)
    return dict(x=px, y=py, theta=theta, vel=vel)


# --------------------------------------------------------------------------------------
# packing helpers shared with the tests (engine dumps 32-bin words, stored order)
# --------------------------------------------------------------------------------------

def pack_bits(field_bool: np.ndarray) -> np.ndarray:
    """(…, R) bool -> (…, ceil(R/32)) uint32, bin b in bit (b & 31) of word b >> 5."""
    R = field_bool.shape[-1]
    W = (R + 31) // 32
    pad = np.zeros(field_bool.shape[:-1] + (W * 32 - R,), bool)
    bits = np.concatenate([field_bool, pad], axis=-1).reshape(field_bool.shape[:-1] + (W, 32))
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    return (bits.astype(np.uint64) * weights).sum(axis=-1).astype(np.uint32)


def unpack_bits(words: np.ndarray, R: int) -> np.ndarray:
    words = np.asarray(words, np.uint32)
    bits = (words[..., :, None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)
    return bits.reshape(words.shape[:-1] + (-1,))[..., :R].astype(bool)


def runs_of(field_bool: np.ndarray):
    """[start, end) index pairs of the 1-runs of a 1-D bool array (SURVEY App. B format)."""
    d = np.diff(np.concatenate([[0], field_bool.astype(np.int8), [0]]))
    return list(zip(np.where(d == 1)[0].tolist(), np.where(d == -1)[0].tolist()))
