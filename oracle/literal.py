"""TEST INFRASTRUCTURE ONLY -- per-pair, loop-structured float64 port of the reference's
visual-flocking agent update, used as the CPU BASELINE (bench.py ``cpu_baseline`` and
``--impl reference``) and as a second, independently structured checker.

Unlike oracle/restate.py (vectorised closed forms) this port keeps the reference's
*cost structure*: one Python iteration per (focal, object) pair, three arg-min scans over
the R bin centres per pair (supcalc.py:8-11 via vf_supcalc.py:102-105), an (N-1, R)
float64 temporary per agent (vf_supcalc.py:38) and numpy scalar calls for every norm /
arccos / arctan -- so its speed is representative of the reference's own CPU path
(SURVEY section 6 measured 40-90 us per pair for the reference; this port lands in the
same range).  It is a restatement, not a copy: names, decomposition and control flow are
ours; the arithmetic follows the cited reference lines so results agree to the last bit.
"""
from __future__ import annotations

import numpy as np

from .restate import VFConfig, reflect_from_walls, teleport_torus, wrap_heading_once

PI2 = 2 * np.pi


def _nearest(grid: np.ndarray, value: float) -> int:
    """supcalc.find_nearest (supcalc.py:8-11): O(R) scan, first minimum."""
    return int(np.abs(grid - value).argmin())


def _signed_angle(a: np.ndarray, b: np.ndarray) -> float:
    """supcalc.angle_between (supcalc.py:19-34)."""
    au = a / np.linalg.norm(a)
    bu = b / np.linalg.norm(b)
    ang = np.arccos(np.clip(np.dot(au, bu), -1.0, 1.0))
    return -ang if au[0] * bu[1] - au[1] * bu[0] < 0 else ang


def _closed_angle(a: np.ndarray, b: np.ndarray) -> float:
    """vf_supcalc.calculate_closed_angle (vf_supcalc.py:142-158)."""
    ang = _signed_angle(a, b) % PI2
    return -ang if 0 <= ang <= np.pi else PI2 - ang


def project_objects(fov, R, focal_pos, focal_r, heading, obj_pos, obj_r, boundary, W, H):
    """vf_supcalc.projection_field (vf_supcalc.py:20-138): (n_obj, R) float64, flipped."""
    out = np.zeros((len(obj_pos), R))
    grid = np.linspace(-np.pi, np.pi, R)
    centre = focal_pos + focal_r
    rim = focal_pos + np.array([1 + np.cos(heading), 1 - np.sin(heading)]) * focal_r
    ahead = rim - centre
    for j in range(len(obj_pos)):
        p = obj_pos[j]
        if p[0] == focal_pos[0] and p[1] == focal_pos[1]:
            continue
        oc = p + obj_r[j]
        sep = oc - centre
        if boundary == "infinite":
            if np.abs(sep[0]) > W / 2:
                if centre[0] < oc[0]:
                    oc[0] -= W
                elif centre[0] > oc[0]:
                    oc[0] += W
            if np.abs(sep[1]) > H / 2:
                if centre[1] < oc[1]:
                    oc[1] -= H
                elif centre[1] > oc[1]:
                    oc[1] += H
            sep = oc - centre
        ca = _closed_angle(ahead, sep)
        dist = np.linalg.norm(oc - centre)
        subtended = 2 * np.arctan(obj_r[j] / (1 * dist))
        k = _nearest(grid, ca)
        lo_px, hi_px = _nearest(grid, fov[0]), _nearest(grid, fov[1])
        width = (subtended / PI2) * R
        a = int(k - np.floor(width / 2))
        b = int(k + np.floor(width / 2))
        if lo_px < a < hi_px or lo_px < b < hi_px:
            if a < 0:
                out[j, R + a:R] = 1
                a = 0
            if b >= R:
                out[j, 0:b - (R - 1)] = 1
                b = R
            out[j, a:b] = 1
    return np.flip(out, axis=1)


def ring_derivative(V: np.ndarray) -> np.ndarray:
    """vf_supcalc.dPhi_V_of (vf_supcalc.py:257-277)."""
    ext = np.pad(V, (1, 1), "wrap")
    d = np.diff(ext)
    return d[0:-1] if (d[0] > 0 and d[-1] > 0) else d[1:]


def flocking_increments(vel, Phi, V, cfg: VFConfig, alp0=None, bet0=None, v0=None):
    """vf_supcalc.VSWRM_flocking_state_variables (vf_supcalc.py:161-254), non-verbose."""
    A0 = cfg.ALP0 if alp0 is None else alp0
    B0 = cfg.BET0 if bet0 is None else bet0
    V0 = cfg.V0 if v0 is None else v0
    spikes = np.square(ring_derivative(V))
    G = -V + 0.0 * np.zeros(len(Phi))
    dvel = cfg.GAM * (V0 - vel) + A0 * np.trapezoid(np.cos(Phi) * G, Phi) + A0 * cfg.ALP1 * np.sum(np.cos(Phi) * spikes)
    dpsi = B0 * np.trapezoid(np.sin(Phi) * G, Phi) + B0 * cfg.BET1 * np.sum(np.sin(Phi) * spikes)
    return dvel, dpsi


def agent_update(i, x, y, theta, vel, radius, cfg: VFConfig, alp0=None, bet0=None, v0=None):
    """VFAgent.update (vf_agent.py:52-80) for agent i from the frozen snapshot (x, y, theta,
    vel, radius arrays).  Returns (stored_field, new_x, new_y, new_theta, new_vel)."""
    N = len(x)
    others = [j for j in range(N) if j != i]                          # vf_agent.py:220
    obj_pos = [np.array([x[j], y[j]], dtype=np.float64) for j in others]
    obj_r = [radius[j] for j in others]
    rows = project_objects(cfg.fov, cfg.R, np.array([x[i], y[i]], dtype=np.float64), radius[i], theta[i],
                           obj_pos, obj_r, cfg.boundary, cfg.width, cfg.height)
    field = rows.sum(axis=0)                                          # vf_agent.py:234-236
    field[field > 0] = 1
    Phi = np.arange(-np.pi, np.pi, PI2 / cfg.R)                       # vf_agent.py:44
    dvel, dpsi = flocking_increments(vel[i], Phi, np.flip(field), cfg,
                                     None if alp0 is None else alp0[i],
                                     None if bet0 is None else bet0[i],
                                     None if v0 is None else v0[i])
    if cfg.limit_movement:
        s = np.sign(dpsi) or 1.0
        if np.abs(dpsi) > cfg.max_th:
            dpsi = cfg.max_th * s
    th = wrap_heading_once(theta[i] + dpsi)
    v = vel[i] + dvel
    if cfg.limit_movement:
        s = np.sign(v) or 1.0
        if np.abs(v) > cfg.max_vel:
            v = cfg.max_vel * s
    nx = x[i] + v * np.cos(th)
    ny = y[i] - v * np.sin(th)
    if cfg.boundary == "walls":
        nx, ny, th = reflect_from_walls(nx, ny, th, radius[i], cfg.width, cfg.height, cfg.window_pad)
    else:
        nx, ny = teleport_torus(nx, ny, radius[i], cfg.width, cfg.height, cfg.window_pad)
    return field, nx, ny, th, v
