"""Drop-in replacements for the function-level interface of the reference's BASE agent
(abm/agent/agent.py, abm/agent/supcalc.py), evaluated by CUDA kernels (no CPU fallback).

  projection_field   Agent.projection_field (agent.py:457-597) as a free function
  F_reloc_LR         supcalc.F_reloc_LR (supcalc.py:81-92)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .vf_supcalc import _pack, _unpack


def projection_field(position, radius, orientation, v_field_res, FOV, obstacles, keep_distance_info=False,
                     non_expl_agents=None, fov=None, visual_exclusion=False, vision_range=None):
    """Agent.projection_field: ``obstacles`` / ``non_expl_agents`` are sequences of (x, y)
    positions (what the reference reads from ``ob.position``); the agent attributes the method
    reads (position, radius, orientation, v_field_res, FOV, visual_exclusion, vision_range) are
    arguments.  Returns the ndarray (v_field_res,) of agent.py:597 (flipped, FOV-masked; values
    1 or, with keep_distance_info, 1 - distance / vision_range)."""
    lib = _lib.load()
    fov = FOV if fov is None else fov
    R = int(v_field_res)
    soc = np.ascontiguousarray(np.asarray(obstacles, np.float64).reshape(-1, 2))
    occ = np.zeros((0, 2)) if non_expl_agents is None else \
        np.ascontiguousarray(np.asarray(non_expl_agents, np.float64).reshape(-1, 2))
    if keep_distance_info and vision_range is None:
        raise TypeError("vision_range is required with keep_distance_info")
    sx, sy = np.ascontiguousarray(soc[:, 0]), np.ascontiguousarray(soc[:, 1])
    ox, oy = np.ascontiguousarray(occ[:, 0]), np.ascontiguousarray(occ[:, 1])
    args = _lib.BaseProjArgs(
        struct_size=C.sizeof(_lib.BaseProjArgs), resolution=R, fov0=float(fov[0]), fov1=float(fov[1]),
        x=float(position[0]), y=float(position[1]), radius=float(radius), orientation=float(orientation),
        n_social=len(soc), n_occluders=len(occ) if non_expl_agents is not None else 0,
        social_x=sx.ctypes.data, social_y=sy.ctypes.data, occluder_x=ox.ctypes.data, occluder_y=oy.ctypes.data,
        visual_exclusion=int(bool(visual_exclusion)), keep_distance_info=int(bool(keep_distance_info)),
        vision_range=float(vision_range or 0.0))
    words = np.zeros((R + 31) // 32, np.uint32)
    amp = C.c_double(1.0)
    _lib.check(lib.abm_base_projection_field(C.byref(args), C.c_void_p(words.ctypes.data), C.byref(amp)),
               "abm_base_projection_field")
    return _unpack(words, R).astype(np.float64) * amp.value


def F_reloc_LR(vel_now, V_now, v_desired=None, reloc_theta_max=0.5, reloc_des_vel=1.0):
    """supcalc.F_reloc_LR; ``reloc_theta_max`` / ``reloc_des_vel`` stand for the module constants
    movement_params.reloc_theta_max / reloc_des_vel the reference reads (supcalc.py:84-90)."""
    lib = _lib.load()
    V = np.asarray(V_now, np.float64)
    nz = V[V != 0]
    amp = float(nz[0]) if nz.size else 1.0
    if nz.size and not np.all(nz == amp):
        raise ValueError("V_now must be a binary field times one amplitude (what projection_field returns)")
    if v_desired is None:
        v_desired = reloc_des_vel
    out = (C.c_double * 2)()
    packed = _pack(V)
    _lib.check(lib.abm_base_reloc_lr(C.c_void_p(packed.ctypes.data), V.shape[0], amp, float(vel_now),
                                     float(v_desired), float(reloc_theta_max), out), "abm_base_reloc_lr")
    return out[0], out[1]
