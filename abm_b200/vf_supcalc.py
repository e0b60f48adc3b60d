"""Drop-in replacements for the function-level interface of the reference's
abm/projects/visual_flocking/vf_agent/vf_supcalc.py -- same names, arguments and return
shapes -- evaluated by the CUDA kernels of libabm_b200.so (no CPU fallback).

  projection_field                  vf_supcalc.py:20-138
  VSWRM_flocking_state_variables    vf_supcalc.py:161-254
  dPhi_V_of                         vf_supcalc.py:257-277
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _pack(field) -> np.ndarray:
    v = np.asarray(field) > 0
    R = v.shape[-1]
    W = (R + 31) // 32
    bits = np.zeros(v.shape[:-1] + (W * 32,), np.uint64)
    bits[..., :R] = v
    words = (bits.reshape(v.shape[:-1] + (W, 32)) << np.arange(32, dtype=np.uint64)).sum(axis=-1)
    return np.ascontiguousarray(words.astype(np.uint32))


def _unpack(words: np.ndarray, R: int) -> np.ndarray:
    bits = (words[..., :, None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)
    return bits.reshape(words.shape[:-1] + (-1,))[..., :R]


def projection_field(fov, v_field_resolution, position, radius, orientation, object_positions,
                     object_sizes=None, boundary_cond="walls", arena_width=None, arena_height=None,
                     vision_range=None, ag_id=0):
    """Visual projection field of one agent: ndarray (n objects, field resolution) of 0./1.,
    flipped along the second axis like the reference's return value."""
    lib = _lib.load()
    R = int(v_field_resolution)
    n = len(object_positions)
    W = (R + 31) // 32
    if n == 0:
        return np.zeros((0, R))
    if boundary_cond == "infinite" and (arena_width is None or arena_height is None):
        raise TypeError("arena_width / arena_height are required for boundary_cond='infinite'")
    ox = np.ascontiguousarray([float(p[0]) for p in object_positions], np.float64)
    oy = np.ascontiguousarray([float(p[1]) for p in object_positions], np.float64)
    osz = None if object_sizes is None else np.ascontiguousarray(object_sizes, np.float64)
    dp = C.POINTER(C.c_double)
    args = _lib.VFProjArgs(
        struct_size=C.sizeof(_lib.VFProjArgs), resolution=R, fov0=float(fov[0]), fov1=float(fov[1]),
        x=float(position[0]), y=float(position[1]), radius=float(radius), orientation=float(orientation),
        n_obj=n, obj_x=ox.ctypes.data_as(dp), obj_y=oy.ctypes.data_as(dp),
        obj_size=None if osz is None else osz.ctypes.data_as(dp),
        boundary=_lib.BOUNDARY_INFINITE if boundary_cond == "infinite" else _lib.BOUNDARY_WALLS,
        arena_width=float(arena_width or 0.0), arena_height=float(arena_height or 0.0),
        vision_range=-1.0 if vision_range is None else float(vision_range))
    rows = np.empty((n, W), np.uint32)
    _lib.check(lib.abm_vf_projection_field(C.byref(args), C.c_void_p(rows.ctypes.data)), "abm_vf_projection_field")
    return _unpack(rows, R).astype(np.float64)


def VSWRM_flocking_state_variables(vel_now, Phi, V_now, vf_params, t_now=None, V_prev=None, t_prev=None,
                                   verbose=False, ALP0=None, BET0=None, V0=None):
    """dvel, dpsi (verbose: + alpha_blob, alpha_edge, beta_blob, beta_edge) of one agent.
    ``Phi`` must be the reference's grid np.arange(-pi, pi, 2pi/len(V_now)) (vf_agent.py:44);
    ``vf_params`` any object with GAM, V0, ALP0, ALP1, BET0, BET1 attributes."""
    lib = _lib.load()
    V = np.asarray(V_now)
    R = V.shape[0]
    if len(Phi) != R:
        raise ValueError("Phi and V_now must have the same length")
    prm = np.array([vf_params.GAM,
                    vf_params.V0 if V0 is None else V0,
                    vf_params.ALP0 if ALP0 is None else ALP0,
                    vf_params.ALP1,
                    vf_params.BET0 if BET0 is None else BET0,
                    vf_params.BET1], np.float64)
    packed = _pack(V)
    out = (C.c_double * 6)()
    _lib.check(lib.abm_vf_flocking_terms(C.c_void_p(packed.ctypes.data), R, float(vel_now),
                                         C.c_void_p(prm.ctypes.data), out), "abm_vf_flocking_terms")
    if not verbose:
        return out[0], out[1]
    return tuple(out)


def dPhi_V_of(Phi, V):
    """Derivative of the visual projection field w.r.t. the visual angle: circular first
    difference with the reference's forward / backward rule (vf_supcalc.py:257-277)."""
    lib = _lib.load()
    Vb = np.asarray(V)
    R = Vb.shape[0]
    packed = _pack(Vb)
    out = np.zeros(R, np.int8)
    _lib.check(lib.abm_vf_dphi(C.c_void_p(packed.ctypes.data), R, C.c_void_p(out.ctypes.data)), "abm_vf_dphi")
    return out.astype(np.float64)
