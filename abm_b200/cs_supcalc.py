"""Drop-in replacement for the projection function of the reference's
abm/projects/cooperative_signaling/cs_agent/cs_supcalc.py -- same name, arguments and return
shape -- evaluated by a CUDA kernel of libabm_b200.so (no CPU fallback).

  projection_field        cs_supcalc.py:204-289
  calculate_closed_angle  cs_supcalc.py:310-326 is the visual-flocking function (vf_supcalc.py:142-158)

Function level only: the cooperative-signaling simulation loop is not part of the engine.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .vf_supcalc import _unpack


def projection_field(fov, v_field_resolution, position, radius, orientation, object_positions,
                     object_meters=None, max_proj_size=None):
    """Visual projection field of one agent: ndarray (n objects, field resolution), flipped along the
    second axis and limited to the FOV like the reference's return value; rows are 0 / 1, or
    0 / object_meters[i] when meters are given."""
    lib = _lib.load()
    R = int(v_field_resolution)
    n = len(object_positions)
    W = (R + 31) // 32
    if n == 0:
        return np.zeros((0, R))
    ox = np.ascontiguousarray([float(p[0]) for p in object_positions], np.float64)
    oy = np.ascontiguousarray([float(p[1]) for p in object_positions], np.float64)
    dp = C.POINTER(C.c_double)
    args = _lib.CSProjArgs(
        struct_size=C.sizeof(_lib.CSProjArgs), resolution=R, fov0=float(fov[0]), fov1=float(fov[1]),
        x=float(position[0]), y=float(position[1]), radius=float(radius), orientation=float(orientation),
        n_obj=n, obj_x=ox.ctypes.data_as(dp), obj_y=oy.ctypes.data_as(dp),
        max_proj_size=-1.0 if max_proj_size is None else max(float(max_proj_size), 0.0))   # < 0 encodes None
    rows = np.empty((n, W), np.uint32)
    _lib.check(lib.abm_cs_projection_field(C.byref(args), C.c_void_p(rows.ctypes.data)), "abm_cs_projection_field")
    out = _unpack(rows, R).astype(np.float64)
    if object_meters is not None:                     # cs_supcalc.py:283-284
        out *= np.asarray(object_meters, np.float64)[:n, None]
    return out
