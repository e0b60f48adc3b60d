"""ctypes binding of libabm_b200.so (C ABI declared in include/abm_b200.h).

There is no CPU fallback: if the library is missing, or no sm_100 device is
usable, the calls raise -- they never route to another implementation.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libabm_b200.so")

ABM_OK = 0
BOUNDARY_WALLS = 0
BOUNDARY_INFINITE = 1
VF_EXACT_FIXUP = 1 << 0
VF_KEEP_FIELDS = 1 << 1
VF_KEEP_TERMS = 1 << 2
VF_SPATIAL_SORT = 1 << 3
VF_NPARAM = 6
VF_IPC_BYTES = 512
VF_TILE_BLOCK = 128


class AbmError(RuntimeError):
    def __init__(self, code, where, detail):
        super().__init__(f"{where} failed with code {code}: {detail}")
        self.code = code


class VFConfig(C.Structure):
    """abm_vf_config_t"""
    _fields_ = [
        ("struct_size", C.c_int32), ("n_replicates", C.c_int32), ("n_agents", C.c_int32),
        ("resolution", C.c_int32), ("fov_px0", C.c_int32), ("fov_px1", C.c_int32),
        ("boundary", C.c_int32), ("limit_movement", C.c_int32),
        ("width", C.c_float), ("height", C.c_float), ("window_pad", C.c_float),
        ("max_vel", C.c_float), ("max_th", C.c_float), ("flags", C.c_uint32),
        ("tile_begin", C.c_int32), ("tile_count", C.c_int32), ("resort_every", C.c_int32),
        ("tile_cycle", C.c_int32), ("tile_phase", C.c_int32),
    ]


class VFProjArgs(C.Structure):
    """abm_vf_proj_args_t"""
    _fields_ = [
        ("struct_size", C.c_int32), ("resolution", C.c_int32),
        ("fov0", C.c_double), ("fov1", C.c_double),
        ("x", C.c_double), ("y", C.c_double), ("radius", C.c_double), ("orientation", C.c_double),
        ("n_obj", C.c_int32),
        ("obj_x", C.POINTER(C.c_double)), ("obj_y", C.POINTER(C.c_double)), ("obj_size", C.POINTER(C.c_double)),
        ("boundary", C.c_int32),
        ("arena_width", C.c_double), ("arena_height", C.c_double), ("vision_range", C.c_double),
    ]


class CSProjArgs(C.Structure):
    """abm_cs_proj_args_t"""
    _fields_ = [
        ("struct_size", C.c_int32), ("resolution", C.c_int32),
        ("fov0", C.c_double), ("fov1", C.c_double),
        ("x", C.c_double), ("y", C.c_double), ("radius", C.c_double), ("orientation", C.c_double),
        ("n_obj", C.c_int32),
        ("obj_x", C.POINTER(C.c_double)), ("obj_y", C.POINTER(C.c_double)),
        ("max_proj_size", C.c_double),
    ]


class BaseConfig(C.Structure):
    """abm_base_config_t"""
    _fields_ = [
        ("struct_size", C.c_int32), ("n_replicates", C.c_int32), ("n_agents", C.c_int32), ("n_patches", C.c_int32),
        ("resolution", C.c_int32), ("tau", C.c_int32), ("visual_exclusion", C.c_int32),
        ("patchwise_exclusion", C.c_int32), ("teleport_exploit", C.c_int32), ("regenerate_patches", C.c_int32),
        ("patch_border_overlap", C.c_int32), ("keep_fields", C.c_int32),
        ("collide_agents", C.c_int32), ("ghost_mode", C.c_int32),
        ("fov0", C.c_double), ("fov1", C.c_double),
        ("width", C.c_double), ("height", C.c_double), ("window_pad", C.c_double),
        ("vision_range", C.c_double), ("agent_radius", C.c_double), ("patch_radius", C.c_double),
        ("min_quality", C.c_double), ("max_quality", C.c_double),
        ("min_units", C.c_int32), ("max_units", C.c_int32), ("seed", C.c_uint64),
    ]


BASE_AGENT_FIELDS = [("x", "f"), ("y", "f"), ("theta", "f"), ("vel", "f"), ("w", "f"), ("u", "f"),
                     ("collected", "f"), ("collected_before", "f"), ("i_priv", "f"),
                     ("env_status", "i"), ("override_mode", "i"), ("mode", "i"), ("patch_id", "i"), ("novelty", "u")]
BASE_PATCH_FIELDS = [("x", "f"), ("y", "f"), ("radius", "f"), ("left", "f"), ("quality", "f"), ("id", "i")]
BASE_NPARAM = 20
BASE_PARAM_NAMES = ["T_w", "Eps_w", "g_w", "B_w", "w_max", "T_u", "Eps_u", "g_u", "B_u", "u_max", "S_wu", "S_uw",
                    "F_N", "F_R", "exp_vel_max", "exp_theta_min", "exp_theta_max", "reloc_theta_max",
                    "exp_stop_ratio", "agent_consumption"]


class BaseProjArgs(C.Structure):
    """abm_base_proj_args_t"""
    _fields_ = [
        ("struct_size", C.c_int32), ("resolution", C.c_int32), ("fov0", C.c_double), ("fov1", C.c_double),
        ("x", C.c_double), ("y", C.c_double), ("radius", C.c_double), ("orientation", C.c_double),
        ("n_social", C.c_int32), ("n_occluders", C.c_int32),
        ("social_x", C.c_void_p), ("social_y", C.c_void_p), ("occluder_x", C.c_void_p), ("occluder_y", C.c_void_p),
        ("visual_exclusion", C.c_int32), ("keep_distance_info", C.c_int32), ("vision_range", C.c_double),
    ]


class BaseAgents(C.Structure):
    """abm_base_agents_t"""
    _fields_ = [(n, C.c_void_p) for n, _ in BASE_AGENT_FIELDS]


class BasePatches(C.Structure):
    """abm_base_patches_t"""
    _fields_ = [(n, C.c_void_p) for n, _ in BASE_PATCH_FIELDS]


# every symbol include/abm_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "abm_version": (C.c_int, []),
    "abm_last_error": (C.c_char_p, []),
    "abm_device_count": (C.c_int, []),
    "abm_field_words": (C.c_int, [C.c_int]),
    "abm_vf_create": (C.c_int, [C.POINTER(VFConfig), C.c_int, C.POINTER(_P)]),
    "abm_destroy": (C.c_int, [_P]),
    "abm_vf_set_params": (C.c_int, [_P, _P, C.c_int]),
    "abm_vf_set_agent_overrides": (C.c_int, [_P, _P, _P, _P, C.c_int, _P]),
    "abm_set_state": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "abm_get_state": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P]),
    "abm_set_state_packed": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "abm_get_state_packed": (C.c_int, [_P, _P, C.c_int, _P]),
    "abm_vf_step_host": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "abm_vf_set_line_map": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _P]),
    "abm_vf_step": (C.c_int, [_P, C.c_int, _P]),
    "abm_get_fields": (C.c_int, [_P, _P, C.c_int, _P]),
    "abm_vf_get_terms": (C.c_int, [_P, _P, C.c_int, _P]),
    "abm_get_counters": (C.c_int, [_P, C.POINTER(C.c_uint64), _P]),
    "abm_vf_record_table": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int)]),
    "abm_vf_last_kernel": (C.c_char_p, [_P]),
    "abm_vf_kernel_stats": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "abm_vf_cluster_launches": (C.c_int, [_P, _P]),
    "abm_vf_cluster_launches": (C.c_int, [_P, _P]),
    "abm_vf_ipc_export": (C.c_int, [_P, _P]),
    "abm_vf_ipc_attach": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "abm_vf_metrics": (C.c_int, [_P, _P, C.c_int, _P]),
    "abm_vf_slow_entries": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _P]),
    "abm_vf_get_permutation": (C.c_int, [_P, _P, C.c_int, _P]),
    "abm_vf_resort": (C.c_int, [_P, _P]),
    "abm_vf_internal_arrays": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "abm_synchronize": (C.c_int, [_P, _P]),
    "abm_vf_projection_field": (C.c_int, [C.POINTER(VFProjArgs), _P]),
    "abm_cs_projection_field": (C.c_int, [C.POINTER(CSProjArgs), _P]),
    "abm_vf_flocking_terms": (C.c_int, [_P, C.c_int, C.c_double, _P, C.POINTER(C.c_double)]),
    "abm_base_create": (C.c_int, [C.POINTER(BaseConfig), C.c_int, C.POINTER(_P)]),
    "abm_base_destroy": (C.c_int, [_P]),
    "abm_base_set_params": (C.c_int, [_P, _P, C.c_int]),
    "abm_base_set_agent_geometry": (C.c_int, [_P, _P, _P, _P, C.c_int]),
    "abm_base_set_agent_radii": (C.c_int, [_P, _P, C.c_int]),
    "abm_base_set_agent_resolution": (C.c_int, [_P, _P, C.c_int]),
    "abm_base_set_agents": (C.c_int, [_P, C.POINTER(BaseAgents), C.c_int, _P]),
    "abm_base_get_agents": (C.c_int, [_P, C.POINTER(BaseAgents), C.c_int, _P]),
    "abm_base_set_patches": (C.c_int, [_P, C.POINTER(BasePatches), C.c_int, _P]),
    "abm_base_get_patches": (C.c_int, [_P, C.POINTER(BasePatches), C.c_int, _P]),
    "abm_base_step": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_uint32, _P]),
    "abm_base_inject_regeneration": (C.c_int, [_P, _P, C.c_int]),
    "abm_base_set_regeneration_params": (C.c_int, [_P, _P, C.c_int]),
    "abm_base_get_fields": (C.c_int, [_P, _P, C.c_int, _P]),
    "abm_base_get_counters": (C.c_int, [_P, C.POINTER(C.c_uint64), _P]),
    "abm_base_metrics": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "abm_base_projection_field": (C.c_int, [C.POINTER(BaseProjArgs), _P, C.POINTER(C.c_double)]),
    "abm_base_reloc_lr": (C.c_int, [_P, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.POINTER(C.c_double)]),
    "abm_vf_dphi": (C.c_int, [_P, C.c_int, _P]),
}

_lib = None


def load():
    """Load the shared library (once) and declare the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(abm_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError here == header / library skew
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, where):
    if rc != ABM_OK:
        detail = load().abm_last_error()
        raise AbmError(rc, where, detail.decode() if detail else "")
