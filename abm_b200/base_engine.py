"""Host-side handle of the BASE / collective-foraging engine (thin wrapper over the C ABI).

State arrays are numpy arrays of shape (n_replicates, n_agents) [agents] or
(n_replicates, n_patches) [patches]; see include/abm_b200.h for the field meanings.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import _current_stream

_DT = {"f": np.float32, "i": np.int32, "u": np.uint32}
PHASE_ENV, PHASE_AGENTS, PHASE_COLLISIONS, PHASE_ALL = 1, 2, 4, 7


class BaseEngine:
    """B replicates x N agents (+ P resource patches each) of the foraging model on one GPU.
    Mirrors the state of Simulation / Agent / Rescource that the main loop touches
    (sims.py:733-864, agent.py:212-283)."""

    def __init__(self, n_replicates: int, n_agents: int, n_patches: int = 0, *, resolution: int = 1200,
                 agent_fov: float = 1.0, width: float = 500.0, height: float = 500.0, window_pad: float = 30.0,
                 vision_range: float = 2000.0, agent_radius: float = 10.0, visual_exclusion: bool = True,
                 patchwise_exclusion: bool = True, teleport_exploit: bool = False, regenerate_patches: bool = True,
                 patch_border_overlap: bool = True, patch_radius: float = 30.0, min_resc_quality: float = 0.25,
                 max_resc_quality: float = -1.0, min_resc_perpatch: int = 100, max_resc_perpatch: int = -1,
                 tau: int = 10, keep_fields: bool = False, collide_agents: bool = False, ghost_mode: bool = True,
                 seed: int = 0, device: int = 0):
        self._lib = _lib.load()
        self.B, self.N, self.P, self.R = int(n_replicates), int(n_agents), int(n_patches), int(resolution)
        self.W = (self.R + 31) // 32
        self._agent_fov, self._vision_range = float(agent_fov), float(vision_range)
        self.tau = int(tau)
        # "negative maximum = use the minimum" rule of sims.py:176-179
        if max_resc_quality < 0:
            max_resc_quality = min_resc_quality
        if max_resc_perpatch < 0:
            max_resc_perpatch = min_resc_perpatch + 1
        cfg = _lib.BaseConfig(
            struct_size=C.sizeof(_lib.BaseConfig), n_replicates=self.B, n_agents=self.N, n_patches=self.P,
            resolution=self.R, tau=self.tau, visual_exclusion=int(bool(visual_exclusion)),
            patchwise_exclusion=int(bool(patchwise_exclusion)), teleport_exploit=int(bool(teleport_exploit)),
            regenerate_patches=int(bool(regenerate_patches)), patch_border_overlap=int(bool(patch_border_overlap)),
            keep_fields=int(bool(keep_fields)), collide_agents=int(bool(collide_agents)),
            ghost_mode=int(bool(ghost_mode)),
            fov0=-float(agent_fov) * np.pi, fov1=float(agent_fov) * np.pi,      # sims.py:160-161
            width=float(width), height=float(height), window_pad=float(window_pad),
            vision_range=float(vision_range), agent_radius=float(agent_radius), patch_radius=float(patch_radius),
            min_quality=float(min_resc_quality), max_quality=float(max_resc_quality),
            min_units=int(min_resc_perpatch), max_units=int(max_resc_perpatch), seed=int(seed))
        self._h = C.c_void_p()
        _lib.check(self._lib.abm_base_create(C.byref(cfg), int(device), C.byref(self._h)), "abm_base_create")
        self.keep_fields = keep_fields

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.abm_base_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- parameters ----------------------------------------------------------------------
    def set_params(self, **kw):
        """Decision / movement parameters by name (_lib.BASE_PARAM_NAMES): scalars, length-B arrays (one set
        per replicate, e.g. a DEC_EPSW sweep) or (B, N) arrays (one set per agent: the heterogeneous agents of
        agent.py:83-108 / sims.py:499-517).  Parameters not named keep the value of the previous call (initially the
        reference's param-module defaults), like assigning single attributes of the reference's param modules."""
        if not hasattr(self, "_param_values"):
            self._param_values = dict(T_w=0.5, Eps_w=3, g_w=0.085, B_w=0, w_max=1, T_u=0.5, Eps_u=3, g_u=0.085, B_u=0,
                                      u_max=1, S_wu=0.25, S_uw=0.01, F_N=2, F_R=1, exp_vel_max=1, exp_theta_min=-0.3,
                                      exp_theta_max=0.3, reloc_theta_max=0.5, exp_stop_ratio=0.08, agent_consumption=1)
        unknown = set(kw) - set(self._param_values)
        if unknown:
            raise TypeError(f"unknown parameter(s): {sorted(unknown)}")
        self._param_values.update(kw)
        vals = [np.asarray(self._param_values[n], np.float64) for n in _lib.BASE_PARAM_NAMES]
        for v in vals:
            if v.shape not in ((), (1,), (self.B,), (self.B, self.N)):
                raise ValueError("parameter arrays must have length 1 or n_replicates, or shape (n_replicates, n_agents)")
        if any(v.ndim == 2 for v in vals):                    # one set per agent, replicate-major
            n = self.B * self.N
            vals = [np.broadcast_to(v[:, None] if v.shape == (self.B,) and self.B != 1 else v,
                                    (self.B, self.N)).reshape(-1) for v in vals]
        else:
            n = max(v.size for v in vals)
            vals = [np.broadcast_to(v.reshape(-1), (n,)) for v in vals]
        tab = np.ascontiguousarray(np.stack(vals, axis=1))
        _lib.check(self._lib.abm_base_set_params(self._h, C.c_void_p(tab.ctypes.data), n), "abm_base_set_params")

    def set_agent_geometry(self, agent_fov=None, vision_range=None):
        """Per-agent FOV (as the fraction of pi the reference's AGENT_FOV is) and vision range, (B, N) or (N,)
        arrays (heterogeneous agents, sims.py:499-517); None keeps the engine-wide value of the constructor.
        Calling it without arguments returns to the engine-wide values."""
        if agent_fov is None and vision_range is None:
            _lib.check(self._lib.abm_base_set_agent_geometry(self._h, None, None, None, 0),
                       "abm_base_set_agent_geometry")
            return
        shape = (self.B, self.N)
        fov = np.broadcast_to(np.asarray(self._agent_fov if agent_fov is None else agent_fov, np.float64), shape)
        vr = np.ascontiguousarray(np.broadcast_to(
            np.asarray(self._vision_range if vision_range is None else vision_range, np.float64), shape))
        f0 = np.ascontiguousarray(-fov * np.pi)                                    # sims.py:506
        f1 = np.ascontiguousarray(fov * np.pi)
        _lib.check(self._lib.abm_base_set_agent_geometry(
            self._h, C.c_void_p(f0.ctypes.data), C.c_void_p(f1.ctypes.data), C.c_void_p(vr.ctypes.data),
            self.B * self.N), "abm_base_set_agent_geometry")

    def set_agent_resolution(self, v_field_res=None):
        """Per-agent field resolution, (B, N) or (N,) integers <= the engine's resolution (heterogeneous agents,
        sims.py:507); None returns to the engine-wide one.  fields() keeps the engine's resolution as its last axis:
        agent i's row holds its own v_field_res[i] bins, the rest is False."""
        if v_field_res is None:
            _lib.check(self._lib.abm_base_set_agent_resolution(self._h, None, 0), "abm_base_set_agent_resolution")
            return
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(v_field_res, np.int32), (self.B, self.N)))
        _lib.check(self._lib.abm_base_set_agent_resolution(self._h, C.c_void_p(r.ctypes.data), self.B * self.N),
                   "abm_base_set_agent_resolution")

    def set_regeneration_params(self, patch_radius=None, min_resc_quality=None, max_resc_quality=None,
                                min_resc_perpatch=None, max_resc_perpatch=None):
        """One set of patch-regeneration parameters per replicate (scalars or length-B arrays; a sweep over the patch
        parameters as one batch); without arguments: back to the constructor's values."""
        vals = (patch_radius, min_resc_quality, max_resc_quality, min_resc_perpatch, max_resc_perpatch)
        if all(v is None for v in vals):
            _lib.check(self._lib.abm_base_set_regeneration_params(self._h, None, 0), "abm_base_set_regeneration_params")
            return
        if any(v is None for v in vals):
            raise ValueError("set_regeneration_params: give all five parameters (or none)")
        tab = np.ascontiguousarray(np.stack([np.broadcast_to(np.asarray(v, np.float64), (self.B,)) for v in vals], axis=1))
        _lib.check(self._lib.abm_base_set_regeneration_params(self._h, C.c_void_p(tab.ctypes.data), self.B),
                   "abm_base_set_regeneration_params")

    def set_agent_radii(self, radius=None):
        """Per-agent radius, (B, N) or (N,) (heterogeneous agents, sims.py:502); None returns to the engine-wide one."""
        if radius is None:
            _lib.check(self._lib.abm_base_set_agent_radii(self._h, None, 0), "abm_base_set_agent_radii")
            return
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, np.float64), (self.B, self.N)))
        _lib.check(self._lib.abm_base_set_agent_radii(self._h, C.c_void_p(r.ctypes.data), self.B * self.N),
                   "abm_base_set_agent_radii")

    # -- state -----------------------------------------------------------------------------
    def _fill(self, struct_cls, fields, arrays, count, keep):
        s = struct_cls()
        for name, kind in fields:
            a = arrays.get(name)
            if a is None:
                continue
            arr = np.ascontiguousarray(np.asarray(a, dtype=_DT[kind]).reshape(-1))
            if arr.size != count:
                raise ValueError(f"{name}: expected {count} elements, got {arr.size}")
            keep.append(arr)
            setattr(s, name, arr.ctypes.data)
        return s

    def set_agents(self, **arrays):
        """x, y, theta required on the first call; the rest default to the constructor state of
        Agent.__init__ (agent.py:68-72, 89-100): zeros, patch_id -1.  ``novelty`` may be a
        (B, N, tau) 0/1 array or packed uint32."""
        nov = arrays.get("novelty")
        if nov is not None and np.asarray(nov).ndim == 3:
            bits = (np.asarray(nov) > 0).astype(np.uint32)
            arrays["novelty"] = (bits << np.arange(bits.shape[-1], dtype=np.uint32)).sum(axis=-1).astype(np.uint32)
        keep = []
        s = self._fill(_lib.BaseAgents, _lib.BASE_AGENT_FIELDS, arrays, self.B * self.N, keep)
        _lib.check(self._lib.abm_base_set_agents(self._h, C.byref(s), 0, C.c_void_p(_current_stream())),
                   "abm_base_set_agents")

    def get_agents(self, names=None) -> dict:
        names = [n for n, _ in _lib.BASE_AGENT_FIELDS] if names is None else list(names)
        kinds = dict(_lib.BASE_AGENT_FIELDS)
        out = {n: np.empty((self.B, self.N), _DT[kinds[n]]) for n in names}
        s = _lib.BaseAgents()
        for n in names:
            setattr(s, n, out[n].ctypes.data)
        _lib.check(self._lib.abm_base_get_agents(self._h, C.byref(s), 0, C.c_void_p(_current_stream())),
                   "abm_base_get_agents")
        return out

    def set_patches(self, **arrays):
        keep = []
        s = self._fill(_lib.BasePatches, _lib.BASE_PATCH_FIELDS, arrays, self.B * self.P, keep)
        _lib.check(self._lib.abm_base_set_patches(self._h, C.byref(s), 0, C.c_void_p(_current_stream())),
                   "abm_base_set_patches")

    def get_patches(self) -> dict:
        kinds = dict(_lib.BASE_PATCH_FIELDS)
        out = {n: np.empty((self.B, self.P), _DT[k]) for n, k in kinds.items()}
        s = _lib.BasePatches()
        for n in out:
            setattr(s, n, out[n].ctypes.data)
        _lib.check(self._lib.abm_base_get_patches(self._h, C.byref(s), 0, C.c_void_p(_current_stream())),
                   "abm_base_get_patches")
        return out

    def step(self, n_steps: int = 1, inject_dtheta=None, phases: int = PHASE_ALL):
        ptr = None
        keep = None
        if inject_dtheta is not None:
            keep = np.ascontiguousarray(np.asarray(inject_dtheta, np.float32).reshape(-1))
            if keep.size != self.B * self.N:
                raise ValueError("inject_dtheta must have n_replicates * n_agents elements")
            ptr = C.c_void_p(keep.ctypes.data)
        _lib.check(self._lib.abm_base_step(self._h, int(n_steps), ptr, 0, int(phases), C.c_void_p(_current_stream())),
                   "abm_base_step")

    def inject_regeneration(self, draws=None):
        """Replace the four random draws of every try of a patch regeneration (sims.py:351-361): ``draws`` of shape
        (B, P, n_tries, 4) = (x, y, units, quality); None returns to the engine's own RNG.  For parity tests."""
        if draws is None:
            _lib.check(self._lib.abm_base_inject_regeneration(self._h, None, 0), "abm_base_inject_regeneration")
            return
        d = np.ascontiguousarray(np.asarray(draws, np.float64))
        if d.ndim != 4 or d.shape[:2] != (self.B, self.P) or d.shape[3] != 4:
            raise ValueError("draws must have shape (n_replicates, n_patches, n_tries, 4)")
        _lib.check(self._lib.abm_base_inject_regeneration(self._h, C.c_void_p(d.ctypes.data), int(d.shape[2])),
                   "abm_base_inject_regeneration")

    def fields(self) -> np.ndarray:
        """(B, N, R) bool, STORED (flipped + FOV-masked) order like Agent.soc_v_field."""
        w = np.empty((self.B, self.N, self.W), np.uint32)
        _lib.check(self._lib.abm_base_get_fields(self._h, C.c_void_p(w.ctypes.data), 0, C.c_void_p(_current_stream())),
                   "abm_base_get_fields")
        bits = (w[..., :, None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)
        return bits.reshape(self.B, self.N, -1)[..., :self.R].astype(bool)

    def metrics(self, reset: bool = False) -> dict:
        """Per-replicate summary metrics reduced on the device (abm/loader/data_loader.py
        calculate_search_efficiency :1294-1353, calculate_relocation_time :1903-1928): dict of (B,) float32 arrays
        `search_efficiency` (mean collected_r / T), `relocation_time`, `explore_time`, `exploit_time`,
        `collide_time` (fractions of the agent-steps logged in that mode) and `mean_collected`; T counts the steps
        since creation or the last `reset`."""
        out = np.empty((self.B, 6), np.float32)
        _lib.check(self._lib.abm_base_metrics(self._h, C.c_void_p(out.ctypes.data), 0, int(reset),
                                              C.c_void_p(_current_stream())), "abm_base_metrics")
        names = ("search_efficiency", "relocation_time", "explore_time", "exploit_time", "collide_time", "mean_collected")
        return {k: out[:, i].copy() for i, k in enumerate(names)}

    def counters(self) -> dict:
        c = (C.c_uint64 * 4)()
        _lib.check(self._lib.abm_base_get_counters(self._h, c, C.c_void_p(_current_stream())), "abm_base_get_counters")
        return dict(patches_regenerated=int(c[0]), regeneration_failed=int(c[1]), launches=int(c[2]), steps=int(c[3]))
