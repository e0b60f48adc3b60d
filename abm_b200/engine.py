"""Host-side handle of the fused visual-flocking engine (thin wrapper over the C ABI).

Arrays may be numpy arrays (host; copied inside the call) or torch CUDA tensors
(device pointers, no copy).  Shapes are (n_replicates, n_agents) or flat.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _fov_pixels(R: int, fov) -> tuple[int, int]:
    """find_nearest(linspace(-pi, pi, R), fov[i]) in float64 (vf_supcalc.py:39, :105)."""
    phis = np.linspace(-np.pi, np.pi, R)
    return int(np.abs(phis - fov[0]).argmin()), int(np.abs(phis - fov[1]).argmin())


def _is_torch_cuda(a) -> bool:
    return hasattr(a, "is_cuda") and a.is_cuda


def _current_stream() -> int:
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_stream().cuda_stream)
    except ImportError:
        pass
    return 0


class VFEngine:
    """B replicates x N agents of the visual-flocking model, resident on one GPU.

    Mirrors what VFSimulation + VFAgent hold per run (vf_sims.py:16-44, vf_agent.py:12-50):
    ``resolution`` is R AFTER the int(R / fov) rescale, ``fov`` the (lo, hi) tuple in radians.
    """

    def __init__(self, n_replicates: int, n_agents: int, *, resolution: int = 1200,
                 fov=(-np.pi, np.pi), boundary: str = "walls", width: float = 900.0, height: float = 900.0,
                 window_pad: float = 30.0, limit_movement: bool = False, max_vel: float = 3.0,
                 max_th: float = 0.1, exact_fixup: bool = True, keep_fields: bool = False,
                 keep_terms: bool = False, device: int = 0, tile: tuple[int, int] | None = None,
                 tile_cycle: tuple[int, int] | None = None, spatial_sort: bool = True, resort_every: int = 32):
        """``tile`` = (begin, count): this engine updates that contiguous range of (internal) agent slots of one large
        swarm; ``tile_cycle`` = (G, r): it updates every G-th block of 128 slots starting with block r (balanced work
        across G GPUs whatever the density distribution); both read all N neighbour records."""
        if boundary not in ("walls", "infinite"):
            raise ValueError(f"boundary must be 'walls' or 'infinite', got {boundary!r}")
        self._lib = _lib.load()
        self.B, self.N, self.R = int(n_replicates), int(n_agents), int(resolution)
        self.W = (self.R + 31) // 32
        self.device = int(device)
        self.tile_begin, self.tile_count = (0, self.N) if tile is None else (int(tile[0]), int(tile[1]))
        cyc, phase = (0, 0) if tile_cycle is None else (int(tile_cycle[0]), int(tile_cycle[1]))
        if cyc > 1:
            if tile is not None:
                raise ValueError("give either tile or tile_cycle")
            self.tile_begin, self.tile_count = 0, self.N // cyc
        self.tile_cycle, self.tile_phase = cyc, phase
        f0, f1 = _fov_pixels(self.R, fov)
        flags = (_lib.VF_EXACT_FIXUP if exact_fixup else 0) | (_lib.VF_KEEP_FIELDS if keep_fields else 0) \
            | (_lib.VF_KEEP_TERMS if keep_terms else 0) | (_lib.VF_SPATIAL_SORT if spatial_sort else 0)
        cfg = _lib.VFConfig(
            struct_size=C.sizeof(_lib.VFConfig), n_replicates=self.B, n_agents=self.N, resolution=self.R,
            fov_px0=f0, fov_px1=f1,
            boundary=_lib.BOUNDARY_INFINITE if boundary == "infinite" else _lib.BOUNDARY_WALLS,
            limit_movement=int(bool(limit_movement)), width=float(width), height=float(height),
            window_pad=float(window_pad), max_vel=float(max_vel), max_th=float(max_th), flags=flags,
            tile_begin=self.tile_begin, tile_count=0 if (tile is None and cyc <= 1) else self.tile_count,
            resort_every=int(resort_every), tile_cycle=cyc, tile_phase=phase)
        self._h = C.c_void_p()
        _lib.check(self._lib.abm_vf_create(C.byref(cfg), self.device, C.byref(self._h)), "abm_vf_create")
        self.keep_fields, self.keep_terms = keep_fields, keep_terms

    # -- life cycle --------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.abm_destroy(self._h)
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

    # -- helpers -----------------------------------------------------------------------
    def _ptr(self, a, dtype, count, keep):
        """(pointer, on_device) of an input array; host arrays are made contiguous/typed."""
        if a is None:
            return None, None
        if _is_torch_cuda(a):
            import torch
            want = {np.float32: torch.float32, np.float64: torch.float64, np.uint32: torch.int32}[dtype]
            if a.dtype != want or not a.is_contiguous() or a.numel() != count:
                raise ValueError("device tensors must be contiguous, of the engine dtype and full size")
            return C.c_void_p(a.data_ptr()), 1
        arr = np.ascontiguousarray(np.asarray(a, dtype=dtype).reshape(-1))
        if arr.size != count:
            raise ValueError(f"expected {count} elements, got {arr.size}")
        keep.append(arr)
        return C.c_void_p(arr.ctypes.data), 0

    @staticmethod
    def _same_side(flags):
        s = {f for f in flags if f is not None}
        if len(s) > 1:
            raise ValueError("all arrays of one call must live on the same side (host or device)")
        return s.pop() if s else 0

    # -- parameters --------------------------------------------------------------------
    def set_params(self, GAM=0.1, V0=1.0, ALP0=1.0, ALP1=0.09, BET0=1.0, BET1=0.09):
        """Scalars (shared) or length-B arrays (one set per replicate: a MetaProtocol sweep)."""
        vals = [np.atleast_1d(np.asarray(v, np.float64)) for v in (GAM, V0, ALP0, ALP1, BET0, BET1)]
        n = max(v.size for v in vals)
        if n not in (1, self.B):
            raise ValueError("parameter arrays must have length 1 or n_replicates")
        tab = np.ascontiguousarray(np.stack([np.broadcast_to(v, (n,)) for v in vals], axis=1))
        _lib.check(self._lib.abm_vf_set_params(self._h, C.c_void_p(tab.ctypes.data), n), "abm_vf_set_params")

    def set_agent_overrides(self, alp0=None, bet0=None, v0=None):
        """Per-agent VFAgent.ALP0 / .BET0 / .V0 (NaN = None = use the replicate's value)."""
        keep = []
        total = self.B * self.N
        ptrs = [self._ptr(a, np.float32, total, keep) for a in (alp0, bet0, v0)]
        side = self._same_side([p[1] for p in ptrs])
        _lib.check(self._lib.abm_vf_set_agent_overrides(self._h, ptrs[0][0], ptrs[1][0], ptrs[2][0], side,
                                                        C.c_void_p(_current_stream())), "abm_vf_set_agent_overrides")
        if not side:
            self.synchronize()

    # -- state -------------------------------------------------------------------------
    def set_state(self, x, y, theta, vel, radius=None, nonblocking=False):
        """``radius``: scalar or (B, N) array; None keeps the radii of the previous call.
        ``nonblocking``: the host arrays are PINNED, C-contiguous float32 and stay alive and untouched until the
        caller has synchronised the stream -- the call only enqueues (ABM_HOST_PINNED_ASYNC)."""
        keep = []
        total = self.B * self.N
        arrays = [x, y, theta, vel]
        if radius is not None:
            if not _is_torch_cuda(radius):
                radius = np.broadcast_to(np.asarray(radius, np.float32), np.asarray(x).shape)
            arrays.append(radius)
        ptrs = [self._ptr(a, np.float32, total, keep) for a in arrays]
        side = self._same_side([p[1] for p in ptrs])
        if nonblocking and not side:
            if any(not isinstance(a, np.ndarray) or k.ctypes.data != a.ctypes.data for k, a in zip(keep, arrays)):
                raise ValueError("nonblocking=True needs C-contiguous float32 arrays of full size (no copies are made)")
            side = 2
        args = [p[0] for p in ptrs] + ([None] if radius is None else [])
        _lib.check(self._lib.abm_set_state(self._h, *args, side, C.c_void_p(_current_stream())), "abm_set_state")
        if not side:
            self.synchronize()   # the host arrays in `keep` may go away after return

    def get_state(self, out=None, nonblocking=False):
        """Returns dict(x, y, theta, vel) of (B, N) float32 numpy arrays (or fills ``out``,
        a dict of numpy arrays / torch CUDA tensors).  ``nonblocking`` (``out`` = pinned host arrays): the copies are
        only enqueued; synchronise the stream before reading them."""
        total = self.B * self.N
        if out is None:
            out = {k: np.empty((self.B, self.N), np.float32) for k in ("x", "y", "theta", "vel")}
        ptrs, sides = [], []
        for k in ("x", "y", "theta", "vel"):
            a = out.get(k)
            if a is None:
                ptrs.append(None)
            elif _is_torch_cuda(a):
                ptrs.append(C.c_void_p(a.data_ptr())); sides.append(1)
            else:
                if a.dtype != np.float32 or not a.flags.c_contiguous or a.size != total:
                    raise ValueError("output arrays must be C-contiguous float32 of full size")
                ptrs.append(C.c_void_p(a.ctypes.data)); sides.append(0)
        side = self._same_side(sides)
        if nonblocking and not side:
            side = 2
        _lib.check(self._lib.abm_get_state(self._h, *ptrs, side, C.c_void_p(_current_stream())), "abm_get_state")
        return out

    def set_state_packed(self, xytv, radius=None, nonblocking=False):
        """The state as ONE (B, N, 4) float32 array of (x, y, theta, vel) rows (numpy / torch CUDA tensor): one copy
        instead of four.  ``radius`` as in `set_state` (scalar or (B, N); None keeps the radii).  ``nonblocking``:
        `xytv` is PINNED host memory that stays alive and untouched until the caller has synchronised the stream."""
        keep = []
        total = self.B * self.N
        p4, side4 = self._ptr(xytv, np.float32, 4 * total, keep)
        pr, side_r = None, None
        if radius is not None:
            if not _is_torch_cuda(radius):
                radius = np.broadcast_to(np.asarray(radius, np.float32), (self.B, self.N))
            pr, side_r = self._ptr(radius, np.float32, total, keep)
        side = self._same_side([side4, side_r])
        if nonblocking and not side:
            if not isinstance(xytv, np.ndarray) or keep[0].ctypes.data != xytv.ctypes.data:
                raise ValueError("nonblocking=True needs a C-contiguous float32 array of full size (no copies are made)")
            side = 2
        _lib.check(self._lib.abm_set_state_packed(self._h, p4, pr, side, C.c_void_p(_current_stream())),
                   "abm_set_state_packed")
        if side == 2 and radius is not None:
            self.synchronize()   # the broadcast radius array in `keep` may go away after return

    def get_state_packed(self, out=None, nonblocking=False):
        """(B, N, 4) float32 rows (x, y, theta, vel); fills ``out`` (numpy array or torch CUDA tensor) if given."""
        total = self.B * self.N
        if out is None:
            out = np.empty((self.B, self.N, 4), np.float32)
        if _is_torch_cuda(out):
            if not out.is_contiguous() or out.numel() != 4 * total:
                raise ValueError("output tensor must be contiguous float32 of full size")
            p, side = C.c_void_p(out.data_ptr()), 1
        else:
            if out.dtype != np.float32 or not out.flags.c_contiguous or out.size != 4 * total:
                raise ValueError("output array must be C-contiguous float32 of full size")
            p, side = C.c_void_p(out.ctypes.data), (2 if nonblocking else 0)
        _lib.check(self._lib.abm_get_state_packed(self._h, p, side, C.c_void_p(_current_stream())),
                   "abm_get_state_packed")
        return out

    def step(self, n_steps: int = 1):
        _lib.check(self._lib.abm_vf_step(self._h, int(n_steps), C.c_void_p(_current_stream())), "abm_vf_step")

    def set_line_map(self, line_map=None, sensor_radius=9, sensor_distance=20):
        """Lines to follow (vf_agent.py:273-276, vf_supcalc.follow_lines_local): ``line_map`` is the reference's
        ``VFAgent.line_map``, shape (WIDTH + window_pad, HEIGHT + window_pad), values in [0, 1]; None switches line
        following off.  Sensor defaults as VFAgent.sensor_size / sensor_distance (vf_agent.py:30-31)."""
        if line_map is None:
            _lib.check(self._lib.abm_vf_set_line_map(self._h, None, 0, 0, 0.0, 0.0, 0, None), "abm_vf_set_line_map")
            return
        m = np.ascontiguousarray(line_map, np.float32)
        if m.ndim != 2:
            raise ValueError("line_map must be two-dimensional")
        _lib.check(self._lib.abm_vf_set_line_map(self._h, C.c_void_p(m.ctypes.data), m.shape[0], m.shape[1],
                                                 float(sensor_radius), float(sensor_distance), 0,
                                                 C.c_void_p(_current_stream())), "abm_vf_set_line_map")

    def step_host(self, xytv_in, xytv_out, n_steps: int = 1):
        """Upload the (B, N, 4) float32 state ``xytv_in``, run ``n_steps`` steps, download the new state into ``xytv_out``
        -- one call (abm_vf_step_host) that does not block: both arrays must be PINNED host memory (e.g. the numpy view
        of a ``torch.empty(..., pin_memory=True)``) and stay alive until ``synchronize()``.  Large batches go through in
        replicate chunks whose copies overlap the other chunks' steps.  The radii are those of an earlier
        ``set_state`` / ``set_state_packed``."""
        total = 4 * self.B * self.N
        for a in (xytv_in, xytv_out):
            if a.dtype != np.float32 or not a.flags.c_contiguous or a.size != total:
                raise ValueError("step_host: arrays must be C-contiguous float32 of shape (B, N, 4)")
        _lib.check(self._lib.abm_vf_step_host(self._h, C.c_void_p(xytv_in.ctypes.data), C.c_void_p(xytv_out.ctypes.data),
                                              int(n_steps), C.c_void_p(_current_stream())), "abm_vf_step_host")

    def synchronize(self):
        _lib.check(self._lib.abm_synchronize(self._h, C.c_void_p(_current_stream())), "abm_synchronize")

    # -- outputs of the last step ----------------------------------------------------------
    def tile_slots(self) -> np.ndarray:
        """Internal slot of every focal agent li of this engine's tile (contiguous range or cyclic blocks)."""
        li = np.arange(self.tile_count)
        if self.tile_cycle > 1:
            blk = _lib.VF_TILE_BLOCK
            return ((li // blk) * self.tile_cycle + self.tile_phase) * blk + li % blk
        return self.tile_begin + li

    def fields_packed(self) -> np.ndarray:
        """(B, tile, W) uint32, STORED (flipped) order like Agent.soc_v_field.  Rows are in the caller's agent order;
        on a TILED engine with the spatial sort on they are in internal slot order (row li = agent
        ``permutation()[b, tile_slots()[li]]``), like `terms()`."""
        out = np.empty((self.B, self.tile_count, self.W), np.uint32)
        _lib.check(self._lib.abm_get_fields(self._h, C.c_void_p(out.ctypes.data), 0, C.c_void_p(_current_stream())),
                   "abm_get_fields")
        return out

    def fields(self) -> np.ndarray:
        """(B, tile, R) bool, STORED order."""
        w = self.fields_packed()
        bits = (w[..., :, None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)
        return bits.reshape(self.B, self.tile_count, -1)[..., :self.R].astype(bool)

    def terms(self) -> np.ndarray:
        """(B, tile, 6) float64: dvel, dpsi, a_blob, a_edge, b_blob, b_edge."""
        out = np.empty((self.B, self.tile_count, 6), np.float64)
        _lib.check(self._lib.abm_vf_get_terms(self._h, C.c_void_p(out.ctypes.data), 0,
                                              C.c_void_p(_current_stream())), "abm_vf_get_terms")
        return out

    def counters(self) -> dict:
        c = (C.c_uint64 * 4)()
        _lib.check(self._lib.abm_get_counters(self._h, c, C.c_void_p(_current_stream())), "abm_get_counters")
        return dict(fp64_pairs=int(c[0]), fp64_inline=int(c[1]), fp32_fp64_differ=int(c[2]), launches=int(c[3]))

    def permutation(self) -> np.ndarray:
        """(B, N) int32: perm[b, slot] = caller's index of the agent in internal slot `slot`."""
        out = np.empty((self.B, self.N), np.int32)
        _lib.check(self._lib.abm_vf_get_permutation(self._h, C.c_void_p(out.ctypes.data), 0,
                                                    C.c_void_p(_current_stream())), "abm_vf_get_permutation")
        return out

    def resort(self):
        _lib.check(self._lib.abm_vf_resort(self._h, C.c_void_p(_current_stream())), "abm_vf_resort")

    def internal_array_ptrs(self) -> tuple[int, int]:
        """Device pointers of the internal heading / speed arrays (B*N float32, internal order)."""
        t, v = C.c_void_p(), C.c_void_p()
        _lib.check(self._lib.abm_vf_internal_arrays(self._h, C.byref(t), C.byref(v)), "abm_vf_internal_arrays")
        return int(t.value), int(v.value)

    def metrics(self) -> dict:
        """Per-replicate summary metrics of the current state, computed on the device (the
        quantities of abm/loader/data_loader.py): dict of (B,) float32 arrays `polarization`, `mean_iid`,
        `mean_nn_dist`, `collision` (1.0 where some pair is closer than 2 * radius), `colliding_agents` (fraction of
        the agents with a higher-indexed agent closer than 2 * radius: the per-agent indicator whose time average is
        the loader's "aacoll")."""
        out = np.empty((self.B, 5), np.float32)
        _lib.check(self._lib.abm_vf_metrics(self._h, C.c_void_p(out.ctypes.data), 0, C.c_void_p(_current_stream())),
                   "abm_vf_metrics")
        return dict(polarization=out[:, 0].copy(), mean_iid=out[:, 1].copy(), mean_nn_dist=out[:, 2].copy(),
                    collision=out[:, 3].copy(), colliding_agents=out[:, 4].copy())

    def slow_entries(self) -> tuple[int, int]:
        """(pairs off the symmetric kernel's fast path so far, number of its launches)."""
        n, l = C.c_uint64(), C.c_uint64()
        _lib.check(self._lib.abm_vf_slow_entries(self._h, C.byref(n), C.byref(l), C.c_void_p(_current_stream())),
                   "abm_vf_slow_entries")
        return int(n.value), int(l.value)

    def ipc_export(self) -> bytes:
        """Handles of this (tiled) engine's record tables and flags for the fused tile exchange."""
        buf = C.create_string_buffer(_lib.VF_IPC_BYTES)
        _lib.check(self._lib.abm_vf_ipc_export(self._h, buf), "abm_vf_ipc_export")
        return buf.raw

    def ipc_attach(self, my_rank: int, exports: list[bytes]):
        """Map the peers' tables (exports of ALL ranks, in rank order): from now on step() exchanges the tiles
        itself over NVLink peer memory and every rank must step in lock step."""
        blob = b"".join(exports)
        _lib.check(self._lib.abm_vf_ipc_attach(self._h, len(exports), int(my_rank), blob), "abm_vf_ipc_attach")

    def kernel_stats(self) -> dict:
        """Step-kernel launches so far, by kernel (and, for the symmetric one, by fast-path width)."""
        c = (C.c_uint64 * 4)()
        _lib.check(self._lib.abm_vf_kernel_stats(self._h, c), "abm_vf_kernel_stats")
        return dict(symmetric=int(c[0]), symmetric_wide=int(c[1]), onesided=int(c[2]), warp=int(c[3]))

    def cluster_launches(self) -> int:
        """Multi-step launches that ran as ONE thread-block cluster (hardware barrier between the steps)."""
        v = C.c_uint64(0)
        _lib.check(self._lib.abm_vf_cluster_launches(self._h, C.byref(v)), "abm_vf_cluster_launches")
        return int(v.value)

    def last_kernel(self) -> str:
        """Name of the step kernel the last step() launched."""
        return (self._lib.abm_vf_last_kernel(self._h) or b"").decode()

    def record_table_ptr(self) -> tuple[int, int]:
        p = C.c_void_p()
        nbytes = C.c_int()
        _lib.check(self._lib.abm_vf_record_table(self._h, C.byref(p), C.byref(nbytes)), "abm_vf_record_table")
        return int(p.value), int(nbytes.value)
