"""Trajectory recorder in the reference's on-disk layout.

The reference keeps per-agent Python lists in RAM (`ifdb.save_agent_data_RAM`, ifdb.py:55-96: posx / posy
int-truncated, orientation, velocity, mode, ...) and writes them at the end of a run as zarr arrays
`ag_posx.zarr`, `ag_posy.zarr`, `ag_ori.zarr`, `ag_vel.zarr`, `ag_mode.zarr` of shape (num_agents, T),
dtype float64, next to `env_params.json` (ifdb.py:470-508, env_saver.py:15-28), which is what
`ExperimentLoader` / `ExperimentReplay` read back.

Here the state of every recorded step is appended on the DEVICE to a ring of `chunk` steps (the engine's unpack
kernel writes straight into the ring, no host round trip per step); a full chunk goes to pinned host memory with one
asynchronous copy on a side stream and is written by a background thread while the simulation runs on.  The arrays
are zarr format 2 directories written without the zarr package (it is not in the image): uncompressed C-order
float64 chunks of shape (num_agents, chunk) named "0.<k>", and a `.zarray` metadata file -- readable by
`zarr.open(path, mode="r")` of any zarr version that reads format 2.  One folder per recorded replicate.
"""
from __future__ import annotations

import json
import os
import queue
import threading

import numpy as np

_FIELDS = ("posx", "posy", "ori", "vel")


def _write_zarray(path, shape, chunks):
    meta = {"chunks": list(chunks), "compressor": None, "dtype": "<f8", "fill_value": 0.0, "filters": None,
            "order": "C", "shape": list(shape), "zarr_format": 2}
    with open(os.path.join(path, ".zarray"), "w") as f:
        json.dump(meta, f, indent=4)


def read_zarr_v2(path) -> np.ndarray:
    """Minimal reader of the arrays this module writes (uncompressed zarr format 2); for tests and for
    users without the zarr package."""
    with open(os.path.join(path, ".zarray")) as f:
        meta = json.load(f)
    assert meta["zarr_format"] == 2 and meta["compressor"] is None and meta["order"] == "C"
    shape, chunks = meta["shape"], meta["chunks"]
    out = np.full(shape, meta["fill_value"], np.dtype(meta["dtype"]))
    n_t = -(-shape[1] // chunks[1])
    for k in range(n_t):
        fn = os.path.join(path, f"0.{k}")
        if not os.path.exists(fn):
            continue
        c = np.fromfile(fn, np.dtype(meta["dtype"])).reshape(chunks)
        t0 = k * chunks[1]
        w = min(chunks[1], shape[1] - t0)
        out[:, t0:t0 + w] = c[:shape[0], :w]
    return out


def agent_row(agent_id: int, n_agents: int) -> int:
    """Row of the `ag_*.zarr` arrays that holds agent ``agent_id``.  The reference writes agent ``ag_id`` into row
    ``ag_id - 1`` (ifdb.py:504-508) while its agents are numbered from 0 (sims.py:528, vf_sims.py:212): agent 0 ends up
    in the LAST row (index -1), agent i in row i - 1.  The recorders here leave exactly that layout on disk (patches are
    numbered from 1, sims.py:362, so patch p is row p - 1 = its slot)."""
    return (agent_id - 1) % n_agents


class VFRecorder:
    """Records a VFEngine's trajectory.  ``replicates``: which replicates to keep (default: all);
    ``every``: record every k-th call of `record()`; ``chunk``: steps per device ring / per zarr chunk."""

    def __init__(self, engine, save_dir, replicates=None, every: int = 1, chunk: int = 64, env_params=None, dirs=None):
        """``dirs``: one output folder per recorded replicate (default: ``save_dir/replicate_<b>``); ``env_params``: one
        dict for all, or a list with one dict per recorded replicate (written as env_params.json, env_saver.py:15-28)."""
        import torch
        self.torch = torch
        self.eng = engine
        self.B, self.N = engine.B, engine.N
        self.reps = list(range(self.B)) if replicates is None else [int(b) for b in replicates]
        self.every, self.chunk = int(every), int(chunk)
        self.save_dir = save_dir
        self.dev = torch.device("cuda", engine.device)
        self.ring = [torch.empty((self.chunk, 4, self.B, self.N), dtype=torch.float32, device=self.dev) for _ in range(2)]
        self.host = [torch.empty((self.chunk, 4, len(self.reps), self.N), dtype=torch.float32).pin_memory()
                     for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.done = [None, None]           # event of the last D2H copy out of ring / into host buffer k
        self.written = [threading.Event(), threading.Event()]
        for w in self.written:
            w.set()
        self.cur, self.fill, self.calls, self.n_recorded, self.n_chunks = 0, 0, 0, 0, 0
        self._sel = torch.tensor(self.reps, device=self.dev, dtype=torch.long)
        self._q: queue.Queue = queue.Queue()
        self._err = None
        self._dirs = []
        for k, b in enumerate(self.reps):
            d = os.path.join(save_dir, f"replicate_{b:05d}") if dirs is None else dirs[k]
            for name in _FIELDS + ("mode",):
                os.makedirs(os.path.join(d, f"ag_{name}.zarr"), exist_ok=True)
            if env_params is not None:
                with open(os.path.join(d, "env_params.json"), "w") as f:    # env_saver.py:15-28
                    json.dump(env_params[k] if isinstance(env_params, (list, tuple)) else env_params, f, indent=4)
            self._dirs.append(d)
        self._thread = threading.Thread(target=self._writer, daemon=True)
        self._thread.start()

    # -- device side -----------------------------------------------------------------------------------------
    def record(self):
        """Append the engine's current state (call after every step)."""
        self.calls += 1
        if (self.calls - 1) % self.every:
            return
        if self.fill == 0:
            # the ring about to be refilled must have been copied out, its host twin written to disk
            if self.done[self.cur] is not None:
                self.torch.cuda.current_stream(self.dev).wait_event(self.done[self.cur])
        slot = self.ring[self.cur][self.fill]
        self.eng.get_state({"x": slot[0], "y": slot[1], "theta": slot[2], "vel": slot[3]})   # device -> device
        self.fill += 1
        self.n_recorded += 1
        if self.fill == self.chunk:
            self._flush()

    def _flush(self):
        if self.fill == 0:
            return
        torch = self.torch
        k, n = self.cur, self.fill
        self.written[k].wait()                      # host buffer k is free again
        self.written[k].clear()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            src = self.ring[k][:n]
            if len(self.reps) != self.B:
                src = src.index_select(2, self._sel)
            self.host[k][:n].copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.done[k] = ev
        self._q.put((k, n, self.n_chunks, ev))
        self.n_chunks += 1
        self.cur ^= 1
        self.fill = 0

    # -- host side -------------------------------------------------------------------------------------------
    def _writer(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            k, n, chunk_idx, ev = item
            try:
                ev.synchronize()
                a = self.host[k][:n].numpy()                                 # (n, 4, R, N) float32
                for ri, d in enumerate(self._dirs):
                    blk = a[:, :, ri, :].astype(np.float64)                  # (n, 4, N)
                    blk[:, 0] = np.trunc(blk[:, 0]); blk[:, 1] = np.trunc(blk[:, 1])   # int(agent.position[.]), ifdb.py:83-84
                    for fi, name in enumerate(_FIELDS):
                        out = np.zeros((self.N, self.chunk), np.float64)
                        out[:, :n] = np.roll(blk[:, fi].T, -1, axis=0)        # row = agent id - 1 (see agent_row)
                        out.tofile(os.path.join(d, f"ag_{name}.zarr", f"0.{chunk_idx}"))
                    np.zeros((self.N, self.chunk), np.float64).tofile(os.path.join(d, "ag_mode.zarr", f"0.{chunk_idx}"))
            except Exception as exc:                                          # surfaced by close()
                self._err = exc
            finally:
                self.written[k].set()

    def close(self):
        """Flush the partial chunk, wait for the writer, write the array metadata.  Returns the folders."""
        self._flush()
        self._q.put(None)
        self._thread.join()
        if self._err is not None:
            raise self._err
        for d in self._dirs:
            for name in _FIELDS + ("mode",):
                _write_zarray(os.path.join(d, f"ag_{name}.zarr"), (self.N, self.n_recorded), (self.N, self.chunk))
        return list(self._dirs)


class BaseRecorder:
    """The same for the BASE / foraging engine (`ifdb.py:83-96, 235-253, 437-508`): per recorded replicate
    `ag_posx/posy/ori/vel/mode/w/u/ipriv/collr/explr.zarr` of shape (num_agents, T) and `res_posx/posy/rad/left/
    qual.zarr` of shape (num_patches, T).  A foraging replicate is small (50 agents, 3 patches), so the state of the
    selected replicates is simply fetched after every recorded step (one bulk download of the batch) and kept in host
    memory like the reference does, then written once by `close()`.  Patch rows are ordered by patch slot (the reference
    orders them by resource id, which a regenerated patch keeps)."""

    _AG = (("posx", "x"), ("posy", "y"), ("ori", "theta"), ("vel", "vel"), ("mode", "mode"), ("w", "w"), ("u", "u"),
           ("ipriv", "i_priv"), ("collr", "collected"), ("explr", "patch_id"))
    _RES = (("posx", "x"), ("posy", "y"), ("rad", "radius"), ("left", "left"), ("qual", "quality"))

    def __init__(self, engine, save_dir, replicates=None, every: int = 1, env_params=None, dirs=None):
        self.eng, self.save_dir, self.every = engine, save_dir, int(every)
        self.reps = list(range(engine.B)) if replicates is None else [int(b) for b in replicates]
        self.env_params, self.dirs = env_params, dirs
        self.calls = 0
        self._ag = {n: [] for n, _ in self._AG}
        self._res = {n: [] for n, _ in self._RES}

    def record(self):
        self.calls += 1
        if (self.calls - 1) % self.every:
            return
        a = self.eng.get_agents([k for _, k in self._AG])
        for n, k in self._AG:
            v = a[k][self.reps].astype(np.float64)
            self._ag[n].append(np.trunc(v) if n in ("posx", "posy") else v)       # int(agent.position[.]), ifdb.py:83-84
        if self.eng.P:
            p = self.eng.get_patches()
            for n, k in self._RES:
                self._res[n].append(p[k][self.reps].astype(np.float64))

    def close(self):
        dirs = []
        T = len(self._ag["posx"])
        for ri, b in enumerate(self.reps):
            d = os.path.join(self.save_dir, f"replicate_{b:05d}") if self.dirs is None else self.dirs[ri]
            os.makedirs(d, exist_ok=True)
            if self.env_params is not None:
                ep = self.env_params[ri] if isinstance(self.env_params, (list, tuple)) else self.env_params
                with open(os.path.join(d, "env_params.json"), "w") as f:
                    json.dump(ep, f, indent=4)
            for prefix, store in (("ag", self._ag), ("res", self._res)):
                for n, steps in store.items():
                    if not steps:
                        continue
                    arr = np.stack([s[ri] for s in steps], axis=1)                          # (num, T)
                    if prefix == "ag":
                        arr = np.roll(arr, -1, axis=0)                                      # row = agent id - 1 (agent_row)
                    arr = np.ascontiguousarray(arr)
                    path = os.path.join(d, f"{prefix}_{n}.zarr")
                    os.makedirs(path, exist_ok=True)
                    arr.tofile(os.path.join(path, "0.0"))
                    _write_zarray(path, arr.shape, arr.shape)      # one chunk, like the reference (ifdb.py:441-443)
            dirs.append(d)
        assert all(len(v) in (0, T) for v in list(self._ag.values()) + list(self._res.values()))
        return dirs
