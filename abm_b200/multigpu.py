"""Multi-GPU plumbing (one process per GPU, torch.distributed):

  * replicate batches shard with NO data-path communication: `replicate_shard`;
  * a single large swarm shards by agent tiles: every rank updates its tile against the full
    neighbour-record table and the tiles are exchanged with ONE all-gather of the 16-byte
    records (x, y, radius, cull^2) per step (`TiledSwarm`); headings and speeds are private
    to the owner of an agent and are never exchanged.
"""
from __future__ import annotations

import numpy as np


def replicate_shard(n_replicates: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [begin, begin + count) of replicates owned by `rank`."""
    base, rem = divmod(n_replicates, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def agent_tile(n_agents: int, world: int, rank: int) -> tuple[int, int]:
    """Equal agent tiles (the in-place all-gather needs equal counts)."""
    if n_agents % world:
        raise ValueError(f"n_agents ({n_agents}) must be divisible by the number of ranks ({world})")
    count = n_agents // world
    return rank * count, count


def gather_tiles(local: "torch.Tensor", world: int, group=None) -> "torch.Tensor":
    """All-gather of equal tiles along dim 0 (works on gloo/CPU and nccl/CUDA)."""
    import torch
    import torch.distributed as dist
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=group)
    return torch.cat(parts, dim=0)


class _DeviceMemory:
    """Zero-copy view of engine-owned device memory for torch (CUDA array interface v2)."""

    def __init__(self, ptr: int, n_float32: int):
        self.__cuda_array_interface__ = {"shape": (n_float32,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class TiledSwarm:
    """One very large swarm (B = 1) across the ranks of a torch.distributed NCCL group."""

    def __init__(self, n_agents: int, *, group=None, **engine_kwargs):
        import torch
        import torch.distributed as dist
        from .engine import VFEngine
        self.torch, self.dist, self.group = torch, dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.N = int(n_agents)
        self.begin, self.count = agent_tile(self.N, self.world, self.rank)
        self.engine = VFEngine(1, self.N, tile=(self.begin, self.count), device=torch.cuda.current_device(),
                               **engine_kwargs)
        self._tables = {}

    def set_params(self, **kw):
        self.engine.set_params(**kw)

    def set_state(self, x, y, theta, vel, radius):
        """Every rank passes the FULL state (identical on all ranks)."""
        self.engine.set_state(x, y, theta, vel, radius)

    def _table(self):
        ptr, nbytes = self.engine.record_table_ptr()
        t = self._tables.get(ptr)
        if t is None:
            t = self.torch.as_tensor(_DeviceMemory(ptr, self.N * nbytes // 4), device="cuda").view(self.N, nbytes // 4)
            self._tables[ptr] = t
        return t

    def step(self, n_steps: int = 1):
        for _ in range(n_steps):
            self.engine.step(1)                       # writes this rank's tile of the next table
            table = self._table()
            tile = table[self.begin:self.begin + self.count]
            self.dist.all_gather_into_tensor(table, tile, group=self.group)   # in place, 16 B / agent

    def _internal(self):
        """torch views of the engine's internal heading / speed arrays (internal order, full size)."""
        tp, vp = self.engine.internal_array_ptrs()
        mk = lambda p: self.torch.as_tensor(_DeviceMemory(p, self.N), device="cuda")
        return mk(tp), mk(vp)

    def resync(self):
        """Make headings / speeds current on every rank (one all-gather each) and re-sort the
        internal order spatially -- call every few hundred steps of a long run."""
        for t in self._internal():
            self.dist.all_gather_into_tensor(t, t[self.begin:self.begin + self.count].clone(), group=self.group)
        self.engine.resort()
        self._tables = {}

    def get_state(self):
        """Full state in the caller's agent order: (x, y) from the gathered record table, headings /
        speeds gathered here (off the step path) from the tiles and un-permuted."""
        st = self.engine.get_state()
        out = {"x": st["x"][0], "y": st["y"][0]}
        perm = self.engine.permutation()[0]
        for k, t in zip(("theta", "vel"), self._internal()):
            full = gather_tiles(t[self.begin:self.begin + self.count].clone(), self.world, self.group).cpu().numpy()
            api = np.empty_like(full)
            api[perm] = full
            out[k] = api
        return out
