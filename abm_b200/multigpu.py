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


def cyclic_slots(n_agents: int, world: int, rank: int, block: int = 128) -> np.ndarray:
    """Slots owned by `rank` when the blocks of `block` consecutive slots are dealt to the ranks round robin (the
    engine's tile_cycle): balanced work whatever the density distribution along the spatially sorted order."""
    if n_agents % (world * block):
        raise ValueError(f"n_agents ({n_agents}) must be a multiple of ranks * {block} ({world * block})")
    li = np.arange(n_agents // world)
    return ((li // block) * world + rank) * block + li % block


def gather_cyclic(local: "torch.Tensor", world: int, block: int = 128, group=None) -> "torch.Tensor":
    """All-gather of cyclic tiles (`local` = this rank's slots in tile order) into slot order."""
    import torch
    import torch.distributed as dist
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=group)
    stacked = torch.stack([p.view(-1, block) for p in parts], dim=1)     # (blocks per rank, world, block)
    return stacked.reshape(-1)


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
    """One very large swarm (B = 1) across the ranks of a torch.distributed NCCL group.

    ``fused=True`` (default): the step kernel itself stores its tile's new records into every peer's record table over
    NVLink peer memory (CUDA IPC mappings of the tables, exchanged once through the process group) and the ranks
    hand-shake through flags in each other's memory -- no collective on the step path.  ``fused=False``: one in-place
    NCCL all-gather of the 16-byte records per step."""

    def __init__(self, n_agents: int, *, group=None, fused: bool = True, **engine_kwargs):
        import torch
        import torch.distributed as dist
        from .engine import VFEngine
        self.torch, self.dist, self.group = torch, dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.N = int(n_agents)
        self.fused = bool(fused) and self.world > 1
        self.begin, self.count = agent_tile(self.N, self.world, self.rank)
        # fused exchange: cyclic blocks of 128 slots (balanced work: a contiguous range of the sorted order is a region
        # of the arena, and the regions of the reference's disc initial condition differ in density by a factor of ten);
        # the in-place all-gather of the NCCL path needs the contiguous range
        self.cyclic = self.fused and self.N % (self.world * 128) == 0
        tile_kw = dict(tile_cycle=(self.world, self.rank)) if self.cyclic else \
            (dict(tile=(self.begin, self.count)) if self.world > 1 else {})
        self.engine = VFEngine(1, self.N, device=torch.cuda.current_device(), **tile_kw, **engine_kwargs)
        self._slots = torch.from_numpy(self.engine.tile_slots()).cuda()
        self._tables = {}
        if self.fused:
            mine = torch.frombuffer(bytearray(self.engine.ipc_export()), dtype=torch.uint8).cuda()
            every = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(every, mine, group=group)
            blob = every.cpu().numpy().tobytes()
            n = mine.numel()
            self.engine.ipc_attach(self.rank, [blob[r * n:(r + 1) * n] for r in range(self.world)])

    def _barrier(self):
        self.torch.cuda.synchronize()
        self.dist.barrier(group=self.group)

    def set_params(self, **kw):
        self.engine.set_params(**kw)

    def set_state(self, x, y, theta, vel, radius):
        """Every rank passes the FULL state (identical on all ranks)."""
        if self.fused:
            self._barrier()                  # nobody may still be stepping (and storing into our tables)
        self.engine.set_state(x, y, theta, vel, radius)
        if self.fused:
            self.engine.resort()             # the initial spatial sort, now: it rewrites both tables
            self._barrier()

    def _table(self):
        ptr, nbytes = self.engine.record_table_ptr()
        t = self._tables.get(ptr)
        if t is None:
            t = self.torch.as_tensor(_DeviceMemory(ptr, self.N * nbytes // 4), device="cuda").view(self.N, nbytes // 4)
            self._tables[ptr] = t
        return t

    def step(self, n_steps: int = 1):
        if self.fused:
            self.engine.step(n_steps)        # the kernels exchange the tiles themselves
            return
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
        if self.fused:
            self._barrier()
        for t in self._internal():
            if self.cyclic:
                t.copy_(gather_cyclic(t[self._slots], self.world, group=self.group))
            else:
                self.dist.all_gather_into_tensor(t, t[self.begin:self.begin + self.count].clone(), group=self.group)
        self.engine.resort()
        self._tables = {}
        if self.fused:
            self._barrier()

    def get_state(self):
        """Full state in the caller's agent order: (x, y) from the gathered record table, headings /
        speeds gathered here (off the step path) from the tiles and un-permuted."""
        if self.fused:
            self._barrier()                  # every rank's last step has landed in our table
        st = self.engine.get_state()
        out = {"x": st["x"][0], "y": st["y"][0]}
        perm = self.engine.permutation()[0]
        for k, t in zip(("theta", "vel"), self._internal()):
            if self.cyclic:
                full = gather_cyclic(t[self._slots], self.world, group=self.group).cpu().numpy()
            else:
                full = gather_tiles(t[self.begin:self.begin + self.count].clone(), self.world, self.group).cpu().numpy()
            api = np.empty_like(full)
            api[perm] = full
            out[k] = api
        return out
