"""Run under torchrun on >= 2 GPUs of one node (NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py
Large-swarm mode: every rank updates its agent tile, tiles are exchanged with one in-place
all-gather of the 16-byte records per step; the result must equal a single-GPU engine's bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    from abm_b200 import VFEngine
    from abm_b200.multigpu import TiledSwarm
    N = 4096 if len(sys.argv) < 2 else int(sys.argv[1])
    steps = 12
    W = float(np.ceil(900 * np.sqrt(N / 100)))
    rng = np.random.default_rng(99)
    th = rng.uniform(0, 2 * np.pi, N).astype(np.float32)
    rho = rng.uniform(0, 1, N) * (W / 2 - 70)
    x = (W / 2 + rho * np.cos(th)).astype(np.float32)
    y = (W / 2 + rho * np.sin(th)).astype(np.float32)
    v = np.zeros(N, np.float32)
    ok = True
    for boundary, fused in (("infinite", True), ("infinite", False), ("walls", True)):
        kw = dict(resolution=1200, width=W, height=W, boundary=boundary)
        ref_state = None
        if rank == 0:
            ref = VFEngine(1, N, **kw)
            ref.set_params(); ref.set_state(x[None], y[None], th[None], v[None], 10.0); ref.step(steps)
            ref_state = ref.get_state()
            ref.close()
        swarm = TiledSwarm(N, fused=fused, **kw)
        swarm.set_params()
        swarm.set_state(x[None], y[None], th[None], v[None], 10.0)
        swarm.step(steps // 2)
        swarm.resync()                          # headings / speeds to every rank + spatial re-sort, mid-run
        swarm.step(steps - steps // 2)
        got = swarm.get_state()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        dist.barrier()
        e0.record(); swarm.step(10); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if rank == 0:
            for k in ("x", "y", "theta", "vel"):
                same = np.array_equal(got[k], ref_state[k][0])
                ok &= same
                print(f"[tiled x{world} {boundary} {'fused peer stores' if fused else 'NCCL all-gather'}] {k}: "
                      f"{'bit-identical' if same else 'MISMATCH'}", flush=True)
            print(f"[tiled x{world} {boundary} {'fused peer stores' if fused else 'NCCL all-gather'}"
                  f"{' cyclic tiles' if swarm.cyclic else ''}] N={N}: {ms:.3f} ms/step", flush=True)
        dist.barrier()
        swarm.engine.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if not bool(flag.item()):
        sys.exit(1)
    if rank == 0:
        print("MULTIGPU_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()
