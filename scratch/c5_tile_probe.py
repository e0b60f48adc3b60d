"""C5 on ONE GPU: what each of the G ranks of the tiled swarm would run -- a tiled engine without peers, N / G focal agents
against the full 65 536-record table; after every step the other tiles' (frozen) records are copied into the table the
step wrote, as the exchange would -- tile by tile, plus the whole swarm.  The per-rank kernel time is what bounds the strong
scaling (the exchange itself is 1 MiB per step).  ABM_VF_WARP_FOCAL=1|2|4|8 forces the focal agents per CTA."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abm_b200 import VFEngine
from abm_b200.multigpu import _DeviceMemory

N = 65536; W = bench.arena_side(N)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
x, y, th, v = bench.synthetic_state(1, N)

def table(eng):
    ptr, nb = eng.record_table_ptr()
    return torch.as_tensor(_DeviceMemory(ptr, N * nb // 4), device="cuda").view(N, nb // 4)

def run(eng, steps, lo, hi, timed):
    ms = 0.0
    for _ in range(steps):
        prev = table(eng)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.step(1); e1.record()
        new = table(eng)
        if lo > 0: new[:lo].copy_(prev[:lo])
        if hi < N: new[hi:].copy_(prev[hi:])
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / steps

eng = VFEngine(1, N, resolution=1200, width=W, height=W)
eng.set_params(**bench.PARAMS); eng.set_state(x, y, th, v, 10.0); run(eng, 3, 0, N, False)
print(f"whole swarm: {run(eng, 20, 0, N, True):.3f} ms/step ({eng.last_kernel()})", flush=True)
eng.close()
cnt = N // G
ts = []
for k in range(G):
    e = VFEngine(1, N, resolution=1200, width=W, height=W, tile=(k * cnt, cnt))
    e.set_params(**bench.PARAMS); e.set_state(x, y, th, v, 10.0); e.resort(); run(e, 3, k * cnt, (k + 1) * cnt, False)
    ts.append(run(e, 20, k * cnt, (k + 1) * cnt, True)); e.close()
print(f"tiles of {cnt} focal agents (1/{G} of the swarm): " + " ".join(f"{t:.3f}" for t in ts) +
      f" ms/step; max {max(ts):.3f}, sum {sum(ts):.3f}", flush=True)
