"""C5 on ONE GPU: what each of the 8 ranks of the tiled swarm would run (a tiled engine without peers: 8192 focal agents
against the full 65 536-record table), tile by tile, plus the whole swarm -- the per-rank kernel time is what bounds the
strong scaling (no exchange cost in here)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abm_b200 import VFEngine

def timed(fn, n):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

N = 65536; W = bench.arena_side(N)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
x, y, th, v = bench.synthetic_state(1, N)
eng = VFEngine(1, N, resolution=1200, width=W, height=W)
eng.set_params(**bench.PARAMS); eng.set_state(x, y, th, v, 10.0); eng.step(3)
print(f"whole swarm: {timed(eng.step, 20):.3f} ms/step ({eng.last_kernel()})")
eng.close()
cnt = N // G
ts = []
for k in range(G):
    e = VFEngine(1, N, resolution=1200, width=W, height=W, tile=(k * cnt, cnt))
    e.set_params(**bench.PARAMS); e.set_state(x, y, th, v, 10.0); e.resort(); e.step(2)
    ts.append(timed(e.step, 20)); e.close()
print(f"tiles of {cnt} focal agents (1/{G} of the swarm, frozen neighbours): " + " ".join(f"{t:.3f}" for t in ts) +
      f" ms/step; max {max(ts):.3f}, sum {sum(ts):.3f}")
