import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abm_b200 import VFEngine
B, N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 1024
W = bench.arena_side(N)
x, y, th, v = bench.synthetic_state(B, N)
eng = VFEngine(B, N, resolution=1200, width=W, height=W)
eng.set_params(**bench.PARAMS)
eng.set_state(x, y, th, v, 10.0)
for _ in range(3): eng.step(1)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); eng.step(5); e1.record(); torch.cuda.synchronize()
print("ms/step", e0.elapsed_time(e1) / 5, eng.counters())
