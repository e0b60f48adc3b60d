import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abm_b200 import VFEngine
N = 65536; W = bench.arena_side(N)
x, y, th, v = bench.synthetic_state(1, N)
eng = VFEngine(1, N, resolution=1200, width=W, height=W)
eng.set_params(**bench.PARAMS); eng.set_state(x, y, th, v, 10.0)
eng.step(6); torch.cuda.synchronize()
