"""C2 (VF, N = 100, one replicate) and small batches: which step kernel is fastest when the grid cannot fill the GPU?"""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abm_b200 import VFEngine
def timed(fn, n):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B, N in [tuple(map(int, a.split("x"))) for a in sys.argv[1:]] or ((1, 100), (1, 1024), (16, 100), (64, 256), (148, 1024), (16, 1024)):
    W = bench.arena_side(N)
    x, y, th, v = bench.synthetic_state(B, N)
    for k in ("symmetric", "onesided", "warp", ""):
        if k: os.environ["ABM_VF_KERNEL"] = k
        else: os.environ.pop("ABM_VF_KERNEL", None)
        eng = VFEngine(B, N, resolution=1200, width=W, height=W)
        eng.set_params(**bench.PARAMS); eng.set_state(x, y, th, v, 10.0); eng.step(10)
        ms = timed(eng.step, 300)
        print(f"B={B} N={N} forced={k or 'auto':10s} {ms * 1e3:8.1f} us/step  kernel {eng.last_kernel()}")
        eng.close()
