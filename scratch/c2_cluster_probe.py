"""C2 (one VF run of 100 agents): us per step of a multi-step launch, cluster barrier vs grid barrier (ABM_VF_NO_CLUSTER)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from abm_b200 import VFEngine
import bench
for N in (100, 64, 16):
    W = bench.arena_side(N)
    x, y, th, v = bench.synthetic_state(1, N)
    for mode in ("cluster", "grid"):
        if mode == "grid": os.environ["ABM_VF_NO_CLUSTER"] = "1"
        else: os.environ.pop("ABM_VF_NO_CLUSTER", None)
        eng = VFEngine(1, N, resolution=1200, width=W, height=W)
        eng.set_params(**bench.PARAMS); eng.set_state(x, y, th, v, 10.0); eng.step(200)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); eng.step(2000); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 2000 * 1e3)
        print(f"N={N} {mode}: {best:.2f} us/step, cluster launches {eng.cluster_launches()}", flush=True)
        eng.close()
