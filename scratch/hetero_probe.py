"""BASELINE configs[3] shape with unequal radii: the symmetric kernel's two-half-width variant against the one-sided
kernel (ABM_VF_KERNEL=onesided) and against equal radii."""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from abm_b200 import VFEngine
import bench
B, N = 296, 1024
W = bench.arena_side(N)
x, y, th, v = bench.synthetic_state(B, N)
rng = np.random.default_rng(1)
rad_het = rng.choice([6.0, 8.0, 10.0, 12.0], (B, N)).astype(np.float32)
for name, rad, kern in (("equal radii, symmetric", 10.0, None), ("unequal radii, symmetric", rad_het, None),
                        ("unequal radii, one-sided", rad_het, "onesided")):
    if kern: os.environ["ABM_VF_KERNEL"] = kern
    else: os.environ.pop("ABM_VF_KERNEL", None)
    eng = VFEngine(B, N, resolution=1200, width=W, height=W)
    eng.set_params(**bench.PARAMS); eng.set_state(x, y, th, v, rad); eng.step(5)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.step(30); e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 30:.3f} ms/step ({B} replicates x {N}), kernel {eng.last_kernel()}", flush=True)
    eng.close()
