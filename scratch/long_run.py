"""Soak run: 2000 steps of the benchmark workload with summary metrics every 250 steps (kernel choice may adapt)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abm_b200 import VFEngine
B, N = 1024, 1024
W = bench.arena_side(N)
x, y, th, v = bench.synthetic_state(B, N)
eng = VFEngine(B, N, resolution=1200, width=W, height=W)
eng.set_params(**bench.PARAMS); eng.set_state(x, y, th, v, 10.0)
t0 = time.perf_counter()
for k in range(8):
    eng.step(250); torch.cuda.synchronize()
    m = eng.metrics(); st = eng.get_state()
    print(f"t={250 * (k + 1):5d}  {time.perf_counter() - t0:6.2f} s  kernel {eng.last_kernel():26s} polarization {m['polarization'].mean():.3f} "
          f"mean iid {m['mean_iid'].mean():7.1f} px  nn dist {m['mean_nn_dist'].mean():5.1f} px  collision frac {m['collision'].mean():.2f}  "
          f"finite {bool(np.isfinite(st['x']).all() and np.isfinite(st['theta']).all())}")
print(eng.counters(), eng.slow_entries())
