"""How much of the symmetric kernel's step is the slow path?  The benchmark scene against a scene without a single
wide interval (agents on a jittered 32 x 32 grid, spacing 125 px > 112 px), same agent count and arena."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abm_b200 import VFEngine
B, N = 1024, 1024
W = 4200.0
rng = np.random.default_rng(3)
def timeit(x, y, th, v, tag):
    eng = VFEngine(B, N, resolution=1200, width=W, height=W)
    eng.set_params(GAM=0.0, V0=0.0, ALP0=0.0, ALP1=0.0, BET0=0.0, BET1=0.0)   # nobody moves: the scene stays what it is
    eng.set_state(x, y, th, v, 10.0)
    for _ in range(3): eng.step(1)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.step(10); e1.record(); torch.cuda.synchronize()
    ent, launches = eng.slow_entries()
    print(tag, "ms/step %.3f" % (e0.elapsed_time(e1) / 10), "kernel", eng.last_kernel(), "slow pairs/step/replicate %.0f" % (ent / launches / B),
          eng.counters())
    eng.close()
gx, gy = np.meshgrid(np.arange(32), np.arange(32))
x = (100 + 125 * gx.ravel()[None, :] + rng.uniform(-4, 4, (B, N))).astype(np.float32)
y = (100 + 125 * gy.ravel()[None, :] + rng.uniform(-4, 4, (B, N))).astype(np.float32)
# shuffle agents inside each replicate so that blocks hold a mix of near and far pairs
for b in range(B):
    p = rng.permutation(N); x[b] = x[b, p]; y[b] = y[b, p]
th = rng.uniform(0, 2 * np.pi, (B, N)).astype(np.float32); v = np.zeros_like(x)
timeit(x, y, th, v, "grid (no wide pair)")
xb, yb, thb, vb = bench.synthetic_state(B, N)
off = (W - bench.arena_side(N)) / 2
timeit(xb + off, yb + off, thb, vb, "benchmark disc")
