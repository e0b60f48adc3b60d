"""Timing probes of the other BASELINE configs (not bench lines): C3 BASE sweep 1024 x 50, C5 swarm N = 65536 (one GPU)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from abm_b200 import BaseEngine, VFEngine
import bench

def timed(fn, n):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# ---- C3: metaprotocol sweep, 1024 replicates x 50 agents foraging, social cues + occlusion + collisions ----
B, N, P, W = 1024, 50, 3, 500.0
rng = np.random.default_rng(3)
x0, y0 = rng.integers(20, 520, (B, N)), rng.integers(20, 520, (B, N))
th0 = rng.uniform(0, 2 * np.pi, (B, N))
pa = dict(x=rng.integers(60, 400, (B, P)), y=rng.integers(60, 400, (B, P)), radius=np.full((B, P), 30.0),
          left=np.full((B, P), 200.0), quality=np.full((B, P), 0.25), id=np.tile(np.arange(P), (B, 1)))
eng = BaseEngine(B, N, P, resolution=1200, width=W, height=W, visual_exclusion=True, collide_agents=True, ghost_mode=False, seed=9)
eps = np.tile(np.array([0, .25, .5, .75, 1, 2, 5, 3], np.float64), B // 8)      # a DEC_EPSW sweep, one set per replicate
eng.set_params(Eps_w=eps, Eps_u=1.0, F_N=0.5, F_R=0.5, exp_vel_max=3.0, exp_theta_min=-0.5, exp_theta_max=0.5,
               reloc_theta_max=1.8, exp_stop_ratio=0.175)
eng.set_agents(x=x0, y=y0, theta=th0); eng.set_patches(**pa)
eng.step(50)
ms = timed(eng.step, 500)
print(f"C3 BASE 1024 x 50, 3 patches, occlusion + collisions: {ms:.4f} ms/step = {B * N / ms * 1e3:.3e} agent-steps/s (3 launches per step)")
a = eng.get_agents(); print("   modes:", np.bincount(a["mode"].ravel(), minlength=4), "finite:", bool(np.isfinite(a["x"]).all()))
eng.close()

# ---- C5: single large swarm N = 65536 on ONE GPU (the multi-GPU run shards the focal agents) ----
N = 65536; W = bench.arena_side(N)
x, y, th, v = bench.synthetic_state(1, N)
for sort in (True,):
    eng = VFEngine(1, N, resolution=1200, width=W, height=W, spatial_sort=sort)
    eng.set_params(**bench.PARAMS); eng.set_state(x, y, th, v, 10.0)
    eng.step(3)
    ms = timed(eng.step, 20)
    print(f"C5 VF swarm N=65536 arena {W:.0f}px one GPU, spatial_sort={sort}: {ms:.3f} ms/step = {N / ms * 1e3:.3e} agent-steps/s, kernel {eng.last_kernel()}")
    eng.close()
# ---- C2: N = 100 single run ----
N = 100; W = bench.arena_side(N)
x, y, th, v = bench.synthetic_state(1, N)
eng = VFEngine(1, N, resolution=1200, width=W, height=W)
eng.set_params(**bench.PARAMS); eng.set_state(x, y, th, v, 10.0); eng.step(10)
ms = timed(eng.step, 2000)
print(f"C2 VF N=100 single run: {ms * 1e3:.1f} us/step = {N / ms * 1e3:.3e} agent-steps/s, kernel {eng.last_kernel()}")
eng.close()
