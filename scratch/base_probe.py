import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from abm_b200 import BaseEngine
B, N, P, W = 1024, 50, 3, 500.0
rng = np.random.default_rng(3)
x0, y0 = rng.integers(20, 520, (B, N)), rng.integers(20, 520, (B, N))
th0 = rng.uniform(0, 2 * np.pi, (B, N))
pa = dict(x=rng.integers(60, 400, (B, P)), y=rng.integers(60, 400, (B, P)), radius=np.full((B, P), 30.0),
          left=np.full((B, P), 200.0), quality=np.full((B, P), 0.25), id=np.tile(np.arange(P), (B, 1)))
eng = BaseEngine(B, N, P, resolution=1200, width=W, height=W, visual_exclusion=True, collide_agents=True, ghost_mode=False, seed=9)
eng.set_params(Eps_w=2.0, Eps_u=1.0, F_N=0.5, F_R=0.5, exp_vel_max=3.0, exp_theta_min=-0.5, exp_theta_max=0.5,
               reloc_theta_max=1.8, exp_stop_ratio=0.175)
eng.set_agents(x=x0, y=y0, theta=th0); eng.set_patches(**pa)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
eng.step(5)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); eng.step(n); e1.record(); torch.cuda.synchronize()
print('BASE C3 probe: %.3f ms/step, %.3g agent-steps/s' % (e0.elapsed_time(e1) / n, B * N * n / e0.elapsed_time(e1) * 1e3))
