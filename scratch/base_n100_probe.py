"""The reference's figure experiments (figExp0/1/2A: N = 100 foragers, occlusion + collisions): one fused launch per step
against one grid per phase (ABM_BASE_FUSED=1 / 0), B replicates."""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from abm_b200 import BaseEngine
P, W = 10, 500.0
N = int(os.environ.get("PROBE_N", "100"))
P = int(os.environ.get("PROBE_P", "10"))
for B in [int(b) for b in (sys.argv[1:] or ["256", "1024"])]:
    rng = np.random.default_rng(3)
    x0, y0 = rng.integers(20, 520, (B, N)), rng.integers(20, 520, (B, N))
    th0 = rng.uniform(0, 2 * np.pi, (B, N))
    pa = dict(x=rng.integers(60, 400, (B, P)), y=rng.integers(60, 400, (B, P)), radius=np.full((B, P), 30.0),
              left=np.full((B, P), 200.0), quality=np.full((B, P), 0.25), id=np.tile(np.arange(P), (B, 1)))
    for mode in ("fused", "separate"):
        os.environ["ABM_BASE_FUSED"] = "0" if mode == "separate" else "1"      # (forced: the engine's own rule picks between them)
        eng = BaseEngine(B, N, P, resolution=1200, width=W, height=W, visual_exclusion=True, collide_agents=True, ghost_mode=False, seed=9)
        eng.set_params(Eps_w=2.0, Eps_u=1.0, F_N=0.5, F_R=0.5, exp_vel_max=3.0, exp_theta_min=-0.5, exp_theta_max=0.5,
                       reloc_theta_max=1.8, exp_stop_ratio=0.175)
        eng.set_agents(x=x0, y=y0, theta=th0); eng.set_patches(**pa)
        eng.step(50); torch.cuda.synchronize()
        l0 = eng.counters()["launches"]
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.step(100); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 100
        print(f"B={B} N={N} {mode}: {ms:.3f} ms/step = {B * N / ms * 1e3:.3g} agent-steps/s, launches per step {(eng.counters()['launches'] - l0) / 100:.0f}", flush=True)
        eng.close()
