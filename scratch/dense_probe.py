import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abm_b200 import VFEngine
B, N = 296, 1024
for radius_px in (1370.0, 900.0, 600.0, 450.0, 300.0, 150.0):
    rng = np.random.default_rng(0)
    ang = rng.uniform(0, 2 * np.pi, (B, N)); rr = np.sqrt(rng.uniform(0, 1, (B, N))) * radius_px
    x = (1440 + rr * np.cos(ang)).astype(np.float32); y = (1440 + rr * np.sin(ang)).astype(np.float32)
    th = rng.uniform(0, 2 * np.pi, (B, N)).astype(np.float32); v = np.zeros_like(x)
    for kern in ("symmetric", "symmetric_wide", "onesided", "auto"):
        import os
        os.environ.pop("ABM_VF_SYM_WIDE", None)
        if kern == "auto": os.environ.pop("ABM_VF_KERNEL", None)
        elif kern.startswith("symmetric"): os.environ["ABM_VF_KERNEL"] = "symmetric"; os.environ["ABM_VF_SYM_WIDE"] = "1" if kern.endswith("wide") else "0"
        else: os.environ["ABM_VF_KERNEL"] = kern
        eng = VFEngine(B, N, resolution=1200, width=2880.0, height=2880.0)
        eng.set_params(**bench.PARAMS); eng.set_state(x, y, th, v, 10.0)
        eng.step(1); torch.cuda.synchronize()
        if kern == "auto": eng.step(3); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        eng.set_state(x, y, th, v, 10.0)
        e0.record(); eng.step(1); e1.record(); torch.cuda.synchronize()
        c = eng.counters(); ne, nl = eng.slow_entries()
        frac = ne / max(nl, 1) / (0.25 * B * N * (N - 1))
        print(f"disc radius {radius_px:6.0f}px {kern:14s}: {e0.elapsed_time(e1):8.3f} ms/step (296 replicates), fp64 pairs {c['fp64_pairs'] // c['launches']}, slow entry fraction {frac:.3f}, last kernel {eng.last_kernel()} {eng.kernel_stats()}")
        eng.close()
