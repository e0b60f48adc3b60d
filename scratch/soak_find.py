"""Find the first step at which the symmetric and the one-sided kernel disagree on a stored field (torus), dump the scene."""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abm_b200 import VFEngine
B, N, T = int(sys.argv[1]), 1024, int(sys.argv[2])
boundary = sys.argv[3] if len(sys.argv) > 3 else "infinite"
W = bench.arena_side(N)
x, y, th, v = bench.synthetic_state(B, N)
engs = {}
for k in ("symmetric", "onesided"):
    os.environ["ABM_VF_KERNEL"] = k
    e = VFEngine(B, N, resolution=1200, width=W, height=W, boundary=boundary, keep_fields=True)
    e.set_params(**bench.PARAMS); e.set_state(x, y, th, v, 10.0); e.step(0)
    engs[k] = e
found = 0
for t in range(T):
    st = engs["symmetric"].get_state()
    os.environ["ABM_VF_KERNEL"] = "onesided"
    engs["onesided"].set_state(st["x"], st["y"], st["theta"], st["vel"])
    engs["onesided"].step(1)
    os.environ["ABM_VF_KERNEL"] = "symmetric"
    engs["symmetric"].step(1)
    fa, fb = engs["symmetric"].fields_packed(), engs["onesided"].fields_packed()
    if not np.array_equal(fa, fb):
        bad = np.argwhere((fa != fb).any(axis=-1))
        print("step", t, "differing (replicate, agent):", bad[:8].tolist(), "of", len(bad))
        for (b, i) in bad[:4]:
            np.savez(f"gpurun_out/soak_diff_{found}.npz", x=st["x"][b], y=st["y"][b], theta=st["theta"][b], vel=st["vel"][b],
                     agent=i, sym=fa[b, i], one=fb[b, i], W=W, step=t)
            found += 1
        if found >= 4: break
print("done", t, found)
