"""Per-step field comparison of two step kernels on the same evolving batch (kernel A's state is fed to kernel B every
step, so rounding-level differences of the fp64 terms cannot make the trajectories drift apart); dumps the scene of the
first differing fields.  usage: soak_find.py B T boundary kernelA kernelB [N] [R] [fov] [hetero]"""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abm_b200 import VFEngine
B, T = int(sys.argv[1]), int(sys.argv[2])
boundary = sys.argv[3] if len(sys.argv) > 3 else "infinite"
kA = sys.argv[4] if len(sys.argv) > 4 else "symmetric"
kB = sys.argv[5] if len(sys.argv) > 5 else "onesided"
N = int(sys.argv[6]) if len(sys.argv) > 6 else 1024
R = int(sys.argv[7]) if len(sys.argv) > 7 else 1200
fov = float(sys.argv[8]) if len(sys.argv) > 8 else 1.0
hetero = len(sys.argv) > 9 and sys.argv[9] == "hetero"
W = bench.arena_side(N)
rng = np.random.default_rng(B + N + R)
x, y, th, v = bench.synthetic_state(B, N)
rad = rng.choice([5.0, 10.0, 14.5], size=(B, N)).astype(np.float32) if hetero else 10.0
engs = {}
for k in (kA, kB):
    os.environ["ABM_VF_KERNEL"] = k
    e = VFEngine(B, N, resolution=R, width=W, height=W, boundary=boundary, fov=(-fov * np.pi, fov * np.pi), keep_fields=True)
    e.set_params(**bench.PARAMS); e.set_state(x, y, th, v, rad); e.step(0)
    engs[k] = e
found = 0
for t in range(T):
    st = engs[kA].get_state()
    os.environ["ABM_VF_KERNEL"] = kB
    engs[kB].set_state(st["x"], st["y"], st["theta"], st["vel"])
    engs[kB].step(1)
    os.environ["ABM_VF_KERNEL"] = kA
    engs[kA].step(1)
    fa, fb = engs[kA].fields_packed(), engs[kB].fields_packed()
    if not np.array_equal(fa, fb):
        bad = np.argwhere((fa != fb).any(axis=-1))
        print("step", t, "differing (replicate, agent):", bad[:8].tolist(), "of", len(bad))
        for (b, i) in bad[:4]:
            np.savez(f"gpurun_out/soak_diff_{found}.npz", x=st["x"][b], y=st["y"][b], theta=st["theta"][b], vel=st["vel"][b],
                     agent=i, A=fa[b, i], B=fb[b, i], W=W, step=t, rad=(rad[b] if hetero else rad), R=R, fov=fov)
            found += 1
        if found >= 4: break
print("done", " ".join(sys.argv[1:]), "| last kernels", engs[kA].last_kernel(), engs[kB].last_kernel(), "| steps", t + 1, "| differing fields found:", found)
