import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from abm_b200 import VFEngine
from oracle import restate as rs
d = np.load('/root/repo/tests/golden/torus_seam_cases.npz')
W = float(d["W"])
for c in (0, 1):
    x, y, th, v = (d[f"{k}{c}"][None, :] for k in ("x", "y", "theta", "vel"))
    i = int(d[f"agent{c}"])
    cfg = rs.VFConfig(R=1200, width=W, height=W, boundary="infinite")
    idx = sorted({i, 0, 17, 500})
    ref = rs.vf_step_frozen(x[0], y[0], th[0], v[0], 10.0, cfg, agents=idx)
    for kernel in ("symmetric", "onesided", "warp"):
        os.environ["ABM_VF_KERNEL"] = kernel
        eng = VFEngine(1, x.shape[1], resolution=1200, width=W, height=W, boundary="infinite", keep_fields=True)
        eng.set_params(); eng.set_state(x, y, th, v, 10.0); eng.step(1)
        f = eng.fields()[0]
        for a in idx:
            r = ref["rows"][a][::-1]
            print("case", c, kernel, "agent", a, "equal", np.array_equal(f[a], r), "diff bins", np.flatnonzero(f[a] != r)[:6], eng.last_kernel())
        eng.close()
