"""A/B probe of a symmetric-kernel variant (scratch/libs/libA.so = the tree's library, libB.so = the variant; the caller
copies one over abm_b200/libabm_b200.so): ms per step on the benchmark workload, and -- `check` -- the variant's fields,
terms and new state against the one-sided kernel's (an independent implementation), bit for bit, over three steps."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abm_b200 import VFEngine

tag = sys.argv[1]; check = len(sys.argv) > 2
N = 1024; W = bench.arena_side(N); B = 1024
x, y, th, v = bench.synthetic_state(B, N)
os.environ['ABM_VF_KERNEL'] = 'symmetric'
eng = VFEngine(B, N, resolution=1200, width=W, height=W)
eng.set_params(**bench.PARAMS); eng.set_state(x, y, th, v, 10.0)
eng.step(5); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); eng.step(40); e1.record(); torch.cuda.synchronize()
print(f"lib {tag}: {e0.elapsed_time(e1) / 40:.4f} ms/step ({eng.last_kernel()})", flush=True)
eng.close()
if check:
    Bc = 96; res = {}
    for kern in ("onesided", "symmetric"):
        os.environ['ABM_VF_KERNEL'] = kern
        e = VFEngine(Bc, N, resolution=1200, width=W, height=W, keep_fields=True, keep_terms=True, spatial_sort=False)
        e.set_params(**bench.PARAMS); e.set_state(x[:Bc], y[:Bc], th[:Bc], v[:Bc], 10.0)
        e.step(1); f1, t1 = e.fields_packed().copy(), e.terms().copy()
        e.step(2)
        res[kern] = (f1, t1, e.fields_packed().copy(), e.terms().copy(), e.get_state(), e.last_kernel(), e.counters())
        e.close()
    a, b = res["onesided"], res["symmetric"]
    ok = all(np.array_equal(a[k], b[k]) for k in range(4)) and all(np.array_equal(a[4][k], b[4][k]) for k in ("x", "y", "theta", "vel"))
    print(f"lib {tag}: {b[5]} vs {a[5]}: fields / terms / state of {Bc} x {N} agents over 3 steps bit-identical: {ok}; fp64 pairs {a[6]['fp64_pairs']} / {b[6]['fp64_pairs']}", flush=True)
