import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from abm_b200 import VFEngine
B, N = 1024, 1024
W = bench.arena_side(N)
x, y, th, v = bench.synthetic_state(B, N)
def morton(xq, yq):
    def part(a):
        a = a.astype(np.uint32) & 0xffff
        a = (a | (a << 8)) & 0x00FF00FF; a = (a | (a << 4)) & 0x0F0F0F0F
        a = (a | (a << 2)) & 0x33333333; a = (a | (a << 1)) & 0x55555555
        return a
    return part(xq) | (part(yq) << 1)
def run(x, y, th, v, label):
    eng = VFEngine(B, N, resolution=1200, width=W, height=W)
    eng.set_params(**bench.PARAMS)
    eng.set_state(x, y, th, v, 10.0)
    for _ in range(3): eng.step(1)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.step(20); e1.record(); torch.cuda.synchronize()
    print(label, e0.elapsed_time(e1) / 20, "ms/step")
    eng.close()
run(x, y, th, v, "random order")
key = morton((x / W * 1023).astype(np.int64), (y / W * 1023).astype(np.int64))
order = np.argsort(key, axis=1)
take = lambda a: np.take_along_axis(a, order, axis=1)
run(take(x), take(y), take(th), take(v), "morton order")
