"""Empirical fp32 error of the symmetric kernel's bin coordinate t = wrap(phi - theta) * (R-1)/2pi + t_half
(numpy float32 emulation of the exact instruction sequence; fma emulated in float64)."""
import numpy as np
f32 = np.float32
def fma(a, b, c): return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)
R = 1200
inv = (R - 1) / (2 * np.pi)
A = [0.9999993443489075, -0.33326515555381775, 0.19881492853164673, -0.13487225770950317,
     0.0838717594742775, -0.037013452500104904, 0.007863515056669712]
ac = [f32(a * inv) for a in A]
half_pi_b, pi_b = f32(0.5 * np.pi * inv), f32(np.pi * inv)
rng = np.random.default_rng(1)
n = 4_000_000
worst = 0
for rep in range(5):
    xi = rng.uniform(30, 2910, n).astype(f32); yi = rng.uniform(30, 2910, n).astype(f32)
    # mix of far and near partners
    d = np.where(rng.random(n) < 0.5, rng.uniform(20, 3000, n), rng.uniform(20, 200, n))
    ang = rng.uniform(0, 2 * np.pi, n)
    ox = (xi + d * np.cos(ang)).astype(f32); oy = (yi + d * np.sin(ang)).astype(f32)
    th = rng.uniform(0, 2 * np.pi, n).astype(f32)
    dx = (ox - xi).astype(f32); dy = (oy - yi).astype(f32)
    au, aw = np.abs(dx), np.abs(dy)
    mx, mn = np.maximum(au, aw), np.minimum(au, aw)
    rcp = (1.0 / mx.astype(np.float64)).astype(f32)
    # MUFU.RCP: up to 1 ulp off; emulate a random +-1 ulp perturbation
    rcp = np.nextafter(rcp, np.where(rng.random(n) < 0.5, f32(np.inf), f32(-np.inf))).astype(f32)
    tq = (mn * rcp).astype(f32)
    z = (tq * tq).astype(f32)
    p = fma(np.full(n, ac[6], f32), z, np.full(n, ac[5], f32))
    for c in (4, 3, 2, 1, 0):
        p = fma(p, z, np.full(n, ac[c], f32))
    p = (p * tq).astype(f32)
    p = np.where(aw > au, (half_pi_b - p).astype(f32), p)
    pr = (pi_b - p).astype(f32)
    pi_abs = np.where(dx < 0, pr, p)
    phi = np.where(-dy < 0, -pi_abs, pi_abs).astype(f32)     # copysign(pi_abs, -dy)
    thb = ((th.astype(np.float64) % (2 * np.pi)) * inv).astype(f32)
    cab = (phi - thb).astype(f32)
    cab = np.where(cab < -pi_b, (cab + f32(2) * pi_b).astype(f32), cab)
    t32 = (cab + f32(0.5)).astype(np.float64)
    # exact
    dxe = ox.astype(np.float64) - xi.astype(np.float64); dye = oy.astype(np.float64) - yi.astype(np.float64)
    ca = np.arctan2(-dye, dxe) - th.astype(np.float64)
    ca = (ca + np.pi) % (2 * np.pi) - np.pi
    te = ca * inv + 0.5
    err = np.abs(t32 - te)
    err = np.minimum(err, np.abs(err - (R - 1)))          # seam
    worst = max(worst, err.max())
    print(rep, "max |t32 - t| =", err.max(), " 99.99% =", np.quantile(err, 0.9999), " mean =", err.mean())
print("worst", worst, "-> current tau_k_sym", 2e-6 * R / (2 * np.pi) + 2.5e-7 * R + 1e-5 + 3e-7 * R)
