"""TEST INFRASTRUCTURE ONLY -- CPU oracle (float64, numpy) for the cooperative-signaling
variant of the projection field, restated from the reference's observable
behaviour: abm/projects/cooperative_signaling/cs_agent/cs_supcalc.py:204-307.

Same import rule as oracle/restate.py: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs may import it; the product path never does.

Pinning: (a) the reference's own golden vectors for exactly this function
(cs_agent/tests/test_cs_supcalc.py:143-168), (b) ``tests/golden/cs_golden.npz``, produced by
executing the UNMODIFIED reference function in the build container
(``tests/golden/make_golden_cs.py``), (c) random scenes against the live reference when the
tree is present (tests/test_oracle_cs.py).

Status: oracle only.  The CUDA path for this variant is not built yet;
what differs from the visual-flocking projection (oracle/restate.py::vf_intervals), and therefore
what the kernel needs as switches:

  * the object's centre is position + the FOCAL radius (:242), there are no object sizes;
  * visibility is decided on the ANGLE, closed interval fov[0] <= angle <= fov[1] (:260), not on
    the interval ends in bin units;
  * projections wider than ``max_proj_size`` bins are dropped whole (:268, :286-296);
  * a row is scaled by the object's meter value (:280-281): rows are float, not bits;
  * after the flip, bins whose linspace angle lies outside the FOV are zeroed (:287-288).
"""
from __future__ import annotations

import numpy as np

from .restate import TWO_PI, bin_grid, closed_angle_vf, heading_vector, nearest_bin


def cs_intervals(fov, R: int, position, radius: float, orientation: float, object_positions, max_proj_size=None):
    """Per-object interval data of cs_supcalc.projection_field (:228-275).

    Returns dict of arrays over the objects: k (centre bin, :256), h (= floor(proj_size / 2), :271-272),
    ps / pe (raw ends), proj_size (float, :265), drawn (passes :233, :260 and :268)."""
    pos = np.asarray(position, np.float64)
    obj = np.asarray(object_positions, np.float64).reshape(-1, 2)
    phis = bin_grid(R)                                                        # :221
    cix, ciy = pos[0] + radius, pos[1] + radius                               # :224
    v1x, v1y = heading_vector(pos[0], pos[1], radius, orientation)            # :227-231
    ojx, ojy = obj[:, 0] + radius, obj[:, 1] + radius                         # :242 focal radius
    v2x, v2y = ojx - cix, ojy - ciy                                           # :245
    valid = ~((obj[:, 0] == pos[0]) & (obj[:, 1] == pos[1]))                  # :233
    with np.errstate(divide="ignore", invalid="ignore"):
        ca = closed_angle_vf(v1x, v1y, v2x, v2y)                              # :248 (same function body as VF)
        dist = np.sqrt(v2x * v2x + v2y * v2y)                                 # :250
        vis_angle = 2 * np.arctan(radius / (1 * dist))                        # :253
    finite = np.isfinite(ca)
    k = nearest_bin(phis, np.where(finite, ca, 0.0))                          # :256
    in_fov = finite & (fov[0] <= ca) & (ca <= fov[1])                         # :260
    proj_size = (vis_angle / TWO_PI) * R                                      # :265
    ok = np.ones(len(obj), bool) if max_proj_size is None else (proj_size <= max_proj_size)   # :286-296
    h = np.floor(np.where(np.isfinite(proj_size), proj_size, 0.0) / 2).astype(np.int64)
    return dict(k=k, h=h, ps=k - h, pe=k + h, proj_size=proj_size, ca=ca, dist=dist,
                drawn=valid & in_fov & ok)


def cs_projection_field(fov, R: int, position, radius: float, orientation: float, object_positions,
                        object_meters=None, max_proj_size=None) -> np.ndarray:
    """Return value of cs_supcalc.projection_field (:204-289): (n_obj, R) float64 rows in STORED
    (flipped) order, scaled by the meters, masked to the FOV."""
    d = cs_intervals(fov, R, position, radius, orientation, object_positions, max_proj_size)
    n = len(d["k"])
    rows = np.zeros((n, R))
    for j in range(n):
        if not d["drawn"][j]:
            continue
        ps, pe = int(d["ps"][j]), int(d["pe"][j])
        if ps < 0:                                                            # :274-276
            rows[j, max(R + ps, 0):R] = 1
            ps = 0
        if pe >= R:                                                           # :277-279
            rows[j, 0:pe - (R - 1)] = 1
            pe = R
        rows[j, ps:pe] = 1                                                    # :281
        if object_meters is not None:                                         # :283-284
            rows[j] *= object_meters[j]
    post = rows[:, ::-1].copy()                                               # :288
    phis = bin_grid(R)
    post[:, phis < fov[0]] = 0                                                # :290-291
    post[:, phis > fov[1]] = 0
    return post
