// Warp-level building blocks of the BASE visual field (Agent.projection_field, agent.py:457-597),
// shared by the agent phase, the collision phase and the stateless function-level kernel.
#pragma once
#include "abm_base.cuh"
#include "abm_vf_device.cuh"

namespace abm {

struct __align__(16) ObjRec {
  int s, e;      // raw interval ends, int() truncated (agent.py:545-546)
  double d;      // centre distance (agent.py:526-528)
};

// per-warp shared-memory work area for up to N objects and a row of W words
struct WarpField {
  ObjRec* raw;      // recorded objects in list order
  ObjRec* sorted;   // ranked by (distance, list order)
  int* key;         // list-order key of raw[m]; bit 30 = social cue (drawn), else occluder only
  int* skey;        // same for sorted[rank]
  uint32_t* row;    // un-flipped field v (agent.py:480), W + 1 words
};
__host__ __device__ inline size_t warp_field_bytes(int N, int W) {
  return (2 * sizeof(ObjRec) * (size_t)N + 2 * sizeof(int) * (size_t)N + sizeof(uint32_t) * (size_t)(W + 1) + 15) / 16 * 16;
}
__device__ __forceinline__ WarpField warp_field_at(unsigned char* base, int N) {
  WarpField wf;
  wf.raw = reinterpret_cast<ObjRec*>(base);
  wf.sorted = wf.raw + N;
  wf.key = reinterpret_cast<int*>(wf.sorted + N);
  wf.skey = wf.key + N;
  wf.row = reinterpret_cast<uint32_t*>(wf.skey + N);
  return wf;
}

// supcalc.distance between agent centres, both with the focal radius (supcalc.py:73-78, agent.py:504-509)
__device__ __forceinline__ double base_centre_distance(const FocalExact& fe, double r, float xj, float yj) {
  const double v2x = __dadd_rn(__dadd_rn((double)xj, r), -fe.cix), v2y = __dadd_rn(__dadd_rn((double)yj, r), -fe.ciy);
  return __dsqrt_rn(__dadd_rn(__dmul_rn(v2x, v2x), __dmul_rn(v2y, v2y)));
}

// the same from a focal centre (cix, ciy) = position + radius already formed in float64, object at (xj, yj) with
// radius rj
__device__ __forceinline__ double base_distance_exact(double cix, double ciy, double rj, float xj, float yj) {
  const double v2x = __dadd_rn(__dadd_rn((double)xj, rj), -cix), v2y = __dadd_rn(__dadd_rn((double)yj, rj), -ciy);
  return __dsqrt_rn(__dadd_rn(__dmul_rn(v2x, v2x), __dmul_rn(v2y, v2y)));
}

// One obstacle of Agent.projection_field (agent.py:501-556), fp64, the reference's operation
// order.  Returns true if the obstacle is recorded (not at the focal position, strictly inside
// the FOV); `dist` is set whenever the position test passes (the loop variable that
// keep_distance_info leaks, agent.py:528 / :590), else left untouched.
__device__ __forceinline__ bool base_interval(const FocalExact& fe, double r, float xi, float yi, float xj, float yj,
                                              double fov0, double fov1, int R, double lin_step, ObjRec& o,
                                              double& dist) {
  if ((xj == xi) && (yj == yi)) return false;                                       // :502
  const double v2x = __dadd_rn(__dadd_rn((double)xj, r), -fe.cix), v2y = __dadd_rn(__dadd_rn((double)yj, r), -fe.ciy);
  const double n2 = __dsqrt_rn(__dadd_rn(__dmul_rn(v2x, v2x), __dmul_rn(v2y, v2y)));
  dist = n2;
  if (!(n2 > 0.0)) return false;
  const double u2x = __ddiv_rn(v2x, n2), u2y = __ddiv_rn(v2y, n2);
  double dot = __dadd_rn(__dmul_rn(fe.u1x, u2x), __dmul_rn(fe.u1y, u2y));
  dot = fmin(1.0, fmax(-1.0, dot));
  double ang = acos(dot);                                                           // supcalc.py:31
  if (__dadd_rn(__dmul_rn(fe.u1x, u2y), -__dmul_rn(fe.u1y, u2x)) < 0.0) ang = -ang;
  if (ang < 0.0) ang = __dadd_rn(ang, ABM_TWO_PI_D);                                // % 2pi (agent.py:516)
  const double ca = (ang > 0.0 && ang < ABM_PI_D) ? -ang : __dadd_rn(ABM_TWO_PI_D, -ang);   // :520-523
  if (!(fov0 < ca && ca < fov1)) return false;                                      // :535
  const int k = nearest_bin_exact(ca, R, lin_step);                                 // :532
  const double vis = __dmul_rn(2.0, atan(__ddiv_rn(r, n2)));                        // :529
  const double size = __dmul_rn(__ddiv_rn(vis, ABM_TWO_PI_D), (double)R);           // :543
  const double half = __ddiv_rn(size, 2.0);
  o.s = (int)__dadd_rn((double)k, -half);                                           // :545-546 int(): toward zero
  o.e = (int)__dadd_rn((double)k, half);
  o.d = n2;
  return true;
}

// ---------------------------------------------------------------------------------------
// fp32 fast path of the same obstacle (the scheme of the visual-flocking kernels): closed angle by rotation into the
// heading frame + polynomial arctangent, bin index and half width by the 1.5 * 2^23 rounding constant, and a guard band
// around every decision the float64 reference takes on a real number -- rounding tie of the bin index, integer
// crossings of the half width (int() truncation of k -+ size / 2), the strict FOV bounds, the dead-ahead / dead-astern
// angles the reference makes invisible (agent.py:520-523).  A pair inside a guard band is re-evaluated by
// base_interval (fp64, the reference's own operation sequence); everything else has the reference's integers.
// ---------------------------------------------------------------------------------------
struct BaseFast {
  float c, ns;                 // cos / -sin of the focal heading
  float r;                     // focal radius (both centres, agent.py:504-509)
  float fov0, fov1;            // strict bounds on the closed angle (agent.py:535)
  float inv_step, t_half;      // (R - 1) / 2pi; 0.5 (even R) or 1.0 (odd R)
  int k_off;                   // R/2 - 1 or (R - 3)/2: k = rint(ca * inv_step + t_half) + k_off
  float y_scale;               // R / 2pi: size / 2 = atan(r / d) * y_scale
  float thr_k, thr_h0, thr_h1; // guard bands in bins (abm_vf_device.cuh: BinConsts)
};
constexpr float kBaseAngleGuard = 4.0e-6f;   // rad: error bound of the fp32 closed angle (polynomial 1e-6 + rotation)

__device__ __forceinline__ BaseFast base_fast_consts(float theta, double r, double fov0, double fov1, int R) {
  BaseFast f;
  float sn, cs;
  sincosf(theta, &sn, &cs);
  f.c = cs; f.ns = -sn;
  f.r = (float)r;
  f.fov0 = (float)fov0; f.fov1 = (float)fov1;
  f.inv_step = (float)((double)(R - 1) / ABM_TWO_PI_D);
  f.t_half = (R & 1) ? 1.0f : 0.5f;
  f.k_off = (R & 1) ? (R - 3) / 2 : R / 2 - 1;
  f.y_scale = (float)((double)R / ABM_TWO_PI_D);
  const float tau_k = (float)(2.0e-6 * (double)R / ABM_TWO_PI_D + 2.5e-7 * (double)R + 1e-5);
  f.thr_k = 0.5f - tau_k;
  f.thr_h0 = 0.5f - 2.0e-5f - 0.5f * 3.0e-6f;
  f.thr_h1 = -3.0e-6f;
  return f;
}

// (dx, dy): position difference object - focal (fp32, one rounding).  Returns true when the pair must be re-evaluated
// in fp64; otherwise `visible` and the raw interval ends (s, e) are the reference's.
__device__ __forceinline__ bool base_interval_fast(float dx, float dy, const BaseFast& f, int& s, int& e, bool& visible) {
  const float u = fmaf(dx, f.c, dy * f.ns);            //  dx cos - dy sin
  const float w = fmaf(dx, f.ns, -(dy * f.c));         // -dx sin - dy cos: closed angle = atan2(w, u)
  const float ca = atan2_fast(w, u);
  const float aca = fabsf(ca);
  bool flagged = (aca < kBaseAngleGuard) | (aca > 3.14159265f - kBaseAngleGuard);        // agent.py:520-523
  flagged |= (fabsf(ca - f.fov0) < kBaseAngleGuard) | (fabsf(ca - f.fov1) < kBaseAngleGuard);
  visible = (ca > f.fov0) & (ca < f.fov1);                                               // :535
  const float t = fmaf(ca, f.inv_step, f.t_half);
  const float tr = t + kMagic;
  const int k = __float_as_int(tr) - kMagicBits + f.k_off;                               // :532
  flagged |= fabsf(t - (tr - kMagic)) > f.thr_k;
  const float d2 = fmaf(dx, dx, dy * dy);
  const float q = f.r * rsqrt_approx(d2);                                                // r / d
  const bool big = q > 1.0f;
  float at = atan_unit(big ? rcp_approx(q) : q);
  if (big) at = 1.57079632679489662f - at;
  const float y = fmaf(at, f.y_scale, -0.5f);                                            // size / 2 - 0.5 (:529, :543)
  const float yr = y + kMagic;
  const int hf = __float_as_int(yr) - kMagicBits;                                        // floor(size / 2)
  flagged |= !(fabsf(y - (yr - kMagic)) <= fmaf(y, f.thr_h1, f.thr_h0));                 // also NaN (d == 0)
  // int(k - size/2), int(k + size/2): truncation toward zero (:545-546); size/2 is not an integer here
  s = (k >= hf + 1) ? k - hf - 1 : k - hf;
  e = k + hf;
  return flagged;
}

// The fp64 re-evaluation of a pair inside a guard band, out of line: ONE copy of the (large: software double-precision
// sincos, acos, atan) code for every caller, and none of it inside the callers' hot loops.
static __device__ __noinline__ bool base_interval_cold(float xi, float yi, double r, float theta, float xj, float yj,
                                                       double fov0, double fov1, int R, double lin_step, int& s, int& e,
                                                       double& dist) {
  const FocalExact fe = vf_focal_exact(xi, yi, (float)r, theta);
  ObjRec o; o.s = 0; o.e = 0; o.d = 0.0;
  const bool rec = base_interval(fe, r, xi, yi, xj, yj, fov0, fov1, R, lin_step, o, dist);
  s = o.s; e = o.e;
  return rec;
}

// append this lane's object (if `rec`) in lane order; returns the new object count
__device__ __forceinline__ int base_record(WarpField& wf, int M, bool rec, const ObjRec& o, int key, int lane) {
  const unsigned mask = __ballot_sync(0xffffffffu, rec);
  if (rec) {
    const int idx = M + __popc(mask & ((1u << lane) - 1u));
    wf.raw[idx] = o;
    wf.key[idx] = key;
  }
  return M + __popc(mask);
}

// numpy basic-slice bounds of v[a:b] on a length-R array (negative indices count from the end)
__device__ __forceinline__ void np_slice(int a, int b, int R, int& lo, int& hi) {
  if (a < 0) a = max(a + R, 0);
  if (b < 0) b = max(b + R, 0);
  lo = min(a, R);
  hi = min(b, R);
}
__device__ __forceinline__ void base_draw(uint32_t* row, int R, int s, int e) {   // agent.py:577-588
  int lo, hi;
  if (s < 0) { np_slice(R + s, R, R, lo, hi); if (hi > lo) set_range<true>(row, 1, lo, hi); s = 0; }
  if (e >= R) { np_slice(0, e - R, R, lo, hi); if (hi > lo) set_range<true>(row, 1, lo, hi); e = R - 1; }
  np_slice(s, e, R, lo, hi);
  if (hi > lo) set_range<true>(row, 1, lo, hi);
}

// Occlusion (Agent.exlude_V_source_data, agent.py:421-445) and fill (agent.py:569-590) of the M
// recorded objects into wf.row (which must be zeroed).  Ends with a __syncwarp.
//
// The reference sorts all objects by distance (stable) and clips every social cue f, in that order, against each
// strictly closer object o: (1) f.s' <= o.s <= f.e' -> f.e' = o.s; (2) f.s' <= o.e <= f.e' -> f.s' = o.e; (3) o covers
// [f.s', f.e'] -> (0, 0), on f's CURRENT interval with o's raw ends.  Every rule needs [o.s, o.e] to meet f's current
// interval, which only ever shrinks inside the original one (or is (0, 0) for good), so only the closer objects that
// OVERLAP f's raw interval matter -- usually none to three of M -- and only their order among themselves.  So nothing
// is sorted: the warp takes the social cues one after the other, the lanes test all objects in parallel (closer and
// overlapping?), and the few hits are applied in (distance, list order) order by repeated warp-wide minimum.
// (Before: an O(M^2) rank count for ALL objects plus a scan over the sorted prefix for every cue -- a third of the
// agent kernel's instructions.)
__device__ __forceinline__ void base_occlude_fill(WarpField& wf, int M, bool visual_exclusion, int R, int lane) {
  if (visual_exclusion && M <= 1024) {
    const int NQ = (M + 31) >> 5;                        // objects per lane: list positions lane + 32 t
    for (int c0 = 0; c0 < M; c0 += 32) {
      const int c = c0 + lane;
      unsigned cues = __ballot_sync(0xffffffffu, c < M && (wf.key[c] & (1 << 30)));   // :447-455 only social cues are drawn
      while (cues) {
        const int ci = c0 + __ffs((int)cues) - 1;
        cues &= cues - 1;
        const ObjRec f = wf.raw[ci];                                                  // (all lanes read the same entry)
        int sx = f.s, ex = f.e;
        unsigned rel = 0u;                               // bit t: object lane + 32 t is strictly closer (:430) and meets f
        for (int t = 0; t < NQ; ++t) {
          const int q = lane + 32 * t;
          if (q < M) {
            const ObjRec o = wf.raw[q];
            if ((o.d < f.d) && (o.s <= f.e) && (o.e >= f.s)) rel |= 1u << t;
          }
        }
        while (__any_sync(0xffffffffu, rel != 0u)) {     // the hits in (distance, list order) order
          // this lane's first remaining hit ... (distances are >= 0: their bit patterns order like the values)
          unsigned bhi = 0xffffffffu, blo = 0xffffffffu, bk = 0xffffffffu;
          int bt = 0, bs = 0, be = 0;
          for (unsigned m = rel; m; m &= m - 1u) {
            const int t = __ffs((int)m) - 1, q = lane + 32 * t;
            const ObjRec o = wf.raw[q];
            const unsigned long long db = (unsigned long long)__double_as_longlong(o.d);
            const unsigned hi = (unsigned)(db >> 32), lo = (unsigned)db, kq = (unsigned)(wf.key[q] & 0x3fffffff);
            if (hi < bhi || (hi == bhi && (lo < blo || (lo == blo && kq < bk)))) {
              bhi = hi; blo = lo; bk = kq; bt = t; bs = o.s; be = o.e;
            }
          }
          // ... and the warp's
          bool cand = rel != 0u;
          const unsigned mhi = __reduce_min_sync(0xffffffffu, cand ? bhi : 0xffffffffu);
          cand = cand && bhi == mhi;
          const unsigned mlo = __reduce_min_sync(0xffffffffu, cand ? blo : 0xffffffffu);
          cand = cand && blo == mlo;
          const unsigned mk = __reduce_min_sync(0xffffffffu, cand ? bk : 0xffffffffu);
          cand = cand && bk == mk;
          const int src = __ffs((int)__ballot_sync(0xffffffffu, cand)) - 1;           // exactly one lane (keys are unique)
          const int os = __shfl_sync(0xffffffffu, bs, src), oe = __shfl_sync(0xffffffffu, be, src);
          if (sx <= os && os <= ex) ex = os;                                          // :432-433
          if (sx <= oe && oe <= ex) sx = oe;                                          // :435-436
          if (os <= sx && oe >= ex) { sx = 0; ex = 0; }                               // :438-440
          if (lane == src) rel &= ~(1u << bt);
        }
        if (lane == (ci & 31)) base_draw(wf.row, R, sx, ex);
        __syncwarp();
      }
    }
  } else if (visual_exclusion) {                         // more than 1024 objects: rank everything, scan the sorted prefix
    for (int m = lane; m < M; m += 32) {   // rank by (distance, list order): the stable sort of :424
      const ObjRec f = wf.raw[m];
      const int kf = wf.key[m] & 0x3fffffff;
      int rank = 0;
      for (int q = 0; q < M; ++q) {
        const double dq = wf.raw[q].d;
        const int kq = wf.key[q] & 0x3fffffff;
        rank += (dq < f.d) || (dq == f.d && kq < kf);
      }
      wf.sorted[rank] = f;
      wf.skey[rank] = wf.key[m];
    }
    __syncwarp();
    for (int p = lane; p < M; p += 32) {
      if (!(wf.skey[p] & (1 << 30))) continue;                                      // :447-455 only social cues are drawn
      const ObjRec f = wf.sorted[p];
      int sx = f.s, ex = f.e;
      for (int q = 0; q < p; ++q) {
        const ObjRec o = wf.sorted[q];
        if (o.d < f.d) {                                                            // :430 strict
          if (sx <= o.s && o.s <= ex) ex = o.s;                                     // :432-433
          if (sx <= o.e && o.e <= ex) sx = o.e;                                     // :435-436
          if (o.s <= sx && o.e >= ex) { sx = 0; ex = 0; }                           // :438-440
        }
      }
      base_draw(wf.row, R, sx, ex);
    }
  } else {
    for (int m = lane; m < M; m += 32)
      if (wf.key[m] & (1 << 30)) base_draw(wf.row, R, wf.raw[m].s, wf.raw[m].e);
  }
  __syncwarp();
}

// number of set bits of row in bins [a, b)
__device__ __forceinline__ int popc_range(const uint32_t* row, int W, int a, int b, int lane) {
  int n = 0;
  for (int w = lane; w < W; w += 32) {
    const int lo = max(a - (w << 5), 0), hi = min(b - (w << 5), 32);
    if (hi > lo) {
      uint32_t m = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
      n += __popc(row[w] & m);
    }
  }
  return __reduce_add_sync(0xffffffffu, n);
}

// set bits of the STORED field (stored[b] = v[R-1-b], kept for b in [mask_lo, mask_hi]) in stored bins [sa, sb)
__device__ __forceinline__ int popc_stored(const uint32_t* row, int R, int W, int mask_lo, int mask_hi, int sa, int sb,
                                           int lane) {
  sa = max(sa, mask_lo);
  sb = min(sb, mask_hi + 1);
  if (sb <= sa) return 0;
  return popc_range(row, W, R - sb, R - sa, lane);
}

}  // namespace abm
