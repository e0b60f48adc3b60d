// Device functions of the visual-flocking (VF) path, shared by the fused step kernel and
// the stateless function-level kernels.  sm_100a only.
//
// Reference behaviour restated here (paths relative to the reference root):
//   vf_supcalc.projection_field          vf_supcalc.py:20-138   (interval per object)
//   vf_supcalc.calculate_closed_angle    vf_supcalc.py:142-158  -> supcalc.angle_between supcalc.py:19-34
//   supcalc.find_nearest                 supcalc.py:8-11        (first arg-min => lower index wins ties)
//   vf_supcalc.dPhi_V_of                 vf_supcalc.py:257-277
//   vf_supcalc.VSWRM_flocking_state_variables  vf_supcalc.py:161-254
//   VFAgent.update_agent_position        vf_agent.py:289-330
//   Agent.reflect_from_walls / prove_orientation   agent.py:347-394, 605-610
//   VFAgent.teleport_infinite_arena      vf_agent.py:188-204
#pragma once
#include <type_traits>

#include "abm_common.cuh"

namespace abm {

// ---------------------------------------------------------------------------------------
// shared-memory / mbarrier / bulk-copy primitives
// ---------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk TMA global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// explicit shared-space accesses with 32-bit addresses: keeps the generic->shared window arithmetic
// (S2UR SR_CgaCtaId / ULEA per access) out of the pair loop
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v));
}

// Pair-loop constants that depend only on the resolution R.  RC > 0: compile-time resolution (they become
// instruction immediates); RC == 0: read from the kernel arguments (constant bank).
template <int RC>
struct PairK {
  static constexpr double kInv = RC > 1 ? (double)(RC - 1) / ABM_TWO_PI_D : 0.0;   // (R - 1) / 2pi
  static constexpr float kTHalf = (RC % 2 == 0) ? 0.5f : 1.0f;
  __device__ __forceinline__ static float ac(const VFKernelArgs& a, int c) {
    constexpr double kA[7] = {0.9999993443489075, -0.33326515555381775, 0.19881492853164673, -0.13487225770950317,
                              0.0838717594742775, -0.037013452500104904, 0.007863515056669712};
    return RC ? (float)(kA[c] * kInv) : a.ac[c];
  }
  __device__ __forceinline__ static float half_pi_b(const VFKernelArgs& a) { return RC ? (float)(0.5 * ABM_PI_D * kInv) : a.half_pi_b; }
  __device__ __forceinline__ static float pi_b(const VFKernelArgs& a) { return RC ? (float)(ABM_PI_D * kInv) : a.pi_b; }
  __device__ __forceinline__ static float t_half(const VFKernelArgs& a) { return RC ? kTHalf : a.t_half; }
  __device__ __forceinline__ static int k_bias(const VFKernelArgs& a) {
    return RC ? ((RC % 2 == 0 ? RC / 2 - 1 : (RC - 3) / 2) - 0x4B400000 + 32) : a.k_bias;
  }
  __device__ __forceinline__ static float y_scale(const VFKernelArgs& a) { return RC ? (float)((double)RC / ABM_TWO_PI_D) : a.y_scale; }
};

// ---------------------------------------------------------------------------------------
// fp32 pair path
// ---------------------------------------------------------------------------------------

// atan(q) for q in [0, 1]: q * P(q^2), minimax in RELATIVE error (7.5e-7 in fp32 Horner).
__device__ __forceinline__ float atan_unit(float q) {
  const float z = q * q;
  float p = 0.007863515056669712f;
  p = fmaf(p, z, -0.037013452500104904f);
  p = fmaf(p, z, 0.0838717594742775f);
  p = fmaf(p, z, -0.13487225770950317f);
  p = fmaf(p, z, 0.19881492853164673f);
  p = fmaf(p, z, -0.33326515555381775f);
  p = fmaf(p, z, 0.9999993443489075f);
  return p * q;
}

// single-instruction MUFU reciprocal / reciprocal square root (no denormal fix-up code)
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// atan2(w, u) in (-pi, pi]; abs error ~1e-6 rad, absorbed by the tau_k guard band.
__device__ __forceinline__ float atan2_fast(float w, float u) {
  const float au = fabsf(u), aw = fabsf(w);
  const float mx = fmaxf(au, aw), mn = fminf(au, aw);
  float p = atan_unit(mn * rcp_approx(mx));
  if (aw > au) p = 1.57079632679489662f - p;
  if (u < 0.0f) p = 3.14159265358979324f - p;
  return copysignf(p, w);
}

// Static per-launch constants of the fp32 bin arithmetic.
//   k = ceil(ca * inv_step + t_frac) + k_off = rint(ca * inv_step + t_half) + k_off   (t_half = t_frac + 0.5)
//   h = floor(y) = rint(y - 0.5),  y = atan(r / d) * y_scale
// Both identities fail only where the argument is within the fp32 error bound of an integer
// -- exactly the pairs that are flagged for fp64 re-evaluation anyway.
struct BinConsts {
  float inv_step, t_half;
  int k_bias;               // k_off - 0x4B400000 (bit pattern of the rounding constant) + 32 (row padding)
  float y_scale;
  float thr_k;              // 0.5 - tau_k           (+inf: guard band off = fp32-only mode)
  float thr_h0, thr_h1;     // 0.5 - tau_h(y') = thr_h0 + thr_h1 * y'
  float ca_guard;           // |ca| beyond this is within the error bound of the +-pi seam
};

// Result of the fp32 evaluation of one (focal, object) pair.
struct PairFast {
  int k, h;        // centre bin IN PADDED POSITIONS (real bin + 32), half width
  bool flagged;    // must be re-evaluated in fp64: k or h within the fp32 error bound of a rounding
                   // boundary, angle on the +-pi seam, overlapping / coincident centres (q > 1, NaN)
};

// (dx, dy): centre difference object - focal (screen coords, y down); (c, ns) = (cos, -sin) of
// the focal heading; r_obj: object radius; d2 = dx^2 + dy^2 (0 -> flagged).
__device__ __forceinline__ PairFast vf_pair_fast(float dx, float dy, float d2, float r_obj,
                                                 float c, float ns, const BinConsts& bc) {
  PairFast o;
  const float MAGIC = 12582912.0f;  // 1.5 * 2^23: (x + MAGIC) - MAGIC == rint(x), low bits of x + MAGIC = rint(x)
  // rotate (dx, -dy) by -theta: closed angle = atan2(w, u)  (== -angle_between remapped, A.1)
  const float u = fmaf(dx, c, dy * ns);          //  dx cos - dy sin
  const float w = fmaf(dx, ns, -(dy * c));       // -dx sin - dy cos
  const float ca = atan2_fast(w, u);
  const float t = fmaf(ca, bc.inv_step, bc.t_half);
  const float tr = t + MAGIC;
  o.k = __float_as_int(tr) + bc.k_bias;
  bool flagged = (fabsf(t - (tr - MAGIC)) > bc.thr_k) | (fabsf(ca) > bc.ca_guard);
  const float q = r_obj * rsqrt_approx(d2);
  flagged |= !(q <= 1.0f);                                 // also catches d2 == 0 (inf / NaN)
  const float y = fmaf(atan_unit(q), bc.y_scale, -0.5f);
  const float yr = y + MAGIC;
  o.h = __float_as_int(yr) - 0x4B400000;
  flagged |= fabsf(y - (yr - MAGIC)) > fmaf(y, bc.thr_h1, bc.thr_h0);
  o.flagged = flagged;
  return o;
}

constexpr float kMagic = 12582912.0f;          // 1.5 * 2^23: low mantissa bits of x + kMagic = rint(x)
constexpr int kMagicBits = 0x4B400000;

// ---- binary angles ----------------------------------------------------------------------------------------------
// Angles live in 32-bit integers, 2^32 to the turn, so differences wrap for free.  The bearing atan2(-dy, dx) of the
// partner comes out of the octant polynomial in units of 2^-25 turn (an octant = 2^22 units: it fits the integer
// window of the 1.5 * 2^23 rounding constant), is unfolded with integer arithmetic and scaled by 128.  With
//   v = bearing - heading + half a turn          (unsigned: the closed angle measured from -pi)
// the nearest index on the linspace(-pi, pi, R) grid (vf_supcalc.py:102) is the high word of v * (R - 1) + 2^31, and
// the low word is the distance from the rounding tie: one IMAD.WIDE yields the bin and its guard band.  The opposite
// direction only differs by half a turn, which is folded into the partner's heading constant.
constexpr uint32_t kQuarter = 1u << 23, kHalf = 1u << 24;      // in units of 2^-25 turn
constexpr uint32_t kMB2 = 2u * (uint32_t)kMagicBits;

// Minimal-image difference o - f on a periodic axis (vf_supcalc.py:70-83), rounded ONCE at the magnitude of the result.
// fl(o - f) of two coordinates far apart carries a rounding error of ulp(period) / 2; after the wrap that is a relative
// error without bound for a close neighbour across the seam (found by scratch/soak_compare.py: 2 wrong bins in 1e11
// directions).  So the period goes first to whichever coordinate it makes SMALLER -- that subtraction is exact --
// and the difference is taken afterwards.
// `tie`: the wrap decision |difference| > half (taken by the reference on the exact difference) cannot be made from
// the rounded one (a neighbour exactly half an arena away: 1 wrong pair in 2.5e11) -> fp64 path.
__device__ __forceinline__ float torus_delta(float o, float f, float period, float half, bool& tie) {
  const float d0 = o - f;
  tie |= fabsf(fabsf(d0) - half) <= half * 2.4e-7f;
  float d = d0;
  if (d0 > half) d = (o - period) - f;
  if (d0 < -half) d = o - (f - period);
  return d;
}
// the same for centres = positions + radii (dr = object radius - focal radius): positions first, then radii
__device__ __forceinline__ float torus_delta_r(float o, float f, float dr, float period, float half, bool& tie) {
  const float d0 = (o - f) + dr;
  tie |= fabsf(fabsf(d0) - half) <= half * 2.4e-7f;
  float d = d0;
  if (d0 > half) d = ((o - period) - f) + dr;
  if (d0 < -half) d = (o - (f - period)) + dr;
  return d;
}

// bits of (kMagic + bearing in 2^-25 turns), bearing in (-half turn, half turn]
__device__ __forceinline__ uint32_t sym_bearing_bits(float dx, float dy, float a6) {
  constexpr double kS = 33554432.0 / ABM_TWO_PI_D;             // 2^25 / 2pi
  const float au = fabsf(dx), aw = fabsf(dy);
  const float tq = fminf(au, aw) * rcp_approx(fmaxf(au, aw));
  const float z = tq * tq;
  float p = fmaf(a6, z, (float)(-0.037013452500104904 * kS));
  p = fmaf(p, z, (float)(0.0838717594742775 * kS));
  p = fmaf(p, z, (float)(-0.13487225770950317 * kS));
  p = fmaf(p, z, (float)(0.19881492853164673 * kS));
  p = fmaf(p, z, (float)(-0.33326515555381775 * kS));
  p = fmaf(p, z, (float)(0.9999993443489075 * kS));
  uint32_t nb = __float_as_uint(fmaf(p, tq, kMagic));          // kMagicBits + rint(octant angle)
  if (aw > au) nb = (kMB2 + kQuarter) - nb;
  if (dx < 0.0f) nb = (kMB2 + kHalf) - nb;
  if (dy > 0.0f) nb = kMB2 - nb;                               // atan2(-dy, dx): screen y points down
  return nb;
}
constexpr float kBearingA6 = (float)(0.007863515056669712 * (33554432.0 / ABM_TWO_PI_D));

// Heading constant of an agent: v = 128 * bearing_bits - heading_const (mod 2^32).
__device__ __forceinline__ uint32_t sym_heading_const(float theta) {
  double turns = (double)theta * (1.0 / ABM_TWO_PI_D);
  turns -= floor(turns);
  const uint32_t th_bam = (uint32_t)(unsigned long long)rint(turns * 4294967296.0);
  return th_bam + 128u * (uint32_t)kMagicBits - 0x80000000u;
}

// One direction on the fast path: nearest linspace index of the closed angle (returned as the padded start position of
// the interval, h folded into `bh`); `slow` is set when the angle is within the fp32 error bound of a rounding tie or
// of the +-pi seam.
// FAST (the pair loops): ONE test covers rounding ties and the seam.  The seam (v = 0) is the centre of bin 0 = bin
// R - 1, i.e. low word = 2^31, so "twice the low word is within 2T of 0 (mod 2^32)" flags every tie AND every bin
// centre; the error bound of the bin coordinate is the same everywhere on the ring, so the tie band T serves both.
// Bin centres other than the seam's are false positives (as rare as ties); the slow path, which uses the separate
// tests, just draws them.  `half64` is the rounding constant 2^31 as a 64-bit register pair the caller keeps alive
// (IMAD.WIDE takes it as its addend).
template <int RC, bool FAST = false>
__device__ __forceinline__ int sym_side_k(const VFKernelArgs& a, uint32_t nb, uint32_t hconst, int bh, bool& slow,
                                          unsigned long long half64 = 0x80000000ull) {
  const uint32_t v = 128u * nb - hconst;
  const uint32_t Rp = RC ? (uint32_t)(RC - 1) : (uint32_t)(a.R - 1);
  const unsigned long long prod = (unsigned long long)v * Rp + half64;
  if (FAST) {
    const uint32_t lo = (uint32_t)prod;
    slow |= (lo + lo + 2u * a.sym_tie32) < 4u * a.sym_tie32;
  } else {
    slow |= ((uint32_t)prod + a.sym_tie32) < 2u * a.sym_tie32;
    slow |= (v + a.sym_seam32) < 2u * a.sym_seam32;
  }
  return (int)(uint32_t)(prod >> 32) + bh;
}

// ---------------------------------------------------------------------------------------
// fp64 exact pair path: the reference's own operation sequence, individually rounded
// ---------------------------------------------------------------------------------------

struct FocalExact {
  float px, py;      // top-left position (exact-coincidence test, vf_supcalc.py:57)
  double cix, ciy;   // centre (vf_supcalc.py:42)
  double u1x, u1y;   // unit heading vector v1 / |v1| (vf_supcalc.py:45-49, supcalc.py:14)
};

__device__ __forceinline__ FocalExact vf_focal_exact(float px, float py, float r, float theta) {
  FocalExact f;
  f.px = px; f.py = py;
  const double x = px, y = py, rr = r, th = theta;
  f.cix = __dadd_rn(x, rr);
  f.ciy = __dadd_rn(y, rr);
  const double ex = __dadd_rn(x, __dmul_rn(__dadd_rn(1.0, cos(th)), rr));
  const double ey = __dadd_rn(y, __dmul_rn(__dadd_rn(1.0, -sin(th)), rr));
  const double v1x = __dadd_rn(ex, -f.cix), v1y = __dadd_rn(ey, -f.ciy);
  const double n1 = __dsqrt_rn(__dadd_rn(__dmul_rn(v1x, v1x), __dmul_rn(v1y, v1y)));
  f.u1x = __ddiv_rn(v1x, n1);
  f.u1y = __ddiv_rn(v1y, n1);
  return f;
}

// first-argmin over the numpy linspace(-pi, pi, R) grid, evaluated on the three candidates
// around the closed-form estimate with the grid values numpy produces
// (y[k] = fl(fl(k * step) + (-pi)), y[R-1] = pi).
__device__ __forceinline__ int nearest_bin_exact(double value, int R, double lin_step) {
  int k0 = (int)ceil((value + ABM_PI_D) / lin_step - 0.5);
  k0 = max(0, min(R - 1, k0));
  int best = k0;
  double bestd = 1e300;
  for (int kk = max(0, k0 - 1); kk <= min(R - 1, k0 + 1); ++kk) {
    const double phi = (kk == R - 1) ? ABM_PI_D : __dadd_rn(__dmul_rn((double)kk, lin_step), -ABM_PI_D);
    const double d = fabs(__dadd_rn(phi, -value));
    if (d < bestd) { bestd = d; best = kk; }
  }
  return best;
}

struct PairExact { int k, h; bool valid; };

// Object at top-left (ox, oy) with radius orad seen by the focal agent `f`.
// boundary / width / height: torus re-centring of vf_supcalc.py:70-83.
static __device__ __noinline__ PairExact vf_pair_exact(const FocalExact& f, float ox, float oy, float orad,
                                                int boundary, double width, double height,
                                                int R, double lin_step) {
  PairExact o;
  const bool same_pos = (ox == f.px) && (oy == f.py);                       // :57
  const double rr = orad;
  double cjx = __dadd_rn((double)ox, rr), cjy = __dadd_rn((double)oy, rr);   // :61-64
  double v2x = __dadd_rn(cjx, -f.cix), v2y = __dadd_rn(cjy, -f.ciy);
  if (boundary == 1) {                                                      // :70-83
    if (fabs(v2x) > width * 0.5) {
      if (f.cix < cjx) cjx = __dadd_rn(cjx, -width); else if (f.cix > cjx) cjx = __dadd_rn(cjx, width);
    }
    if (fabs(v2y) > height * 0.5) {
      if (f.ciy < cjy) cjy = __dadd_rn(cjy, -height); else if (f.ciy > cjy) cjy = __dadd_rn(cjy, height);
    }
    v2x = __dadd_rn(cjx, -f.cix); v2y = __dadd_rn(cjy, -f.ciy);
  }
  const double n2 = __dsqrt_rn(__dadd_rn(__dmul_rn(v2x, v2x), __dmul_rn(v2y, v2y)));
  const double u2x = __ddiv_rn(v2x, n2), u2y = __ddiv_rn(v2y, n2);
  double dot = __dadd_rn(__dmul_rn(f.u1x, u2x), __dmul_rn(f.u1y, u2y));
  dot = fmin(1.0, fmax(-1.0, dot));
  double ang = acos(dot);                                                   // supcalc.py:31
  if (__dadd_rn(__dmul_rn(f.u1x, u2y), -__dmul_rn(f.u1y, u2x)) < 0.0) ang = -ang;
  if (ang < 0.0) ang = __dadd_rn(ang, ABM_TWO_PI_D);                        // % 2pi (vf_supcalc.py:150)
  const double ca = (ang >= 0.0 && ang <= ABM_PI_D) ? -ang : __dadd_rn(ABM_TWO_PI_D, -ang);   // :154-157
  const double dist = n2;                                                   // :88
  const double vis = __dmul_rn(2.0, atan(__ddiv_rn(rr, dist)));             // :99
  const double proj = __dmul_rn(__ddiv_rn(vis, ABM_TWO_PI_D), (double)R);   // :114
  o.k = nearest_bin_exact(ca, R, lin_step);                                 // :102
  o.h = (int)floor(__ddiv_rn(proj, 2.0));                                   // :116
  o.valid = (n2 > 0.0) && !same_pos;
  return o;
}

// ---------------------------------------------------------------------------------------
// drawing an interval into a bit-packed UN-flipped row (vf_supcalc.py:119-129)
// ---------------------------------------------------------------------------------------

// Row word w of the calling thread lives at f[w * stride].
template <bool ATOMIC>
__device__ __forceinline__ void or_word(uint32_t* p, uint32_t m) {
  if (ATOMIC) atomicOr(p, m); else *p |= m;
}

// set bins [a, b), 0 <= a < b <= R
template <bool ATOMIC>
__device__ __forceinline__ void set_range(uint32_t* f, int stride, int a, int b) {
  const int w0 = a >> 5, w1 = (b - 1) >> 5;
  const uint32_t m0 = 0xffffffffu << (a & 31);
  const uint32_t m1 = 0xffffffffu >> ((32 - b) & 31);
  if (w0 == w1) {
    or_word<ATOMIC>(f + w0 * stride, m0 & m1);
  } else {
    or_word<ATOMIC>(f + w0 * stride, m0);
    for (int w = w0 + 1; w < w1; ++w) or_word<ATOMIC>(f + w * stride, 0xffffffffu);
    or_word<ATOMIC>(f + w1 * stride, m1);
  }
}

// Full drawing rule incl. the visibility test on pixel indices and the VF wrap quirks.
template <bool ATOMIC>
__device__ __forceinline__ void vf_draw(uint32_t* f, int stride, int R, int fov0, int fov1, int k, int h) {
  int ps = k - h, pe = k + h;
  const bool vis = (fov0 < ps && ps < fov1) || (fov0 < pe && pe < fov1);   // :119
  if (!vis || h <= 0) return;
  if (ps < 0) {                                                             // :122-124
    set_range<ATOMIC>(f, stride, max(R + ps, 0), R);
    ps = 0;
  }
  if (pe >= R) {                                                            // :125-127
    const int e = min(pe - (R - 1), R);
    if (e > 0) set_range<ATOMIC>(f, stride, 0, e);
    pe = R;
  }
  if (pe > ps) set_range<ATOMIC>(f, stride, ps, pe);                        // :129
}

// The same drawing rule on a row addressed in the SHARED state space (32-bit address of real word 0,
// byte stride between the words of a row), with fire-and-forget reductions: for code that is not
// inlined into the kernel (a generic pointer would turn every atomic into a generic-space ATOM).
__device__ __forceinline__ void red_or_shared(uint32_t addr, uint32_t m) {
  asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(m) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_shared(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void set_range_shared(uint32_t row, uint32_t stride_b, int a, int b) {
  const int w0 = a >> 5, w1 = (b - 1) >> 5;
  const uint32_t m0 = 0xffffffffu << (a & 31);
  const uint32_t m1 = 0xffffffffu >> ((32 - b) & 31);
  if (w0 == w1) {
    red_or_shared(row + (uint32_t)w0 * stride_b, m0 & m1);
  } else {
    red_or_shared(row + (uint32_t)w0 * stride_b, m0);
    for (int w = w0 + 1; w < w1; ++w) red_or_shared(row + (uint32_t)w * stride_b, 0xffffffffu);
    red_or_shared(row + (uint32_t)w1 * stride_b, m1);
  }
}
__device__ __forceinline__ void vf_draw_shared(uint32_t row, uint32_t stride_b, int R, int fov0, int fov1, int k, int h) {
  int ps = k - h, pe = k + h;
  const bool vis = (fov0 < ps && ps < fov1) || (fov0 < pe && pe < fov1);   // :119
  if (!vis || h <= 0) return;
  if (ps < 0) {                                                             // :122-124
    set_range_shared(row, stride_b, max(R + ps, 0), R);
    ps = 0;
  }
  if (pe >= R) {                                                            // :125-127
    const int e = min(pe - (R - 1), R);
    if (e > 0) set_range_shared(row, stride_b, 0, e);
    pe = R;
  }
  if (pe > ps) set_range_shared(row, stride_b, ps, pe);                     // :129
}

// Hot-path drawing into a PADDED private row: word 0 of the padded row holds the virtual bins
// [-32, 0), words 1.. the real bins, and one more word follows the last real word, so an
// interval that leaves [0, R) by at most 32 bins is drawn without any wrap logic; the padding
// is folded back onto the ring once per step (vf_fold_padding).  `f` points at padded word 0.
//   VF wrap rule (vf_supcalc.py:122-127): bins below 0 wrap to R + bin; an interval reaching
//   pe >= R additionally sets bin pe - R (the slice end is pe - (R - 1)), i.e. it is drawn as
//   [ps, pe + 1) on the ring.
// An interval of at most 32 bins touches at most two words: one funnel shift each.
// Interval [ps, ps + 2h) in PADDED bit positions (real bin + 32), 1 <= h <= 16.  The extra
// wrap bin of the reference is added once per step in vf_fold_padding, not per interval.
__device__ __forceinline__ void vf_draw_short(unsigned char* row_bytes, int stride_bytes, int ps_padded, int h) {
  const uint32_t t = 0xffffffffu >> (32 - 2 * h);
  uint32_t* p = reinterpret_cast<uint32_t*>(row_bytes + (ps_padded >> 5) * stride_bytes);
  const uint32_t lo = __funnelshift_l(0u, t, ps_padded);   // t << (ps & 31)
  const uint32_t hi = __funnelshift_l(t, 0u, ps_padded);   // bits shifted out into the next word
  *p |= lo;
  if (hi) *reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(p) + stride_bytes) |= hi;
}

// Wide interval (h > 16) that still fits in the padded row: first / middle / last word.
// [ps, pe) in padded positions, 0 <= ps, pe <= R + 62.
__device__ __forceinline__ void vf_draw_wide(unsigned char* row_bytes, int stride_bytes, int ps_padded, int pe_padded) {
  const int w0 = ps_padded >> 5, w1 = (pe_padded - 1) >> 5;   // w1 > w0 because the interval has > 32 bins
  unsigned char* p = row_bytes + w0 * stride_bytes;
  *reinterpret_cast<uint32_t*>(p) |= __funnelshift_l(0u, 0xffffffffu, ps_padded);
#pragma unroll 1
  for (int w = w0 + 1; w < w1; ++w) {
    p += stride_bytes;
    *reinterpret_cast<uint32_t*>(p) = 0xffffffffu;
  }
  p += stride_bytes;
  *reinterpret_cast<uint32_t*>(p) |= 0xffffffffu >> ((32 - pe_padded) & 31);
}

// Fold the padding words back onto the ring and clear them.  `f` points at padded word 0.
// Intervals drawn by vf_draw_short that reach pe >= R cover bin R-1 and spill into [R, R+32);
// the reference gives each of them the slice [0, pe - R + 1), so their union is the spilled
// bits shifted up by one bin with bin 0 added (vf_supcalc.py:125-127).  (An interval drawn by
// the general rule with ps < 0 also covers bin R-1, but it always covers bin 0 as well, so
// taking it for a reaching interval is harmless.)
__device__ __forceinline__ void vf_fold_padding(uint32_t* f, int stride, int R, int W) {
  const bool reach = (f[(((R - 1) >> 5) + 1) * stride] >> ((R - 1) & 31)) & 1u;
  // overflow: ring bins [R, R + 32) (padded positions [R + 32, R + 64)) -> bins [0, 32) = padded word 1
  {
    const int pos = R + 32, w = pos >> 5, sh = pos & 31;
    const uint32_t lo = f[w * stride];
    const uint32_t hi = (w + 1 < W + 2) ? f[(w + 1) * stride] : 0u;
    const uint32_t ov = sh ? __funnelshift_r(lo, hi, sh) : lo;
    if (reach) f[stride] |= (ov << 1) | 1u;
    // clear everything at or beyond ring bin R
    f[w * stride] = sh ? (lo & ((1u << sh) - 1u)) : 0u;
    if (w + 1 < W + 2) f[(w + 1) * stride] = 0u;
  }
  // underflow: virtual bins [-32, 0) -> ring bins [R - 32, R) = padded positions [R, R + 32)
  const uint32_t un = f[0];
  if (un) {
    const int sh = R & 31, w = R >> 5;
    f[w * stride] |= un << sh;
    if (sh) f[(w + 1) * stride] |= un >> (32 - sh);
    f[0] = 0u;
  }
}

// ---------------------------------------------------------------------------------------
// epilogue: dV/dphi, flocking integrals, kinematics (fp64; tiny next to the pair loop)
// ---------------------------------------------------------------------------------------

struct FlockTerms { double dvel, dpsi, a_blob, a_edge, b_blob, b_edge; };

// V (un-flipped row) word w at f[w * stride].  Closed form of vf_supcalc.py:199-254:
//   blob = trapz(cos(Phi) * (-V), Phi) = dphi * (-sum_k cos_k V_k + (cos_0 V_0 + cos_{R-1} V_{R-1}) / 2)
//   edge = sum_k cos_k * dPhi_V_of(V)_k^2, attributed to bin k-1 (forward diff) or k (backward diff,
//          iff V[0] - V[R-1] > 0; vf_supcalc.py:268-272) of each ring boundary between k-1 and k.
// sum_k cos_k V_k over a run [a, b) = pc[b] - pc[a] (prefix table).
__device__ __forceinline__ FlockTerms vf_flock_terms(const uint32_t* f, int stride, int R, int W,
                                                     const PhiLut* __restrict__ lut, double dphi,
                                                     double vel, const VFParams6& prm,
                                                     double A0, double B0, double V0) {
  const uint32_t last_valid = (R & 31) ? ((1u << (R & 31)) - 1u) : 0xffffffffu;
  const uint32_t w_last = f[(W - 1) * stride] & last_valid;
  const uint32_t v_first = f[0] & 1u;
  const uint32_t v_last = (w_last >> ((R - 1) & 31)) & 1u;
  const bool backward = (v_first == 1u) && (v_last == 0u);
  double Sc = 0.0, Ss = 0.0, Ec = 0.0, Es = 0.0;
  uint32_t carry = v_last;   // ring predecessor of bin 0
  for (int w = 0; w < W; ++w) {
    uint32_t cur = f[w * stride];
    if (w == W - 1) cur &= last_valid;
    uint32_t diff = cur ^ ((cur << 1) | carry);   // bit b: V[k] != V[k-1], k = 32w + b
    if (w == W - 1) diff &= last_valid;
    carry = (w == W - 1) ? 0u : (cur >> 31);
    while (diff) {
      const int b = __ffs(diff) - 1;
      diff &= diff - 1;
      const int k = (w << 5) + b;
      const int ke = backward ? k : (k == 0 ? R - 1 : k - 1);
      const double2 cs = *reinterpret_cast<const double2*>(&lut[ke].c);
      Ec += cs.x; Es += cs.y;
      const double2 pp = *reinterpret_cast<const double2*>(&lut[k].pc);
      if ((cur >> b) & 1u) { Sc -= pp.x; Ss -= pp.y; }   // run starts at k
      else                 { Sc += pp.x; Ss += pp.y; }   // run ended at k
    }
  }
  if (v_last) { Sc += lut[R].pc; Ss += lut[R].ps; }      // run reaching the end of the row
  const double c0 = lut[0].c, s0 = lut[0].s, cl = lut[R - 1].c, sl = lut[R - 1].s;
  const double endc = 0.5 * (c0 * (double)v_first + cl * (double)v_last);
  const double ends = 0.5 * (s0 * (double)v_first + sl * (double)v_last);
  FlockTerms t;
  t.a_blob = A0 * (dphi * (endc - Sc));
  t.a_edge = A0 * prm.alp1 * Ec;
  t.b_blob = B0 * (dphi * (ends - Ss));
  t.b_edge = B0 * prm.bet1 * Es;
  t.dvel = prm.gam * (V0 - vel) + t.a_blob + t.a_edge;
  t.dpsi = t.b_blob + t.b_edge;
  return t;
}

// The same terms for the step kernels, from the ring edges alone.  With E_k = exp(i Phi_k) and Phi_k = Phi_0 + k d:
//   sum_{m<k} E_m = kappa (E_k - E_0),  kappa = 1 / (exp(i d) - 1)            (geometric series)
// so the blob sums over all runs collapse to kappa (Z_fall - Z_rise) (+ the prefix at R if the last bin is set,
// because trapz works on the linear array), and the edge sums are Z_rise + Z_fall, rotated by exp(-i d) under the
// forward-difference rule (edge attributed to bin k - 1).  One 16-byte table read and four fp64 operations per edge.
// etab_s: shared-space address of a copy of (cos Phi_k, sin Phi_k), k < R, or 0 (read the global table).
__device__ __forceinline__ double2 lds_d2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
struct EdgeSums {
  double zsr, zsi, zdr, zdi;   // Z_rise + Z_fall, Z_fall - Z_rise (real, imaginary)
  uint32_t v_first, v_last;
};

// Edge sums of K rows per thread, interleaved in ONE flat loop for the whole warp (all 32 lanes call this, `active` or
// not): in every iteration a lane first moves each of its rows on to the next word if the current one has no edge
// left, then takes ONE edge per row; the warp reconverges between the two halves, so each half always runs with as
// many lanes as have work.  (A loop over the edges of a word inside a loop over the words makes the warp wait, in
// every word, for the lane with the most edges there; K > 1 gives the latency-bound loop independent work.)
// edges taken per row and loop iteration (the vote, the two reconvergence points and the word bookkeeping are paid once)
constexpr int kEdgesPerIteration = 2;   // (3: 1.906 ms, 4: 1.925 ms against 1.891; two word advances per iteration: slower)
template <int K>
__device__ __forceinline__ void vf_edge_sums(const uint32_t* const (&f)[K], int stride, const VFKernelArgs& a,
                                             uint32_t etab_s, const bool (&active)[K], EdgeSums (&out)[K]) {
  const int R = a.R, W = a.W;
  const PhiLut* __restrict__ lut = a.lut;
  const uint32_t last_valid = (R & 31) ? ((1u << (R & 31)) - 1u) : 0xffffffffu;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const uint32_t w_last = active[j] ? f[j][(W - 1) * stride] & last_valid : 0u;
    out[j].v_first = active[j] ? f[j][0] & 1u : 0u;
    out[j].v_last = (w_last >> ((R - 1) & 31)) & 1u;
    out[j].zsr = out[j].zsi = out[j].zdr = out[j].zdi = 0.0;
  }
  // (vf_fold_padding has cleared the bits at and beyond bin R of the last word; only the difference needs the mask.)
  auto edge_loop = [&](auto smem_tab) {
    int w[K];
    uint32_t cur[K], diff[K], carry[K];
    bool done[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { w[j] = -1; cur[j] = diff[j] = 0u; carry[j] = out[j].v_last; done[j] = !active[j]; }
    for (;;) {
      bool all_done = true;
#pragma unroll
      for (int j = 0; j < K; ++j) all_done &= done[j];
      if (!__any_sync(0xffffffffu, !all_done)) break;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (!done[j] && diff[j] == 0u) {
          if (++w[j] >= W) {
            done[j] = true;
          } else {
            cur[j] = f[j][w[j] * stride];
            diff[j] = cur[j] ^ ((cur[j] << 1) | carry[j]);        // bit b: V[k] != V[k-1], k = 32w + b
            if (w[j] == W - 1) diff[j] &= last_valid;
            carry[j] = cur[j] >> 31;                              // ring predecessor of the next word's bin 0
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int rep = 0; rep < kEdgesPerIteration; ++rep)
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (diff[j]) {
          const int b = __ffs(diff[j]) - 1;
          diff[j] &= diff[j] - 1;
          const int k = (w[j] << 5) + b;
          double2 e;
          if (decltype(smem_tab)::value) e = lds_d2(etab_s + 16u * (uint32_t)k);
          else e = *reinterpret_cast<const double2*>(&lut[k].c);
          // rising edge (run starts at k, V[k] = 1) counts into Z_rise: negative in Z_fall - Z_rise
          const int sgn = (int)((cur[j] >> b) << 31);
          out[j].zsr += e.x; out[j].zsi += e.y;
          out[j].zdr += __hiloint2double(__double2hiint(e.x) ^ sgn, __double2loint(e.x));
          out[j].zdi += __hiloint2double(__double2hiint(e.y) ^ sgn, __double2loint(e.y));
        }
      }
      __syncwarp();
    }
  };
  if (etab_s) edge_loop(std::true_type{}); else edge_loop(std::false_type{});
}

__device__ __forceinline__ FlockTerms vf_terms_from_edges(const EdgeSums& z, const VFKernelArgs& a, double vel,
                                                          const VFParams6& prm, double A0, double B0, double V0) {
  const int R = a.R;
  const PhiLut* __restrict__ lut = a.lut;
  const bool backward = (z.v_first == 1u) && (z.v_last == 0u);
  double Sc = a.kappa_r * z.zdr - a.kappa_i * z.zdi, Ss = a.kappa_r * z.zdi + a.kappa_i * z.zdr;
  if (z.v_last) { Sc += lut[R].pc; Ss += lut[R].ps; }      // run reaching the end of the row
  const double rc = backward ? 1.0 : a.rot_c, rs = backward ? 0.0 : -a.rot_s;   // exp(-i d) under the forward rule
  const double Ec = rc * z.zsr - rs * z.zsi, Es = rc * z.zsi + rs * z.zsr;
  const double c0 = lut[0].c, s0 = lut[0].s, cl = lut[R - 1].c, sl = lut[R - 1].s;
  const double endc = 0.5 * (c0 * (double)z.v_first + cl * (double)z.v_last);
  const double ends = 0.5 * (s0 * (double)z.v_first + sl * (double)z.v_last);
  FlockTerms t;
  t.a_blob = A0 * (a.dphi * (endc - Sc));
  t.a_edge = A0 * prm.alp1 * Ec;
  t.b_blob = B0 * (a.dphi * (ends - Ss));
  t.b_edge = B0 * prm.bet1 * Es;
  t.dvel = prm.gam * (V0 - vel) + t.a_blob + t.a_edge;
  t.dpsi = t.b_blob + t.b_edge;
  return t;
}

__device__ __forceinline__ double wrap_heading_once(double th) {   // agent.py:605-610
  if (th < 0.0) th = ABM_TWO_PI_D + th;
  if (th > ABM_TWO_PI_D) th = th - ABM_TWO_PI_D;
  return th;
}

__device__ __forceinline__ double limit_abs(double v, double lim) {   // vf_agent.py:311-330
  const double sg = (v < 0.0) ? -1.0 : 1.0;
  return (fabs(v) > lim) ? lim * sg : v;
}

// Agent.reflect_from_walls (agent.py:347-394): tests use the centre BEFORE any fix.  `turn`: pi/2 (the mirror-like rule
// of the base class); a VFAgent that has lines to follow turns by pi instead (vf_agent.py:80-129).
__device__ __forceinline__ void reflect_from_walls(double& x, double& y, double& th, double r,
                                                   double width, double height, double pad, double turn = ABM_PI_D / 2.0) {
  const double bx0 = pad, bx1 = pad + width, by0 = pad, by1 = pad + height;
  const double cx = x + r, cy = y + r;
  const double PI = ABM_PI_D, H = ABM_PI_D / 2.0, T3 = 3.0 * ABM_PI_D / 2.0, TWO = 2.0 * ABM_PI_D;
  if (cx < bx0) {
    x = bx0 - r;
    if (H <= th && th < PI) th -= turn; else if (PI <= th && th <= T3) th += turn;
    th = wrap_heading_once(th);
  }
  if (cx > bx1) {
    x = bx1 - r - 1.0;
    if (T3 <= th && th < TWO) th -= turn; else if (0.0 <= th && th <= H) th += turn;
    th = wrap_heading_once(th);
  }
  if (cy < by0) {
    y = by0 - r;
    if (H <= th && th <= PI) th += turn; else if (0.0 <= th && th < H) th -= turn;
    th = wrap_heading_once(th);
  }
  if (cy > by1) {
    y = by1 - r - 1.0;
    if (T3 <= th && th <= TWO) th += turn; else if (PI <= th && th < T3) th -= turn;
    th = wrap_heading_once(th);
  }
}

// VFAgent.teleport_infinite_arena (vf_agent.py:188-204)
__device__ __forceinline__ void teleport_torus(double& x, double& y, double r, double width, double height,
                                               double pad) {
  const double bx0 = pad, bx1 = pad + width, by0 = pad, by1 = pad + height;
  const double cx = x + r, cy = y + r;
  if (cx < bx0) x = bx1 - r; else if (cx > bx1) x = bx0 + r;
  if (cy < by0) y = by1 - r; else if (cy > by1) y = by0 + r;
}

// stored word ws of the flipped field: stored[s] = V[R - 1 - s]
__device__ __forceinline__ uint32_t flipped_word(const uint32_t* f, int stride, int R, int W, int ws) {
  // bits s = 32ws .. 32ws+31  <-  V[hi .. hi-31], hi = R - 1 - 32ws
  const int hi = R - 1 - (ws << 5);
  const int wh = hi >> 5;            // word holding V[hi]   (hi >= 0 for ws < W)
  const int sh = 31 - (hi & 31);     // left shift that brings V[hi] to bit 31
  const uint32_t a = f[wh * stride];
  const uint32_t lo = (wh > 0) ? f[(wh - 1) * stride] : 0u;
  const uint32_t v = sh ? ((a << sh) | (lo >> (32 - sh))) : a;   // bit 31 = V[hi], bit 0 = V[hi-31]
  uint32_t out = __brev(v);                                       // bit 0 = V[hi]
  const int nvalid = min(32, hi + 1);
  if (nvalid < 32) out &= (1u << nvalid) - 1u;
  return out;
}

// vf_supcalc.follow_lines_local (vf_supcalc.py:293-328): two sensors ahead-left / ahead-right of the agent read the mean of
// the line map over a square window (numpy basic slices of int()-truncated bounds: negative bounds wrap, out-of-range
// ones clip, an empty window gives nan and a heading change of 0); the agent steers towards the brighter sensor.
// map: [d0][d1] as VFAgent.line_map (first axis x).  fp64 with the reference's operation order; the window sums run in
// row order (numpy's pairwise order is not reproduced: 1e-16 relative).
__device__ __forceinline__ void np_slice_bounds(int a, int b, int n, int& lo, int& hi) {
  if (a < 0) a = max(a + n, 0);
  if (b < 0) b = max(b + n, 0);
  lo = min(a, n); hi = min(b, n);
}
__device__ __forceinline__ double vf_line_window_mean(const float* map, int d0, int d1, double s0, double s1, double sr) {
  int a, b, c, d;
  np_slice_bounds((int)(s1 - sr), (int)(s1 + sr), d0, a, b);      // :310 first axis: sensor_pos[1]
  np_slice_bounds((int)(s0 - sr), (int)(s0 + sr), d1, c, d);      //      second axis: sensor_pos[0]
  if (b <= a || d <= c) return __longlong_as_double(0x7ff8000000000000ll);   // empty slice: nan
  double sum = 0.0;
  int cnt = 0;
  for (int i = a; i < b; ++i) {
    const float* row = map + (size_t)i * d1;
    for (int j = c; j < d; ++j) {
      const float v = __ldg(row + j);
      if (v == v) { sum += (double)v; ++cnt; }                    // nanmean
    }
  }
  return cnt ? sum / (double)cnt : __longlong_as_double(0x7ff8000000000000ll);
}
__device__ __forceinline__ double vf_follow_lines(const VFKernelArgs& a, double x, double y, double r, double th, double vel) {
  const double a34 = 3.0 * ABM_PI_D / 4.0, sd = a.lm_sd;
  const double base0 = (y + r) - sd, base1 = (x + r) - sd;       // :295-302
  const double s1_0 = base0 + (1.0 + sin(th + a34)) * sd, s1_1 = base1 + (1.0 - cos(th + a34)) * sd;
  const double s2_0 = base0 + (1.0 + sin(th - a34)) * sd, s2_1 = base1 + (1.0 - cos(th - a34)) * sd;
  const double m1 = vf_line_window_mean(a.line_map, a.lm_d0, a.lm_d1, s1_0, s1_1, a.lm_sr);
  const double m2 = vf_line_window_mean(a.line_map, a.lm_d0, a.lm_d1, s2_0, s2_1, a.lm_sr);
  if (m1 != m1 || m2 != m2) return 0.0;                          // :314-315
  const double ori_change = (vel != 0.0) ? 0.5 * (m2 - m1) : 0.0;   // np.sign(agvel) truthy (:317-320)
  if (m1 != m2) return ori_change;                               // :321-324
  return (m1 != 0.0) ? 0.01 : 0.0;                               // :326-329
}

// Epilogue shared by the step kernels, K agents per thread: fold the row padding, edges + flocking integrals, heading /
// speed / position update, walls or torus, outputs.
// b: replicate, i: agent index in the replicate, li: index in this engine's tile, padrow: padded word 0 of the agent's
// row (stride in words), me: (x, y, radius, cull^2).
// Called by ALL lanes of the warp (the edge loop reconverges the warp with full-mask votes); lanes without an agent
// pass active = false and touch no memory.
template <bool TORUS, int K>
__device__ __forceinline__ void vf_agent_epilogue(const VFKernelArgs& a, int b, const int (&i)[K], const int (&li)[K],
                                                  uint32_t* const (&padrow)[K], int stride, const float4 (&me)[K],
                                                  const float (&th)[K], uint32_t etab_s, const bool (&active)[K]) {
  const VFParams6 prm = *reinterpret_cast<const VFParams6*>(a.params + (size_t)b * a.param_stride);
  const uint32_t* rows[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (active[j]) vf_fold_padding(padrow[j], stride, a.R, a.W);
    rows[j] = padrow[j] + stride;      // real word 0
  }
  EdgeSums z[K];
  if (a.phi_ok) vf_edge_sums<K>(rows, stride, a, etab_s, active, z);
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (!active[j]) continue;
    const size_t gi = (size_t)b * a.N + i[j];
    double A0 = prm.alp0, B0 = prm.bet0, V0 = prm.v0;            // vf_supcalc.py:191-196
    if (a.ov_alp0) { const float v = a.ov_alp0[gi]; if (v == v) A0 = v; }
    if (a.ov_bet0) { const float v = a.ov_bet0[gi]; if (v == v) B0 = v; }
    if (a.ov_v0)   { const float v = a.ov_v0[gi];   if (v == v) V0 = v; }
    const double vel0 = a.vel[gi];
    FlockTerms ft;
    if (a.phi_ok) {
      ft = vf_terms_from_edges(z[j], a, vel0, prm, A0, B0, V0);
    } else {   // len(PHI) != len(soc_v_field): the reference skips the calculation (vf_agent.py:282-284)
      ft.dvel = ft.dpsi = ft.a_blob = ft.a_edge = ft.b_blob = ft.b_edge = 0.0;
    }
    double dpsi = ft.dpsi, dvel = ft.dvel;
    if (a.line_map)                                               // lines to follow replace the heading change (:273-276)
      dpsi = vf_follow_lines(a, (double)me[j].x, (double)me[j].y, (double)me[j].z, (double)th[j], vel0);
    if (a.limit_movement) dpsi = limit_abs(dpsi, a.max_th);       // vf_agent.py:293-294
    double nth = wrap_heading_once((double)th[j] + dpsi);         // :295-296
    double nv = vel0 + dvel;                                      // :298
    if (a.limit_movement) nv = limit_abs(nv, a.max_vel);          // :299-300
    double sn, cn;
    sincos(nth, &sn, &cn);
    double nx = (double)me[j].x + nv * cn;                        // :303-306
    double ny = (double)me[j].y - nv * sn;
    if (!TORUS) reflect_from_walls(nx, ny, nth, (double)me[j].z, a.width_d, a.height_d, a.pad_d,
                                   a.line_map ? ABM_PI_D : ABM_PI_D / 2.0);       // vf_agent.py:80-129
    else teleport_torus(nx, ny, (double)me[j].z, a.width_d, a.height_d, a.pad_d);

    const float4 rec_new = make_float4((float)nx, (float)ny, me[j].z, me[j].w);
    a.rec_out[gi] = rec_new;
    for (int p = 0; p < a.n_peers; ++p) a.peer_rec_out[p][gi] = rec_new;   // NVLink peer stores (fused tile exchange)
    a.theta[gi] = (float)nth;
    a.vel[gi] = (float)nv;

    // outputs in API order when the engine keeps an internal (spatially sorted) order
    const size_t oi = a.perm ? (size_t)b * a.N + a.perm[gi] : (size_t)b * a.tile_count + li[j];
    if (a.terms_out) {
      double* t = a.terms_out + oi * 6;
      t[0] = ft.dvel; t[1] = ft.dpsi; t[2] = ft.a_blob; t[3] = ft.a_edge; t[4] = ft.b_blob; t[5] = ft.b_edge;
    }
    if (a.fields_out) {
      uint32_t* out = a.fields_out + oi * a.W;
      for (int ws = 0; ws < a.W; ++ws) out[ws] = flipped_word(rows[j], stride, a.R, a.W, ws);
    }
  }
}

}  // namespace abm
