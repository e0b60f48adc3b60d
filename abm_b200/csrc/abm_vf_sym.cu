// Symmetric visual-flocking step kernel for sm_100a: every UNORDERED pair {i, j} of a replicate
// is evaluated once and drawn into BOTH agents' rows.
//
// What the two directions share (equal radii): the centre distance, hence rsqrt, atan(r/d) and
// the half width h, and the absolute bearing phi = atan2(-dy, dx); the closed angles are then
//   ca_i = wrap(phi - theta_i)          and          ca_j = wrap(phi + pi - theta_j)
// (vf_supcalc.py:86-117 evaluated from both ends), so both polynomial arctangents and both MUFU
// operations are paid once per unordered pair; what is left per direction is one subtraction, the
// rounding to a bin with its guard band, and the read-modify-write of one or two row words.
//
// Mapping.  CTA = one replicate, 16 warps at N = 1024.  ALL rows of the replicate live in the CTA's shared memory
// ([word][agent]: bank == agent mod 32; 164 B per agent at R = 1200 -> 1024 agents in 164 KB).  The 32-agent blocks
// are paired like the rounds of a round-robin tournament (circle method): P (P - 1) / 2 block pairs, dealt to the
// warps; lane l owns agent l of block I and visits agent (l XOR s) of block J in step s, so own-row and partner-row
// accesses of a warp instruction hit 32 different banks and the partner's addresses are one LOP3 away from a
// per-round base.  Draws are shared-memory reductions (red.shared.or runs at the rate of a plain store), so nobody
// owns a row and the warps never wait for each other between the staging barrier and the epilogue.  The diagonal
// blocks (pairs inside a block) are evaluated once per unordered pair as well, two blocks per warp.
//
// Fast path: intervals of 1..32 bins (1 <= h <= 16) whose bin arithmetic is clear of every fp32 guard band are drawn
// with straight-line code: the inner loop of a round (16 iterations of two unordered pairs per lane) has no branch,
// vote or call.  Everything else is "slow" (1-2 % of the pairs of the benchmark scene): its draw is redirected to a
// scratch word and the pair remembered as a flag bit; after the round the flagged pairs are compacted into the warp's
// queue and worked off 32 at a time by the warp itself -- wide intervals by the general rule, guard-band hits by the
// fp64 path (the reference's own operation sequence, abm_vf_device.cuh), also per warp and 32 at a time -- while the
// other warps keep running their pair loops.
//
// Used when no distance culling is wanted (unequal radii: the HET variants, two half widths per pair), the engine owns whole replicates, the rows fit in
// shared memory and the batch is large enough to fill the GPU with one CTA per replicate; otherwise vf_step_kernel
// (abm_vf.cu) or vf_step_warp_kernel (abm_vf_warp.cu) runs (abm_api.cu).
#include "abm_vf_device.cuh"

namespace abm {

constexpr int kSymWarpQ = 64;        // entries per warp: worked off 32 at a time, as soon as 32 are there
constexpr int kSymQueueCap = 3072;   // fp64 queue entries of a CTA, split evenly among its warps (192 each at 16 warps; a
                                     // batch of 32 slow pairs adds at most 64); afterwards the exp(i Phi) table lives there

// Loop-invariant values of the pair loop that must live in registers (the compiler would otherwise
// re-materialise them with a MOV in every step).
struct SymConsts {
  float rS;            // radius * R/2pi
  float c3s, c5s, c7s; // -1/(3 S^2), 1/(5 S^4), -1/(7 S^6), S = R/2pi: atan(q) S = qs (1 + c3s qs^2 + c5s qs^4 + c7s qs^6)
  float qs_max;        // fast-path limit on qs (VFKernelArgs::sym_qs_max)
  float a6;            // leading coefficient of the bearing polynomial
  int scratch_pos;     // padded position of the scratch word
  unsigned long long half64;   // 2^31: rounding constant of the bin index, addend of its IMAD.WIDE
};
// acc |= m under a predicate: ONE predicated LOP3 (the compiler's own choice is SEL + LOP3)
__device__ __forceinline__ void or_if(uint32_t& acc, bool p, uint32_t m) {
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q or.b32 %0, %0, %2;\n\t}" : "+r"(acc) : "r"((uint32_t)p), "r"(m));
}

struct SymShared {
  float4* ag;          // [Np] (x, y, heading constant, the same + half a turn); padding agents beyond N are far away
  uint32_t rad_s;      // shared-space address of the radii [Np] (unequal radii: the HET variants), else 0
  uint32_t* rows;      // [W + 5][Np] padded rows (W + 2 words) + three scratch words (draws of slow directions land there)
  uint32_t* queue;     // [kSymQueueCap][2]  directions waiting for fp64 (focal << 16 | object, k << 16 | h), per warp
  uint32_t* warpq;     // [warps][kSymWarpQ] per-warp queue of pairs with directions off the fast path
  int* qcount;         // [0]: fp64 entries (statistics), [1]: pairs off the fast path (statistics), [2 + w]: fp64 entries
                       // waiting in warp w's part of the queue
  uint32_t ag_s, rows_s, queue_s, qcount_s;   // shared-space addresses of the above (code that is not inlined)
  int Np, N;
  int fq_cap;          // fp64 queue entries per warp
  const float4* rep_in; const float* th_in;   // the replicate's records and headings in global memory (fp64 path)
};

// Out-of-line fp64 evaluation + atomic draw of one ordered pair.  Returns 1 if the fp64 indices
// differ from the fp32 ones.
static __device__ __noinline__ unsigned sym_exact_and_draw(const VFKernelArgs& a, uint32_t row_s, uint32_t stride_b,
                                                           float4 f4, float fth, float4 o, int k32, int h32) {
  const FocalExact fe = vf_focal_exact(f4.x, f4.y, f4.z, fth);
  const PairExact pe = vf_pair_exact(fe, o.x, o.y, o.z, a.boundary, a.width_d, a.height_d, a.R, a.lin_step);
  if (pe.valid) vf_draw_shared(row_s, stride_b, a.R, a.fov_px0, a.fov_px1, pe.k, pe.h);
  return (pe.valid && ((pe.k != k32) | (pe.h != h32))) ? 1u : 0u;
}

// Slow path of one pair {i, j}; dirs bit 0: i sees j is off the fast path, bit 1: j sees i.  The complete fp32
// evaluation (full-range arctangent for the half width, relative guard band), shared by both directions like on the
// fast path; guard-band hits are queued for fp64, anything else visible is drawn by the general rule (wide
// intervals, wrap quirks) with reductions.
template <int RC>
__device__ __forceinline__ void sym_slow_dir(const VFKernelArgs& a, uint32_t rows_s, uint32_t queue_s, uint32_t qcount_s,
                                             int fq_cap, uint32_t stride_b, int f, int o, uint32_t nb, uint32_t hconst,
                                             int h, bool flagged) {
  const int R = RC ? RC : a.R;
  const int k = sym_side_k<RC>(a, nb, hconst, 0, flagged);           // bin index
  const uint32_t row_s = rows_s + stride_b + 4u * (uint32_t)f;        // real word 0
  if (flagged) {                                                      // into this warp's part of the fp64 queue
    const int slot = (int)atom_add_shared(qcount_s, 1u);
    if (slot < fq_cap) {
      sts_u32(queue_s + 8u * (uint32_t)slot, ((uint32_t)f << 16) | (uint32_t)o);
      sts_u32(queue_s + 8u * (uint32_t)slot + 4u, ((uint32_t)k << 16) | ((uint32_t)h & 0xffffu));
    } else {                                                          // full (cannot happen with batches of 32 pairs)
      const size_t g = (size_t)blockIdx.x * a.N;
      const unsigned diff = sym_exact_and_draw(a, row_s, stride_b, a.rec_in[g + f], a.theta[g + f], a.rec_in[g + o], k, h);
      atomicAdd(&a.counters[0], 1ull);
      atomicAdd(&a.counters[1], 1ull);
      if (diff) atomicAdd(&a.counters[2], 1ull);
    }
  } else {
    vf_draw_shared(row_s, stride_b, R, a.fov_px0, a.fov_px1, k, h);
  }
}
template <bool TORUS, int RC>
static __device__ __noinline__ void sym_slow_pair(const VFKernelArgs& a, uint32_t ag_s, uint32_t rad_s, uint32_t rows_s, uint32_t queue_s,
                                                  uint32_t qcount_s, int fq_cap, int Np, int i, int j, uint32_t dirs) {
  using K = PairK<RC>;
  if ((i >= a.N) | (j >= a.N)) return;                                // padding agents
  const float4 ia = lds_f4(ag_s + 16u * (uint32_t)i), ja = lds_f4(ag_s + 16u * (uint32_t)j);
  if ((ia.x == ja.x) & (ia.y == ja.y)) return;                        // vf_supcalc.py:57
  // unequal radii: centre = position + own radius (vf_supcalc.py:42, 60-63), and i sees j under j's radius
  const float ri = rad_s ? lds_f32(rad_s + 4u * (uint32_t)i) : a.sym_radius;
  const float rj = rad_s ? lds_f32(rad_s + 4u * (uint32_t)j) : a.sym_radius;
  const float dr = rj - ri;
  float dx = (ja.x - ia.x) + dr, dy = (ja.y - ia.y) + dr;             // positions first (exact for close neighbours), then radii
  bool wrap_tie = false;
  if (TORUS) {
    dx = torus_delta_r(ja.x, ia.x, dr, a.width, a.half_w, wrap_tie);
    dy = torus_delta_r(ja.y, ia.y, dr, a.height, a.half_h, wrap_tie);
  }
  const float d2 = fmaf(dx, dx, dy * dy);
  const float rs = rsqrt_approx(d2);
  const uint32_t nb = sym_bearing_bits(dx, dy, kBearingA6);           // bearing of j seen from i
  const uint32_t stride_b = 4u * (uint32_t)Np;
#pragma unroll 1
  for (int dir = 0; dir < 2; ++dir) {                                  // 0: i sees j (j's radius), 1: j sees i
    if (!((dirs >> dir) & 1u)) continue;
    const float q = (dir ? ri : rj) * rs;
    const float y = fmaf(atan_unit(q), K::y_scale(a), -0.5f);
    const float yr = y + kMagic;
    const int h = __float_as_int(yr) - kMagicBits;
    const bool flagged = !(q <= 1.0f) | (fmaf(y, a.nthr_h1, fabsf(y - (yr - kMagic))) > a.thr_h0) | wrap_tie;
    if (dir == 0) sym_slow_dir<RC>(a, rows_s, queue_s, qcount_s, fq_cap, stride_b, i, j, nb, __float_as_uint(ia.z), h, flagged);
    else sym_slow_dir<RC>(a, rows_s, queue_s, qcount_s, fq_cap, stride_b, j, i, nb, __float_as_uint(ja.w), h, flagged);
  }
}

// One entry of a warp's slow queue = one unordered pair with directions off the fast path: agent i | agent j << 10 |
// dirs << 20 (bit 0: i sees j, bit 1: j sees i; 0 = no entry).  A batch is 32 entries, one per lane of the converged
// warp, so the out-of-line evaluation always runs with as many lanes as there are entries.
// Guard-band hits wait in the warp's own part of the fp64 queue and are worked off by the warp itself, 32 at a time
// (`all`: whatever is left), as soon as a slow batch has left at least 32 there: like the slow batches, the
// latency-bound fp64 evaluation (the reference's own operation sequence) then overlaps the other warps' pair loops
// instead of forming a serial phase of the CTA.  Called by the converged warp.
struct SymWarpStats { unsigned n_fp64, n_mismatch; };
template <bool TORUS, int RC>
static __device__ __noinline__ void sym_fp64_drain(const VFKernelArgs& a, const SymShared& sh, SymWarpStats& ws, bool all) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t cnt_s = sh.qcount_s + 8u + 4u * (uint32_t)warp;
  const uint32_t q_s = sh.queue_s + 8u * (uint32_t)(warp * sh.fq_cap);
  const uint32_t stride_b = 4u * (uint32_t)sh.Np;
  int n = min((int)lds_u32(cnt_s), sh.fq_cap);
  __syncwarp();
  while (n >= 32 || (all && n > 0)) {
    const int take = min(n, 32);
    n -= take;
    if (lane < take) {
      const uint32_t q0 = lds_u32(q_s + 8u * (uint32_t)(n + lane)), q1 = lds_u32(q_s + 8u * (uint32_t)(n + lane) + 4u);
      const int f = (int)(q0 >> 16), o = (int)(q0 & 0xffffu);
      ws.n_mismatch += sym_exact_and_draw(a, sh.rows_s + stride_b + 4u * (uint32_t)f, stride_b, sh.rep_in[f], sh.th_in[f],
                                          sh.rep_in[o], (int)(short)(q1 >> 16), (int)(short)(q1 & 0xffffu));
    }
    ws.n_fp64 += (unsigned)take;                               // (every lane counts the batch; lane 0 reports)
    __syncwarp();
  }
  if (lane == 0) sts_u32(cnt_s, (uint32_t)n);
  __syncwarp();
}

template <bool TORUS, int RC>
static __device__ __noinline__ void sym_slow_batch(const VFKernelArgs& a, const SymShared& sh, uint32_t ent) {
  if (a.flags & (1u << 30)) return;   // timing probe (ABM_VF_DEBUG_SKIP_SLOW, results are wrong): what the slow pairs cost
  const uint32_t dirs = ent >> 20;
  const int warp = threadIdx.x >> 5;
  if (dirs) sym_slow_pair<TORUS, RC>(a, sh.ag_s, sh.rad_s, sh.rows_s, sh.queue_s + 8u * (uint32_t)(warp * sh.fq_cap),
                                     sh.qcount_s + 8u + 4u * (uint32_t)warp, sh.fq_cap, sh.Np, (int)(ent & 1023u),
                                     (int)((ent >> 10) & 1023u), dirs);
  __syncwarp();
}

// Work off full batches of the warp's queue (wcount entries, warp-uniform).
template <bool TORUS, int RC>
__device__ __forceinline__ void sym_drain(const VFKernelArgs& a, const SymShared& sh, uint32_t wq_s, int& wcount, int lane,
                                          SymWarpStats& ws) {
  while (wcount >= 32) {
    wcount -= 32;
    if (lane == 0) atom_add_shared(sh.qcount_s + 4u, 32u);     // statistics: pairs off the fast path (kernel choice)
    __syncwarp();
    sym_slow_batch<TORUS, RC>(a, sh, lds_u32(wq_s + 4u * (uint32_t)(wcount + lane)));
    if (lds_u32(sh.qcount_s + 8u + 4u * (uint32_t)(threadIdx.x >> 5)) >= 32u) sym_fp64_drain<TORUS, RC>(a, sh, ws, false);
    __syncwarp();
  }
}

// Diagonal blocks: append this lane's pair (if it has flagged directions) to the warp's queue.  Draws are reductions,
// so a batch can be worked off at any time -- the latency-bound out-of-line evaluation overlaps the other warps' pair
// loops instead of forming a serial phase.  wq_s: shared-space address of the warp's queue; wcount: entries in it
// (warp-uniform).  Called by the converged warp.
template <bool TORUS, int RC>
__device__ __forceinline__ void sym_push(const VFKernelArgs& a, const SymShared& sh, uint32_t wq_s, int& wcount, int lane,
                                         int i, int j, bool f0, bool f1, SymWarpStats& ws) {
  const bool any = f0 | f1;
  const uint32_t bal = __ballot_sync(0xffffffffu, any);
  if (bal) {                                                   // warp-uniform
    const uint32_t entry = (uint32_t)i | ((uint32_t)j << 10) | ((f0 ? 1u : 0u) << 20) | ((f1 ? 2u : 0u) << 20);
    if (any) sts_u32(wq_s + 4u * (uint32_t)(wcount + __popc(bal & ((1u << lane) - 1u))), entry);
    wcount += __popc(bal);
    sym_drain<TORUS, RC>(a, sh, wq_s, wcount, lane, ws);
  }
}

// The off-diagonal pair loop does not push inside its 16 iterations: every lane collects the slow flags of the round's
// 32 steps as bit s of two accumulators (own direction / partner's direction; one predicated OR each), and the round's
// pairs are appended to the warp's queue here, once per round: a warp scan of the lanes' counts gives every lane its
// place, then each lane stores its own entries (no ballot per entry).  A round that would overflow the queue (crowded
// scenes) takes the ballot loop, 32 entries at a time.
template <bool TORUS, int RC>
__device__ __forceinline__ void sym_push_round(const VFKernelArgs& a, const SymShared& sh, uint32_t wq_s, int& wcount,
                                               int lane, int i, int j0, uint32_t fi, uint32_t fj, SymWarpStats& ws) {
  uint32_t pend = fi | fj;                                     // bit s: the pair of step s has flagged directions
  const int cnt = __popc(pend);
  int incl = cnt;                                              // inclusive warp scan
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (wcount + total <= kSymWarpQ) {                           // warp-uniform; the usual case
    uint32_t pos = wq_s + 4u * (uint32_t)(wcount + incl - cnt);
    while (pend) {
      const int st = __ffs((int)pend) - 1;
      pend &= pend - 1u;
      const uint32_t dirs = ((fi >> st) & 1u) | (((fj >> st) & 1u) << 1);
      sts_u32(pos, (uint32_t)i | ((uint32_t)(j0 ^ st) << 10) | (dirs << 20));
      pos += 4u;
    }
    wcount += total;
    __syncwarp();
    sym_drain<TORUS, RC>(a, sh, wq_s, wcount, lane, ws);
  } else {
    while (__any_sync(0xffffffffu, pend != 0u)) {
      const bool have = pend != 0u;
      const int st = have ? __ffs((int)pend) - 1 : 0;
      pend &= pend - 1u;
      sym_push<TORUS, RC>(a, sh, wq_s, wcount, lane, i, j0 ^ st, have && ((fi >> st) & 1u), have && ((fj >> st) & 1u), ws);
    }
  }
}

// Result of the fp32 evaluation of one unordered pair: padded start positions of the two intervals
// (already redirected to the scratch word when the direction is off the fast path), the 2h-ones mask.
struct SymStep {
  int ps_i, ps_j;
  uint32_t mask, mask_hi;   // 2h ones: bits 0..31 / 32..63 (the lane's own direction; both when all radii are equal)
  uint32_t mask_j, mask_hi_j;   // the partner's direction (unequal radii: its half width comes from the OTHER radius)
  bool slow_i, slow_j;
};

// Pure arithmetic (no shared-memory access): the compiler interleaves two of these.
// BOTH: evaluate both directions; otherwise only the lane's own.
// One direction's half width: h = floor(atan(r / d) R / 2pi) from qs = (r R / 2pi) / d (vf_supcalc.py:96-99, :114-117).
struct SymHalf { uint32_t hraw, mask, mask_hi; bool slow; };
template <int RC, bool WIDE3>
__device__ __forceinline__ SymHalf sym_half_width(const VFKernelArgs& a, float qs, const SymConsts& c) {
  // Four-term series in q = r/d (truncation error 3e-6 bins at h = 32, q = 0.17, R = 1200); larger q -> slow path.
  // (With the fast path ending at h = 16 -- intervals of two words -- the pairs between 56 and 112 px, three
  // quarters of all wide ones, went through the out-of-line slow path: 29 % of the kernel's warp samples.)
  SymHalf r;
  const float zs = qs * qs;
  float p;
  if (WIDE3) { p = fmaf(zs, c.c7s, c.c5s); p = fmaf(p, zs, c.c3s); }   // four terms: exact to the guard band up to q = 0.18
  else p = fmaf(zs, c.c5s, c.c3s);                                      // three: up to q = 0.12
  p = fmaf(p, zs, 1.0f);
  const float y = fmaf(qs, p, -0.5f);
  const float yr = y + kMagic;
  r.hraw = __float_as_uint(yr);                              // h + kMagicBits
  // (the limit is on qs, not on y: beyond q = 0.18 / 0.12 the truncated series is inexact, and the four-term one finally
  // changes sign; at R = 1200 the two-word limit h <= 16 is q < 0.087 and the three-term series only grows: the
  // benchmark's variant keeps its immediate compare)
  const bool too_wide = (!WIDE3 && RC == 1200) ? !(y < 16.5f) : !(qs < c.qs_max);   // also d2 == 0 (inf / NaN)
  r.slow = too_wide | (fabsf(y - (yr - kMagic)) > a.sym_thr_h);
  const uint32_t w2 = 2u * r.hraw - 2u * (uint32_t)kMagicBits; // 2h
  asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(r.mask) : "r"(w2));                                      // min(2h, 32) ones
  r.mask_hi = 0u;
  if (WIDE3) asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(r.mask_hi) : "r"((uint32_t)max((int)w2 - 32, 0)));   // the ones beyond 32
  return r;
}

// Pure arithmetic (no shared-memory access): the compiler interleaves two of these.
// BOTH: evaluate both directions; otherwise only the lane's own.
// HET: unequal radii -- centres are positions + OWN radii (vf_supcalc.py:42, 60-63: distance and bearing are still shared by
// the two directions, bit for bit: every term of the other direction's difference is the exact negative), but i sees j under
// j's radius and j sees i under i's: two half widths.  rS_o / rS_me: the partner's / the lane's radius times R / 2pi,
// dr: partner's radius - own.
template <bool TORUS, bool FULL_FOV, int RC, bool BOTH, bool WIDE3, bool HET = false>
__device__ __forceinline__ SymStep sym_eval(const VFKernelArgs& a, float4 o, float xi, float yi, uint32_t hc_i,
                                            const SymConsts& c, float rS_o = 0.f, float rS_me = 0.f, float dr = 0.f) {
  SymStep r;
  float dx = o.x - xi, dy = o.y - yi;
  if (HET) { dx += dr; dy += dr; }                           // positions first (exact for close neighbours), then radii
  bool wrap_tie = false;
  if (TORUS) {                                               // vf_supcalc.py:70-83
    if (HET) {
      dx = torus_delta_r(o.x, xi, dr, a.width, a.half_w, wrap_tie);
      dy = torus_delta_r(o.y, yi, dr, a.height, a.half_h, wrap_tie);
    } else {
      dx = torus_delta(o.x, xi, a.width, a.half_w, wrap_tie);
      dy = torus_delta(o.y, yi, a.height, a.half_h, wrap_tie);
    }
  }
  if (HET) wrap_tie |= (o.x == xi) & (o.y == yi);            // coincident POSITIONS are skipped (vf_supcalc.py:57): slow path
  const float d2 = fmaf(dx, dx, dy * dy);
  // ---- half width(s): shared by both directions when all radii are equal ----
  const float rs = rsqrt_approx(d2);
  const SymHalf hi_ = sym_half_width<RC, WIDE3>(a, rs * (HET ? rS_o : c.rS), c);   // q * R/2pi
  const bool slow_h = hi_.slow | wrap_tie;
  r.mask = hi_.mask; r.mask_hi = hi_.mask_hi;
  const int bh = 32 + kMagicBits - (int)hi_.hraw;            // ps = bin index + 32 - h
  // ---- bearing (shared), bin index of both directions ----
  const uint32_t nb = sym_bearing_bits(dx, dy, c.a6);
  r.slow_i = slow_h;
  const int ps_i = sym_side_k<RC, true>(a, nb, hc_i, bh, r.slow_i, c.half64);
  bool draw_i = !r.slow_i;
  if (!FULL_FOV) {
    const int pe = ps_i + 2 * ((int)hi_.hraw - kMagicBits);
    draw_i &= ((unsigned)(ps_i - a.fov0p) < a.span) | ((unsigned)(pe - a.fov0p) < a.span);   // vf_supcalc.py:119
  }
  r.ps_i = draw_i ? ps_i : c.scratch_pos;
  r.slow_j = false;
  r.ps_j = c.scratch_pos;
  r.mask_j = r.mask; r.mask_hi_j = r.mask_hi;
  if (BOTH) {
    uint32_t hraw_j = hi_.hraw;
    int bh_j = bh;
    r.slow_j = slow_h;
    if (HET) {                                               // the partner sees this lane's agent under ITS radius
      const SymHalf hj = sym_half_width<RC, WIDE3>(a, rs * rS_me, c);
      r.slow_j = hj.slow | wrap_tie;
      r.mask_j = hj.mask; r.mask_hi_j = hj.mask_hi;
      hraw_j = hj.hraw;
      bh_j = 32 + kMagicBits - (int)hj.hraw;
    }
    const int ps_j = sym_side_k<RC, true>(a, nb, __float_as_uint(o.w), bh_j, r.slow_j, c.half64);   // o.w: heading constant + half a turn
    bool draw_j = !r.slow_j;
    if (!FULL_FOV) {
      const int pe = ps_j + 2 * ((int)hraw_j - kMagicBits);
      draw_j &= ((unsigned)(ps_j - a.fov0p) < a.span) | ((unsigned)(pe - a.fov0p) < a.span);
    }
    r.ps_j = draw_j ? ps_j : c.scratch_pos;
  }
  return r;
}

// OR an interval of <= 32 bins into the one or two row words it touches with shared-memory reductions (RED.OR runs
// at the rate of a plain shared store: one wavefront per conflict-free warp instruction, measured on B200 --
// scratch/red_bench.cu).  Nobody has to own a row, draws need no ordering, and the second word is written
// unconditionally (its mask is usually 0): that is cheaper than a predicate.
template <bool WIDE3>
__device__ __forceinline__ void sym_red(uint32_t row, uint32_t stride_b, int ps, uint32_t m, uint32_t mh) {
  uint32_t wa;   // row + (ps >> 5) * stride_b as ONE multiply-add (with a power-of-two immediate stride the compiler
                 // would otherwise build it from a shift, a mask and an add)
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(wa) : "r"((uint32_t)(ps >> 5)), "r"(stride_b), "r"(row));
  asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(wa), "r"(__funnelshift_l(0u, m, ps)));
  if (WIDE3) {   // up to 64 bins: three words
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(wa + stride_b), "r"(__funnelshift_l(m, mh, ps)));
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(wa + 2u * stride_b), "r"(__funnelshift_l(mh, 0u, ps)));
  } else {
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(wa + stride_b), "r"(__funnelshift_l(m, 0u, ps)));
  }
}

size_t vf_sym_smem_bytes(int Np, int W, bool wide3, bool het) {
  return sizeof(float4) * (size_t)Np + sizeof(uint32_t) * (size_t)(W + (wide3 ? 5 : 3)) * Np + 2 * sizeof(uint32_t) * kSymQueueCap +
         sizeof(uint32_t) * kSymWarpQ * (size_t)(Np / 64) + 96 +   // counters: 2 + one per warp (<= 16)
         (het ? sizeof(float) * (size_t)Np + 128 : 0);            // the radii (128-byte aligned)
}

// NPC > 0: compile-time padded replicate size (row stride becomes an immediate), 0: the run-time argument.
// WIDE3: the fast path takes intervals of up to 64 bins (three row words, four-term series) instead of 32 -- a few more
// instructions on every pair, but in a crowded scene the pairs between 56 and 112 px stay out of the slow path (the
// engine switches on the measured share of slow pairs, abm_api.cu).
// HET: unequal radii (two half widths per pair, sym_eval); the radii live in shared memory next to the records.
template <bool TORUS, bool FULL_FOV, int RC, int NPC, bool WIDE3, bool HET>
__global__ void __launch_bounds__(512, 1) vf_step_sym_kernel(const __grid_constant__ VFKernelArgs a, int Np_arg) {
  const int Np = NPC ? NPC : Np_arg;
  using K = PairK<RC>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SymShared sh;
  sh.ag = reinterpret_cast<float4*>(smem_raw);
  sh.rows = reinterpret_cast<uint32_t*>(sh.ag + Np);
  constexpr int kRowExtra = WIDE3 ? 5 : 3;   // padding (2) + scratch words for redirected draws (1 per word of a draw)
  sh.queue = sh.rows + (size_t)(a.W + kRowExtra) * Np;
  sh.warpq = sh.queue + 2 * kSymQueueCap;
  sh.qcount = reinterpret_cast<int*>(sh.warpq + kSymWarpQ * (Np / 64));
  // [Np] (HET only), on a 128-byte boundary: the pair loop walks it with the same XOR as the rows
  float* rad = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sh.qcount + 24) + 127) & ~uintptr_t(127));
  sh.rad_s = HET ? smem_u32(rad) : 0u;
  sh.Np = Np; sh.N = a.N;
  sh.fq_cap = kSymQueueCap / (int)(blockDim.x >> 5);
  sh.rep_in = a.rec_in + (size_t)blockIdx.x * a.N; sh.th_in = a.theta + (size_t)blockIdx.x * a.N;
  sh.ag_s = smem_u32(sh.ag); sh.rows_s = smem_u32(sh.rows); sh.queue_s = smem_u32(sh.queue);
  sh.qcount_s = smem_u32(sh.qcount);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = blockDim.x;
  const int b = blockIdx.x;
  const int N = a.N;
  const int P = Np >> 5;                      // number of 32-agent blocks (even)
  const float4* rep_in = a.rec_in + (size_t)b * N;
  const float* th_in = a.theta + (size_t)b * N;

  // ---- stage the replicate: (x, y, heading in bins of the linspace grid, radius) ----
  for (int j = tid; j < Np; j += T) {
    float4 v;
    if (j < N) {
      const float4 r4 = rep_in[j];
      const uint32_t hc = sym_heading_const(th_in[j]);
      v = make_float4(r4.x, r4.y, __uint_as_float(hc), __uint_as_float(hc - 0x80000000u));
    } else {   // padding: far away (half width 0), all distinct
      v = make_float4(-1.0e6f - 4096.0f * (float)(j - N), -1.0e6f, 0.f, 0.f);
    }
    sh.ag[j] = v;
    if (HET) rad[j] = (j < N) ? rep_in[j].z : a.sym_radius;
  }
  for (int w = tid; w < (a.W + kRowExtra) * Np; w += T) sh.rows[w] = 0u;
  if (tid < 18) sh.qcount[tid] = 0;
  __syncthreads();

  const float S = K::y_scale(a);
  SymConsts c;
  c.rS = a.sym_radius * S;
  c.c3s = -1.0f / (3.0f * S * S);
  c.qs_max = 0.f; c.c7s = 0.f;
  if (WIDE3 || RC != 1200) c.qs_max = __uint_as_float(__float_as_uint(WIDE3 ? a.sym_qs_max : a.sym_qs_max2) | a.opaque_zero);
  if (WIDE3) c.c7s = __uint_as_float(__float_as_uint(-1.0f / (7.0f * S * S * S * S * S * S)) | a.opaque_zero);
  // OR-ed with a kernel argument that is always 0: a value ptxas cannot rebuild with one MOV, so it stays in a register
  c.c5s = __uint_as_float(__float_as_uint(1.0f / (5.0f * S * S * S * S)) | a.opaque_zero);
  c.a6 = __uint_as_float(__float_as_uint(kBearingA6) | a.opaque_zero);
  c.scratch_pos = 32 * (a.W + 2);
  c.half64 = ((unsigned long long)a.opaque_zero << 32) | (0x80000000u | a.opaque_zero);
  const uint32_t ag_s = smem_u32(sh.ag), rows_s = smem_u32(sh.rows);
  const uint32_t stride_b = 4u * (uint32_t)Np;

  const uint32_t wq = smem_u32(sh.warpq + kSymWarpQ * warp);   // this warp's queue of slow entries
  int wcount = 0;                              // entries in it (warp-uniform)
  SymWarpStats ws{0u, 0u};                     // fp64 entries worked off by this warp (warp-uniform) / mismatches (per lane)

  // ---- diagonal blocks, two per warp: the pair {x, x ^ s} of a block is taken, in step s, by the one of its two
  //      lanes whose bit hb(s) (the highest set bit of s) is clear -- for the warp's first block; for its second block
  //      by the lane whose bit is set.  So every lane has a pair in every step, each unordered pair of both blocks is
  //      evaluated once, and own / partner rows of a warp instruction still hit 32 different banks. ----
  {
    const float4 me0 = sh.ag[(2 * warp << 5) + lane], me1 = sh.ag[((2 * warp + 1) << 5) + lane];
    const float r0 = HET ? rad[(2 * warp << 5) + lane] : 0.f, r1 = HET ? rad[((2 * warp + 1) << 5) + lane] : 0.f;
#pragma unroll 1
    for (int s = 1; s < 32; ++s) {
      const int sel = (lane >> (31 - __clz(s))) & 1;
      const int i = ((2 * warp + sel) << 5) + lane, j = i ^ s;
      const float4 me = sel ? me1 : me0;
      const float rme = sel ? r1 : r0, rj = HET ? rad[j] : 0.f;
      const SymStep A = sym_eval<TORUS, FULL_FOV, RC, true, WIDE3, HET>(a, lds_f4(ag_s + 16u * (uint32_t)j), me.x, me.y,
                                                                 __float_as_uint(me.z), c, rj * S, rme * S, rj - rme);
      sym_red<WIDE3>(rows_s + 4u * (uint32_t)i, stride_b, A.ps_i, A.mask, A.mask_hi);
      sym_red<WIDE3>(rows_s + 4u * (uint32_t)j, stride_b, A.ps_j, A.mask_j, A.mask_hi_j);
      sym_push<TORUS, RC>(a, sh, wq, wcount, lane, i, j, A.slow_i, A.slow_j, ws);
    }
  }

  // ---- off-diagonal block pairs: the P (P - 1) / 2 pairs are dealt to the warps like the rounds of a round-robin
  //      tournament (circle method); draws are reductions, so the warps never wait for each other ----
  const int m = P - 1;
  for (int round = 0; round < m; ++round) {
    int I, J;
    if (warp == 0) { I = m; J = round; }
    else { I = (round + warp) % m; J = (round - warp + m) % m; }
    const int i = (I << 5) + lane, j0 = (J << 5) + lane;
    const float4 me = sh.ag[i];
    const uint32_t rec_j0 = ag_s + 16u * (uint32_t)j0, row_j0 = rows_s + 4u * (uint32_t)j0;
    const uint32_t row_i = rows_s + 4u * (uint32_t)i;
    // two steps per iteration: both evaluations are independent arithmetic (instruction-level parallelism for the
    // 4 warps per scheduler this kernel runs with)
    float4 oA = lds_f4(rec_j0), oB = lds_f4(rec_j0 ^ 16u);
    uint32_t fi = 0u, fj = 0u, tbit = 1u;                  // slow flags of the round (own / partner's direction), bit s = step s
    const uint32_t rec_j1 = rec_j0 ^ 16u, row_j1 = row_j0 ^ 4u;
    const float rme = HET ? rad[i] : 0.f, rSme = rme * S;
    const uint32_t rad_j0 = sh.rad_s + 4u * (uint32_t)j0;      // (the radii are 4-byte words like the rows: same XOR walk)
#pragma unroll 1
    for (uint32_t o16 = 32u; o16 <= 512u; o16 += 32u) {       // o16 = 16 (s + 2), s = 0, 2, .. 30: ONE induction variable
      // prefetch the next two partner records (the last iteration reads block J ^ 1: harmless)
      const float4 nA = lds_f4(rec_j0 ^ o16), nB = lds_f4(rec_j1 ^ o16);
      const uint32_t o4 = (o16 >> 2) - 8u;                     // 4 s
      const float rA = HET ? lds_f32(rad_j0 ^ o4) : 0.f, rB = HET ? lds_f32(rad_j0 ^ 4u ^ o4) : 0.f;
      const SymStep A = sym_eval<TORUS, FULL_FOV, RC, true, WIDE3, HET>(a, oA, me.x, me.y, __float_as_uint(me.z), c, rA * S, rSme, rA - rme);
      const SymStep B = sym_eval<TORUS, FULL_FOV, RC, true, WIDE3, HET>(a, oB, me.x, me.y, __float_as_uint(me.z), c, rB * S, rSme, rB - rme);
      sym_red<WIDE3>(row_i, stride_b, A.ps_i, A.mask, A.mask_hi);
      sym_red<WIDE3>(row_j0 ^ o4, stride_b, A.ps_j, A.mask_j, A.mask_hi_j);
      sym_red<WIDE3>(row_i, stride_b, B.ps_i, B.mask, B.mask_hi);
      sym_red<WIDE3>(row_j1 ^ o4, stride_b, B.ps_j, B.mask_j, B.mask_hi_j);
      // ---- off the fast path (~1 % of the directions): remembered, pushed after the round ----
      or_if(fi, A.slow_i, tbit);
      or_if(fj, A.slow_j, tbit);
      or_if(fi, B.slow_i, tbit + tbit);
      or_if(fj, B.slow_j, tbit + tbit);
      tbit <<= 2;
      oA = nA; oB = nB;
    }
    sym_push_round<TORUS, RC>(a, sh, wq, wcount, lane, i, j0, fi, fj, ws);
  }
  // the rest of the warp's queue (fewer than 32 entries), then the rest of its fp64 queue
  __syncwarp();
  if (lane == 0) atom_add_shared(sh.qcount_s + 4u, (uint32_t)wcount);
  sym_slow_batch<TORUS, RC>(a, sh, lane < wcount ? lds_u32(wq + 4u * (uint32_t)lane) : 0u);
  sym_fp64_drain<TORUS, RC>(a, sh, ws, true);
  {
    const unsigned nm = __reduce_add_sync(0xffffffffu, ws.n_mismatch);
    if (lane == 0 && ws.n_fp64) atomicAdd(&a.counters[0], (unsigned long long)ws.n_fp64);
    if (lane == 0 && nm) atomicAdd(&a.counters[2], (unsigned long long)nm);
  }
  __syncthreads();
  if (tid == 0 && sh.qcount[1]) atomicAdd(&a.counters[4], (unsigned long long)sh.qcount[1]);

  // ---- epilogue: one agent per thread and pass (bank == lane) ----
  // the fp64 queue is dead now: it takes a copy of exp(i Phi_k) for the edge sums (19 KB at R = 1200)
  uint32_t etab_s = 0u;
  if ((size_t)a.R * sizeof(double2) <= 2 * sizeof(uint32_t) * kSymQueueCap) {
    double2* etab = reinterpret_cast<double2*>(sh.queue);
    for (int k = tid; k < a.R; k += T) etab[k] = *reinterpret_cast<const double2*>(&a.lut[k].c);
    __syncthreads();
    etab_s = sh.queue_s;
  }
  for (int i0 = 0; i0 < N; i0 += 2 * T) {      // all lanes of a warp go through the epilogue together; two agents per
    int ii[2]; bool aa[2]; uint32_t* pr[2];    // thread give its latency-bound edge loop independent work
    float4 mm[2]; float tt[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int i = i0 + j * T + tid;
      aa[j] = i < N;
      ii[j] = aa[j] ? i : 0;
      pr[j] = sh.rows + ii[j];
      mm[j] = rep_in[ii[j]];
      tt[j] = th_in[ii[j]];
    }
    uint32_t* const prc[2] = {pr[0], pr[1]};
    vf_agent_epilogue<TORUS, 2>(a, b, ii, ii, prc, Np, mm, tt, etab_s, aa);
  }
}

bool vf_sym_applicable(const VFKernelArgs& a, bool uniform_r, bool cull, size_t smem_limit) {
  if (cull) return false;
  if (a.tile_begin != 0 || a.tile_count != a.N) return false;
  const int Np = (a.N + 63) / 64 * 64;
  if (Np > 1024) return false;   // 16 warps at most; queue entries hold 16-bit agent indices
  return vf_sym_smem_bytes(Np, a.W, true, !uniform_r) <= smem_limit;
}

template <bool TORUS, bool FULL_FOV, int RC, int NPC, bool WIDE3, bool HET>
static void launch_sym_variant(const VFKernelArgs& a, int Np, int threads, size_t smem, cudaStream_t stream) {
  static SmemOptIn optin;   // per device (abm_common.cuh)
  optin.ensure(vf_step_sym_kernel<TORUS, FULL_FOV, RC, NPC, WIDE3, HET>, smem);
  vf_step_sym_kernel<TORUS, FULL_FOV, RC, NPC, WIDE3, HET><<<a.B, threads, smem, stream>>>(a, Np);
}

template <bool TORUS, bool WIDE3>
static void launch_sym_fov(const VFKernelArgs& a, int Np, int threads, size_t smem, bool het, cudaStream_t stream) {
  if (het) {   // unequal radii: the generic instantiations only
    if (a.full_fov) launch_sym_variant<TORUS, true, 0, 0, WIDE3, true>(a, Np, threads, smem, stream);
    else launch_sym_variant<TORUS, false, 0, 0, WIDE3, true>(a, Np, threads, smem, stream);
  } else if (a.full_fov && a.R == 1200 && Np == 1024) launch_sym_variant<TORUS, true, 1200, 1024, WIDE3, false>(a, Np, threads, smem, stream);
  else if (a.full_fov && a.R == 1200) launch_sym_variant<TORUS, true, 1200, 0, WIDE3, false>(a, Np, threads, smem, stream);
  else if (a.full_fov) launch_sym_variant<TORUS, true, 0, 0, WIDE3, false>(a, Np, threads, smem, stream);
  else launch_sym_variant<TORUS, false, 0, 0, WIDE3, false>(a, Np, threads, smem, stream);
}

void launch_vf_step_sym(const VFKernelArgs& a, bool wide3, bool uniform_r, cudaStream_t stream) {
  const int Np = (a.N + 63) / 64 * 64;
  const int threads = 32 * (Np / 64);
  const bool het = !uniform_r;
  const size_t smem = vf_sym_smem_bytes(Np, a.W, wide3, het);
  if (a.boundary == 1) { if (wide3) launch_sym_fov<true, true>(a, Np, threads, smem, het, stream); else launch_sym_fov<true, false>(a, Np, threads, smem, het, stream); }
  else { if (wide3) launch_sym_fov<false, true>(a, Np, threads, smem, het, stream); else launch_sym_fov<false, false>(a, Np, threads, smem, het, stream); }
}

}  // namespace abm
