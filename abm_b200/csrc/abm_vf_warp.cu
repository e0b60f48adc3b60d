// Warp-per-focal-agent visual-flocking step kernel for sm_100a: the kernel for ONE LARGE, SPARSE swarm and for its
// agent tiles on several GPUs.
//
// Why: with one thread per focal agent (abm_vf.cu) a 65 536-agent swarm is only 2048 warps -- 3.5 per scheduler on one
// B200, 0.4 on each of 8 -- and every thread walks thousands of neighbour records in sequence: latency-bound (ncu: 32 %
// of the issue slots).  Here a WARP owns a focal agent: its lanes stride over the neighbour records that survive the
// tile-level culling (coalesced 512-byte reads straight from the L2-resident record table), evaluate the pair with the
// same arithmetic as the symmetric kernel (32-bit binary angles, one IMAD.WIDE for bin + guard band; full-range
// arctangent for the half width) and OR the interval into the warp's row in shared memory with RED.OR reductions
// (abm_vf_sym.cu explains why those are cheap).  65 536 warps per step instead of 2048; per-lane culling costs no
// divergence.  Guard-band hits are re-evaluated on the spot in fp64 (the reference's own operation sequence).
// The epilogue is warp-cooperative: lanes take the words of the row, edge sums are reduced with shuffles, lane 0
// finishes (terms, kinematics, walls / torus, outputs, peer stores of the fused tile exchange).
//
// CTA = 8 warps = 8 consecutive focal agents of one replicate (spatially close when the engine keeps its Morton
// order), sharing one list of record tiles to visit (bounding boxes, as in abm_vf.cu).
#include "abm_vf_device.cuh"

namespace abm {

constexpr int kWarpsPerCta = 8;

size_t vf_warp_smem_bytes(int W) {
  return sizeof(uint32_t) * (size_t)(W + 1) * kWarpsPerCta + sizeof(int) * (kMaxTileList + 4) + 64;
}

template <bool TORUS, bool CULL>
__global__ void __launch_bounds__(kWarpsPerCta * 32) vf_step_warp_kernel(const __grid_constant__ VFKernelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw);                        // [warps][W + 1]
  int* tile_list = reinterpret_cast<int*>(rows + (size_t)(a.W + 1) * kWarpsPerCta);   // [kMaxTileList] + count
  float* fbox = reinterpret_cast<float*>(tile_list + kMaxTileList + 4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per_rep = (a.tile_count + kWarpsPerCta - 1) / kWarpsPerCta;
  const int b = blockIdx.x / per_rep;
  const int li = (blockIdx.x - b * per_rep) * kWarpsPerCta + warp;   // index inside this engine's focal tile
  const bool active = li < a.tile_count;
  const int i = a.tile_begin + (active ? li : 0);
  const size_t gi = (size_t)b * a.N + i;
  const float4* rep_in = a.rec_in + (size_t)b * a.N;
  const int R = a.R, W = a.W;

  uint32_t* row = rows + (size_t)(W + 1) * warp;
  for (int w = lane; w < W + 1; w += 32) row[w] = 0u;
  if (a.n_peers > 0 && tid <= a.n_peers) {   // fused tile exchange: wait for every rank's previous step (abm_vf.cu)
    const uint32_t* f = a.xflags + tid;
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    } while ((int)(v - a.step_no) < 0);
  }
  const float4 me = rep_in[i];
  const float th = a.theta[gi];

  // ---- record tiles to visit: bounding box of the CTA's focal agents against the tiles' boxes ----
  const int tile_sz = a.tile_bbox != nullptr ? a.cull_tile : kRecTile;
  const int n_tiles = (a.N + tile_sz - 1) / tile_sz;
  const bool use_list = CULL && a.tile_bbox != nullptr && n_tiles <= kMaxTileList;
  int n_stage = n_tiles;
  if (use_list) {
    if (tid < 4) fbox[tid] = (tid < 2) ? 3.0e38f : -3.0e38f;
    if (tid == 0) tile_list[kMaxTileList] = 0;
    __syncthreads();
    if (lane == 0 && active) {
      atomicMin(reinterpret_cast<int*>(&fbox[0]), __float_as_int(fmaxf(me.x, 0.0f)));
      atomicMin(reinterpret_cast<int*>(&fbox[1]), __float_as_int(fmaxf(me.y, 0.0f)));
      atomicMax(reinterpret_cast<int*>(&fbox[2]), __float_as_int(fmaxf(me.x, 0.0f)));
      atomicMax(reinterpret_cast<int*>(&fbox[3]), __float_as_int(fmaxf(me.y, 0.0f)));
    }
    __syncthreads();
    const float fx0 = fbox[0], fy0 = fbox[1], fx1 = fbox[2], fy1 = fbox[3];
    const float4* bb = a.tile_bbox + (size_t)b * n_tiles;
    const float* c2 = a.tile_cull2 + (size_t)b * n_tiles;
    for (int t = tid; t < n_tiles; t += blockDim.x) {
      const float4 q = bb[t];
      float gx = fmaxf(0.0f, fmaxf(q.x - fx1, fx0 - q.z));
      float gy = fmaxf(0.0f, fmaxf(q.y - fy1, fy0 - q.w));
      if (TORUS) {   // minimal image: the tile shifted by one period either way
        gx = fminf(gx, fmaxf(0.0f, fmaxf(q.x + a.width - fx1, fx0 - (q.z + a.width))));
        gx = fminf(gx, fmaxf(0.0f, fmaxf(q.x - a.width - fx1, fx0 - (q.z - a.width))));
        gy = fminf(gy, fmaxf(0.0f, fmaxf(q.y + a.height - fy1, fy0 - (q.w + a.height))));
        gy = fminf(gy, fmaxf(0.0f, fmaxf(q.y - a.height - fy1, fy0 - (q.w - a.height))));
      }
      const float reach = sqrtf(c2[t]) + a.bbox_slack + 1.0f;
      if (gx * gx + gy * gy <= reach * reach) tile_list[atomicAdd(&tile_list[kMaxTileList], 1)] = t;
    }
    __syncthreads();
    n_stage = tile_list[kMaxTileList];
  } else {
    __syncthreads();
  }

  // ---- pair loop: lanes stride over the records of the visited tiles ----
  unsigned n_fp64 = 0, n_differ = 0;
  if (active) {
    const uint32_t hc = sym_heading_const(th);
    const FocalExact fe = vf_focal_exact(me.x, me.y, me.z, th);
    const uint32_t row_s = smem_u32(row);
    for (int st = 0; st < n_stage; ++st) {
      const int t = use_list ? tile_list[st] : st;
      const int j0 = t * tile_sz, j1 = min(a.N, j0 + tile_sz);
      for (int j = j0 + lane; j < j1; j += 32) {
        const float4 o = __ldg(rep_in + j);
        const float dr = o.z - me.z;
        float dx = (o.x - me.x) + dr, dy = (o.y - me.y) + dr;   // positions first (exact for close neighbours), then radii
        bool wrap_tie = false;
        if (TORUS) {                                           // vf_supcalc.py:70-83
          dx = torus_delta_r(o.x, me.x, dr, a.width, a.half_w, wrap_tie);
          dy = torus_delta_r(o.y, me.y, dr, a.height, a.half_h, wrap_tie);
        }
        const float d2 = fmaf(dx, dx, dy * dy);
        if (CULL) { if (d2 > o.w) continue; }                  // beyond it the half width is 0
        if ((o.x == me.x) & (o.y == me.y)) continue;           // self / coincident positions (vf_supcalc.py:57)
        const float q = o.z * rsqrt_approx(d2);
        const float y = fmaf(atan_unit(q), a.y_scale, -0.5f);
        const float yr = y + kMagic;
        int h = __float_as_int(yr) - kMagicBits;
        bool flagged = !(q <= 1.0f) | (fmaf(y, a.nthr_h1, fabsf(y - (yr - kMagic))) > a.thr_h0) | wrap_tie;
        int k = sym_side_k<0>(a, sym_bearing_bits(dx, dy, kBearingA6), hc, 0, flagged);   // bin index
        if (flagged) {                                         // fp64, the reference's own operation sequence
          const PairExact pe = vf_pair_exact(fe, o.x, o.y, o.z, a.boundary, a.width_d, a.height_d, R, a.lin_step);
          ++n_fp64;
          if (!pe.valid) continue;
          if ((pe.k != k) | (pe.h != h)) ++n_differ;
          k = pe.k; h = pe.h;
        }
        const int ps = k - h, pe_ = k + h;
        if (((unsigned)(h - 1) < 16u) & (ps >= 0) & (pe_ < R) &
            (((a.fov_px0 < ps) & (ps < a.fov_px1)) | ((a.fov_px0 < pe_) & (pe_ < a.fov_px1)))) {
          // interval of <= 32 bins inside the row: at most two words
          const uint32_t m = 0xffffffffu >> (32 - 2 * h);
          const uint32_t wa = row_s + 4u * (uint32_t)(ps >> 5);
          red_or_shared(wa, __funnelshift_l(0u, m, ps));
          const uint32_t hi = __funnelshift_l(m, 0u, ps);
          if (hi) red_or_shared(wa + 4u, hi);
        } else {
          vf_draw_shared(row_s, 4u, R, a.fov_px0, a.fov_px1, k, h);   // wide / wrapping / outside the FOV: general rule
        }
      }
    }
  }
  __syncwarp();
  {
    const unsigned nf = __reduce_add_sync(0xffffffffu, n_fp64), nd = __reduce_add_sync(0xffffffffu, n_differ);
    if (lane == 0 && nf) atomicAdd(&a.counters[0], (unsigned long long)nf);
    if (lane == 0 && nd) atomicAdd(&a.counters[2], (unsigned long long)nd);
  }

  // ---- epilogue: lanes take the words of the row, edge sums reduced with shuffles ----
  if (active) {
    const uint32_t last_valid = (R & 31) ? ((1u << (R & 31)) - 1u) : 0xffffffffu;
    const uint32_t v_first = row[0] & 1u;
    const uint32_t v_last = (row[W - 1] >> ((R - 1) & 31)) & 1u;
    double zsr = 0.0, zsi = 0.0, zdr = 0.0, zdi = 0.0;
    if (a.phi_ok) {
      for (int w = lane; w < W; w += 32) {
        const uint32_t cur = row[w];
        const uint32_t carry = w ? (row[w - 1] >> 31) : v_last;     // ring predecessor of the word's bin 0
        uint32_t diff = cur ^ ((cur << 1) | carry);
        if (w == W - 1) diff &= last_valid;
        while (diff) {
          const int bit = __ffs(diff) - 1;
          diff &= diff - 1;
          const double2 e = *reinterpret_cast<const double2*>(&a.lut[(w << 5) + bit].c);
          const int sgn = (int)((cur >> bit) << 31);                 // rising edge: negative in Z_fall - Z_rise
          zsr += e.x; zsi += e.y;
          zdr += __hiloint2double(__double2hiint(e.x) ^ sgn, __double2loint(e.x));
          zdi += __hiloint2double(__double2hiint(e.y) ^ sgn, __double2loint(e.y));
        }
      }
      for (int off = 16; off > 0; off >>= 1) {
        zsr += __shfl_xor_sync(0xffffffffu, zsr, off); zsi += __shfl_xor_sync(0xffffffffu, zsi, off);
        zdr += __shfl_xor_sync(0xffffffffu, zdr, off); zdi += __shfl_xor_sync(0xffffffffu, zdi, off);
      }
    }
    if (lane == 0) {
      const VFParams6 prm = *reinterpret_cast<const VFParams6*>(a.params + (size_t)b * a.param_stride);
      double A0 = prm.alp0, B0 = prm.bet0, V0 = prm.v0;            // vf_supcalc.py:191-196
      if (a.ov_alp0) { const float v = a.ov_alp0[gi]; if (v == v) A0 = v; }
      if (a.ov_bet0) { const float v = a.ov_bet0[gi]; if (v == v) B0 = v; }
      if (a.ov_v0)   { const float v = a.ov_v0[gi];   if (v == v) V0 = v; }
      const double vel0 = a.vel[gi];
      FlockTerms ft;
      if (a.phi_ok) {
        EdgeSums z;
        z.zsr = zsr; z.zsi = zsi; z.zdr = zdr; z.zdi = zdi; z.v_first = v_first; z.v_last = v_last;
        ft = vf_terms_from_edges(z, a, vel0, prm, A0, B0, V0);
      } else {   // len(PHI) != len(soc_v_field): the reference skips the calculation (vf_agent.py:282-284)
        ft.dvel = ft.dpsi = ft.a_blob = ft.a_edge = ft.b_blob = ft.b_edge = 0.0;
      }
      double dpsi = ft.dpsi, dvel = ft.dvel;
      if (a.limit_movement) dpsi = limit_abs(dpsi, a.max_th);       // vf_agent.py:293-294
      double nth = wrap_heading_once((double)th + dpsi);            // :295-296
      double nv = vel0 + dvel;                                      // :298
      if (a.limit_movement) nv = limit_abs(nv, a.max_vel);          // :299-300
      double sn, cn;
      sincos(nth, &sn, &cn);
      double nx = (double)me.x + nv * cn;                           // :303-306
      double ny = (double)me.y - nv * sn;
      if (!TORUS) reflect_from_walls(nx, ny, nth, (double)me.z, a.width_d, a.height_d, a.pad_d);
      else teleport_torus(nx, ny, (double)me.z, a.width_d, a.height_d, a.pad_d);
      const float4 rec_new = make_float4((float)nx, (float)ny, me.z, me.w);
      a.rec_out[gi] = rec_new;
      if (a.n_peers > 0) {                                            // NVLink peer stores (fused tile exchange)
        for (int p = 0; p < a.n_peers; ++p) a.peer_rec_out[p][gi] = rec_new;
      }
      a.theta[gi] = (float)nth;
      a.vel[gi] = (float)nv;
      if (a.terms_out) {
        const size_t oi = a.perm ? (size_t)b * a.N + a.perm[gi] : (size_t)b * a.tile_count + li;
        double* t = a.terms_out + oi * 6;
        t[0] = ft.dvel; t[1] = ft.dpsi; t[2] = ft.a_blob; t[3] = ft.a_edge; t[4] = ft.b_blob; t[5] = ft.b_edge;
      }
    }
    if (a.fields_out) {
      const size_t oi = a.perm ? (size_t)b * a.N + a.perm[gi] : (size_t)b * a.tile_count + li;
      uint32_t* out = a.fields_out + oi * W;
      for (int ws = lane; ws < W; ws += 32) out[ws] = flipped_word(row, 1, R, W, ws);
    }
  }

}

template <bool TORUS, bool CULL>
static void launch_warp_variant(const VFKernelArgs& a, cudaStream_t stream) {
  const int per_rep = (a.tile_count + kWarpsPerCta - 1) / kWarpsPerCta;
  const size_t smem = vf_warp_smem_bytes(a.W);
  vf_step_warp_kernel<TORUS, CULL><<<(unsigned)((size_t)a.B * per_rep), kWarpsPerCta * 32, smem, stream>>>(a);
}

void launch_vf_step_warp(const VFKernelArgs& a, bool cull, cudaStream_t stream) {
  if (a.boundary == 1) { if (cull) launch_warp_variant<true, true>(a, stream); else launch_warp_variant<true, false>(a, stream); }
  else { if (cull) launch_warp_variant<false, true>(a, stream); else launch_warp_variant<false, false>(a, stream); }
}

}  // namespace abm
