// Fused visual-flocking step kernel for sm_100a.
//
// One launch advances every agent of every replicate by one synchronous step:
//   neighbour records staged in shared memory by 1-D bulk TMA (cp.async.bulk + mbarrier,
//   double buffered) -> per-pair interval in fp32 -> private bit-packed row in shared
//   memory (word w of thread t at [w][t]: bank == lane, no conflicts, no atomics) ->
//   deferred fp64 re-evaluation of the pairs whose bin index is within the fp32 error
//   bound of a rounding boundary -> edges, flocking integrals, kinematics, walls / torus
//   -> new record + heading + speed (+ optional packed field / terms dump).
//
// Mapping: CTA = (replicate, tile of <= 256 focal agents), one thread per focal agent; all
// lanes of a warp read the same neighbour record (shared-memory broadcast, LDS.128).
// Replaces VFSimulation.step_sim -> VFAgent.update for all agents (vf_sims.py:291-302,
// vf_agent.py:52-80).
#include "abm_vf_device.cuh"

namespace abm {

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

size_t vf_step_smem_bytes(int threads, int W) {
  return 2 * sizeof(float4) * kRecTile            // record stages
         + sizeof(uint32_t) * (size_t)(W + 2) * threads   // padded rows
         + 2 * sizeof(uint32_t) * kQueueCap         // deferred-pair queue
         + 64;                                      // mbarriers + queue counter
}

int vf_step_threads(int tile_count) {
  int t = (tile_count + 31) / 32 * 32;
  return t > kMaxThreads ? kMaxThreads : t;
}

// Out-of-line fp64 evaluation + atomic draw of one pair (queue overflow / deferred pairs).
// Returns 1 if the fp64 indices differ from the fp32 ones (k32, h32).
static __device__ __noinline__ unsigned vf_exact_and_draw(const VFKernelArgs& a, uint32_t* row, int stride,
                                                   float4 f4, float fth, float4 o, int k32, int h32) {
  const FocalExact fe = vf_focal_exact(f4.x, f4.y, f4.z, fth);
  const PairExact pe = vf_pair_exact(fe, o.x, o.y, o.z, a.boundary, a.width_d, a.height_d, a.R, a.lin_step);
  if (pe.valid) vf_draw<true>(row, stride, a.R, a.fov_px0, a.fov_px1, pe.k, pe.h);
  return (pe.valid && ((pe.k != k32) | (pe.h != h32))) ? 1u : 0u;
}

// Queue overflow: evaluate a flagged pair on the spot (out of line: rare).
static __device__ __noinline__ void vf_exact_inline(const VFKernelArgs& a, uint32_t* myrow, int stride, float4 me, size_t gi,
                                             float4 o, int k, int h) {
  const unsigned diff = vf_exact_and_draw(a, myrow, stride, me, a.theta[gi], o, k, h);
  atomicAdd(&a.counters[1], 1ull);
  if (diff) atomicAdd(&a.counters[2], 1ull);
}

// Interval that would leave the row padding (h > 16 right at the seam): general rule, out of line.
static __device__ __noinline__ void vf_draw_general(const VFKernelArgs& a, uint32_t* myrow, int stride, int k, int h) {
  vf_draw<false>(myrow, stride, a.R, a.fov_px0, a.fov_px1, k, h);
}

// TORUS: BOUNDARY == infinite.  UNIFORM_R: all radii equal (centre difference == position
// difference).  CULL: skip pairs beyond the distance at which the half width becomes 0.
template <bool TORUS, bool UNIFORM_R, bool CULL>
__global__ void __launch_bounds__(kMaxThreads, 3)
vf_step_kernel(const VFKernelArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* recs = reinterpret_cast<float4*>(smem_raw);                       // [2][kRecTile]
  uint32_t* rows = reinterpret_cast<uint32_t*>(recs + 2 * kRecTile);        // [W + 2][T], padded (vf_draw_fast)
  const int T = blockDim.x;
  uint32_t* queue = rows + (size_t)(a.W + 2) * T;                           // [kQueueCap][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(queue + 2 * kQueueCap);      // [2]
  int* qcount = reinterpret_cast<int*>(bars + 2);

  const int tid = threadIdx.x;
  const int tiles_per_rep = (a.tile_count + T - 1) / T;
  const int b = blockIdx.x / tiles_per_rep;
  const int tile = blockIdx.x - b * tiles_per_rep;
  const int li = tile * T + tid;                  // index inside this engine's focal tile
  const bool active = li < a.tile_count;
  const int i = a.tile_begin + li;                // agent index inside the replicate
  const float4* rep_in = a.rec_in + (size_t)b * a.N;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    *qcount = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int w = 0; w < a.W + 2; ++w) rows[w * T + tid] = 0u;
  __syncthreads();

  const int n_stage = (a.N + kRecTile - 1) / kRecTile;
  if (tid == 0) {
    const int n0 = min(kRecTile, a.N);
    mbar_expect_tx(&bars[0], n0 * (uint32_t)sizeof(float4));
    tma_load_1d(recs, rep_in, n0 * (uint32_t)sizeof(float4), &bars[0]);
  }

  float4 me = make_float4(0.f, 0.f, 1.f, 0.f);
  float th = 0.f;
  if (active) {
    me = rep_in[i];
    th = a.theta[(size_t)b * a.N + i];
  }
  float c, ns;
  {
    double sd, cd;
    sincos((double)th, &sd, &cd);
    c = (float)cd; ns = (float)(-sd);
  }
  const BinConsts bc{a.inv_step, a.t_half, a.k_bias, a.y_scale, a.thr_k, a.thr_h0, a.thr_h1, a.ca_guard};
  uint32_t* padrow = rows + tid;        // padded word 0 (virtual bins [-32, 0))
  uint32_t* myrow = rows + T + tid;     // real word 0
  const int R = a.R;
  unsigned char* padrow_b = reinterpret_cast<unsigned char*>(padrow);
  const int stride_b = 4 * T;
  const int fov0p = a.fov0p;           // first visible padded position
  const unsigned span = a.span;        // number of visible positions (vf_supcalc.py:119)
  const float width = a.width, height = a.height, half_w = a.half_w, half_h = a.half_h;
  unsigned n_mismatch = 0;

  for (int st = 0; st < n_stage; ++st) {
    if (tid == 0 && st + 1 < n_stage) {
      const int n1 = min(kRecTile, a.N - (st + 1) * kRecTile);
      uint64_t* bar = &bars[(st + 1) & 1];
      mbar_expect_tx(bar, n1 * (uint32_t)sizeof(float4));
      tma_load_1d(recs + ((st + 1) & 1) * kRecTile, rep_in + (size_t)(st + 1) * kRecTile,
                  n1 * (uint32_t)sizeof(float4), bar);
    }
    mbar_wait(&bars[st & 1], (st >> 1) & 1);
    const float4* tile_recs = recs + (st & 1) * kRecTile;
    const int nj = min(kRecTile, a.N - st * kRecTile);
    if (active) {
      const float4* rend = tile_recs + nj;
#pragma unroll 2
      for (const float4* rp = tile_recs; rp < rend; ++rp) {
        const float4 o = *rp;                                // broadcast LDS.128
        float dx, dy;
        if (UNIFORM_R) {
          dx = o.x - me.x; dy = o.y - me.y;
        } else {   // positions first (exact for close neighbours), then radii
          const float dr = o.z - me.z;
          dx = (o.x - me.x) + dr; dy = (o.y - me.y) + dr;
        }
        if (TORUS) {                                         // vf_supcalc.py:70-83
          if (fabsf(dx) > half_w) dx -= copysignf(width, dx);
          if (fabsf(dy) > half_h) dy -= copysignf(height, dy);
        }
        const float d2 = fmaf(dx, dx, dy * dy);
        if (CULL) { if (d2 > o.w) continue; }                // o.w: beyond it the half width is 0
        if (!UNIFORM_R) { if ((o.x == me.x) & (o.y == me.y)) continue; }   // vf_supcalc.py:57
        const PairFast pf = vf_pair_fast(dx, dy, d2, o.z, c, ns, bc);
        const int ps = pf.k - pf.h, pe = pf.k + pf.h;        // padded positions
        const bool vis = ((unsigned)(ps - fov0p) < span) | ((unsigned)(pe - fov0p) < span);   // vf_supcalc.py:119
        if (!pf.flagged & vis & ((unsigned)(pf.h - 1) < 16u)) {
          vf_draw_short(padrow_b, stride_b, ps, pf.h);
        } else if (pf.flagged) {
          // deferred to fp64 (self / exactly coincident positions are skipped: vf_supcalc.py:57)
          if (!((o.x == me.x) & (o.y == me.y))) {
            const int slot = atomicAdd(qcount, 1);           // keeps counting past the capacity
            if (slot < kQueueCap) {
              queue[2 * slot] = ((uint32_t)tid << 24) | (uint32_t)(st * kRecTile + (int)(rp - tile_recs));
              queue[2 * slot + 1] = ((uint32_t)(pf.k - 32) << 16) | ((uint32_t)pf.h & 0xffffu);
            } else {
              vf_exact_inline(a, myrow, T, me, (size_t)b * a.N + i, o, pf.k - 32, pf.h);
            }
          }
        } else if (vis & (pf.h > 16)) {
          if ((ps >= 0) & (pe <= R + 62)) vf_draw_wide(padrow_b, stride_b, ps, pe);
          else vf_draw_general(a, myrow, T, pf.k - 32, pf.h);
        }
      }
    }
    __syncthreads();   // everyone is done with this stage before it is refilled
  }

  // ---- deferred pairs: fp64, the reference's own operation sequence ----
  {
    const int nq = min(*qcount, kQueueCap);
    for (int e = tid; e < nq; e += T) {
      const uint32_t q0 = queue[2 * e], q1 = queue[2 * e + 1];
      const int ft = (int)(q0 >> 24);
      const int j = (int)(q0 & 0xffffffu);
      const int fi = a.tile_begin + tile * T + ft;
      n_mismatch += vf_exact_and_draw(a, rows + T + ft, T, rep_in[fi], a.theta[(size_t)b * a.N + fi], rep_in[j],
                                      (int)(q1 >> 16), (int)(short)(q1 & 0xffffu));
    }
  }
  __syncthreads();

  // ---- counters (one atomic per warp) ----
  {
    const unsigned nm = __reduce_add_sync(0xffffffffu, n_mismatch);
    if ((tid & 31) == 0 && nm) atomicAdd(&a.counters[2], (unsigned long long)nm);
    if (tid == 0 && *qcount) atomicAdd(&a.counters[0], (unsigned long long)*qcount);   // all flagged pairs
  }

  if (!active) return;
  vf_fold_padding(padrow, T, a.R, a.W);

  // ---- epilogue: edges, integrals, kinematics (fp64) ----
  const size_t gi = (size_t)b * a.N + i;
  const VFParams6 prm = *reinterpret_cast<const VFParams6*>(a.params + (size_t)b * a.param_stride);
  double A0 = prm.alp0, B0 = prm.bet0, V0 = prm.v0;            // vf_supcalc.py:191-196
  if (a.ov_alp0) { const float v = a.ov_alp0[gi]; if (v == v) A0 = v; }
  if (a.ov_bet0) { const float v = a.ov_bet0[gi]; if (v == v) B0 = v; }
  if (a.ov_v0)   { const float v = a.ov_v0[gi];   if (v == v) V0 = v; }
  const double vel0 = a.vel[gi];
  FlockTerms ft;
  if (a.phi_ok) {
    ft = vf_flock_terms(myrow, T, a.R, a.W, a.lut, a.dphi, vel0, prm, A0, B0, V0);
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

  a.rec_out[gi] = make_float4((float)nx, (float)ny, me.z, me.w);
  a.theta[gi] = (float)nth;
  a.vel[gi] = (float)nv;

  const size_t oi = (size_t)b * a.tile_count + li;
  if (a.terms_out) {
    double* t = a.terms_out + oi * 6;
    t[0] = ft.dvel; t[1] = ft.dpsi; t[2] = ft.a_blob; t[3] = ft.a_edge; t[4] = ft.b_blob; t[5] = ft.b_edge;
  }
  if (a.fields_out) {
    uint32_t* out = a.fields_out + oi * a.W;
    for (int ws = 0; ws < a.W; ++ws) out[ws] = flipped_word(myrow, T, a.R, a.W, ws);
  }
}

template <bool TORUS, bool UNIFORM_R, bool CULL>
static void launch_variant(const VFKernelArgs& a, unsigned grid, int T, size_t smem, cudaStream_t stream) {
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(vf_step_kernel<TORUS, UNIFORM_R, CULL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
    configured = smem;
  }
  vf_step_kernel<TORUS, UNIFORM_R, CULL><<<grid, T, smem, stream>>>(a);
}

void launch_vf_step(const VFKernelArgs& a, bool uniform_r, bool cull, cudaStream_t stream) {
  const int T = vf_step_threads(a.tile_count);
  const int tiles_per_rep = (a.tile_count + T - 1) / T;
  const size_t smem = vf_step_smem_bytes(T, a.W);
  const unsigned grid = (unsigned)((size_t)a.B * tiles_per_rep);
  const int v = (a.boundary == 1 ? 4 : 0) | (uniform_r ? 2 : 0) | (cull ? 1 : 0);
  switch (v) {
    case 0: launch_variant<false, false, false>(a, grid, T, smem, stream); break;
    case 1: launch_variant<false, false, true>(a, grid, T, smem, stream); break;
    case 2: launch_variant<false, true, false>(a, grid, T, smem, stream); break;
    case 3: launch_variant<false, true, true>(a, grid, T, smem, stream); break;
    case 4: launch_variant<true, false, false>(a, grid, T, smem, stream); break;
    case 5: launch_variant<true, false, true>(a, grid, T, smem, stream); break;
    case 6: launch_variant<true, true, false>(a, grid, T, smem, stream); break;
    default: launch_variant<true, true, true>(a, grid, T, smem, stream); break;
  }
}

// ---------------------------------------------------------------------------------------
// stateless function-level kernels
// ---------------------------------------------------------------------------------------

// vf_supcalc.projection_field: one thread per object, private row in global memory.
__global__ void vf_projection_kernel(const VFProjArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n_obj) return;
  uint32_t* row = a.rows + (size_t)j * a.W;   // used un-flipped first, flipped in place at the end
  uint32_t tmp[128];                          // W <= 128 (R <= 4096) for this entry point
  for (int w = 0; w < a.W; ++w) tmp[w] = 0u;
  const float ox = a.ox[j], oy = a.oy[j];
  const float orad = a.osz ? a.osz[j] : a.fr;
  const bool same = (ox == a.fx) & (oy == a.fy);              // vf_supcalc.py:57
  if (!same) {
    const float dr = orad - a.fr;
    float dx = (ox - a.fx) + dr, dy = (oy - a.fy) + dr;
    if (a.boundary == 1) {
      if (fabsf(dx) > a.half_w) dx -= copysignf(a.width, dx);
      if (fabsf(dy) > a.half_h) dy -= copysignf(a.height, dy);
    }
    const float d2 = fmaf(dx, dx, dy * dy);
    bool in_range = true;
    if (a.vision_range >= 0.0) {                              // :91-93 (decided in fp64)
      const double ddx = dx, ddy = dy;
      in_range &= !(sqrt(ddx * ddx + ddy * ddy) > a.vision_range);
    }
    if (in_range) {
      double sd, cd;
      sincos((double)a.ftheta, &sd, &cd);
      const BinConsts bc{a.inv_step, a.t_half, a.k_bias, a.y_scale, a.thr_k, a.thr_h0, a.thr_h1, a.ca_guard};
      const PairFast pf = vf_pair_fast(dx, dy, d2, orad, (float)cd, (float)(-sd), bc);
      int k = pf.k - 32, h = pf.h;
      if (pf.flagged) {
        const FocalExact fe = vf_focal_exact(a.fx, a.fy, a.fr, a.ftheta);
        const PairExact pe = vf_pair_exact(fe, ox, oy, orad, a.boundary, a.width_d, a.height_d, a.R, a.lin_step);
        k = pe.k; h = pe.valid ? pe.h : 0;
      }
      vf_draw<false>(tmp, 1, a.R, a.fov_px0, a.fov_px1, k, h);
    }
  }
  for (int ws = 0; ws < a.W; ++ws) row[ws] = flipped_word(tmp, 1, a.R, a.W, ws);   // :134
}

void launch_vf_projection(const VFProjArgs& a, cudaStream_t stream) {
  const int threads = 64;
  vf_projection_kernel<<<(a.n_obj + threads - 1) / threads, threads, 0, stream>>>(a);
}

__global__ void vf_terms_kernel(const uint32_t* packed_v, int R, int W, double vel, const VFParams6* prm,
                                const PhiLut* lut, double dphi, double* out6) {
  const VFParams6 p = *prm;
  const FlockTerms t = vf_flock_terms(packed_v, 1, R, W, lut, dphi, vel, p, p.alp0, p.bet0, p.v0);
  out6[0] = t.dvel; out6[1] = t.dpsi; out6[2] = t.a_blob; out6[3] = t.a_edge; out6[4] = t.b_blob; out6[5] = t.b_edge;
}

void launch_vf_terms(const uint32_t* packed_v, int R, int W, double vel, const VFParams6* prm,
                     const PhiLut* lut, double dphi, double* out6, cudaStream_t stream) {
  vf_terms_kernel<<<1, 1, 0, stream>>>(packed_v, R, W, vel, prm, lut, dphi, out6);
}

// SoA host-facing state <-> packed neighbour records
__global__ void pack_records_kernel(const float* x, const float* y, const float* r, float cull_scale,
                                    float4* rec, unsigned* radius_minmax, long long n) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float rr = 0.f;
  if (g < n) {
    rr = r[g];
    rec[g] = make_float4(x[g], y[g], rr, rr * rr * cull_scale);
  }
  // min / max radius of the batch (non-negative floats order like their bit patterns)
  unsigned lo = (g < n) ? __float_as_uint(fmaxf(rr, 0.f)) : 0x7f800000u;
  unsigned hi = (g < n) ? __float_as_uint(fmaxf(rr, 0.f)) : 0u;
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&radius_minmax[0], lo);
    atomicMax(&radius_minmax[1], hi);
  }
}
__global__ void unpack_records_kernel(const float4* rec, float* x, float* y, long long n) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) {
    const float4 v = rec[g];
    if (x) x[g] = v.x;
    if (y) y[g] = v.y;
  }
}
void launch_pack_records(const float* x, const float* y, const float* r, float cull_scale, float4* rec,
                         unsigned* radius_minmax, long long n, cudaStream_t stream) {
  const int threads = 256;
  pack_records_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(x, y, r, cull_scale, rec,
                                                                                      radius_minmax, n);
}
void launch_unpack_records(const float4* rec, float* x, float* y, long long n, cudaStream_t stream) {
  const int threads = 256;
  unpack_records_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(rec, x, y, n);
}

}  // namespace abm
