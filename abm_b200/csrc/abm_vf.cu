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
#include <climits>
#include "abm_vf_device.cuh"

namespace abm {

size_t vf_step_smem_bytes(int threads, int W) {
  return 2 * sizeof(float4) * kRecTile            // record stages
         + sizeof(uint32_t) * (size_t)(W + 2) * threads   // padded rows
         + 2 * sizeof(uint32_t) * kQueueCap         // deferred-pair queue
         + 64                                       // mbarriers + queue counter
         + sizeof(int) * (kMaxTileList + 4) + 64;   // tile list of the culling variants, focal bounding box
}

// Focal agents per CTA: 256, or fewer when that would leave the GPU short of CTAs (one large swarm, or its tile on one
// of several GPUs: 65 536 agents are only 256 CTAs of 256) -- at least ~3 CTAs per SM are wanted.
int vf_step_threads(int tile_count, int n_replicates, int n_sms) {
  int t = (tile_count + 31) / 32 * 32;
  if (t > kMaxThreads) t = kMaxThreads;
  while (t > 64 && (long long)n_replicates * ((tile_count + t - 1) / t) < 3LL * n_sms) t >>= 1;
  return (t + 31) / 32 * 32;
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
template <bool TORUS, bool UNIFORM_R, bool CULL, bool FULL_FOV, int RC>
__global__ void __launch_bounds__(kMaxThreads, 3)
vf_step_kernel(const __grid_constant__ VFKernelArgs a) {
  using K = PairK<RC>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* recs = reinterpret_cast<float4*>(smem_raw);                       // [2][kRecTile]
  uint32_t* rows = reinterpret_cast<uint32_t*>(recs + 2 * kRecTile);        // [W + 2][T], padded (vf_draw_fast)
  const int T = blockDim.x;
  uint32_t* queue = rows + (size_t)(a.W + 2) * T;                           // [kQueueCap][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(queue + 2 * kQueueCap);      // [2]
  int* qcount = reinterpret_cast<int*>(bars + 2);
  int* tile_list = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(bars) + 64);   // [kMaxTileList] + count
  float* fbox = reinterpret_cast<float*>(tile_list + kMaxTileList + 4);                     // focal bbox [4]

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
  if (a.n_peers > 0 && tid <= a.n_peers) {
    // fused tile exchange: wait until every rank (thread r watches rank r, this rank included) has published the
    // previous step -- its records have landed in our table and nobody reads the table this launch writes into any more
    const uint32_t* f = a.xflags + tid;
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    } while ((int)(v - a.step_no) < 0);
  }
  __syncthreads();

  uint32_t* padrow = rows + tid;        // padded word 0 (virtual bins [-32, 0))
  uint32_t* myrow = rows + T + tid;     // real word 0
  const int R = RC ? RC : a.R;
  unsigned char* padrow_b = reinterpret_cast<unsigned char*>(padrow);
  const int stride_b = 4 * T;
  const uint32_t row_a = smem_u32(padrow);   // shared-space address of padded word 0
  const int fov0p = a.fov0p;           // first visible padded position
  const unsigned span = a.span;        // number of visible positions (vf_supcalc.py:119)
  const float width = a.width, height = a.height, half_w = a.half_w, half_h = a.half_h;
  unsigned n_mismatch = 0;

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
  // ---- which record tiles to visit: all of them, or (culling variants on spatially sorted state) only those
  //      whose bounding box comes within the cull distance of the bounding box of this CTA's focal agents ----
  const int n_tiles = (a.N + kRecTile - 1) / kRecTile;
  int n_stage = n_tiles;
  const bool use_list = CULL && a.tile_bbox != nullptr && n_tiles <= kMaxTileList;
  if (use_list) {
    if (tid < 4) fbox[tid] = (tid < 2) ? 3.0e38f : -3.0e38f;
    if (tid == 0) tile_list[kMaxTileList] = 0;
    __syncthreads();
    float x0 = active ? me.x : 3.0e38f, y0 = active ? me.y : 3.0e38f;
    float x1 = active ? me.x : -3.0e38f, y1 = active ? me.y : -3.0e38f;
    for (int off = 16; off > 0; off >>= 1) {
      x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, off)); y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, off));
      x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, off)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, off));
    }
    if ((tid & 31) == 0) {   // positive floats order like their bit patterns; negatives are clamped by the atomics' int view
      atomicMin(reinterpret_cast<int*>(&fbox[0]), __float_as_int(fmaxf(x0, 0.0f)));
      atomicMin(reinterpret_cast<int*>(&fbox[1]), __float_as_int(fmaxf(y0, 0.0f)));
      atomicMax(reinterpret_cast<int*>(&fbox[2]), __float_as_int(fmaxf(x1, 0.0f)));
      atomicMax(reinterpret_cast<int*>(&fbox[3]), __float_as_int(fmaxf(y1, 0.0f)));
    }
    __syncthreads();
    const float fx0 = fbox[0], fy0 = fbox[1], fx1 = fbox[2], fy1 = fbox[3];
    const float4* bb = a.tile_bbox + (size_t)b * n_tiles;
    const float* c2 = a.tile_cull2 + (size_t)b * n_tiles;
    for (int t = tid; t < n_tiles; t += T) {
      const float4 q = bb[t];
      float gx = fmaxf(0.0f, fmaxf(q.x - fx1, fx0 - q.z));
      float gy = fmaxf(0.0f, fmaxf(q.y - fy1, fy0 - q.w));
      if (TORUS) {   // minimal image: the tile shifted by one period either way
        gx = fminf(gx, fmaxf(0.0f, fmaxf(q.x + width - fx1, fx0 - (q.z + width))));
        gx = fminf(gx, fmaxf(0.0f, fmaxf(q.x - width - fx1, fx0 - (q.z - width))));
        gy = fminf(gy, fmaxf(0.0f, fmaxf(q.y + height - fy1, fy0 - (q.w + height))));
        gy = fminf(gy, fmaxf(0.0f, fmaxf(q.y - height - fy1, fy0 - (q.w - height))));
      }
      const float reach = sqrtf(c2[t]) + a.bbox_slack + 1.0f;
      if (gx * gx + gy * gy <= reach * reach) tile_list[atomicAdd(&tile_list[kMaxTileList], 1)] = t;
    }
    __syncthreads();
    n_stage = tile_list[kMaxTileList];
  }
  auto tile_of = [&](int st) { return use_list ? tile_list[st] : st; };
  if (tid == 0 && n_stage > 0) {
    const int t0 = tile_of(0);
    const int n0 = min(kRecTile, a.N - t0 * kRecTile);
    mbar_expect_tx(&bars[0], n0 * (uint32_t)sizeof(float4));
    tma_load_1d(recs, rep_in + (size_t)t0 * kRecTile, n0 * (uint32_t)sizeof(float4), &bars[0]);
  }

  for (int st = 0; st < n_stage; ++st) {
    if (tid == 0 && st + 1 < n_stage) {
      const int t1 = tile_of(st + 1);
      const int n1 = min(kRecTile, a.N - t1 * kRecTile);
      uint64_t* bar = &bars[(st + 1) & 1];
      mbar_expect_tx(bar, n1 * (uint32_t)sizeof(float4));
      tma_load_1d(recs + ((st + 1) & 1) * kRecTile, rep_in + (size_t)t1 * kRecTile, n1 * (uint32_t)sizeof(float4), bar);
    }
    mbar_wait(&bars[st & 1], (st >> 1) & 1);
    const int tile_j0 = tile_of(st) * kRecTile;            // first record index of this stage
    const float4* tile_recs = recs + (st & 1) * kRecTile;
    const int nj = min(kRecTile, a.N - tile_j0);
    if (active) {
      const uint32_t rec0 = smem_u32(tile_recs), rec_end = rec0 + 16u * (uint32_t)nj;
#pragma unroll 2
      for (uint32_t ra = rec0; ra < rec_end; ra += 16u) {
        const float4 o = lds_f4(ra);                         // broadcast LDS.128
        float dx, dy;
        if (UNIFORM_R) {
          dx = o.x - me.x; dy = o.y - me.y;
        } else {   // positions first (exact for close neighbours), then radii
          const float dr = o.z - me.z;
          dx = (o.x - me.x) + dr; dy = (o.y - me.y) + dr;
        }
        bool wrap_tie = false;
        if (TORUS) {                                         // vf_supcalc.py:70-83
          const float dr = UNIFORM_R ? 0.0f : o.z - me.z;
          dx = torus_delta_r(o.x, me.x, dr, width, half_w, wrap_tie);
          dy = torus_delta_r(o.y, me.y, dr, height, half_h, wrap_tie);
        }
        const float d2 = fmaf(dx, dx, dy * dy);
        if (CULL) { if (d2 > o.w) continue; }                // o.w: beyond it the half width is 0
        if (!UNIFORM_R) { if ((o.x == me.x) & (o.y == me.y)) continue; }   // vf_supcalc.py:57
        // ---- closed angle in BINS: rotate (dx, -dy) by -theta, octant-reduced polynomial atan2 whose
        //      coefficients carry the 1/step factor (vf_supcalc.py:86, :102) ----
        const float u = fmaf(dx, c, dy * ns);
        const float w = fmaf(dx, ns, -(dy * c));
        const float au = fabsf(u), aw = fabsf(w);
        const float tq = fminf(au, aw) * rcp_approx(fmaxf(au, aw));
        const float z = tq * tq;
        float p = fmaf(K::ac(a, 6), z, K::ac(a, 5));
        p = fmaf(p, z, K::ac(a, 4));
        p = fmaf(p, z, K::ac(a, 3));
        p = fmaf(p, z, K::ac(a, 2));
        p = fmaf(p, z, K::ac(a, 1));
        p = fmaf(p, z, K::ac(a, 0));
        p *= tq;
        if (aw > au) p = K::half_pi_b(a) - p;
        if (u < 0.0f) p = K::pi_b(a) - p;
        const float cab = copysignf(p, w);
        const float MAGIC = 12582912.0f;                     // 1.5 * 2^23
        const float tk = cab + K::t_half(a);
        const float tr = tk + MAGIC;                         // rint(tk) in the low mantissa bits
        const int k = __float_as_int(tr) + K::k_bias(a);     // centre bin, padded position
        const float dfk = tk - (tr - MAGIC);
        bool flagged = (fabsf(dfk) > a.thr_k) | (fabsf(cab) > a.seam_b) | wrap_tie;
        // ---- half width h = floor(atan(r / d) * R / 2pi) (vf_supcalc.py:96-99, :114-117) ----
        const float q = o.z * rsqrt_approx(d2);
        flagged |= !(q <= 1.0f);                             // also d2 == 0 (inf / NaN)
        const float y = fmaf(atan_unit(q), K::y_scale(a), -0.5f);
        const float yr = y + MAGIC;
        const int h = __float_as_int(yr) - 0x4B400000;
        flagged |= fmaf(y, a.nthr_h1, fabsf(y - (yr - MAGIC))) > a.thr_h0;
        const int ps = k - h, pe = k + h;
        const bool vis = FULL_FOV ? true : (((unsigned)(ps - fov0p) < span) | ((unsigned)(pe - fov0p) < span));
        if (!flagged & vis & ((unsigned)(h - 1) < 16u)) {
          // interval of <= 32 bins: at most two words of the private row (bank == lane)
          const uint32_t t = 0xffffffffu >> (32 - 2 * h);
          const uint32_t wa = row_a + (uint32_t)(ps >> 5) * (uint32_t)stride_b;
          const uint32_t lo = __funnelshift_l(0u, t, ps);
          const uint32_t hi = __funnelshift_l(t, 0u, ps);
          sts_u32(wa, lds_u32(wa) | lo);
          if (hi) sts_u32(wa + stride_b, lds_u32(wa + stride_b) | hi);
        } else if (flagged) {
          // deferred to fp64 (self / exactly coincident positions are skipped: vf_supcalc.py:57)
          if (!((o.x == me.x) & (o.y == me.y))) {
            const int slot = atomicAdd(qcount, 1);           // keeps counting past the capacity
            if (slot < kQueueCap) {
              queue[2 * slot] = ((uint32_t)tid << 24) | (uint32_t)(tile_j0 + (int)((ra - rec0) >> 4));
              queue[2 * slot + 1] = ((uint32_t)(k - 32) << 16) | ((uint32_t)h & 0xffffu);
            } else {
              vf_exact_inline(a, myrow, T, me, (size_t)b * a.N + i, o, k - 32, h);
            }
          }
        } else if (vis & (h > 16)) {
          if ((ps >= 0) & (pe <= R + 62)) vf_draw_wide(padrow_b, stride_b, ps, pe);
          else vf_draw_general(a, myrow, T, k - 32, h);
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

  {
    const int ii[1] = {active ? i : a.tile_begin}, lli[1] = {active ? li : 0};
    uint32_t* const pr[1] = {padrow};
    const float4 mm[1] = {me};
    const float tt[1] = {th};
    const bool aa[1] = {active};
    vf_agent_epilogue<TORUS, 1>(a, b, ii, lli, pr, T, mm, tt, 0u, aa);
  }
}

template <bool TORUS, bool UNIFORM_R, bool CULL, bool FULL_FOV, int RC>
static void launch_variant(const VFKernelArgs& a, unsigned grid, int T, size_t smem, cudaStream_t stream) {
  static SmemOptIn optin;   // per device (abm_common.cuh)
  optin.ensure(vf_step_kernel<TORUS, UNIFORM_R, CULL, FULL_FOV, RC>, smem);
  vf_step_kernel<TORUS, UNIFORM_R, CULL, FULL_FOV, RC><<<grid, T, smem, stream>>>(a);
}

// R = 1200 (the resolution of 99 of the reference's 108 experiment files) with full FOV gets a
// kernel with compile-time bin constants; everything else runs the run-time-R variants.
template <bool TORUS, bool UNIFORM_R, bool CULL>
static void launch_fov(const VFKernelArgs& a, unsigned grid, int T, size_t smem, cudaStream_t stream) {
  if (a.full_fov && a.R == 1200) launch_variant<TORUS, UNIFORM_R, CULL, true, 1200>(a, grid, T, smem, stream);
  else if (a.full_fov) launch_variant<TORUS, UNIFORM_R, CULL, true, 0>(a, grid, T, smem, stream);
  else launch_variant<TORUS, UNIFORM_R, CULL, false, 0>(a, grid, T, smem, stream);
}

void launch_vf_step(const VFKernelArgs& a, bool uniform_r, bool cull, cudaStream_t stream) {
  const int T = vf_step_threads(a.tile_count, a.B, a.n_sms);
  const int tiles_per_rep = (a.tile_count + T - 1) / T;
  const size_t smem = vf_step_smem_bytes(T, a.W);
  const unsigned grid = (unsigned)((size_t)a.B * tiles_per_rep);
  const int v = (a.boundary == 1 ? 4 : 0) | (uniform_r ? 2 : 0) | (cull ? 1 : 0);
  switch (v) {
    case 0: launch_fov<false, false, false>(a, grid, T, smem, stream); break;
    case 1: launch_fov<false, false, true>(a, grid, T, smem, stream); break;
    case 2: launch_fov<false, true, false>(a, grid, T, smem, stream); break;
    case 3: launch_fov<false, true, true>(a, grid, T, smem, stream); break;
    case 4: launch_fov<true, false, false>(a, grid, T, smem, stream); break;
    case 5: launch_fov<true, false, true>(a, grid, T, smem, stream); break;
    case 6: launch_fov<true, true, false>(a, grid, T, smem, stream); break;
    default: launch_fov<true, true, true>(a, grid, T, smem, stream); break;
  }
}

// Fused tile exchange: publish the finished step to every rank.  Launched behind the step kernel on the same stream, so
// all of its records -- the peer stores included -- are in place (kernel boundary); one release store per rank.
__global__ void vf_publish_kernel(uint32_t* own_flags, int my_rank, uint32_t done, int n_peers, VFPeerFlags pf) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(own_flags + my_rank), "r"(done) : "memory");
    for (int p = 0; p < n_peers; ++p)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.p[p] + my_rank), "r"(done) : "memory");
  }
}
void launch_vf_publish(const VFKernelArgs& a, cudaStream_t stream) {
  VFPeerFlags pf;
  for (int p = 0; p < 7; ++p) pf.p[p] = a.peer_flags[p];
  vf_publish_kernel<<<1, 32, 0, stream>>>(a.xflags, a.my_rank, a.step_no + 1u, a.n_peers, pf);
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
    bool wrap_tie = false;
    if (a.boundary == 1) {
      dx = torus_delta_r(ox, a.fx, dr, a.width, a.half_w, wrap_tie);
      dy = torus_delta_r(oy, a.fy, dr, a.height, a.half_h, wrap_tie);
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
      if (pf.flagged | wrap_tie) {
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

// cs_supcalc.projection_field (cooperative_signaling/cs_agent/cs_supcalc.py:204-289): one thread per
// object, the reference's operation sequence in float64 on float64 inputs.  Differences from the VF
// function above: object centre = position + FOCAL radius (:242); visibility decided on the angle,
// fov0 <= angle <= fov1 (:260); projections wider than max_proj_size bins dropped (:268); after the
// flip, bins whose linspace angle is outside the FOV are cleared (:290-291, `keep` = their mask).
// The meter amplitudes (:283-284) scale whole rows and are applied by the host.
__global__ void cs_projection_kernel(const CSProjArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n_obj) return;
  uint32_t* row = a.rows + (size_t)j * a.W;
  uint32_t tmp[128];                          // W <= 128 (R <= 4096) for this entry point
  for (int w = 0; w < a.W; ++w) tmp[w] = 0u;
  const double ox = a.ox[j], oy = a.oy[j];
  if (!(ox == a.fx && oy == a.fy)) {                                          // :233
    const double cix = __dadd_rn(a.fx, a.fr), ciy = __dadd_rn(a.fy, a.fr);    // :224
    const double ex = __dadd_rn(a.fx, __dmul_rn(__dadd_rn(1.0, cos(a.ftheta)), a.fr));   // :227-228
    const double ey = __dadd_rn(a.fy, __dmul_rn(__dadd_rn(1.0, -sin(a.ftheta)), a.fr));
    const double v1x = __dadd_rn(ex, -cix), v1y = __dadd_rn(ey, -ciy);        // :231
    const double n1 = __dsqrt_rn(__dadd_rn(__dmul_rn(v1x, v1x), __dmul_rn(v1y, v1y)));
    const double u1x = __ddiv_rn(v1x, n1), u1y = __ddiv_rn(v1y, n1);
    const double v2x = __dadd_rn(__dadd_rn(ox, a.fr), -cix);                  // :242, :245
    const double v2y = __dadd_rn(__dadd_rn(oy, a.fr), -ciy);
    const double n2 = __dsqrt_rn(__dadd_rn(__dmul_rn(v2x, v2x), __dmul_rn(v2y, v2y)));
    const double u2x = __ddiv_rn(v2x, n2), u2y = __ddiv_rn(v2y, n2);
    double dot = __dadd_rn(__dmul_rn(u1x, u2x), __dmul_rn(u1y, u2y));
    dot = fmin(1.0, fmax(-1.0, dot));
    double ang = acos(dot);                                                   // supcalc.py:31
    if (__dadd_rn(__dmul_rn(u1x, u2y), -__dmul_rn(u1y, u2x)) < 0.0) ang = -ang;
    if (ang < 0.0) ang = __dadd_rn(ang, ABM_TWO_PI_D);                        // :317
    const double ca = (ang >= 0.0 && ang <= ABM_PI_D) ? -ang : __dadd_rn(ABM_TWO_PI_D, -ang);   // :321-324
    const double vis = __dmul_rn(2.0, atan(__ddiv_rn(a.fr, n2)));             // :253
    const double proj = __dmul_rn(__ddiv_rn(vis, ABM_TWO_PI_D), (double)a.R); // :265
    const bool in_fov = (a.fov0 <= ca) && (ca <= a.fov1);                     // :260 (false for NaN)
    const bool size_ok = (a.max_proj_size < 0.0) || (proj <= a.max_proj_size);   // :286-296
    if (in_fov && size_ok && n2 > 0.0) {
      const int k = nearest_bin_exact(ca, a.R, a.lin_step);                   // :256
      const int h = (int)floor(__ddiv_rn(proj, 2.0));                         // :271-272
      vf_draw<false>(tmp, 1, a.R, INT_MIN, INT_MAX, k, h);                    // :274-281 (no test on the ends)
    }
  }
  for (int ws = 0; ws < a.W; ++ws) row[ws] = flipped_word(tmp, 1, a.R, a.W, ws) & a.keep[ws];   // :288-291
}

void launch_cs_projection(const CSProjArgs& a, cudaStream_t stream) {
  const int threads = 64;
  cs_projection_kernel<<<(a.n_obj + threads - 1) / threads, threads, 0, stream>>>(a);
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

// Bounding box and largest cull^2 of every tile of TILE records (culling variants, every step): one warp per 32
// records, one CTA per tile.
template <int TILE>
__global__ void __launch_bounds__(TILE) vf_tile_bbox_kernel(const float4* rec, int N, int n_tiles, float4* bbox,
                                                             float* cull2) {
  __shared__ float red[5][TILE / 32];
  const int bt = blockIdx.x, b = bt / n_tiles, t = bt - b * n_tiles;
  const int j = t * TILE + threadIdx.x;
  float x0 = 3.0e38f, y0 = 3.0e38f, x1 = -3.0e38f, y1 = -3.0e38f, c = 0.0f;
  if (j < N) {
    const float4 v = rec[(size_t)b * N + j];
    x0 = x1 = v.x; y0 = y1 = v.y; c = v.w;
  }
  for (int off = 16; off > 0; off >>= 1) {
    x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, off)); y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, off));
    x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, off)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, off));
    c = fmaxf(c, __shfl_xor_sync(0xffffffffu, c, off));
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][w] = x0; red[1][w] = y0; red[2][w] = x1; red[3][w] = y1; red[4][w] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < TILE / 32; ++k) {
      x0 = fminf(x0, red[0][k]); y0 = fminf(y0, red[1][k]); x1 = fmaxf(x1, red[2][k]); y1 = fmaxf(y1, red[3][k]);
      c = fmaxf(c, red[4][k]);
    }
    bbox[bt] = make_float4(x0, y0, x1, y1);
    cull2[bt] = c;
  }
}
void launch_tile_bbox(const float4* rec, int B, int N, int tile, float4* bbox, float* cull2, cudaStream_t stream) {
  const int n_tiles = (N + tile - 1) / tile;
  if (tile == kWarpTile) vf_tile_bbox_kernel<kWarpTile><<<(unsigned)((size_t)B * n_tiles), kWarpTile, 0, stream>>>(rec, N, n_tiles, bbox, cull2);
  else vf_tile_bbox_kernel<kRecTile><<<(unsigned)((size_t)B * n_tiles), kRecTile, 0, stream>>>(rec, N, n_tiles, bbox, cull2);
}

// SoA host-facing state <-> packed neighbour records
__global__ void pack_records_kernel(const float* x, const float* y, const float* r, const int* perm, int N,
                                    float cull_scale, float4* rec, unsigned* radius_minmax, long long n) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float rr = 0.f;
  if (g < n) {
    const long long o = perm ? (g - (g % N)) + perm[g] : g;   // caller's order -> internal slot g
    rr = r[o];
    rec[g] = make_float4(x[o], y[o], rr, rr * rr * cull_scale);
  }
  if (radius_minmax == nullptr) return;   // the radii did not change (abm_set_state with radius == NULL)
  // min / max radius of the batch (non-negative floats order like their bit patterns)
  unsigned lo = (g < n) ? __float_as_uint(fmaxf(rr, 0.f)) : 0x7f800000u;
  unsigned hi = (g < n) ? __float_as_uint(fmaxf(rr, 0.f)) : 0u;
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0) {   // (an atomic only where it would change the result: 65 000 warps on two words otherwise)
    if (lo < __ldcg(&radius_minmax[0])) atomicMin(&radius_minmax[0], lo);
    if (hi > __ldcg(&radius_minmax[1])) atomicMax(&radius_minmax[1], hi);
  }
}
// The same from ONE interleaved array of (x, y, heading, speed) per agent (abm_set_state_packed): record, heading and
// speed of internal slot g come from the caller's agent o.  radius_minmax == nullptr: the radii did not change.
__global__ void pack_state4_kernel(const float4* __restrict__ s4, const float* __restrict__ r, const int* perm, int N,
                                   float cull_scale, float4* __restrict__ rec, float* __restrict__ theta,
                                   float* __restrict__ vel, unsigned* radius_minmax, long long n) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float rr = 0.f;
  if (g < n) {
    const long long o = perm ? (g - (g % N)) + perm[g] : g;
    const float4 s = s4[o];
    rr = r[o];
    rec[g] = make_float4(s.x, s.y, rr, rr * rr * cull_scale);
    theta[g] = s.z;
    vel[g] = s.w;
  }
  if (radius_minmax == nullptr) return;
  unsigned lo = (g < n) ? __float_as_uint(fmaxf(rr, 0.f)) : 0x7f800000u;
  unsigned hi = (g < n) ? __float_as_uint(fmaxf(rr, 0.f)) : 0u;
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0) {
    if (lo < __ldcg(&radius_minmax[0])) atomicMin(&radius_minmax[0], lo);
    if (hi > __ldcg(&radius_minmax[1])) atomicMax(&radius_minmax[1], hi);
  }
}
__global__ void unpack_state4_kernel(const float4* __restrict__ rec, const float* __restrict__ theta,
                                     const float* __restrict__ vel, const int* perm, int N, float4* __restrict__ out,
                                     long long n) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) {
    const float4 v = rec[g];
    const long long o = perm ? (g - (g % N)) + perm[g] : g;
    out[o] = make_float4(v.x, v.y, theta[g], vel[g]);
  }
}
void launch_pack_state4(const float4* s4, const float* r, const int* perm, int N, float cull_scale, float4* rec, float* theta,
                        float* vel, unsigned* radius_minmax, long long n, cudaStream_t stream) {
  const int threads = 256;
  pack_state4_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(s4, r, perm, N, cull_scale, rec, theta,
                                                                                     vel, radius_minmax, n);
}
void launch_unpack_state4(const float4* rec, const float* theta, const float* vel, const int* perm, int N, float4* out,
                          long long n, cudaStream_t stream) {
  const int threads = 256;
  unpack_state4_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(rec, theta, vel, perm, N, out, n);
}
__global__ void unpack_records_kernel(const float4* rec, const int* perm, int N, float* x, float* y, long long n) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) {
    const float4 v = rec[g];
    const long long o = perm ? (g - (g % N)) + perm[g] : g;
    if (x) x[o] = v.x;
    if (y) y[o] = v.y;
  }
}
void launch_pack_records(const float* x, const float* y, const float* r, const int* perm, int N, float cull_scale,
                         float4* rec, unsigned* radius_minmax, long long n, cudaStream_t stream) {
  const int threads = 256;
  pack_records_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(x, y, r, perm, N, cull_scale, rec,
                                                                                      radius_minmax, n);
}
void launch_unpack_records(const float4* rec, const int* perm, int N, float* x, float* y, long long n,
                           cudaStream_t stream) {
  const int threads = 256;
  unpack_records_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(rec, perm, N, x, y, n);
}

}  // namespace abm
