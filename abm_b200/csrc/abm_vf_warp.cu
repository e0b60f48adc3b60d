// Visual-flocking step kernel for ONE LARGE, SPARSE swarm, for its agent tiles on several GPUs, and for batches too small
// to fill the GPU with a CTA per replicate (sm_100a).
//
// Mapping.  CTA = F consecutive focal agents of one replicate (spatially close when the engine keeps its Morton order)
// x 256 threads.  The CTA builds ONE list of record tiles to visit (bounding box of its focal agents against the tiles'
// boxes); its threads stride over the RECORDS of the listed tiles -- one coalesced 16-byte load per record from the
// L2-resident table -- and evaluate every loaded record against all F focal agents (focal constants broadcast from
// shared memory): F independent pair evaluations per load give the latency-bound loop its instruction-level
// parallelism, all 8 warps of the CTA carry the same load whatever the focal agents see, and the work of a focal
// agent in a dense region is spread over 256 threads instead of the 32 lanes of one warp.  F is chosen from the number
// of focal agents of the launch (8 for a whole 65 536-agent swarm; 2 for an eighth of it on each of 8 GPUs; 1 for one
// small run) so that the grid keeps every SM busy with several CTAs: the first version of this kernel gave each focal
// agent one warp and an eighth of the swarm took almost as long as the whole (the warps of the agents in the dense
// centre of the disc ran alone at the end).
// Pair arithmetic as in the symmetric kernel (32-bit binary angles, one IMAD.WIDE for bin + guard band; full-range
// arctangent for the half width); intervals are OR-ed into the focal agent's row in shared memory with RED.OR
// (abm_vf_sym.cu explains why those are cheap); guard-band hits are re-evaluated on the spot in fp64 (the reference's own
// operation sequence).  Epilogue: warp f takes focal agent f -- lanes take the words of the row, edge sums are reduced
// with shuffles, lane 0 finishes (terms, kinematics, walls / torus, outputs, peer stores of the fused tile exchange).
//
// Fused tile exchange (abm_vf_ipc_attach; one process per GPU): a launch starts when every rank has published the
// previous step (flags in this GPU's memory); the epilogue stores the new records into the peers' tables over NVLink;
// the LAST CTA of the launch (ticket counter) computes the bounding boxes of this rank's record tiles for the next step,
// stores them into every rank's box table, and publishes the step to all ranks (st.release.sys) -- one launch per step,
// no collective, and nothing reads a table a peer may still be writing.
#include "abm_vf_device.cuh"

namespace abm {

constexpr int kWarpThreads = 256;      // default threads per CTA
constexpr int kWarpMaxThreads = 512;   // F = 8 on a grid that would otherwise be a partial wave
constexpr int kWarpMaxFocal = 8;     // focal agents per CTA <= warps per CTA (a warp per focal agent in the epilogue)
constexpr int kChunk = 128;          // records per pass of half the CTA (= kWarpTile)
constexpr int kWarpQueue = 1024;     // pairs off the fast path waiting for the CTA's slow pass

// per focal agent of the CTA, in shared memory: (x, y, radius, heading constant) read with one LDS.128 per pair
struct WarpShared {
  float4* focal;       // [F]
  float* theta;        // [F] heading (fp64 path only)
  uint32_t* rows;      // [F][W + 3]: padded rows (word 0 = virtual bins [-32, 0), words 1 .. W the real bins, one word of
                       // overflow; abm_vf_device.cuh: vf_fold_padding) + the spill word of the second reduction
  int* tile_list;      // [kMaxTileList] + count
  int* fbox;           // focal bounding box (ordered ints)
  uint32_t* queue;     // [kWarpQueue] focal << 24 | record: pairs off the fast path; count at tile_list[kMaxTileList + 1]
};
size_t vf_warp_smem_bytes(int W, int F) {
  return sizeof(float4) * F + sizeof(float) * kWarpMaxFocal + sizeof(uint32_t) * (size_t)(W + 3) * F +
         sizeof(int) * (kMaxTileList + 4) + 64 + sizeof(uint32_t) * kWarpQueue;
}

// ordered-integer image of a float (monotonic for all finite values): min / max by integer atomics
__device__ __forceinline__ int float_ordered(float v) {
  const int b = __float_as_int(v);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// Off the fast path (wide interval, guard-band hit, coincident positions): the complete evaluation -- full-range
// arctangent for the half width, separate tie / seam tests, guard-band hits re-evaluated in fp64 with the reference's own
// operation sequence -- and the general drawing rule into the REAL words of the row.
template <bool TORUS>
static __device__ __noinline__ void warp_pair_slow(const VFKernelArgs& a, float4 f, float fth, uint32_t row_real_s,
                                                   float4 o, unsigned& n_fp64, unsigned& n_differ) {
  if ((o.x == f.x) & (o.y == f.y)) return;               // self / coincident positions (vf_supcalc.py:57)
  const int R = a.R;
  const float dr = o.z - f.z;
  float dx = (o.x - f.x) + dr, dy = (o.y - f.y) + dr;
  bool wrap_tie = false;
  if (TORUS) {
    dx = torus_delta_r(o.x, f.x, dr, a.width, a.half_w, wrap_tie);
    dy = torus_delta_r(o.y, f.y, dr, a.height, a.half_h, wrap_tie);
  }
  const float d2 = fmaf(dx, dx, dy * dy);
  const float q = o.z * rsqrt_approx(d2);
  const float y = fmaf(atan_unit(q), a.y_scale, -0.5f);
  const float yr = y + kMagic;
  int h = __float_as_int(yr) - kMagicBits;
  bool flagged = !(q <= 1.0f) | (fmaf(y, a.nthr_h1, fabsf(y - (yr - kMagic))) > a.thr_h0) | wrap_tie;
  int k = sym_side_k<0>(a, sym_bearing_bits(dx, dy, kBearingA6), __float_as_uint(f.w), 0, flagged);   // bin index
  if (flagged) {                                         // fp64, the reference's own operation sequence
    const FocalExact fe = vf_focal_exact(f.x, f.y, f.z, fth);
    const PairExact pe = vf_pair_exact(fe, o.x, o.y, o.z, a.boundary, a.width_d, a.height_d, R, a.lin_step);
    ++n_fp64;
    if (!pe.valid) return;
    if ((pe.k != k) | (pe.h != h)) ++n_differ;
    k = pe.k; h = pe.h;
  }
  vf_draw_shared(row_real_s, 4u, R, a.fov_px0, a.fov_px1, k, h);
}

// One (focal, record) pair on the fast path: interval of 1 .. 32 bins clear of every fp32 guard band, drawn into the
// padded row with two reductions (the arithmetic of the symmetric kernel's fast path, one direction).
template <bool TORUS, bool CULL, bool FULL_FOV, bool UNIFORM_R>
__device__ __forceinline__ void warp_pair(const VFKernelArgs& a, const float4 f, uint32_t row_pad_s, const float4 o,
                                          float S, float c3s, float c5s, float c7s, uint32_t qentry, uint32_t queue_s, uint32_t qcount_s,
                                          const WarpShared& sh, unsigned& n_fp64, unsigned& n_differ) {
  float dx, dy;
  bool slow = false;
  if (UNIFORM_R) {
    dx = o.x - f.x; dy = o.y - f.y;
    if (TORUS) { dx = torus_delta(o.x, f.x, a.width, a.half_w, slow); dy = torus_delta(o.y, f.y, a.height, a.half_h, slow); }
  } else {
    const float dr = o.z - f.z;
    dx = (o.x - f.x) + dr; dy = (o.y - f.y) + dr;        // positions first (exact for close neighbours), then radii
    if (TORUS) {
      dx = torus_delta_r(o.x, f.x, dr, a.width, a.half_w, slow);
      dy = torus_delta_r(o.y, f.y, dr, a.height, a.half_h, slow);
    }
  }
  const float d2 = fmaf(dx, dx, dy * dy);
  if (CULL) { if (d2 > o.w) return; }                    // beyond it the half width is 0
  slow |= (o.x == f.x) & (o.y == f.y);
  // half width h = floor(atan(r / d) R / 2pi): four-term series in q = r / d, exact to the guard band up to q = 0.18
  // (abm_vf_sym.cu); the limit on qs also keeps the interval within two row words (h <= 16)
  const float qs = rsqrt_approx(d2) * (o.z * S);
  const float zs = qs * qs;
  float p = fmaf(zs, c7s, c5s);
  p = fmaf(p, zs, c3s);
  p = fmaf(p, zs, 1.0f);
  const float y = fmaf(qs, p, -0.5f);
  const float yr = y + kMagic;
  const uint32_t hraw = __float_as_uint(yr);             // h + kMagicBits
  slow |= !(qs < a.warp_qs_max) | (fabsf(y - (yr - kMagic)) > a.sym_thr_h);   // wide, near an integer, inf / NaN
  uint32_t mask;
  asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(mask) : "r"(2u * hraw - 2u * (uint32_t)kMagicBits));   // 2h ones
  const int bh = 32 + kMagicBits - (int)hraw;            // ps (padded) = bin index + 32 - h
  const int ps = sym_side_k<0, true>(a, sym_bearing_bits(dx, dy, kBearingA6), __float_as_uint(f.w), bh, slow);
  if (slow) {
    // Rare (1-2 per thousand), expensive (fp64 with the reference's operation sequence) and divergent: not here, where 31
    // lanes would wait for one -- the pair goes to the CTA's queue and the slow pass after the loop takes the queued pairs
    // one per THREAD.  (Evaluated on the spot, the slow pairs cost as much as all the fast ones together.)
    const uint32_t slot = atom_add_shared(qcount_s, 1u);
    if (slot < (uint32_t)kWarpQueue) sts_u32(queue_s + 4u * slot, qentry);
    else warp_pair_slow<TORUS>(a, f, sh.theta[qentry >> 24], row_pad_s + 4u, o, n_fp64, n_differ);
    return;
  }
  if (!FULL_FOV) {                                       // vf_supcalc.py:119 on padded positions
    const int pe = ps + 2 * ((int)hraw - kMagicBits);
    if (!(((unsigned)(ps - a.fov0p) < a.span) | ((unsigned)(pe - a.fov0p) < a.span))) return;
  }
  const uint32_t wa = row_pad_s + 4u * (uint32_t)(ps >> 5);
  asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(wa), "r"(__funnelshift_l(0u, mask, ps)));
  asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(wa + 4u), "r"(__funnelshift_l(mask, 0u, ps)));
}

// MULTI: n_steps time steps in ONE launch, a barrier over all CTAs between steps -- for runs so small that a step is
// shorter than a kernel launch (one run of 100 agents: BASELINE configs[1]); the record tables swap roles every step.
// Without peers, culling lists or re-sorting (abm_api.cu checks).
//   1: cooperative launch, the barrier is a ticket counter in global memory (an atomic + an acquire poll per step);
//   2: the whole grid is ONE thread-block cluster (one replicate of at most 16 * 4 agents): the barrier is the hardware's
//      barrier.cluster (release / acquire at cluster scope); CTAs beyond the last focal agent only take part in the
//      barrier.  Measured (scratch/c2_cluster_probe.py, us per step, cluster / grid barrier): N = 16 5.9 / 6.5, N = 64
//      8.3 / 8.7; with 8 focal agents per CTA (N = 100 in 13 CTAs) 10.1 / 9.2 -- the longer pair loop per CTA costs more
//      than the barrier saves, so larger replicates keep mode 1.  Either way a step of a small run is a chain of
//      dependent latencies (state loads from L2, the few fp64 pairs, the fp64 epilogue), not the barrier.
template <bool TORUS, bool CULL, bool FULL_FOV, bool UNIFORM_R, int MULTI>
__global__ void __launch_bounds__(kWarpMaxThreads, 2) vf_step_warp_kernel(const __grid_constant__ VFKernelArgs a, const int F,
                                                                       const int n_steps_arg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpShared sh;
  sh.focal = reinterpret_cast<float4*>(smem_raw);
  sh.theta = reinterpret_cast<float*>(sh.focal + F);
  sh.rows = reinterpret_cast<uint32_t*>(sh.theta + kWarpMaxFocal);
  sh.tile_list = reinterpret_cast<int*>(sh.rows + (size_t)(a.W + 3) * F);
  sh.fbox = sh.tile_list + kMaxTileList + 4;
  sh.queue = reinterpret_cast<uint32_t*>(sh.fbox + 16);
  __shared__ int s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = blockDim.x;                                // 256 or 512 threads
  const int per_rep = (a.tile_count + F - 1) / F;
  const bool idle = MULTI == 2 && (int)blockIdx.x >= per_rep * a.B;   // cluster padding: barrier only
  const int b = idle ? 0 : blockIdx.x / per_rep;
  const int li0 = idle ? 0 : (blockIdx.x - b * per_rep) * F;   // first focal agent of the CTA inside this engine's tile
  const int nf = idle ? 0 : min(F, a.tile_count - li0);    // focal agents of this CTA
  const int R = a.R, W = a.W;
  const int row_words = W + 3;
  int* tile_list = sh.tile_list;
  int* fbox = sh.fbox;
  const int n_steps = MULTI ? n_steps_arg : 1;
#pragma unroll 1
  for (int step_i = 0; step_i < n_steps; ++step_i) {
  const float4* rec_in_t = (MULTI && (step_i & 1)) ? a.rec_out : a.rec_in;
  float4* rec_out_t = (MULTI && (step_i & 1)) ? const_cast<float4*>(a.rec_in) : a.rec_out;
  const float4* rep_in = rec_in_t + (size_t)b * a.N;

  for (int w = tid; w < row_words * F; w += T) sh.rows[w] = 0u;
  if (a.n_peers > 0 && tid <= a.n_peers) {   // fused tile exchange: wait for every rank's previous step (thread r: rank r)
    const uint32_t* fl = a.xflags + tid;
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(fl) : "memory");
    } while ((int)(v - a.step_no) < 0);
  }
  if (tid < 4) fbox[tid] = (tid < 2) ? 0x7fffffff : (int)0x80000000;
  if (tid == 0) { tile_list[kMaxTileList] = 0; tile_list[kMaxTileList + 1] = 0; }
  __syncthreads();                                          // (also: nobody reads the record table before the hand-shake)
  if (tid < nf) {                                           // focal constants, one thread per focal agent
    const int i = vf_tile_slot(a, li0 + tid);
    const float4 me = __ldcg(rep_in + i);
    const float th = a.theta[(size_t)b * a.N + i];
    sh.focal[tid] = make_float4(me.x, me.y, me.z, __uint_as_float(sym_heading_const(th)));
    sh.theta[tid] = th;
    atomicMin(&fbox[0], float_ordered(me.x)); atomicMin(&fbox[1], float_ordered(me.y));
    atomicMax(&fbox[2], float_ordered(me.x)); atomicMax(&fbox[3], float_ordered(me.y));
  }
  __syncthreads();
  const float fx0 = ordered_float(fbox[0]), fy0 = ordered_float(fbox[1]);
  const float fx1 = ordered_float(fbox[2]), fy1 = ordered_float(fbox[3]);

  // ---- record tiles to visit: bounding box of the CTA's focal agents against the tiles' boxes ----
  const int tile_sz = a.tile_bbox != nullptr ? a.cull_tile : kRecTile;   // a power of two >= kChunk
  const int tile_sh = 31 - __clz(tile_sz);
  const int n_tiles = (a.N + tile_sz - 1) / tile_sz;
  const bool use_list = CULL && a.tile_bbox != nullptr && n_tiles <= kMaxTileList;
  int n_stage = idle ? 0 : n_tiles;
  if (use_list) {
    const float4* bb = a.tile_bbox + (size_t)b * n_tiles;
    const float* c2 = a.tile_cull2 + (size_t)b * n_tiles;
    for (int t = tid; t < n_tiles; t += T) {
      const float4 q = __ldcg(bb + t);
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
  }

  // ---- pair loop: the two halves of the CTA take alternate chunks of 128 records of the visited tiles, a thread per
  //      record; every record meets all focal agents of the CTA ----
  unsigned n_fp64 = 0, n_differ = 0;
  {
    const uint32_t rows_s = smem_u32(sh.rows), focal_s = smem_u32(sh.focal), queue_s = smem_u32(sh.queue);
    const uint32_t qcount_s = smem_u32(tile_list + kMaxTileList + 1);
    const float S = a.y_scale;
    const float c3s = -1.0f / (3.0f * S * S), c5s = 1.0f / (5.0f * S * S * S * S), c7s = -1.0f / (7.0f * S * S * S * S * S * S);
    const int cpt_sh = tile_sh - 7;                          // chunks per tile = 2^cpt_sh
    const int n_chunks = n_stage << cpt_sh;
    const int off = tid & (kChunk - 1);
    for (int c = tid >> 7; c < n_chunks; c += T / kChunk) {
      const int st = c >> cpt_sh;
      const int j = ((((use_list ? tile_list[st] : st) << cpt_sh) + (c & ((1 << cpt_sh) - 1))) << 7) + off;
      if (j >= a.N) continue;
      const float4 o = __ldcg(rep_in + j);
      if (CULL && !TORUS && UNIFORM_R) {
        // beyond reach of the whole focal box: beyond reach of every focal agent in it (same expression, monotonic)
        const float gx = fmaxf(0.0f, fmaxf(o.x - fx1, fx0 - o.x)), gy = fmaxf(0.0f, fmaxf(o.y - fy1, fy0 - o.y));
        if (fmaf(gx, gx, gy * gy) > o.w) continue;
      }
      // four focal agents per pass: four independent evaluations in flight per thread (the loop is latency-bound
      // otherwise -- measured), without the code size of a fully unrolled F = 8
#pragma unroll 1
      for (int f0 = 0; f0 < nf; f0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int f = f0 + u;
          if (f < nf)
            warp_pair<TORUS, CULL, FULL_FOV, UNIFORM_R>(a, lds_f4(focal_s + 16u * (uint32_t)f),
                                                         rows_s + 4u * (uint32_t)(row_words * f), o, S, c3s, c5s, c7s,
                                                         ((uint32_t)f << 24) | (uint32_t)j, queue_s, qcount_s, sh, n_fp64,
                                                         n_differ);
        }
      }
    }
    // ---- slow pass: the queued pairs, one per thread ----
    __syncthreads();
    const int nq = min(tile_list[kMaxTileList + 1], kWarpQueue);
    for (int q = tid; q < nq; q += T) {
      const uint32_t ent = sh.queue[q];
      const int f = (int)(ent >> 24), j = (int)(ent & 0xffffffu);
      warp_pair_slow<TORUS>(a, sh.focal[f], sh.theta[f], rows_s + 4u * (uint32_t)(row_words * f) + 4u, __ldcg(rep_in + j),
                            n_fp64, n_differ);
    }
  }
  {
    const unsigned nfp = __reduce_add_sync(0xffffffffu, n_fp64), nd = __reduce_add_sync(0xffffffffu, n_differ);
    if (lane == 0 && nfp) atomicAdd(&a.counters[0], (unsigned long long)nfp);
    if (lane == 0 && nd) atomicAdd(&a.counters[2], (unsigned long long)nd);
  }
  __syncthreads();

  // ---- epilogue: warp f takes focal agent f; lanes take the words of the row, edge sums reduced with shuffles ----
  if (warp < nf) {
    const int li = li0 + warp;
    const int i = vf_tile_slot(a, li);
    const size_t gi = (size_t)b * a.N + i;
    uint32_t* padrow = sh.rows + (size_t)row_words * warp;
    if (lane == 0) vf_fold_padding(padrow, 1, R, W);      // padding words back onto the ring (vf_supcalc.py:122-127)
    __syncwarp();
    const uint32_t* row = padrow + 1;                      // real word 0
    const float4 me = sh.focal[warp];
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
      const float th = sh.theta[warp];
      FlockTerms ft;
      if (a.phi_ok) {
        EdgeSums z;
        z.zsr = zsr; z.zsi = zsi; z.zdr = zdr; z.zdi = zdi; z.v_first = v_first; z.v_last = v_last;
        ft = vf_terms_from_edges(z, a, vel0, prm, A0, B0, V0);
      } else {   // len(PHI) != len(soc_v_field): the reference skips the calculation (vf_agent.py:282-284)
        ft.dvel = ft.dpsi = ft.a_blob = ft.a_edge = ft.b_blob = ft.b_edge = 0.0;
      }
      double dpsi = ft.dpsi, dvel = ft.dvel;
      if (a.line_map)                                               // lines to follow replace the heading change (:273-276)
        dpsi = vf_follow_lines(a, (double)me.x, (double)me.y, (double)me.z, (double)th, vel0);
      if (a.limit_movement) dpsi = limit_abs(dpsi, a.max_th);       // vf_agent.py:293-294
      double nth = wrap_heading_once((double)th + dpsi);            // :295-296
      double nv = vel0 + dvel;                                      // :298
      if (a.limit_movement) nv = limit_abs(nv, a.max_vel);          // :299-300
      double sn, cn;
      sincos(nth, &sn, &cn);
      double nx = (double)me.x + nv * cn;                           // :303-306
      double ny = (double)me.y - nv * sn;
      if (!TORUS) reflect_from_walls(nx, ny, nth, (double)me.z, a.width_d, a.height_d, a.pad_d,
                                     a.line_map ? ABM_PI_D : ABM_PI_D / 2.0);     // vf_agent.py:80-129
      else teleport_torus(nx, ny, (double)me.z, a.width_d, a.height_d, a.pad_d);
      const float4 rec_new = make_float4((float)nx, (float)ny, me.z, rep_in[i].w);
      rec_out_t[gi] = rec_new;
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

  if (MULTI == 2) {   // one cluster: every CTA's records of this step are in place before anybody reads them
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
    continue;
  }
  if (MULTI == 1) {   // grid-wide barrier
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(a.step_ticket, 1u);
      const uint32_t target = (uint32_t)(step_i + 1) * gridDim.x;
      uint32_t v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.step_ticket) : "memory");
      } while (v < target);
    }
    __syncthreads();
    continue;
  }
  // ---- fused tile exchange: the LAST CTA of the launch closes the step for this rank ----
  if (a.step_ticket == nullptr) return;
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();                                   // this CTA's records (local and peer stores) before its ticket
    s_last = (atomicAdd(a.step_ticket, 1u) == gridDim.x - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();                                            // every CTA's records are visible to this one
  if (a.bbox_out != nullptr) {
    // bounding boxes of this rank's record tiles in the table just written (B == 1), a warp per tile
    // (own tiles: the contiguous range, or every tile_cycle-th tile from tile_phase on)
    const bool cyc = a.tile_cycle > 1;
    const int t0 = cyc ? a.tile_phase : a.tile_begin >> tile_sh;
    const int t1 = cyc ? n_tiles : (a.tile_begin + a.tile_count + tile_sz - 1) >> tile_sh;
    const int dt = cyc ? a.tile_cycle : 1;
    for (int t = t0 + warp * dt; t < t1; t += (T / 32) * dt) {
      float x0 = 3.0e38f, y0 = 3.0e38f, x1 = -3.0e38f, y1 = -3.0e38f;
      for (int j = (t << tile_sh) + lane; j < min(a.N, (t + 1) << tile_sh); j += 32) {
        const float4 v = __ldcg(a.rec_out + j);
        x0 = fminf(x0, v.x); y0 = fminf(y0, v.y); x1 = fmaxf(x1, v.x); y1 = fmaxf(y1, v.y);
      }
      for (int off = 16; off > 0; off >>= 1) {
        x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, off)); y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, off));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, off)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, off));
      }
      if (lane == 0) {
        const float4 box = make_float4(x0, y0, x1, y1);
        a.bbox_out[t] = box;
        for (int p = 0; p < a.n_peers; ++p) a.peer_bbox_out[p][t] = box;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    *a.step_ticket = 0u;                                      // for the next launch (stream order)
    __threadfence_system();
    const uint32_t done = a.step_no + 1u;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.xflags + a.my_rank), "r"(done) : "memory");
    for (int p = 0; p < a.n_peers; ++p)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_flags[p] + a.my_rank), "r"(done) : "memory");
  }
  }   // step loop (one pass unless MULTI)
}

// focal agents per CTA: as many as keep >= 8 CTAs per SM (two waves of the 4 resident ones) in the grid -- measured on
// an eighth of the 65 536-agent swarm (scratch/c5_tile_probe.py): F = 4 1.64, F = 2 1.87, F = 8 1.84 ms summed over the
// eight tiles; the whole swarm: F = 8 1.15, F = 4 1.28 ms
int vf_warp_focal_per_cta(long long focal_total, int n_sms) {
  int F = kWarpMaxFocal;
  while (F > 1 && focal_total / F < 8LL * n_sms) F >>= 1;
  return F;
}

static int warp_threads_choice(const VFKernelArgs& a, int F) {
  const char* env = getenv("ABM_VF_WARP_THREADS");            // measurement probes only
  if (env) return atoi(env) == 512 ? 512 : 256;
  (void)a; (void)F;
  return kWarpThreads;
}

template <bool TORUS, bool CULL, bool FULL_FOV, bool UNIFORM_R>
static void launch_warp_variant(const VFKernelArgs& a, int F, cudaStream_t stream) {
  const int per_rep = (a.tile_count + F - 1) / F;
  const size_t smem = vf_warp_smem_bytes(a.W, F);
  static SmemOptIn optin;
  if (smem > 48 * 1024) optin.ensure(vf_step_warp_kernel<TORUS, CULL, FULL_FOV, UNIFORM_R, 0>, smem);
  vf_step_warp_kernel<TORUS, CULL, FULL_FOV, UNIFORM_R, 0>
      <<<(unsigned)((size_t)a.B * per_rep), warp_threads_choice(a, F), smem, stream>>>(a, F, 1);
}
template <bool TORUS, bool CULL>
static void launch_warp_fr(const VFKernelArgs& a, int F, bool uniform_r, cudaStream_t stream) {
  if (a.full_fov) { if (uniform_r) launch_warp_variant<TORUS, CULL, true, true>(a, F, stream); else launch_warp_variant<TORUS, CULL, true, false>(a, F, stream); }
  else { if (uniform_r) launch_warp_variant<TORUS, CULL, false, true>(a, F, stream); else launch_warp_variant<TORUS, CULL, false, false>(a, F, stream); }
}

static int warp_focal_choice(const VFKernelArgs& a) {
  const char* env = getenv("ABM_VF_WARP_FOCAL");              // measurement probes only
  int F = env ? atoi(env) : vf_warp_focal_per_cta((long long)a.B * a.tile_count, a.n_sms);
  if (F != 1 && F != 2 && F != 4 && F != 8) F = 1;
  return F;
}

void launch_vf_step_warp(const VFKernelArgs& a, bool cull, bool uniform_r, cudaStream_t stream) {
  const int F = warp_focal_choice(a);
  if (a.boundary == 1) { if (cull) launch_warp_fr<true, true>(a, F, uniform_r, stream); else launch_warp_fr<true, false>(a, F, uniform_r, stream); }
  else { if (cull) launch_warp_fr<false, true>(a, F, uniform_r, stream); else launch_warp_fr<false, false>(a, F, uniform_r, stream); }
}

// n_steps steps in one cooperative launch (no culling, no peers); false: the grid cannot be co-resident -- nothing launched
template <bool TORUS, bool FULL_FOV, bool UNIFORM_R>
static bool launch_warp_multi_variant(const VFKernelArgs& a, int F, int n_steps, cudaStream_t stream) {
  auto kernel = vf_step_warp_kernel<TORUS, false, FULL_FOV, UNIFORM_R, 1>;
  const int per_rep = (a.tile_count + F - 1) / F;
  const unsigned grid = (unsigned)((size_t)a.B * per_rep);
  const size_t smem = vf_warp_smem_bytes(a.W, F);
  static SmemOptIn optin;
  if (smem > 48 * 1024) optin.ensure(kernel, smem);
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarpThreads, smem) != cudaSuccess) return false;
  if ((long long)per_sm * a.n_sms < (long long)grid) return false;
  VFKernelArgs args = a;
  int f = F, n = n_steps;
  void* params[3] = {&args, &f, &n};
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(grid), dim3(kWarpThreads), params, smem, stream) ==
         cudaSuccess;
}
// n_steps steps of ONE small replicate in one launch whose grid is a single thread-block cluster (at most 16 CTAs of 4
// focal agents); false: not applicable / the cluster cannot be scheduled -- nothing launched
template <bool TORUS, bool FULL_FOV, bool UNIFORM_R>
static bool launch_warp_cluster_variant(const VFKernelArgs& a, int F, int cluster, int n_steps, cudaStream_t stream) {
  auto kernel = vf_step_warp_kernel<TORUS, false, FULL_FOV, UNIFORM_R, 2>;
  const size_t smem = vf_warp_smem_bytes(a.W, F);
  static SmemOptIn optin;
  if (smem > 48 * 1024) optin.ensure(kernel, smem);
  if (cluster > 8 &&
      cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
    return false;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cluster); cfg.blockDim = dim3(kWarpThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, a, F, n_steps) == cudaSuccess;
}
bool launch_vf_step_warp_cluster(const VFKernelArgs& a, bool uniform_r, int n_steps, cudaStream_t stream) {
  constexpr int kClusterMaxFocal = 4;
  if (a.B != 1 || a.tile_count != a.N || a.N > 16 * kClusterMaxFocal || getenv("ABM_VF_NO_CLUSTER")) return false;
  int F = 1;
  while (F < kClusterMaxFocal && (a.N + F - 1) / F > 16) F <<= 1;   // as many CTAs as a cluster can hold
  const int ctas = (a.N + F - 1) / F;
  const int cluster = ctas <= 1 ? 1 : (ctas <= 2 ? 2 : (ctas <= 4 ? 4 : (ctas <= 8 ? 8 : 16)));
  if (a.boundary == 1) {
    if (a.full_fov) return uniform_r ? launch_warp_cluster_variant<true, true, true>(a, F, cluster, n_steps, stream) : launch_warp_cluster_variant<true, true, false>(a, F, cluster, n_steps, stream);
    return uniform_r ? launch_warp_cluster_variant<true, false, true>(a, F, cluster, n_steps, stream) : launch_warp_cluster_variant<true, false, false>(a, F, cluster, n_steps, stream);
  }
  if (a.full_fov) return uniform_r ? launch_warp_cluster_variant<false, true, true>(a, F, cluster, n_steps, stream) : launch_warp_cluster_variant<false, true, false>(a, F, cluster, n_steps, stream);
  return uniform_r ? launch_warp_cluster_variant<false, false, true>(a, F, cluster, n_steps, stream) : launch_warp_cluster_variant<false, false, false>(a, F, cluster, n_steps, stream);
}

bool launch_vf_step_warp_multi(const VFKernelArgs& a, bool uniform_r, int n_steps, cudaStream_t stream) {
  const int F = warp_focal_choice(a);
  if (a.boundary == 1) {
    if (a.full_fov) return uniform_r ? launch_warp_multi_variant<true, true, true>(a, F, n_steps, stream) : launch_warp_multi_variant<true, true, false>(a, F, n_steps, stream);
    return uniform_r ? launch_warp_multi_variant<true, false, true>(a, F, n_steps, stream) : launch_warp_multi_variant<true, false, false>(a, F, n_steps, stream);
  }
  if (a.full_fov) return uniform_r ? launch_warp_multi_variant<false, true, true>(a, F, n_steps, stream) : launch_warp_multi_variant<false, true, false>(a, F, n_steps, stream);
  return uniform_r ? launch_warp_multi_variant<false, false, true>(a, F, n_steps, stream) : launch_warp_multi_variant<false, false, false>(a, F, n_steps, stream);
}

}  // namespace abm
