// BASE (collective foraging) path for sm_100a: environment phase (agent-patch interaction)
// and agent phase (social visual field with distance-ordered occlusion, decision process,
// mode machine, kinematics, walls).
//
// Reference behaviour restated here (paths relative to the reference root):
//   Agent.calc_social_V_proj            agent.py:396-419
//   Agent.projection_field              agent.py:457-597
//   Agent.exlude_V_source_data          agent.py:421-445
//   Agent.calc_I_priv / update_decision_processes / update     agent.py:168-283
//   supcalc.F_reloc_LR / random_walk    supcalc.py:38-47, 81-92
//   sims.notify_agent / refine_ar_overlap_group / patch loop   sims.py:29-56, 790-858
//   Simulation.bias_agent_towards_res_center / add_new_resource_patch   sims.py:544-552, 332-374
//   Rescource.deplete                   rescource.py:118-133
//
// The pair arithmetic of this variant runs in fp64 and follows the reference's own operation
// sequence (unit vectors -> arccos -> first-arg-min bin -> int() truncation), so the integer
// interval ends agree with the float64 reference without guard bands; with N <= a few hundred
// agents per replicate the all-pairs work is small next to the occlusion pass.
#include "abm_base.cuh"
#include "abm_vf_device.cuh"

namespace abm {

// ---------------------------------------------------------------------------------------
// counter-based RNG (Philox-4x32-10): draws are a pure function of (seed, replicate, agent /
// patch, step, purpose), so replicate batches are reproducible regardless of scheduling.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {   // 53-bit uniform in [0, 1)
  return (double)((((unsigned long long)a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------------------------------
// environment phase: one thread per replicate, sequential in patch / agent order
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void notify(const BaseAgentPtrs& ag, size_t g, int status, int res_id, uint32_t tau_mask) {
  const int before = ag.env_status[g];                       // sims.py:31-32
  ag.env_status[g] = status;
  const uint32_t bit = (status - before > 0) ? 1u : 0u;      // :34-37
  ag.novelty[g] = ((ag.novelty[g] << 1) | bit) & tau_mask;   // np.roll(novelty, 1); novelty[0] = bit
  ag.patch_id[g] = res_id;                                   // :39-42 (None -> -1)
}

__global__ void base_env_kernel(const BaseKernelArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const BaseParams prm = *reinterpret_cast<const BaseParams*>(a.params + (size_t)b * a.param_stride);
  const uint32_t tau_mask = (a.Tau >= 32) ? 0xffffffffu : ((1u << a.Tau) - 1u);
  const size_t a0 = (size_t)b * a.N, p0 = (size_t)b * a.P;
  const double r = a.radius;
  // "agent is on some patch" marks live in bit 31 of a scratch copy of nothing: recomputed below
  for (int p = 0; p < a.P; ++p) {
    const double prad = a.pa.radius[p0 + p];
    const double pcx = (double)a.pa.x[p0 + p] + prad, pcy = (double)a.pa.y[p0 + p] + prad;
    bool destroy = false;
    for (int i = 0; i < a.N; ++i) {
      const size_t g = a0 + i;
      const double ddx = ((double)a.ag.x[g] + r) - pcx, ddy = ((double)a.ag.y[g] + r) - pcy;
      if (!(sqrt(ddx * ddx + ddy * ddy) < prad)) continue;                         // sims.py:45-56
      // bias_agent_towards_res_center (sims.py:544-552): no wrap of the heading
      {
        const double dx = pcx - ((double)a.ag.x[g] + r), dy = pcy - ((double)a.ag.y[g] + r);
        const double th = a.ag.theta[g];
        double cl = fmod(atan2(dy, dx) + th, ABM_TWO_PI_D);
        if (cl < 0.0) cl += ABM_TWO_PI_D;
        a.ag.theta[g] = (float)(th + (cl - ABM_PI_D) * 0.02);
      }
      if (destroy) {
        notify(a.ag, g, -1, -1, tau_mask);                                          // :812-813
      } else {
        notify(a.ag, g, 1, a.pa.id[p0 + p], tau_mask);                              // :816-818 (pooling_time == 0)
        if (a.teleport_exploit) {                                                   // :820-821
          a.ag.x[g] = (float)((double)a.pa.x[p0 + p] + prad - r);
          a.ag.y[g] = (float)((double)a.pa.y[p0 + p] + prad - r);
        }
        if (a.ag.override_mode[g] == OV_EXPLOIT) {                                  // :824-828
          double take = fmin(prm.consumption, (double)a.pa.quality[p0 + p]);        // rescource.py:121-122
          double left = a.pa.left[p0 + p];
          if (left >= take) left -= take; else { take = left; left = 0.0; }
          a.pa.left[p0 + p] = (float)left;
          destroy = !(left > 0.0);
          const float c = a.ag.collected[g];
          a.ag.collected_before[g] = c;
          a.ag.collected[g] = (float)((double)c + take);
          if (destroy) {                                                            // :829-836
            for (int i2 = 0; i2 < a.N; ++i2) {
              const size_t g2 = a0 + i2;
              const double ex = ((double)a.ag.x[g2] + r) - pcx, ey = ((double)a.ag.y[g2] + r) - pcy;
              if (sqrt(ex * ex + ey * ey) < prad) notify(a.ag, g2, -1, -1, tau_mask);
            }
          }
        }
      }
      a.ag.mode[g] |= 0x100;   // scratch mark: on a patch this step
    }
    if (destroy && a.regenerate) {   // kill_resource + add_new_resource_patch(force_id) (sims.py:321-374)
      bool placed = false;
      for (unsigned t = 0; t < 10000u && !placed; ++t) {
        const uint4 rn = philox4x32(make_uint4((uint32_t)b, (uint32_t)p, a.step, t),
                                    make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ 0x50415443u));
        const double R_ = a.patch_radius;
        double lox, hix, loy, hiy;
        if (a.border_overlap) { lox = a.pad - R_; hix = a.width + a.pad - R_; loy = a.pad - R_; hiy = a.height + a.pad - R_; }
        else { lox = a.pad; hix = a.width + a.pad - 2 * R_; loy = a.pad; hiy = a.height + a.pad - 2 * R_; }
        const uint4 rn2 = philox4x32(make_uint4((uint32_t)b, (uint32_t)p, a.step, t), make_uint2((uint32_t)a.seed, 0x51554c54u));
        const double nx = floor(lox + floor(hix - lox) * u01(rn.x, rn.y));          // np.random.randint(lo, hi)
        const double ny = floor(loy + floor(hiy - loy) * u01(rn.z, rn.w));
        bool ok = true;
        for (int p2 = 0; p2 < a.P; ++p2) {                                          // proove_sprite: no patch-patch overlap
          if (p2 == p) continue;
          const double r2 = a.pa.radius[p0 + p2];
          const double ex = (nx + R_) - ((double)a.pa.x[p0 + p2] + r2), ey = (ny + R_) - ((double)a.pa.y[p0 + p2] + r2);
          if (ex * ex + ey * ey <= (R_ + r2) * (R_ + r2)) { ok = false; break; }
        }
        if (!ok) continue;
        const int units = a.min_units + (int)floor((double)(a.max_units - a.min_units) * u01(rn2.x, rn2.y));
        const double q = a.min_quality + (a.max_quality - a.min_quality) * u01(rn2.z, rn2.w);
        a.pa.x[p0 + p] = (float)nx; a.pa.y[p0 + p] = (float)ny; a.pa.radius[p0 + p] = (float)R_;
        a.pa.left[p0 + p] = (float)units; a.pa.quality[p0 + p] = (float)q;
        placed = true;
        atomicAdd(&a.counters[0], 1ull);
      }
      if (!placed) atomicAdd(&a.counters[1], 1ull);
    } else if (destroy) {
      a.pa.radius[p0 + p] = 0.0f;   // resource.kill() without regeneration: the patch no longer exists
    }
  }
  // agents on no patch (and not colliding) are told so (sims.py:847-855, pooling_time == 0)
  for (int i = 0; i < a.N; ++i) {
    const size_t g = a0 + i;
    const int m = a.ag.mode[g];
    if (m & 0x100) a.ag.mode[g] = m & 0xff;
    else if (a.ag.override_mode[g] != OV_COLLIDE) notify(a.ag, g, -1, -1, tau_mask);
    // frozen snapshot for the agent phase: every agent sees the same positions / modes
    a.ag.snap_x[g] = a.ag.x[g];
    a.ag.snap_y[g] = a.ag.y[g];
    a.ag.snap_override[g] = a.ag.override_mode[g];
  }
}

void launch_base_env(const BaseKernelArgs& a, cudaStream_t stream) {
  const int threads = 64;
  base_env_kernel<<<(a.B + threads - 1) / threads, threads, 0, stream>>>(a);
}

// ---------------------------------------------------------------------------------------
// agent phase: one warp per focal agent
// ---------------------------------------------------------------------------------------
struct __align__(16) ObjRec {
  int s, e;      // raw interval ends, int() truncated (agent.py:545-546)
  double d;      // centre distance (agent.py:526-528)
};

size_t base_agents_smem_bytes(int N, int W, int warps) {
  const size_t per_warp = 2 * sizeof(ObjRec) * (size_t)N + 2 * sizeof(int) * (size_t)N + sizeof(uint32_t) * (size_t)(W + 1);
  return ((per_warp + 15) / 16 * 16) * warps;
}
int base_agents_warps(int N, int W, size_t smem_limit) {
  int w = 8;
  while (w > 1 && base_agents_smem_bytes(N, W, w) > smem_limit) w >>= 1;
  return w;
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

__global__ void __launch_bounds__(256) base_agent_kernel(const BaseKernelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const size_t per_warp = (2 * sizeof(ObjRec) * (size_t)a.N + 2 * sizeof(int) * (size_t)a.N +
                           sizeof(uint32_t) * (size_t)(a.W + 1) + 15) / 16 * 16;
  unsigned char* base = smem_raw + per_warp * wib;
  ObjRec* raw = reinterpret_cast<ObjRec*>(base);
  ObjRec* sorted = raw + a.N;
  int* key = reinterpret_cast<int*>(sorted + a.N);      // tie-break key of raw[m]; bit 30 = social cue
  int* skey = key + a.N;                                // same for sorted[rank]
  uint32_t* row = reinterpret_cast<uint32_t*>(skey + a.N);   // un-flipped field v (agent.py:480)

  const long long gw = (long long)blockIdx.x * wpb + wib;   // global warp = (replicate, focal agent)
  if (gw >= (long long)a.B * a.N) return;
  const int b = (int)(gw / a.N), i = (int)(gw - (long long)b * a.N);
  const size_t a0 = (size_t)b * a.N, gi = a0 + i;
  const int R = a.R, W = a.W, N = a.N;
  for (int w = lane; w < W + 1; w += 32) row[w] = 0u;

  const float xi_f = a.ag.snap_x[gi], yi_f = a.ag.snap_y[gi];
  const double xi = xi_f, yi = yi_f, r = a.radius;
  const FocalExact fe = vf_focal_exact(xi_f, yi_f, (float)a.radius, a.ag.theta[gi]);   // same v1 construction (agent.py:484-495)
  const int my_patch = a.ag.patch_id[gi];

  // ---- candidates, classes, raw intervals (agent.py:396-419, 497-556) ----
  int M = 0;
  for (int j0 = 0; j0 < N; j0 += 32) {
    const int j = j0 + lane;
    bool rec = false, social = false;
    ObjRec o; o.s = 0; o.e = 0; o.d = 0.0;
    int k2 = 0;
    if (j < N) {
      const size_t gj = a0 + j;
      const float xj_f = a.ag.snap_x[gj], yj_f = a.ag.snap_y[gj];
      const double cjx = __dadd_rn((double)xj_f, r), cjy = __dadd_rn((double)yj_f, r);
      const double v2x = __dadd_rn(cjx, -fe.cix), v2y = __dadd_rn(cjy, -fe.ciy);
      const double n2 = __dsqrt_rn(__dadd_rn(__dmul_rn(v2x, v2x), __dmul_rn(v2y, v2y)));
      const bool in_range = n2 <= a.vision_range;                                   // agent.py:400
      const bool is_expl = (j != i) && (a.ag.snap_override[gj] == OV_EXPLOIT);      // :402-403
      int cls = 0;   // 0 none, 1 social, 2 occluder (other), 3 occluder (same-patch exploiter)
      if (in_range) {
        if (is_expl) {
          const int pj = a.ag.patch_id[gj];
          if (a.patchwise_exclusion && pj == my_patch) cls = 3;                     // :406-409
          else if (pj != -1) cls = 1;                                               // :410, :413
        } else {
          cls = 2;                                                                  // :405 (self included, skipped below)
        }
      }
      if (!a.visual_exclusion && cls != 1) cls = 0;                                 // :415-419
      const bool same = (xj_f == xi_f) && (yj_f == yi_f);                           // :502
      if (cls != 0 && !same && n2 > 0.0) {
        const double u2x = __ddiv_rn(v2x, n2), u2y = __ddiv_rn(v2y, n2);
        double dot = __dadd_rn(__dmul_rn(fe.u1x, u2x), __dmul_rn(fe.u1y, u2y));
        dot = fmin(1.0, fmax(-1.0, dot));
        double ang = acos(dot);                                                     // supcalc.py:31
        if (__dadd_rn(__dmul_rn(fe.u1x, u2y), -__dmul_rn(fe.u1y, u2x)) < 0.0) ang = -ang;
        if (ang < 0.0) ang = __dadd_rn(ang, ABM_TWO_PI_D);                          // % 2pi (agent.py:516)
        const double ca = (ang > 0.0 && ang < ABM_PI_D) ? -ang : __dadd_rn(ABM_TWO_PI_D, -ang);   // :520-523
        if (a.fov0 < ca && ca < a.fov1) {                                           // :535
          const int k = nearest_bin_exact(ca, R, a.lin_step);                       // :532
          const double vis = __dmul_rn(2.0, atan(__ddiv_rn(r, n2)));                // :529
          const double size = __dmul_rn(__ddiv_rn(vis, ABM_TWO_PI_D), (double)R);   // :543
          const double half = __ddiv_rn(size, 2.0);
          o.s = (int)__dadd_rn((double)k, -half);                                   // :545-546 int(): toward zero
          o.e = (int)__dadd_rn((double)k, half);
          o.d = n2;
          social = (cls == 1);
          // list order of the reference: social cues, then other occluders, then same-patch
          // exploiters, each in agent order (agent.py:402-410, 472-477)
          k2 = ((cls == 1) ? 0 : (cls == 2 ? 1 : 2)) * N + j;
          rec = true;
        }
      }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, rec);
    if (rec) {
      const int idx = M + __popc(mask & ((1u << lane) - 1u));
      raw[idx] = o;
      key[idx] = k2 | (social ? (1 << 30) : 0);
    }
    M += __popc(mask);
  }
  __syncwarp();

  // ---- occlusion (agent.py:421-445) and fill (agent.py:569-590) ----
  if (a.visual_exclusion) {
    for (int m = lane; m < M; m += 32) {   // rank by (distance, list order): stable sort of :424
      const ObjRec f = raw[m];
      const int kf = key[m] & 0x3fffffff;
      int rank = 0;
      for (int q = 0; q < M; ++q) {
        const double dq = raw[q].d;
        const int kq = key[q] & 0x3fffffff;
        rank += (dq < f.d) || (dq == f.d && kq < kf);
      }
      sorted[rank] = f;
      skey[rank] = key[m];
    }
    __syncwarp();
    for (int p = lane; p < M; p += 32) {
      if (!(skey[p] & (1 << 30))) continue;                                         // :447-455 only social cues are drawn
      const ObjRec f = sorted[p];
      int sx = f.s, ex = f.e;
      for (int q = 0; q < p; ++q) {
        const ObjRec o = sorted[q];
        if (o.d < f.d) {                                                            // :430 strict
          if (sx <= o.s && o.s <= ex) ex = o.s;                                     // :432-433
          if (sx <= o.e && o.e <= ex) sx = o.e;                                     // :435-436
          if (o.s <= sx && o.e >= ex) { sx = 0; ex = 0; }                           // :438-440
        }
      }
      base_draw(row, R, sx, ex);
    }
  } else {
    for (int m = lane; m < M; m += 32) base_draw(row, R, raw[m].s, raw[m].e);
  }
  __syncwarp();

  // ---- flip + FOV mask (agent.py:593-595): stored[b] = v[R-1-b], kept for b in [mask_lo, mask_hi] ----
  const int h = R / 2;                                       // int(V_field_len / 2) (supcalc.py:86-88)
  const int va = R - 1 - a.mask_hi, vb = R - a.mask_lo;      // kept bins in v coordinates [va, vb)
  const int n_left = popc_range(row, W, max(va, R - h), min(vb, R), lane);    // stored[0:h]  <-> v[R-h:R]
  const int n_right = popc_range(row, W, max(va, 0), min(vb, R - h), lane);   // stored[h:]   <-> v[0:R-h]
  if (a.fields_out) {
    uint32_t* out = a.fields_out + gi * W;
    for (int ws = lane; ws < W; ws += 32) {
      uint32_t word = flipped_word(row, 1, R, W, ws);
      const int lo = max(a.mask_lo - (ws << 5), 0), hi = min(a.mask_hi + 1 - (ws << 5), 32);
      uint32_t m = 0u;
      if (hi > lo) m = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
      out[ws] = word & m;
    }
  }
  if (lane != 0) return;

  // ---- decision process, mode machine, kinematics (agent.py:168-283), fp64 ----
  const BaseParams prm = *reinterpret_cast<const BaseParams*>(a.params + (size_t)b * a.param_stride);
  const double mean_all = (double)(n_left + n_right) / (double)R;
  const double collected = a.ag.collected[gi];
  const double I_priv = prm.F_N * ((a.ag.novelty[gi] != 0u) ? 1.0 : 0.0) +
                        prm.F_R * (collected - (double)a.ag.collected_before[gi]);   // :168-175
  double w = a.ag.w[gi], u = a.ag.u[gi];
  const double w_p = (w > prm.T_w) ? w : 0.0, u_p = (u > prm.T_u) ? u : 0.0;         // :196-197
  const double dw = prm.Eps_w * mean_all - prm.g_w * (w - prm.B_w) - u_p * prm.S_uw;  // :198-199
  const double du = prm.Eps_u * I_priv - prm.g_u * (u - prm.B_u) - w_p * prm.S_wu;    // :200
  w += dw; u += du;
  if (w > prm.w_max) w = prm.w_max;
  if (w < -prm.w_max) w = -prm.w_max;
  if (u > prm.u_max) u = prm.u_max;
  if (u < -prm.u_max) u = -prm.u_max;
  const bool Wt = w > prm.T_w, Ut = u > prm.T_w;             // tr_u compares with T_w (agent.py:652-657)
  int override = a.ag.override_mode[gi], mode = a.ag.mode[gi] & 0xff;
  const double vel0 = a.ag.vel[gi], th0 = a.ag.theta[gi];
  const bool env1 = a.ag.env_status[gi] == 1;
  double dvel = 0.0, dth = 0.0;
  if (override != OV_COLLIDE) {                              // :233
    if ((!Wt && !Ut) || (Ut && !Wt && !env1)) {              // random_walk (supcalc.py:38-47)
      double rnd;
      if (a.inject_dtheta) rnd = a.inject_dtheta[gi];
      else {
        const uint4 rn = philox4x32(make_uint4((uint32_t)b, (uint32_t)i, a.step, 0u),
                                    make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
        rnd = prm.exp_theta_min + (prm.exp_theta_max - prm.exp_theta_min) * u01(rn.x, rn.y);
      }
      dvel = prm.exp_vel_max; dth = rnd;
      override = OV_NONE; mode = MODE_EXPLORE;
    } else if (Ut && env1) {                                 // :238-240, :254-256
      dvel = -vel0 * prm.exp_stop_ratio; dth = 0.0;
      override = OV_EXPLOIT; mode = MODE_EXPLOIT;
    } else {                                                 // F_reloc_LR (supcalc.py:81-92)
      const double left = (double)n_left / (double)h, right = (double)n_right / (double)(R - h);
      dvel = prm.exp_vel_max - vel0;
      dth = (left - right) * prm.reloc_theta_max;
      override = OV_NONE; mode = MODE_RELOCATE;
    }
  }
  double th = wrap_heading_once(th0 + dth);                  // :269-270
  double vel = vel0 + dvel;                                  // :271
  if (override == OV_NONE && !Wt) {                          // prove_velocity (agent.py:612-620)
    if (fabs(vel) > 1.0) vel = prm.exp_vel_max;
  }
  double sn, cn;
  sincos(th, &sn, &cn);
  double nx = xi + vel * cn, ny = yi - vel * sn;             // :275-276
  reflect_from_walls(nx, ny, th, r, a.width, a.height, a.pad);   // :279
  a.ag.x[gi] = (float)nx; a.ag.y[gi] = (float)ny;
  a.ag.theta[gi] = (float)th; a.ag.vel[gi] = (float)vel;
  a.ag.w[gi] = (float)w; a.ag.u[gi] = (float)u; a.ag.i_priv[gi] = (float)I_priv;
  a.ag.override_mode[gi] = override; a.ag.mode[gi] = mode;
  a.ag.collected_before[gi] = (float)collected;              // :283
}

void launch_base_agents(const BaseKernelArgs& a, cudaStream_t stream) {
  int dev = 0;
  cudaGetDevice(&dev);
  int smem_max = 48 * 1024;
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int warps = base_agents_warps(a.N, a.W, (size_t)smem_max);
  const size_t smem = base_agents_smem_bytes(a.N, a.W, warps);
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(base_agent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  const long long total = (long long)a.B * a.N;
  const unsigned grid = (unsigned)((total + warps - 1) / warps);
  base_agent_kernel<<<grid, warps * 32, smem, stream>>>(a);
}

}  // namespace abm
