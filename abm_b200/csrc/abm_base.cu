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
#include <algorithm>

#include "abm_base.cuh"
#include "abm_base_device.cuh"

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
// environment phase: one warp per replicate, patches and agent chunks in the reference's order
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void notify(const BaseAgentPtrs& ag, size_t g, int status, int res_id, uint32_t tau_mask) {
  const int before = ag.env_status[g];                       // sims.py:31-32
  ag.env_status[g] = status;
  const uint32_t bit = (status - before > 0) ? 1u : 0u;      // :34-37
  ag.novelty[g] = ((ag.novelty[g] << 1) | bit) & tau_mask;   // np.roll(novelty, 1); novelty[0] = bit
  ag.patch_id[g] = res_id;                                   // :39-42 (None -> -1)
}

// radius of agent g (global index): its own (heterogeneous agents, sims.py:502) or the engine-wide one
__device__ __forceinline__ double base_radius_of(const BaseKernelArgs& a, size_t g) {
  return a.ag.radius ? (double)a.ag.radius[g] : a.radius;
}

// Parameter set of agent i of replicate b: one set for the whole batch, one per replicate (sweeps), or one per agent
// (heterogeneous agents, agent.py:83-108).
__device__ __forceinline__ const BaseParams& base_params_of(const BaseKernelArgs& a, int b, int i) {
  return *reinterpret_cast<const BaseParams*>(a.params + (size_t)b * a.param_stride + (size_t)i * a.param_stride_agent);
}

// Field resolution of agent g (heterogeneous agents: each agent's own v_field_res, agent.py:58, 480-481; rows keep the
// engine-wide stride of W words, the bins beyond the agent's resolution stay 0).
__device__ __forceinline__ int base_res_of(const BaseKernelArgs& a, size_t g) { return a.agent_geo ? a.agent_geo[g].res : a.R; }
__device__ __forceinline__ double base_lin_step_of(const BaseKernelArgs& a, size_t g) {
  return a.agent_geo ? a.agent_geo[g].lin_step : a.lin_step;
}

// One WARP per replicate: lanes take the agents of a chunk of 32 in parallel (membership, bias, notify, teleport),
// chunks and patches go in the reference's order, and the one truly sequential piece -- the depletion of a patch by
// its exploiting agents in agent order (sims.py:824-836) -- is replayed by all lanes in lock step over the ballot of
// the chunk's exploiters.  Per agent the notify sequence is the reference's: on a patch destroyed by agent d in this
// step, agents i <= d see (+1, -1), agents i > d see (-1, -1); the second notification of each is the destroy loop's
// (sims.py:829-836), applied after the patch's pass.
__device__ __forceinline__ void base_env_replicate(const BaseKernelArgs& a, int b, int lane, unsigned step) {
  const uint32_t tau_mask = (a.Tau >= 32) ? 0xffffffffu : ((1u << a.Tau) - 1u);
  const size_t a0 = (size_t)b * a.N, p0 = (size_t)b * a.P;
  // The reference visits every patch (sims.py:790-858), but a patch with nobody on it is a no-op, and this loop is the one
  // sequential piece of the step: with the 10 .. 100 patches of the reference's figure experiments most (patch, chunk)
  // iterations find nobody.  The lanes keep their agents' centres in registers (fp32; replicates of up to 128 agents) and
  // reject a whole patch with one compare per chunk and a vote -- conservatively (0.5 px of margin); a patch somebody may
  // stand on goes through the exact float64 code below.  The copies follow the teleports.
  constexpr int kStaged = 4;                    // chunks of 32 agents kept in registers
  const bool staged = a.N <= 32 * kStaged;
  float scx[kStaged], scy[kStaged];
  if (staged) {
#pragma unroll
    for (int c = 0; c < kStaged; ++c) {
      const int i = c * 32 + lane;
      scx[c] = scy[c] = 3.0e18f;                // no agent: far from every patch
      if (i < a.N) {
        const float r = (float)base_radius_of(a, a0 + i);
        scx[c] = a.ag.x[a0 + i] + r; scy[c] = a.ag.y[a0 + i] + r;
      }
    }
  }
  for (int p = 0; p < a.P; ++p) {
    const double prad = a.pa.radius[p0 + p];
    const double pcx = (double)a.pa.x[p0 + p] + prad, pcy = (double)a.pa.y[p0 + p] + prad;
    unsigned near_chunks = 0xffffffffu;         // chunks of 32 agents with somebody near the patch (warp-uniform)
    if (staged) {
      near_chunks = 0u;
      const float fx = (float)pcx, fy = (float)pcy, reach = (float)prad + 0.5f, reach2 = reach * reach;
#pragma unroll
      for (int c = 0; c < kStaged; ++c) {
        const float dx = scx[c] - fx, dy = scy[c] - fy;
        if (__any_sync(0xffffffffu, dx * dx + dy * dy < reach2)) near_chunks |= 1u << c;
      }
      if (!near_chunks) continue;               // nobody can be a member: the iteration below would change nothing
    }
    bool destroy = false;                       // warp-uniform
    double left = a.pa.left[p0 + p];            // warp-uniform running value (stored as float after every take)
    for (int c0 = 0; c0 < a.N; c0 += 32) {
      if (staged && !((near_chunks >> (c0 >> 5)) & 1u)) continue;   // (the teleports of THIS patch go to its centre: already near)
      const int i = c0 + lane;
      const size_t g = a0 + (i < a.N ? i : 0);
      bool member = false;
      const double r = base_radius_of(a, g);       // the agent's own radius: its centre is position + radius
      if (i < a.N) {
        const double ddx = ((double)a.ag.x[g] + r) - pcx, ddy = ((double)a.ag.y[g] + r) - pcy;
        member = sqrt(ddx * ddx + ddy * ddy) < prad;                                 // sims.py:45-56
      }
      if (member) {   // bias_agent_towards_res_center (sims.py:544-552): no wrap of the heading
        const double dx = pcx - ((double)a.ag.x[g] + r), dy = pcy - ((double)a.ag.y[g] + r);
        const double th = a.ag.theta[g];
        double cl = fmod(atan2(dy, dx) + th, ABM_TWO_PI_D);
        if (cl < 0.0) cl += ABM_TWO_PI_D;
        a.ag.theta[g] = (float)(th + (cl - ABM_PI_D) * 0.02);
      }
      // ---- depletion in agent order among this chunk's exploiters (only while the patch still exists) ----
      const bool expl = member && a.ag.override_mode[g] == OV_EXPLOIT;
      uint32_t eb = destroy ? 0u : __ballot_sync(0xffffffffu, expl);
      int d_lane = destroy ? -1 : 32;           // lanes <= d_lane still see the patch; 32: nobody destroyed it (yet)
      double my_take = 0.0;
      const double my_consumption = base_params_of(a, b, i < a.N ? i : 0).consumption;   // agent.consumption
      while (eb) {
        const int l = __ffs(eb) - 1;
        eb &= eb - 1;
        double take = fmin(__shfl_sync(0xffffffffu, my_consumption, l), (double)a.pa.quality[p0 + p]);   // rescource.py:121-122
        if (left >= take) left -= take; else { take = left; left = 0.0; }
        left = (double)(float)left;                                                   // resc_left lives in fp32 state
        if (lane == l) my_take = take;
        if (!(left > 0.0)) { destroy = true; d_lane = l; eb = 0u; }
      }
      if (member) {
        if (lane > d_lane) {
          notify(a.ag, g, -1, -1, tau_mask);                                          // :812-813
        } else {
          notify(a.ag, g, 1, a.pa.id[p0 + p], tau_mask);                              // :816-818 (pooling_time == 0)
          if (a.teleport_exploit) {                                                   // :820-821
            a.ag.x[g] = (float)((double)a.pa.x[p0 + p] + prad - r);
            a.ag.y[g] = (float)((double)a.pa.y[p0 + p] + prad - r);
            if (staged) {                                                             // the register copy follows
#pragma unroll
              for (int c = 0; c < kStaged; ++c)
                if (c0 == 32 * c) { scx[c] = (float)pcx; scy[c] = (float)pcy; }
            }
          }
          if (expl) {                                                                 // :824-828
            const float c = a.ag.collected[g];
            a.ag.collected_before[g] = c;
            a.ag.collected[g] = (float)((double)c + my_take);
          }
        }
        a.ag.mode[g] |= 0x100;   // scratch mark: on a patch this step
      }
    }
    if (lane == 0) a.pa.left[p0 + p] = (float)left;
    if (destroy) {                                                                    // :829-836, for every agent on it
      __syncwarp();
      for (int c0 = 0; c0 < a.N; c0 += 32) {
        const int i = c0 + lane;
        if (i < a.N) {
          const size_t g = a0 + i;
          const double r = base_radius_of(a, g);
          const double ex = ((double)a.ag.x[g] + r) - pcx, ey = ((double)a.ag.y[g] + r) - pcy;
          if (sqrt(ex * ex + ey * ey) < prad) notify(a.ag, g, -1, -1, tau_mask);
        }
      }
    }
    if (destroy && a.regenerate) {   // kill_resource + add_new_resource_patch(force_id) (sims.py:321-374)
      if (lane == 0) {
        bool placed = false;
        for (unsigned t = 0; t < 10000u && !placed; ++t) {      // max_retries, sims.py:335
          const double* rt = a.regen_tab ? a.regen_tab + (size_t)b * 5 : nullptr;   // this replicate's own patch parameters
          const double R_ = rt ? rt[0] : a.patch_radius;
          double nx, ny, q;
          int units;
          if (a.regen_draws) {                                  // injected draws (parity tests): try t of this slot
            if ((int)t >= a.regen_tries) break;
            const double* d = a.regen_draws + (((size_t)b * a.P + p) * a.regen_tries + t) * 4;
            nx = d[0]; ny = d[1]; units = (int)d[2]; q = d[3];
          } else {
            const uint4 rn = philox4x32(make_uint4((uint32_t)b, (uint32_t)p, step, t),
                                        make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ 0x50415443u));
            double lox, hix, loy, hiy;
            if (a.border_overlap) { lox = a.pad - R_; hix = a.width + a.pad - R_; loy = a.pad - R_; hiy = a.height + a.pad - R_; }
            else { lox = a.pad; hix = a.width + a.pad - 2 * R_; loy = a.pad; hiy = a.height + a.pad - 2 * R_; }
            const uint4 rn2 = philox4x32(make_uint4((uint32_t)b, (uint32_t)p, step, t), make_uint2((uint32_t)a.seed, 0x51554c54u));
            nx = floor(lox + floor(hix - lox) * u01(rn.x, rn.y));          // np.random.randint(lo, hi)
            ny = floor(loy + floor(hiy - loy) * u01(rn.z, rn.w));
            const int u_lo = rt ? (int)rt[3] : a.min_units, u_hi = rt ? (int)rt[4] : a.max_units;
            const double q_lo = rt ? rt[1] : a.min_quality, q_hi = rt ? rt[2] : a.max_quality;
            units = u_lo + (int)floor((double)(u_hi - u_lo) * u01(rn2.x, rn2.y));
            q = q_lo + (q_hi - q_lo) * u01(rn2.z, rn2.w);
          }
          bool ok = true;
          for (int p2 = 0; p2 < a.P; ++p2) {                                          // proove_sprite: no patch-patch overlap
            if (p2 == p) continue;
            const double r2 = a.pa.radius[p0 + p2];
            if (!(r2 > 0.0)) continue;                                              // a killed patch is not in the group
            const double ex = (nx + R_) - ((double)a.pa.x[p0 + p2] + r2), ey = (ny + R_) - ((double)a.pa.y[p0 + p2] + r2);
            if (ex * ex + ey * ey <= (R_ + r2) * (R_ + r2)) { ok = false; break; }
          }
          if (!ok) continue;
          a.pa.x[p0 + p] = (float)nx; a.pa.y[p0 + p] = (float)ny; a.pa.radius[p0 + p] = (float)R_;
          a.pa.left[p0 + p] = (float)units; a.pa.quality[p0 + p] = (float)q;
          placed = true;
          atomicAdd(&a.counters[0], 1ull);
        }
        if (!placed) atomicAdd(&a.counters[1], 1ull);
      }
    } else if (destroy) {
      if (lane == 0) a.pa.radius[p0 + p] = 0.0f;   // resource.kill() without regeneration: the patch no longer exists
    }
    __syncwarp();   // the next patch reads what lane 0 wrote (and the agents' new positions / headings)
  }
  // agents on no patch (and not colliding) are told so (sims.py:847-855, pooling_time == 0)
  for (int i = lane; i < a.N; i += 32) {
    const size_t g = a0 + i;
    const int m = a.ag.mode[g];
    if (m & 0x100) a.ag.mode[g] = m & 0xff;
    else if (!a.ag.collided[g]) notify(a.ag, g, -1, -1, tau_mask);
    // frozen snapshot for the agent phase: every agent sees the same positions / modes
    a.ag.snap_x[g] = a.ag.x[g];
    a.ag.snap_y[g] = a.ag.y[g];
    a.ag.snap_override[g] = a.ag.override_mode[g];
  }
}
__global__ void __launch_bounds__(128) base_env_kernel(const BaseKernelArgs a) {
  const int b = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (b >= a.B) return;
  base_env_replicate(a, b, lane, a.step);
}

void launch_base_env(const BaseKernelArgs& a, cudaStream_t stream) {
  const int threads = 128;   // 4 replicates (warps) per CTA
  base_env_kernel<<<(unsigned)(((size_t)a.B * 32 + threads - 1) / threads), threads, 0, stream>>>(a);
}

// ---------------------------------------------------------------------------------------
// agent phase: one warp per focal agent
// ---------------------------------------------------------------------------------------
size_t base_agents_smem_bytes(int N, int W, int warps) { return warp_field_bytes(N, W) * warps; }
int base_agents_warps(int N, int W, size_t smem_limit) {
  int w = 8;
  while (w > 1 && base_agents_smem_bytes(N, W, w) > smem_limit) w >>= 1;
  return w;
}

// Field phase of one focal agent gw = (replicate, agent), run by a whole warp: returns the set bins of the stored field's
// left and right halves (what the decision process needs of it).
__device__ __forceinline__ void base_agent_field(const BaseKernelArgs& a, WarpField& wf, long long gw, int lane, int& n_left_out,
                                                 int& n_right_out) {
  uint32_t* row = wf.row;
  const int b = (int)(gw / a.N), i = (int)(gw - (long long)b * a.N);
  const size_t a0 = (size_t)b * a.N, gi = a0 + i;
  const int W = a.W, N = a.N;
  int R = a.R;
  double lin_step = a.lin_step;
  for (int w = lane; w < W + 1; w += 32) row[w] = 0u;

  const float xi_f = a.ag.snap_x[gi], yi_f = a.ag.snap_y[gi];
  const float th_f = a.ag.theta[gi];
  // heterogeneous radii (sims.py:499-517): the candidate distance is between the agents' own centres (agent.py:400 ->
  // supcalc.distance), the projection uses the FOCAL radius for both centres (agent.py:504-509)
  const double r = a.ag.radius ? (double)a.ag.radius[gi] : a.radius;
  const double cix = __dadd_rn((double)xi_f, r), ciy = __dadd_rn((double)yi_f, r);
  const int my_patch = a.ag.patch_id[gi];
  double fov0 = a.fov0, fov1 = a.fov1, vision_range = a.vision_range;
  int mask_lo = a.mask_lo, mask_hi = a.mask_hi;
  if (a.agent_geo) {                                                                // this agent's own FOV / range
    const BaseAgentGeo g = a.agent_geo[gi];
    fov0 = g.fov0; fov1 = g.fov1; vision_range = g.vision_range; mask_lo = g.mask_lo; mask_hi = g.mask_hi;
    R = g.res; lin_step = g.lin_step;
  }
  const BaseFast bf = base_fast_consts(th_f, r, fov0, fov1, R);

  // ---- candidates, classes, raw intervals (agent.py:396-419, 497-556) ----
  int M = 0;
  for (int j0 = 0; j0 < N; j0 += 32) {
    const int j = j0 + lane;
    bool rec = false;
    ObjRec o; o.s = 0; o.e = 0; o.d = 0.0;
    int k2 = 0;
    if (j < N) {
      const size_t gj = a0 + j;
      const float xj_f = a.ag.snap_x[gj], yj_f = a.ag.snap_y[gj];
      // centre distance in float64 with the reference's operation sequence: it decides the candidate set (<=) and
      // orders the occlusion (strict <), both on the rounded value
      double n2 = base_distance_exact(cix, ciy, r, xj_f, yj_f);                     // both centres with the focal radius
      double n2c = n2;
      if (a.ag.radius) {
        const double rj = a.ag.radius[gj];
        n2c = base_distance_exact(cix, ciy, rj, xj_f, yj_f);                        // own radii (agent.py:400)
      }
      const bool in_range = n2c <= vision_range;                                    // agent.py:400
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
      if (cls != 0 && !((xj_f == xi_f) && (yj_f == yi_f))) {                        // :502
        bool vis;
        if (!base_interval_fast(xj_f - xi_f, yj_f - yi_f, bf, o.s, o.e, vis)) {
          rec = vis; o.d = n2;
        } else {                                                                    // inside a guard band: fp64
          double dist;
          rec = base_interval_cold(xi_f, yi_f, r, th_f, xj_f, yj_f, fov0, fov1, R, lin_step, o.s, o.e, dist);
          o.d = dist;
          atomicAdd(&a.counters[2], 1ull);
        }
        // list order of the reference: social cues, then other occluders, then same-patch
        // exploiters, each in agent order (agent.py:402-410, 472-477)
        k2 = (((cls == 1) ? 0 : (cls == 2 ? 1 : 2)) * N + j) | ((cls == 1) ? (1 << 30) : 0);
      }
    }
    M = base_record(wf, M, rec, o, k2, lane);
  }
  __syncwarp();
  base_occlude_fill(wf, M, a.visual_exclusion != 0, R, lane);   // agent.py:421-445, 569-590
  __syncwarp();

  // ---- flip + FOV mask (agent.py:593-595): stored[b] = v[R-1-b], kept for b in [mask_lo, mask_hi] ----
  const int h = R / 2;                                       // int(V_field_len / 2) (supcalc.py:86-88)
  const int va = R - 1 - mask_hi, vb = R - mask_lo;          // kept bins in v coordinates [va, vb)
  n_left_out = popc_range(row, W, max(va, R - h), min(vb, R), lane);    // stored[0:h]  <-> v[R-h:R]
  n_right_out = popc_range(row, W, max(va, 0), min(vb, R - h), lane);   // stored[h:]   <-> v[0:R-h]
  if (a.fields_out) {
    uint32_t* out = a.fields_out + gi * W;
    for (int ws = lane; ws < W; ws += 32) {
      const int lo = max(mask_lo - (ws << 5), 0), hi = min(mask_hi + 1 - (ws << 5), 32);
      uint32_t m = 0u;
      if (hi > lo) m = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
      out[ws] = m ? (flipped_word(row, 1, R, W, ws) & m) : 0u;   // (words beyond the agent's own resolution: 0)
    }
  }
}

// Decision process, mode machine, kinematics (agent.py:168-283) of one focal agent, fp64, one THREAD per agent: the
// threads 0 .. warps-1 of the CTA take the agents of its warps after a barrier, so that this serial code runs with as many
// lanes as the CTA has agents instead of on lane 0 of every warp.
__device__ __forceinline__ void base_agent_decide(const BaseKernelArgs& a, long long gw, int n_left, int n_right, unsigned step) {
  const int b = (int)(gw / a.N), i = (int)(gw - (long long)b * a.N);
  const size_t gi = (size_t)b * a.N + i;
  const int R = base_res_of(a, gi), h = R / 2;
  const double xi = a.ag.snap_x[gi], yi = a.ag.snap_y[gi], r = base_radius_of(a, gi);   // own radius at the walls
  const BaseParams prm = base_params_of(a, b, i);
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
        const uint4 rn = philox4x32(make_uint4((uint32_t)b, (uint32_t)i, step, 0u),
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
  atomicAdd(&a.mode_steps[(size_t)b * 4 + mode], 1u);        // the mode this agent is logged with at the end of the step
}

__global__ void __launch_bounds__(256, 3) base_agent_kernel(const BaseKernelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int halves[2 * 8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const long long total = (long long)a.B * a.N;
  const long long gw = (long long)blockIdx.x * wpb + wib;   // global warp = (replicate, focal agent)
  if (gw < total) {
    WarpField wf = warp_field_at(smem_raw + warp_field_bytes(a.N, a.W) * wib, a.N);
    int n_left = 0, n_right = 0;
    base_agent_field(a, wf, gw, lane, n_left, n_right);
    if (lane == 0) { halves[2 * wib] = n_left; halves[2 * wib + 1] = n_right; }
  }
  __syncthreads();
  if (threadIdx.x < wpb) {
    const long long g2 = (long long)blockIdx.x * wpb + threadIdx.x;
    if (g2 < total) base_agent_decide(a, g2, halves[2 * threadIdx.x], halves[2 * threadIdx.x + 1], a.step);
  }
}

// ---------------------------------------------------------------------------------------
// collision phase (sims.py:736-783, 421-468): the colliding (a1, a2) pairs in group order -- an
// agent that collides with several others is turned once per partner, each time from its
// already-turned heading.  PARITY UNPINNED for the
// pair detection (pygame.sprite.collide_circle on int-truncated rect centres, not in the
// reference tree); the proximity field itself is Agent.projection_field (agent.py:457-597).
// ---------------------------------------------------------------------------------------
// Mapping: one CTA per replicate, its warps take the HIT agents a2 = warp, warp + warps, ...  The reference walks the
// ordered pairs (a1, a2) sequentially, but everything an event (a1, a2) reads besides a2's own heading is fixed during
// the phase -- positions do not move, and the only override mode that is tested, "exploit", is neither set nor cleared
// here -- so the events of different a2 are independent and only those of the same a2 (its heading turns with every
// hit) have to stay in a1 order.  `collided_agents` is a set: the warps mark its members with plain stores.
__device__ __forceinline__ void base_collision_replicate(const BaseKernelArgs& a, int b, unsigned char* smem_raw) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int N = a.N, W = a.W;
  const size_t per_warp = warp_field_bytes(N, W);
  WarpField wf = warp_field_at(smem_raw + per_warp * wib, N);
  int* ov = reinterpret_cast<int*>(smem_raw + per_warp * wpb);          // override_mode (entry a2: its warp's)
  int* md = ov + N;                                                      // mode
  float* th = reinterpret_cast<float*>(md + N);                          // heading
  int* col = reinterpret_cast<int*>(th + N);                             // member of collided_agents
  float* px = reinterpret_cast<float*>(col + N);                         // positions (this phase does not move anybody)
  float* py = px + N;
  float* rad = py + N;                                                   // radii (own: heterogeneous agents)
  const uint32_t tau_mask = (a.Tau >= 32) ? 0xffffffffu : ((1u << a.Tau) - 1u);
  const size_t a0 = (size_t)b * N;
  int* work = reinterpret_cast<int*>(rad + N);                           // [N] hit agents, then: front count, back count, next
  // "exploits at the start of the phase" as its own read-only array: ov[] of a hit agent is written by ITS warp
  // (-> collide, never to or from exploit) while other warps test their hitting agents for exploit -- the answer cannot
  // change, but reading it from ov[] was a read/write hazard between warps (racecheck)
  int* ex = work + N + 4;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    ov[i] = a.ag.override_mode[a0 + i]; md[i] = a.ag.mode[a0 + i]; th[i] = a.ag.theta[a0 + i]; col[i] = 0;
    px[i] = a.ag.x[a0 + i]; py[i] = a.ag.y[a0 + i]; rad[i] = (float)base_radius_of(a, a0 + i);
    ex[i] = (ov[i] == OV_EXPLOIT) ? 1 : 0;
  }
  __syncthreads();
  // pygame's collide_circle (documented): circles of the sprites' radii (+2 each, sims.py:739-752) around the rect
  // centres = int-truncated position + half the integer rect size
  auto hits = [&](int p, int q) {
    const float dx = (truncf(px[p]) + truncf(rad[p])) - (truncf(px[q]) + truncf(rad[q]));
    const float dy = (truncf(py[p]) + truncf(rad[p])) - (truncf(py[q]) + truncf(rad[q]));
    const float rs = rad[p] + rad[q] + 4.0f;
    return dx * dx + dy * dy <= rs * rs;
  };

  // hit agents first (cheap), then the warps pull them from a list: the events of one hit agent are a sequential chain
  // of fp64 latency, and with a fixed assignment the CTA waited at the final barrier for its unluckiest warp (more than
  // half of the kernel's warp time)
  // (agents hit by several others -- the long chains -- are listed from the front and taken first, the rest from the back)
  if (threadIdx.x == 0) { work[N] = 0; work[N + 1] = 0; work[N + 2] = 0; }
  __syncthreads();
  for (int a2 = wib; a2 < N; a2 += wpb) {
    int n_hit = 0;
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int jj = j0 + lane;
      const bool hit = jj < N && jj != a2 && hits(jj, a2);
      n_hit += __popc(__ballot_sync(0xffffffffu, hit));
    }
    if (lane == 0 && n_hit >= 2) work[atomicAdd(&work[N], 1)] = a2;
    if (lane == 0 && n_hit == 1) work[N - 1 - atomicAdd(&work[N + 1], 1)] = a2;
  }
  __syncthreads();
  const int n_front = work[N], n_work = n_front + work[N + 1];
  for (;;) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(&work[N + 2], 1);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= n_work) break;
    const int a2 = work[slot < n_front ? slot : N - 1 - (slot - n_front)];
    const bool expl2 = ex[a2] != 0;                                      // (fixed during the phase)
    const double r = rad[a2];                                              // the hit agent is the focal agent of its LIDAR field
    const int R = base_res_of(a, a0 + a2), h = R / 2;                      // ... at its own resolution
    const double lin_step = base_lin_step_of(a, a0 + a2);
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int jj = j0 + lane;
      const bool hit = jj < N && jj != a2 && hits(jj, a2);                 // rect.x = int(position) (agent.py:303-304)
      unsigned hitm = __ballot_sync(0xffffffffu, hit);
      while (hitm) {                                                       // the agents a1 that hit a2, in group order
        const int a1 = j0 + __ffs(hitm) - 1;
        hitm &= hitm - 1;
        const bool expl1 = ex[a1] != 0;
        // ---- agent_agent_collision_proximity(a1, a2) (sims.py:421-468) ----
        bool do_coll = true;
        if (a.ghost_mode) do_coll = !expl2 && !expl1;
        if (do_coll) {
          __syncwarp();
          if (lane == 0 && !expl2) { ov[a2] = OV_COLLIDE; md[a2] = MODE_COLLIDE; }   // :442-443
          for (int w = lane; w < W + 1; w += 32) wf.row[w] = 0u;
          const float x2 = px[a2], y2 = py[a2], th2 = th[a2];
          const double cix = __dadd_rn((double)x2, r), ciy = __dadd_rn((double)y2, r);
          const BaseFast bf = base_fast_consts(th2, r, -ABM_PI_D, ABM_PI_D, R);
          int M = 0, last_j = -1;
          double last_d = 0.0;
          for (int v0 = 0; v0 < N; v0 += 32) {
            const int j = v0 + lane;
            bool rec = false, counted = false;
            ObjRec o; o.s = 0; o.e = 0; o.d = 0.0;
            double dist = 0.0;
            if (j < N && j != a2) {
              const float xj = px[j], yj = py[j];
              // vicinity: distance between the agents' OWN centres (supcalc.distance); the field's distance: both
              // centres with the focal radius (agent.py:504-509)
              const double n2 = base_distance_exact(cix, ciy, r, xj, yj);
              const double n2own = a.ag.radius ? base_distance_exact(cix, ciy, (double)rad[j], xj, yj) : n2;
              if (n2own < 2.0 * r + 20.0) {                                         // :446-447
                counted = !((xj == x2) && (yj == y2));
                if (counted) {
                  dist = n2;
                  bool vis;
                  if (!base_interval_fast(xj - x2, yj - y2, bf, o.s, o.e, vis)) {
                    rec = vis; o.d = n2;
                  } else {                                                          // inside a guard band: fp64
                    rec = base_interval_cold(x2, y2, r, th2, xj, yj, -ABM_PI_D, ABM_PI_D, R, lin_step, o.s, o.e, dist);
                    o.d = dist;
                  }
                }
              }
            }
            // the loop variable `distance` that keep_distance_info leaks (agent.py:590): last obstacle in list order
            const unsigned cm = __ballot_sync(0xffffffffu, counted);
            if (cm) {
              const int src = 31 - __clz(cm);
              last_j = v0 + src;
              last_d = __shfl_sync(0xffffffffu, dist, src);
            }
            M = base_record(wf, M, rec, o, j | (1 << 30), lane);
          }
          __syncwarp();
          base_occlude_fill(wf, M, a.visual_exclusion != 0, R, lane);
          const int n_left = popc_range(wf.row, W, R - h, R, lane);                 // stored[0:h]
          const int n_right = popc_range(wf.row, W, 0, R - h, lane);                // stored[h:]
          int flo = h - 100, fhi = h + 100;                                         // :462 with numpy slice semantics
          if (flo < 0) flo = max(flo + R, 0);
          flo = min(flo, R); fhi = min(fhi, R);
          const int n_front = (fhi > flo) ? popc_range(wf.row, W, R - fhi, R - flo, lane) : 0;
          if (lane == 0) {
            const double vr2 = a.agent_geo ? a.agent_geo[a0 + a2].vision_range : a.vision_range;   // agent2's own
            const double amp = (last_j >= 0) ? 1.0 - last_d / vr2 : 1.0;
            const double left = amp * (double)n_left / (double)h, right = amp * (double)n_right / (double)(R - h);
            double D = (left > right) ? 1.0 : ((left < right) ? -1.0 : 0.0);
            if (D == 0.0) D = -1.0;                                                  // :456
            if (!expl2) th[a2] = (float)((double)th[a2] - D * 0.2);                 // :458-459 (not wrapped)
            if (amp * (double)n_front > 0.0) a.ag.vel[a0 + a2] = 0.0f;              // :462-463
            else if (!expl2) a.ag.vel[a0 + a2] = (float)base_params_of(a, b, a2).exp_vel_max;   // agent2.max_exp_vel (:465)
          }
          __syncwarp();
        }
        // ---- collided_agents bookkeeping (sims.py:759-776) ----
        if (lane == 0) {
          if (a.teleport_exploit) { if (!expl1) col[a1] = 1; if (!expl2) col[a2] = 1; }
          else if (!a.ghost_mode) { col[a1] = 1; col[a2] = 1; }
          else if (!expl1 && !expl2) { col[a1] = 1; col[a2] = 1; }
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {                               // sims.py:778-783
    const size_t g = a0 + i;
    if (!col[i] && ov[i] == OV_COLLIDE) { ov[i] = OV_NONE; md[i] = MODE_EXPLORE; }
    if (col[i] && ov[i] == OV_COLLIDE) notify(a.ag, g, -1, -1, tau_mask);
    a.ag.override_mode[g] = ov[i]; a.ag.mode[g] = md[i]; a.ag.theta[g] = th[i]; a.ag.collided[g] = col[i];
  }
}
__global__ void __launch_bounds__(512) base_collision_kernel(const BaseKernelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  base_collision_replicate(a, blockIdx.x, smem_raw);
}

// ---------------------------------------------------------------------------------------
// agent phase of one replicate by the whole CTA (N <= 128 agents, R <= 8192): the work is dealt to the THREADS by item,
// not to warps by focal agent -- (focal, object) pairs, then (focal, social cue) occlusion scans, then (focal, word)
// popcounts -- so every lane has work whatever N is (a warp per focal agent runs 49 objects as two ragged passes of 32
// lanes and spends most of its instructions on per-focal bookkeeping).  Focal agents go in groups of kBlkGroup (the pair
// table of a group lives in shared memory).
// ---------------------------------------------------------------------------------------
constexpr int kBlkMaxGroup = 32;    // focal agents per group (the pair table of a group lives in shared memory)
constexpr int kBlkMaxN = 128;       // objects per focal agent fit a 128-bit mask (the reference's figure experiments: N <= 100)

struct BlkShared {
  float *x, *y, *th;                // [N] frozen snapshot positions, headings
  int *ov, *pid;                    // [N] snapshot override mode, patch id
  float* rad;                       // [N] per-agent radius (or the engine-wide one)
  BaseFast* bf;                     // [G] fast-path constants of the group's focal agents
  double2* fovd;                    // [G] (fov0, fov1) in float64 (fp64 fallback)
  int2* mask;                       // [G] stored bins kept by the FOV mask (mask_lo, mask_hi)
  int* res;                         // [G] the focal agent's own field resolution (<= R)
  double* lin;                      // [G] its linspace step
  float* vr2;                       // [G] vision range squared
  uint32_t* se;                     // [G][N] raw interval ends: (uint16)s | (uint16)e << 16
  double* d;                        // [G][N] centre distance, float64 with the reference's operation sequence (:526-528):
                                    // what the occlusion orders by (strict <, stable in list order)
  unsigned char* cls;               // [G][N] 0 not recorded, 1 social cue, 2 occluder, 3 occluder (same-patch exploiter)
  unsigned short* cues;             // [G * N] list of (focal slot << 8 | object) of the group's social cues
  uint32_t* rows;                   // [G][W + 1] un-flipped fields
  int* halves;                      // [N][2] set bins of the stored field's left / right half
  int* ncue;                        // [1]
};
size_t base_block_smem_bytes(int N, int W, int G) {
  size_t b = 6 * 4 * (size_t)N;                                        // x, y, th, ov, pid, rad
  b = (b + 15) / 16 * 16 + sizeof(BaseFast) * G + 16 * G + 8 * G + 8 * G + 4 * G + 4 * G;   // bf, fovd, lin, mask, vr2, res
  b = (b + 15) / 16 * 16 + (8 + 4 + 1 + 2) * (size_t)G * N + 16;        // d, se, cls, cues
  b = (b + 15) / 16 * 16 + 4 * (size_t)G * (W + 1) + 8 * (size_t)N + 16;
  return b + 64;
}
// focal agents per group: equal groups, as few as fit `budget` bytes of shared memory per CTA
int base_block_group(int N, int W, size_t budget) {
  for (int n_groups = 1; n_groups <= N; ++n_groups) {
    const int G = (N + n_groups - 1) / n_groups;
    if (G <= kBlkMaxGroup && base_block_smem_bytes(N, W, G) <= budget) return G;
  }
  return 1;
}
__device__ __forceinline__ BlkShared base_block_carve(unsigned char* p, int N, int W, int G) {
  auto up = [](unsigned char* q) { return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(q) + 15) & ~uintptr_t(15)); };
  BlkShared s;
  s.x = reinterpret_cast<float*>(p); s.y = s.x + N; s.th = s.y + N;
  s.ov = reinterpret_cast<int*>(s.th + N); s.pid = s.ov + N; s.rad = reinterpret_cast<float*>(s.pid + N);
  p = up(reinterpret_cast<unsigned char*>(s.rad + N));
  s.fovd = reinterpret_cast<double2*>(p); p += 16 * G;
  s.lin = reinterpret_cast<double*>(p); p += 8 * G;
  s.bf = reinterpret_cast<BaseFast*>(p); p += sizeof(BaseFast) * G;
  s.mask = reinterpret_cast<int2*>(p); p += 8 * G;
  s.vr2 = reinterpret_cast<float*>(p); p += 4 * G;
  s.res = reinterpret_cast<int*>(p); p += 4 * G;
  p = up(p);
  s.d = reinterpret_cast<double*>(p); p += 8 * (size_t)G * N;
  s.se = reinterpret_cast<uint32_t*>(p); p += 4 * (size_t)G * N;
  s.cues = reinterpret_cast<unsigned short*>(p); p += 2 * (size_t)G * N;
  s.cls = p; p += (size_t)G * N;
  p = up(p);
  s.rows = reinterpret_cast<uint32_t*>(p); p += 4 * (size_t)G * (W + 1);
  s.halves = reinterpret_cast<int*>(p); p += 8 * (size_t)N;
  s.ncue = reinterpret_cast<int*>(p);
  return s;
}

// MASK2: replicates of 65 .. 128 agents (two 64-bit words for the set of closer objects; a kernel of its own, so that the
// usual sizes do not pay for the second word: 7 % of the configs[2] step)
template <bool MASK2>
__device__ __forceinline__ void base_agents_block(const BaseKernelArgs& a, int b, unsigned char* smem_raw, unsigned step,
                                                  const int G) {
  const int N = a.N, W = a.W;
  const int tid = threadIdx.x, T = blockDim.x;
  const size_t a0 = (size_t)b * N;
  BlkShared sh = base_block_carve(smem_raw, N, W, G);
  for (int i = tid; i < N; i += T) {
    sh.x[i] = a.ag.snap_x[a0 + i]; sh.y[i] = a.ag.snap_y[a0 + i]; sh.th[i] = a.ag.theta[a0 + i];
    sh.ov[i] = a.ag.snap_override[a0 + i]; sh.pid[i] = a.ag.patch_id[a0 + i];
    sh.rad[i] = a.ag.radius ? a.ag.radius[a0 + i] : (float)a.radius;
    sh.halves[2 * i] = 0; sh.halves[2 * i + 1] = 0;
  }
  for (int g0 = 0; g0 < N; g0 += G) {
    const int ng = min(G, N - g0);
    __syncthreads();                                         // staging done / previous group's rows and table are free
    // ---- per focal agent of the group: constants of the fast path, FOV, zeroed row ----
    if (tid < ng) {
      const int i = g0 + tid;
      double fov0 = a.fov0, fov1 = a.fov1, vr = a.vision_range;
      int mlo = a.mask_lo, mhi = a.mask_hi, R = a.R;
      double lin = a.lin_step;
      if (a.agent_geo) {
        const BaseAgentGeo q = a.agent_geo[a0 + i];
        fov0 = q.fov0; fov1 = q.fov1; vr = q.vision_range; mlo = q.mask_lo; mhi = q.mask_hi; R = q.res; lin = q.lin_step;
      }
      sh.bf[tid] = base_fast_consts(sh.th[i], a.ag.radius ? (double)sh.rad[i] : a.radius, fov0, fov1, R);
      sh.res[tid] = R; sh.lin[tid] = lin;
      sh.fovd[tid] = make_double2(fov0, fov1);
      sh.mask[tid] = make_int2(mlo, mhi);
      sh.vr2[tid] = (float)(vr * vr);
    }
    for (int w = tid; w < ng * (W + 1); w += T) sh.rows[w] = 0u;
    if (tid == 0) *sh.ncue = 0;
    __syncthreads();
    // ---- (focal, object) pairs: class and raw interval (agent.py:396-419, 497-556) ----
    for (int it = tid; it < ng * N; it += T) {
      const int gi = it / N, j = it - gi * N, i = g0 + gi;
      const float xi = sh.x[i], yi = sh.y[i], xj = sh.x[j], yj = sh.y[j];
      const double ri = a.ag.radius ? (double)sh.rad[i] : a.radius;
      const float dxf = xj - xi, dyf = yj - yi;
      const float d2 = fmaf(dxf, dxf, dyf * dyf);
      // candidate test (agent.py:400) on the distance between the agents' OWN centres
      bool in_range;
      {
        float dc2 = d2;
        if (a.ag.radius) { const float dr = sh.rad[j] - sh.rad[i]; const float ex = dxf + dr, ey = dyf + dr; dc2 = fmaf(ex, ex, ey * ey); }
        const float v2 = sh.vr2[gi];
        if (dc2 < v2 * (1.0f - 2.0e-6f)) in_range = true;
        else if (dc2 > v2 * (1.0f + 2.0e-6f)) in_range = false;
        else {   // within rounding of the range: float64, the reference's operation sequence
          const double rj = a.ag.radius ? (double)sh.rad[j] : a.radius;
          const double vr = a.agent_geo ? a.agent_geo[a0 + i].vision_range : a.vision_range;
          in_range = base_distance_exact(__dadd_rn((double)xi, ri), __dadd_rn((double)yi, ri), rj, xj, yj) <= vr;
        }
      }
      const bool is_expl = (j != i) && (sh.ov[j] == OV_EXPLOIT);                     // :402-403
      int cls = 0;
      if (in_range) {
        if (is_expl) {
          const int pj = sh.pid[j];
          if (a.patchwise_exclusion && pj == sh.pid[i]) cls = 3;                     // :406-409
          else if (pj != -1) cls = 1;                                                // :410, :413
        } else {
          cls = 2;                                                                   // :405
        }
      }
      if (!a.visual_exclusion && cls != 1) cls = 0;                                  // :415-419
      uint32_t se = 0u;
      double dist_ij = 0.0;
      if (cls != 0) {
        if ((xj == xi) && (yj == yi)) cls = 0;                                       // :502 (self, coincident positions)
        else {
          int s_, e_;
          bool vis;
          // (every pair's distance here, in parallel: the occlusion scans below only compare)
          if (a.visual_exclusion)
            dist_ij = base_distance_exact(__dadd_rn((double)xi, ri), __dadd_rn((double)yi, ri), ri, xj, yj);
          if (base_interval_fast(dxf, dyf, sh.bf[gi], s_, e_, vis)) {                // inside a guard band: float64
            double dist;
            vis = base_interval_cold(xi, yi, ri, sh.th[i], xj, yj, sh.fovd[gi].x, sh.fovd[gi].y, sh.res[gi], sh.lin[gi], s_, e_, dist);
            atomicAdd(&a.counters[2], 1ull);
          }
          if (!vis) cls = 0;
          se = ((uint32_t)s_ & 0xffffu) | ((uint32_t)e_ << 16);
        }
      }
      sh.se[it] = se; sh.d[it] = dist_ij; sh.cls[it] = (unsigned char)cls;
      if (cls == 1) sh.cues[atomicAdd(sh.ncue, 1)] = (unsigned short)((gi << 8) | j);
    }
    __syncthreads();
    // ---- (focal, social cue): occlusion by the strictly closer objects that meet the cue's interval, applied in
    //      (distance, list order) order (agent.py:421-445), then the fill (agent.py:569-590) ----
    const int n_cues = *sh.ncue;
    for (int c = tid; c < n_cues; c += T) {
      const int gi = sh.cues[c] >> 8, jc = sh.cues[c] & 0xff;
      const uint32_t sec = sh.se[gi * N + jc];
      const int fs = (int)(short)(sec & 0xffffu), fe_ = (int)(short)(sec >> 16);
      int sx = fs, ex = fe_;
      if (a.visual_exclusion) {
        const double dc = sh.d[gi * N + jc];
        unsigned long long rel0 = 0ull, rel1 = 0ull;         // objects strictly closer than the cue that meet its raw interval
        const unsigned char* cl = sh.cls + gi * N;           // (two words: N <= 128)
        const uint32_t* sep = sh.se + gi * N;
        const double* dp = sh.d + gi * N;
#pragma unroll 2
        for (int j = 0; j < N; ++j) {
          const uint32_t so = sep[j];
          const int os = (int)(short)(so & 0xffffu), oe = (int)(short)(so >> 16);
          if ((cl[j] != 0) & (os <= fe_) & (oe >= fs) & (j != jc)) {
            if (dp[j] < dc) { if (!MASK2 || j < 64) rel0 |= 1ull << j; else rel1 |= 1ull << (j - 64); }   // :430 strict, on the float64 values
          }
        }
        while (MASK2 ? (rel0 | rel1) != 0ull : rel0 != 0ull) {   // usually none to three
          int best = -1;
          if (MASK2 ? (__popcll(rel0) + __popcll(rel1) == 1) : ((rel0 & (rel0 - 1ull)) == 0ull)) {
            best = (!MASK2 || rel0) ? __ffsll((long long)rel0) - 1 : 63 + __ffsll((long long)rel1);
          } else {                                           // several: nearest first, ties in list order (stable sort, :424)
            double bd = 0.0; int bk = 0;
#pragma unroll
            for (int half = 0; half < (MASK2 ? 2 : 1); ++half) {
              for (unsigned long long m = half ? rel1 : rel0; m; m &= m - 1ull) {
                const int j = 64 * half + __ffsll((long long)m) - 1;
                const double dj = dp[j];
                const int cj = cl[j];
                const int kj = ((cj == 1) ? 0 : (cj == 2 ? 1 : 2)) * N + j;
                if (best < 0 || dj < bd || (dj == bd && kj < bk)) { best = j; bd = dj; bk = kj; }
              }
            }
          }
          if (!MASK2 || best < 64) rel0 &= ~(1ull << best); else rel1 &= ~(1ull << (best - 64));
          const uint32_t so = sep[best];
          const int os = (int)(short)(so & 0xffffu), oe = (int)(short)(so >> 16);
          if (sx <= os && os <= ex) ex = os;                                          // :432-433
          if (sx <= oe && oe <= ex) sx = oe;                                          // :435-436
          if (os <= sx && oe >= ex) { sx = 0; ex = 0; }                               // :438-440
        }
      }
      base_draw(sh.rows + (size_t)gi * (W + 1), sh.res[gi], sx, ex);
    }
    __syncthreads();
    // ---- (focal, word): set bins of the stored field's halves (flip + FOV mask, agent.py:593-595; supcalc.py:86-91)
    //      and the packed stored field ----
    for (int it = tid; it < ng * W; it += T) {
      const int gi = it / W, w = it - gi * W, i = g0 + gi;
      const uint32_t* row = sh.rows + (size_t)gi * (W + 1);
      const int mlo = sh.mask[gi].x, mhi = sh.mask[gi].y;
      const int R = sh.res[gi], h = R / 2;                   // int(V_field_len / 2) (supcalc.py:86-88)
      const int va = R - 1 - mhi, vb = R - mlo;              // kept bins in v coordinates [va, vb)
      const uint32_t word = row[w];
      if (word) {
        auto cnt = [&](int lo_, int hi_) {
          const int lo = max(lo_ - (w << 5), 0), hi = min(hi_ - (w << 5), 32);
          if (hi <= lo) return 0;
          const uint32_t m = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
          return __popc(word & m);
        };
        const int nl = cnt(max(va, R - h), min(vb, R));      // stored[0:h]  <-> v[R-h:R]
        const int nr = cnt(max(va, 0), min(vb, R - h));      // stored[h:]   <-> v[0:R-h]
        if (nl) atomicAdd(&sh.halves[2 * i], nl);
        if (nr) atomicAdd(&sh.halves[2 * i + 1], nr);
      }
      if (a.fields_out) {
        const int lo = max(mlo - (w << 5), 0), hi = min(mhi + 1 - (w << 5), 32);
        uint32_t m = 0u;
        if (hi > lo) m = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
        a.fields_out[(a0 + i) * W + w] = m ? (flipped_word(row, 1, R, W, w) & m) : 0u;   // (beyond the agent's resolution: 0)
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < N; i += T) base_agent_decide(a, (long long)a0 + i, sh.halves[2 * i], sh.halves[2 * i + 1], step);
}

// ---------------------------------------------------------------------------------------
// the fused step: ONE launch per time step, one CTA per replicate (sims.py:733-864 in its own order: collisions,
// agent-patch interaction, Agent.update of every agent from one frozen snapshot).  The phases of a replicate are
// sequential by construction and each of them is a chain of dependent latencies; with a CTA per replicate and all
// replicates of a sweep resident at once, the SMs overlap the chains of different replicates instead of running three
// grids one after the other (each with its own launch and its own tail).
// ---------------------------------------------------------------------------------------
size_t base_step_smem_bytes(int N, int W, int warps) {
  // the collision phase's layout (base_collision_replicate: the warps' fields, nine per-agent arrays, 4 counters) + N spare
  return warp_field_bytes(N, W) * warps + 9 * sizeof(int) * (size_t)N + 4 * sizeof(int) + sizeof(int) * (size_t)N;
}

template <bool MASK2>
__global__ void __launch_bounds__(128, 7) base_step_kernel(const __grid_constant__ BaseKernelArgs a, unsigned phases, int collide, int n_steps,
                                                           int blk_group) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x, N = a.N;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const size_t a0 = (size_t)b * N;
  // Replicates never interact: the CTA runs ALL n_steps time steps of its replicate in this one launch (one barrier
  // between steps instead of a launch; a single small run -- config 1 -- is bound by nothing else).
  for (int s = 0; s < n_steps; ++s) {
    const unsigned step = a.step + (unsigned)s;
    if (s) __syncthreads();
    if (collide) {                                   // sims.py:736-783
      base_collision_replicate(a, b, smem_raw);
      __syncthreads();
    }
    if (phases & 1u) {                               // sims.py:790-858 (+ the frozen snapshot of the agent phase)
      if (wib == 0) base_env_replicate(a, b, lane, step);
    } else {                                         // agent phase alone: the snapshot is the current state
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        a.ag.snap_x[a0 + i] = a.ag.x[a0 + i];
        a.ag.snap_y[a0 + i] = a.ag.y[a0 + i];
        a.ag.snap_override[a0 + i] = a.ag.override_mode[a0 + i];
      }
    }
    __syncthreads();
    if (!(phases & 2u)) continue;
    // sims.py:861: Agent.update of every agent from the frozen snapshot
    base_agents_block<MASK2>(a, b, smem_raw, step, blk_group);   // the whole CTA, work dealt by item
  }
}

// opt-in shared memory per block of the current device, asked once per device (the attribute query costs more host time
// than a launch, and a foraging step is three launches of a few hundred microseconds together)
static int base_smem_optin() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!cached[dev]) {
    int v = 48 * 1024;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cached[dev] = v;
  }
  return cached[dev];
}

void launch_base_collisions(const BaseKernelArgs& a, cudaStream_t stream) {
  const int smem_max = base_smem_optin();
  const size_t per_warp = warp_field_bytes(a.N, a.W), shared = 9 * sizeof(int) * (size_t)a.N + 4 * sizeof(int);
  int warps = 16;
  while (warps > 1 && (warps / 2 >= a.N || per_warp * warps + shared > (size_t)smem_max)) warps >>= 1;
  const size_t smem = per_warp * warps + shared;
  static SmemOptIn optin;   // per device (abm_common.cuh)
  optin.ensure(base_collision_kernel, smem);
  base_collision_kernel<<<a.B, warps * 32, smem, stream>>>(a);
}

// ---------------------------------------------------------------------------------------
// stateless function-level kernels
// ---------------------------------------------------------------------------------------
__global__ void base_projection_kernel(const BaseProjArgs a) {   // one warp
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const int n = a.n_social + a.n_occ;
  WarpField wf = warp_field_at(smem_raw, n > 0 ? n : 1);
  for (int w = lane; w < a.W + 1; w += 32) wf.row[w] = 0u;
  const double cix = __dadd_rn((double)a.fx, a.radius), ciy = __dadd_rn((double)a.fy, a.radius);
  const BaseFast bf = base_fast_consts(a.ftheta, a.radius, a.fov0, a.fov1, a.R);
  int M = 0, last_j = -1;
  double last_d = 0.0;
  const int n_used = a.visual_exclusion ? n : a.n_social;            // agent.py:475-477, :415-419
  for (int j0 = 0; j0 < n_used; j0 += 32) {
    const int j = j0 + lane;
    bool rec = false, counted = false;
    ObjRec o; o.s = 0; o.e = 0; o.d = 0.0;
    double dist = 0.0;
    if (j < n_used) {
      counted = !((a.ox[j] == a.fx) && (a.oy[j] == a.fy));
      if (counted) {
        dist = base_distance_exact(cix, ciy, a.radius, a.ox[j], a.oy[j]);
        bool vis;
        if (!base_interval_fast(a.ox[j] - a.fx, a.oy[j] - a.fy, bf, o.s, o.e, vis)) {
          rec = vis; o.d = dist;
        } else {                                                      // inside a guard band: fp64
          rec = base_interval_cold(a.fx, a.fy, a.radius, a.ftheta, a.ox[j], a.oy[j], a.fov0, a.fov1, a.R, a.lin_step, o.s,
                                   o.e, dist);
          o.d = dist;
        }
      }
    }
    const unsigned cm = __ballot_sync(0xffffffffu, counted);
    if (cm) { const int src = 31 - __clz(cm); last_j = j0 + src; last_d = __shfl_sync(0xffffffffu, dist, src); }
    M = base_record(wf, M, rec, o, j | ((j < a.n_social) ? (1 << 30) : 0), lane);
  }
  __syncwarp();
  base_occlude_fill(wf, M, a.visual_exclusion != 0, a.R, lane);
  for (int ws = lane; ws < a.W; ws += 32) {
    uint32_t word = flipped_word(wf.row, 1, a.R, a.W, ws);
    const int lo = max(a.mask_lo - (ws << 5), 0), hi = min(a.mask_hi + 1 - (ws << 5), 32);
    uint32_t m = 0u;
    if (hi > lo) m = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
    a.field[ws] = word & m;
  }
  if (lane == 0) *a.amplitude = (a.keep_distance && last_j >= 0) ? 1.0 - last_d / a.vision_range : 1.0;
}
void launch_base_projection(const BaseProjArgs& a, cudaStream_t stream) {
  const int n = a.n_social + a.n_occ;
  const size_t smem = warp_field_bytes(n > 0 ? n : 1, a.W);
  cudaFuncSetAttribute(base_projection_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  base_projection_kernel<<<1, 32, smem, stream>>>(a);
}

__global__ void base_reloc_lr_kernel(const uint32_t* field, int R, int W, double amp, double vel, double vdes,
                                     double thmax, double* out2) {   // one warp; field in STORED order
  const int lane = threadIdx.x, h = R / 2;
  const int n_left = popc_range(field, W, 0, h, lane), n_right = popc_range(field, W, h, R, lane);
  if (lane == 0) {
    out2[0] = vdes - vel;                                                           // supcalc.py:92
    out2[1] = (amp * (double)n_left / (double)h - amp * (double)n_right / (double)(R - h)) * thmax;   // :86-91
  }
}
void launch_base_reloc_lr(const uint32_t* field, int R, int W, double amp, double vel, double vdes, double thmax,
                          double* out2, cudaStream_t stream) {
  base_reloc_lr_kernel<<<1, 32, 0, stream>>>(field, R, W, amp, vel, vdes, thmax, out2);
}

__global__ void vf_dphi_kernel(const uint32_t* v, int R, int W, signed char* out) {   // vf_supcalc.py:257-277
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= R) return;
  auto bit = [&](int i) { return (int)((v[i >> 5] >> (i & 31)) & 1u); };
  const bool backward = bit(0) - bit(R - 1) > 0;
  out[k] = backward ? (signed char)(bit(k) - bit(k == 0 ? R - 1 : k - 1))
                    : (signed char)(bit(k == R - 1 ? 0 : k + 1) - bit(k));
}
void launch_vf_dphi(const uint32_t* v, int R, int W, signed char* out, cudaStream_t stream) {
  vf_dphi_kernel<<<(R + 127) / 128, 128, 0, stream>>>(v, R, W, out);
}

// One launch for the whole step when a CTA per replicate fills the GPU (sweeps); false: the caller launches the phases
// as separate grids (few replicates of many agents: a warp per focal agent over the whole GPU).
bool launch_base_step(const BaseKernelArgs& a, unsigned phases, bool collide, int n_steps, int n_sms, cudaStream_t stream) {
  const int smem_max = base_smem_optin();
  // 4 warps per replicate, 72 registers: 7 CTAs per SM, so the 1024 replicates of a sweep are resident all at once and the
  // SMs interleave their (latency-bound, sequential) phase chains
  int warps = 4;
  while (warps > 1 && (warps / 2 >= a.N || base_step_smem_bytes(a.N, a.W, warps) > (size_t)smem_max)) warps >>= 1;
  // (larger replicates, or resolutions beyond the packed interval ends: one grid per phase, a warp per focal agent)
  if (a.N > kBlkMaxN || a.R > 8192) return false;
  size_t smem = base_step_smem_bytes(a.N, a.W, warps);
  // agent phase by the whole CTA: as few groups of focal agents as 7 resident CTAs per SM (1024 replicates on 148 SMs) leave room for
  const int blk_group = base_block_group(a.N, a.W, (size_t)(224 * 1024) / 7 - 1024);
  smem = std::max(smem, base_block_smem_bytes(a.N, a.W, blk_group));
  if (smem > (size_t)smem_max) return false;
  static SmemOptIn optin, optin2;
  if (a.N > 64) {
    optin2.ensure(base_step_kernel<true>, smem);
    base_step_kernel<true><<<a.B, warps * 32, smem, stream>>>(a, phases, collide ? 1 : 0, n_steps, blk_group);
  } else {
    optin.ensure(base_step_kernel<false>, smem);
    base_step_kernel<false><<<a.B, warps * 32, smem, stream>>>(a, phases, collide ? 1 : 0, n_steps, blk_group);
  }
  return true;
}

void launch_base_agents(const BaseKernelArgs& a, cudaStream_t stream) {
  const int smem_max = base_smem_optin();
  const int warps = base_agents_warps(a.N, a.W, (size_t)smem_max);
  const size_t smem = base_agents_smem_bytes(a.N, a.W, warps);
  static SmemOptIn optin;   // per device (abm_common.cuh)
  optin.ensure(base_agent_kernel, smem);
  const long long total = (long long)a.B * a.N;
  const unsigned grid = (unsigned)((total + warps - 1) / warps);
  base_agent_kernel<<<grid, warps * 32, smem, stream>>>(a);
}

// ---------------------------------------------------------------------------------------
// summary metrics of the foraging runs, one warp per replicate:
//   search efficiency  mean over agents of collected_r / T        (data_loader.py calculate_search_efficiency :1294-1353)
//   relocation time    fraction of the agent-steps logged in mode "relocate" (calculate_relocation_time :1903-1928);
//                      the other three modes come with it
// ---------------------------------------------------------------------------------------
__global__ void base_metrics_kernel(const float* __restrict__ collected, const unsigned int* __restrict__ mode_steps, int B,
                                    int N, unsigned long long steps, float* __restrict__ out) {
  const int b = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  double sum = 0.0;
  for (int i = lane; i < N; i += 32) sum += (double)collected[(size_t)b * N + i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if (lane == 0) {
    const double T = steps ? (double)steps : 1.0, NT = T * (double)N;
    float* o = out + (size_t)b * 6;
    o[0] = (float)(sum / (double)N / T);
    o[1] = (float)((double)mode_steps[b * 4 + MODE_RELOCATE] / NT);
    o[2] = (float)((double)mode_steps[b * 4 + MODE_EXPLORE] / NT);
    o[3] = (float)((double)mode_steps[b * 4 + MODE_EXPLOIT] / NT);
    o[4] = (float)((double)mode_steps[b * 4 + MODE_COLLIDE] / NT);
    o[5] = (float)(sum / (double)N);
  }
}
void launch_base_metrics(const float* collected, const unsigned int* mode_steps, int B, int N, unsigned long long steps,
                         float* out, cudaStream_t stream) {
  base_metrics_kernel<<<(B + 3) / 4, 128, 0, stream>>>(collected, mode_steps, B, N, steps, out);
}

}  // namespace abm
