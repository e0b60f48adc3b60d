// Declarations of the BASE (collective foraging) path.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "abm_common.cuh"

namespace abm {

// override codes (Agent.overriding_mode, agent.py:671-693) and logged mode codes (ifdb.py:197-206)
enum { OV_NONE = 0, OV_EXPLOIT = 1, OV_COLLIDE = 3 };
enum { MODE_EXPLORE = 0, MODE_EXPLOIT = 1, MODE_RELOCATE = 2, MODE_COLLIDE = 3 };

// per-replicate decision / movement parameters, order of ABM_BASE_* in include/abm_b200.h
struct BaseParams {
  double T_w, Eps_w, g_w, B_w, w_max;
  double T_u, Eps_u, g_u, B_u, u_max;
  double S_wu, S_uw, F_N, F_R;
  double exp_vel_max, exp_theta_min, exp_theta_max, reloc_theta_max, exp_stop_ratio;
  double consumption;
};
constexpr int kBaseNParam = 20;

struct BaseAgentPtrs {
  float *x, *y, *theta, *vel, *w, *u, *collected, *collected_before, *i_priv;
  int32_t *env_status, *override_mode, *mode, *patch_id;
  uint32_t* novelty;   // Tau-bit shift register: bit t = novelty[t] (sims.py:33-37)
  // frozen snapshot read by the agent phase (Jacobi update): written at the end of the environment phase
  float *snap_x, *snap_y;
  int32_t* snap_override;
  int32_t* collided;   // 1 = member of `collided_agents` of this step (sims.py:754-783); all 0 without collisions
  const float* radius; // nullable, B*N: per-agent radius (heterogeneous agents, sims.py:502); else BaseKernelArgs::radius
};
struct BasePatchPtrs {
  float *x, *y, *radius, *left, *quality;
  int32_t* id;
};

// Per-agent field geometry (heterogeneous agents, sims.py:499-517: FOV and vision_range are constructor arguments of
// every agent): FOV in radians, vision range, the stored bins the FOV mask keeps (agent.py:594-595), and the agent's own
// field resolution (v_field_res, agent.py:58; <= BaseKernelArgs::R, which is the row stride) with its linspace step.
struct BaseAgentGeo {
  double fov0, fov1, vision_range;
  double lin_step;
  int mask_lo, mask_hi;
  int res, pad_;
};

struct BaseKernelArgs {
  int B, N, P, R, W, Tau;
  int visual_exclusion, patchwise_exclusion, teleport_exploit, regenerate, border_overlap, ghost_mode;
  double fov0, fov1;           // agent FOV in radians (strict test on the closed angle, agent.py:535)
  int mask_lo, mask_hi;        // stored bins kept by the FOV mask: phis[b] >= fov0 && phis[b] <= fov1 (agent.py:594-595)
  double lin_step;             // numpy linspace step (see nearest_bin_exact)
  double width, height, pad, vision_range, radius;
  // patch regeneration (sims.py:332-374)
  double patch_radius, min_quality, max_quality;
  int min_units, max_units;
  // nullable, B x 5 doubles (patch radius, min / max quality, min / max units): one set per replicate -- a sweep over the
  // patch parameters as one batch (abm_base_set_regeneration_params)
  const double* regen_tab;
  unsigned long long seed;
  unsigned step;               // time step index, part of the RNG counter
  BaseAgentPtrs ag;            // B*N, updated in place
  BasePatchPtrs pa;            // B*P, updated in place
  const double* params;        // n_sets * kBaseNParam
  int param_stride;            // doubles between the sets of consecutive replicates: 0, kBaseNParam or N * kBaseNParam
  int param_stride_agent;      // doubles between the sets of consecutive agents: 0, or kBaseNParam with one set per
                               // agent (heterogeneous agents: agent.py:83-108 behave_params, sims.py:499-517)
  const BaseAgentGeo* agent_geo;   // nullable, B*N: replaces fov0 / fov1 / mask_lo / mask_hi / vision_range per focal agent
  const float* inject_dtheta;  // nullable, B*N: replaces the random-walk draw (parity tests)
  const double* regen_draws;   // nullable, B*P*regen_tries*4: (x, y, units, quality) of every try of a regeneration of
  int regen_tries;             // patch slot p of replicate b, replacing the four draws of sims.py:351-361 (parity tests)
  uint32_t* fields_out;        // nullable, B*N*W, stored order
  unsigned long long* counters;   // [0] patches regenerated, [1] regeneration retries exhausted
  unsigned int* mode_steps;    // B*4: agent-steps spent in mode explore / exploit / relocate / collide since the last
                               // reset (the mode an agent is logged with, ifdb.py:197-206); summary metrics
};

bool launch_base_step(const BaseKernelArgs& a, unsigned phases, bool collide, int n_steps, int n_sms, cudaStream_t stream);
void launch_base_env(const BaseKernelArgs& a, cudaStream_t stream);
void launch_base_agents(const BaseKernelArgs& a, cudaStream_t stream);
void launch_base_collisions(const BaseKernelArgs& a, cudaStream_t stream);
void launch_base_metrics(const float* collected, const unsigned int* mode_steps, int B, int N, unsigned long long steps,
                         float* out, cudaStream_t stream);

struct BaseProjArgs {
  int R, W, n_social, n_occ, visual_exclusion, keep_distance;
  int mask_lo, mask_hi;
  double fov0, fov1, lin_step, radius, vision_range;
  float fx, fy, ftheta;
  const float* ox; const float* oy;   // n_social + n_occ positions
  uint32_t* field; double* amplitude;
};
void launch_base_projection(const BaseProjArgs& a, cudaStream_t stream);
void launch_base_reloc_lr(const uint32_t* field, int R, int W, double amp, double vel, double vdes, double thmax,
                          double* out2, cudaStream_t stream);
void launch_vf_dphi(const uint32_t* v, int R, int W, signed char* out, cudaStream_t stream);
size_t base_agents_smem_bytes(int N, int W, int warps);
int base_agents_warps(int N, int W, size_t smem_limit);

}  // namespace abm
