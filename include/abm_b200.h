/*
 * abm_b200.h -- C ABI of libabm_b200.so: the B200-native engine for the one
 * data-parallel hot path of scioip34/ABM (per-agent visual-field projection +
 * vision-driven movement / flocking / foraging update, batched over agents and
 * independent replicate simulations).
 *
 * The reference is pure Python and has NO FFI boundary; the
 * entry points below are what a binding for this path would call, each citing
 * the reference interface it replaces (paths relative to the reference root).
 * The Python host (abm_b200/*.py) loads this library with ctypes; see
 * INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *   - every function returns 0 on success or a negative ABM_E_* code and never
 *     throws; abm_last_error() returns a thread-local message for the last failure.
 *   - the engine owns its device state; the caller owns every buffer it passes.
 *   - `on_device` == 1: the pointers are device pointers (same device as the
 *     engine); == 0: host pointers, calls that fill host buffers return when they are
 *     filled; == 2 (ABM_HOST_PINNED_ASYNC, abm_set_state / abm_get_state): PINNED host
 *     pointers, the call only enqueues on `stream` and never blocks -- the caller
 *     keeps the buffers alive and synchronises before touching them (two engines on
 *     two streams overlap one batch's copies with the other's step).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     all calls are asynchronous on that stream except create / destroy and calls
 *     that fill HOST buffers from pageable memory.
 *   - arrays over agents are SoA, length n_replicates * n_agents, replicate-major.
 *   - screen coordinates (y down); `x`,`y` are the TOP-LEFT corner of the agent
 *     sprite exactly like `Agent.position` (agent.py:54); centre = position + radius.
 *   - packed visual fields: bin b of the STORED (flipped) field `soc_v_field`
 *     (vf_supcalc.py:134, agent.py:593) is bit (b & 31) of word (b >> 5);
 *     words per field = abm_field_words(R) = ceil(R / 32).
 */
#ifndef ABM_B200_H
#define ABM_B200_H
#define ABM_HOST_PINNED_ASYNC 2

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABM_B200_VERSION 100 /* 0.1.0 */

enum {
  ABM_OK = 0,
  ABM_E_INVALID = -1,   /* bad argument / config */
  ABM_E_CUDA = -2,      /* CUDA runtime error (message has the detail) */
  ABM_E_NO_DEVICE = -3, /* no usable sm_100 device: there is NO CPU fallback */
  ABM_E_STATE = -4      /* call made in the wrong state (e.g. step before set_state) */
};

enum { ABM_BOUNDARY_WALLS = 0, ABM_BOUNDARY_INFINITE = 1 }; /* vf_params.py:20 BOUNDARY */

/* abm_vf_config_t.flags */
enum {
  ABM_VF_EXACT_FIXUP = 1u << 0, /* re-evaluate pairs whose fp32 bin index is within the error
                                   bound of a rounding boundary in fp64 (default on) */
  ABM_VF_KEEP_FIELDS = 1u << 1, /* keep the packed stored fields of the last step (abm_get_fields) */
  ABM_VF_KEEP_TERMS = 1u << 2,  /* keep the six flocking terms of the last step (abm_vf_get_terms) */
  ABM_VF_SPATIAL_SORT = 1u << 3 /* keep the agents of every replicate in Morton order internally (refreshed every
                                   `resort_every` steps); invisible through this ABI, which always speaks the
                                   caller's agent order -- except abm_vf_record_table, see there */
};

typedef struct abm_engine abm_engine_t;

/* Scalar configuration of a visual-flocking batch.
 * Replaces: VFSimulation.__init__ kwargs (vf_sims.py:16-44 -> sims.py:60-68) and the
 * module constants of vf_contrib/vf_params.py:12-23 that are per-run scalars. */
typedef struct {
  int32_t struct_size;    /* = sizeof(abm_vf_config_t) */
  int32_t n_replicates;   /* B independent simulations (metarunner.py:251-254 runs them one by one) */
  int32_t n_agents;       /* N agents per replicate (sims.py N) */
  int32_t resolution;     /* R, AFTER the int(R / fov) rescale of vf_sims.py:41-44 */
  int32_t fov_px0;        /* find_nearest(linspace(-pi,pi,R), fov[0]) (vf_supcalc.py:105) */
  int32_t fov_px1;        /* find_nearest(linspace(-pi,pi,R), fov[1]) */
  int32_t boundary;       /* ABM_BOUNDARY_* */
  int32_t limit_movement; /* VF_LIMIT_MOVEMENT (vf_agent.py:293-300) */
  float width;            /* ENV_WIDTH */
  float height;           /* ENV_HEIGHT */
  float window_pad;       /* hard-coded 30 in app_visual_flocking.py */
  float max_vel;          /* VF_MAX_VEL */
  float max_th;           /* VF_MAX_TH */
  uint32_t flags;         /* ABM_VF_* */
  /* large-swarm tiling: this engine updates agents
   * [tile_begin, tile_begin + tile_count) of every replicate and reads all n_agents
   * neighbour records; tile_count == 0 means the whole replicate. */
  int32_t tile_begin;
  int32_t tile_count;
  int32_t resort_every;   /* with ABM_VF_SPATIAL_SORT: re-sort after this many steps (0: only when the state is set
                             or abm_vf_resort is called) */
  /* Cyclic agent tiles (load balance of one large swarm across G GPUs: a contiguous range of the spatially sorted
   * order is a REGION of the arena, and regions differ in density by an order of magnitude on the reference's disc
   * initial condition): with tile_cycle = G > 1 this engine owns the blocks of ABM_VF_TILE_BLOCK consecutive internal
   * slots whose block index is congruent to tile_phase modulo G; tile_begin is ignored, tile_count must be
   * n_agents / G and n_agents a multiple of G * ABM_VF_TILE_BLOCK.  Focal agent li of the tile is slot
   * ((li / BLOCK) * G + tile_phase) * BLOCK + li % BLOCK.  0 / 1: the contiguous tile above. */
  int32_t tile_cycle;
  int32_t tile_phase;
} abm_vf_config_t;
#define ABM_VF_TILE_BLOCK 128

/* The six per-replicate flocking parameters, in this order (vf_params.py:12-19;
 * ALP2/BET2 multiply an all-zero dt_V at vf_supcalc.py:199 and do not exist here). */
enum { ABM_VF_GAM = 0, ABM_VF_V0, ABM_VF_ALP0, ABM_VF_ALP1, ABM_VF_BET0, ABM_VF_BET1, ABM_VF_NPARAM };

/* ---- library ---- */
int abm_version(void);
const char* abm_last_error(void);
int abm_device_count(void);
int abm_field_words(int resolution);

/* ---- engine life cycle ---- */
/* Replaces: VFSimulation(**kwargs) + prepare_start() (vf_sims.py:16, :342). */
int abm_vf_create(const abm_vf_config_t* cfg, int device, abm_engine_t** out);
int abm_destroy(abm_engine_t* e);

/* params: n_sets x ABM_VF_NPARAM doubles, n_sets == 1 (shared) or n_replicates (a sweep:
 * metarunner.py:168-214 writes one .env per combination).  Host pointer. */
int abm_vf_set_params(abm_engine_t* e, const double* params, int n_sets);

/* Per-agent overrides VFAgent.ALP0 / .BET0 / .V0 (vf_agent.py:24-26; None there == NaN here).
 * Any pointer may be NULL (= no override for that parameter). */
int abm_vf_set_agent_overrides(abm_engine_t* e, const float* alp0, const float* bet0, const float* v0,
                               int on_device, void* stream);

/* Agent state in / out.  Replaces writing / reading Agent.position, .orientation,
 * .velocity, .radius (agent.py:52-68).  set: radius may be NULL after the first call (the radii stay what they were:
 * they never change in the reference either).  get: any pointer may be NULL. */
int abm_set_state(abm_engine_t* e, const float* x, const float* y, const float* theta, const float* vel,
                  const float* radius, int on_device, void* stream);
int abm_get_state(abm_engine_t* e, float* x, float* y, float* theta, float* vel, int on_device, void* stream);

/* The same state as ONE interleaved array: (x, y, theta, vel) per agent, n_replicates * n_agents * 4 floats,
 * replicate-major -- one copy per direction instead of four (a host-driven loop that uploads the state, steps and
 * downloads it every step pays per-copy overhead and PCIe turn-arounds; with ABM_HOST_PINNED_ASYNC neither call blocks
 * or reads anything back).  radius: as in abm_set_state (SoA, NULL = keep).  Replaces the same attribute reads /
 * writes (agent.py:52-68) for a caller that keeps its agents as rows of a table. */
int abm_set_state_packed(abm_engine_t* e, const float* xytv, const float* radius, int on_device, void* stream);
int abm_get_state_packed(abm_engine_t* e, float* xytv, int on_device, void* stream);

/* n_steps synchronous (Jacobi) steps, one fused kernel launch per step.
 * Replaces: VFSimulation.step_sim -> agents.update (vf_sims.py:291-302) ->
 * VFAgent.update (vf_agent.py:52-80) for every agent of every replicate. */
int abm_vf_step(abm_engine_t* e, int n_steps, void* stream);

/* Line following: a VFAgent with lines to follow (`len(self.lines) != 0`, vf_agent.py:273-276) takes its
 * heading change from vf_supcalc.follow_lines_local (vf_supcalc.py:293-328: two sensors read window means of
 * `agent.line_map`) instead of the flocking integral -- the speed change stays the flocking one -- and turns back by pi
 * at the walls instead of pi/2 (vf_agent.py:80-129).  map: dim0 x dim1 float32, row-major, first axis x -- the
 * reference's `line_map` of shape (WIDTH + window_pad, HEIGHT + window_pad) (vf_agent.py:32), the same for every agent
 * and replicate; sensor_radius / sensor_distance: VFAgent.sensor_size = 9 / sensor_distance = 20 (vf_agent.py:30-31).
 * map == NULL: no lines again.  The six terms of abm_vf_get_terms stay the flocking ones. */
int abm_vf_set_line_map(abm_engine_t* e, const float* map, int dim0, int dim1, double sensor_radius, double sensor_distance,
                        int on_device, void* stream);

/* The host-driven loop of the reference in ONE call: upload the state (xytv_in, as abm_set_state_packed with radius ==
 * NULL: an earlier call passed the radii), n_steps steps, download the new state (xytv_out, as abm_get_state_packed).
 * Both arrays must be PINNED host memory (ABM_HOST_PINNED_ASYNC semantics: the call does not block; xytv_in may be
 * reused and xytv_out read after abm_synchronize(e, stream), or after any later work on `stream`).  When the batch runs
 * the symmetric kernel (a CTA per replicate) and is at least three waves of CTAs, the replicates go through in chunks of
 * whole waves: the upload of chunk c + 1 and the download of chunk c - 1 overlap the steps of chunk c on two copy
 * streams of the engine, so that one batch in flight already hides the PCIe time (only the first chunk's upload and the
 * last chunk's download are exposed).  Results are identical to the three separate calls.
 * Replaces: the per-step read of agent.position / orientation / velocity around VFSimulation.step_sim
 * (vf_sims.py:291-302) for a caller that keeps its agents on the host. */
int abm_vf_step_host(abm_engine_t* e, const float* xytv_in, float* xytv_out, int n_steps, void* stream);

/* Packed STORED fields of the last step, n_replicates*tile*abm_field_words(R) words
 * (needs ABM_VF_KEEP_FIELDS).  Replaces reading Agent.soc_v_field (vf_agent.py:209).
 * Row order: the caller's agent order for an engine that owns whole replicates.  A TILED engine (tile_count != 0)
 * with ABM_VF_SPATIAL_SORT owns the internal slots [tile_begin, tile_begin + tile_count) (or the cyclic blocks of
 * tile_cycle): its rows (and those of abm_vf_get_terms) are in the order of its focal agents li, row li = the agent
 * abm_vf_get_permutation reports for that focal agent's slot. */
int abm_get_fields(abm_engine_t* e, uint32_t* packed, int on_device, void* stream);

/* (dvel, dpsi, a_blob, a_edge, b_blob, b_edge) per agent of the last step, 6 doubles each
 * (needs ABM_VF_KEEP_TERMS).  Replaces VFAgent.dv/.dphi/.ablob/.aedge/.bblob/.bedge
 * (vf_agent.py:278-281). */
int abm_vf_get_terms(abm_engine_t* e, double* terms, int on_device, void* stream);

/* counters[0] = pairs re-evaluated in fp64 since creation, [1] = of those, handled inline
 * because the per-CTA queue was full, [2] = pairs whose fp32 and fp64 bin indices differed,
 * [3] = kernel launches since creation. */
int abm_get_counters(abm_engine_t* e, uint64_t counters[4], void* stream);

/* Device pointer of the neighbour-record table for the NEXT step: n_replicates*n_agents
 * float4 (x, y, radius, cull^2).  A tiled engine writes only its own tile; the host
 * completes the table with an all-gather (torch.distributed / NCCL) before the next
 * abm_vf_step.  *bytes_per_agent = 16. */
int abm_vf_record_table(abm_engine_t* e, void** dev_ptr, int* bytes_per_agent);

/* With ABM_VF_SPATIAL_SORT the record table, and the tile of a tiled engine, are in INTERNAL order.
 * abm_vf_get_permutation: perm[b * n_agents + slot] = caller's index of the agent in that slot.
 * abm_vf_resort: re-sort now (a tiled engine needs the full heading / speed arrays to be current on
 * every rank first: abm_vf_internal_arrays exposes them for the all-gather). */
int abm_vf_get_permutation(abm_engine_t* e, int32_t* perm, int on_device, void* stream);

/* Fused tile exchange of one large swarm across the GPUs of a node, instead of an all-gather per step: the
 * step kernel of a tiled engine stores every new record of its tile straight into the peers' record tables over NVLink
 * (CUDA IPC mappings), a launch starts only when every rank has published the previous step (flags in each GPU's
 * memory, written by the peers' last CTAs) and publishes its own at its end.  One process per GPU:
 *   abm_vf_ipc_export  -> ABM_VF_IPC_BYTES bytes describing this engine's tables and flags; exchange them between the
 *                         ranks (e.g. torch.distributed.all_gather);
 *   abm_vf_ipc_attach  <- the exports of ALL ranks in rank order (this engine's own entry at `my_rank` is skipped).
 * All ranks must then call abm_vf_step the same number of times; abm_set_state on an attached engine needs a barrier
 * between the ranks before the next step (the host's job). */
#define ABM_VF_IPC_BYTES 512
int abm_vf_ipc_export(abm_engine_t* e, void* out);
int abm_vf_ipc_attach(abm_engine_t* e, int n_ranks, int my_rank, const void* exports);
int abm_vf_resort(abm_engine_t* e, void* stream);
int abm_vf_internal_arrays(abm_engine_t* e, void** theta_dev, void** vel_dev);

/* Name of the step kernel the last abm_vf_step launched ("abm::vf_step_sym_kernel": every unordered pair once, all rows
 * of a replicate in one CTA's shared memory; "abm::vf_step_kernel": one thread per focal agent; "abm::vf_step_warp_kernel":
 * a CTA per few focal agents, a thread per neighbour record -- large sparse swarms, their tiles, small batches), ""
 * before the first step.  The choice is made per step from the state; the environment variable
 * ABM_VF_KERNEL=onesided|symmetric|warp overrides it for tests. */
const char* abm_vf_last_kernel(abm_engine_t* e);
/* Step-kernel launches so far: [0] symmetric with the two-word fast path, [1] symmetric with the three-word fast path
 * (crowded scenes: chosen when more than 4 % of the pairs of a two-word step left its fast path), [2] one thread per
 * focal agent, [3] warp kernel (a multi-step cooperative launch counts once). */
int abm_vf_kernel_stats(abm_engine_t* e, uint64_t stats[4]);
/* Multi-step launches (abm_vf_step with n_steps > 1 on one small replicate) whose grid was ONE thread-block cluster: the
 * steps of the launch are separated by the hardware cluster barrier instead of a grid-wide barrier in global memory
 * (at most 16 CTAs x 4 focal agents = 64 agents; ABM_VF_NO_CLUSTER=1 disables it).  Diagnostic, like the stats above. */
int abm_vf_cluster_launches(abm_engine_t* e, uint64_t* out);
/* Statistics behind the automatic choice: unordered pairs that left the symmetric kernel's (two-word) fast path -- wide
 * intervals, guard-band hits -- summed over its launches so far, and the number of those launches. */
int abm_vf_slow_entries(abm_engine_t* e, uint64_t* entries, uint64_t* sym_launches, void* stream);

/* Summary metrics of the current state, per replicate (the quantities abm/loader/data_loader.py computes
 * offline from logged trajectories: calculate_polarization :1761-1836, calculate_interindividual_distance :1367-1460,
 * calculate_mean_NN_dist :1461-1488, calculate_collision_time :1838-1869).  out[b * 5 + k], k = 0: polarization
 * |sum_i (cos theta_i, sin theta_i)| / N; 1: mean distance over pairs i < j (minimal image with BOUNDARY infinite);
 * 2: mean over agents of the distance to the nearest other agent; 3: 1.0 if any pair is closer than 2 * radius (the
 * reference's agent-agent collision criterion), else 0.0; 4: the fraction of agents i that have an agent j > i (caller's
 * order) at 0 < distance < 2 * radius -- the loader's own per-agent indicator (it works on the upper triangle of the
 * distance matrix); its mean over time is an experiment's "aacoll" value. */
int abm_vf_metrics(abm_engine_t* e, float* out, int on_device, void* stream);

int abm_synchronize(abm_engine_t* e, void* stream);

/* ---- stateless function-level entry points (host pointers, synchronous) ----
 * They run on the CURRENT CUDA device of the calling thread (cudaSetDevice is the caller's business, as for any CUDA
 * library call without a handle); that device must be an sm_100 one. */

/* vf_supcalc.projection_field (vf_supcalc.py:20-138): one focal agent, n_obj objects.
 * Inputs are float64 like the reference's; they are rounded to the engine's fp32 state.
 * out_rows: n_obj * abm_field_words(R) words, STORED (flipped) order, one row per object. */
typedef struct {
  int32_t struct_size;
  int32_t resolution;
  double fov0, fov1;
  double x, y, radius, orientation; /* focal agent */
  int32_t n_obj;
  const double* obj_x;
  const double* obj_y;
  const double* obj_size; /* NULL: every object has the focal radius (vf_supcalc.py:61-62) */
  int32_t boundary;
  double arena_width, arena_height;
  double vision_range; /* < 0: None */
} abm_vf_proj_args_t;
int abm_vf_projection_field(const abm_vf_proj_args_t* args, uint32_t* out_rows);

/* cooperative-signaling projection_field (abm/projects/cooperative_signaling/cs_agent/cs_supcalc.py:204-289):
 * one focal agent, n_obj objects of the FOCAL radius; visibility decided on the angle (fov0 <= angle <= fov1),
 * projections wider than max_proj_size bins dropped, bins outside the FOV cleared after the flip.  Evaluated in
 * float64 on the float64 inputs.  out_rows as above (bits); the meter amplitudes of the reference
 * (object_meters, :283-284) scale whole rows and are applied by the caller. */
typedef struct {
  int32_t struct_size;
  int32_t resolution;
  double fov0, fov1;
  double x, y, radius, orientation; /* focal agent */
  int32_t n_obj;
  const double* obj_x;
  const double* obj_y;
  double max_proj_size; /* < 0: None */
} abm_cs_proj_args_t;
int abm_cs_projection_field(const abm_cs_proj_args_t* args, uint32_t* out_rows);

/* VSWRM_flocking_state_variables (vf_supcalc.py:161-254), verbose form, on a packed
 * UN-flipped field V_now (what vf_agent.py:270 passes: np.flip(soc_v_field)).
 * params: ABM_VF_NPARAM doubles.  out: dvel, dpsi, a_blob, a_edge, b_blob, b_edge. */
int abm_vf_flocking_terms(const uint32_t* packed_v_now, int resolution, double vel_now,
                          const double* params, double out[6]);


/* =====================================================================================
 * BASE / collective-foraging variant (Simulation + Agent of abm/simulation/sims.py and
 * abm/agent/agent.py): social visual field with distance-ordered occlusion, decision
 * process (w, u), mode machine, kinematics, walls, and agent-patch exploitation.
 * ===================================================================================== */

typedef struct abm_base_engine abm_base_engine_t;

/* Replaces: Simulation.__init__ kwargs (sims.py:60-68) that the hot path reads. */
typedef struct {
  int32_t struct_size;          /* = sizeof(abm_base_config_t) */
  int32_t n_replicates;         /* B independent simulations */
  int32_t n_agents;             /* N (sims.py N) */
  int32_t n_patches;            /* N_resc; 0 = no environment */
  int32_t resolution;           /* v_field_res */
  int32_t tau;                  /* DEC_TAU: length of the novelty memory, <= 32 (decision_params.py:39) */
  int32_t visual_exclusion;     /* VISUAL_EXCLUSION (agent.py:415, :559) */
  int32_t patchwise_exclusion;  /* PATCHWISE_SOCIAL_EXCLUSION (agent.py:406-410) */
  int32_t teleport_exploit;     /* TELEPORT_TO_MIDDLE (sims.py:820-821) */
  int32_t regenerate_patches;   /* REGENERATE_PATCHES (sims.py:321-330) */
  int32_t patch_border_overlap; /* PATCH_BORDER_OVERLAP (sims.py:351-358) */
  int32_t keep_fields;          /* keep packed stored fields of the last step (abm_base_get_fields) */
  int32_t collide_agents;       /* AGENT_AGENT_COLLISION (sims.py:736-783); pair detection restates pygame's
                                   documented collide_circle -- parity unpinned, see DESIGN.md */
  int32_t ghost_mode;           /* GHOST_WHILE_EXPLOIT (sims.py:431-436, 771-776) */
  double fov0, fov1;            /* agent_fov in radians: (-fov*pi, fov*pi) (sims.py:160-161) */
  double width, height, window_pad;
  double vision_range;          /* VISION_RANGE (agent.py:400) */
  double agent_radius;          /* RADIUS_AGENT (one radius per batch, as in the reference's homogeneous runs) */
  double patch_radius;          /* RADIUS_RESOURCE, used when a depleted patch is re-created */
  double min_quality, max_quality;   /* MIN/MAX_RESOURCE_QUALITY (negatives resolved by the host, sims.py:176-179) */
  int32_t min_units, max_units;      /* MIN/MAX_RESOURCE_PER_PATCH */
  uint64_t seed;                /* counter-based RNG key (random walk, patch regeneration) */
} abm_base_config_t;

/* One parameter set (per batch, replicate or agent: abm_base_set_params), in this order (decision_params.py:13-42, movement_params.py:13-23,
 * sims.py agent_consumption). */
enum {
  ABM_BASE_T_W = 0, ABM_BASE_EPS_W, ABM_BASE_G_W, ABM_BASE_B_W, ABM_BASE_W_MAX,
  ABM_BASE_T_U, ABM_BASE_EPS_U, ABM_BASE_G_U, ABM_BASE_B_U, ABM_BASE_U_MAX,
  ABM_BASE_S_WU, ABM_BASE_S_UW, ABM_BASE_F_N, ABM_BASE_F_R,
  ABM_BASE_EXP_VEL_MAX, ABM_BASE_EXP_TH_MIN, ABM_BASE_EXP_TH_MAX, ABM_BASE_REL_TH_MAX, ABM_BASE_STOP_RATIO,
  ABM_BASE_CONSUMPTION, ABM_BASE_NPARAM
};

/* override_mode: Agent.overriding_mode (agent.py:671-693): 0 None, 1 "exploit", 3 "collide".
 * mode: last Agent.mode, logged code of ifdb.py:197-206: 0 explore, 1 exploit, 2 relocate, 3 collide.
 * novelty: bit t = Agent.novelty[t] (sims.py:33-37).  SoA, n_replicates * n_agents each;
 * any member may be NULL (skipped). */
typedef struct {
  float *x, *y, *theta, *vel, *w, *u, *collected, *collected_before, *i_priv;
  int32_t *env_status, *override_mode, *mode, *patch_id;
  uint32_t* novelty;
} abm_base_agents_t;

/* Rescource fields (rescource.py:15-47): top-left position, radius, resc_left, unit_per_timestep, id.
 * n_replicates * n_patches each. */
typedef struct {
  float *x, *y, *radius, *left, *quality;
  int32_t* id;
} abm_base_patches_t;

enum { ABM_BASE_PHASE_ENV = 1, ABM_BASE_PHASE_AGENTS = 2, ABM_BASE_PHASE_COLLISIONS = 4, ABM_BASE_PHASE_ALL = 7 };

/* Replaces: Simulation(**kwargs) + create_agents / create_resources (sims.py:526-541). */
int abm_base_create(const abm_base_config_t* cfg, int device, abm_base_engine_t** out);
int abm_base_destroy(abm_base_engine_t* e);
/* params: n_sets * ABM_BASE_NPARAM doubles.  n_sets = 1 (one set for the batch), n_replicates (one per replicate:
 * parameter sweeps) or n_replicates * n_agents (one per agent, replicate-major: heterogeneous agents, the
 * behave_params of agent.py:83-108 / agent_behave_param_list of sims.py:499-517 -- decision parameters,
 * exp_vel_max, exp_stop_ratio, agent_consumption; FOV and vision range: abm_base_set_agent_geometry). */
int abm_base_set_params(abm_base_engine_t* e, const double* params, int n_sets);
/* Per-agent field geometry of heterogeneous agents (sims.py:499-517: every Agent gets its own FOV and vision_range):
 * fov0 / fov1 in radians (the reference's (-agent_fov * pi, agent_fov * pi)) and vision_range, n = n_replicates *
 * n_agents values each, replicate-major, host pointers.  n = 0 returns to the engine-wide values of the config.
 * The radius: abm_base_set_agent_radii; the resolution: abm_base_set_agent_resolution. */
int abm_base_set_agent_geometry(abm_base_engine_t* e, const double* fov0, const double* fov1,
                                const double* vision_range, int n);
/* Per-agent field resolution of heterogeneous agents (sims.py:507: Agent(v_field_res = behave_params["v_field_res"]);
 * agent.py:58, 480-481, 543, 577-588: the agent's own linspace grid, projection size and wrap; supcalc.py:86-91 and
 * agent.py:196-199: the means over ITS field length; sims.py:449-462: the collision LIDAR field of the hit agent):
 * n = n_replicates * n_agents values in [2, cfg.resolution], replicate-major, host pointer; n = 0 returns to
 * cfg.resolution for everybody.  cfg.resolution stays the row stride of abm_base_fields: agent i's row holds its
 * resolution[i] stored bins, the bits beyond are 0. */
int abm_base_set_agent_resolution(abm_base_engine_t* e, const int32_t* resolution, int n);
/* Per-agent radius of heterogeneous agents (sims.py:502: Agent(radius = behave_params["agent_radius"])): n =
 * n_replicates * n_agents values, replicate-major, host pointer; n = 0 returns to the config's agent_radius.  As in the
 * reference the candidate test (agent.py:400), the patch membership (sims.py:45-56), the wall reflection
 * (agent.py:347-394) and the collision circles use each agent's OWN radius, the projection the FOCAL agent's radius for
 * both centres and for the projected size (agent.py:504-509, :529). */
int abm_base_set_agent_radii(abm_base_engine_t* e, const double* radius, int n);
int abm_base_set_agents(abm_base_engine_t* e, const abm_base_agents_t* src, int on_device, void* stream);
int abm_base_get_agents(abm_base_engine_t* e, const abm_base_agents_t* dst, int on_device, void* stream);
int abm_base_set_patches(abm_base_engine_t* e, const abm_base_patches_t* src, int on_device, void* stream);
int abm_base_get_patches(abm_base_engine_t* e, const abm_base_patches_t* dst, int on_device, void* stream);

/* n_steps time steps of the main loop body (sims.py:733-864): environment phase
 * (agent-patch interaction, notify; sims.py:790-858) then agent phase (Agent.update for all
 * agents from one frozen snapshot; sims.py:861 -> agent.py:212-283).
 * inject_dtheta: NULL, or n_replicates*n_agents floats that replace the random-walk draw
 * np.random.uniform(exp_theta_min, exp_theta_max) (supcalc.py:45) -- used with n_steps == 1
 * by the parity tests.  phases: ABM_BASE_PHASE_*. */
int abm_base_step(abm_base_engine_t* e, int n_steps, const float* inject_dtheta, int inject_on_device,
                  uint32_t phases, void* stream);

/* Patch regeneration with GIVEN draws (parity tests; the counterpart of inject_dtheta): Simulation.add_new_resource_patch
 * (sims.py:332-374) draws x, y (np.random.randint), units (randint) and quality (uniform) for every try and retries
 * while the new patch overlaps another one.  draws: n_replicates * n_patches * n_tries * 4 doubles -- (x, y, units,
 * quality) of try t of a regeneration of patch slot p of replicate b, host pointer; a regeneration that exhausts its
 * n_tries counts as failed (abm_base_get_counters [1]) and leaves the slot dead.  Applies to all following steps;
 * draws == NULL or n_tries <= 0 returns to the engine's counter-based RNG. */
int abm_base_inject_regeneration(abm_base_engine_t* e, const double* draws, int n_tries);
/* One set of patch-regeneration parameters per replicate (a MetaProtocol sweep over RADIUS_RESOURCE,
 * MIN/MAX_RESOURCE_PER_PATCH, MIN/MAX_RESOURCE_QUALITY as ONE batch; add_new_resource_patch, sims.py:332-374, reads them
 * from the simulation object): table = n_replicates x 5 doubles (patch radius, min quality, max quality, min units,
 * max units), host pointer; n = 0 returns to the values of the config. */
int abm_base_set_regeneration_params(abm_base_engine_t* e, const double* table, int n);

/* Packed STORED fields (flipped + FOV-masked: Agent.soc_v_field, agent.py:593-597) of the last step. */
int abm_base_get_fields(abm_base_engine_t* e, uint32_t* packed, int on_device, void* stream);

/* counters[0] = patches regenerated, [1] = regenerations that exhausted their retries,
 * [2] = kernel launches, [3] = steps. */
int abm_base_get_counters(abm_base_engine_t* e, uint64_t counters[4], void* stream);

/* Summary metrics of the foraging runs, per replicate, reduced on the device (the quantities
 * abm/loader/data_loader.py computes offline from the logged collresource / mode arrays): out[b * 6 + k], k = 0: search
 * efficiency = mean over agents of collected_r / T (calculate_search_efficiency :1294-1353 with t_start = 0); 1: relative
 * relocation time = fraction of the agent-steps logged in mode "relocate" (calculate_relocation_time :1903-1928; the mode
 * is the one ifdb.save_agent_data_RAM logs at the end of a step, codes of ifdb.mode_to_int :197-206); 2, 3, 4: the same
 * for explore, exploit, collide; 5: mean collected_r.  T = steps with an agent phase since creation or the last reset.
 * out == NULL with reset != 0 only restarts the time window (collected_r keeps counting, as in the reference). */
int abm_base_metrics(abm_base_engine_t* e, float* out, int on_device, int reset, void* stream);

/* ---- stateless function-level entry points of the BASE variant (host pointers, synchronous) ---- */

/* Agent.projection_field(obstacles, keep_distance_info, non_expl_agents, fov) (agent.py:457-597)
 * of one focal agent.  Objects are [social..., occluders...] in list order; occluders are only
 * used when visual_exclusion != 0 (they clip, they are never drawn).  out_field: packed STORED
 * field (flipped + FOV-masked), abm_field_words(R) words.  out_amplitude: the value the set bins
 * carry: 1, or 1 - distance_last / vision_range with keep_distance_info (agent.py:590). */
typedef struct {
  int32_t struct_size;
  int32_t resolution;
  double fov0, fov1;
  double x, y, radius, orientation;
  int32_t n_social;
  int32_t n_occluders;
  const double* social_x;
  const double* social_y;
  const double* occluder_x;
  const double* occluder_y;
  int32_t visual_exclusion;
  int32_t keep_distance_info;
  double vision_range;
} abm_base_proj_args_t;
int abm_base_projection_field(const abm_base_proj_args_t* args, uint32_t* out_field, double* out_amplitude);

/* supcalc.F_reloc_LR(vel_now, V_now, v_desired) (supcalc.py:81-92) on a packed STORED field:
 * out[0] = v_desired - vel_now, out[1] = (mean(V[:R/2]) - mean(V[R/2:])) * amplitude * reloc_theta_max. */
int abm_base_reloc_lr(const uint32_t* packed_field, int resolution, double amplitude, double vel_now,
                      double v_desired, double reloc_theta_max, double out[2]);

/* vf_supcalc.dPhi_V_of(Phi, V) (vf_supcalc.py:257-277) on a packed field: out[k] in {-1, 0, 1}. */
int abm_vf_dphi(const uint32_t* packed_v, int resolution, int8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* ABM_B200_H */
