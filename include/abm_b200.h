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
 *   - `on_device` != 0: the pointers are device pointers (same device as the
 *     engine); == 0: host pointers (pinned for truly asynchronous copies).
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
  ABM_VF_KEEP_TERMS = 1u << 2   /* keep the six flocking terms of the last step (abm_vf_get_terms) */
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
} abm_vf_config_t;

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
 * .velocity, .radius (agent.py:52-68).  get: any pointer may be NULL. */
int abm_set_state(abm_engine_t* e, const float* x, const float* y, const float* theta, const float* vel,
                  const float* radius, int on_device, void* stream);
int abm_get_state(abm_engine_t* e, float* x, float* y, float* theta, float* vel, int on_device, void* stream);

/* n_steps synchronous (Jacobi) steps, one fused kernel launch per step.
 * Replaces: VFSimulation.step_sim -> agents.update (vf_sims.py:291-302) ->
 * VFAgent.update (vf_agent.py:52-80) for every agent of every replicate. */
int abm_vf_step(abm_engine_t* e, int n_steps, void* stream);

/* Packed STORED fields of the last step, n_replicates*tile*abm_field_words(R) words
 * (needs ABM_VF_KEEP_FIELDS).  Replaces reading Agent.soc_v_field (vf_agent.py:209). */
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

int abm_synchronize(abm_engine_t* e, void* stream);

/* ---- stateless function-level entry points (host pointers, synchronous) ---- */

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

/* VSWRM_flocking_state_variables (vf_supcalc.py:161-254), verbose form, on a packed
 * UN-flipped field V_now (what vf_agent.py:270 passes: np.flip(soc_v_field)).
 * params: ABM_VF_NPARAM doubles.  out: dvel, dpsi, a_blob, a_edge, b_blob, b_edge. */
int abm_vf_flocking_terms(const uint32_t* packed_v_now, int resolution, double vel_now,
                          const double* params, double out[6]);

#ifdef __cplusplus
}
#endif
#endif /* ABM_B200_H */
