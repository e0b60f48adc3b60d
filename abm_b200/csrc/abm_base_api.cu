// C ABI of the BASE / collective-foraging engine (declared in include/abm_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <utility>
#include <new>
#include <string>
#include <vector>

#include "abm_api_util.cuh"
#include "abm_base.cuh"

using abm::DevBuf;

struct abm_base_engine {
  abm_base_config_t cfg;
  int device = 0, n_sms = 148;
  int W = 0, mask_lo = 0, mask_hi = -1;
  double lin_step = 0.0;
  size_t n_agents_total = 0, n_patches_total = 0;
  DevBuf<float> x, y, theta, vel, w, u, collected, collected_before, i_priv, snap_x, snap_y;
  DevBuf<int32_t> env_status, override_mode, mode, patch_id, snap_override, collided;
  DevBuf<uint32_t> novelty, fields;
  DevBuf<float> px, py, pradius, pleft, pquality;
  DevBuf<int32_t> pid;
  DevBuf<double> params;
  DevBuf<abm::BaseAgentGeo> agent_geo;   // B*N once abm_base_set_agent_geometry / _resolution was called
  bool has_agent_geo = false;
  std::vector<double> h_fov0, h_fov1, h_vr;   // host copies of what the two calls set (either may be empty)
  std::vector<int> h_res;
  DevBuf<double> regen_tab;              // B x 5 once abm_base_set_regeneration_params was called
  bool has_regen_tab = false;
  DevBuf<double> regen_draws;            // abm_base_inject_regeneration
  int regen_tries = 0;
  DevBuf<float> agent_radius;            // B*N once abm_base_set_agent_radii was called
  bool has_agent_radius = false;
  DevBuf<float> inject;
  DevBuf<unsigned long long> counters;
  DevBuf<unsigned int> mode_steps;      // B*4, see BaseKernelArgs
  DevBuf<float> metrics;                // B*6 staging of abm_base_metrics
  unsigned long long metric_steps = 0;  // full steps since the last metrics reset
  int n_param_sets = 1;
  bool agents_set = false, patches_set = false;
  unsigned step = 0;
  unsigned long long launches = 0, steps = 0;
};

namespace {

inline int fail(int code, const std::string& msg) { return abm::api_fail(code, msg); }

template <typename T>
int xfer(T* dev, const T* host_or_dev, T* out, size_t n, bool to_device, int on_device, cudaStream_t st) {
  // to_device: copy src(host_or_dev) -> dev ; else dev -> out
  if (to_device) {
    if (!host_or_dev) return ABM_OK;
    ABM_CUDA(cudaMemcpyAsync(dev, host_or_dev, sizeof(T) * n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  } else {
    if (!out) return ABM_OK;
    ABM_CUDA(cudaMemcpyAsync(out, dev, sizeof(T) * n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  }
  return ABM_OK;
}

}  // namespace

extern "C" {

int abm_base_create(const abm_base_config_t* cfg, int device, abm_base_engine_t** out) {
  if (!cfg || !out) return fail(ABM_E_INVALID, "abm_base_create: null argument");
  if (cfg->struct_size != (int32_t)sizeof(abm_base_config_t))
    return fail(ABM_E_INVALID, "abm_base_create: struct_size mismatch (header / library version skew)");
  if (cfg->n_replicates < 1 || cfg->n_agents < 1 || cfg->n_patches < 0)
    return fail(ABM_E_INVALID, "abm_base_create: bad batch shape");
  if (cfg->resolution < 8 || cfg->resolution > 65535) return fail(ABM_E_INVALID, "abm_base_create: bad resolution");
  if (cfg->tau < 1 || cfg->tau > 32) return fail(ABM_E_INVALID, "abm_base_create: tau must be in [1, 32]");
  if (!(cfg->agent_radius > 0.0)) return fail(ABM_E_INVALID, "abm_base_create: agent_radius must be > 0");
  int ndev = 0;
  ABM_CUDA(cudaGetDeviceCount(&ndev));
  if (ndev < 1 || device < 0 || device >= ndev) return fail(ABM_E_NO_DEVICE, "abm_base_create: no such CUDA device");
  ABM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ABM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(ABM_E_NO_DEVICE, "abm_base_create: device is not sm_100");
  abm_base_engine* e = new (std::nothrow) abm_base_engine();
  if (!e) return fail(ABM_E_INVALID, "abm_base_create: out of host memory");
  e->cfg = *cfg;
  e->device = device;
  e->n_sms = prop.multiProcessorCount;
  const int R = cfg->resolution;
  e->W = (R + 31) / 32;
  e->lin_step = ABM_TWO_PI_D / (double)(R - 1);
  // FOV mask in stored coordinates (agent.py:594-595) on the numpy linspace grid
  e->mask_lo = R; e->mask_hi = -1;
  for (int k = 0; k < R; ++k) {
    const double phi = (k == R - 1) ? ABM_PI_D : ((double)k * e->lin_step + (-ABM_PI_D));
    if (!(phi < cfg->fov0) && !(phi > cfg->fov1)) {
      if (k < e->mask_lo) e->mask_lo = k;
      if (k > e->mask_hi) e->mask_hi = k;
    }
  }
  if (abm::base_agents_smem_bytes(cfg->n_agents, e->W, 1) > (size_t)prop.sharedMemPerBlockOptin) {
    delete e;
    return fail(ABM_E_INVALID, "abm_base_create: n_agents too large for the per-warp occlusion buffers");
  }
  const size_t na = e->n_agents_total = (size_t)cfg->n_replicates * cfg->n_agents;
  const size_t np = e->n_patches_total = (size_t)cfg->n_replicates * cfg->n_patches;
  cudaError_t err = cudaSuccess;
  auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
  for (DevBuf<float>* b : {&e->x, &e->y, &e->theta, &e->vel, &e->w, &e->u, &e->collected, &e->collected_before,
                           &e->i_priv, &e->snap_x, &e->snap_y, &e->inject}) A(b->alloc(na));
  for (DevBuf<int32_t>* b : {&e->env_status, &e->override_mode, &e->mode, &e->patch_id, &e->snap_override, &e->collided})
    A(b->alloc(na));
  A(e->novelty.alloc(na));
  if (cfg->keep_fields) A(e->fields.alloc(na * e->W));
  for (DevBuf<float>* b : {&e->px, &e->py, &e->pradius, &e->pleft, &e->pquality}) A(b->alloc(np));
  A(e->pid.alloc(np));
  A(e->params.alloc((size_t)cfg->n_replicates * cfg->n_agents * abm::kBaseNParam));   // up to one set per agent
  A(e->counters.alloc(4));
  if (err == cudaSuccess) err = cudaMemset(e->counters.p, 0, 4 * sizeof(unsigned long long));
  A(e->mode_steps.alloc(4 * (size_t)cfg->n_replicates));
  A(e->metrics.alloc(6 * (size_t)cfg->n_replicates));
  if (err == cudaSuccess) err = cudaMemset(e->mode_steps.p, 0, 4 * sizeof(unsigned int) * (size_t)cfg->n_replicates);
  // defaults of decision_params.py / movement_params.py
  const double defaults[abm::kBaseNParam] = {0.5, 3, 0.085, 0, 1, 0.5, 3, 0.085, 0, 1, 0.25, 0.01, 2, 1,
                                             1, -0.3, 0.3, 0.5, 0.08, 1};
  if (err == cudaSuccess) err = cudaMemcpy(e->params.p, defaults, sizeof(defaults), cudaMemcpyHostToDevice);
  for (DevBuf<float>* b : {&e->w, &e->u, &e->collected, &e->collected_before, &e->i_priv, &e->vel})
    if (err == cudaSuccess) err = cudaMemset(b->p, 0, sizeof(float) * na);
  for (DevBuf<int32_t>* b : {&e->env_status, &e->override_mode, &e->mode, &e->collided})
    if (err == cudaSuccess) err = cudaMemset(b->p, 0, sizeof(int32_t) * na);
  if (err == cudaSuccess) err = cudaMemset(e->patch_id.p, 0xff, sizeof(int32_t) * na);
  if (err == cudaSuccess) err = cudaMemset(e->novelty.p, 0, sizeof(uint32_t) * na);
  if (err != cudaSuccess) {
    std::string m = std::string("abm_base_create: ") + cudaGetErrorString(err);
    abm_base_destroy(e);
    return fail(ABM_E_CUDA, m);
  }
  *out = e;
  return ABM_OK;
}

int abm_base_destroy(abm_base_engine_t* e) {
  if (!e) return ABM_OK;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (DevBuf<float>* b : {&e->x, &e->y, &e->theta, &e->vel, &e->w, &e->u, &e->collected, &e->collected_before,
                           &e->i_priv, &e->snap_x, &e->snap_y, &e->inject, &e->px, &e->py, &e->pradius, &e->pleft,
                           &e->pquality}) b->release();
  for (DevBuf<int32_t>* b : {&e->env_status, &e->override_mode, &e->mode, &e->patch_id, &e->snap_override, &e->pid,
                             &e->collided}) b->release();
  e->novelty.release(); e->fields.release(); e->params.release(); e->agent_geo.release(); e->regen_tab.release(); e->regen_draws.release(); e->agent_radius.release(); e->counters.release(); e->mode_steps.release(); e->metrics.release();
  delete e;
  return ABM_OK;
}

int abm_base_set_params(abm_base_engine_t* e, const double* params, int n_sets) {
  if (!e || !params) return fail(ABM_E_INVALID, "abm_base_set_params: null argument");
  if (n_sets != 1 && n_sets != e->cfg.n_replicates &&
      (long long)n_sets != (long long)e->cfg.n_replicates * e->cfg.n_agents)
    return fail(ABM_E_INVALID, "abm_base_set_params: n_sets must be 1, n_replicates or n_replicates * n_agents");
  ABM_CUDA(cudaSetDevice(e->device));
  ABM_CUDA(cudaMemcpy(e->params.p, params, sizeof(double) * abm::kBaseNParam * n_sets, cudaMemcpyHostToDevice));
  e->n_param_sets = n_sets;
  return ABM_OK;
}

static int agents_xfer(abm_base_engine_t* e, const abm_base_agents_t* s, bool to_dev, int on_device, cudaStream_t st) {
  const size_t n = e->n_agents_total;
  int rc;
#define F(member, buf) if ((rc = xfer(e->buf.p, s->member, s->member, n, to_dev, on_device, st))) return rc;
  F(x, x) F(y, y) F(theta, theta) F(vel, vel) F(w, w) F(u, u) F(collected, collected)
  F(collected_before, collected_before) F(i_priv, i_priv) F(env_status, env_status)
  F(override_mode, override_mode) F(mode, mode) F(patch_id, patch_id) F(novelty, novelty)
#undef F
  return ABM_OK;
}

// The per-agent geometry table from whatever abm_base_set_agent_geometry and abm_base_set_agent_resolution hold (the
// other call's part falls back to the engine-wide values of the config).
static int base_upload_geo(abm_base_engine_t* e) {
  const size_t n = e->n_agents_total;
  const bool has_fov = !e->h_fov0.empty(), has_res = !e->h_res.empty();
  if (!has_fov && !has_res) { e->has_agent_geo = false; return ABM_OK; }   // back to the engine-wide values
  std::vector<abm::BaseAgentGeo> geo(n);
  std::map<std::tuple<double, double, int>, std::pair<int, int>> masks;   // few distinct (FOV, resolution): one scan each
  for (size_t i = 0; i < n; ++i) {
    const double f0 = has_fov ? e->h_fov0[i] : e->cfg.fov0, f1 = has_fov ? e->h_fov1[i] : e->cfg.fov1;
    const int R = has_res ? e->h_res[i] : e->cfg.resolution;
    const double lin = ABM_TWO_PI_D / (double)(R - 1);       // numpy.linspace step of the agent's own grid (agent.py:481)
    const auto key = std::make_tuple(f0, f1, R);
    auto it = masks.find(key);
    if (it == masks.end()) {
      int lo = R, hi = -1;   // FOV mask in stored coordinates (agent.py:594-595) on the numpy linspace grid
      for (int k = 0; k < R; ++k) {
        const double phi = (k == R - 1) ? ABM_PI_D : ((double)k * lin + (-ABM_PI_D));
        if (!(phi < f0) && !(phi > f1)) { if (k < lo) lo = k; if (k > hi) hi = k; }
      }
      it = masks.emplace(key, std::make_pair(lo, hi)).first;
    }
    geo[i] = abm::BaseAgentGeo{f0, f1, has_fov ? e->h_vr[i] : e->cfg.vision_range, lin, it->second.first, it->second.second, R, 0};
  }
  ABM_CUDA(cudaSetDevice(e->device));
  if (!e->agent_geo.p) ABM_CUDA(e->agent_geo.alloc(n));
  ABM_CUDA(cudaMemcpy(e->agent_geo.p, geo.data(), sizeof(abm::BaseAgentGeo) * n, cudaMemcpyHostToDevice));
  e->has_agent_geo = true;
  return ABM_OK;
}

int abm_base_set_agent_geometry(abm_base_engine_t* e, const double* fov0, const double* fov1,
                                const double* vision_range, int n) {
  if (!e) return fail(ABM_E_INVALID, "abm_base_set_agent_geometry: null engine");
  if (n == 0) { e->h_fov0.clear(); e->h_fov1.clear(); e->h_vr.clear(); return base_upload_geo(e); }
  if (!fov0 || !fov1 || !vision_range) return fail(ABM_E_INVALID, "abm_base_set_agent_geometry: null argument");
  if ((size_t)n != e->n_agents_total)
    return fail(ABM_E_INVALID, "abm_base_set_agent_geometry: n must be n_replicates * n_agents (or 0)");
  e->h_fov0.assign(fov0, fov0 + n); e->h_fov1.assign(fov1, fov1 + n); e->h_vr.assign(vision_range, vision_range + n);
  return base_upload_geo(e);
}

int abm_base_set_agent_resolution(abm_base_engine_t* e, const int32_t* resolution, int n) {
  if (!e) return fail(ABM_E_INVALID, "abm_base_set_agent_resolution: null engine");
  if (n == 0) { e->h_res.clear(); return base_upload_geo(e); }
  if (!resolution) return fail(ABM_E_INVALID, "abm_base_set_agent_resolution: null argument");
  if ((size_t)n != e->n_agents_total)
    return fail(ABM_E_INVALID, "abm_base_set_agent_resolution: n must be n_replicates * n_agents (or 0)");
  for (int i = 0; i < n; ++i)
    if (resolution[i] < 2 || resolution[i] > e->cfg.resolution)
      return fail(ABM_E_INVALID, "abm_base_set_agent_resolution: every value must be in [2, the engine's resolution]");
  e->h_res.assign(resolution, resolution + n);
  return base_upload_geo(e);
}

int abm_base_set_agent_radii(abm_base_engine_t* e, const double* radius, int n) {
  if (!e) return fail(ABM_E_INVALID, "abm_base_set_agent_radii: null engine");
  if (n == 0) { e->has_agent_radius = false; return ABM_OK; }   // back to the engine-wide radius of the config
  if (!radius) return fail(ABM_E_INVALID, "abm_base_set_agent_radii: null argument");
  if ((size_t)n != e->n_agents_total)
    return fail(ABM_E_INVALID, "abm_base_set_agent_radii: n must be n_replicates * n_agents (or 0)");
  std::vector<float> h((size_t)n);
  for (int i = 0; i < n; ++i) {
    if (!(radius[i] > 0.0)) return fail(ABM_E_INVALID, "abm_base_set_agent_radii: radii must be > 0");
    h[i] = (float)radius[i];
  }
  ABM_CUDA(cudaSetDevice(e->device));
  if (!e->agent_radius.p) ABM_CUDA(e->agent_radius.alloc((size_t)n));
  ABM_CUDA(cudaMemcpy(e->agent_radius.p, h.data(), sizeof(float) * (size_t)n, cudaMemcpyHostToDevice));
  e->has_agent_radius = true;
  return ABM_OK;
}

int abm_base_set_agents(abm_base_engine_t* e, const abm_base_agents_t* src, int on_device, void* stream) {
  if (!e || !src) return fail(ABM_E_INVALID, "abm_base_set_agents: null argument");
  if (!e->agents_set && (!src->x || !src->y || !src->theta))
    return fail(ABM_E_INVALID, "abm_base_set_agents: x, y, theta are required on the first call");
  ABM_CUDA(cudaSetDevice(e->device));
  int rc = agents_xfer(e, src, true, on_device, (cudaStream_t)stream);
  if (rc) return rc;
  if (!on_device) ABM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  e->agents_set = true;
  return ABM_OK;
}

int abm_base_get_agents(abm_base_engine_t* e, const abm_base_agents_t* dst, int on_device, void* stream) {
  if (!e || !dst) return fail(ABM_E_INVALID, "abm_base_get_agents: null argument");
  if (!e->agents_set) return fail(ABM_E_STATE, "abm_base_get_agents: no agent state has been set");
  ABM_CUDA(cudaSetDevice(e->device));
  int rc = agents_xfer(e, dst, false, on_device, (cudaStream_t)stream);
  if (rc) return rc;
  if (!on_device) ABM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return ABM_OK;
}

static int patches_xfer(abm_base_engine_t* e, const abm_base_patches_t* s, bool to_dev, int on_device, cudaStream_t st) {
  const size_t n = e->n_patches_total;
  int rc;
#define F(member, buf) if ((rc = xfer(e->buf.p, s->member, s->member, n, to_dev, on_device, st))) return rc;
  F(x, px) F(y, py) F(radius, pradius) F(left, pleft) F(quality, pquality) F(id, pid)
#undef F
  return ABM_OK;
}

int abm_base_set_patches(abm_base_engine_t* e, const abm_base_patches_t* src, int on_device, void* stream) {
  if (!e || !src) return fail(ABM_E_INVALID, "abm_base_set_patches: null argument");
  ABM_CUDA(cudaSetDevice(e->device));
  int rc = patches_xfer(e, src, true, on_device, (cudaStream_t)stream);
  if (rc) return rc;
  if (!on_device) ABM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  e->patches_set = true;
  return ABM_OK;
}

int abm_base_get_patches(abm_base_engine_t* e, const abm_base_patches_t* dst, int on_device, void* stream) {
  if (!e || !dst) return fail(ABM_E_INVALID, "abm_base_get_patches: null argument");
  ABM_CUDA(cudaSetDevice(e->device));
  int rc = patches_xfer(e, dst, false, on_device, (cudaStream_t)stream);
  if (rc) return rc;
  if (!on_device) ABM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return ABM_OK;
}

int abm_base_step(abm_base_engine_t* e, int n_steps, const float* inject_dtheta, int inject_on_device,
                  uint32_t phases, void* stream) {
  if (!e) return fail(ABM_E_INVALID, "abm_base_step: null engine");
  if (!e->agents_set) return fail(ABM_E_STATE, "abm_base_step: abm_base_set_agents has not been called");
  if (e->cfg.n_patches > 0 && !e->patches_set)
    return fail(ABM_E_STATE, "abm_base_step: abm_base_set_patches has not been called");
  if (n_steps < 0) return fail(ABM_E_INVALID, "abm_base_step: n_steps < 0");
  if (inject_dtheta && n_steps > 1) return fail(ABM_E_INVALID, "abm_base_step: inject_dtheta needs n_steps == 1");
  ABM_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const abm_base_config_t& c = e->cfg;
  abm::BaseKernelArgs a;
  memset(&a, 0, sizeof(a));
  a.B = c.n_replicates; a.N = c.n_agents; a.P = c.n_patches; a.R = c.resolution; a.W = e->W; a.Tau = c.tau;
  a.visual_exclusion = c.visual_exclusion; a.patchwise_exclusion = c.patchwise_exclusion;
  a.teleport_exploit = c.teleport_exploit; a.regenerate = c.regenerate_patches; a.border_overlap = c.patch_border_overlap;
  a.ghost_mode = c.ghost_mode;
  a.fov0 = c.fov0; a.fov1 = c.fov1; a.mask_lo = e->mask_lo; a.mask_hi = e->mask_hi; a.lin_step = e->lin_step;
  a.width = c.width; a.height = c.height; a.pad = c.window_pad; a.vision_range = c.vision_range; a.radius = c.agent_radius;
  a.patch_radius = c.patch_radius; a.min_quality = c.min_quality; a.max_quality = c.max_quality;
  a.min_units = c.min_units; a.max_units = c.max_units; a.seed = c.seed;
  a.regen_tab = e->has_regen_tab ? e->regen_tab.p : nullptr;
  a.ag = abm::BaseAgentPtrs{e->x.p, e->y.p, e->theta.p, e->vel.p, e->w.p, e->u.p, e->collected.p,
                            e->collected_before.p, e->i_priv.p, e->env_status.p, e->override_mode.p, e->mode.p,
                            e->patch_id.p, e->novelty.p, e->snap_x.p, e->snap_y.p, e->snap_override.p, e->collided.p,
                            e->has_agent_radius ? e->agent_radius.p : nullptr};
  a.pa = abm::BasePatchPtrs{e->px.p, e->py.p, e->pradius.p, e->pleft.p, e->pquality.p, e->pid.p};
  a.params = e->params.p;
  a.agent_geo = e->has_agent_geo ? e->agent_geo.p : nullptr;
  if (e->n_param_sets == 1) { a.param_stride = 0; a.param_stride_agent = 0; }
  else if (e->n_param_sets == e->cfg.n_replicates && e->cfg.n_agents != 1) { a.param_stride = abm::kBaseNParam; a.param_stride_agent = 0; }
  else { a.param_stride = e->cfg.n_agents * abm::kBaseNParam; a.param_stride_agent = abm::kBaseNParam; }
  a.fields_out = e->fields.p; a.counters = e->counters.p; a.mode_steps = e->mode_steps.p;
  a.regen_draws = e->regen_tries > 0 ? e->regen_draws.p : nullptr; a.regen_tries = e->regen_tries;
  if (inject_dtheta) {
    const float* src = inject_dtheta;
    if (!inject_on_device) {
      ABM_CUDA(cudaMemcpyAsync(e->inject.p, inject_dtheta, sizeof(float) * e->n_agents_total, cudaMemcpyHostToDevice, st));
      src = e->inject.p;
    }
    a.inject_dtheta = src;
  }
  // One CTA per replicate (the fused kernel) or one grid per phase with a warp per focal agent?  A replicate's three
  // phases are a chain of dependent latencies: the fused kernel wins when the OTHER replicates fill the SMs meanwhile
  // (or when the run is so small that a step is shorter than three launches); a batch of few, larger replicates is faster
  // with its agents spread over the whole GPU.  Measured crossovers on B200 (scratch/base_n100_probe.py, occlusion +
  // collisions): N = 10 fused at any B (0.021 against 0.030 ms per step at B = 1), N = 25 a tie up to B = 148,
  // N = 50 fused from B ~ 400 (B = 148: 0.154 against 0.103; B = 1024: 0.22 against 0.42), N = 100 from B ~ 600
  // (B = 148: 0.49 against 0.22; B = 1024: 1.13 against 1.56).  ABM_BASE_FUSED=1|0 forces the choice.
  bool separate = getenv("ABM_BASE_SEPARATE_PHASES") != nullptr;
  {
    const char* f = getenv("ABM_BASE_FUSED");
    const long long B = c.n_replicates, N = c.n_agents;
    if (f) separate = separate || atoi(f) == 0;
    else if (!separate) separate = !(N <= 25 || B * 148 >= (long long)e->n_sms * (200 + 4 * N));
  }
  const bool collide = (phases & ABM_BASE_PHASE_COLLISIONS) && c.collide_agents;
  // ONE launch for all n_steps (a CTA per replicate runs the phases in the reference's order, step after step: replicates
  // never interact) whenever the batch fills the GPU that way; otherwise one grid per phase and step
  // (the decision process of a step reads the mode marks the SAME step's environment phase leaves: both in the launch)
  // All n_steps in one launch only while every CTA has an SM to itself: with several replicates per SM, CTAs that drift
  // apart in the step loop execute different phases of this (large) kernel at the same time and evict each other's
  // code from the instruction cache (measured: 0.35 against 0.30 ms per step at 1024 x 50); a launch per step keeps
  // the CTAs of an SM in the same phase.
  const int per_launch = (getenv("ABM_BASE_ONE_STEP_PER_LAUNCH") || c.n_replicates > e->n_sms) ? 1 : n_steps;
  int done = 0;
  a.step = e->step;
  while (!separate && done < n_steps && (phases & (ABM_BASE_PHASE_ENV | ABM_BASE_PHASE_AGENTS))) {
    const int chunk = std::min(per_launch, n_steps - done);
    a.step = e->step;
    if (!abm::launch_base_step(a, phases, collide, chunk, e->n_sms, st)) break;
    ++e->launches;
    if (phases & ABM_BASE_PHASE_AGENTS) e->metric_steps += (unsigned long long)chunk;
    e->step += (unsigned)chunk; e->steps += (unsigned long long)chunk;
    done += chunk;
  }
  for (int s = done; s < n_steps; ++s) {
    a.step = e->step;
    if (collide) { abm::launch_base_collisions(a, st); ++e->launches; }
    if (phases & ABM_BASE_PHASE_ENV) { abm::launch_base_env(a, st); ++e->launches; }
    else {   // agent phase alone: the snapshot is the current state
      ABM_CUDA(cudaMemcpyAsync(e->snap_x.p, e->x.p, sizeof(float) * e->n_agents_total, cudaMemcpyDeviceToDevice, st));
      ABM_CUDA(cudaMemcpyAsync(e->snap_y.p, e->y.p, sizeof(float) * e->n_agents_total, cudaMemcpyDeviceToDevice, st));
      ABM_CUDA(cudaMemcpyAsync(e->snap_override.p, e->override_mode.p, sizeof(int32_t) * e->n_agents_total,
                               cudaMemcpyDeviceToDevice, st));
    }
    if (phases & ABM_BASE_PHASE_AGENTS) { abm::launch_base_agents(a, st); ++e->launches; ++e->metric_steps; }
    ++e->step; ++e->steps;
  }
  ABM_CUDA(cudaGetLastError());
  if (inject_dtheta && !inject_on_device) ABM_CUDA(cudaStreamSynchronize(st));
  return ABM_OK;
}

int abm_base_set_regeneration_params(abm_base_engine_t* e, const double* table, int n) {
  if (!e) return fail(ABM_E_INVALID, "abm_base_set_regeneration_params: null engine");
  if (n == 0) { e->has_regen_tab = false; return ABM_OK; }      // back to the config's values
  if (!table) return fail(ABM_E_INVALID, "abm_base_set_regeneration_params: null argument");
  if (n != e->cfg.n_replicates)
    return fail(ABM_E_INVALID, "abm_base_set_regeneration_params: n must be n_replicates (or 0)");
  for (int b = 0; b < n; ++b) {
    const double* t = table + (size_t)b * 5;
    if (!(t[0] > 0.0) || t[2] < t[1] || t[4] < t[3])
      return fail(ABM_E_INVALID, "abm_base_set_regeneration_params: radius > 0, max >= min expected");
  }
  ABM_CUDA(cudaSetDevice(e->device));
  if (!e->regen_tab.p) ABM_CUDA(e->regen_tab.alloc((size_t)n * 5));
  ABM_CUDA(cudaMemcpy(e->regen_tab.p, table, sizeof(double) * (size_t)n * 5, cudaMemcpyHostToDevice));
  e->has_regen_tab = true;
  return ABM_OK;
}

int abm_base_inject_regeneration(abm_base_engine_t* e, const double* draws, int n_tries) {
  if (!e) return fail(ABM_E_INVALID, "abm_base_inject_regeneration: null engine");
  if (!draws || n_tries <= 0) { e->regen_tries = 0; return ABM_OK; }   // back to the engine's own counter-based RNG
  ABM_CUDA(cudaSetDevice(e->device));
  const size_t n = e->n_patches_total * (size_t)n_tries * 4;
  ABM_CUDA(e->regen_draws.alloc(n));
  ABM_CUDA(cudaMemcpy(e->regen_draws.p, draws, sizeof(double) * n, cudaMemcpyHostToDevice));
  e->regen_tries = n_tries;
  return ABM_OK;
}

int abm_base_get_fields(abm_base_engine_t* e, uint32_t* packed, int on_device, void* stream) {
  if (!e || !packed) return fail(ABM_E_INVALID, "abm_base_get_fields: null argument");
  if (!e->fields.p) return fail(ABM_E_STATE, "abm_base_get_fields: engine created without keep_fields");
  ABM_CUDA(cudaSetDevice(e->device));
  ABM_CUDA(cudaMemcpyAsync(packed, e->fields.p, sizeof(uint32_t) * e->n_agents_total * e->W,
                           on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  if (!on_device) ABM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return ABM_OK;
}

int abm_base_get_counters(abm_base_engine_t* e, uint64_t counters[4], void* stream) {
  if (!e || !counters) return fail(ABM_E_INVALID, "abm_base_get_counters: null argument");
  ABM_CUDA(cudaSetDevice(e->device));
  unsigned long long h[4];
  ABM_CUDA(cudaMemcpyAsync(h, e->counters.p, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  ABM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  counters[0] = h[0]; counters[1] = h[1]; counters[2] = e->launches; counters[3] = e->steps;
  return ABM_OK;
}

int abm_base_metrics(abm_base_engine_t* e, float* out, int on_device, int reset, void* stream) {
  if (!e) return fail(ABM_E_INVALID, "abm_base_metrics: null engine");
  ABM_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int B = e->cfg.n_replicates;
  if (out) {
    float* dst = on_device ? out : e->metrics.p;
    abm::launch_base_metrics(e->collected.p, e->mode_steps.p, B, e->cfg.n_agents, e->metric_steps, dst, st);
    ABM_CUDA(cudaGetLastError());
    if (!on_device) {
      ABM_CUDA(cudaMemcpyAsync(out, dst, sizeof(float) * 6 * (size_t)B, cudaMemcpyDeviceToHost, st));
      ABM_CUDA(cudaStreamSynchronize(st));
    }
  }
  if (reset) {   // the time window of the mode fractions starts again (collected_r keeps counting, like the reference's)
    ABM_CUDA(cudaMemsetAsync(e->mode_steps.p, 0, 4 * sizeof(unsigned int) * (size_t)B, st));
    e->metric_steps = 0;
  }
  return ABM_OK;
}

// ---- stateless function-level entry points ----

int abm_base_projection_field(const abm_base_proj_args_t* args, uint32_t* out_field, double* out_amplitude) {
  if (!args || !out_field) return fail(ABM_E_INVALID, "abm_base_projection_field: null argument");
  if (args->struct_size != (int32_t)sizeof(abm_base_proj_args_t))
    return fail(ABM_E_INVALID, "abm_base_projection_field: struct_size mismatch");
  const int R = args->resolution;
  if (R < 8 || R > 65535) return fail(ABM_E_INVALID, "abm_base_projection_field: bad resolution");
  const int ns = args->n_social, no = args->n_occluders, n = ns + no;
  if (ns < 0 || no < 0 || (ns > 0 && (!args->social_x || !args->social_y)) || (no > 0 && (!args->occluder_x || !args->occluder_y)))
    return fail(ABM_E_INVALID, "abm_base_projection_field: bad object lists");
  const int W = (R + 31) / 32;
  std::vector<float> hx(n > 0 ? n : 1), hy(n > 0 ? n : 1);
  for (int j = 0; j < ns; ++j) { hx[j] = (float)args->social_x[j]; hy[j] = (float)args->social_y[j]; }
  for (int j = 0; j < no; ++j) { hx[ns + j] = (float)args->occluder_x[j]; hy[ns + j] = (float)args->occluder_y[j]; }
  abm::BaseProjArgs a;
  memset(&a, 0, sizeof(a));
  a.R = R; a.W = W; a.n_social = ns; a.n_occ = no;
  a.visual_exclusion = args->visual_exclusion; a.keep_distance = args->keep_distance_info;
  a.fov0 = args->fov0; a.fov1 = args->fov1; a.lin_step = ABM_TWO_PI_D / (double)(R - 1);
  a.radius = args->radius; a.vision_range = args->vision_range;
  a.fx = (float)args->x; a.fy = (float)args->y; a.ftheta = (float)args->orientation;
  a.mask_lo = R; a.mask_hi = -1;
  for (int k = 0; k < R; ++k) {
    const double phi = (k == R - 1) ? ABM_PI_D : ((double)k * a.lin_step + (-ABM_PI_D));
    if (!(phi < a.fov0) && !(phi > a.fov1)) { if (k < a.mask_lo) a.mask_lo = k; if (k > a.mask_hi) a.mask_hi = k; }
  }
  DevBuf<float> dx, dy; DevBuf<uint32_t> df; DevBuf<double> da;
  ABM_CUDA(dx.alloc(hx.size())); ABM_CUDA(dy.alloc(hy.size())); ABM_CUDA(df.alloc(W)); ABM_CUDA(da.alloc(1));
  int rc = ABM_OK;
  cudaError_t ce;
  double amp = 1.0;
  if ((ce = cudaMemcpy(dx.p, hx.data(), sizeof(float) * hx.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (ce = cudaMemcpy(dy.p, hy.data(), sizeof(float) * hy.size(), cudaMemcpyHostToDevice)) != cudaSuccess) {
    rc = fail(ABM_E_CUDA, cudaGetErrorString(ce));
  } else {
    a.ox = dx.p; a.oy = dy.p; a.field = df.p; a.amplitude = da.p;
    abm::launch_base_projection(a, 0);
    if ((ce = cudaGetLastError()) != cudaSuccess ||
        (ce = cudaMemcpy(out_field, df.p, sizeof(uint32_t) * W, cudaMemcpyDeviceToHost)) != cudaSuccess ||
        (ce = cudaMemcpy(&amp, da.p, sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess)
      rc = fail(ABM_E_CUDA, cudaGetErrorString(ce));
  }
  if (out_amplitude) *out_amplitude = amp;
  dx.release(); dy.release(); df.release(); da.release();
  return rc;
}

int abm_base_reloc_lr(const uint32_t* packed_field, int resolution, double amplitude, double vel_now, double v_desired,
                      double reloc_theta_max, double out[2]) {
  if (!packed_field || !out) return fail(ABM_E_INVALID, "abm_base_reloc_lr: null argument");
  if (resolution < 8 || resolution > 65535) return fail(ABM_E_INVALID, "abm_base_reloc_lr: bad resolution");
  const int W = (resolution + 31) / 32;
  DevBuf<uint32_t> df; DevBuf<double> dout;
  ABM_CUDA(df.alloc(W)); ABM_CUDA(dout.alloc(2));
  int rc = ABM_OK;
  cudaError_t ce;
  if ((ce = cudaMemcpy(df.p, packed_field, sizeof(uint32_t) * W, cudaMemcpyHostToDevice)) != cudaSuccess) {
    rc = fail(ABM_E_CUDA, cudaGetErrorString(ce));
  } else {
    abm::launch_base_reloc_lr(df.p, resolution, W, amplitude, vel_now, v_desired, reloc_theta_max, dout.p, 0);
    if ((ce = cudaGetLastError()) != cudaSuccess ||
        (ce = cudaMemcpy(out, dout.p, sizeof(double) * 2, cudaMemcpyDeviceToHost)) != cudaSuccess)
      rc = fail(ABM_E_CUDA, cudaGetErrorString(ce));
  }
  df.release(); dout.release();
  return rc;
}

int abm_vf_dphi(const uint32_t* packed_v, int resolution, int8_t* out) {
  if (!packed_v || !out) return fail(ABM_E_INVALID, "abm_vf_dphi: null argument");
  if (resolution < 2 || resolution > 65535) return fail(ABM_E_INVALID, "abm_vf_dphi: bad resolution");
  const int W = (resolution + 31) / 32;
  DevBuf<uint32_t> dv; DevBuf<signed char> dout;
  ABM_CUDA(dv.alloc(W)); ABM_CUDA(dout.alloc(resolution));
  int rc = ABM_OK;
  cudaError_t ce;
  if ((ce = cudaMemcpy(dv.p, packed_v, sizeof(uint32_t) * W, cudaMemcpyHostToDevice)) != cudaSuccess) {
    rc = fail(ABM_E_CUDA, cudaGetErrorString(ce));
  } else {
    abm::launch_vf_dphi(dv.p, resolution, W, dout.p, 0);
    if ((ce = cudaGetLastError()) != cudaSuccess ||
        (ce = cudaMemcpy(out, dout.p, resolution, cudaMemcpyDeviceToHost)) != cudaSuccess)
      rc = fail(ABM_E_CUDA, cudaGetErrorString(ce));
  }
  dv.release(); dout.release();
  return rc;
}

}  // extern "C"
