// Shared declarations for the abm_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ABM_PI_D 3.141592653589793238462643383279502884
#define ABM_TWO_PI_D (2.0 * ABM_PI_D)

#include <mutex>

namespace abm {

// cudaFuncAttributeMaxDynamicSharedMemorySize is an attribute of a kernel ON ONE DEVICE: a process that creates
// engines on several devices has to opt in on each of them.  One instance per kernel (variant), e.g. a function-local
// static of the launch function; thread-safe.
struct SmemOptIn {
  std::mutex mu;
  size_t configured[64] = {};
  template <typename Kernel>
  void ensure(Kernel kernel, size_t smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    const bool tracked = dev >= 0 && dev < 64;
    if (!tracked || smem > configured[dev]) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (tracked) configured[dev] = smem;
    }
  }
};

constexpr int kMaxThreads = 256;     // focal agents per CTA (one thread each)
constexpr int kRecTile = 512;        // neighbour records per shared-memory stage (one-thread-per-focal-agent kernel)
constexpr int kWarpTile = 128;       // records per culling tile of the warp-per-focal-agent kernel (finer boxes, fewer candidates)
constexpr int kQueueCap = 1024;      // per-CTA queue of pairs deferred to the fp64 path

// per-replicate flocking parameters, order of ABM_VF_* in include/abm_b200.h
struct VFParams6 { double gam, v0, alp0, alp1, bet0, bet1; };

// one entry of the integration-grid table (Phi = arange(-pi, pi, 2pi/R), vf_agent.py:44)
struct __align__(32) PhiLut {
  double c;   // cos(Phi_k)
  double s;   // sin(Phi_k)
  double pc;  // sum_{m<k} cos(Phi_m)
  double ps;  // sum_{m<k} sin(Phi_m)
};

struct VFKernelArgs {
  int B, N, R, W;                 // replicates, agents / replicate, bins, words / field
  int tile_begin, tile_count;     // focal agents handled by this engine
  int tile_cycle, tile_phase;     // > 1: cyclic blocks of kWarpTile slots instead of the contiguous range (vf_tile_slot)
  int fov_px0, fov_px1;
  int boundary, limit_movement;
  int phi_ok;                     // len(arange grid) == R (vf_agent.py:267); else dv = dphi = 0
  uint32_t flags;
  // fp32 pair path (see BinConsts in abm_vf_device.cuh); kept as plain kernel arguments so that
  // the hot loop reads them straight from the constant bank
  float inv_step;                 // (R-1) / 2pi
  float t_half;                   // 0.5 (even R) or 1.0 (odd R)
  int k_bias;                     // k_off - 0x4B400000 + 32
  float y_scale;                  // R / 2pi  (proj_size / 2 = atan(r/d) * y_scale)
  float thr_k, thr_h0, thr_h1, ca_guard;
  // the same constants pre-scaled to bins so that every hot-loop instruction needs at most ONE
  // run-time constant (read straight from the constant bank, no registers, no LDC)
  float ac[7];                    // atan polynomial coefficients * inv_step (ascending powers of z)
  float half_pi_b, pi_b;          // pi/2 * inv_step, pi * inv_step
  float seam_b;                   // ca_guard * inv_step
  float nthr_h1;                  // -thr_h1
  // symmetric kernel (abm_vf_sym.cu): wider k guard band (the closed angle is a difference of two rounded bin
  // angles), absolute h guard band of the fast path (h <= 16), the common radius
  float sym_thr_h, sym_radius;
  // fast-path limits on qs = (r / d) * R / 2pi: the four-term arctangent series is exact to the guard band up to q = 0.18,
  // and the interval must fit three (symmetric kernel: h <= 32) or two (warp kernel: h <= 16) row words; also catches
  // NaN / inf (coincident centres) and the sign change of the truncated series at large q
  float sym_qs_max, sym_qs_max2, warp_qs_max;   // three-word / two-word fast path of the symmetric kernel, warp kernel
  // line following (vf_supcalc.follow_lines_local): the rasterised line map [lm_d0][lm_d1] (first axis x) or nullptr
  const float* line_map;
  int lm_d0, lm_d1;
  double lm_sr, lm_sd;            // sensor radius (VFAgent.sensor_size) and distance (vf_agent.py:30-31)
  uint32_t sym_tie32, sym_seam32; // guard bands of the binary-angle bin index, in 2^-32 bins / 2^-32 turns
  uint32_t opaque_zero;           // always 0: OR-ed into loop constants so that ptxas keeps them in registers
  int full_fov;                   // fov covers every bin: any interval with h >= 1 is visible
  int fov0p;                      // fov_px0 + 33: first visible padded position
  unsigned span;                  // fov_px1 - fov_px0 - 1: number of visible positions
  float width, height, half_w, half_h;
  float cull_scale;               // cot^2(2pi/R) * (1 + margin): cull^2 = r^2 * cull_scale
  // fp64 exact path / epilogue
  double lin_step;                // linspace step fl(2*pi_fl / (R-1))
  double dphi;                    // 2pi / R
  double kappa_r, kappa_i;        // 1 / (exp(i d) - 1), d = step of the integration grid (vf_flock_terms_edges)
  double rot_c, rot_s;            // cos d, sin d
  double width_d, height_d, pad_d, max_vel, max_th;
  // state
  const float4* rec_in;           // (x, y, r, cull^2) per agent, B*N
  float4* rec_out;
  float* theta;                   // B*N, updated in place (only the owner thread touches it)
  float* vel;
  const double* params;           // n_sets * 6
  int param_stride;               // 0 (shared) or 6
  const float* ov_alp0;           // nullable per-agent overrides, NaN = none
  const float* ov_bet0;
  const float* ov_v0;
  const int* perm;                // nullable: internal slot -> API index (outputs are written in API order)
  // distance culling by whole record tiles (CULL variants, spatially sorted state): bounding box (xmin, ymin,
  // xmax, ymax) and largest cull^2 of every tile of kRecTile records; nullptr: every tile is visited
  int cull_tile;                  // records per culling tile (kRecTile or kWarpTile)
  const float4* tile_bbox;
  const float* tile_cull2;
  float bbox_slack;               // 2 * (r_max - r_min): positions are top-left corners, distances are between centres
  // fused tile exchange of one large swarm over NVLink peer memory (abm_vf_ipc_attach): the epilogue stores every new
  // record into the peers' record tables as well; a kernel starts only when every peer has published the previous step
  // (flags in this GPU's memory, written by the peers), and a one-thread kernel behind it publishes the step to all ranks
  int n_sms;                      // multiprocessors of the device (grid shaping)
  int n_peers;                    // 0: off
  int my_rank;                    // index of this engine among the attached engines
  uint32_t step_no;               // number of steps completed before this launch
  float4* peer_rec_out[7];        // the peers' tables being written in this step
  uint32_t* peer_flags[7];        // the peers' flag arrays (entry my_rank is ours to write)
  uint32_t* xflags;               // this GPU's flag array [8]: xflags[r] = steps published by rank r
  // the warp kernel closes a step of the fused exchange itself (abm_vf_warp.cu): its last CTA (ticket counter) computes
  // the bounding boxes of this rank's record tiles in the table just written, stores them into every rank's box table
  // of the next step and publishes the step
  uint32_t* step_ticket;          // nullptr: no ticket / publish in this launch
  float4* bbox_out;               // nullptr: the boxes of the next step are not produced by this launch
  float4* peer_bbox_out[7];
  const PhiLut* lut;              // R + 1 entries
  uint32_t* fields_out;           // nullable, B*tile*W
  double* terms_out;              // nullable, B*tile*6
  unsigned long long* counters;   // 4
};

// internal slot of focal agent li of this engine's tile
__host__ __device__ inline int vf_tile_slot(const VFKernelArgs& a, int li) {
  if (a.tile_cycle <= 1) return a.tile_begin + li;
  return (((li / kWarpTile) * a.tile_cycle + a.tile_phase) * kWarpTile) + (li % kWarpTile);
}

void launch_vf_step(const VFKernelArgs& a, bool uniform_r, bool cull, cudaStream_t stream);
void launch_tile_bbox(const float4* rec, int B, int N, int tile, float4* bbox, float* cull2, cudaStream_t stream);
constexpr int kMaxTileList = 1024;   // tiles per replicate the culling list can hold (N <= 524288 at 512 records per tile,
                                     // N <= 131072 at 128)
size_t vf_step_smem_bytes(int threads, int W);
// symmetric kernel (abm_vf_sym.cu): every unordered pair once, all rows of a replicate in one CTA
size_t vf_sym_smem_bytes(int Np, int W, bool wide3, bool het = false);
bool vf_sym_applicable(const VFKernelArgs& a, bool uniform_r, bool cull, size_t smem_limit);
void launch_vf_step_sym(const VFKernelArgs& a, bool wide3, bool uniform_r, cudaStream_t stream);
int vf_step_threads(int tile_count, int n_replicates, int n_sms);
// warp-per-focal-agent kernel (abm_vf_warp.cu): one large sparse swarm and its tiles
void launch_vf_step_warp(const VFKernelArgs& a, bool cull, bool uniform_r, cudaStream_t stream);
int vf_warp_focal_per_cta(long long focal_total, int n_sms);
bool launch_vf_step_warp_multi(const VFKernelArgs& a, bool uniform_r, int n_steps, cudaStream_t stream);
bool launch_vf_step_warp_cluster(const VFKernelArgs& a, bool uniform_r, int n_steps, cudaStream_t stream);
struct VFPeerFlags { uint32_t* p[7]; };
void launch_vf_publish(const VFKernelArgs& a, cudaStream_t stream);   // fused tile exchange: step done -> all ranks

struct VFProjArgs {
  int R, W, n_obj, boundary;
  int fov_px0, fov_px1;
  float inv_step, t_half; int k_bias; float y_scale, thr_k, thr_h0, thr_h1, ca_guard;
  float width, height, half_w, half_h;
  double lin_step, width_d, height_d;
  float fx, fy, fr, ftheta;       // focal agent (fp32 state)
  double vision_range;            // < 0: none
  const float* ox; const float* oy; const float* osz;
  uint32_t* rows;                 // n_obj * W, stored order
};
void launch_vf_projection(const VFProjArgs& a, cudaStream_t stream);
struct CSProjArgs {
  int R, W, n_obj;
  double lin_step, fov0, fov1;
  double fx, fy, fr, ftheta;      // focal agent, float64 as passed
  double max_proj_size;           // < 0: none
  const double* ox; const double* oy;
  const uint32_t* keep;           // W words, stored order: bins with fov0 <= linspace angle <= fov1
  uint32_t* rows;                 // n_obj * W, stored order
};
void launch_cs_projection(const CSProjArgs& a, cudaStream_t stream);
void launch_vf_terms(const uint32_t* packed_v, int R, int W, double vel, const VFParams6* prm,
                     const PhiLut* lut, double dphi, double* out6, cudaStream_t stream);
// radius_minmax: 2 uints (bit patterns of min / max radius; init 0x7f800000 / 0)
// perm (nullable): internal slot -> caller's index; the records are written in internal order
void launch_pack_records(const float* x, const float* y, const float* r, const int* perm, int N, float cull_scale,
                         float4* rec, unsigned* radius_minmax, long long n, cudaStream_t stream);
// the same from / to ONE interleaved (x, y, heading, speed) array in the caller's order (abm_set_state_packed / get)
void launch_pack_state4(const float4* s4, const float* r, const int* perm, int N, float cull_scale, float4* rec, float* theta,
                        float* vel, unsigned* radius_minmax, long long n, cudaStream_t stream);
void launch_unpack_state4(const float4* rec, const float* theta, const float* vel, const int* perm, int N, float4* out,
                          long long n, cudaStream_t stream);
// perm (nullable): internal slot -> API index inside the replicate; x / y are written in API order
void launch_unpack_records(const float4* rec, const int* perm, int N, float* x, float* y, long long n,
                           cudaStream_t stream);

// summary metrics (abm_metrics.cu): out[b * 5 + {0: polarization, 1: mean inter-individual distance, 2: mean nearest-
// neighbour distance, 3: collision flag, 4: fraction of colliding agents}]; perm (nullable): internal slot -> caller's index
void launch_vf_metrics(const float4* rec, const float* theta, const int* perm, int B, int N, int torus, float width,
                       float height, float* out, cudaStream_t stream);

// spatial re-ordering (abm_vf_sort.cu)
size_t vf_sort_temp_bytes(int B, int N);
cudaError_t vf_sort_order(const float4* rec, int B, int N, float x0, float y0, float extent, void* temp, size_t temp_bytes,
                          uint32_t* keys_in, uint32_t* keys_out, int* vals_in, int* order, int* offsets,
                          cudaStream_t stream);
void launch_gather_f4(const float4* in, const int* order, float4* out, int N, long long n, cudaStream_t s);
void launch_gather_f32(const float* in, const int* order, float* out, int N, long long n, cudaStream_t s);
void launch_gather_i32(const int* in, const int* order, int* out, int N, long long n, cudaStream_t s);
void launch_scatter_f32(const float* in, const int* perm, float* out, int N, long long n, cudaStream_t s);
void launch_iota(int* p, int N, long long n, cudaStream_t s);

}  // namespace abm
