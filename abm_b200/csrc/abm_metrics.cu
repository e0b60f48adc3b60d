// On-device summary metrics of a batch of replicates: the per-time-step quantities the
// reference's offline analysis computes from the logged trajectories (abm/loader/data_loader.py):
//   polarization            |sum_i (cos theta_i, sin theta_i)| / N                       (calculate_polarization :1761-1836)
//   mean inter-individual distance  mean over i < j of |p_i - p_j|                     (calculate_interindividual_distance
//                                   (minimal image on the torus)                        :1367-1460, upper triangle :1440-1452)
//   mean nearest-neighbour distance mean_i min_{j != i} |p_i - p_j|                     (calculate_mean_NN_dist :1461-1488)
//   collision flag          any pair with 0 < |p_i - p_j| < 2 r
//   colliding agents        fraction of the agents i with some agent j > i (caller's order) at 0 < |p_i - p_j| < 2 r:
//                           the loader's own indicator (calculate_collision_time :1838-1869 works on the upper triangle
//                           of the distance matrix); its average over time is the "aacoll" value of an experiment
// One CTA per replicate: positions staged in shared memory, thread t takes agents t, t + T, ... against all others.
// fp32 pair distances, fp64 accumulation.  sm_100a only.
#include "abm_common.cuh"

namespace abm {

constexpr int kMetricThreads = 256;

__global__ void __launch_bounds__(kMetricThreads)
vf_metrics_kernel(const float4* __restrict__ rec, const float* __restrict__ theta, const int* __restrict__ perm, int N,
                  int torus, float width, float height, float* __restrict__ out) {
  extern __shared__ float2 pos[];                       // [N]
  __shared__ double red[6][kMetricThreads / 32];
  int* api = reinterpret_cast<int*>(pos + N);           // [N] caller's index of the agent in internal slot i
  const int b = blockIdx.x, tid = threadIdx.x;
  const float4* r = rec + (size_t)b * N;
  const float* th = theta + (size_t)b * N;
  double px = 0.0, py = 0.0;
  float rad = 0.f;
  for (int i = tid; i < N; i += kMetricThreads) {
    const float4 v = r[i];
    pos[i] = make_float2(v.x, v.y);
    api[i] = perm ? perm[(size_t)b * N + i] : i;
    rad = fmaxf(rad, v.z);
    double s, c;
    sincos((double)th[i], &s, &c);
    px += c; py += s;
  }
  __syncthreads();
  const float half_w = 0.5f * width, half_h = 0.5f * height;
  double sum_d = 0.0, sum_nn = 0.0, n_coll_agents = 0.0;
  float coll = 0.f;
  for (int i = tid; i < N; i += kMetricThreads) {
    const float2 pi = pos[i];
    const int ai = api[i];
    bool mine = false;
    float nn = 3.0e38f, acc = 0.f;
    const float lim = rad + rad;                        // uniform radius assumed by the reference's criterion (2 * RADIUS_AGENT)
    for (int j = 0; j < N; ++j) {
      const float2 pj = pos[j];                         // broadcast read
      float dx = pj.x - pi.x, dy = pj.y - pi.y;
      if (torus) {
        if (fabsf(dx) > half_w) dx -= copysignf(width, dx);
        if (fabsf(dy) > half_h) dy -= copysignf(height, dy);
      }
      const float d = sqrtf(fmaf(dx, dx, dy * dy));
      if (j > i) acc += d;
      if (j != i) nn = fminf(nn, d);
      if (d > 0.f && d < lim) { coll = 1.f; mine |= api[j] > ai; }
      if ((j & 255) == 255) { sum_d += (double)acc; acc = 0.f; }   // keep the fp32 partial sums short
    }
    sum_d += (double)acc;
    if (N > 1) sum_nn += (double)nn;
    if (mine) n_coll_agents += 1.0;
  }
  double v[6] = {px, py, sum_d, sum_nn, (double)coll, n_coll_agents};
  for (int k = 0; k < 6; ++k) {
    double x = v[k];
    for (int off = 16; off > 0; off >>= 1) {
      const double y = __shfl_xor_sync(0xffffffffu, x, off);
      x = (k == 4) ? fmax(x, y) : x + y;
    }
    if ((tid & 31) == 0) red[k][tid >> 5] = x;
  }
  __syncthreads();
  if (tid == 0) {
    double t[6];
    for (int k = 0; k < 6; ++k) {
      t[k] = red[k][0];
      for (int w = 1; w < kMetricThreads / 32; ++w) t[k] = (k == 4) ? fmax(t[k], red[k][w]) : t[k] + red[k][w];
    }
    const double n = (double)N, npairs = 0.5 * n * (n - 1.0);
    float* o = out + 5 * (size_t)b;
    o[0] = (float)(sqrt(t[0] * t[0] + t[1] * t[1]) / n);
    o[1] = npairs > 0.0 ? (float)(t[2] / npairs) : 0.f;
    o[2] = N > 1 ? (float)(t[3] / n) : 0.f;
    o[3] = (float)t[4];
    o[4] = (float)(t[5] / n);
  }
}

void launch_vf_metrics(const float4* rec, const float* theta, const int* perm, int B, int N, int torus, float width,
                       float height, float* out, cudaStream_t stream) {
  vf_metrics_kernel<<<B, kMetricThreads, (sizeof(float2) + sizeof(int)) * (size_t)N, stream>>>(rec, theta, perm, N, torus, width,
                                                                                                 height, out);
}

}  // namespace abm
