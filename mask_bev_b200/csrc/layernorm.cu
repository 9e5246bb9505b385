// K3+LN — scatter to the BEV canvas fused with the LayerNorm that follows it (sm_100a). SURVEY.md §8 row f1.
//
// Replaces mask_bev_encoders.py:91-92: `pseudo_img = self.middle_encode(...)` followed by
// `self._layer_norm(pseudo_img)` with `nn.LayerNorm([C, ny, nx], eps=1e-3)` (:75) — per frame a normalisation
// over ALL C*ny*nx elements with an element-wise affine of that same shape. Unfused, torch reads the canvas, the
// weight and the bias and writes the canvas again (4x the scatter's traffic). Here:
//   * the statistics come from the pillar features alone (every other cell is an exact zero):
//     sum = sum_p sum_c f, sumsq likewise, N = C*ny*nx; fp64, fixed-order two-stage reduction (deterministic);
//   * one streaming pass writes y = ((x - mean_b) * rstd_b) * w + bias with x = feature or 0: a warp owns a run
//     of 512 cells and a chunk of channels; per channel it loads the run's weight / bias ONCE and then walks the
//     frames of the batch (table and features from L1/L2), so weight + bias are read once per batch and the
//     canvas is written once: B*C*G*4 + 2*C*G*4 bytes instead of ~4*B*C*G*4.
// Forward only (inference / no-grad); the autograd path keeps torch's LayerNorm after K3.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace mbev {
namespace {

constexpr int kThreads = 256;
constexpr int kRun = 512;      // cells per warp task
constexpr int kStatBlocks = 64;  // partial sums per frame

// partial (sum, sumsq) of the feature rows of frame b, slice j of kStatBlocks: one warp per pillar row at a time
__global__ void __launch_bounds__(kThreads)
k_ln_partials(const float *__restrict__ feats, const int *__restrict__ pillar_base, const int C,
              double2 *__restrict__ partial) {
  __shared__ double s_a[kThreads / 32], s_b[kThreads / 32];
  const int b = blockIdx.y, j = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p0 = pillar_base[b], p1 = pillar_base[b + 1];
  const long long n = p1 - p0;
  const int lo = p0 + static_cast<int>(n * j / kStatBlocks), hi = p0 + static_cast<int>(n * (j + 1) / kStatBlocks);
  double a = 0.0, q = 0.0;
  for (int p = lo + warp; p < hi; p += kThreads / 32) {
    const float *row = feats + static_cast<size_t>(p) * C;
    for (int c = lane; c < C; c += 32) {
      const double v = static_cast<double>(__ldg(row + c));
      a += v;
      q += v * v;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    s_a[warp] = a;
    s_b[warp] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tq = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      ta += s_a[w];
      tq += s_b[w];
    }
    partial[b * kStatBlocks + j] = make_double2(ta, tq);
  }
}

// stats[b] = (mean, rstd) — biased variance as nn.LayerNorm
__global__ void k_ln_finalize(const double2 *__restrict__ partial, const int batch, const double count,
                              const double eps, float2 *__restrict__ stats) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double a = 0.0, q = 0.0;
  for (int j = 0; j < kStatBlocks; ++j) {
    const double2 v = partial[b * kStatBlocks + j];
    a += v.x;
    q += v.y;
  }
  const double mean = a / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[b] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + eps)));
}

__global__ void __launch_bounds__(kThreads)
k_scatter_ln(const float *__restrict__ feats, const int *__restrict__ table, const float2 *__restrict__ stats,
             const float *__restrict__ lnw, const float *__restrict__ lnb, const int batch, const int C, const int G,
             const int runs, const int csplit, float *__restrict__ out) {
  __shared__ float2 s_stats[MBEV_MAX_BATCH];
  for (int i = threadIdx.x; i < batch; i += kThreads) s_stats[i] = stats[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nw = gridDim.x * (kThreads / 32);
  const int cper = (C + csplit - 1) / csplit;
  for (int task = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); task < runs * csplit; task += nw) {
    const int run = task / csplit;
    const int ch0 = (task - run * csplit) * cper, ch1 = min(C, ch0 + cper);
    const int g0 = run * kRun + 4 * lane;
    bool inb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) inb[k] = g0 + 128 * k < G;
    for (int ch = ch0; ch < ch1; ++ch) {
      float4 w[4], bi[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        w[k] = bi[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inb[k]) {
          w[k] = __ldg(reinterpret_cast<const float4 *>(lnw + static_cast<size_t>(ch) * G + g0 + 128 * k));
          bi[k] = __ldg(reinterpret_cast<const float4 *>(lnb + static_cast<size_t>(ch) * G + g0 + 128 * k));
        }
      }
      for (int b = 0; b < batch; ++b) {
        const float mean = s_stats[b].x, rstd = s_stats[b].y;
        const int *tb = table + static_cast<size_t>(b) * G + g0;
        float *o = out + (static_cast<size_t>(b) * C + ch) * G + g0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (!inb[k]) continue;
          const int4 pid = __ldg(reinterpret_cast<const int4 *>(tb + 128 * k));
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if ((pid.x & pid.y & pid.z & pid.w) >= 0) {
            if (pid.x >= 0) x.x = __ldg(feats + static_cast<size_t>(pid.x) * C + ch);
            if (pid.y >= 0) x.y = __ldg(feats + static_cast<size_t>(pid.y) * C + ch);
            if (pid.z >= 0) x.z = __ldg(feats + static_cast<size_t>(pid.z) * C + ch);
            if (pid.w >= 0) x.w = __ldg(feats + static_cast<size_t>(pid.w) * C + ch);
          }
          float4 y;  // ((x - mean) * rstd) * w + b, the operation order of torch's LayerNorm kernel
          y.x = __fmaf_rn(__fmul_rn(__fsub_rn(x.x, mean), rstd), w[k].x, bi[k].x);
          y.y = __fmaf_rn(__fmul_rn(__fsub_rn(x.y, mean), rstd), w[k].y, bi[k].y);
          y.z = __fmaf_rn(__fmul_rn(__fsub_rn(x.z, mean), rstd), w[k].z, bi[k].z);
          y.w = __fmaf_rn(__fmul_rn(__fsub_rn(x.w, mean), rstd), w[k].w, bi[k].w);
          st_global_v4_stream_nc(o + 128 * k, y);
        }
      }
    }
  }
}

struct LnWs {
  double2 *partial;
  size_t bytes;
};

LnWs carve(void *ws, int batch) {
  Carver c(ws);
  LnWs w;
  w.partial = c.take<double2>(static_cast<size_t>(batch) * kStatBlocks);
  w.bytes = c.off;
  return w;
}

}  // namespace
}  // namespace mbev

using namespace mbev;

extern "C" int mbev_scatter_layernorm_supported(int batch, int c_out, int ny, int nx, const float *out,
                                                const float *ln_weight, const float *ln_bias) {
  if (batch < 1 || batch > MBEV_MAX_BATCH || c_out < 1 || ny < 1 || nx < 1) return 0;
  const int64_t G = static_cast<int64_t>(ny) * nx;
  if ((G & 3) || G * batch > 0x7fffffffLL) return 0;
  if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(ln_weight) | reinterpret_cast<uintptr_t>(ln_bias)) & 15)
    return 0;
  return 1;
}

extern "C" int mbev_scatter_layernorm_workspace_bytes(int batch, size_t *bytes) {
  if (!bytes || batch < 1 || batch > MBEV_MAX_BATCH) return MBEV_ERR_BAD_ARG;
  *bytes = carve(nullptr, batch).bytes;
  return MBEV_OK;
}

extern "C" int mbev_scatter_layernorm_forward(const float *feats, const int32_t *cell_table, const int32_t *pillar_base,
                                              int batch, int c_out, int ny, int nx, const float *ln_weight,
                                              const float *ln_bias, float eps, float *out, float *stats_out,
                                              void *workspace, size_t workspace_bytes, void *stream_) {
  if (!cell_table || !pillar_base || !ln_weight || !ln_bias || !out || !stats_out || !workspace) return MBEV_ERR_BAD_ARG;
  if (!mbev_scatter_layernorm_supported(batch, c_out, ny, nx, out, ln_weight, ln_bias)) return MBEV_ERR_UNSUPPORTED;
  if (!(eps >= 0.f)) return MBEV_ERR_BAD_ARG;
  const LnWs w = carve(workspace, batch);
  if (workspace_bytes < w.bytes) return MBEV_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int G = ny * nx;
  k_ln_partials<<<dim3(kStatBlocks, batch), kThreads, 0, stream>>>(feats, pillar_base, c_out, w.partial);
  MBEV_CHECK_LAUNCH();
  float2 *stats = reinterpret_cast<float2 *>(stats_out);
  k_ln_finalize<<<(batch + 127) / 128, 128, 0, stream>>>(w.partial, batch, static_cast<double>(c_out) * G,
                                                        static_cast<double>(eps), stats);
  MBEV_CHECK_LAUNCH();
  const int runs = (G + kRun - 1) / kRun;
  const int want_warps = kNumSMs * 3 * (kThreads / 32);
  int csplit = 1;
  while (csplit < 32 && runs * csplit < want_warps && c_out / (2 * csplit) >= 1) csplit *= 2;
  const int tasks = runs * csplit;
  const int blocks = std::min((tasks + kThreads / 32 - 1) / (kThreads / 32), kNumSMs * 3);
  k_scatter_ln<<<blocks, kThreads, 0, stream>>>(feats, cell_table, stats, ln_weight, ln_bias, batch, c_out, G, runs,
                                                csplit, out);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}
