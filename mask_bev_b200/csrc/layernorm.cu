// K3+LN — scatter to the BEV canvas fused with the LayerNorm that follows it (sm_100a). SURVEY.md §8 row f1.
//
// Replaces mask_bev_encoders.py:91-92: `pseudo_img = self.middle_encode(...)` followed by
// `self._layer_norm(pseudo_img)` with `nn.LayerNorm([C, ny, nx], eps=1e-3)` (:75) — per frame a normalisation
// over ALL C*ny*nx elements with an element-wise affine of that same shape. Unfused, torch reads the canvas, the
// weight and the bias and writes the canvas again (4x the scatter's traffic). Here:
//   * the statistics come from the pillar features alone (every other cell is an exact zero):
//     sum = sum_p sum_c f, sumsq likewise, N = C*ny*nx; fp64, fixed-order two-stage reduction (deterministic);
//   * one streaming pass writes y = ((x - mean_b) * rstd_b) * w + bias with x = feature or 0, composed as in
//     k_scatter_warp; tasks are ordered so that all frames of a run are in flight together and the run's weight /
//     bias come from HBM once and from L2 for the other frames: ~B*C*G*4 + 2*C*G*4 bytes of DRAM traffic instead of
//     ~4*B*C*G*4 (measured history: a frame-inner loop with dependent table/feature loads per store 3.5-4.0 ms).
// Forward only (inference / no-grad); the autograd path keeps torch's LayerNorm after K3.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace mbev {
namespace {

constexpr int kThreads = 256;
constexpr int kRun = 256;      // cells per warp task of k_scatter_ln
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

// A warp owns (run of 256 cells, channel chunk, ONE frame) and composes x exactly as k_scatter_warp does (pillar
// ids in registers, feature values requested one plane ahead), multiplies by the run's weight / adds its bias and
// streams the result out. Weight and bias of plane ch+1 are requested one plane ahead as well: they come from L2
// (or HBM for the first frame that touches them) and were the kernel's main stall when loaded at the point of use
// (ncu: long-scoreboard 11 per issue on the FMA lines, 16 warps per SM at 128 registers; hence the 256-cell run:
// half the registers per warp, 24 warps per SM). Tasks are ordered run-major / frame-minor, so the resident warps
// work on the same runs for all frames at once: the weight / bias lines of a run are fetched from HBM once and hit
// L2 for the other frames (ncu: 886 MB of DRAM reads = weight + bias + features + table, each once).
constexpr int kLnK = 2;  // 128-cell groups per run

__global__ void __launch_bounds__(kThreads, 3)
k_scatter_ln(const float *__restrict__ feats, const int *__restrict__ table, const float2 *__restrict__ stats,
             const float *__restrict__ lnw, const float *__restrict__ lnb, const int batch, const int C, const int G,
             const int runs, const int csplit, float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int nw = gridDim.x * (kThreads / 32);
  const int cper = (C + csplit - 1) / csplit;
  const int per_run = batch * csplit;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int task = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); task < runs * per_run; task += nw) {
    const int run = task / per_run;
    const int rem = task - run * per_run;
    const int cs = rem / batch, b = rem - cs * batch;
    const int ch0 = cs * cper, ch1 = min(C, ch0 + cper);
    const int g0 = run * kRun + 4 * lane;
    const float2 st = __ldg(stats + b);
    const float mean = st.x, rstd = st.y;
    int4 pid[kLnK];
    bool inb[kLnK], any = false;
#pragma unroll
    for (int k = 0; k < kLnK; ++k) {
      inb[k] = g0 + 128 * k < G;
      pid[k] = inb[k] ? __ldg(reinterpret_cast<const int4 *>(table + static_cast<size_t>(b) * G + g0 + 128 * k))
                      : make_int4(-1, -1, -1, -1);
      any |= (pid[k].x & pid[k].y & pid[k].z & pid[k].w) >= 0;
    }
    auto load_plane = [&](int ch, float4 (&v)[kLnK], float4 (&w)[kLnK], float4 (&bi)[kLnK]) {
#pragma unroll
      for (int k = 0; k < kLnK; ++k) {
        v[k] = w[k] = bi[k] = z;
        if (inb[k]) {
          w[k] = __ldg(reinterpret_cast<const float4 *>(lnw + static_cast<size_t>(ch) * G + g0 + 128 * k));
          bi[k] = __ldg(reinterpret_cast<const float4 *>(lnb + static_cast<size_t>(ch) * G + g0 + 128 * k));
        }
        if (any) {
          if (pid[k].x >= 0) v[k].x = __ldg(feats + static_cast<size_t>(pid[k].x) * C + ch);
          if (pid[k].y >= 0) v[k].y = __ldg(feats + static_cast<size_t>(pid[k].y) * C + ch);
          if (pid[k].z >= 0) v[k].z = __ldg(feats + static_cast<size_t>(pid[k].z) * C + ch);
          if (pid[k].w >= 0) v[k].w = __ldg(feats + static_cast<size_t>(pid[k].w) * C + ch);
        }
      }
    };
    float4 xn[kLnK], wn[kLnK], bn[kLnK];
    load_plane(ch0, xn, wn, bn);
    float *o = out + (static_cast<size_t>(b) * C) * G + g0;
    for (int ch = ch0; ch < ch1; ++ch) {
      float4 x[kLnK], w[kLnK], bi[kLnK];
#pragma unroll
      for (int k = 0; k < kLnK; ++k) {
        x[k] = xn[k];
        w[k] = wn[k];
        bi[k] = bn[k];
      }
      if (ch + 1 < ch1) load_plane(ch + 1, xn, wn, bn);
#pragma unroll
      for (int k = 0; k < kLnK; ++k) {
        if (!inb[k]) continue;
        float4 y;  // ((x - mean) * rstd) * w + b, the operation order of torch's LayerNorm kernel
        y.x = __fmaf_rn(__fmul_rn(__fsub_rn(x[k].x, mean), rstd), w[k].x, bi[k].x);
        y.y = __fmaf_rn(__fmul_rn(__fsub_rn(x[k].y, mean), rstd), w[k].y, bi[k].y);
        y.z = __fmaf_rn(__fmul_rn(__fsub_rn(x[k].z, mean), rstd), w[k].z, bi[k].z);
        y.w = __fmaf_rn(__fmul_rn(__fsub_rn(x[k].w, mean), rstd), w[k].w, bi[k].w);
        st_global_v4_stream_nc(o + static_cast<size_t>(ch) * G + 128 * k, y);
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
  const int want_warps = kNumSMs * 3 * (kThreads / 32) * 4;  // several tasks per resident warp (tail balance)
  int csplit = 1;
  while (csplit < 32 && static_cast<int64_t>(runs) * batch * csplit < want_warps && c_out / (2 * csplit) >= 1) csplit *= 2;
  const int64_t tasks64 = static_cast<int64_t>(runs) * batch * csplit;
  if (tasks64 > 0x7fffffffLL) return MBEV_ERR_UNSUPPORTED;
  const int tasks = static_cast<int>(tasks64);
  const int blocks = (tasks + kThreads / 32 - 1) / (kThreads / 32);  // no grid-stride: the CTA scheduler balances
  k_scatter_ln<<<blocks, kThreads, 0, stream>>>(feats, cell_table, stats, ln_weight, ln_bias, batch, c_out, G, runs,
                                                csplit, out);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}
