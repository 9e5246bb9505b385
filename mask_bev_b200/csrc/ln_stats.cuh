// Per-frame statistics of nn.LayerNorm([C, ny, nx]) over the scattered canvas, from the pillar features alone (every
// other cell is an exact zero): shared by the fused scatter + LayerNorm (layernorm.cu) and the pillar patch embedding
// (patch_embed.cu). fp64, fixed-order two-stage reduction (run-to-run identical).
#pragma once
#include "common.cuh"

namespace mbev {
namespace {

constexpr int kStatThreads = 256;
constexpr int kStatBlocks = 64;  // partial sums per frame

// partial (sum, sumsq) of the feature rows of frame b, slice j of kStatBlocks: one warp per pillar row at a time
__global__ void __launch_bounds__(kStatThreads)
k_ln_partials(const float *__restrict__ feats, const int *__restrict__ pillar_base, const int C,
              double2 *__restrict__ partial) {
  __shared__ double s_a[kStatThreads / 32], s_b[kStatThreads / 32];
  const int b = blockIdx.y, j = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p0 = pillar_base[b], p1 = pillar_base[b + 1];
  const long long n = p1 - p0;
  const int lo = p0 + static_cast<int>(n * j / kStatBlocks), hi = p0 + static_cast<int>(n * (j + 1) / kStatBlocks);
  double a = 0.0, q = 0.0;
  for (int p = lo + warp; p < hi; p += kStatThreads / 32) {
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
    for (int w = 0; w < kStatThreads / 32; ++w) {
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


}  // namespace
}  // namespace mbev
