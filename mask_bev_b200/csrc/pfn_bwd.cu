// K2' — backward of the pillar feature net w.r.t. its parameters (sm_100a).
//
// The reference gets this from autograd over the dense (P, T, .) op sequence: max-backward
// (one-hot to the arg-max slot), ReLU mask, BatchNorm backward (two more reductions over P*T), Linear backward
// (dW = X^T dY, dX = dY W). Raw points carry no gradient.
//
// Here the same algebra runs in compact ROW SPACE: R = N_k real rows + one virtual row per pillar with padding
// (weight w = T - n_p). Because all padded slots of a pillar share their forward values and the backward is linear
// in the per-slot incoming gradient, the aggregated virtual row is exact (DESIGN.md "Backward"):
//     dy_v = scale * (DZ_v - w * (S1/M + xhat_v * S2/M)),   S1 = sum dz,  S2 = sum dz*xhat over all slots.
// Layer-wise pipeline over global row-major buffers (recomputes the forward first, so nothing has to be kept
// from the forward call):
//   k_row_offsets / k_fill_rows   compact row index: pillar -> [row_off[p], row_off[p+1])
//   k_decorate_rows               X_0
//   per layer  l = 0..L-1 :  Y_l = X_l W_l^T ; m_l = segmented max of relu(bn(Y_l)) ; X_{l+1} = [a_l || m_l]
//                            (every X_l is kept: dW_l needs it again and rebuilding it was a full row-space pass)
//   per layer  l = L-1..0 :  k_dz (arg-max routing + ReLU mask + BN sums) ; k_bn_finalize (dgamma, dbeta) ;
//                            k_dy ; dW_l = dY^T X_l (k_gemm_dw: split over row slices, fixed-order reduce) ; dX = dY W_l
// Every reduction has a fixed order => run-to-run identical gradients. Products: dX (always) and the eval-mode recompute
// run as 3xTF32 tcgen05 GEMMs (rows_gemm_tc.cuh); the train-mode recompute and dW stay on the fp32 FMA pipe (1e-5 parity
// of gradients that BatchNorm's backward amplifies). The element-wise passes handle four units per thread.

#include <algorithm>

#include "common.cuh"
#include "rows_gemm_tc.cuh"

namespace mbev {
namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------------
// row space
// ---------------------------------------------------------------------------------------------------
// Compact row index: row_off[p] = sum over pillars before p of (n + [n < T]). Two launches of 256-thread CTAs that own
// 2048 consecutive pillars each (8 per thread): per-CTA totals, then every CTA re-reduces the totals before it and scans
// its own pillars. (The single-CTA scan took 86 us for 88 k pillars and 0.3 ms for the 353 k of a 16-frame batch.)
constexpr int kRowOffPer = 8;
constexpr int kRowOffTile = 256 * kRowOffPer;

__device__ __forceinline__ int row_count(const int *__restrict__ num_points, const int p, const int P, const int T) {
  if (p >= P) return 0;
  const int n = num_points[p];
  return n + (n < T ? 1 : 0);
}

__global__ void __launch_bounds__(256)
k_row_part(const int *__restrict__ num_points, const int *__restrict__ num_pillars, const int T, int *__restrict__ tot) {
  __shared__ int s_w[8];
  const int P = *num_pillars;
  const int p0 = blockIdx.x * kRowOffTile + threadIdx.x * kRowOffPer;
  if (blockIdx.x * kRowOffTile >= P) {  // grid sized by the capacity: nothing here
    if (threadIdx.x == 0) tot[blockIdx.x] = 0;
    return;
  }
  int sum = 0;
#pragma unroll
  for (int j = 0; j < kRowOffPer; ++j) sum += row_count(num_points, p0 + j, P, T);
  sum = __reduce_add_sync(0xffffffffu, sum);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_w[i];
    tot[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256)
k_row_offsets(const int *__restrict__ num_points, const int *__restrict__ num_pillars, const int T,
              const int *__restrict__ tot, int *__restrict__ row_off, int *__restrict__ num_rows) {
  __shared__ int s_w[8], s_b[8];
  const int P = *num_pillars;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int first = blockIdx.x * kRowOffTile;
  if (first > P) return;  // (first == P: the CTA that writes the total when P is a multiple of the tile)
  int base = 0;
  for (int j = tid; j < static_cast<int>(blockIdx.x); j += 256) base += __ldg(tot + j);
  base = __reduce_add_sync(0xffffffffu, base);
  const int p0 = first + tid * kRowOffPer;
  int v[kRowOffPer], sum = 0;
#pragma unroll
  for (int j = 0; j < kRowOffPer; ++j) {
    v[j] = row_count(num_points, p0 + j, P, T);
    sum += v[j];
  }
  int x = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += u;
  }
  if (lane == 31) s_w[warp] = x;
  if (lane == 0) s_b[warp] = base;
  __syncthreads();
  int excl = x - sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    excl += s_b[i];
    if (i < warp) excl += s_w[i];
  }
#pragma unroll
  for (int j = 0; j < kRowOffPer; ++j) {
    if (p0 + j < P) row_off[p0 + j] = excl;
    if (p0 + j == P) {  // one past the last pillar: the total
      row_off[P] = excl;
      *num_rows = excl;  // R, at a host-known address
    }
    excl += v[j];
  }
}

// one thread per pillar: row tables, cluster mean, decorated rows X0 (R, D0)
struct DecoK {
  int C, D0, T, cluster, vcenter, dist, legacy, vcd;
  float vx, vy, vz, xo, yo, zo;
};

__global__ void __launch_bounds__(kThreads)
k_decorate_rows(const float *__restrict__ rows_src, const int *__restrict__ kept_idx, const int *__restrict__ num_points,
                const int *__restrict__ coors, const int *__restrict__ num_pillars, const int *__restrict__ row_off,
                const __grid_constant__ DecoK k, int *__restrict__ row_pillar, float *__restrict__ row_w,
                float *__restrict__ x0) {
  const int P = *num_pillars;
  for (int p = blockIdx.x * kThreads + threadIdx.x; p < P; p += gridDim.x * kThreads) {
    const int n = num_points[p];
    const int r0 = row_off[p];
    const int4 c = reinterpret_cast<const int4 *>(coors)[p];
    const float cx = __fadd_rn(__fmul_rn(static_cast<float>(c.w), k.vx), k.xo);
    const float cy = __fadd_rn(__fmul_rn(static_cast<float>(c.z), k.vy), k.yo);
    const float cz = __fadd_rn(__fmul_rn(static_cast<float>(c.y), k.vz), k.zo);
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int t = 0; t < n; ++t) {
      const size_t slot = static_cast<size_t>(p) * k.T + t;
      const float *q = rows_src + static_cast<size_t>(kept_idx ? kept_idx[slot] : static_cast<int>(slot)) * k.C;
      sx = __fadd_rn(sx, q[0]);
      sy = __fadd_rn(sy, q[1]);
      sz = __fadd_rn(sz, q[2]);
    }
    const float fn = static_cast<float>(n);
    const float mx = __fdiv_rn(sx, fn), my = __fdiv_rn(sy, fn), mz = __fdiv_rn(sz, fn);
    for (int t = 0; t < n; ++t) {
      const size_t slot = static_cast<size_t>(p) * k.T + t;
      const float *q = rows_src + static_cast<size_t>(kept_idx ? kept_idx[slot] : static_cast<int>(slot)) * k.C;
      const float x = q[0], y = q[1], z = q[2];
      const float ex = __fsub_rn(x, cx), ey = __fsub_rn(y, cy), ez = __fsub_rn(z, cz);
      const bool alias = k.vcenter && k.legacy;
      const float a0 = alias ? ex : x, a1 = alias ? ey : y, a2 = (alias && k.vcd > 2) ? ez : z;
      float *o = x0 + static_cast<size_t>(r0 + t) * k.D0;
      int d = 0;
      o[d++] = a0; o[d++] = a1; o[d++] = a2;
      for (int cc = 3; cc < k.C; ++cc) o[d++] = q[cc];
      if (k.cluster) { o[d++] = __fsub_rn(x, mx); o[d++] = __fsub_rn(y, my); o[d++] = __fsub_rn(z, mz); }
      if (k.vcenter) { o[d++] = ex; o[d++] = ey; if (k.vcd > 2) o[d++] = ez; }
      if (k.dist) o[d++] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)), __fmul_rn(a2, a2)));
      row_pillar[r0 + t] = p;
      row_w[r0 + t] = 1.f;
    }
    if (n < k.T) {
      float *o = x0 + static_cast<size_t>(r0 + n) * k.D0;
      for (int d = 0; d < k.D0; ++d) o[d] = 0.f;
      row_pillar[r0 + n] = p;
      row_w[r0 + n] = static_cast<float>(k.T - n);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// generic fp32 tiled GEMM: C(m,n) = sum_k A(m,k) * B(k,n), element strides; 64x64 tile, BK 16, 4x4 per thread.
// M may come from device memory (row count). Optional split-K over gridDim.z into Cpart[z].
// ---------------------------------------------------------------------------------------------------
struct GemmK {
  const float *A, *B;
  float *C;
  long long sAm, sAk, sBk, sBn, ldc;
  int M, N, K;        // static extents; if m_dev / k_dev set they override M / K
  const int *m_dev, *k_dev;
  long long split_stride;  // elements between split-K partial outputs
};

__global__ void __launch_bounds__(256)
k_gemm(const __grid_constant__ GemmK g) {
  __shared__ float sA[16][64 + 4];
  __shared__ float sB[16][64 + 4];
  const int M = g.m_dev ? *g.m_dev : g.M;
  const int K = g.k_dev ? *g.k_dev : g.K;
  const int N = g.N;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int nsplit = gridDim.z;
  const int kper = ((K + nsplit - 1) / nsplit + 15) / 16 * 16;
  const int k0 = blockIdx.z * kper, k1 = min(K, k0 + kper);
  float *C = g.C + blockIdx.z * g.split_stride;
  for (int mt = blockIdx.x; mt * 64 < M; mt += gridDim.x) {
    const int m0 = mt * 64, n0 = blockIdx.y * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int kb = k0; kb < k1; kb += 16) {
      // load A tile (64 x 16) and B tile (16 x 64); fastest-varying thread index follows the unit stride
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = tid + it * 256;
        int mm, kk;
        if (g.sAk == 1) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
        const int m = m0 + mm, k = kb + kk;
        sA[kk][mm] = (m < M && k < k1) ? __ldg(g.A + m * g.sAm + k * g.sAk) : 0.f;
        int nn, kb2;
        if (g.sBn == 1) { nn = idx & 63; kb2 = idx >> 6; } else { kb2 = idx & 15; nn = idx >> 4; }
        const int n = n0 + nn, k2 = kb + kb2;
        sB[kb2][nn] = (n < N && k2 < k1) ? __ldg(g.B + k2 * g.sBk + n * g.sBn) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const float4 a = *reinterpret_cast<const float4 *>(&sA[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4 *>(&sB[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n < N) C[m * g.ldc + n] = acc[i][j];
      }
    }
  }
}

// Row-streaming form for the two GEMMs whose M is the compact row count (forward recompute Y = X W^T, dX = dY W):
// 128 x BN tile, BK 16, 256 threads, 8 x (BN/16) outputs per thread (16 FMAs per shared-memory load instead of 8),
// the next k-block's global loads are in flight while the current one is multiplied. A must be row-major
// (sAk == 1). Every output element still accumulates k = 0, 1, 2, ... in order with one fmaf per term, so the
// result is bit-identical to k_gemm and to the forward kernel the statistics came from.
template <int BN>
__global__ void __launch_bounds__(256)
k_gemm_rows(const __grid_constant__ GemmK g) {
  constexpr int BM = 128, BK = 16, TN = BN / 16;
  __shared__ float sA[BK][BM + 4];
  __shared__ float sB[BK][BN + 4];
  const int M = g.m_dev ? *g.m_dev : g.M;
  const int K = g.K, N = g.N;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.y * BN;
  float ra[8], rb[BN / 16];
  auto gload = [&](int m0, int kb) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {  // A tile 128 x 16: 16 consecutive k of one row per 16 threads
        const int idx = tid + it * 256;
        const int kk = idx & 15, mm = idx >> 4;
        const int m = m0 + mm, k = kb + kk;
        ra[it] = (m < M && k < K) ? __ldg(g.A + static_cast<long long>(m) * g.sAm + k) : 0.f;
      }
#pragma unroll
      for (int it = 0; it < BN / 16; ++it) {  // B tile 16 x BN, fastest thread index along the unit stride
        const int idx = tid + it * 256;
        int nn, kk;
        if (g.sBn == 1) { nn = idx % BN; kk = idx / BN; } else { kk = idx & 15; nn = idx >> 4; }
        const int n = n0 + nn, k = kb + kk;
        rb[it] = (n < N && k < K) ? __ldg(g.B + static_cast<long long>(k) * g.sBk + static_cast<long long>(n) * g.sBn) : 0.f;
      }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = tid + it * 256;
      sA[idx & 15][idx >> 4] = ra[it];
    }
#pragma unroll
    for (int it = 0; it < BN / 16; ++it) {
      const int idx = tid + it * 256;
      if (g.sBn == 1) sB[idx / BN][idx % BN] = rb[it]; else sB[idx & 15][idx >> 4] = rb[it];
    }
  };
  if (static_cast<int>(blockIdx.x) * BM < M) gload(blockIdx.x * BM, 0);
  for (int mt = blockIdx.x; mt * BM < M; mt += gridDim.x) {
    const int m0 = mt * BM;
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int kb = 0; kb < K; kb += BK) {
      __syncthreads();  // the previous block's readers are done with the tiles
      sstore();
      __syncthreads();
      // the next k-block's rows — or, under the last k-block, the NEXT TILE's first rows (K = 64 / 128 is only 4 - 8
      // k-blocks: no load latency exposed at the start of a tile) — are in flight during the products
      if (kb + BK < K) gload(m0, kb + BK);
      else if ((mt + static_cast<int>(gridDim.x)) * BM < M) gload((mt + gridDim.x) * BM, 0);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4 *>(&sA[kk][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4 *>(&sA[kk][ty * 8 + 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float bv[TN];  // columns tx*4 .. tx*4+3 of every 64-column half: consecutive lanes read consecutive 16 bytes
#pragma unroll
        for (int j4 = 0; j4 < TN; j4 += 4) {
          const float4 b = *reinterpret_cast<const float4 *>(&sB[kk][tx * 4 + 16 * j4]);
          bv[j4] = b.x; bv[j4 + 1] = b.y; bv[j4 + 2] = b.z; bv[j4 + 3] = b.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + ty * 8 + i;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * 4 + 16 * (j & ~3) + (j & 3);
        if (n < N) g.C[static_cast<long long>(m) * g.ldc + n] = acc[i][j];
      }
    }
  }
}

// dW slice: part[z] (U, Kin) = sum over the rows of slice z of dY[r][:]^T X[r][:] — both factors row-major with the
// reduction along rows, so a k-block is 16 whole rows of each (16-byte loads, no transposition). One CTA owns the WHOLE
// (<= 128 x 128) output: BM x 128 tile, (BM/16) x 8 outputs per thread (16 FMAs per shared-memory load against 8 in
// k_gemm's 4 x 4), next k-block's rows in flight while the current one is multiplied. Rows accumulate in order with one
// fmaf per term, as in k_gemm: same partials bit for bit.
template <int BM>
__global__ void __launch_bounds__(256)
k_gemm_dw(const float *__restrict__ DY, const float *__restrict__ X, const int U, const int Kin,
          const int *__restrict__ num_rows, float *__restrict__ part, const long long split_stride) {
  constexpr int BN = 128, BK = 16, TM = BM / 16, NA = BM / 64;
  __shared__ __align__(16) float sA[BK][BM + 4];
  __shared__ __align__(16) float sB[BK][BN + 4];
  const int R = *num_rows;
  const int nsplit = gridDim.x;
  const int kper = ((R + nsplit - 1) / nsplit + BK - 1) / BK * BK;
  const int k0 = min(R, static_cast<int>(blockIdx.x) * kper), k1 = min(R, k0 + kper);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[TM][8];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ra[NA], rb[2];
  auto gload = [&](int kb) {
#pragma unroll
    for (int it = 0; it < NA; ++it) {
      const int idx = tid + it * 256, row = idx / (BM / 4), c = (idx % (BM / 4)) * 4, r = kb + row;
      ra[it] = (r < k1 && c < U) ? __ldg(reinterpret_cast<const float4 *>(DY + static_cast<size_t>(r) * U + c)) : z4;
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + it * 256, row = idx >> 5, c = (idx & 31) * 4, r = kb + row;
      rb[it] = (r < k1 && c < Kin) ? __ldg(reinterpret_cast<const float4 *>(X + static_cast<size_t>(r) * Kin + c)) : z4;
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int it = 0; it < NA; ++it) {
      const int idx = tid + it * 256;
      *reinterpret_cast<float4 *>(&sA[idx / (BM / 4)][(idx % (BM / 4)) * 4]) = ra[it];
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + it * 256;
      *reinterpret_cast<float4 *>(&sB[idx >> 5][(idx & 31) * 4]) = rb[it];
    }
  };
  if (k0 < k1) gload(k0);
  for (int kb = k0; kb < k1; kb += BK) {
    __syncthreads();
    sstore();
    __syncthreads();
    if (kb + BK < k1) gload(kb + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[8];
#pragma unroll
      for (int i4 = 0; i4 < TM; i4 += 4) {
        const float4 a = *reinterpret_cast<const float4 *>(&sA[kk][ty * TM + i4]);
        av[i4] = a.x; av[i4 + 1] = a.y; av[i4 + 2] = a.z; av[i4 + 3] = a.w;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 b = *reinterpret_cast<const float4 *>(&sB[kk][tx * 4 + 64 * h]);
        bv[4 * h] = b.x; bv[4 * h + 1] = b.y; bv[4 * h + 2] = b.z; bv[4 * h + 3] = b.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  float *C = part + static_cast<long long>(blockIdx.x) * split_stride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int u = ty * TM + i;
    if (u >= U) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = tx * 4 + 64 * (j >> 2) + (j & 3);
      if (n < Kin) C[static_cast<long long>(u) * Kin + n] = acc[i][j];
    }
  }
}

// dW slice of a NARROW layer (Kin <= 16: layer 0, whose input is the 9..13 decorated point features): the 64 x 64 tiles
// of k_gemm waste 5/6 of their columns there (129 us per 4-frame step for 0.4 GFLOP). One thread per (four consecutive
// units, input column k): U / 4 * Kin <= 512 threads; rows through shared memory in tiles of 32 (16-byte loads of whole
// dY rows, all of a tile in flight at once); per row a thread issues one 16-byte and one 4-byte shared-memory load and
// four fmaf, rows in order (same partials as k_gemm). Earlier forms: thread = (u, column group) with 8 predicated
// accumulators compiled to ~80 instructions per row and warp (196 - 321 us); one thread per output element was bound by
// the shared-memory pipe (two wavefronts per 32 products, 93 us). Needs U % 4 == 0, U <= 128, Kin <= 16.
__global__ void __launch_bounds__(512)
k_dw_narrow(const float *__restrict__ DY, const float *__restrict__ X, const int U, const int Kin,
            const int *__restrict__ num_rows, float *__restrict__ part, const long long split_stride) {
  constexpr int TR = 32;
  __shared__ __align__(16) float sDY[TR * 128];
  __shared__ float sX[TR * 16];
  const int R = *num_rows;
  const int nsplit = gridDim.x;
  const int kper = ((R + nsplit - 1) / nsplit + TR - 1) / TR * TR;
  const int k0 = min(R, static_cast<int>(blockIdx.x) * kper), k1 = min(R, k0 + kper);
  const int tid = threadIdx.x, nthr = blockDim.x, U4 = U >> 2;
  const bool has = tid < U4 * Kin;
  const int k = has ? tid / U4 : 0, ug = has ? tid % U4 : 0;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int kb = k0; kb < k1; kb += TR) {
    const int rows = min(TR, k1 - kb);
    const float4 *src = reinterpret_cast<const float4 *>(DY + static_cast<size_t>(kb) * U);
    for (int i = tid; i < rows * U4; i += nthr) reinterpret_cast<float4 *>(sDY)[i] = __ldg(src + i);
    const float *xs = X + static_cast<size_t>(kb) * Kin;
    for (int i = tid; i < rows * Kin; i += nthr) sX[i] = __ldg(xs + i);
    __syncthreads();
    if (has) {
#pragma unroll 8
      for (int r = 0; r < rows; ++r) {
        const float4 dy = reinterpret_cast<const float4 *>(sDY)[r * U4 + ug];
        const float x = sX[r * Kin + k];
        acc.x = fmaf(dy.x, x, acc.x);
        acc.y = fmaf(dy.y, x, acc.y);
        acc.z = fmaf(dy.z, x, acc.z);
        acc.w = fmaf(dy.w, x, acc.w);
      }
    }
    __syncthreads();
  }
  if (has) {
    float *C = part + static_cast<long long>(blockIdx.x) * split_stride + static_cast<long long>(4 * ug) * Kin + k;
    C[0] = acc.x;
    C[Kin] = acc.y;
    C[2 * Kin] = acc.z;
    C[3 * Kin] = acc.w;
  }
}

// fixed-order reduction of split-K partials into dst (count elements): one WARP per element — lane j sums partials j,
// j + 32, ... in that order, then a fixed xor-shuffle tree (deterministic; one thread per element walking all partials was
// a chain of `nsplit` L2 round trips)
__global__ void __launch_bounds__(256)
k_reduce_splits(const float *__restrict__ part, const int nsplit, const long long stride, const int count,
                float *__restrict__ dst) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= count) return;
  float s = 0.f;
#pragma unroll 4
  for (int z = lane; z < nsplit; z += 32) s += __ldg(part + z * stride + i);
#pragma unroll
  for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0) dst[i] = s;
}

// ---------------------------------------------------------------------------------------------------
// forward recompute pieces
// ---------------------------------------------------------------------------------------------------
// one thread per (pillar, unit): m[p][u] = max over the pillar's rows of relu(y*scale+shift)
__global__ void __launch_bounds__(kThreads)
k_act_max(const float *__restrict__ Y, const int U, const float *__restrict__ scale, const float *__restrict__ shift,
          const int *__restrict__ row_off, const int *__restrict__ num_pillars, float *__restrict__ Mx) {
  const long long total = static_cast<long long>(*num_pillars) * U;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int p = static_cast<int>(i / U), u = static_cast<int>(i - static_cast<long long>(p) * U);
    const float sc = scale[u], sh = shift[u];
    float m = 0.f;
    for (int r = row_off[p]; r < row_off[p + 1]; ++r) m = fmaxf(m, fmaf(Y[static_cast<size_t>(r) * U + u], sc, sh));
    Mx[i] = m;
  }
}

// The same with four units per thread (U % 4 == 0): 16-byte accesses, four times the bytes in flight per thread — these
// row-space passes are latency-bound, not bandwidth-bound (the scalar forms moved 1.8 - 2.5 TB/s). Same arithmetic per
// element, so the results are bit-identical to the scalar kernels.
// arg (optional): offset inside the pillar of the FIRST row attaining the max (torch.max routes the gradient there;
// all-non-positive columns have m = 0 and route to row 0, whose ReLU mask then zeroes the gradient anyway).
__global__ void __launch_bounds__(kThreads)
k_act_max4(const float *__restrict__ Y, const int U, const float *__restrict__ scale, const float *__restrict__ shift,
           const int *__restrict__ row_off, const int *__restrict__ num_pillars, float *__restrict__ Mx,
           uchar4 *__restrict__ arg) {
  const int U4 = U >> 2;
  const long long total = static_cast<long long>(*num_pillars) * U4;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int p = static_cast<int>(i / U4), q = static_cast<int>(i - static_cast<long long>(p) * U4);
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale) + q);
    const float4 sh = __ldg(reinterpret_cast<const float4 *>(shift) + q);
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    uchar4 a = make_uchar4(0, 0, 0, 0);
    const int r0 = row_off[p], r1 = row_off[p + 1];
    for (int r = r0; r < r1; ++r) {
      const float4 y = __ldg(reinterpret_cast<const float4 *>(Y + static_cast<size_t>(r) * U) + q);
      const float zx = fmaf(y.x, sc.x, sh.x), zy = fmaf(y.y, sc.y, sh.y), zz = fmaf(y.z, sc.z, sh.z),
                  zw = fmaf(y.w, sc.w, sh.w);
      const unsigned char t = static_cast<unsigned char>(r - r0);
      if (zx > m.x) { m.x = zx; a.x = t; }
      if (zy > m.y) { m.y = zy; a.y = t; }
      if (zz > m.z) { m.z = zz; a.z = t; }
      if (zw > m.w) { m.w = zw; a.w = t; }
    }
    reinterpret_cast<float4 *>(Mx)[i] = m;
    if (arg) arg[i] = a;
  }
}

// k_act_max4 and the row pass X_{l+1} = [a_l || m_l] (k_build_x) in one pillar pass (layers that feed another layer): the second walk over the pillar's rows
// re-reads Y from L1 / L2 and writes X_{l+1}[r] = [ relu(bn(Y_l[r])) || m ] — one DRAM read of Y and one launch less.
__global__ void __launch_bounds__(kThreads)
k_act_max_build4(const float *__restrict__ Y, const int U, const float *__restrict__ scale,
                 const float *__restrict__ shift, const int *__restrict__ row_off, const int *__restrict__ num_pillars,
                 float *__restrict__ Mx, uchar4 *__restrict__ arg, float *__restrict__ X) {
  const int U4 = U >> 2;
  const long long total = static_cast<long long>(*num_pillars) * U4;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int p = static_cast<int>(i / U4), q = static_cast<int>(i - static_cast<long long>(p) * U4);
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale) + q);
    const float4 sh = __ldg(reinterpret_cast<const float4 *>(shift) + q);
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    uchar4 a = make_uchar4(0, 0, 0, 0);
    const int r0 = row_off[p], r1 = row_off[p + 1];
    for (int r = r0; r < r1; ++r) {
      const float4 y = __ldg(reinterpret_cast<const float4 *>(Y + static_cast<size_t>(r) * U) + q);
      const float zx = fmaf(y.x, sc.x, sh.x), zy = fmaf(y.y, sc.y, sh.y), zz = fmaf(y.z, sc.z, sh.z),
                  zw = fmaf(y.w, sc.w, sh.w);
      const unsigned char t = static_cast<unsigned char>(r - r0);
      if (zx > m.x) { m.x = zx; a.x = t; }
      if (zy > m.y) { m.y = zy; a.y = t; }
      if (zz > m.z) { m.z = zz; a.z = t; }
      if (zw > m.w) { m.w = zw; a.w = t; }
    }
    reinterpret_cast<float4 *>(Mx)[i] = m;
    arg[i] = a;
    for (int r = r0; r < r1; ++r) {
      const float4 y = __ldg(reinterpret_cast<const float4 *>(Y + static_cast<size_t>(r) * U) + q);
      float4 *x = reinterpret_cast<float4 *>(X + static_cast<size_t>(r) * 2 * U);
      x[q] = make_float4(fmaxf(fmaf(y.x, sc.x, sh.x), 0.f), fmaxf(fmaf(y.y, sc.y, sh.y), 0.f),
                         fmaxf(fmaf(y.z, sc.z, sh.z), 0.f), fmaxf(fmaf(y.w, sc.w, sh.w), 0.f));
      x[U4 + q] = m;
    }
  }
}

// X_{l+1}[r] = [ relu(bn(Y_l[r])) || m_l[pillar(r)] ]   (R, 2U)
__global__ void __launch_bounds__(kThreads)
k_build_x(const float *__restrict__ Y, const int U, const float *__restrict__ scale, const float *__restrict__ shift,
          const float *__restrict__ Mx, const int *__restrict__ row_pillar, const int *__restrict__ num_rows,
          float *__restrict__ X) {
  const long long total = static_cast<long long>(*num_rows) * 2 * U;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int r = static_cast<int>(i / (2 * U)), c = static_cast<int>(i - static_cast<long long>(r) * 2 * U);
    X[i] = (c < U) ? fmaxf(fmaf(Y[static_cast<size_t>(r) * U + c], scale[c], shift[c]), 0.f)
                   : Mx[static_cast<size_t>(row_pillar[r]) * U + (c - U)];
  }
}

// Train mode: the batch statistics of layer l, recomputed from the Y_l THIS pass computed (sum over all P*T slots =
// sum over compact rows of row_w * y). The forward may have run on the tensor cores (3xTF32) while this recompute runs
// on the FMA pipe: the two Y agree to ~1e-6, but BatchNorm's backward subtracts sums that cancel, so normalising the
// FMA rows with the tensor-core pass's mean / variance put train-mode gradients at 2x torch's own fp32 error. With
// the statistics taken from the same rows the backward differentiates is self-consistent again. Block = (U, NL): lane
// y walks rows ra + y, ra + y + NL, ...; fp64, fixed order.
constexpr int kStatBlocks = 512;

__global__ void __launch_bounds__(1024)
k_row_stats(const float *__restrict__ Y, const int U, const float *__restrict__ row_w, const int *__restrict__ num_rows,
            double *__restrict__ partials) {
  extern __shared__ double s_sum[];  // [NL][2][U]
  const int u = threadIdx.x, yl = threadIdx.y, NL = blockDim.y;
  const int R = *num_rows;
  const int per = (R + gridDim.x - 1) / gridDim.x;
  const int ra = min(R, blockIdx.x * per), rb = min(R, ra + per);
  double s1 = 0.0, s2 = 0.0;
  for (int r = ra + yl; r < rb; r += NL) {
    const double y = static_cast<double>(Y[static_cast<size_t>(r) * U + u]);
    const double wy = static_cast<double>(row_w[r]) * y;
    s1 += wy;
    s2 += wy * y;
  }
  s_sum[(yl * 2 + 0) * U + u] = s1;
  s_sum[(yl * 2 + 1) * U + u] = s2;
  __syncthreads();
  if (yl == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int j = 0; j < NL; ++j) {
      t1 += s_sum[(j * 2 + 0) * U + u];
      t2 += s_sum[(j * 2 + 1) * U + u];
    }
    partials[(static_cast<size_t>(blockIdx.x) * 2 + 0) * U + u] = t1;
    partials[(static_cast<size_t>(blockIdx.x) * 2 + 1) * U + u] = t2;
  }
}

// one warp per unit (fixed-order lane-strided sums + shuffle tree). gamma / beta are recovered from the folded
// scale / shift and the statistics the forward used: gamma = scale * sqrt(var + eps), beta = shift + mean * scale.
__global__ void __launch_bounds__(256)
k_row_stats_finalize(const double *__restrict__ partials, const int nblocks, const int U,
                     const int *__restrict__ num_pillars, const int T, const float eps,
                     const float *__restrict__ scale_fwd, const float *__restrict__ shift_fwd,
                     const float *__restrict__ mean_fwd, const float *__restrict__ var_fwd, float *__restrict__ scale,
                     float *__restrict__ shift, float *__restrict__ mean, float *__restrict__ var) {
  const int u = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (u >= U) return;
  double s1 = 0.0, s2 = 0.0;
  for (int b = lane; b < nblocks; b += 32) {
    s1 += partials[(static_cast<size_t>(b) * 2 + 0) * U + u];
    s2 += partials[(static_cast<size_t>(b) * 2 + 1) * U + u];
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
  if (lane) return;
  const double M = static_cast<double>(*num_pillars) * T;
  const double mu = M > 0 ? s1 / M : 0.0;
  double v = M > 0 ? s2 / M - mu * mu : 0.0;
  if (v < 0) v = 0;
  const double gamma = static_cast<double>(scale_fwd[u]) * sqrt(static_cast<double>(var_fwd[u]) + static_cast<double>(eps));
  const double beta = static_cast<double>(shift_fwd[u]) + static_cast<double>(mean_fwd[u]) * static_cast<double>(scale_fwd[u]);
  const double sc = gamma / sqrt(v + static_cast<double>(eps));
  scale[u] = static_cast<float>(sc);
  shift[u] = static_cast<float>(beta - mu * sc);
  mean[u] = static_cast<float>(mu);
  var[u] = static_cast<float>(v);
}

// ---------------------------------------------------------------------------------------------------
// backward pieces
// ---------------------------------------------------------------------------------------------------
// Block = (U, NL) threads: threadIdx.x = unit, threadIdx.y = pillar lane. The block owns a contiguous range of
// pillars; lane y walks pillars pa + y, pa + y + NL, ... (NL independent load chains per unit instead of one — the
// single-lane form of this kernel was latency-bound at 7.6 ms per layer on kitti_b16). For each (pillar, unit):
//   dm = dfeats (last layer) or sum over the pillar's rows of dXnext[r][U + u]   (gradient of the broadcast max)
//   route dm to the FIRST row attaining the max (torch.max semantics; real rows precede the virtual row),
//   add the per-row gradient dXnext[r][u], apply the ReLU mask -> dz; accumulate S1 = sum dz, S2 = sum dz*xhat.
// The lanes' sums are combined in lane order through shared memory: fixed order, run-to-run identical.
constexpr int kDzThreads = 1024;

__global__ void __launch_bounds__(kDzThreads)
k_dz(const float *__restrict__ Y, const int U, const float *__restrict__ scale, const float *__restrict__ shift,
     const float *__restrict__ mean, const float *__restrict__ var, const float eps, const float *__restrict__ Mx,
     const float *__restrict__ dfeats, const float *__restrict__ dXnext, const int ldx,
     const int *__restrict__ row_off, const int *__restrict__ num_pillars, float *__restrict__ DZ,
     double *__restrict__ partials) {
  extern __shared__ double s_sum[];  // [NL][2][U]
  const int u = threadIdx.x, yl = threadIdx.y, NL = blockDim.y;
  const int P = *num_pillars;
  // the split follows the ACTUAL pillar count (device side), not the capacity: every block has work
  const int pillars_per_block = (P + gridDim.x - 1) / gridDim.x;
  const int pa = min(P, blockIdx.x * pillars_per_block), pb = min(P, pa + pillars_per_block);
  const float sc = scale[u], sh = shift[u];
  const float mu = mean[u], rstd = rsqrtf(var[u] + eps);
  double s1 = 0.0, s2 = 0.0;
  for (int p = pa + yl; p < pb; p += NL) {
    const int r0 = row_off[p], r1 = row_off[p + 1];
    float dm;
    if (dXnext == nullptr) {
      dm = dfeats[static_cast<size_t>(p) * U + u];
    } else {
      dm = 0.f;
      for (int r = r0; r < r1; ++r) dm += dXnext[static_cast<size_t>(r) * ldx + U + u];
    }
    const float m = Mx[static_cast<size_t>(p) * U + u];
    bool routed = false;
    for (int r = r0; r < r1; ++r) {
      const float y = Y[static_cast<size_t>(r) * U + u];
      const float z = fmaf(y, sc, sh);
      const float a = fmaxf(z, 0.f);
      float dA = dXnext ? dXnext[static_cast<size_t>(r) * ldx + u] : 0.f;
      if (!routed && a == m) {
        dA += dm;
        routed = true;
      }
      const float dz = z > 0.f ? dA : 0.f;
      DZ[static_cast<size_t>(r) * U + u] = dz;
      s1 += static_cast<double>(dz);
      s2 += static_cast<double>(dz) * static_cast<double>((y - mu) * rstd);
    }
  }
  s_sum[(yl * 2 + 0) * U + u] = s1;
  s_sum[(yl * 2 + 1) * U + u] = s2;
  __syncthreads();
  if (yl == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int j = 0; j < NL; ++j) {
      t1 += s_sum[(j * 2 + 0) * U + u];
      t2 += s_sum[(j * 2 + 1) * U + u];
    }
    partials[(static_cast<size_t>(blockIdx.x) * 2 + 0) * U + u] = t1;
    partials[(static_cast<size_t>(blockIdx.x) * 2 + 1) * U + u] = t2;
  }
}

// Four units per thread (U % 4 == 0): block = (U / 4, NL), 512 threads. Same per-element arithmetic and the same
// first-row arg-max routing as k_dz; the fp64 lane sums are folded in lane order (fixed order, run-to-run identical).
constexpr int kDz4Threads = 512;

__global__ void __launch_bounds__(kDz4Threads)
k_dz4(const float *__restrict__ Y, const int U, const float *__restrict__ scale, const float *__restrict__ shift,
      const float *__restrict__ mean, const float *__restrict__ var, const float eps, const float *__restrict__ Mx,
      const float *__restrict__ dfeats, const float *__restrict__ dXnext, const int ldx,
      const int *__restrict__ row_off, const int *__restrict__ num_pillars, float *__restrict__ DZ,
      double *__restrict__ partials) {
  extern __shared__ double s_sum[];  // [NL][2][U]
  const int q = threadIdx.x, yl = threadIdx.y, NL = blockDim.y;
  const int P = *num_pillars;
  const int pillars_per_block = (P + gridDim.x - 1) / gridDim.x;
  const int pa = min(P, blockIdx.x * pillars_per_block), pb = min(P, pa + pillars_per_block);
  const float4 sc4 = __ldg(reinterpret_cast<const float4 *>(scale) + q);
  const float4 sh4 = __ldg(reinterpret_cast<const float4 *>(shift) + q);
  const float4 mu4 = __ldg(reinterpret_cast<const float4 *>(mean) + q);
  const float4 va4 = __ldg(reinterpret_cast<const float4 *>(var) + q);
  const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
  const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w};
  const float rstd[4] = {rsqrtf(va4.x + eps), rsqrtf(va4.y + eps), rsqrtf(va4.z + eps), rsqrtf(va4.w + eps)};
  double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
  for (int p = pa + yl; p < pb; p += NL) {
    const int r0 = row_off[p], r1 = row_off[p + 1];
    float dm[4];
    if (dXnext == nullptr) {
      const float4 d = __ldg(reinterpret_cast<const float4 *>(dfeats + static_cast<size_t>(p) * U) + q);
      dm[0] = d.x; dm[1] = d.y; dm[2] = d.z; dm[3] = d.w;
    } else {
      dm[0] = dm[1] = dm[2] = dm[3] = 0.f;
      for (int r = r0; r < r1; ++r) {
        const float4 d = __ldg(reinterpret_cast<const float4 *>(dXnext + static_cast<size_t>(r) * ldx + U) + q);
        dm[0] += d.x; dm[1] += d.y; dm[2] += d.z; dm[3] += d.w;
      }
    }
    const float4 m4 = __ldg(reinterpret_cast<const float4 *>(Mx + static_cast<size_t>(p) * U) + q);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w};
    bool routed[4] = {false, false, false, false};
    for (int r = r0; r < r1; ++r) {
      const float4 y4 = __ldg(reinterpret_cast<const float4 *>(Y + static_cast<size_t>(r) * U) + q);
      float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dXnext) a4 = __ldg(reinterpret_cast<const float4 *>(dXnext + static_cast<size_t>(r) * ldx) + q);
      const float y[4] = {y4.x, y4.y, y4.z, y4.w};
      float dA[4] = {a4.x, a4.y, a4.z, a4.w}, dz[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float z = fmaf(y[j], sc[j], sh[j]);
        const float a = fmaxf(z, 0.f);
        if (!routed[j] && a == m[j]) {
          dA[j] += dm[j];
          routed[j] = true;
        }
        dz[j] = z > 0.f ? dA[j] : 0.f;
        s1[j] += static_cast<double>(dz[j]);
        s2[j] += static_cast<double>(dz[j]) * static_cast<double>((y[j] - mu[j]) * rstd[j]);
      }
      reinterpret_cast<float4 *>(DZ + static_cast<size_t>(r) * U)[q] = make_float4(dz[0], dz[1], dz[2], dz[3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s_sum[(yl * 2 + 0) * U + 4 * q + j] = s1[j];
    s_sum[(yl * 2 + 1) * U + 4 * q + j] = s2[j];
  }
  __syncthreads();
  const int t = yl * blockDim.x + q;  // the first 2U threads fold one (sum, unit) column each
  if (t < 2 * U) {
    const int which = t / U, u = t - which * U;
    double acc = 0.0;
    for (int j = 0; j < NL; ++j) acc += s_sum[(j * 2 + which) * U + u];
    partials[(static_cast<size_t>(blockIdx.x) * 2 + which) * U + u] = acc;
  }
}

// k_dz4 split in two, so that no thread walks a pillar's rows twice behind dependent loads (k_dz4 stayed at the scalar
// kernel's 174 us per layer of a 4-frame step: latency, not bytes):
//   k_dm4      : dm[p][u] = sum over the pillar's rows of dXnext[r][U + u]  (gradient of the broadcast max; row order)
//   k_dz_rows4 : one row per thread and iteration — dz = relu'(z) * (dXnext[r][u] + [r is the arg-max row] * dm[p][u]),
//                BN sums in fp64 per thread, folded per block in lane order (fixed order, run-to-run identical).
// The arg-max row comes from the forward (k_act_max4's `arg`), which is the first row attaining the max, as k_dz routes.
__global__ void __launch_bounds__(kThreads)
k_dm4(const float *__restrict__ dXnext, const int ldx, const int U, const int *__restrict__ row_off,
      const int *__restrict__ num_pillars, float *__restrict__ DM) {
  const int U4 = U >> 2;
  const long long total = static_cast<long long>(*num_pillars) * U4;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int p = static_cast<int>(i / U4), q = static_cast<int>(i - static_cast<long long>(p) * U4);
    const int r0 = row_off[p], r1 = row_off[p + 1];
    float4 dm = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = r0; r < r1; ++r) {
      const float4 d = __ldg(reinterpret_cast<const float4 *>(dXnext + static_cast<size_t>(r) * ldx + U) + q);
      dm.x += d.x; dm.y += d.y; dm.z += d.z; dm.w += d.w;
    }
    reinterpret_cast<float4 *>(DM)[i] = dm;
  }
}

__global__ void __launch_bounds__(kDz4Threads, 2)  // <= 64 registers: two 512-thread CTAs per SM (83 registers left one)
k_dz_rows4(const float *__restrict__ Y, const int U, const float *__restrict__ scale, const float *__restrict__ shift,
           const float *__restrict__ mean, const float *__restrict__ var, const float eps,
           const uchar4 *__restrict__ arg, const float *__restrict__ dm_src, const float *__restrict__ dXnext,
           const int ldx, const int *__restrict__ row_off, const int *__restrict__ row_pillar,
           const int *__restrict__ num_rows, float *__restrict__ DZ, double *__restrict__ partials) {
  extern __shared__ double s_sum[];  // [NL][2][U]
  const int q = threadIdx.x, yl = threadIdx.y, NL = blockDim.y, U4 = U >> 2;
  const int R = *num_rows;
  const int per = (R + gridDim.x - 1) / gridDim.x;
  const int ra = min(R, static_cast<int>(blockIdx.x) * per), rb = min(R, ra + per);
  const float4 sc4 = __ldg(reinterpret_cast<const float4 *>(scale) + q);
  const float4 sh4 = __ldg(reinterpret_cast<const float4 *>(shift) + q);
  const float4 mu4 = __ldg(reinterpret_cast<const float4 *>(mean) + q);
  const float4 va4 = __ldg(reinterpret_cast<const float4 *>(var) + q);
  const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
  const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w};
  const float rstd[4] = {rsqrtf(va4.x + eps), rsqrtf(va4.y + eps), rsqrtf(va4.z + eps), rsqrtf(va4.w + eps)};
  double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 2
  for (int r = ra + yl; r < rb; r += NL) {
    const int p = __ldg(row_pillar + r);
    const int local = r - __ldg(row_off + p);
    const float4 y4 = __ldg(reinterpret_cast<const float4 *>(Y + static_cast<size_t>(r) * U) + q);
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dXnext) a4 = __ldg(reinterpret_cast<const float4 *>(dXnext + static_cast<size_t>(r) * ldx) + q);
    const float4 dm4 = __ldg(reinterpret_cast<const float4 *>(dm_src + static_cast<size_t>(p) * U) + q);
    const uchar4 g4 = __ldg(arg + static_cast<size_t>(p) * U4 + q);
    const float y[4] = {y4.x, y4.y, y4.z, y4.w}, dm[4] = {dm4.x, dm4.y, dm4.z, dm4.w};
    const int g[4] = {g4.x, g4.y, g4.z, g4.w};
    float dA[4] = {a4.x, a4.y, a4.z, a4.w}, dz[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float z = fmaf(y[j], sc[j], sh[j]);
      if (local == g[j]) dA[j] += dm[j];
      dz[j] = z > 0.f ? dA[j] : 0.f;
      s1[j] += static_cast<double>(dz[j]);
      s2[j] += static_cast<double>(dz[j]) * static_cast<double>((y[j] - mu[j]) * rstd[j]);
    }
    reinterpret_cast<float4 *>(DZ + static_cast<size_t>(r) * U)[q] = make_float4(dz[0], dz[1], dz[2], dz[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s_sum[(yl * 2 + 0) * U + 4 * q + j] = s1[j];
    s_sum[(yl * 2 + 1) * U + 4 * q + j] = s2[j];
  }
  __syncthreads();
  const int t = yl * blockDim.x + q;  // the first 2U threads fold one (sum, unit) column each
  if (t < 2 * U) {
    const int which = t / U, u = t - which * U;
    double acc = 0.0;
    for (int j = 0; j < NL; ++j) acc += s_sum[(j * 2 + which) * U + u];
    partials[(static_cast<size_t>(blockIdx.x) * 2 + which) * U + u] = acc;
  }
}

// One WARP per unit: lane j sums partials j, j+32, ... in that order, then a fixed xor-shuffle tree folds the 32 lane
// sums — deterministic, and ~32x shorter than one thread walking all kDzBlocks partials (the serial form was the top
// user kernel of a training step: ~99 us per launch).
__global__ void __launch_bounds__(256)
k_bn_finalize(const double *__restrict__ partials, const int nblocks, const int U, const int *__restrict__ num_pillars,
              const int T, const int train, float *__restrict__ dgamma, float *__restrict__ dbeta,
              float *__restrict__ c12) {
  const int u = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (u >= U) return;
  double s1 = 0.0, s2 = 0.0;
  for (int b = lane; b < nblocks; b += 32) {
    s1 += partials[(static_cast<size_t>(b) * 2 + 0) * U + u];
    s2 += partials[(static_cast<size_t>(b) * 2 + 1) * U + u];
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
  if (lane) return;
  dbeta[u] = static_cast<float>(s1);
  dgamma[u] = static_cast<float>(s2);
  const double M = static_cast<double>(*num_pillars) * T;
  c12[u] = (train && M > 0) ? static_cast<float>(s1 / M) : 0.f;
  c12[U + u] = (train && M > 0) ? static_cast<float>(s2 / M) : 0.f;
}

// dy = scale * (dz - w_row * (c1 + xhat * c2)), in place over DZ
__global__ void __launch_bounds__(kThreads)
k_dy(const float *__restrict__ Y, const int U, const float *__restrict__ scale, const float *__restrict__ mean,
     const float *__restrict__ var, const float eps, const float *__restrict__ c12, const float *__restrict__ row_w,
     const int *__restrict__ num_rows, float *__restrict__ DZ) {
  const long long total = static_cast<long long>(*num_rows) * U;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int r = static_cast<int>(i / U), u = static_cast<int>(i - static_cast<long long>(r) * U);
    const float xhat = (Y[i] - mean[u]) * rsqrtf(var[u] + eps);
    DZ[i] = scale[u] * (DZ[i] - row_w[r] * (c12[u] + xhat * c12[U + u]));
  }
}

__global__ void __launch_bounds__(kThreads)
k_dy4(const float *__restrict__ Y, const int U, const float *__restrict__ scale, const float *__restrict__ mean,
      const float *__restrict__ var, const float eps, const float *__restrict__ c12, const float *__restrict__ row_w,
      const int *__restrict__ num_rows, float *__restrict__ DZ) {
  const int U4 = U >> 2;
  const long long total = static_cast<long long>(*num_rows) * U4;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int r = static_cast<int>(i / U4), q = static_cast<int>(i - static_cast<long long>(r) * U4);
    const float4 y = __ldg(reinterpret_cast<const float4 *>(Y) + i);
    float4 d = reinterpret_cast<const float4 *>(DZ)[i];
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale) + q);
    const float4 mu = __ldg(reinterpret_cast<const float4 *>(mean) + q);
    const float4 va = __ldg(reinterpret_cast<const float4 *>(var) + q);
    const float4 c1 = __ldg(reinterpret_cast<const float4 *>(c12) + q);
    const float4 c2 = __ldg(reinterpret_cast<const float4 *>(c12 + U) + q);
    const float w = row_w[r];
    d.x = sc.x * (d.x - w * (c1.x + ((y.x - mu.x) * rsqrtf(va.x + eps)) * c2.x));
    d.y = sc.y * (d.y - w * (c1.y + ((y.y - mu.y) * rsqrtf(va.y + eps)) * c2.y));
    d.z = sc.z * (d.z - w * (c1.z + ((y.z - mu.z) * rsqrtf(va.z + eps)) * c2.z));
    d.w = sc.w * (d.w - w * (c1.w + ((y.w - mu.w) * rsqrtf(va.w + eps)) * c2.w));
    reinterpret_cast<float4 *>(DZ)[i] = d;
  }
}

struct BwdWs {
  int *row_off, *row_pillar, *num_rows, *row_tot;
  float *row_w, *X[MBEV_MAX_LAYERS], *DZ, *DX, *Mx[MBEV_MAX_LAYERS], *Y[MBEV_MAX_LAYERS], *c12, *wpart;  // X[l]: input rows of layer l, kept for dW_l
  float *stats;  // (L, 4, MBEV_MAX_UNITS): scale, shift, mean, var recomputed by this pass (train mode)
  float *DM;                         // (cap, umax): gradient of a pillar's broadcast max (k_dm4)
  uchar4 *Arg[MBEV_MAX_LAYERS];      // (cap, U_l / 4): arg-max row offset inside the pillar (k_act_max4)
  double *partials;
  float *img_fwd[MBEV_MAX_LAYERS], *img_dx[MBEV_MAX_LAYERS];  // hi / lo TF32 images of W_l and W_l^T (tensor-core GEMMs)
  size_t bytes;
};

constexpr int kDzBlocks = 1024;
constexpr int kSplitK = 296;  // dW split-K: 2 x 2 output tiles x 296 row slices = 8 CTAs per SM

BwdWs carve_bwd(void *ws, const MbevPfnParams *p, int64_t cap, int64_t rows_cap) {
  Carver c(ws);
  BwdWs w;
  int umax = 0, inmax = 0;
  for (int l = 0; l < p->num_layers; ++l) {
    umax = std::max(umax, p->units[l]);
    inmax = std::max(inmax, p->in_dim[l]);
  }
  w.row_off = c.take<int>(static_cast<size_t>(cap) + 1);
  w.num_rows = c.take<int>(1);
  w.row_tot = c.take<int>(static_cast<size_t>(cap) / kRowOffTile + 2);
  w.row_pillar = c.take<int>(static_cast<size_t>(rows_cap));
  w.row_w = c.take<float>(static_cast<size_t>(rows_cap));
  for (int l = 0; l < p->num_layers; ++l) w.X[l] = c.take<float>(static_cast<size_t>(rows_cap) * p->in_dim[l]);
  w.DZ = c.take<float>(static_cast<size_t>(rows_cap) * umax);
  w.DX = c.take<float>(static_cast<size_t>(rows_cap) * inmax);
  for (int l = 0; l < p->num_layers; ++l) {
    w.Y[l] = c.take<float>(static_cast<size_t>(rows_cap) * p->units[l]);
    w.Mx[l] = c.take<float>(static_cast<size_t>(cap) * p->units[l]);
  }
  w.c12 = c.take<float>(2 * umax);
  w.DM = c.take<float>(static_cast<size_t>(cap) * umax);
  for (int l = 0; l < p->num_layers; ++l) w.Arg[l] = c.take<uchar4>(static_cast<size_t>(cap) * ((p->units[l] + 3) / 4));
  w.wpart = c.take<float>(static_cast<size_t>(kSplitK) * umax * inmax);
  w.partials = c.take<double>(static_cast<size_t>(kDzBlocks) * 2 * umax);
  w.stats = c.take<float>(static_cast<size_t>(MBEV_MAX_LAYERS) * 4 * MBEV_MAX_UNITS);
  for (int l = 0; l < p->num_layers; ++l) {
    w.img_fwd[l] = c.take<float>(static_cast<size_t>(2) * p->units[l] * p->in_dim[l]);
    w.img_dx[l] = c.take<float>(static_cast<size_t>(2) * p->units[l] * p->in_dim[l]);
  }
  w.bytes = c.off;
  return w;
}

// The two row-space products of a layer, Y = X W^T (K = in, N = units) and dX = dY W (K = units, N = in), run as 3xTF32
// tcgen05 GEMMs (rows_gemm_tc.cuh, the patch embedding's kernel without the gather) when the shapes fit: K in {64, 128},
// N a multiple of 32 — every layer but the first of the configured stacks. fp32 parity as in the forward (a_l * w_l
// dropped: <= 2^-22 relative); the FMA kernels remain for layer 0 (K = 10..12) and as `gemm_path = MBEV_GEMM_FMA`.
// In TRAIN mode the recompute Y = X W^T stays on the FMA kernel: BatchNorm's backward subtracts sums that cancel and
// amplifies the row error a thousandfold (torch's own fp32 gradients are 2e-3 from float64 on the [128,128,128] test
// case); tensor-core rows — also with the fourth product a_l * w_l and a rounded a_l, measured — put dgamma at 6.4e-3,
// twice the allowance, so it is the accumulation inside the tensor core, not the split, that costs the accuracy there.
bool tc_rows_ok(const MbevPfnParams *p, int K, int N) {
  if (p->gemm_path == MBEV_GEMM_FMA) return false;
  PeArgs a{};
  return (K == 64 || K == 128) && pe_gemm_plan(a, N, K);
}

int launch_gemm(const GemmK &g, int m_tiles_cap, int n, int splits, cudaStream_t stream) {
  if (splits == 1 && g.sAk == 1 && g.m_dev != nullptr && (g.sBn == 1 || g.sBk == 1)) {
    const int tiles = std::max(1, std::min((m_tiles_cap + 1) / 2, kNumSMs * 4));
    // 64-column tiles also for N = 128: the 128-column instantiation needs 165 registers = ONE 8-warp CTA per SM (343 us
    // for the 128 -> 128 layer of a 4-frame step against 322 us for two co-resident 64-column CTAs; ncu launch lists); the
    // A tile is read twice, from L2. Same k order per output element: bit-identical.
    k_gemm_rows<64><<<dim3(tiles, (n + 63) / 64), 256, 0, stream>>>(g);
    MBEV_CHECK_LAUNCH();
    return MBEV_OK;
  }
  dim3 grid(std::max(1, std::min(m_tiles_cap, kNumSMs * 8)), (n + 63) / 64, splits);
  k_gemm<<<grid, 256, 0, stream>>>(g);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}


// Statistics of the train-mode FORWARD in row space (mbev_pfn_forward_train_rows): gamma / beta are the layer's own
// parameters; mean / biased variance over all P*T slots; the folded scale / shift go to the pass's own block (what the
// backward reads) and to the caller's output blocks (running-statistics update, same contract as mbev_pfn_forward_train).
__global__ void __launch_bounds__(256)
k_row_stats_finalize_fwd(const double *__restrict__ partials, const int nblocks, const int U,
                         const int *__restrict__ num_pillars, const int T, const float eps,
                         const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ scale,
                         float *__restrict__ shift, float *__restrict__ mean, float *__restrict__ var,
                         float *__restrict__ scale_out, float *__restrict__ shift_out, float *__restrict__ mean_out,
                         float *__restrict__ var_out) {
  const int u = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (u >= U) return;
  double s1 = 0.0, s2 = 0.0;
  for (int b = lane; b < nblocks; b += 32) {
    s1 += partials[(static_cast<size_t>(b) * 2 + 0) * U + u];
    s2 += partials[(static_cast<size_t>(b) * 2 + 1) * U + u];
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
  if (lane) return;
  const double M = static_cast<double>(*num_pillars) * T;
  const double mu = M > 0 ? s1 / M : 0.0;
  double v = M > 0 ? s2 / M - mu * mu : 0.0;
  if (v < 0) v = 0;
  const double sc = static_cast<double>(gamma[u]) / sqrt(v + static_cast<double>(eps));
  const float fsc = static_cast<float>(sc), fsh = static_cast<float>(static_cast<double>(beta[u]) - mu * sc);
  const float fmu = static_cast<float>(mu), fv = static_cast<float>(v);
  scale[u] = fsc; shift[u] = fsh; mean[u] = fmu; var[u] = fv;
  scale_out[u] = fsc; shift_out[u] = fsh; mean_out[u] = fmu; var_out[u] = fv;
}

__global__ void __launch_bounds__(kThreads)
k_copy_feats(const float4 *__restrict__ src, const int U4, const int *__restrict__ num_pillars, float4 *__restrict__ dst) {
  const long long total = static_cast<long long>(*num_pillars) * U4;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads)
    dst[i] = src[i];
}

// One row-space pass = forward (keeps X_l, Y_l, m_l of every layer in the workspace) + backward over what it kept.
// Three users: mbev_pfn_backward (recompute, then backward), mbev_pfn_forward_train_rows (the forward half as THE
// train-mode forward of a training step) and mbev_pfn_backward_rows (the backward half over that workspace).
struct RowsPass {
  const MbevPfnParams *params;
  BwdWs w;
  DecoK dk;
  int L, T, train;
  int64_t rows_cap, cap;
  float eps;
  const int32_t *num_pillars_dev;
  const float *scale_shift, *batch_stats;       // the fold / statistics of an earlier forward (recompute mode), else null
  const float *const *gamma, *const *beta;      // forward mode: the layers' own BatchNorm parameters
  float *scale_shift_out, *batch_stats_out, *feats_out;  // forward mode outputs
  bool caller_aligned;
  cudaStream_t stream;

  const float *FSC(int l) const { return scale_shift + (2 * l) * MBEV_MAX_UNITS; }
  const float *FSH(int l) const { return scale_shift + (2 * l + 1) * MBEV_MAX_UNITS; }
  const float *FMEAN(int l) const { return batch_stats + (2 * l) * MBEV_MAX_UNITS; }
  const float *FVAR(int l) const { return batch_stats + (2 * l + 1) * MBEV_MAX_UNITS; }
  float *RST(int l, int which) const { return w.stats + (4 * l + which) * MBEV_MAX_UNITS; }
  // train mode: scale / shift / mean / var come from THIS pass's rows (k_row_stats); eval mode: the forward's (running
  // statistics, constants of the graph)
  const float *SC(int l) const { return train ? RST(l, 0) : FSC(l); }
  const float *SH(int l) const { return train ? RST(l, 1) : FSH(l); }
  const float *MEAN(int l) const { return train ? RST(l, 2) : FMEAN(l); }
  const float *VAR(int l) const { return train ? RST(l, 3) : FVAR(l); }
  // four-units-per-thread element-wise kernels: U % 4 == 0 (<= 256 for k_dz4's fold) and 16-byte aligned caller tensors
  bool vec4_ok(int U) const { return caller_aligned && U % 4 == 0 && U >= 4 && U <= 256 && kDz4Threads % (U / 4) == 0; }
};

constexpr int kEwBlocks = kNumSMs * 8;

int rows_forward(const RowsPass &c, const float *rows, const int32_t *kept_idx, const int32_t *num_points,
                 const int32_t *coors) {
  const BwdWs &w = c.w;
  cudaStream_t stream = c.stream;
  const MbevPfnParams *params = c.params;
  const int L = c.L, T = c.T;
  const int m_tiles_cap = static_cast<int>(std::min<int64_t>((c.rows_cap + 63) / 64, 1 << 30));
  {  // grid by the capacity (the pillar count lives on the device); one tile more so that row_off[P] always has an owner
    const int tiles = static_cast<int>(c.cap / kRowOffTile + 1);
    k_row_part<<<tiles, 256, 0, stream>>>(num_points, c.num_pillars_dev, T, w.row_tot);
    MBEV_CHECK_LAUNCH();
    k_row_offsets<<<tiles, 256, 0, stream>>>(num_points, c.num_pillars_dev, T, w.row_tot, w.row_off, w.num_rows);
    MBEV_CHECK_LAUNCH();
  }
  k_decorate_rows<<<kEwBlocks, kThreads, 0, stream>>>(rows, kept_idx, num_points, coors, c.num_pillars_dev, w.row_off,
                                                      c.dk, w.row_pillar, w.row_w, w.X[0]);
  MBEV_CHECK_LAUNCH();
  for (int l = 0; l < L; ++l) {
    const int U = params->units[l], K = params->in_dim[l];
    GemmK g{};  // Y_l (R,U) = X_l (R,K) * W_l^T ; W_l is (U,K) row-major => B(k,n) = W[n*K + k]
    g.A = w.X[l]; g.sAm = K; g.sAk = 1;
    g.B = params->weight[l]; g.sBk = 1; g.sBn = K;
    g.C = w.Y[l]; g.ldc = U;
    g.M = 0; g.N = U; g.K = K; g.m_dev = w.num_rows; g.k_dev = nullptr; g.split_stride = 0;
    int st;
    if (!c.train && tc_rows_ok(params, K, U)) {
      k_pe_prep_weights<<<dim3((U * K + 255) / 256, 1), 256, 0, stream>>>(params->weight[l], U, K, 1, w.img_fwd[l], 0);
      MBEV_CHECK_LAUNCH();
      st = launch_rows_gemm(w.X[l], w.num_rows, c.rows_cap, K, U, w.img_fwd[l], w.Y[l], stream);
    } else {
      st = launch_gemm(g, m_tiles_cap, U, 1, stream);
    }
    if (st) return st;
    if (c.train) {
      const int nls = std::max(1, 1024 / U);
      k_row_stats<<<kStatBlocks, dim3(U, nls), sizeof(double) * 2 * U * nls, stream>>>(w.Y[l], U, w.row_w, w.num_rows,
                                                                                      w.partials);
      MBEV_CHECK_LAUNCH();
      if (c.gamma) {
        k_row_stats_finalize_fwd<<<(U + 7) / 8, 256, 0, stream>>>(
            w.partials, kStatBlocks, U, c.num_pillars_dev, T, c.eps, c.gamma[l], c.beta[l], c.RST(l, 0), c.RST(l, 1),
            c.RST(l, 2), c.RST(l, 3), c.scale_shift_out + (2 * l) * MBEV_MAX_UNITS,
            c.scale_shift_out + (2 * l + 1) * MBEV_MAX_UNITS, c.batch_stats_out + (2 * l) * MBEV_MAX_UNITS,
            c.batch_stats_out + (2 * l + 1) * MBEV_MAX_UNITS);
      } else {
        k_row_stats_finalize<<<(U + 7) / 8, 256, 0, stream>>>(w.partials, kStatBlocks, U, c.num_pillars_dev, T, c.eps,
                                                              c.FSC(l), c.FSH(l), c.FMEAN(l), c.FVAR(l), c.RST(l, 0),
                                                              c.RST(l, 1), c.RST(l, 2), c.RST(l, 3));
      }
      MBEV_CHECK_LAUNCH();
    }
    const bool v4 = c.vec4_ok(U);
    if (v4 && l + 1 < L) {  // max, arg-max row and X_{l+1} = [a_l || m_l] in one pillar pass
      k_act_max_build4<<<kEwBlocks, kThreads, 0, stream>>>(w.Y[l], U, c.SC(l), c.SH(l), w.row_off, c.num_pillars_dev,
                                                          w.Mx[l], w.Arg[l], w.X[l + 1]);
      MBEV_CHECK_LAUNCH();
      continue;
    }
    if (v4) k_act_max4<<<kEwBlocks, kThreads, 0, stream>>>(w.Y[l], U, c.SC(l), c.SH(l), w.row_off, c.num_pillars_dev, w.Mx[l], w.Arg[l]);
    else k_act_max<<<kEwBlocks, kThreads, 0, stream>>>(w.Y[l], U, c.SC(l), c.SH(l), w.row_off, c.num_pillars_dev, w.Mx[l]);
    MBEV_CHECK_LAUNCH();
    if (l + 1 < L) {  // X_{l+1} = [a_l || m_l], kept until dW_{l+1} has been taken
      k_build_x<<<kEwBlocks, kThreads, 0, stream>>>(w.Y[l], U, c.SC(l), c.SH(l), w.Mx[l], w.row_pillar, w.num_rows, w.X[l + 1]);
      MBEV_CHECK_LAUNCH();
    }
  }
  return MBEV_OK;
}

int rows_backward(const RowsPass &c, const float *dfeats, float *const *dweight, float *const *dgamma,
                  float *const *dbeta) {
  const BwdWs &w = c.w;
  cudaStream_t stream = c.stream;
  const MbevPfnParams *params = c.params;
  const int L = c.L, T = c.T;
  const int m_tiles_cap = static_cast<int>(std::min<int64_t>((c.rows_cap + 63) / 64, 1 << 30));
  for (int l = L - 1; l >= 0; --l) {  // top layer first
    const int U = params->units[l], K = params->in_dim[l];
    const float *dxn = (l == L - 1) ? nullptr : w.DX;
    const int ldx = (l == L - 1) ? 0 : params->in_dim[l + 1];
    const bool v4 = c.vec4_ok(U) && (l != L - 1 || (reinterpret_cast<uintptr_t>(dfeats) & 15) == 0);
    if (v4 && T < 255) {  // (the arg-max row offset is a byte: <= T rows + the virtual one)
      const int nl = kDz4Threads / (U / 4);
      if (dxn) {
        k_dm4<<<kEwBlocks, kThreads, 0, stream>>>(dxn, ldx, U, w.row_off, c.num_pillars_dev, w.DM);
        MBEV_CHECK_LAUNCH();
      }
      k_dz_rows4<<<kDzBlocks, dim3(U / 4, nl), sizeof(double) * 2 * U * nl, stream>>>(
          w.Y[l], U, c.SC(l), c.SH(l), c.MEAN(l), c.VAR(l), c.eps, w.Arg[l], dxn ? w.DM : dfeats, dxn, ldx, w.row_off,
          w.row_pillar, w.num_rows, w.DZ, w.partials);
    } else if (v4) {
      const int nl = kDz4Threads / (U / 4);
      k_dz4<<<kDzBlocks, dim3(U / 4, nl), sizeof(double) * 2 * U * nl, stream>>>(
          w.Y[l], U, c.SC(l), c.SH(l), c.MEAN(l), c.VAR(l), c.eps, w.Mx[l], dfeats, dxn, ldx, w.row_off,
          c.num_pillars_dev, w.DZ, w.partials);
    } else {
      const int nl = std::max(1, kDzThreads / U);
      k_dz<<<kDzBlocks, dim3(U, nl), sizeof(double) * 2 * U * nl, stream>>>(
          w.Y[l], U, c.SC(l), c.SH(l), c.MEAN(l), c.VAR(l), c.eps, w.Mx[l], dfeats, dxn, ldx, w.row_off,
          c.num_pillars_dev, w.DZ, w.partials);
    }
    MBEV_CHECK_LAUNCH();
    k_bn_finalize<<<(U + 7) / 8, 256, 0, stream>>>(w.partials, kDzBlocks, U, c.num_pillars_dev, T, c.train, dgamma[l],
                                                       dbeta[l], w.c12);
    MBEV_CHECK_LAUNCH();
    if (v4) k_dy4<<<kEwBlocks, kThreads, 0, stream>>>(w.Y[l], U, c.SC(l), c.MEAN(l), c.VAR(l), c.eps, w.c12, w.row_w, w.num_rows, w.DZ);
    else k_dy<<<kEwBlocks, kThreads, 0, stream>>>(w.Y[l], U, c.SC(l), c.MEAN(l), c.VAR(l), c.eps, w.c12, w.row_w, w.num_rows, w.DZ);
    MBEV_CHECK_LAUNCH();
    {  // dW_l (U,K) = dY^T (U,R) * X_l (R,K): split over row slices, fixed-order reduce
      const long long split_stride = static_cast<long long>(U) * K;
      int nsplit = kSplitK;
      if (U % 4 == 0 && K % 4 == 0 && U <= 128 && K <= 128 && K >= 32) {  // whole output in one CTA tile
        if (U > 64) k_gemm_dw<128><<<kSplitK, 256, 0, stream>>>(w.DZ, w.X[l], U, K, w.num_rows, w.wpart, split_stride);
        else k_gemm_dw<64><<<kSplitK, 256, 0, stream>>>(w.DZ, w.X[l], U, K, w.num_rows, w.wpart, split_stride);
        MBEV_CHECK_LAUNCH();
      } else if (K <= 16 && U <= 128 && U % 4 == 0) {
        int umax = 0, inmax = 0;  // as many row slices as the partial-sum buffer holds, up to 6 small CTAs per SM
        for (int j = 0; j < L; ++j) {
          umax = std::max(umax, params->units[j]);
          inmax = std::max(inmax, params->in_dim[j]);
        }
        nsplit = static_cast<int>(std::min<long long>(kNumSMs * 6, static_cast<long long>(kSplitK) * umax * inmax / split_stride));
        const int threads = (U / 4 * K + 31) / 32 * 32;
        k_dw_narrow<<<nsplit, threads, 0, stream>>>(w.DZ, w.X[l], U, K, w.num_rows, w.wpart, split_stride);
        MBEV_CHECK_LAUNCH();
      } else {
        GemmK g{};
        g.A = w.DZ; g.sAm = 1; g.sAk = U;   // A(m=u, k=r) = DY[r*U + u]
        g.B = w.X[l]; g.sBk = K; g.sBn = 1;  // B(k=r, n) = X_l[r*K + n]
        g.C = w.wpart; g.ldc = K;
        g.M = U; g.N = K; g.K = 0; g.m_dev = nullptr; g.k_dev = w.num_rows;
        g.split_stride = split_stride;
        int st = launch_gemm(g, (U + 63) / 64, K, kSplitK, stream);
        if (st) return st;
      }
      k_reduce_splits<<<(U * K + 7) / 8, 256, 0, stream>>>(w.wpart, nsplit, split_stride, U * K, dweight[l]);
      MBEV_CHECK_LAUNCH();
    }
    if (l > 0) {  // dX (R,K) = dY (R,U) * W_l (U,K)
      GemmK g{};
      g.A = w.DZ; g.sAm = U; g.sAk = 1;
      g.B = params->weight[l]; g.sBk = K; g.sBn = 1;
      g.C = w.DX; g.ldc = K;
      g.M = 0; g.N = K; g.K = U; g.m_dev = w.num_rows; g.k_dev = nullptr; g.split_stride = 0;
      int st;
      if (tc_rows_ok(params, U, K)) {  // B = W_l^T: (N = in, K = units)
        k_pe_prep_weights<<<dim3((U * K + 255) / 256, 1), 256, 0, stream>>>(params->weight[l], K, U, 1, w.img_dx[l], 1);
        MBEV_CHECK_LAUNCH();
        st = launch_rows_gemm(w.DZ, w.num_rows, c.rows_cap, U, K, w.img_dx[l], w.DX, stream);
      } else {
        st = launch_gemm(g, m_tiles_cap, K, 1, stream);
      }
      if (st) return st;
    }
  }
  return MBEV_OK;
}

// R = sum_p (n_p + [n_p < T]) is only known on the device. Its host-side bound sizes the row buffers:
// the caller may pass a tight `rows_capacity_hint` (e.g. points + pillar capacity); otherwise P_cap * (T + 1).
int64_t rows_capacity(int64_t pillar_capacity, int T, int64_t hint) {
  return hint > 0 ? hint : pillar_capacity * (static_cast<int64_t>(T) + 1);
}

int check_layers(const MbevPfnParams *params, int C, int T) {
  if (!params) return MBEV_ERR_BAD_ARG;
  const int L = params->num_layers;
  if (L < 1 || L > MBEV_MAX_LAYERS || C < 3 || C > MBEV_MAX_POINT_DIM || T < 1) return MBEV_ERR_BAD_ARG;
  for (int l = 0; l < L; ++l) {
    if (!params->weight[l]) return MBEV_ERR_BAD_ARG;
    if (params->units[l] > MBEV_MAX_UNITS || params->units[l] < 1) return MBEV_ERR_UNSUPPORTED;
    if (params->in_dim[l] != (l ? 2 * params->units[l - 1] : params->in_dim[0])) return MBEV_ERR_BAD_ARG;
  }
  return MBEV_OK;
}

int make_deco(const MbevPfnParams *params, int C, int T, DecoK *out) {
  DecoK dk;
  dk.C = C; dk.T = T;
  dk.cluster = params->with_cluster_center != 0;
  dk.vcenter = params->with_voxel_center != 0;
  dk.dist = params->with_distance != 0;
  dk.legacy = params->legacy != 0;
  dk.vcd = params->voxel_center_dims;
  dk.D0 = C + (dk.cluster ? 3 : 0) + (dk.vcenter ? dk.vcd : 0) + (dk.dist ? 1 : 0);
  if (dk.D0 != params->in_dim[0]) return MBEV_ERR_BAD_ARG;
  dk.vx = params->vx; dk.vy = params->vy; dk.vz = params->vz;
  dk.xo = params->x_offset; dk.yo = params->y_offset; dk.zo = params->z_offset;
  *out = dk;
  return MBEV_OK;
}

}  // namespace
}  // namespace mbev

using namespace mbev;

extern "C" int mbev_pfn_backward_workspace_bytes(const MbevPfnParams *params, int T, int64_t pillar_capacity,
                                                 int64_t rows_capacity_hint, size_t *bytes) {
  if (!params || !bytes || params->num_layers < 1 || params->num_layers > MBEV_MAX_LAYERS || pillar_capacity < 0)
    return MBEV_ERR_BAD_ARG;
  *bytes = carve_bwd(nullptr, params, std::max<int64_t>(pillar_capacity, 1),
                     std::max<int64_t>(rows_capacity(pillar_capacity, T, rows_capacity_hint), 1)).bytes;
  return MBEV_OK;
}

extern "C" int mbev_pfn_backward(const float *rows, int C, const int32_t *kept_idx, const int32_t *num_points,
                                 const int32_t *coors, const int32_t *num_pillars_dev, int64_t pillar_capacity, int T,
                                 int64_t rows_capacity_hint, const MbevPfnParams *params, const float *scale_shift,
                                 const float *batch_stats, float eps, int train, const float *dfeats,
                                 float *const *dweight, float *const *dgamma, float *const *dbeta, void *workspace,
                                 size_t workspace_bytes, void *stream_) {
  if (!params || !num_points || !coors || !num_pillars_dev || !scale_shift || !batch_stats || !dfeats || !dweight ||
      !dgamma || !dbeta || !workspace)
    return MBEV_ERR_BAD_ARG;
  int st = check_layers(params, C, T);
  if (st) return st;
  const int L = params->num_layers;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  for (int l = 0; l < L; ++l)
    if (!dweight[l] || !dgamma[l] || !dbeta[l]) return MBEV_ERR_BAD_ARG;
  if (pillar_capacity <= 0) {
    for (int l = 0; l < L; ++l) {
      MBEV_CUDA(cudaMemsetAsync(dweight[l], 0, sizeof(float) * params->units[l] * params->in_dim[l], stream));
      MBEV_CUDA(cudaMemsetAsync(dgamma[l], 0, sizeof(float) * params->units[l], stream));
      MBEV_CUDA(cudaMemsetAsync(dbeta[l], 0, sizeof(float) * params->units[l], stream));
    }
    return MBEV_OK;
  }
  if (!rows) return MBEV_ERR_BAD_ARG;
  RowsPass c{};
  c.rows_cap = rows_capacity(pillar_capacity, T, rows_capacity_hint);
  c.cap = pillar_capacity;
  if (c.rows_cap * MBEV_MAX_UNITS > 0x7fffffffLL * 4) return MBEV_ERR_UNSUPPORTED;
  c.w = carve_bwd(workspace, params, pillar_capacity, c.rows_cap);
  if (workspace_bytes < c.w.bytes) return MBEV_ERR_WORKSPACE;
  st = make_deco(params, C, T, &c.dk);
  if (st) return st;
  c.params = params; c.L = L; c.T = T; c.train = train; c.eps = eps; c.num_pillars_dev = num_pillars_dev;
  c.scale_shift = scale_shift; c.batch_stats = batch_stats; c.stream = stream;
  c.caller_aligned = ((reinterpret_cast<uintptr_t>(scale_shift) | reinterpret_cast<uintptr_t>(batch_stats)) & 15) == 0;
  st = rows_forward(c, rows, kept_idx, num_points, coors);  // recompute: keeps X_l, Y_l and m_l of every layer
  if (st) return st;
  return rows_backward(c, dfeats, dweight, dgamma, dbeta);
}

// The train-mode forward of a TRAINING STEP: the row-space forward above run once, as the forward — statistics from the
// layers' own gamma / beta, features out — with every activation left in the workspace, which the caller keeps until
// mbev_pfn_backward_rows. The step then computes the PFN forward once instead of three times (tensor-core forward with its
// L statistics passes + K2''s recompute), and forward and backward see the very same rows (fp32 FMA products).
extern "C" int mbev_pfn_forward_train_rows(const float *rows, int C, const int32_t *kept_idx, const int32_t *num_points,
                                           const int32_t *coors, const int32_t *num_pillars_dev, int64_t pillar_capacity,
                                           int T, int64_t rows_capacity_hint, const MbevPfnParams *params,
                                           const float *const *gamma, const float *const *beta, float eps, float *feats,
                                           float *scale_shift_out, float *batch_stats_out, void *workspace,
                                           size_t workspace_bytes, void *stream_) {
  if (!params || !num_points || !coors || !num_pillars_dev || !gamma || !beta || !feats || !scale_shift_out ||
      !batch_stats_out || !workspace)
    return MBEV_ERR_BAD_ARG;
  int st = check_layers(params, C, T);
  if (st) return st;
  const int L = params->num_layers;
  for (int l = 0; l < L; ++l)
    if (!gamma[l] || !beta[l]) return MBEV_ERR_BAD_ARG;
  if (pillar_capacity <= 0) return MBEV_OK;
  if (!rows) return MBEV_ERR_BAD_ARG;
  RowsPass c{};
  c.rows_cap = rows_capacity(pillar_capacity, T, rows_capacity_hint);
  c.cap = pillar_capacity;
  if (c.rows_cap * MBEV_MAX_UNITS > 0x7fffffffLL * 4) return MBEV_ERR_UNSUPPORTED;
  c.w = carve_bwd(workspace, params, pillar_capacity, c.rows_cap);
  if (workspace_bytes < c.w.bytes) return MBEV_ERR_WORKSPACE;
  st = make_deco(params, C, T, &c.dk);
  if (st) return st;
  c.params = params; c.L = L; c.T = T; c.train = 1; c.eps = eps; c.num_pillars_dev = num_pillars_dev;
  c.gamma = gamma; c.beta = beta; c.scale_shift_out = scale_shift_out; c.batch_stats_out = batch_stats_out;
  c.stream = static_cast<cudaStream_t>(stream_);
  c.caller_aligned = true;  // statistics live in the workspace
  st = rows_forward(c, rows, kept_idx, num_points, coors);
  if (st) return st;
  const int U = params->units[L - 1];
  if (U % 4 == 0 && (reinterpret_cast<uintptr_t>(feats) & 15) == 0) {
    k_copy_feats<<<kEwBlocks, kThreads, 0, c.stream>>>(reinterpret_cast<const float4 *>(c.w.Mx[L - 1]), U / 4,
                                                       num_pillars_dev, reinterpret_cast<float4 *>(feats));
    MBEV_CHECK_LAUNCH();
  } else {
    MBEV_CUDA(cudaMemcpyAsync(feats, c.w.Mx[L - 1], sizeof(float) * static_cast<size_t>(pillar_capacity) * U,
                              cudaMemcpyDeviceToDevice, c.stream));
  }
  return MBEV_OK;
}

extern "C" int mbev_pfn_backward_rows(const int32_t *num_pillars_dev, int64_t pillar_capacity, int C, int T,
                                      int64_t rows_capacity_hint, const MbevPfnParams *params, float eps,
                                      const float *dfeats, float *const *dweight, float *const *dgamma,
                                      float *const *dbeta, void *workspace, size_t workspace_bytes, void *stream_) {
  if (!params || !num_pillars_dev || !dfeats || !dweight || !dgamma || !dbeta || !workspace) return MBEV_ERR_BAD_ARG;
  int st = check_layers(params, C, T);
  if (st) return st;
  const int L = params->num_layers;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  for (int l = 0; l < L; ++l)
    if (!dweight[l] || !dgamma[l] || !dbeta[l]) return MBEV_ERR_BAD_ARG;
  if (pillar_capacity <= 0) {
    for (int l = 0; l < L; ++l) {
      MBEV_CUDA(cudaMemsetAsync(dweight[l], 0, sizeof(float) * params->units[l] * params->in_dim[l], stream));
      MBEV_CUDA(cudaMemsetAsync(dgamma[l], 0, sizeof(float) * params->units[l], stream));
      MBEV_CUDA(cudaMemsetAsync(dbeta[l], 0, sizeof(float) * params->units[l], stream));
    }
    return MBEV_OK;
  }
  RowsPass c{};
  c.rows_cap = rows_capacity(pillar_capacity, T, rows_capacity_hint);
  c.cap = pillar_capacity;
  if (c.rows_cap * MBEV_MAX_UNITS > 0x7fffffffLL * 4) return MBEV_ERR_UNSUPPORTED;
  c.w = carve_bwd(workspace, params, pillar_capacity, c.rows_cap);  // the SAME carve as the forward's: same arguments
  if (workspace_bytes < c.w.bytes) return MBEV_ERR_WORKSPACE;
  c.params = params; c.L = L; c.T = T; c.train = 1; c.eps = eps; c.num_pillars_dev = num_pillars_dev;
  c.stream = stream;
  c.caller_aligned = true;
  return rows_backward(c, dfeats, dweight, dgamma, dbeta);
}
