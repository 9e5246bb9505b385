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
//   per layer  l = 0..L-1 :  Y_l = X_l W_l^T (k_gemm) ; m_l = segmented max of relu(bn(Y_l)) ; X_{l+1} = [a_l || m_l]
//   per layer  l = L-1..0 :  k_dz (arg-max routing + ReLU mask + BN sums) ; k_bn_finalize (dgamma, dbeta) ;
//                            k_dy ; dW_l = dY^T X_l (split-K, fixed-order reduce) ; dX = dY W_l
// Every reduction has a fixed order => run-to-run identical gradients. fp32 FMA throughout (1e-5 parity).

#include <algorithm>

#include "common.cuh"
#include "rows_gemm_tc.cuh"

namespace mbev {
namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------------
// row space
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_row_offsets(const int *__restrict__ num_points, const int *__restrict__ num_pillars, const int T,
              int *__restrict__ row_off, int *__restrict__ num_rows) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int P = *num_pillars;
  const int tid = threadIdx.x;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < P; base += 1024) {
    const int p = base + tid;
    int v = 0;
    if (p < P) {
      const int n = num_points[p];
      v = n + (n < T ? 1 : 0);
    }
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, x, d);
      if ((tid & 31) >= d) x += u;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = x;
    __syncthreads();
    if (tid < 32) {
      int y = s_warp[tid];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, y, d);
        if (tid >= d) y += u;
      }
      s_warp[tid] = y;
    }
    __syncthreads();
    const int excl = s_carry + x - v + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0);
    if (p < P) row_off[p] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) {
    row_off[P] = s_carry;
    *num_rows = s_carry;  // R, at a host-known address
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
  for (int mt = blockIdx.x; mt * BM < M; mt += gridDim.x) {
    const int m0 = mt * BM;
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float ra[8], rb[BN / 16];
    auto gload = [&](int kb) {
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
    gload(0);
    for (int kb = 0; kb < K; kb += BK) {
      __syncthreads();  // the previous block's readers are done with the tiles
      sstore();
      __syncthreads();
      if (kb + BK < K) gload(kb + BK);
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

// fixed-order reduction of split-K partials into dst (count elements)
__global__ void k_reduce_splits(const float *__restrict__ part, const int nsplit, const long long stride,
                                const int count, float *__restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int z = 0; z < nsplit; ++z) s += part[z * stride + i];
  dst[i] = s;
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

struct BwdWs {
  int *row_off, *row_pillar, *num_rows;
  float *row_w, *X, *DZ, *DX, *Mx[MBEV_MAX_LAYERS], *Y[MBEV_MAX_LAYERS], *c12, *wpart;
  float *stats;  // (L, 4, MBEV_MAX_UNITS): scale, shift, mean, var recomputed by this pass (train mode)
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
  w.row_pillar = c.take<int>(static_cast<size_t>(rows_cap));
  w.row_w = c.take<float>(static_cast<size_t>(rows_cap));
  w.X = c.take<float>(static_cast<size_t>(rows_cap) * inmax);
  w.DZ = c.take<float>(static_cast<size_t>(rows_cap) * umax);
  w.DX = c.take<float>(static_cast<size_t>(rows_cap) * inmax);
  for (int l = 0; l < p->num_layers; ++l) {
    w.Y[l] = c.take<float>(static_cast<size_t>(rows_cap) * p->units[l]);
    w.Mx[l] = c.take<float>(static_cast<size_t>(cap) * p->units[l]);
  }
  w.c12 = c.take<float>(2 * umax);
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
    if (n > 64) {
      k_gemm_rows<128><<<dim3(tiles, (n + 127) / 128), 256, 0, stream>>>(g);
    } else {
      k_gemm_rows<64><<<dim3(tiles, (n + 63) / 64), 256, 0, stream>>>(g);
    }
    MBEV_CHECK_LAUNCH();
    return MBEV_OK;
  }
  dim3 grid(std::max(1, std::min(m_tiles_cap, kNumSMs * 8)), (n + 63) / 64, splits);
  k_gemm<<<grid, 256, 0, stream>>>(g);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

}  // namespace
}  // namespace mbev

using namespace mbev;

// R = sum_p (n_p + [n_p < T]) is only known on the device. Its host-side bound sizes the row buffers:
// the caller may pass a tight `rows_capacity_hint` (e.g. points + pillar capacity); otherwise P_cap * (T + 1).
static int64_t rows_capacity(int64_t pillar_capacity, int T, int64_t hint) {
  return hint > 0 ? hint : pillar_capacity * (static_cast<int64_t>(T) + 1);
}

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
  const int L = params->num_layers;
  if (L < 1 || L > MBEV_MAX_LAYERS || C < 3 || C > MBEV_MAX_POINT_DIM || T < 1) return MBEV_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  for (int l = 0; l < L; ++l) {
    if (!dweight[l] || !dgamma[l] || !dbeta[l] || !params->weight[l]) return MBEV_ERR_BAD_ARG;
    if (params->units[l] > MBEV_MAX_UNITS || params->units[l] < 1) return MBEV_ERR_UNSUPPORTED;
    if (params->in_dim[l] != (l ? 2 * params->units[l - 1] : params->in_dim[0])) return MBEV_ERR_BAD_ARG;
  }
  if (pillar_capacity <= 0) {
    for (int l = 0; l < L; ++l) {
      MBEV_CUDA(cudaMemsetAsync(dweight[l], 0, sizeof(float) * params->units[l] * params->in_dim[l], stream));
      MBEV_CUDA(cudaMemsetAsync(dgamma[l], 0, sizeof(float) * params->units[l], stream));
      MBEV_CUDA(cudaMemsetAsync(dbeta[l], 0, sizeof(float) * params->units[l], stream));
    }
    return MBEV_OK;
  }
  if (!rows) return MBEV_ERR_BAD_ARG;
  const int64_t rows_cap = rows_capacity(pillar_capacity, T, rows_capacity_hint);
  if (rows_cap * MBEV_MAX_UNITS > 0x7fffffffLL * 4) return MBEV_ERR_UNSUPPORTED;
  const BwdWs w = carve_bwd(workspace, params, pillar_capacity, rows_cap);
  if (workspace_bytes < w.bytes) return MBEV_ERR_WORKSPACE;

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

  const int ew_blocks = kNumSMs * 8;
  const int m_tiles_cap = static_cast<int>(std::min<int64_t>((rows_cap + 63) / 64, 1 << 30));
  // train mode: scale / shift / mean / var come from THIS pass's rows (k_row_stats); eval mode: the forward's (running
  // statistics, constants of the graph)
  auto FSC = [&](int l) { return scale_shift + (2 * l) * MBEV_MAX_UNITS; };
  auto FSH = [&](int l) { return scale_shift + (2 * l + 1) * MBEV_MAX_UNITS; };
  auto FMEAN = [&](int l) { return batch_stats + (2 * l) * MBEV_MAX_UNITS; };
  auto FVAR = [&](int l) { return batch_stats + (2 * l + 1) * MBEV_MAX_UNITS; };
  auto RST = [&](int l, int which) { return w.stats + (4 * l + which) * MBEV_MAX_UNITS; };
  auto SC = [&](int l) -> const float * { return train ? RST(l, 0) : FSC(l); };
  auto SH = [&](int l) -> const float * { return train ? RST(l, 1) : FSH(l); };
  auto MEAN = [&](int l) -> const float * { return train ? RST(l, 2) : FMEAN(l); };
  auto VAR = [&](int l) -> const float * { return train ? RST(l, 3) : FVAR(l); };

  // ---- row space + forward recompute (keeps Y_l and m_l of every layer) --------------------------------
  k_row_offsets<<<1, 1024, 0, stream>>>(num_points, num_pillars_dev, T, w.row_off, w.num_rows);
  MBEV_CHECK_LAUNCH();
  k_decorate_rows<<<ew_blocks, kThreads, 0, stream>>>(rows, kept_idx, num_points, coors, num_pillars_dev, w.row_off, dk,
                                                      w.row_pillar, w.row_w, w.X);
  MBEV_CHECK_LAUNCH();
  for (int l = 0; l < L; ++l) {
    const int U = params->units[l], K = params->in_dim[l];
    GemmK g{};  // Y_l (R,U) = X_l (R,K) * W_l^T ; W_l is (U,K) row-major => B(k,n) = W[n*K + k]
    g.A = w.X; g.sAm = K; g.sAk = 1;
    g.B = params->weight[l]; g.sBk = 1; g.sBn = K;
    g.C = w.Y[l]; g.ldc = U;
    g.M = 0; g.N = U; g.K = K; g.m_dev = w.num_rows; g.k_dev = nullptr; g.split_stride = 0;
    int st;
    if (!train && tc_rows_ok(params, K, U)) {
      k_pe_prep_weights<<<dim3((U * K + 255) / 256, 1), 256, 0, stream>>>(params->weight[l], U, K, 1, w.img_fwd[l], 0);
      MBEV_CHECK_LAUNCH();
      st = launch_rows_gemm(w.X, w.num_rows, rows_cap, K, U, w.img_fwd[l], w.Y[l], stream);
    } else {
      st = launch_gemm(g, m_tiles_cap, U, 1, stream);
    }
    if (st) return st;
    if (train) {
      const int nls = std::max(1, 1024 / U);
      k_row_stats<<<kStatBlocks, dim3(U, nls), sizeof(double) * 2 * U * nls, stream>>>(w.Y[l], U, w.row_w, w.num_rows,
                                                                                      w.partials);
      MBEV_CHECK_LAUNCH();
      k_row_stats_finalize<<<(U + 7) / 8, 256, 0, stream>>>(w.partials, kStatBlocks, U, num_pillars_dev, T, eps, FSC(l),
                                                            FSH(l), FMEAN(l), FVAR(l), RST(l, 0), RST(l, 1), RST(l, 2),
                                                            RST(l, 3));
      MBEV_CHECK_LAUNCH();
    }
    k_act_max<<<ew_blocks, kThreads, 0, stream>>>(w.Y[l], U, SC(l), SH(l), w.row_off, num_pillars_dev, w.Mx[l]);
    MBEV_CHECK_LAUNCH();
    if (l + 1 < L) {
      k_build_x<<<ew_blocks, kThreads, 0, stream>>>(w.Y[l], U, SC(l), SH(l), w.Mx[l], w.row_pillar, w.num_rows, w.X);
      MBEV_CHECK_LAUNCH();
    }
  }
  // ---- backward, top layer first ------------------------------------------------------------------------
  for (int l = L - 1; l >= 0; --l) {
    const int U = params->units[l], K = params->in_dim[l];
    const int nl = std::max(1, kDzThreads / U);
    k_dz<<<kDzBlocks, dim3(U, nl), sizeof(double) * 2 * U * nl, stream>>>(
        w.Y[l], U, SC(l), SH(l), MEAN(l), VAR(l), eps, w.Mx[l], dfeats, (l == L - 1) ? nullptr : w.DX,
        (l == L - 1) ? 0 : params->in_dim[l + 1], w.row_off, num_pillars_dev, w.DZ, w.partials);
    MBEV_CHECK_LAUNCH();
    k_bn_finalize<<<(U + 7) / 8, 256, 0, stream>>>(w.partials, kDzBlocks, U, num_pillars_dev, T, train, dgamma[l],
                                                       dbeta[l], w.c12);
    MBEV_CHECK_LAUNCH();
    k_dy<<<ew_blocks, kThreads, 0, stream>>>(w.Y[l], U, SC(l), MEAN(l), VAR(l), eps, w.c12, w.row_w, w.num_rows, w.DZ);
    MBEV_CHECK_LAUNCH();
    // X_l: layer 0 = decorated rows, else [a_{l-1} || m_{l-1}]
    if (l == 0) {
      k_decorate_rows<<<ew_blocks, kThreads, 0, stream>>>(rows, kept_idx, num_points, coors, num_pillars_dev, w.row_off,
                                                          dk, w.row_pillar, w.row_w, w.X);
    } else {
      k_build_x<<<ew_blocks, kThreads, 0, stream>>>(w.Y[l - 1], params->units[l - 1], SC(l - 1), SH(l - 1), w.Mx[l - 1],
                                                    w.row_pillar, w.num_rows, w.X);
    }
    MBEV_CHECK_LAUNCH();
    {  // dW_l (U,K) = dY^T (U,R) * X_l (R,K): split-K over rows, fixed-order reduce
      GemmK g{};
      g.A = w.DZ; g.sAm = 1; g.sAk = U;   // A(m=u, k=r) = DY[r*U + u]
      g.B = w.X; g.sBk = K; g.sBn = 1;    // B(k=r, n) = X[r*K + n]
      g.C = w.wpart; g.ldc = K;
      g.M = U; g.N = K; g.K = 0; g.m_dev = nullptr; g.k_dev = w.num_rows;
      g.split_stride = static_cast<long long>(U) * K;
      int st = launch_gemm(g, (U + 63) / 64, K, kSplitK, stream);
      if (st) return st;
      k_reduce_splits<<<(U * K + 255) / 256, 256, 0, stream>>>(w.wpart, kSplitK, g.split_stride, U * K, dweight[l]);
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
        st = launch_rows_gemm(w.DZ, w.num_rows, rows_cap, U, K, w.img_dx[l], w.DX, stream);
      } else {
        st = launch_gemm(g, m_tiles_cap, K, 1, stream);
      }
      if (st) return st;
    }
  }
  return MBEV_OK;
}
