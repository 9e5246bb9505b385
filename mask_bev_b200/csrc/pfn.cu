// K2 — pillar feature net forward (sm_100a), all PFN layers fused per pillar-aligned row chunk.
//
// Replaces mmdet3d PillarFeatureNet.forward / PFNLayer.forward (mask_bev_encoders.py:119-120): decoration
// (cluster offset, legacy in-place centre offset, distance), then L x [Linear(no bias) -> BatchNorm1d ->
// ReLU -> max over the T slots -> concat(x, max)].
//
// Upstream runs dense over all P*T slots (88-97 % of them zero padding) and materialises every
// intermediate in HBM. Here:
//   * rows = the N_k stored points + ONE weighted virtual row per pillar that has padding (A.4 identity):
//     the virtual row enters layer 0 as zeros, flows through every layer like a real row, takes part in the
//     max, and counts (T - n_p) times in the train-mode batch statistics. Results are unchanged.
//   * the concat input [a_row || max_p] is split: the max_p half of each Linear is a per-PILLAR product
//     (pillar term), only the a_row half is per row — halves the per-row MACs for layers >= 1.
//   * a persistent CTA takes groups of consecutive pillars, packs them greedily into chunks of <= ROWS rows,
//     stages gathered points in shared memory, and keeps every activation on chip; per chunk the only HBM
//     traffic is the gathered points in and the (pillar, C_out) features out.
//   * the per-row GEMMs are register-tiled fp32 FMA (8 rows x 4 units per thread) out of shared memory:
//     fp32 parity at 1e-5 rules out plain TF32.
// Train mode: batch statistics need a grid-wide reduction per layer, so the same kernel is launched in
// STATS mode for layer s = 0..L-1 (layers < s applied with the statistics already known, layer s only
// accumulates sum / sum-of-squares in fp64, fixed reduction order => run-to-run identical), then once in
// FULL mode.
#include <algorithm>

#include "common.cuh"
#include "pfn_tcw2.cuh"

namespace mbev {
namespace {

constexpr int kThreads = 256;
constexpr int kGroup = 64;   // pillars per work unit
constexpr int kPtPitch = 8;  // floats per staged point row (>= MBEV_MAX_POINT_DIM)
constexpr int kMaxTilesPerThread = 2;

struct PfnK {
  int L;
  int K[MBEV_MAX_LAYERS];  // per-row inner dim: K0 = D0, Kl = U_{l-1}
  int U[MBEV_MAX_LAYERS];
  const float *wa[MBEV_MAX_LAYERS];  // (K_l, U_l) transposed per-row half  (workspace)
  const float *wb[MBEV_MAX_LAYERS];  // (U_{l-1}, U_l) transposed pillar half, l >= 1 (workspace)
  const float *scale[MBEV_MAX_LAYERS];
  const float *shift[MBEV_MAX_LAYERS];
  int C, D0, T;
  int cluster, vcenter, dist, legacy, vcd;
  float vx, vy, vz, xo, yo, zo;
  int rows, pcap, xp, kx, um;  // chunk limits and pitches
  // shared-memory offsets in 4-byte words
  int o_w[MBEV_MAX_LAYERS], o_x[2], o_m[2], o_pt, o_ss, o_pts, o_mean, o_ctr, o_roww, o_int;
  int smem_bytes;
  int stat_layer;    // -1: full forward; s: accumulate statistics of layer s and stop
  double *partials;  // (gridDim.x, 2, um) per-CTA sums for STATS mode
};

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

__global__ void __launch_bounds__(kThreads, 1)
k_pfn(const float *__restrict__ rows_src, const int *__restrict__ kept_idx, const int *__restrict__ num_points,
      const int *__restrict__ coors, const int *__restrict__ num_pillars, float *__restrict__ feats,
      const __grid_constant__ PfnK k) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int P = *num_pillars;
  const int T = k.T, ROWS = k.rows, XP = k.xp, UM = k.um;
  int *s_n = reinterpret_cast<int *>(smem + k.o_int);
  int *s_prow0 = s_n + kGroup;
  int *s_cstart = s_prow0 + kGroup;          // kGroup + 1
  int *s_rowp = s_cstart + kGroup + 1;       // ROWS
  int *s_rowsrc = s_rowp + ROWS;             // ROWS
  int *s_misc = s_rowsrc + ROWS;             // [0] = nchunks
  float *s_roww = smem + k.o_roww;
  float *s_pts = smem + k.o_pts;
  float *s_mean = smem + k.o_mean;
  float *s_ctr = smem + k.o_ctr;
  float *s_ss = smem + k.o_ss;
  float *s_ptm = smem + k.o_pt;

  // one-time: weights (per-row halves) and folded BN scale/shift into shared memory
  for (int l = 0; l < k.L; ++l) {
    const int n = k.K[l] * k.U[l];
    float *dst = smem + k.o_w[l];
    for (int i = tid; i < n; i += kThreads) dst[i] = __ldg(k.wa[l] + i);
    if (k.stat_layer < 0 || l < k.stat_layer) {
      for (int i = tid; i < k.U[l]; i += kThreads) {
        s_ss[(2 * l) * UM + i] = __ldg(k.scale[l] + i);
        s_ss[(2 * l + 1) * UM + i] = __ldg(k.shift[l] + i);
      }
    }
  }
  double st1[kMaxTilesPerThread][4], st2[kMaxTilesPerThread][4];
#pragma unroll
  for (int j = 0; j < kMaxTilesPerThread; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) st1[j][c] = st2[j][c] = 0.0;
  __syncthreads();

  const int last_layer = (k.stat_layer >= 0) ? k.stat_layer : k.L - 1;
  const int ngroups = (P + kGroup - 1) / kGroup;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int gp0 = grp * kGroup;
    const int np = min(kGroup, P - gp0);
    __syncthreads();
    if (tid < kGroup) s_n[tid] = (tid < np) ? __ldg(num_points + gp0 + tid) : 0;
    __syncthreads();
    if (tid == 0) {  // greedy packing of whole pillars into chunks of <= ROWS rows and <= pcap pillars
      int nch = 0, r = 0, cnt = 0;
      s_cstart[0] = 0;
      for (int p = 0; p < np; ++p) {
        const int need = s_n[p] + (s_n[p] < T ? 1 : 0);
        if (r + need > ROWS || cnt == k.pcap) {
          s_cstart[++nch] = p;
          r = 0;
          cnt = 0;
        }
        s_prow0[p] = r;
        r += need;
        ++cnt;
      }
      s_cstart[++nch] = np;
      s_misc[0] = nch;
    }
    __syncthreads();
    const int nchunks = s_misc[0];
    for (int ch = 0; ch < nchunks; ++ch) {
      const int p0 = s_cstart[ch], p1 = s_cstart[ch + 1];
      const int npil = p1 - p0;
      const int nlast = s_n[p1 - 1];
      const int nrows = s_prow0[p1 - 1] + nlast + (nlast < T ? 1 : 0);
      __syncthreads();  // previous chunk fully consumed
      // ---- row tables -------------------------------------------------------------------------------
      for (int idx = tid; idx < npil * (T + 1); idx += kThreads) {
        const int pl = idx / (T + 1), t = idx - pl * (T + 1);
        const int n = s_n[p0 + pl];
        const int row = s_prow0[p0 + pl] + t;
        if (t < n) {
          const size_t slot = static_cast<size_t>(gp0 + p0 + pl) * T + t;
          s_rowp[row] = pl;
          s_rowsrc[row] = kept_idx ? __ldg(kept_idx + slot) : static_cast<int>(slot);
          s_roww[row] = 1.f;
        } else if (t == n && n < T) {
          s_rowp[row] = pl;
          s_rowsrc[row] = -1;  // virtual row standing for the T - n zero-padded slots
          s_roww[row] = static_cast<float>(T - n);
        }
      }
      __syncthreads();
      // ---- gather points ----------------------------------------------------------------------------
      if (tid < nrows) {
        const int src = s_rowsrc[tid];
        float *dst = s_pts + tid * kPtPitch;
        if (src >= 0) {
          const float *p = rows_src + static_cast<size_t>(src) * k.C;
          if (k.C == 4) {
            *reinterpret_cast<float4 *>(dst) = __ldg(reinterpret_cast<const float4 *>(p));
          } else {
            for (int c = 0; c < k.C; ++c) dst[c] = __ldg(p + c);
          }
        }
      } else if (tid >= 128 && tid - 128 < npil) {  // pillar centres, on otherwise idle threads
        const int pl = tid - 128;
        const int4 c = __ldg(reinterpret_cast<const int4 *>(coors) + gp0 + p0 + pl);  // (b, z, y, x)
        // upstream: coors.type_as(features) * vx + x_offset — float32 multiply THEN add (no FMA contraction)
        s_ctr[pl * 4 + 0] = __fadd_rn(__fmul_rn(static_cast<float>(c.w), k.vx), k.xo);
        s_ctr[pl * 4 + 1] = __fadd_rn(__fmul_rn(static_cast<float>(c.z), k.vy), k.yo);
        s_ctr[pl * 4 + 2] = __fadd_rn(__fmul_rn(static_cast<float>(c.y), k.vz), k.zo);
      }
      __syncthreads();
      if (tid < npil) {  // cluster mean: sum over the pillar's slots in slot order / num_points
        const int n = s_n[p0 + tid];
        const float *p = s_pts + s_prow0[p0 + tid] * kPtPitch;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int t = 0; t < n; ++t) {
          sx = __fadd_rn(sx, p[t * kPtPitch + 0]);
          sy = __fadd_rn(sy, p[t * kPtPitch + 1]);
          sz = __fadd_rn(sz, p[t * kPtPitch + 2]);
        }
        const float fn = static_cast<float>(n);
        s_mean[tid * 4 + 0] = __fdiv_rn(sx, fn);
        s_mean[tid * 4 + 1] = __fdiv_rn(sy, fn);
        s_mean[tid * 4 + 2] = __fdiv_rn(sz, fn);
      }
      __syncthreads();
      // ---- decoration -> X0 (transposed: [d][row]) --------------------------------------------------
      float *x_in = smem + k.o_x[0];
      if (tid < ROWS) {
        const int row = tid;
        if (row < nrows && s_rowsrc[row] >= 0) {
          const int pl = s_rowp[row];
          const float *p = s_pts + row * kPtPitch;
          const float x = p[0], y = p[1], z = p[2];
          const float ex = __fsub_rn(x, s_ctr[pl * 4 + 0]), ey = __fsub_rn(y, s_ctr[pl * 4 + 1]),
                      ez = __fsub_rn(z, s_ctr[pl * 4 + 2]);
          const bool alias = k.vcenter && k.legacy;  // legacy: centre offset written in place over xyz
          const float r0 = alias ? ex : x, r1 = alias ? ey : y, r2 = (alias && k.vcd > 2) ? ez : z;  // 2-channel centre: z stays raw
          int d = 0;
          x_in[(d++) * XP + row] = r0;
          x_in[(d++) * XP + row] = r1;
          x_in[(d++) * XP + row] = r2;
          for (int c = 3; c < k.C; ++c) x_in[(d++) * XP + row] = p[c];
          if (k.cluster) {
            x_in[(d++) * XP + row] = __fsub_rn(x, s_mean[pl * 4 + 0]);
            x_in[(d++) * XP + row] = __fsub_rn(y, s_mean[pl * 4 + 1]);
            x_in[(d++) * XP + row] = __fsub_rn(z, s_mean[pl * 4 + 2]);
          }
          if (k.vcenter) {
            x_in[(d++) * XP + row] = ex;
            x_in[(d++) * XP + row] = ey;
            if (k.vcd > 2) x_in[(d++) * XP + row] = ez;
          }
          if (k.dist) x_in[(d++) * XP + row] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r0, r0), __fmul_rn(r1, r1)), __fmul_rn(r2, r2)));
        } else {
          for (int d = 0; d < k.D0; ++d) x_in[d * XP + row] = 0.f;  // virtual rows and chunk padding
        }
      }
      // ---- layers -----------------------------------------------------------------------------------
      int xi = 0, mi = 0;
      for (int l = 0; l <= last_layer; ++l) {
        const int K = k.K[l], U = k.U[l];
        const bool stats = (l == k.stat_layer);
        const bool last = (l == k.L - 1);
        float *m_cur = smem + k.o_m[mi];
        const float *m_prev = smem + k.o_m[mi ^ 1];
        const float *xin = smem + k.o_x[xi];
        float *xout = smem + k.o_x[xi ^ 1];
        const float *wt = smem + k.o_w[l];
        __syncthreads();  // x_in / m_prev complete
        // pillar term: pt[pl][u] = sum_k m_prev[pl][k] * Wb[k][u]   (weights via L1, coalesced over u)
        if (l > 0) {
          const int Kp = k.U[l - 1];
          const int nct = U >> 2;
          for (int it = tid; it < npil * nct; it += kThreads) {
            const int pl = it / nct, ct = it - pl * nct;
            const float *mp = m_prev + pl * UM;
            const float *wb = k.wb[l] + ct * 4;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int kk = 0; kk < Kp; ++kk) {
              const float mv = mp[kk];
              const float4 w = __ldg(reinterpret_cast<const float4 *>(wb + static_cast<size_t>(kk) * U));
              acc.x = fmaf(mv, w.x, acc.x);
              acc.y = fmaf(mv, w.y, acc.y);
              acc.z = fmaf(mv, w.z, acc.z);
              acc.w = fmaf(mv, w.w, acc.w);
            }
            *reinterpret_cast<float4 *>(s_ptm + pl * UM + ct * 4) = acc;
          }
        }
        if (!stats)
          for (int i = tid; i < npil * U; i += kThreads) m_cur[(i / U) * UM + (i % U)] = 0.f;
        __syncthreads();
        // per-row GEMM, 8 rows x 4 units per thread
        const int nct = U >> 2, nrt = ROWS >> 3;
        const int ntiles = nct * nrt;
#pragma unroll
        for (int j = 0; j < kMaxTilesPerThread; ++j) {
          const int ti = tid + j * kThreads;
          if (ti >= ntiles) break;
          const int rt = ti / nct, ct = ti - rt * nct;
          if (rt * 8 >= nrows) continue;
          float acc[8][4];
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
          const float *xa = xin + rt * 8;
          const float *wa = wt + ct * 4;
#pragma unroll 4
          for (int kk = 0; kk < K; ++kk) {
            const float4 a0 = ld4(xa + kk * XP), a1 = ld4(xa + kk * XP + 4);
            const float4 w = ld4(wa + kk * U);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], wv[c], acc[r][c]);
          }
          // epilogue
          float sc[4], sh[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            sc[c] = stats ? 1.f : s_ss[(2 * l) * UM + ct * 4 + c];
            sh[c] = stats ? 0.f : s_ss[(2 * l + 1) * UM + ct * 4 + c];
          }
          int run_pl = -1;
          float run_max[4] = {0.f, 0.f, 0.f, 0.f};
          float outv[4][8];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int row = rt * 8 + r;
            const bool valid = row < nrows;
            const int pl = valid ? s_rowp[row] : -1;
            float y[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) y[c] = acc[r][c];
            if (l > 0 && valid) {
              const float4 pt = ld4(s_ptm + pl * UM + ct * 4);
              y[0] += pt.x; y[1] += pt.y; y[2] += pt.z; y[3] += pt.w;
            }
            if (stats) {
              if (valid) {
                const double w = static_cast<double>(s_roww[row]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const double yd = static_cast<double>(y[c]);
                  st1[j][c] += w * yd;
                  st2[j][c] += w * yd * yd;
                }
              }
              continue;
            }
            float a[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              a[c] = fmaxf(fmaf(y[c], sc[c], sh[c]), 0.f);
              outv[c][r] = a[c];
            }
            if (valid) {
              if (pl != run_pl) {
                if (run_pl >= 0) {
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                    atomicMax(reinterpret_cast<int *>(m_cur + run_pl * UM + ct * 4 + c), __float_as_int(run_max[c]));
                }
                run_pl = pl;
#pragma unroll
                for (int c = 0; c < 4; ++c) run_max[c] = a[c];
              } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) run_max[c] = fmaxf(run_max[c], a[c]);
              }
            }
          }
          if (!stats) {
            if (run_pl >= 0) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                atomicMax(reinterpret_cast<int *>(m_cur + run_pl * UM + ct * 4 + c), __float_as_int(run_max[c]));
            }
            if (!last) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                float *o = xout + (ct * 4 + c) * XP + rt * 8;
                *reinterpret_cast<float4 *>(o) = make_float4(outv[c][0], outv[c][1], outv[c][2], outv[c][3]);
                *reinterpret_cast<float4 *>(o + 4) = make_float4(outv[c][4], outv[c][5], outv[c][6], outv[c][7]);
              }
            }
          }
        }
        xi ^= 1;
        mi ^= 1;
      }
      if (k.stat_layer < 0) {
        __syncthreads();
        const int U = k.U[k.L - 1];
        const float *m_fin = smem + k.o_m[mi ^ 1];
        float *dst = feats + static_cast<size_t>(gp0 + p0) * U;
        for (int i = tid; i < npil * U; i += kThreads) dst[i] = m_fin[(i / U) * UM + (i % U)];
      }
    }
  }
  if (k.stat_layer >= 0) {
    // deterministic CTA reduction: every thread parks its fp64 partials, then one thread per unit sums the
    // row-tiles that share its column in fixed order
    __syncthreads();
    double *red = reinterpret_cast<double *>(smem + k.o_x[0]);  // kThreads * kMaxTilesPerThread * 8 doubles
    const int U = k.U[k.stat_layer];
    const int nct = U >> 2, nrt = ROWS >> 3, ntiles = nct * nrt;
#pragma unroll
    for (int j = 0; j < kMaxTilesPerThread; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        red[((j * kThreads + tid) * 4 + c) * 2 + 0] = st1[j][c];
        red[((j * kThreads + tid) * 4 + c) * 2 + 1] = st2[j][c];
      }
    __syncthreads();
    for (int u = tid; u < U; u += kThreads) {
      const int ct = u >> 2, c = u & 3;
      double a = 0.0, b = 0.0;
      for (int rt = 0; rt < nrt; ++rt) {
        const int ti = rt * nct + ct;
        if (ti >= ntiles) break;
        a += red[(ti * 4 + c) * 2 + 0];
        b += red[(ti * 4 + c) * 2 + 1];
      }
      k.partials[(static_cast<size_t>(blockIdx.x) * 2 + 0) * UM + u] = a;
      k.partials[(static_cast<size_t>(blockIdx.x) * 2 + 1) * UM + u] = b;
    }
  }
}

// Split / transpose the nn.Linear weights: wa[l] = W_l[:, :K_l]^T (K_l, U_l), wb[l] = W_l[:, K_l:]^T.
__global__ void k_prep_weights(const float *__restrict__ w, const int U, const int in_dim, const int K,
                               float *__restrict__ wa, float *__restrict__ wb) {
  const int n = U * in_dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int u = i / in_dim, kk = i - u * in_dim;
    const float v = __ldg(w + i);
    if (kk < K) wa[kk * U + u] = v;
    else wb[(kk - K) * U + u] = v;
  }
}

// mean / biased variance over M = P*T slots, folded scale & shift (fixed summation order over CTAs).
__global__ void k_stats_finalize(const double *__restrict__ partials, const int nctas, const int um, const int U,
                                 const int *__restrict__ num_pillars, const int T, const float *__restrict__ gamma,
                                 const float *__restrict__ beta, const float eps, float *__restrict__ scale,
                                 float *__restrict__ shift, float *__restrict__ mean_out, float *__restrict__ var_out) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  double a = 0.0, b = 0.0;
  for (int c = 0; c < nctas; ++c) {
    a += partials[(static_cast<size_t>(c) * 2 + 0) * um + u];
    b += partials[(static_cast<size_t>(c) * 2 + 1) * um + u];
  }
  const double M = static_cast<double>(*num_pillars) * T;
  const double mean = M > 0 ? a / M : 0.0;
  double var = M > 0 ? b / M - mean * mean : 0.0;
  if (var < 0) var = 0;
  const double sc = static_cast<double>(gamma[u]) / sqrt(var + static_cast<double>(eps));
  scale[u] = static_cast<float>(sc);
  shift[u] = static_cast<float>(static_cast<double>(beta[u]) - mean * sc);
  mean_out[u] = static_cast<float>(mean);
  var_out[u] = static_cast<float>(var);
}

struct Plan {
  PfnK k;
  size_t ws_bytes;
  float *wa[MBEV_MAX_LAYERS], *wb[MBEV_MAX_LAYERS];
  float *scale[MBEV_MAX_LAYERS], *shift[MBEV_MAX_LAYERS];
  double *partials;
  int grid;
};

int make_plan(const MbevPfnParams *p, int C, int T, void *ws, Plan *out) {
  if (!p || p->num_layers < 1 || p->num_layers > MBEV_MAX_LAYERS) return MBEV_ERR_BAD_ARG;
  if (C < 3 || C > MBEV_MAX_POINT_DIM || T < 1) return MBEV_ERR_UNSUPPORTED;
  PfnK &k = out->k;
  k = PfnK();
  k.L = p->num_layers;
  k.C = C;
  k.T = T;
  k.cluster = p->with_cluster_center != 0;
  k.vcenter = p->with_voxel_center != 0;
  k.dist = p->with_distance != 0;
  k.legacy = p->legacy != 0;
  k.vcd = p->voxel_center_dims;
  if (k.vcenter && k.vcd != 2 && k.vcd != 3) return MBEV_ERR_BAD_ARG;
  k.D0 = C + (k.cluster ? 3 : 0) + (k.vcenter ? k.vcd : 0) + (k.dist ? 1 : 0);
  k.vx = p->vx; k.vy = p->vy; k.vz = p->vz;
  k.xo = p->x_offset; k.yo = p->y_offset; k.zo = p->z_offset;
  int um = 0, kx = k.D0;
  for (int l = 0; l < k.L; ++l) {
    const int U = p->units[l];
    if (U < 4 || (U & 3) || U > MBEV_MAX_UNITS) return MBEV_ERR_UNSUPPORTED;
    k.U[l] = U;
    k.K[l] = (l == 0) ? k.D0 : p->units[l - 1];
    const int expect_in = (l == 0) ? k.D0 : 2 * p->units[l - 1];
    if (p->in_dim[l] != expect_in) return MBEV_ERR_BAD_ARG;
    um = std::max(um, U);
    if (l < k.L - 1) kx = std::max(kx, U);
  }
  k.um = um;
  k.kx = kx;
  // workspace: transposed weights, train-mode scale/shift, per-CTA statistic partials
  Carver cw(ws);
  for (int l = 0; l < k.L; ++l) {
    out->wa[l] = cw.take<float>(static_cast<size_t>(k.K[l]) * k.U[l]);
    out->wb[l] = cw.take<float>(l ? static_cast<size_t>(k.U[l - 1]) * k.U[l] : 1);
    out->scale[l] = cw.take<float>(um);
    out->shift[l] = cw.take<float>(um);
    k.wa[l] = out->wa[l];
    k.wb[l] = out->wb[l];
  }
  out->grid = kNumSMs;
  out->partials = cw.take<double>(static_cast<size_t>(out->grid) * 2 * um);
  k.partials = out->partials;
  out->ws_bytes = cw.off;
  // shared-memory plan: largest chunk that fits 227 KB
  const int limit = 227 * 1024;
  for (int rows = 128; rows >= 8; rows >>= 1) {
    if (rows < T + 1) return MBEV_ERR_UNSUPPORTED;  // a pillar (T rows + its virtual row) must fit one chunk
    k.rows = rows;
    k.pcap = std::min(32, rows);
    k.xp = rows + 4;
    if ((rows / 8) * (um / 4) > kThreads * kMaxTilesPerThread) continue;
    int o = 0;
    for (int l = 0; l < k.L; ++l) { k.o_w[l] = o; o += (k.K[l] * k.U[l] + 3) & ~3; }
    const int xwords = std::max(kx * k.xp, kThreads * kMaxTilesPerThread * 8 * 2);  // also the fp64 reduce scratch
    k.o_x[0] = o; o += (xwords + 3) & ~3;
    k.o_x[1] = o; o += (kx * k.xp + 3) & ~3;
    k.o_m[0] = o; o += k.pcap * um;
    k.o_m[1] = o; o += k.pcap * um;
    k.o_pt = o; o += k.pcap * um;
    k.o_ss = o; o += k.L * 2 * um;
    k.o_pts = o; o += rows * kPtPitch;
    k.o_mean = o; o += k.pcap * 4;
    k.o_ctr = o; o += k.pcap * 4;
    k.o_roww = o; o += rows;
    k.o_int = o; o += 3 * kGroup + 1 + 2 * rows + 4;
    k.smem_bytes = o * 4;
    if (k.smem_bytes <= limit) return MBEV_OK;
  }
  return MBEV_ERR_UNSUPPORTED;
}

int launch_prep(const MbevPfnParams *p, const Plan &pl, cudaStream_t stream) {
  for (int l = 0; l < pl.k.L; ++l) {
    if (!p->weight[l]) return MBEV_ERR_BAD_ARG;
    const int n = pl.k.U[l] * p->in_dim[l];
    k_prep_weights<<<(n + 255) / 256, 256, 0, stream>>>(p->weight[l], pl.k.U[l], p->in_dim[l], pl.k.K[l], pl.wa[l],
                                                        pl.wb[l]);
    MBEV_CHECK_LAUNCH();
  }
  return MBEV_OK;
}

int launch_pfn(const Plan &pl, const float *rows, const int32_t *kept_idx, const int32_t *num_points,
               const int32_t *coors, const int32_t *num_pillars_dev, int64_t cap, float *feats, int stat_layer,
               cudaStream_t stream) {
  PfnK k = pl.k;
  k.stat_layer = stat_layer;
  // per device and cheap: set on every launch (no process-global "done" flag)
  MBEV_CUDA(cudaFuncSetAttribute(k_pfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int64_t groups = (cap + kGroup - 1) / kGroup;
  // STATS launches always use the full grid so that the partials array is fully written
  const int grid = stat_layer >= 0 ? pl.grid : static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(groups, pl.grid)));
  k_pfn<<<grid, kThreads, k.smem_bytes, stream>>>(rows, kept_idx, num_points, coors, num_pillars_dev, feats, k);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

// Which implementation runs the Linear layers of this stack: MBEV_GEMM_TCGEN05[_BF16] / MBEV_GEMM_FMA, or < 0.
int select_path(const MbevPfnParams *p, int C, int T) {  // capacity-independent
  if (!p) return MBEV_ERR_BAD_ARG;
  if (p->gemm_path < MBEV_GEMM_AUTO || p->gemm_path > MBEV_GEMM_TCGEN05_BF16) return MBEV_ERR_BAD_ARG;
  if (p->gemm_path == MBEV_GEMM_FMA) return MBEV_GEMM_FMA;
  tc::Plan tp;
  const int st = tc::make_plan(p, C, T, 1, nullptr, &tp);
  if (p->gemm_path == MBEV_GEMM_TCGEN05_BF16) {  // the warp-local two-pipeline kernel only
    if (st) return st;
    tc::Kargs k2 = tp.k;
    return tc::tcw2_plan(k2) ? MBEV_GEMM_TCGEN05_BF16 : MBEV_ERR_UNSUPPORTED;
  }
  if (st == MBEV_OK) return MBEV_GEMM_TCGEN05;
  if (st == MBEV_ERR_UNSUPPORTED && p->gemm_path == MBEV_GEMM_AUTO) return MBEV_GEMM_FMA;
  return st;
}

int raw_point_dim(const MbevPfnParams *p) {
  // C only affects D0 consistency, which in_dim[0] pins: recover it from in_dim[0]
  const int extra = (p->with_cluster_center ? 3 : 0) + (p->with_voxel_center ? p->voxel_center_dims : 0) +
                    (p->with_distance ? 1 : 0);
  return p->in_dim[0] - extra;
}

}  // namespace
}  // namespace mbev

using namespace mbev;

extern "C" int mbev_pfn_path(const MbevPfnParams *params, int T) {
  if (!params) return MBEV_ERR_BAD_ARG;
  const int path = select_path(params, raw_point_dim(params), T);
  if (path == MBEV_GEMM_FMA) {  // validate the stack against the FMA kernel's own limits
    Plan pl;
    const int st = make_plan(params, raw_point_dim(params), T, nullptr, &pl);
    if (st) return st;
  }
  return path;
}

extern "C" int mbev_pfn_workspace_bytes(const MbevPfnParams *params, int T, int64_t pillar_capacity, int train,
                                        size_t *bytes) {
    (void)train;
  if (!bytes || !params) return MBEV_ERR_BAD_ARG;
  const int C = raw_point_dim(params);
  const int path = select_path(params, C, T);
  if (path < 0) return path;
  if (path == MBEV_GEMM_TCGEN05 || path == MBEV_GEMM_TCGEN05_BF16) {
    tc::Plan tp;
    const int st = tc::make_plan(params, C, T, pillar_capacity, nullptr, &tp);
    if (st) return st;
    *bytes = tp.ws_bytes;
    return MBEV_OK;
  }
  Plan pl;
  const int st = make_plan(params, C, T, nullptr, &pl);
  if (st) return st;
  *bytes = pl.ws_bytes;
  return MBEV_OK;
}

extern "C" int mbev_pfn_forward(const float *rows, int C, const int32_t *kept_idx, const int32_t *num_points,
                                const int32_t *coors, const int32_t *num_pillars_dev, int64_t pillar_capacity, int T,
                                const MbevPfnParams *params, float *feats, void *workspace, size_t workspace_bytes,
                                void *stream_) {
  if (!num_points || !coors || !num_pillars_dev || !feats || !workspace) return MBEV_ERR_BAD_ARG;
  if (pillar_capacity <= 0) return MBEV_OK;
  if (!rows) return MBEV_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int path = select_path(params, C, T);
  if (path < 0) return path;
  if (path == MBEV_GEMM_TCGEN05 || path == MBEV_GEMM_TCGEN05_BF16) {
    tc::Plan tp;
    int st = tc::make_plan(params, C, T, pillar_capacity, workspace, &tp);
    if (st) return st;
    if (workspace_bytes < tp.ws_bytes) return MBEV_ERR_WORKSPACE;
    for (int l = 0; l < tp.k.L; ++l) {
      if (!params->scale[l] || !params->shift[l]) return MBEV_ERR_BAD_ARG;
      tp.k.scale[l] = params->scale[l];
      tp.k.shift[l] = params->shift[l];
    }
    st = tc::launch_prep(params, tp, num_points, num_pillars_dev, stream);
    if (st) return st;
    return tc::launch(tp, rows, kept_idx, num_points, coors, feats, -1, stream);
  }
  Plan pl;
  int st = make_plan(params, C, T, workspace, &pl);
  if (st) return st;
  if (workspace_bytes < pl.ws_bytes) return MBEV_ERR_WORKSPACE;
  for (int l = 0; l < pl.k.L; ++l) {
    if (!params->scale[l] || !params->shift[l]) return MBEV_ERR_BAD_ARG;
    pl.k.scale[l] = params->scale[l];
    pl.k.shift[l] = params->shift[l];
  }
  st = launch_prep(params, pl, stream);
  if (st) return st;
  return launch_pfn(pl, rows, kept_idx, num_points, coors, num_pillars_dev, pillar_capacity, feats, -1, stream);
}

extern "C" int mbev_pfn_forward_train(const float *rows, int C, const int32_t *kept_idx, const int32_t *num_points,
                                      const int32_t *coors, const int32_t *num_pillars_dev, int64_t pillar_capacity,
                                      int T, const MbevPfnParams *params, const float *const *gamma,
                                      const float *const *beta, float eps, float *feats, float *scale_shift_out,
                                      float *batch_stats_out, void *workspace, size_t workspace_bytes, void *stream_) {
  if (!num_points || !coors || !num_pillars_dev || !feats || !workspace || !gamma || !beta || !scale_shift_out ||
      !batch_stats_out)
    return MBEV_ERR_BAD_ARG;
  if (pillar_capacity <= 0) return MBEV_OK;
  if (!rows) return MBEV_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int path = select_path(params, C, T);
  if (path < 0) return path;
  if (path == MBEV_GEMM_TCGEN05_BF16) return MBEV_ERR_UNSUPPORTED;  // batch statistics need the fp32-accurate path
  if (path == MBEV_GEMM_TCGEN05) {
    tc::Plan tp;
    int st = tc::make_plan(params, C, T, pillar_capacity, workspace, &tp);
    if (st) return st;
    if (workspace_bytes < tp.ws_bytes) return MBEV_ERR_WORKSPACE;
    st = tc::launch_prep(params, tp, num_points, num_pillars_dev, stream);
    if (st) return st;
    for (int l = 0; l < tp.k.L; ++l) {
      if (!gamma[l] || !beta[l]) return MBEV_ERR_BAD_ARG;
      tp.k.scale[l] = scale_shift_out + (2 * l) * MBEV_MAX_UNITS;
      tp.k.shift[l] = scale_shift_out + (2 * l + 1) * MBEV_MAX_UNITS;
    }
    for (int s = 0; s < tp.k.L; ++s) {
      st = tc::launch(tp, rows, kept_idx, num_points, coors, feats, s, stream);
      if (st) return st;
      const int U = tp.k.U[s];
      k_stats_finalize<<<(U + 127) / 128, 128, 0, stream>>>(
          tp.k.partials, tp.grid, tp.k.um, U, num_pillars_dev, T, gamma[s], beta[s], eps,
          scale_shift_out + (2 * s) * MBEV_MAX_UNITS, scale_shift_out + (2 * s + 1) * MBEV_MAX_UNITS,
          batch_stats_out + (2 * s) * MBEV_MAX_UNITS, batch_stats_out + (2 * s + 1) * MBEV_MAX_UNITS);
      MBEV_CHECK_LAUNCH();
    }
    return tc::launch(tp, rows, kept_idx, num_points, coors, feats, -1, stream);
  }
  Plan pl;
  int st = make_plan(params, C, T, workspace, &pl);
  if (st) return st;
  if (workspace_bytes < pl.ws_bytes) return MBEV_ERR_WORKSPACE;
  st = launch_prep(params, pl, stream);
  if (st) return st;
  for (int l = 0; l < pl.k.L; ++l) {
    if (!gamma[l] || !beta[l]) return MBEV_ERR_BAD_ARG;
    // the library writes the folded scale/shift it uses straight into the caller's output block
    pl.k.scale[l] = scale_shift_out + (2 * l) * MBEV_MAX_UNITS;
    pl.k.shift[l] = scale_shift_out + (2 * l + 1) * MBEV_MAX_UNITS;
  }
  for (int s = 0; s < pl.k.L; ++s) {
    st = launch_pfn(pl, rows, kept_idx, num_points, coors, num_pillars_dev, pillar_capacity, feats, s, stream);
    if (st) return st;
    const int U = pl.k.U[s];
    k_stats_finalize<<<(U + 127) / 128, 128, 0, stream>>>(
        pl.partials, pl.grid, pl.k.um, U, num_pillars_dev, T, gamma[s], beta[s], eps,
        scale_shift_out + (2 * s) * MBEV_MAX_UNITS, scale_shift_out + (2 * s + 1) * MBEV_MAX_UNITS,
        batch_stats_out + (2 * s) * MBEV_MAX_UNITS, batch_stats_out + (2 * s + 1) * MBEV_MAX_UNITS);
    MBEV_CHECK_LAUNCH();
  }
  return launch_pfn(pl, rows, kept_idx, num_points, coors, num_pillars_dev, pillar_capacity, feats, -1, stream);
}

#ifdef MBEV_K2_TRACE
extern "C" __attribute__((visibility("default"))) int mbev_debug_k2_trace(long long *host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, mbev::tc::g_k2_trace, sizeof(long long) * 18 * 8 * 24));
}
#endif
