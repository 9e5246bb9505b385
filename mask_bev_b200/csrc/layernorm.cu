// K3+LN — scatter to the BEV canvas fused with the LayerNorm that follows it (sm_100a). SURVEY.md §8 row f1.
//
// Replaces mask_bev_encoders.py:91-92: `pseudo_img = self.middle_encode(...)` followed by
// `self._layer_norm(pseudo_img)` with `nn.LayerNorm([C, ny, nx], eps=1e-3)` (:75) — per frame a normalisation
// over ALL C*ny*nx elements with an element-wise affine of that same shape. Unfused, torch reads the canvas, the
// weight and the bias and writes the canvas again (4x the scatter's traffic). Here:
//   * the statistics come from the pillar features alone (every other cell is an exact zero):
//     sum = sum_p sum_c f, sumsq likewise, N = C*ny*nx; fp64, fixed-order two-stage reduction (deterministic);
//   * one streaming pass writes y = ((x - mean_b) * rstd_b) * w + bias with x = feature or 0, composed as in
//     k_scatter_run; tasks are ordered so that all frames of a run are in flight together and the run's weight /
//     bias come from HBM once and from L2 for the other frames: ~B*C*G*4 + 2*C*G*4 bytes of DRAM traffic instead of
//     ~4*B*C*G*4 (measured history: a frame-inner loop with dependent table/feature loads per store 3.5-4.0 ms).
// The backward (mbev_scatter_layernorm_backward, lower half of this file) streams dy once for dweight / dbias and the
// two per-frame sums, and finishes dfeats in the compact pillar space.
#include <algorithm>
#include <cmath>

#include "common.cuh"

#include "ln_stats.cuh"

namespace mbev {
namespace {

constexpr int kThreads = 256;
constexpr int kRun = 256;      // cells per warp task of k_scatter_ln

// A warp owns (run of 256 cells, channel chunk, ONE frame) and composes x exactly as k_scatter_run does (pillar
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

// ---- backward of the fused scatter + LayerNorm ------------------------------------------------------------------
// With xh = (x - mean_b) * rstd_b (x = pillar feature or 0), g = dy * w and M = C*G elements per frame:
//   dbias = sum_b dy,  dweight = sum_b dy * xh                       (dense: an empty cell has xh = -mean_b*rstd_b != 0)
//   dx = rstd_b * (g - S1_b / M - xh * S2_b / M),  S1_b = sum g,  S2_b = sum g * xh   (sums over the whole frame)
//   dfeats[p, :] = dx at the pillar's cell (the scatter's gather backward).
// Pass 1 (k_ln_bwd_dense_async) streams dy once: a warp owns (128 cells, 4 channels) and walks the frames with dweight /
// dbias in registers, so they are written once and dy / weight are read once; at occupied cells it parks g in the
// dfeats row of the pillar (4 channels = one 16-byte piece) and it leaves per-(frame, CTA) partial sums of S1, S2
// (fp32 over a lane's 16 terms, fp64 from there on, fixed order => run-to-run identical). Pass 2 folds the partials,
// pass 3 (k_ln_bwd_feats) turns the parked g into dfeats in place, in the compact pillar space.
constexpr int kBwdRun = 128;  // cells per warp task: lane owns 4 consecutive cells
constexpr int kBwdCh = 4;     // channels per warp task

__device__ __forceinline__ float4 ld_global_v4_stream(const float *p) {
  float4 v;  // read once, never again: keep it out of the way of weight / table / feature lines in L1 / L2
  asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// dy and the table rows are staged through shared memory by cp.async. (Round 1's first form kept ONE frame ahead in
// registers: 16 warps per SM at 128 registers = 48 % of the DRAM peak, long-scoreboard 5 per issue,
// profiles/r1g_lnbwd_full.txt.) Here every lane
// copies its own 16-byte pieces of the next kStages-1 frames into its own ring slots (no cross-thread sharing: the
// only synchronisation is cp.async.wait_group), so the bytes in flight no longer cost registers; the feature rows of
// frame b+1 are requested while frame b is processed, xh is one FMA, a run without any pillar in a frame skips the
// feature path warp-uniformly, and the per-frame sums are folded in fp32 inside the warp (512 terms), fp64 above it.
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int kStages>
__global__ void __launch_bounds__(kThreads, 2)
k_ln_bwd_dense_async(const float *__restrict__ dy, const float *__restrict__ feats, const int *__restrict__ table,
                     const float2 *__restrict__ stats, const float *__restrict__ lnw, const int batch, const int C,
                     const int G, const int nchunks, const long long tasks, float *__restrict__ dw,
                     float *__restrict__ db, float *__restrict__ gfeat, double2 *__restrict__ partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4 *ring = reinterpret_cast<float4 *>(smem_raw);  // [kStages][kBwdCh + 1][kThreads]; plane kBwdCh = table row
  double *s_sum = reinterpret_cast<double *>(smem_raw + sizeof(float4) * kStages * (kBwdCh + 1) * kThreads);  // [batch][warp][2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long task = static_cast<long long>(blockIdx.x) * (kThreads / 32) + warp;
  // channel chunk fastest: the 8 warps of a CTA share a table row and read whole 128-byte lines of the feature rows.
  // (K3's order — run fastest, sequential streams per plane — is WORSE here: forward 1.42 -> 1.86 ms, backward 1.51 ->
  // 2.07 ms on kitti_b16; probe build of the third session.)
  const int run = static_cast<int>(task / nchunks);
  const int ch0 = static_cast<int>(task - static_cast<long long>(run) * nchunks) * kBwdCh;
  const int g0 = run * kBwdRun + 4 * lane;
  const bool inb = task < tasks && g0 < G;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 w[kBwdCh], aw[kBwdCh], ab[kBwdCh];
#pragma unroll
  for (int k = 0; k < kBwdCh; ++k) {
    w[k] = inb ? __ldg(reinterpret_cast<const float4 *>(lnw + static_cast<size_t>(ch0 + k) * G + g0)) : z;
    aw[k] = ab[k] = z;
  }
  auto slot = [&](int stage, int k) { return ring + (stage * (kBwdCh + 1) + k) * kThreads + tid; };
  const size_t frame_stride = static_cast<size_t>(C) * G;
  const float *dy_lane = dy + static_cast<size_t>(ch0) * G + g0;
  auto issue = [&](int b) {
    const int st = b % kStages;
    cp_async16(slot(st, kBwdCh), table + static_cast<size_t>(b) * G + g0);
    const float *src = dy_lane + static_cast<size_t>(b) * frame_stride;
#pragma unroll
    for (int k = 0; k < kBwdCh; ++k) cp_async16(slot(st, k), src + static_cast<size_t>(k) * G);
  };
  auto features = [&](const int4 &pid, float4 (&f)[4]) {  // f[cell] = the 4 channels of that cell's pillar
    f[0] = pid.x >= 0 ? __ldg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(pid.x) * C + ch0)) : z;
    f[1] = pid.y >= 0 ? __ldg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(pid.y) * C + ch0)) : z;
    f[2] = pid.z >= 0 ? __ldg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(pid.z) * C + ch0)) : z;
    f[3] = pid.w >= 0 ? __ldg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(pid.w) * C + ch0)) : z;
  };
#pragma unroll
  for (int s0 = 0; s0 < kStages - 1; ++s0) {
    if (inb && s0 < batch) issue(s0);
    cp_async_commit();
  }
  cp_async_wait<kStages - 2>();  // frame 0 has landed
  int4 pid_n = make_int4(-1, -1, -1, -1);
  float4 f_n[4] = {z, z, z, z};
  bool occ_n = false;  // some cell of this warp's run holds a pillar in the frame
  {
    if (inb) pid_n = *reinterpret_cast<const int4 *>(slot(0, kBwdCh));
    occ_n = __any_sync(0xffffffffu, (pid_n.x & pid_n.y & pid_n.z & pid_n.w) >= 0);
    if (occ_n) features(pid_n, f_n);
  }
  float2 st_n = __ldg(stats);
  for (int b = 0; b < batch; ++b) {
    if (inb && b + kStages - 1 < batch) issue(b + kStages - 1);
    cp_async_commit();
    cp_async_wait<kStages - 2>();  // frames <= b + 1 have landed
    const int4 pid = pid_n;
    const bool occ = occ_n;
    const float2 st = st_n;
    float4 f[4] = {f_n[0], f_n[1], f_n[2], f_n[3]};
    if (b + 1 < batch) {  // warp-uniform
      st_n = __ldg(stats + b + 1);
      pid_n = make_int4(-1, -1, -1, -1);
      if (inb) pid_n = *reinterpret_cast<const int4 *>(slot((b + 1) % kStages, kBwdCh));
      occ_n = __any_sync(0xffffffffu, (pid_n.x & pid_n.y & pid_n.z & pid_n.w) >= 0);
      if (occ_n) features(pid_n, f_n);
    }
    float s1 = 0.f, s2 = 0.f;
    if (inb) {
      const float rstd = st.y, e = -st.x * st.y;  // xh = x * rstd + e; an empty cell has xh = e
      const int stg = b % kStages;
      if (!occ) {
#pragma unroll
        for (int k = 0; k < kBwdCh; ++k) {
          const float4 d = *slot(stg, k);
          ab[k].x += d.x;
          ab[k].y += d.y;
          ab[k].z += d.z;
          ab[k].w += d.w;
          aw[k].x = fmaf(d.x, e, aw[k].x);
          aw[k].y = fmaf(d.y, e, aw[k].y);
          aw[k].z = fmaf(d.z, e, aw[k].z);
          aw[k].w = fmaf(d.w, e, aw[k].w);
          s1 += fmaf(d.x, w[k].x, d.y * w[k].y) + fmaf(d.z, w[k].z, d.w * w[k].w);
        }
        s2 = e * s1;
      } else {
        const float xk[kBwdCh][4] = {{f[0].x, f[1].x, f[2].x, f[3].x},
                                     {f[0].y, f[1].y, f[2].y, f[3].y},
                                     {f[0].z, f[1].z, f[2].z, f[3].z},
                                     {f[0].w, f[1].w, f[2].w, f[3].w}};  // xk[channel][cell]
        float4 g[kBwdCh];
#pragma unroll
        for (int k = 0; k < kBwdCh; ++k) {
          const float4 d = *slot(stg, k);
          float4 xh;
          xh.x = fmaf(xk[k][0], rstd, e);
          xh.y = fmaf(xk[k][1], rstd, e);
          xh.z = fmaf(xk[k][2], rstd, e);
          xh.w = fmaf(xk[k][3], rstd, e);
          ab[k].x += d.x;
          ab[k].y += d.y;
          ab[k].z += d.z;
          ab[k].w += d.w;
          aw[k].x = fmaf(d.x, xh.x, aw[k].x);
          aw[k].y = fmaf(d.y, xh.y, aw[k].y);
          aw[k].z = fmaf(d.z, xh.z, aw[k].z);
          aw[k].w = fmaf(d.w, xh.w, aw[k].w);
          g[k].x = d.x * w[k].x;
          g[k].y = d.y * w[k].y;
          g[k].z = d.z * w[k].z;
          g[k].w = d.w * w[k].w;
          s1 += (g[k].x + g[k].y) + (g[k].z + g[k].w);
          s2 += fmaf(g[k].x, xh.x, g[k].y * xh.y) + fmaf(g[k].z, xh.z, g[k].w * xh.w);
        }
        if (pid.x >= 0)
          *reinterpret_cast<float4 *>(gfeat + static_cast<size_t>(pid.x) * C + ch0) = make_float4(g[0].x, g[1].x, g[2].x, g[3].x);
        if (pid.y >= 0)
          *reinterpret_cast<float4 *>(gfeat + static_cast<size_t>(pid.y) * C + ch0) = make_float4(g[0].y, g[1].y, g[2].y, g[3].y);
        if (pid.z >= 0)
          *reinterpret_cast<float4 *>(gfeat + static_cast<size_t>(pid.z) * C + ch0) = make_float4(g[0].z, g[1].z, g[2].z, g[3].z);
        if (pid.w >= 0)
          *reinterpret_cast<float4 *>(gfeat + static_cast<size_t>(pid.w) * C + ch0) = make_float4(g[0].w, g[1].w, g[2].w, g[3].w);
      }
    }
    // both sums in one butterfly: the lower half-warp keeps s1 and hands s2 over, the upper half the other way round;
    // lane 0 ends up with S1, lane 16 with S2 (fixed order)
    const bool upper = lane & 16;
    float v = (upper ? s2 : s1) + __shfl_xor_sync(0xffffffffu, upper ? s1 : s2, 16);
#pragma unroll
    for (int o = 8; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((lane & 15) == 0) s_sum[(b * (kThreads / 32) + warp) * 2 + (lane >> 4)] = static_cast<double>(v);
  }
  if (inb) {
#pragma unroll
    for (int k = 0; k < kBwdCh; ++k) {
      st_global_v4_stream(dw + static_cast<size_t>(ch0 + k) * G + g0, aw[k]);
      st_global_v4_stream(db + static_cast<size_t>(ch0 + k) * G + g0, ab[k]);
    }
  }
  __syncthreads();
  for (int b = tid; b < batch; b += kThreads) {
    double a = 0.0, q = 0.0;
#pragma unroll
    for (int wv = 0; wv < kThreads / 32; ++wv) {
      a += s_sum[(b * (kThreads / 32) + wv) * 2];
      q += s_sum[(b * (kThreads / 32) + wv) * 2 + 1];
    }
    partial[static_cast<size_t>(b) * gridDim.x + blockIdx.x] = make_double2(a, q);
  }
}

inline size_t bwd_async_smem(int stages, int batch) {
  return sizeof(float4) * stages * (kBwdCh + 1) * kThreads + sizeof(double2) * static_cast<size_t>(batch) * (kThreads / 32);
}

// msum[b] = (S1_b / M, S2_b / M): one CTA per frame, fixed-order strided sums + shared-memory tree
__global__ void __launch_bounds__(kThreads)
k_ln_bwd_finalize(const double2 *__restrict__ partial, const int nblocks, const double count,
                  float2 *__restrict__ msum) {
  __shared__ double s_a[kThreads], s_q[kThreads];
  const int b = blockIdx.x, t = threadIdx.x;
  double a = 0.0, q = 0.0;
  for (int j = t; j < nblocks; j += kThreads) {
    const double2 v = partial[static_cast<size_t>(b) * nblocks + j];
    a += v.x;
    q += v.y;
  }
  s_a[t] = a;
  s_q[t] = q;
  __syncthreads();
  for (int o = kThreads / 2; o; o >>= 1) {
    if (t < o) {
      s_a[t] += s_a[t + o];
      s_q[t] += s_q[t + o];
    }
    __syncthreads();
  }
  if (t == 0) msum[b] = make_float2(static_cast<float>(s_a[0] / count), static_cast<float>(s_q[0] / count));
}

// dfeats[p, :] = rstd_b * (g - S1_b/M - xh * S2_b/M), in place over the g parked by pass 1; one warp per pillar row.
// Rows that are not in the cell table (none for a table from mbev_voxelize; duplicates of a hand-made coors list) and
// rows at or beyond the pillar count get zeros.
__global__ void __launch_bounds__(kThreads)
k_ln_bwd_feats(const float *__restrict__ feats, const int *__restrict__ coors, const int *__restrict__ num_pillars,
               const int *__restrict__ table, const float2 *__restrict__ stats, const float2 *__restrict__ msum,
               const long long capacity, const int batch, const int C, const int ny, const int nx,
               float *__restrict__ dfeats) {
  const int lane = threadIdx.x & 31;
  const int P = *num_pillars;
  const int G = ny * nx;
  const long long nwarps = static_cast<long long>(gridDim.x) * (kThreads / 32);
  for (long long p = static_cast<long long>(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5); p < capacity;
       p += nwarps) {
    bool ok = false;
    int b = 0;
    if (p < P) {
      const int4 c = __ldg(reinterpret_cast<const int4 *>(coors) + p);  // (b, z, y, x)
      if (c.x >= 0 && c.x < batch && c.z >= 0 && c.z < ny && c.w >= 0 && c.w < nx) {
        b = c.x;
        ok = __ldg(table + static_cast<size_t>(b) * G + c.z * nx + c.w) == static_cast<int>(p);
      }
    }
    float *row = dfeats + static_cast<size_t>(p) * C;
    if (!ok) {
      for (int ch = lane; ch < C; ch += 32) row[ch] = 0.f;
      continue;
    }
    const float2 st = __ldg(stats + b), ms = __ldg(msum + b);
    const float *x = feats + static_cast<size_t>(p) * C;
    for (int ch = lane; ch < C; ch += 32) {
      const float xh = (__ldg(x + ch) - st.x) * st.y;
      row[ch] = st.y * ((row[ch] - ms.x) - xh * ms.y);
    }
  }
}

// ---- K3+LN forward, frame-walking schedule (walk = MBEV_LN_WALK_FRAMES). Same task shape as the backward's streaming pass: a warp
// owns (128 cells, 4 channels) and walks the B frames with weight and bias IN REGISTERS, so they are read from HBM once
// and never again from L2 (k_scatter_ln re-reads them from L2 for every frame: 15 x 0.66 GB of L2 -> SM traffic per
// kitti_b16 batch, and two 16-byte loads per plane and group in its inner loop). Table rows are requested two frames
// ahead, the feature rows (one 16-byte load per occupied cell) one frame ahead; a (run, frame) without a pillar writes
// fma(e_b, w, bias) warp-uniformly. The arithmetic per element is the very expression of k_scatter_ln, so the two
// forms agree bit for bit.
__global__ void __launch_bounds__(kThreads, 2)
k_scatter_ln_frames(const float *__restrict__ feats, const int *__restrict__ table, const float2 *__restrict__ stats,
                    const float *__restrict__ lnw, const float *__restrict__ lnb, const int batch, const int C,
                    const int G, const int nchunks, const long long tasks, float *__restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long task = static_cast<long long>(blockIdx.x) * (kThreads / 32) + warp;
  const int run = static_cast<int>(task / nchunks);  // channel chunk fastest (see k_ln_bwd_dense_async)
  const int ch0 = static_cast<int>(task - static_cast<long long>(run) * nchunks) * kBwdCh;
  const int g0 = run * kBwdRun + 4 * lane;
  const bool inb = task < tasks && g0 < G;  // G % 4 == 0: a lane's four cells are in or out together
  if (task >= tasks) return;                // warp-uniform (no block-level synchronisation in this kernel)
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const int4 none = make_int4(-1, -1, -1, -1);
  float4 w[kBwdCh], bi[kBwdCh];
#pragma unroll
  for (int k = 0; k < kBwdCh; ++k) {
    w[k] = inb ? __ldg(reinterpret_cast<const float4 *>(lnw + static_cast<size_t>(ch0 + k) * G + g0)) : z;
    bi[k] = inb ? __ldg(reinterpret_cast<const float4 *>(lnb + static_cast<size_t>(ch0 + k) * G + g0)) : z;
  }
  auto row = [&](int b) {
    return (inb && b < batch) ? __ldg(reinterpret_cast<const int4 *>(table + static_cast<size_t>(b) * G + g0)) : none;
  };
  auto features = [&](const int4 &pid, float4 (&f)[4]) {  // f[cell] = the 4 channels of that cell's pillar
    f[0] = pid.x >= 0 ? __ldg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(pid.x) * C + ch0)) : z;
    f[1] = pid.y >= 0 ? __ldg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(pid.y) * C + ch0)) : z;
    f[2] = pid.z >= 0 ? __ldg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(pid.z) * C + ch0)) : z;
    f[3] = pid.w >= 0 ? __ldg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(pid.w) * C + ch0)) : z;
  };
  int4 pid_n = row(0), pid_nn = row(1);
  float4 f_n[4] = {z, z, z, z};
  bool occ_n = __any_sync(0xffffffffu, (pid_n.x & pid_n.y & pid_n.z & pid_n.w) >= 0);
  if (occ_n) features(pid_n, f_n);
  float2 st_n = __ldg(stats);
  float *o = out + static_cast<size_t>(ch0) * G + g0;
  const size_t frame_stride = static_cast<size_t>(C) * G;
  for (int b = 0; b < batch; ++b) {
    const bool occ = occ_n;
    const float2 st = st_n;
    float4 f[4] = {f_n[0], f_n[1], f_n[2], f_n[3]};
    if (b + 1 < batch) {  // warp-uniform
      st_n = __ldg(stats + b + 1);
      pid_n = pid_nn;
      pid_nn = row(b + 2);
      occ_n = __any_sync(0xffffffffu, (pid_n.x & pid_n.y & pid_n.z & pid_n.w) >= 0);
      if (occ_n) features(pid_n, f_n);
    }
    if (inb) {
      const float mean = st.x, rstd = st.y;
      float *ob = o + static_cast<size_t>(b) * frame_stride;
      if (!occ) {
        const float e = __fmul_rn(__fsub_rn(0.f, mean), rstd);  // ((0 - mean) * rstd), the general expression at x = 0
#pragma unroll
        for (int k = 0; k < kBwdCh; ++k) {
          float4 y;
          y.x = __fmaf_rn(e, w[k].x, bi[k].x);
          y.y = __fmaf_rn(e, w[k].y, bi[k].y);
          y.z = __fmaf_rn(e, w[k].z, bi[k].z);
          y.w = __fmaf_rn(e, w[k].w, bi[k].w);
          st_global_v4_stream_nc(ob + static_cast<size_t>(k) * G, y);
        }
      } else {
        const float xk[kBwdCh][4] = {{f[0].x, f[1].x, f[2].x, f[3].x},
                                     {f[0].y, f[1].y, f[2].y, f[3].y},
                                     {f[0].z, f[1].z, f[2].z, f[3].z},
                                     {f[0].w, f[1].w, f[2].w, f[3].w}};  // xk[channel][cell]
#pragma unroll
        for (int k = 0; k < kBwdCh; ++k) {
          float4 y;  // ((x - mean) * rstd) * w + b, the operation order of torch's LayerNorm kernel
          y.x = __fmaf_rn(__fmul_rn(__fsub_rn(xk[k][0], mean), rstd), w[k].x, bi[k].x);
          y.y = __fmaf_rn(__fmul_rn(__fsub_rn(xk[k][1], mean), rstd), w[k].y, bi[k].y);
          y.z = __fmaf_rn(__fmul_rn(__fsub_rn(xk[k][2], mean), rstd), w[k].z, bi[k].z);
          y.w = __fmaf_rn(__fmul_rn(__fsub_rn(xk[k][3], mean), rstd), w[k].w, bi[k].w);
          st_global_v4_stream_nc(ob + static_cast<size_t>(k) * G, y);
        }
      }
    }
  }
}

struct LnWs {
  double2 *partial;
  size_t bytes;
};

struct LnBwdWs {
  double2 *partial;  // (batch, blocks of pass 1)
  float2 *msum;      // (batch)
  size_t bytes;
  long long tasks;
  int blocks, nchunks;
};

LnBwdWs carve_bwd(void *ws, int batch, int c_out, int G) {
  Carver c(ws);
  LnBwdWs w;
  w.nchunks = c_out / kBwdCh;
  w.tasks = static_cast<long long>((G + kBwdRun - 1) / kBwdRun) * w.nchunks;
  w.blocks = static_cast<int>((w.tasks + kThreads / 32 - 1) / (kThreads / 32));
  w.partial = c.take<double2>(static_cast<size_t>(batch) * w.blocks);
  w.msum = c.take<float2>(static_cast<size_t>(batch));
  w.bytes = c.off;
  return w;
}

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
                                              const float *ln_bias, float eps, int walk, float *out, float *stats_out,
                                              void *workspace, size_t workspace_bytes, void *stream_) {
  if (!cell_table || !pillar_base || !ln_weight || !ln_bias || !out || !stats_out || !workspace) return MBEV_ERR_BAD_ARG;
  if (!mbev_scatter_layernorm_supported(batch, c_out, ny, nx, out, ln_weight, ln_bias)) return MBEV_ERR_UNSUPPORTED;
  if (!(eps >= 0.f)) return MBEV_ERR_BAD_ARG;
  if (walk != MBEV_LN_WALK_RUNS && walk != MBEV_LN_WALK_FRAMES) return MBEV_ERR_BAD_ARG;
  if (walk == MBEV_LN_WALK_FRAMES && (c_out % kBwdCh || (reinterpret_cast<uintptr_t>(feats) & 15)))
    return MBEV_ERR_UNSUPPORTED;
  const LnWs w = carve(workspace, batch);
  if (workspace_bytes < w.bytes) return MBEV_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int G = ny * nx;
  k_ln_partials<<<dim3(kStatBlocks, batch), kStatThreads, 0, stream>>>(feats, pillar_base, c_out, w.partial);
  MBEV_CHECK_LAUNCH();
  float2 *stats = reinterpret_cast<float2 *>(stats_out);
  k_ln_finalize<<<(batch + 127) / 128, 128, 0, stream>>>(w.partial, batch, static_cast<double>(c_out) * G,
                                                        static_cast<double>(eps), stats);
  MBEV_CHECK_LAUNCH();
  // MBEV_LN_WALK_RUNS: k_scatter_ln, a warp per (256-cell run, channel chunk, frame); MBEV_LN_WALK_FRAMES:
  // k_scatter_ln_frames, a warp per (128 cells, 4 channels) walking the frames with weight / bias in registers
  if (walk == MBEV_LN_WALK_FRAMES) {
    const int nchunks = c_out / kBwdCh;
    const long long ftasks = static_cast<long long>((G + kBwdRun - 1) / kBwdRun) * nchunks;
    const int fblocks = static_cast<int>((ftasks + kThreads / 32 - 1) / (kThreads / 32));
    k_scatter_ln_frames<<<fblocks, kThreads, 0, stream>>>(feats, cell_table, stats, ln_weight, ln_bias, batch, c_out, G,
                                                         nchunks, ftasks, out);
    MBEV_CHECK_LAUNCH();
    return MBEV_OK;
  }
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

extern "C" int mbev_scatter_layernorm_backward_supported(int batch, int c_out, int ny, int nx) {
  if (batch < 1 || batch > MBEV_MAX_BATCH || c_out < kBwdCh || (c_out % kBwdCh) || ny < 1 || nx < 1) return 0;
  const int64_t G = static_cast<int64_t>(ny) * nx;
  if ((G & 3) || G * batch > 0x7fffffffLL) return 0;
  return 1;
}

extern "C" int mbev_scatter_layernorm_backward_workspace_bytes(int batch, int c_out, int ny, int nx, size_t *bytes) {
  if (!bytes) return MBEV_ERR_BAD_ARG;
  if (!mbev_scatter_layernorm_backward_supported(batch, c_out, ny, nx)) return MBEV_ERR_UNSUPPORTED;
  *bytes = carve_bwd(nullptr, batch, c_out, ny * nx).bytes;
  return MBEV_OK;
}

extern "C" int mbev_scatter_layernorm_backward(const float *dout, const float *feats, const int32_t *cell_table,
                                               const int32_t *coors, const int32_t *num_pillars_dev,
                                               int64_t pillar_capacity, int batch, int c_out, int ny, int nx,
                                               const float *ln_weight, const float *stats, float *dfeats,
                                               float *dweight, float *dbias, void *workspace, size_t workspace_bytes,
                                               void *stream_) {
  if (!dout || !cell_table || !num_pillars_dev || !ln_weight || !stats || !dweight || !dbias || !workspace)
    return MBEV_ERR_BAD_ARG;
  if (pillar_capacity < 0 || (pillar_capacity > 0 && (!feats || !coors || !dfeats))) return MBEV_ERR_BAD_ARG;
  if (!mbev_scatter_layernorm_backward_supported(batch, c_out, ny, nx)) return MBEV_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(feats) | reinterpret_cast<uintptr_t>(ln_weight) |
       reinterpret_cast<uintptr_t>(dfeats) | reinterpret_cast<uintptr_t>(dweight) | reinterpret_cast<uintptr_t>(dbias) |
       reinterpret_cast<uintptr_t>(coors) | reinterpret_cast<uintptr_t>(cell_table)) & 15)
    return MBEV_ERR_UNSUPPORTED;
  const int G = ny * nx;
  const LnBwdWs w = carve_bwd(workspace, batch, c_out, G);
  if (workspace_bytes < w.bytes) return MBEV_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const float2 *st = reinterpret_cast<const float2 *>(stats);
  // dy / table rows staged by cp.async, 5 ring stages (4 when batch > 64: two CTAs must fit an SM). The opt-in to
  // > 48 KB of dynamic shared memory is per device and cheap: set on every launch (no process-global flag).
  if (batch <= 64) {
    MBEV_CUDA(cudaFuncSetAttribute(k_ln_bwd_dense_async<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(bwd_async_smem(5, 64))));
    k_ln_bwd_dense_async<5><<<w.blocks, kThreads, bwd_async_smem(5, batch), stream>>>(
        dout, feats, cell_table, st, ln_weight, batch, c_out, G, w.nchunks, w.tasks, dweight, dbias, dfeats, w.partial);
  } else {
    MBEV_CUDA(cudaFuncSetAttribute(k_ln_bwd_dense_async<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(bwd_async_smem(4, MBEV_MAX_BATCH))));
    k_ln_bwd_dense_async<4><<<w.blocks, kThreads, bwd_async_smem(4, batch), stream>>>(
        dout, feats, cell_table, st, ln_weight, batch, c_out, G, w.nchunks, w.tasks, dweight, dbias, dfeats, w.partial);
  }
  MBEV_CHECK_LAUNCH();
  k_ln_bwd_finalize<<<batch, kThreads, 0, stream>>>(w.partial, w.blocks, static_cast<double>(c_out) * G, w.msum);
  MBEV_CHECK_LAUNCH();
  if (pillar_capacity > 0) {
    const int64_t want = (pillar_capacity + kThreads / 32 - 1) / (kThreads / 32);
    const int blocks = static_cast<int>(std::min<int64_t>(want, kNumSMs * 16));
    k_ln_bwd_feats<<<blocks, kThreads, 0, stream>>>(feats, coors, num_pillars_dev, cell_table, st, w.msum,
                                                   pillar_capacity, batch, c_out, ny, nx, dfeats);
    MBEV_CHECK_LAUNCH();
  }
  return MBEV_OK;
}
