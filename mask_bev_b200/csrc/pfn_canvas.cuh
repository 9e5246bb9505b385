// K2+K3 fused — pillar feature net forward that also writes the dense BEV canvas (sm_100a).
//
// Separate kernels leave the two halves of the path on different rooflines one after the other: k_pfn_tcw2 is
// bound by the tensor/epilogue pipes and touches almost no DRAM, k_scatter_warp is a pure HBM write stream. Here
// they run in ONE persistent kernel (k_pfn_tcw2<true>, pfn_tcw2.cuh): the PFN warps walk the pillars in CELL order
// (not in pillar-id order), so the canvas region behind them is complete as soon as they have passed it, and 14
// writer warps per SM stream that region out (zeros, then the features dropped in) while the next chunks compute.
// mmdet3d PointPillarsScatter.forward_batch (mask_bev_encoders.py:122-123) is thereby hidden under
// PillarFeatureNet.forward (:119-120).
//
// This file holds the pre-pass that turns the cell table into the walk order and the balanced partition:
//   strip        128 consecutive global cells (gc = b * G + y * nx + x); the writer's unit (512 B per plane)
//   k_strip_sums per strip: pillars in it and their compact rows (n + [n < T]); per-CTA totals
//   k_strip_order exclusive prefix over strips (two-level, every CTA re-reduces the ~1 k CTA totals), then
//                  ord[j] = (pillar id, num_points) of the j-th occupied cell, and the CTA bounds: boundary c is the
//                  first strip whose cost prefix reaches c / 148 of the total, cost = 256 * rows + kappa * strips
//                  (rows load the PFN warps, strips load the writers / HBM).
//   k_sub_bounds  the 8 sub-ranges of each CTA: equal compact rows (the writers of a CTA share its strips).
#pragma once
#include "pfn_tcw2.cuh"

namespace mbev {
namespace tc {

constexpr int kCoStrips = 64;  // strips per pre-pass CTA (8 per warp)

struct CanvasPlan {
  int *packed;  // (NS) rows << 8 | pillars of each strip
  int2 *blk;    // (NB) (pillars, rows) of each pre-pass CTA
  int2 *ord;    // (capacity)
  int *sb, *pb; // (NSUB + 1)
  int *spre;    // (NS + 1) occupied cells before each strip
  int *rpre;    // (NS + 1) compact rows before each strip
  int *cb;      // (grid + 1) first strip of each CTA
  int NS, NB, NSUB, NC, G;
  size_t ws_bytes;  // including the tc::Plan part
};

__global__ void __launch_bounds__(256)
k_strip_sums(const int *__restrict__ table, const int *__restrict__ num_points, const int T, const int NC, const int NS,
             int *__restrict__ packed, int2 *__restrict__ blk) {
  __shared__ int s_c[8], s_r[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int csum = 0, rsum = 0;
#pragma unroll 2
  for (int i = 0; i < kCoStrips / 8; ++i) {
    const int s = blockIdx.x * kCoStrips + 8 * i + warp;
    if (s >= NS) break;
    const int gc = s * kCvStrip + 4 * lane;
    const int4 pid = (gc < NC) ? __ldg(reinterpret_cast<const int4 *>(table + gc)) : make_int4(-1, -1, -1, -1);
    int c = 0, r = 0;
    const int p[4] = {pid.x, pid.y, pid.z, pid.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (p[j] >= 0) {
        const int n = __ldg(num_points + p[j]);
        ++c;
        r += n + (n < T ? 1 : 0);
      }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    r = __reduce_add_sync(0xffffffffu, r);
    if (lane == 0) packed[s] = (r << 8) | c;
    csum += c;
    rsum += r;
  }
  if (lane == 0) {
    s_c[warp] = csum;
    s_r[warp] = rsum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0, r = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      c += s_c[w];
      r += s_r[w];
    }
    blk[blockIdx.x] = make_int2(c, r);
  }
}

__global__ void __launch_bounds__(256)
k_strip_order(const int *__restrict__ table, const int *__restrict__ num_points, const int NC, const int NS,
              const int *__restrict__ packed, const int2 *__restrict__ blk, const int NB, const int NSUB,
              const int kappa_q, int2 *__restrict__ ord, int *__restrict__ cb, int *__restrict__ spre,
              int *__restrict__ rpre) {
  __shared__ int s_red[4][8];
  __shared__ int s_pc[kCoStrips + 1], s_pr[kCoStrips + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, bid = blockIdx.x;
  // (a) pillars / rows before this CTA and in total
  int v[4] = {0, 0, 0, 0};  // base_c, base_r, tot_c, tot_r
  for (int j = tid; j < NB; j += 256) {
    const int2 b = __ldg(blk + j);
    v[2] += b.x;
    v[3] += b.y;
    if (j < bid) {
      v[0] += b.x;
      v[1] += b.y;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    v[q] = __reduce_add_sync(0xffffffffu, v[q]);
    if (lane == 0) s_red[q][warp] = v[q];
  }
  // (b) exclusive prefix over this CTA's strips
  const int s_lo = bid * kCoStrips;
  const int nloc = min(kCoStrips, NS - s_lo);
  if (warp == 0) {
    int c[2], r[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int s = 2 * lane + i;
      const int pk = (s < nloc) ? __ldg(packed + s_lo + s) : 0;
      c[i] = pk & 0xff;
      r[i] = pk >> 8;
    }
    int ic = c[0] + c[1], ir = r[0] + r[1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int tc = __shfl_up_sync(0xffffffffu, ic, o), tr = __shfl_up_sync(0xffffffffu, ir, o);
      if (lane >= o) {
        ic += tc;
        ir += tr;
      }
    }
    const int ec = ic - c[0] - c[1], er = ir - r[0] - r[1];
    s_pc[2 * lane] = ec;
    s_pr[2 * lane] = er;
    s_pc[2 * lane + 1] = ec + c[0];
    s_pr[2 * lane + 1] = er + r[0];
    if (lane == 31) {
      s_pc[kCoStrips] = ic;
      s_pr[kCoStrips] = ir;
    }
  }
  __syncthreads();
  int base_c = 0, base_r = 0, tot_c = 0, tot_r = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    base_c += s_red[0][w];
    base_r += s_red[1][w];
    tot_c += s_red[2][w];
    tot_r += s_red[3][w];
  }
  for (int i = tid; i < nloc; i += 256) {
    spre[s_lo + i] = base_c + s_pc[i];
    rpre[s_lo + i] = base_r + s_pr[i];
  }
  if (bid == NB - 1 && tid == 0) {
    spre[NS] = tot_c;
    rpre[NS] = tot_r;
  }
  // (c) the walk order: occupied cells of strip s land at ord[base_c + s_pc[s] ...] in cell order
  for (int i = warp; i < nloc; i += 8) {
    const int gc = (s_lo + i) * kCvStrip + 4 * lane;
    const int4 pid = (gc < NC) ? __ldg(reinterpret_cast<const int4 *>(table + gc)) : make_int4(-1, -1, -1, -1);
    const int mine = (pid.x >= 0) + (pid.y >= 0) + (pid.z >= 0) + (pid.w >= 0);
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    int at = base_c + s_pc[i] + inc - mine;
    const int p[4] = {pid.x, pid.y, pid.z, pid.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (p[j] >= 0) ord[at++] = make_int2(p[j], __ldg(num_points + p[j]));
  }
  // (d) CTA bounds that fall inside this CTA's strips (NSUB here = number of fused-kernel CTAs)
  const long long ctot = 256LL * tot_r + static_cast<long long>(kappa_q) * NS;
  const long long c_lo = 256LL * base_r + static_cast<long long>(kappa_q) * s_lo;
  const long long c_hi = 256LL * (base_r + s_pr[nloc]) + static_cast<long long>(kappa_q) * (s_lo + nloc);
  for (int kb = tid; kb <= NSUB; kb += 256) {
    if (kb == 0) {
      if (bid == 0) cb[0] = 0;
    } else if (kb == NSUB) {
      if (bid == NB - 1) cb[NSUB] = NS;
    } else {
      const long long target = (ctot * kb + NSUB - 1) / NSUB;
      if (c_lo < target && target <= c_hi) {
        int lo = 1, hi = nloc;  // smallest i in [1, nloc] with cost(i) >= target
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const long long cm = 256LL * (base_r + s_pr[mid]) + static_cast<long long>(kappa_q) * (s_lo + mid);
          if (cm >= target) hi = mid; else lo = mid + 1;
        }
        cb[kb] = s_lo + lo;
      }
    }
  }
}

// Second level: the 8 sub-ranges of a CTA split ITS strips by compact rows only — the PFN warp pairs walk their
// sub-range sequentially (so rows must balance), while the CTA's writers share all its strips dynamically.
__global__ void k_sub_bounds(const int *__restrict__ cb, const int *__restrict__ spre, const int *__restrict__ rpre,
                             const int ncta, int *__restrict__ sb, int *__restrict__ pb) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > 8 * ncta) return;
  const int c = min(t >> 3, ncta - 1), j = t - 8 * c;  // j = 8 only for the very last bound
  const int s0 = __ldg(cb + c), s1 = __ldg(cb + c + 1);
  int s = s1;
  if (j == 0) {
    s = s0;
  } else if (j < 8) {
    const long long r0 = __ldg(rpre + s0), r1 = __ldg(rpre + s1);
    const long long target = r0 + ((r1 - r0) * j + 7) / 8;
    int lo = s0, hi = s1;  // smallest strip index in [s0, s1] with rpre >= target
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(rpre + mid) >= target) hi = mid; else lo = mid + 1;
    }
    s = lo;
  }
  sb[t] = s;
  pb[t] = __ldg(spre + s);
}

// Is the fused kernel usable for this stack / canvas? (capacity-independent)
inline bool canvas_supported(const Plan &pl, int batch, int ny, int nx, const void *canvas) {
  Kargs k = pl.k;
  if (!tcw2_plan(k, true)) return false;
  const int64_t G = static_cast<int64_t>(ny) * nx;
  if (G < kCvStrip || (G & 3)) return false;  // a float4 group must not straddle a frame; a strip at most one
  if (G * batch > 0x7fffffffLL - 2 * kCvStrip) return false;
  if (reinterpret_cast<uintptr_t>(canvas) & 15) return false;
  return true;
}

// Does the fused batch entry (mbev_encode_batch) take this kernel by default? Not yet: on B200 (kitti_b16) it is
// correct but slower than K2 followed by K3 (DESIGN.md "K2+K3 fused"); MBEV_FUSED_CANVAS=1 opts in.
inline bool canvas_default_on() {
  static const bool on = getenv("MBEV_FUSED_CANVAS") != nullptr && atoi(getenv("MBEV_FUSED_CANVAS")) != 0;
  return on;
}

inline void make_canvas_plan(const Plan &pl, int batch, int ny, int nx, int64_t pillar_capacity, void *ws, CanvasPlan *cp) {
  Carver cw(ws);
  cw.off = align_up(pl.ws_bytes);
  cp->G = ny * nx;
  cp->NC = batch * cp->G;
  cp->NS = (cp->NC + kCvStrip - 1) / kCvStrip;
  cp->NB = (cp->NS + kCoStrips - 1) / kCoStrips;
  cp->NSUB = 8 * pl.grid;
  cp->packed = cw.take<int>(cp->NS);
  cp->blk = cw.take<int2>(cp->NB);
  cp->ord = cw.take<int2>(static_cast<size_t>(std::max<int64_t>(pillar_capacity, 1)));
  cp->sb = cw.take<int>(cp->NSUB + 1);
  cp->pb = cw.take<int>(cp->NSUB + 1);
  cp->spre = cw.take<int>(cp->NS + 1);
  cp->rpre = cw.take<int>(cp->NS + 1);
  cp->cb = cw.take<int>(pl.grid + 1);
  cp->ws_bytes = cw.off;
}

inline int launch_canvas(const MbevPfnParams *p, Plan &pl, const CanvasPlan &cp, const float *rows, const int32_t *kept_idx,
                         const int32_t *num_points, const int32_t *coors, const int32_t *table, float *feats,
                         float *canvas, cudaStream_t stream) {
  for (int l = 0; l < pl.k.L; ++l) {
    if (!p->weight[l]) return MBEV_ERR_BAD_ARG;
    pl.prep.w[l] = p->weight[l];
  }
  k_prep_weights_tc<<<dim3(16, pl.k.L), 256, 0, stream>>>(pl.prep);
  MBEV_CHECK_LAUNCH();
  // kappa: cost of one strip (64 KB of canvas) in compact rows, fixed point x256 (MBEV_CANVAS_KAPPA = rows per strip)
  static const double kappa = getenv("MBEV_CANVAS_KAPPA") ? atof(getenv("MBEV_CANVAS_KAPPA")) : 16.0;
  const int kappa_q = std::max(1, static_cast<int>(kappa * 256.0 + 0.5));
  k_strip_sums<<<cp.NB, 256, 0, stream>>>(table, num_points, pl.k.T, cp.NC, cp.NS, cp.packed, cp.blk);
  MBEV_CHECK_LAUNCH();
  k_strip_order<<<cp.NB, 256, 0, stream>>>(table, num_points, cp.NC, cp.NS, cp.packed, cp.blk, cp.NB, pl.grid, kappa_q,
                                           cp.ord, cp.cb, cp.spre, cp.rpre);
  MBEV_CHECK_LAUNCH();
  k_sub_bounds<<<(cp.NSUB + 1 + 255) / 256, 256, 0, stream>>>(cp.cb, cp.spre, cp.rpre, pl.grid, cp.sb, cp.pb);
  MBEV_CHECK_LAUNCH();
  Kargs k = pl.k;
  k.stat_layer = -1;
  static const int dbg = getenv("MBEV_TC_DBG") ? atoi(getenv("MBEV_TC_DBG")) : 0;
  k.dbg = dbg;
  if (!tcw2_plan(k, true)) return MBEV_ERR_UNSUPPORTED;
  static bool attr_done = false;
  if (!attr_done) {
    MBEV_CUDA(cudaFuncSetAttribute(k_pfn_tcw2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_done = true;
  }
  CanvasArgs cv;
  cv.table = table;
  cv.ord = cp.ord;
  cv.sb = cp.sb;
  cv.pb = cp.pb;
  cv.spre = cp.spre;
  cv.canvas = canvas;
  cv.G = cp.G;
  cv.NC = cp.NC;
  cv.Cout = k.U[k.L - 1];
  k_pfn_tcw2<true><<<pl.grid, kCvThreads, k.smem_bytes, stream>>>(rows, kept_idx, num_points, coors, nullptr, feats, k, cv);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

}  // namespace tc
}  // namespace mbev
