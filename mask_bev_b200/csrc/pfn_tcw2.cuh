// K2-TCW2 — two chunk pipelines per SM (the FULL-forward tcgen05 kernel when T <= 32).
//
// With the whole layer input A (K = 2U = 128 columns, hi + lo = 256) plus the accumulators in the 512 TMEM columns
// only one chunk fits per SM, and the tensor pipe idles while the epilogue warps work and vice versa (round 1's
// one-pipeline kernel: ~3.2 k cycles of epilogue and ~3.2 k of exposed MMA per layer). Here K is split in TIME: the
// x half of the next layer's input goes through the A columns first, and once those MMAs have retired
// (tcgen05.commit -> "x consumed") the replicated-max half is written into the SAME columns and accumulated on
// top. A chunk then needs A_hi 64 + A_lo 64 + D 128 = 256 columns, and two independent chunk pipelines ("sets")
// share the SM: 2 x (8 epilogue warps + 1 MMA-issuer warp). While one set waits for its MMAs the other one runs
// its epilogue, so the two pipes overlap without any cross-set synchronisation. The segmented max of a layer is
// computed while its x-part MMAs run.
//
// Row layout (what removes every block barrier and shared-memory transposition from the layer loop): a chunk is four
// WINDOWS of 32 rows, one per TMEM lane quadrant. Each quadrant walks its own contiguous, row-balanced pillar
// sub-range (8 per CTA: 2 sets x 4 quadrants) and packs whole pillars into its window, so a pillar never straddles a
// warp and the per-pillar max is a segmented max-scan over the lanes of one warp (log2(longest pillar) shuffle steps).
// Two epilogue warps share a quadrant and split the columns of every layer; each recomputes the window's packing on
// its own, so they never synchronise with each other, only with the set's MMA issuer through mbarriers. The cluster
// mean re-adds a pillar's points in slot order through shuffles (bit-identical to upstream's sequential sum). The
// layer-0 input is built in statically indexed registers (wide column order, pfn_tc.cuh): the kernel's shared memory
// is the resident weight image + scale / shift + barriers (~208 KB for [128,128,128]), which leaves room on the SM
// for the 17 KB CTA of the TMA-engine scatter (k_scatter_bulk) — K3 of batch i runs under K2 of batch i+1.
#pragma once
#include "pfn_tc.cuh"

namespace mbev {
namespace tc {

constexpr int kW2Sets = 2;
constexpr int kW2EpiWarps = 8 * kW2Sets;               // per set: 4 quadrants x 2 column halves
constexpr int kW2Threads = (kW2EpiWarps + 4) * 32;  // + one warpgroup: the two MMA issuers and two idle warps (setmaxnreg is warpgroup-wide)
constexpr int kW2SetCols = 256;                         // TMEM columns per set: A_hi [0,64) A_lo [64,128) D [128,256)
constexpr int kW2BarW = 0, kW2BarSet = 1, kW2BarsPerSet = 5, kW2NumBars = kW2BarSet + kW2Sets * kW2BarsPerSet;
enum { kW2X0 = 0, kW2D = 1, kW2XC = 2, kW2EVX = 3, kW2EVM = 4 };

#ifdef MBEV_K2_TRACE
// developer build only: cycle stamps of CTA 0 (18 warps x 8 chunks x 24 slots), read back by mbev_debug_k2_trace
__device__ long long g_k2_trace[18 * 8 * 24];
#define MBEV_TR(slot)                                                                                   \
  do {                                                                                                  \
    if (blockIdx.x == 0 && lane == 0 && c >= 6 && c < 14) g_k2_trace[(warp * 8 + (c - 6)) * 24 + (slot)] = clock64(); \
  } while (0)
#else
#define MBEV_TR(slot) do {} while (0)
#endif

struct Window {
  int cnt, nrows;  // pillars / rows packed into this window
  int pil;         // global pillar of this lane's row
  int s0, s1, t, n, maxlen;
  bool inwin, real;
};

// pack whole pillars cursor, cursor+1, ... into a 32-row window (lane r = row r); identical on every warp that
// calls it with the same cursor
// num_points of the 32 pillars a window starting at `cursor` may take (lane i: pillar cursor + i), requested a whole
// chunk before pack_window consumes it
__device__ __forceinline__ int load_np(const int *num_points, int cursor, int pend, int lane) {
  return cursor + lane < pend ? num_points[cursor + lane] : 0;  // plain load: np_sorted is written by this kernel
}
__device__ __forceinline__ int load_ord(const int *order, int cursor, int pend, int lane) {
  return cursor + lane < pend ? order[cursor + lane] : 0;
}

// Local counting sort of a sub-range's pillars by row count, longest first (stable, no atomics: ranks inside a group of
// 32 come from __match_any_sync, so the two warps that share a quadrant build the SAME order independently and write
// identical values). The per-pillar max is a log2(longest pillar of the window)-round shuffle all-reduce, a third of
// K2's time; pillars come numbered by first appearance, i.e. lengths are mixed at random and nearly every 32-row window
// holds a long pillar: 3-4 rounds everywhere. Sorted, the windows are uniform — on LiDAR frames a third of all rows sit
// in 2-row pillars (one point + the virtual row) and need ONE round. Sorting inside the sub-range keeps the row-balanced
// partition and the locality of the kept_idx / coors / feats rows. Results do not depend on the order at all.
constexpr int kSortClasses = 64;
__device__ __forceinline__ int sort_class(int n, int T) { return min(kSortClasses - 1, max(0, 33 - (n + (n < T ? 1 : 0)))); }
__device__ __forceinline__ void sort_subrange(const int *__restrict__ num_points, int p0, int pend, int T, int lane,
                                              int *s_cls, int *order, int *np_sorted) {
  s_cls[lane] = 0;
  s_cls[lane + 32] = 0;
  __syncwarp();
#pragma unroll 4
  for (int b = p0; b < pend; b += 32) {
    const int p = b + lane;
    const bool valid = p < pend;
    const int cls = valid ? sort_class(__ldg(num_points + p), T) : kSortClasses + lane;
    const unsigned m = __match_any_sync(0xffffffffu, cls);
    if (valid && lane == __ffs(m) - 1) s_cls[cls] += __popc(m);
    __syncwarp();
  }
  {  // exclusive prefix over the 64 classes
    const int a = s_cls[2 * lane], c = s_cls[2 * lane + 1];
    int inc = a + c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    __syncwarp();
    s_cls[2 * lane] = inc - a - c;
    s_cls[2 * lane + 1] = inc - c;
    __syncwarp();
  }
#pragma unroll 2
  for (int b = p0; b < pend; b += 32) {
    const int p = b + lane;
    const bool valid = p < pend;
    const int n = valid ? __ldg(num_points + p) : 0;
    const int cls = valid ? sort_class(n, T) : kSortClasses + lane;
    const unsigned m = __match_any_sync(0xffffffffu, cls);
    if (valid) {
      const int slot = p0 + s_cls[cls] + __popc(m & ((1u << lane) - 1u));
      order[slot] = p;
      np_sorted[slot] = n;
    }
    __syncwarp();
    if (valid && lane == __ffs(m) - 1) s_cls[cls] += __popc(m);
    __syncwarp();
  }
  __syncwarp();
}
// np / ord: num_points and pillar id of the candidate at sorted position cursor + lane
__device__ __forceinline__ Window pack_window(int np, int ord, int cursor, int pend, int T, int lane) {
  Window w;
  const bool cand = cursor + lane < pend;
  const int need = cand ? np + (np < T ? 1 : 0) : 0;
  int incl = need;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const bool taken = cand && incl <= 32;  // prefix-closed
  w.cnt = __popc(__ballot_sync(0xffffffffu, taken));
  w.nrows = w.cnt ? __shfl_sync(0xffffffffu, incl, w.cnt - 1) : 0;
  // bit r of `starts` = a pillar starts at row r: the pillar / first row / last row of row `lane` are bit counts and bit
  // positions of that mask (no shuffle search)
  const unsigned starts = __reduce_or_sync(0xffffffffu, taken ? (1u << (incl - need)) : 0u);
  const unsigned upto = starts & (0xffffffffu >> (31 - lane));  // starts at rows <= lane
  const unsigned above = starts & ~(0xffffffffu >> (31 - lane));
  const int pi = max(__popc(upto) - 1, 0);
  w.inwin = lane < w.nrows;
  w.n = __shfl_sync(0xffffffffu, np, pi);
  w.s0 = w.inwin ? 31 - __clz(upto) : lane;  // window padding: a segment of its own
  w.s1 = w.inwin ? (above ? __ffs(above) - 2 : w.nrows - 1) : lane;
  w.t = lane - w.s0;
  w.real = w.inwin && w.t < w.n;
  w.pil = __shfl_sync(0xffffffffu, ord, pi);
  w.maxlen = __reduce_max_sync(0xffffffffu, w.s1 - w.s0 + 1);
  return w;
}

// The gather of a chunk's points runs one chunk ahead of their use (software pipeline over the chunks of a quadrant):
// kept_idx row of the NEXT window right after this chunk's rendezvous, the point and the pillar's coordinates under the
// last layer's MMAs, decoration at the top of the next chunk — no global-memory latency on the chunk's critical chain.
// The bytes in flight are staged in shared memory by cp.async (48 bytes per lane: 8 point features + the pillar's
// (b, z, y, x)), not in registers: a prefetched REGISTER that ptxas spills turns the prefetch into a synchronous load.
constexpr int kStageFloats = 16;  // per lane: the raw point (8) + coordinates (4) in flight, then the decorated row (16)
struct Gathered {
  float pv[MBEV_MAX_POINT_DIM];
  int4 cc;
};
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ int gather_index(const Kargs &k, const Window &w, const int *__restrict__ kept_idx) {
  if (!w.real) return -1;
  const size_t slot = static_cast<size_t>(w.pil) * k.T + w.t;
  return kept_idx ? __ldg(kept_idx + slot) : static_cast<int>(slot);
}
// request the point of this lane's row and its pillar's coordinates into the lane's staging slot
__device__ __forceinline__ void gather_issue(const Kargs &k, int src, int pil, bool inwin, const float *__restrict__ rows_src,
                                             const int *__restrict__ coors, uint32_t slot) {
  if (src >= 0) {
    const float *pp = rows_src + static_cast<size_t>(src) * k.C;
    if (k.C == 4) {
      cp_async16(slot, pp);
    } else {
#pragma unroll
      for (int j = 0; j < MBEV_MAX_POINT_DIM; ++j)
        if (j < k.C) cp_async4(slot + 4u * j, pp + j);
    }
  }
  if (inwin) cp_async16(slot + 4u * MBEV_MAX_POINT_DIM, reinterpret_cast<const int4 *>(coors) + pil);
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void gather_take(const Kargs &k, const Window &w, const float *stage, Gathered &g) {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int j = 0; j < MBEV_MAX_POINT_DIM; ++j) g.pv[j] = 0.f;
  if (w.real) {
    const float4 v = *reinterpret_cast<const float4 *>(stage);
    g.pv[0] = v.x; g.pv[1] = v.y; g.pv[2] = v.z; g.pv[3] = v.w;
    if (k.C > 4) {
      const float4 u = *reinterpret_cast<const float4 *>(stage + 4);
      g.pv[4] = u.x; g.pv[5] = u.y; g.pv[6] = u.z; g.pv[7] = u.w;
#pragma unroll
      for (int j = 4; j < MBEV_MAX_POINT_DIM; ++j)
        if (j >= k.C) g.pv[j] = 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j >= k.C) g.pv[j] = 0.f;
    }
  }
  g.cc = make_int4(0, 0, 0, 0);
  if (w.inwin) g.cc = *reinterpret_cast<const int4 *>(stage + MBEV_MAX_POINT_DIM);
}

// decorate the gathered point (mmdet3d PillarFeatureNet.forward) into the layer-0 input row (wide column order: every
// index below is a compile-time constant, so the row lives in registers) and park it in the lane's staging slot
__device__ __forceinline__ void decorate_x0(const Kargs &k, const Window &w, const Gathered &g, float *stage) {
  const float (&pv)[MBEV_MAX_POINT_DIM] = g.pv;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (w.inwin) {
    const int4 cc = g.cc;
    // upstream: coors.type_as(features) * vx + x_offset — float32 multiply THEN add (no FMA contraction)
    cx = __fadd_rn(__fmul_rn(static_cast<float>(cc.w), k.vx), k.xo);
    cy = __fadd_rn(__fmul_rn(static_cast<float>(cc.z), k.vy), k.yo);
    cz = __fadd_rn(__fmul_rn(static_cast<float>(cc.y), k.vz), k.zo);
  }
  // cluster mean: the pillar's points re-added in slot order by every lane of the pillar, / num_points
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int tt = 0; tt < w.maxlen; ++tt) {
    const float vx = __shfl_sync(0xffffffffu, pv[0], (w.s0 + tt) & 31);
    const float vy = __shfl_sync(0xffffffffu, pv[1], (w.s0 + tt) & 31);
    const float vz = __shfl_sync(0xffffffffu, pv[2], (w.s0 + tt) & 31);
    if (tt < w.n) {
      sx = __fadd_rn(sx, vx);
      sy = __fadd_rn(sy, vy);
      sz = __fadd_rn(sz, vz);
    }
  }
  float xd[kK0Pad];
#pragma unroll
  for (int d = 0; d < kK0Pad; ++d) xd[d] = 0.f;  // virtual rows, window padding and the unused slots
  if (w.real) {
    const float fn = static_cast<float>(w.n);
    const float mx = __fdiv_rn(sx, fn), my = __fdiv_rn(sy, fn), mz = __fdiv_rn(sz, fn);
    const float x = pv[0], y = pv[1], z = pv[2];
    const float ex = __fsub_rn(x, cx), ey = __fsub_rn(y, cy), ez = __fsub_rn(z, cz);
    const bool alias = k.vcenter && k.legacy;  // legacy: centre offset written in place over xyz
    const float r0 = alias ? ex : x, r1 = alias ? ey : y, r2 = (alias && k.vcd > 2) ? ez : z;  // 2-channel centre: z stays raw
    xd[0] = r0;
    xd[1] = r1;
    xd[2] = r2;
#pragma unroll
    for (int j = 3; j < MBEV_MAX_POINT_DIM; ++j) xd[kWideExtra + j - 3] = pv[j];  // zero beyond C (and a zero weight column)
    if (k.cluster) {
      xd[kWideCluster + 0] = __fsub_rn(x, mx);
      xd[kWideCluster + 1] = __fsub_rn(y, my);
      xd[kWideCluster + 2] = __fsub_rn(z, mz);
    }
    if (k.vcenter) {
      xd[kWideCentre + 0] = ex;
      xd[kWideCentre + 1] = ey;
      if (k.vcd > 2) xd[kWideCentre + 2] = ez;
    }
    if (k.dist) xd[kWideDist] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r0, r0), __fmul_rn(r1, r1)), __fmul_rn(r2, r2)));
  }
  float4 *st4 = reinterpret_cast<float4 *>(stage);
#pragma unroll
  for (int q = 0; q < 4; ++q) st4[q] = make_float4(xd[4 * q], xd[4 * q + 1], xd[4 * q + 2], xd[4 * q + 3]);
}
// the parked layer-0 input row -> split -> tensor memory
__device__ __forceinline__ void store_x0(const float *stage, uint32_t t_hi, uint32_t t_lo) {
  const float4 *st4 = reinterpret_cast<const float4 *>(stage);
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = st4[q];
    split_tf32_alu(v.x, hi[4 * q + 0], lo[4 * q + 0]);
    split_tf32_alu(v.y, hi[4 * q + 1], lo[4 * q + 1]);
    split_tf32_alu(v.z, hi[4 * q + 2], lo[4 * q + 2]);
    split_tf32_alu(v.w, hi[4 * q + 3], lo[4 * q + 3]);
  }
  tmem_st16(t_hi, hi);
  tmem_st16(t_lo, lo);
  tc_wait_st();
}

// D batch -> folded BatchNorm + ReLU
__device__ __forceinline__ void load_bn_relu(uint32_t taddr, const float *sc, const float *sh, float (&a)[16]) {
  uint32_t v[16];
  tmem_ld16(taddr, v);
  tc_wait_ld();
  const float4 *sc4 = reinterpret_cast<const float4 *>(sc), *sh4 = reinterpret_cast<const float4 *>(sh);
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 c = sc4[j4], s = sh4[j4];
    a[4 * j4 + 0] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 0]), c.x, s.x), 0.f);
    a[4 * j4 + 1] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 1]), c.y, s.y), 0.f);
    a[4 * j4 + 2] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 2]), c.z, s.z), 0.f);
    a[4 * j4 + 3] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 3]), c.w, s.w), 0.f);
  }
}

__device__ __forceinline__ void load_bn_relu32(uint32_t taddr, const float *sc, const float *sh, float (&a)[32]) {
  uint32_t v[32];
  tmem_ld32(taddr, v);
  tc_wait_ld();
  const float4 *sc4 = reinterpret_cast<const float4 *>(sc), *sh4 = reinterpret_cast<const float4 *>(sh);
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 c = sc4[j4], s = sh4[j4];
    a[4 * j4 + 0] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 0]), c.x, s.x), 0.f);
    a[4 * j4 + 1] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 1]), c.y, s.y), 0.f);
    a[4 * j4 + 2] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 2]), c.z, s.z), 0.f);
    a[4 * j4 + 3] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 3]), c.w, s.w), 0.f);
  }
}

__device__ __forceinline__ void split_store16(uint32_t t_hi, uint32_t t_lo, const float (&a)[16]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) split_tf32_alu(a[j], hi[j], lo[j]);
  tmem_st16(t_hi, hi);
  tmem_st16(t_lo, lo);
}

// Per-pillar max over the lanes of the window as a CYCLIC DOUBLING all-reduce: in round d every lane takes the max with
// the lane d rows further down its own pillar, wrapping to the pillar's first row — max is idempotent, so after
// ceil(log2(len)) rounds every lane of the pillar holds the max over the whole pillar, for any mix of pillar lengths in
// one warp (a lane whose pillar is not longer than d reads itself). No select, no broadcast round: one shuffle + one
// max per register and round. N registers are scanned together so that N independent shuffles are in flight.
__device__ __forceinline__ int cyc_src(const Window &w, int lane, int d) {
  const int len = w.s1 - w.s0 + 1;
  const int t2 = w.t + d;
  return d < len ? (t2 < len ? lane + d : lane + d - len) : lane;
}
template <int N>
__device__ __forceinline__ void seg_allmax(const Window &w, int lane, float (&a)[N]) {
  for (int d = 1; d < w.maxlen; d <<= 1) {
    const int src = cyc_src(w, lane, d);
#pragma unroll
    for (int j = 0; j < N; ++j) a[j] = fmaxf(a[j], __shfl_sync(0xffffffffu, a[j], src));
  }
}
// bf16 form: rounding to bf16 is monotonic, so max(bf16(a)) == bf16(max(a)) — the all-reduce can run on the PACKED
// registers that go to tensor memory anyway: half the shuffles and half the max instructions.
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
template <int N>
__device__ __forceinline__ void seg_allmax_bf16(const Window &w, int lane, uint32_t (&v)[N]) {
  for (int d = 1; d < w.maxlen; d <<= 1) {
    const int src = cyc_src(w, lane, d);
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = max_bf16x2(v[j], __shfl_sync(0xffffffffu, v[j], src));
  }
}

// Register budget: 20 warps = 5 on each of the four SM sub-partitions (16 384 registers each), so the kernel starts
// with 96 registers per thread; the last warpgroup (MMA issuers: uniform-register work only) then hands its registers
// back and the epilogue warpgroups grow to 112 (setmaxnreg): 4 x 32 x 112 + 32 x 24 per sub-partition; the grow must fit into what the shrink released: 128 x 72 >= 512 x 16.
#ifndef MBEV_W2_REGS_LAUNCH
#define MBEV_W2_REGS_LAUNCH 96
#define MBEV_W2_REGS_EPI 112
#endif
constexpr int kW2RegsLaunch = MBEV_W2_REGS_LAUNCH, kW2RegsEpi = MBEV_W2_REGS_EPI, kW2RegsIssuer = 24;
template <bool kBf16>
__global__ void __maxnreg__(kW2RegsLaunch)
k_pfn_tcw2(const float *__restrict__ rows_src, const int *__restrict__ kept_idx, const int *__restrict__ num_points_nat,
           int *np_sorted, int *order, const int *__restrict__ coors, const int *__restrict__ bounds8,
           float *__restrict__ feats, const __grid_constant__ Kargs k) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float *s_ss = reinterpret_cast<float *>(smem_raw + k.o_ss);     // [L][2][128]
  int *s_live = reinterpret_cast<int *>(smem_raw + k.o_tab);      // [2 sets][4] live windows of chunk c (ring)
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + k.o_bar);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + kW2NumBars);
  const uint32_t bar0 = smem_u32(s_bar);
  const uint32_t bar_w = bar0 + 8 * kW2BarW;
  const uint32_t smem_base = smem_u32(smem_raw);
  const int L = k.L;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < kW2Sets; ++s) {
      const uint32_t b = bar0 + 8 * (kW2BarSet + kW2BarsPerSet * s);
      mbar_init(b + 8 * kW2X0, 8);
      mbar_init(b + 8 * kW2D, 1);
      mbar_init(b + 8 * kW2XC, 1);
      mbar_init(b + 8 * kW2EVX, 8);
      mbar_init(b + 8 * kW2EVM, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 4 * kW2Sets) s_live[tid] = 0;
  __syncwarp();
  if (warp == kW2EpiWarps) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  for (int l = 0; l < L; ++l) {
    for (int i = tid; i < k.U[l]; i += kW2Threads) {
      s_ss[(2 * l) * MBEV_MAX_UNITS + i] = __ldg(k.scale[l] + i);
      s_ss[(2 * l + 1) * MBEV_MAX_UNITS + i] = __ldg(k.shift[l] + i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp >= kW2EpiWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kW2RegsIssuer));
  }
  if (warp >= kW2EpiWarps + kW2Sets) {
    // idle warps of the issuer warpgroup
  } else if (warp >= kW2EpiWarps) {
    // =========================================== MMA issuer of set `set` =====================================
    // Every operand of tcgen05.mma below is WARP-UNIFORM and provably so for the compiler (kernel parameters, a
    // constant-lane shuffle of the warp index and of the TMEM base, uniform loop counters), and the issuing thread is
    // chosen by elect.sync: the descriptors then live in uniform registers. With `if (lane == 0)` and per-thread
    // operands ptxas wraps every single MMA in an ELECT / 6 x R2UR / branch waterfall (~90 cycles per instruction —
    // three times the tensor pipe's own 32 cycles for N = 64).
    const int set = __shfl_sync(0xffffffffu, warp - kW2EpiWarps, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    if (set == 0) {
      if (elect_one()) {
        mbar_expect_tx(bar_w, k.w_bytes);  // weights: global image -> shared memory, resident, shared by both sets
        for (uint32_t off = 0; off < k.w_bytes; off += 32768u)
          bulk_g2s(smem_base + off, reinterpret_cast<const char *>(k.w_img) + off, min(32768u, k.w_bytes - off), bar_w);
      }
    }
    __syncwarp();
    mbar_wait(bar_w, 0);
    const uint32_t bs = bar0 + 8 * (kW2BarSet + kW2BarsPerSet * set);
    const uint32_t t_ah = tmem_u + kW2SetCols * set, t_al = t_ah + 64, t_d = t_ah + 128;
    uint32_t par_x0 = 0, par_evx = 0, par_evm = 0;
    for (int c = 0;; ++c) {
      mbar_wait(bs + 8 * kW2X0, par_x0);
      par_x0 ^= 1u;
      if (*reinterpret_cast<volatile int *>(s_live + 4 * set + (c & 3)) == 0) break;
      tc_fence_after();
      MBEV_TR(0);
      for (int l = 0; l < L; ++l) {
        const int U = k.U[l];
        const bool half = kBf16 && l > 0;  // this layer's operands are bf16: one MMA per 16 K, no lo image
        const uint32_t idesc = half ? make_idesc_bf16(U) : make_idesc(U);
        const uint32_t lbo = static_cast<uint32_t>(U) * 16u;
        const uint64_t dh0 = make_bdesc(smem_base + k.w_off[l][0], lbo, 128u);
        const uint64_t dl0 = make_bdesc(smem_base + k.w_off[l][1], lbo, 128u);
        const uint32_t dstep = lbo >> 3;  // one K-step (two 16-byte K-chunks) in the descriptor address field (low word)
        const int nks = (l == 0) ? (k.Kp[0] >> 3) : (k.U[l - 1] >> (half ? 4 : 3));
        const int nparts = (l == 0) ? 1 : 2;
        uint32_t acc = 0;
        for (int part = 0; part < nparts; ++part) {  // x K-half, then the replicated-max K-half through the same A columns
          if (l > 0) {
            if (part == 0) {
              mbar_wait(bs + 8 * kW2EVX, par_evx);
              par_evx ^= 1u;
            } else {
              mbar_wait(bs + 8 * kW2EVM, par_evm);
              par_evm ^= 1u;
            }
            tc_fence_after();
          }
          MBEV_TR(1 + 4 * l + 2 * part);
          if (elect_one()) {
            const uint32_t koff = dstep * static_cast<uint32_t>(part * nks);
            if (half) {
#pragma unroll 2
              for (int j = 0; j < nks; ++j) {
                mma_bf16_ts(t_d, t_ah + 8u * j, dh0 + (koff + dstep * j), idesc, acc);
                acc = 1;
              }
            } else {
#pragma unroll 2
              for (int j = 0; j < nks; ++j) {  // al*wh, ah*wl, ah*wh : small terms first
                const uint64_t dh = dh0 + (koff + dstep * j), dl = dl0 + (koff + dstep * j);
                mma_tf32_ts(t_d, t_al + 8u * j, dh, idesc, acc);
                mma_tf32_ts(t_d, t_ah + 8u * j, dl, idesc, 1u);
                mma_tf32_ts(t_d, t_ah + 8u * j, dh, idesc, 1u);
                acc = 1;
              }
            }
            tc_commit(bs + 8 * ((l > 0 && part == 0) ? kW2XC : kW2D));
          }
          __syncwarp();
          MBEV_TR(1 + 4 * l + 2 * part + 1);
        }
      }
    }
  } else {
    // =========================================== epilogue warps ===============================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kW2RegsEpi));
    const int set = warp >> 3, quad = warp & 3, h = (warp >> 2) & 1;
    const uint32_t tl = tmem + (static_cast<uint32_t>(quad * 32) << 16) + kW2SetCols * set;  // lane quadrant, set columns
    const uint32_t t_ah = tl, t_al = tl + 64, t_d = tl + 128;
    const uint32_t bs = bar0 + 8 * (kW2BarSet + kW2BarsPerSet * set);
    const int sub = 8 * blockIdx.x + 4 * set + quad;
    int cursor = __ldg(bounds8 + sub);
    const int pend = __ldg(bounds8 + sub + 1);
    uint32_t par_d = 0, par_x0 = 0, par_xc = 0;
    int *live = s_live + 4 * set;

    // prologue of the gather pipeline: window 0 and its points, num_points of window 1
    sort_subrange(num_points_nat, cursor, pend, k.T, lane, reinterpret_cast<int *>(smem_raw + k.o_scr) + 4 * kW2Sets * 32 * kStageFloats + warp * kSortClasses,
                  order, np_sorted);
    const int *num_points = np_sorted;
    Window w = pack_window(load_np(num_points, cursor, pend, lane), load_ord(order, cursor, pend, lane), cursor, pend, k.T, lane);
    int cnext = cursor + w.cnt;
    int np_next = load_np(num_points, cnext, pend, lane), ord_next = load_ord(order, cnext, pend, lane);
    float *stage = reinterpret_cast<float *>(smem_raw + k.o_scr) + ((4 * set + quad) * 32 + lane) * kStageFloats;
    const uint32_t stage_u = smem_u32(stage);
    if (h == 0) {
      gather_issue(k, gather_index(k, w, kept_idx), w.pil, w.inwin, rows_src, coors, stage_u);
      Gathered g;
      gather_take(k, w, stage, g);
      decorate_x0(k, w, g, stage);
    }

    for (int c = 0;; ++c) {
      MBEV_TR(0);
      MBEV_TR(1);
      if (h == 0) store_x0(stage, t_ah, t_al);
      MBEV_TR(2);
      // ---- set rendezvous: every window's layer-0 input is in TMEM, nobody reads the previous chunk's D ------
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if ((warp & 7) == 0) live[(c + 2) & 3] = 0;  // everybody read this slot two rendezvous ago
        if (h == 0 && w.cnt > 0) atomicAdd(live + (c & 3), 1);
        mbar_arrive(bs + 8 * kW2X0);
      }
      mbar_wait(bs + 8 * kW2X0, par_x0);
      par_x0 ^= 1u;
      if (*reinterpret_cast<volatile int *>(live + (c & 3)) == 0) break;
      MBEV_TR(3);
      Window wn;
      int cnn = cnext, src_n = -1;

      for (int l = 0; l < L; ++l) {
        const int U = k.U[l];
        const bool last = (l == L - 1);
        // The next chunk's input is prepared in the three places where this warp would otherwise only wait for MMAs:
        // (A) window + kept_idx entries, (B) point / coordinate requests, (C) decoration into the staging slot
        if (l == max(L - 3, 0)) {
          // next window (its num_points were requested a chunk ago), the kept_idx entries of its rows, num_points of the
          // window after it
          wn = pack_window(np_next, ord_next, cnext, pend, k.T, lane);
          cnn = cnext + wn.cnt;
          np_next = load_np(num_points, cnn, pend, lane);
          ord_next = load_ord(order, cnn, pend, lane);
          if (h == 0) src_n = gather_index(k, wn, kept_idx);
        }
        if (h == 0) {
          if (l == max(L - 2, 0)) gather_issue(k, src_n, wn.pil, wn.inwin, rows_src, coors, stage_u);
          if (last) {
            Gathered g;
            gather_take(k, wn, stage, g);
            decorate_x0(k, wn, g, stage);
          }
        }
        mbar_wait(bs + 8 * kW2D, par_d);
        par_d ^= 1u;
        tc_fence_after();
        MBEV_TR(4 + 5 * l);
        const int Uh = U >> 1;
        const int nbat = Uh >> 4;
        const float *sc = s_ss + (2 * l) * MBEV_MAX_UNITS + h * Uh, *sh = s_ss + (2 * l + 1) * MBEV_MAX_UNITS + h * Uh;
        if (last) {
          // two 16-column batches per iteration: 32 independent registers through the all-reduce, 128 contiguous bytes
          // per pillar and store
#pragma unroll 1
          for (int b = 0; b + 1 < nbat; b += 2) {
            const int col0 = h * Uh + 16 * b;
            float a[32];
            load_bn_relu32(t_d + static_cast<uint32_t>(col0), sc + 16 * b, sh + 16 * b, a);
            seg_allmax<32>(w, lane, a);
            if (w.inwin && lane == w.s1) {
              float4 *out = reinterpret_cast<float4 *>(feats + static_cast<size_t>(w.pil) * U + col0);
#pragma unroll
              for (int q = 0; q < 8; ++q) out[q] = make_float4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
            }
          }
          if (nbat & 1) {
            const int col0 = h * Uh + 16 * (nbat - 1);
            float a[16];
            load_bn_relu(t_d + static_cast<uint32_t>(col0), sc + 16 * (nbat - 1), sh + 16 * (nbat - 1), a);
            seg_allmax<16>(w, lane, a);
            if (w.inwin && lane == w.s1) {
              float4 *out = reinterpret_cast<float4 *>(feats + static_cast<size_t>(w.pil) * U + col0);
#pragma unroll
              for (int q = 0; q < 4; ++q) out[q] = make_float4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
            }
          }
        } else {
          // non-last layers have U <= 64: at most two 16-column batches per warp, kept in registers across the
          // "x consumed" wait
          const uint32_t c0 = static_cast<uint32_t>(h * Uh);
          if (kBf16) {
            // next layer's operands are bf16: pack once, store the x half, scan the packed registers, store the max half
            const uint32_t ca = c0 >> 1;  // A column of K element c0 (two bf16 per column)
            float a[16];
            uint32_t v0[8], v1[8];
            load_bn_relu(t_d + c0, sc, sh, a);
            pack_bf16x16(a, v0);
            tmem_st8(t_ah + ca, v0);
            if (nbat > 1) {
              load_bn_relu(t_d + c0 + 16, sc + 16, sh + 16, a);
              pack_bf16x16(a, v1);
              tmem_st8(t_ah + ca + 8, v1);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bs + 8 * kW2EVX);
            if (nbat > 1) {  // runs while the x-part MMAs do
              uint32_t v[16];
#pragma unroll
              for (int j = 0; j < 8; ++j) { v[j] = v0[j]; v[8 + j] = v1[j]; }
              seg_allmax_bf16<16>(w, lane, v);
#pragma unroll
              for (int j = 0; j < 8; ++j) { v0[j] = v[j]; v1[j] = v[8 + j]; }
            } else {
              seg_allmax_bf16<8>(w, lane, v0);
            }
            mbar_wait(bs + 8 * kW2XC, par_xc);  // the x-part MMAs have read A: its columns are free again
            par_xc ^= 1u;
            tc_fence_after();
            tmem_st8(t_ah + ca, v0);  // max half: K index = U + unit index, same A columns
            if (nbat > 1) tmem_st8(t_ah + ca + 8, v1);
          } else {
            float a0[16], a1[16];
            load_bn_relu(t_d + c0, sc, sh, a0);
            split_store16(t_ah + c0, t_al + c0, a0);  // x half: K index = unit index
            if (nbat > 1) {
              load_bn_relu(t_d + c0 + 16, sc + 16, sh + 16, a1);
              split_store16(t_ah + c0 + 16, t_al + c0 + 16, a1);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bs + 8 * kW2EVX);
            MBEV_TR(4 + 5 * l + 1);
            if (nbat > 1) {  // runs while the x-part MMAs do
              float a[32];
#pragma unroll
              for (int j = 0; j < 16; ++j) { a[j] = a0[j]; a[16 + j] = a1[j]; }
              seg_allmax<32>(w, lane, a);
#pragma unroll
              for (int j = 0; j < 16; ++j) { a0[j] = a[j]; a1[j] = a[16 + j]; }
            } else {
              seg_allmax<16>(w, lane, a0);
            }
            MBEV_TR(4 + 5 * l + 2);
            mbar_wait(bs + 8 * kW2XC, par_xc);  // the x-part MMAs have read A: its columns are free again
            par_xc ^= 1u;
            tc_fence_after();
            MBEV_TR(4 + 5 * l + 3);
            split_store16(t_ah + c0, t_al + c0, a0);  // max half: K index = U + unit index, same A columns
            if (nbat > 1) split_store16(t_ah + c0 + 16, t_al + c0 + 16, a1);
          }
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bs + 8 * kW2EVM);
          MBEV_TR(4 + 5 * l + 4);
        }
      }
      MBEV_TR(20);
      w = wn;
      cnext = cnn;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kW2EpiWarps) tmem_dealloc(tmem, kTmemCols);
}

// does the stack fit k_pfn_tcw2 (T <= 32, every U a multiple of 32, non-last layers <= 64 units), and its
// shared-memory plan: [weight image][scale / shift][live ring][gather staging][barriers + TMEM slot]
inline bool tcw2_plan(Kargs &k) {
  if (k.T > 32) return false;
  for (int l = 0; l < k.L; ++l)
    if (k.U[l] % 32) return false;
  for (int l = 0; l + 1 < k.L; ++l)
    if (k.U[l] > 64) return false;
  uint32_t o = (k.w_bytes + 127u) & ~127u;
  k.o_ss = o; o += k.L * 2 * MBEV_MAX_UNITS * 4;
  k.o_tab = o; o += 4 * kW2Sets * 4;
  o = (o + 15u) & ~15u;
  k.o_scr = o; o += 4 * kW2Sets * 32 * kStageFloats * 4;  // gather staging: one 64-byte slot per row of every window
  o += kW2EpiWarps * kSortClasses * 4;                  // class counters of the sub-range sort, one set per epilogue warp
  k.o_bar = o; o += 8 * kW2NumBars + 16;
  k.smem_bytes = static_cast<int>(o);
  return k.smem_bytes <= kSmemLimit;
}

// FULL forward: the two-pipeline warp-local kernel when it applies, else (and for every STATS launch) the block-level
// kernel. The opt-in to > 48 KB of dynamic shared memory is per device and cheap, so it is set on every launch (no
// process-global "done" flag: a second GPU in the same process needs it too).
inline int launch(const Plan &pl, const float *rows, const int32_t *kept_idx, const int32_t *num_points,
                  const int32_t *coors, float *feats, int stat_layer, cudaStream_t stream) {
  Kargs k = pl.k;
  k.stat_layer = stat_layer;
  Kargs k2 = k;
  if (k.bf16) {  // single-pass bf16 layers: k_pfn_tcw2 only, eval-mode only (the STATS launches are 3xTF32 images)
    if (stat_layer >= 0 || !tcw2_plan(k2)) return MBEV_ERR_UNSUPPORTED;
    MBEV_CUDA(cudaFuncSetAttribute(k_pfn_tcw2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    k_pfn_tcw2<true><<<pl.grid, kW2Threads, k2.smem_bytes, stream>>>(rows, kept_idx, num_points, pl.np_sorted, pl.order, coors, pl.bounds, feats, k2);
  } else if (stat_layer < 0 && tcw2_plan(k2)) {
    MBEV_CUDA(cudaFuncSetAttribute(k_pfn_tcw2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    k_pfn_tcw2<false><<<pl.grid, kW2Threads, k2.smem_bytes, stream>>>(rows, kept_idx, num_points, pl.np_sorted, pl.order, coors, pl.bounds, feats, k2);
  } else {
    MBEV_CUDA(cudaFuncSetAttribute(k_pfn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    k_pfn_tc<<<pl.grid, kThreads, k.smem_bytes, stream>>>(rows, kept_idx, num_points, coors, pl.bounds, feats, k);
  }
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

}  // namespace tc
}  // namespace mbev
