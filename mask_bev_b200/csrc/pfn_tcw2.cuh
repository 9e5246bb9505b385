// K2-TCW2 — two chunk pipelines per SM (the FULL-forward tcgen05 kernel when T <= 32).
//
// k_pfn_tcw keeps one chunk in flight per SM: the layer input A (K = 2U = 128 columns, hi + lo = 256) plus the
// accumulators fill the 512 TMEM columns, so the tensor pipe idles while the epilogue warps work and vice versa
// (cycle stamps: ~3.2 k cycles of epilogue and ~3.2 k of exposed MMA per layer). Here K is split in TIME: the
// x half of the next layer's input goes through the A columns first, and once those MMAs have retired
// (tcgen05.commit -> "x consumed") the replicated-max half is written into the SAME columns and accumulated on
// top. A chunk then needs A_hi 64 + A_lo 64 + D 128 = 256 columns, and two independent chunk pipelines ("sets")
// share the SM: 2 x (8 epilogue warps + 1 MMA-issuer warp). While one set waits for its MMAs the other one runs
// its epilogue, so the two pipes overlap without any cross-set synchronisation. The segmented max of a layer is
// computed while its x-part MMAs run.
//
// Everything else is k_pfn_tcw: 32-row windows per TMEM lane quadrant, warp-local segmented max by shuffles,
// 3xTF32, weights resident in shared memory (shared by both sets), row-balanced sub-ranges (8 per CTA).
#pragma once
#include <cstring>

#include "pfn_tcw.cuh"

namespace mbev {
namespace tc {

constexpr int kW2Sets = 2;
constexpr int kW2EpiWarps = 8 * kW2Sets;               // per set: 4 quadrants x 2 column halves
constexpr int kW2Threads = (kW2EpiWarps + kW2Sets) * 32;
// canvas form (K2 + K3 in one kernel): 1024 threads = 16 epilogue warps | 2 issuers + 14 canvas writers. The CTA
// launches at 64 registers per thread; setmaxnreg then moves registers inside that allocation (it cannot take more
// from the SM): epilogue warpgroups 96, the others 32 — 512 x 96 + 512 x 32 = 1024 x 64. One warp sustains only a few
// bytes per clock of streaming stores (measured: 2 writer warps per SM -> 1.6 TB/s, 8 -> 3.4 TB/s), so every warp
// the register file can hold next to the epilogue is a writer.
constexpr int kCvWriterWarp0 = kW2EpiWarps + kW2Sets;
constexpr int kCvWriters = 14;
constexpr int kCvThreads = (kCvWriterWarp0 + kCvWriters) * 32;
constexpr int kCvRegsEpi = 96, kCvRegsRest = 32;
constexpr int kCvRun = 4;       // strips per writer claim: zeros go out as 2 KB bulk copies per plane
constexpr int kCvZeroBytes = kCvRun * 128 * 4;  // shared-memory zero source of the bulk copies
constexpr int kCvStrip = 128;  // cells per strip: one warp-wide float4 store covers a strip of one plane (512 B)

// Canvas side of the fused kernel. Cells are numbered globally, gc = b * G + y * nx + x; a STRIP is 128 consecutive
// global cells. Sub-range `s` (8 per CTA) owns the strips [sb[s], sb[s+1]) and the pillars ord[pb[s] .. pb[s+1]) that
// live in them (ord lists the pillars in global cell order).
struct CanvasArgs {
  const int *table;   // (NC) cell -> pillar id, -1 empty
  const int2 *ord;    // (P) (pillar id, num_points) in cell order
  const int *sb;      // (8 * grid + 1) first strip of each sub-range
  const int *pb;      // (8 * grid + 1) first ord index of each sub-range
  const int *spre;    // (NS) occupied cells (= ord index) before each strip
  float *canvas;      // (B, C_out, G)
  int G, NC, Cout;
};
constexpr int kW2SetCols = 256;                         // TMEM columns per set: A_hi [0,64) A_lo [64,128) D [128,256)
constexpr int kW2BarW = 0, kW2BarSet = 1, kW2BarsPerSet = 5, kW2NumBars = kW2BarSet + kW2Sets * kW2BarsPerSet;
enum { kW2X0 = 0, kW2D = 1, kW2XC = 2, kW2EVX = 3, kW2EVM = 4 };

struct Window {
  int cnt, nrows;  // pillars / rows packed into this window
  int pil;         // global pillar of this lane's row
  int s0, s1, t, n, maxlen;
  bool inwin, real;
};

// pack whole pillars cursor, cursor+1, ... into a 32-row window (lane r = row r); identical on every warp that
// calls it with the same cursor
// `ord` (optional): walk order, ord[j] = (pillar id, num_points) — the canvas kernel walks pillars in CELL order
__device__ __forceinline__ Window pack_window(const int *__restrict__ num_points, const int2 *__restrict__ ord, int cursor,
                                              int pend, int T, int lane) {
  Window w;
  const bool cand = cursor + lane < pend;
  int np = 0, pid = cursor + lane;
  if (cand) {
    if (ord) {
      const int2 o = __ldg(ord + cursor + lane);
      pid = o.x;
      np = o.y;
    } else {
      np = __ldg(num_points + cursor + lane);
    }
  }
  const int need = cand ? np + (np < T ? 1 : 0) : 0;
  int incl = need;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  w.cnt = __popc(__ballot_sync(0xffffffffu, cand && incl <= 32));  // prefix-closed
  w.nrows = w.cnt ? __shfl_sync(0xffffffffu, incl, w.cnt - 1) : 0;
  const int excl = incl - need;
  int pi = 0;  // the pillar of row `lane` is the last i < cnt with excl_i <= lane
#pragma unroll
  for (int bit = 16; bit; bit >>= 1) {
    const int j = pi + bit;
    const int e = __shfl_sync(0xffffffffu, excl, j & 31);
    if (j < w.cnt && e <= lane) pi = j;
  }
  w.inwin = lane < w.nrows;
  w.s0 = __shfl_sync(0xffffffffu, excl, pi);
  int nd = __shfl_sync(0xffffffffu, need, pi);
  w.n = __shfl_sync(0xffffffffu, np, pi);
  if (!w.inwin) {  // window padding: a segment of its own
    w.s0 = lane;
    nd = 1;
  }
  w.s1 = w.s0 + nd - 1;
  w.t = lane - w.s0;
  w.real = w.inwin && w.t < w.n;
  w.pil = ord ? __shfl_sync(0xffffffffu, pid, pi) : cursor + pi;
  int ml = w.inwin ? nd : 1;
#pragma unroll
  for (int o = 16; o; o >>= 1) ml = max(ml, __shfl_xor_sync(0xffffffffu, ml, o));
  w.maxlen = ml;
  return w;
}

// gather this row's point, decorate (mmdet3d PillarFeatureNet.forward), split and store the layer-0 input row
__device__ __forceinline__ void build_x0(const Kargs &k, const Window &w, const float *__restrict__ rows_src,
                                         const int *__restrict__ kept_idx, const int *__restrict__ coors, float *xd,
                                         uint32_t t_hi, uint32_t t_lo) {
  float pv[MBEV_MAX_POINT_DIM];
#pragma unroll
  for (int j = 0; j < MBEV_MAX_POINT_DIM; ++j) pv[j] = 0.f;
  if (w.real) {
    const size_t slot = static_cast<size_t>(w.pil) * k.T + w.t;
    const int src = kept_idx ? __ldg(kept_idx + slot) : static_cast<int>(slot);
    const float *pp = rows_src + static_cast<size_t>(src) * k.C;
    if (k.C == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(pp));
      pv[0] = v.x; pv[1] = v.y; pv[2] = v.z; pv[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < MBEV_MAX_POINT_DIM; ++j)
        if (j < k.C) pv[j] = __ldg(pp + j);
    }
  }
  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (w.inwin) {
    const int4 cc = __ldg(reinterpret_cast<const int4 *>(coors) + w.pil);  // (b, z, y, x)
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
#pragma unroll
  for (int d = 0; d < kK0Pad; ++d) xd[d] = 0.f;  // virtual rows, window padding and the K padding
  if (w.real) {
    const float fn = static_cast<float>(w.n);
    const float mx = __fdiv_rn(sx, fn), my = __fdiv_rn(sy, fn), mz = __fdiv_rn(sz, fn);
    const float x = pv[0], y = pv[1], z = pv[2];
    const float ex = __fsub_rn(x, cx), ey = __fsub_rn(y, cy), ez = __fsub_rn(z, cz);
    const bool alias = k.vcenter && k.legacy;  // legacy: centre offset written in place over xyz
    const float r0 = alias ? ex : x, r1 = alias ? ey : y, r2 = (alias && k.vcd > 2) ? ez : z;  // 2-channel centre: z stays raw
    int d = 0;
    xd[d++] = r0;
    xd[d++] = r1;
    xd[d++] = r2;
#pragma unroll
    for (int j = 3; j < MBEV_MAX_POINT_DIM; ++j)
      if (j < k.C) xd[d++] = pv[j];
    if (k.cluster) {
      xd[d++] = __fsub_rn(x, mx);
      xd[d++] = __fsub_rn(y, my);
      xd[d++] = __fsub_rn(z, mz);
    }
    if (k.vcenter) {
      xd[d++] = ex;
      xd[d++] = ey;
      if (k.vcd > 2) xd[d++] = ez;
    }
    if (k.dist) xd[d++] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r0, r0), __fmul_rn(r1, r1)), __fmul_rn(r2, r2)));
  }
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 v = *reinterpret_cast<const float4 *>(xd + 4 * j4);
    split_tf32_alu(v.x, hi[4 * j4 + 0], lo[4 * j4 + 0]);
    split_tf32_alu(v.y, hi[4 * j4 + 1], lo[4 * j4 + 1]);
    split_tf32_alu(v.z, hi[4 * j4 + 2], lo[4 * j4 + 2]);
    split_tf32_alu(v.w, hi[4 * j4 + 3], lo[4 * j4 + 3]);
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

__device__ __forceinline__ void split_store16(uint32_t t_hi, uint32_t t_lo, const float (&a)[16]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) split_tf32_alu(a[j], hi[j], lo[j]);
  tmem_st16(t_hi, hi);
  tmem_st16(t_lo, lo);
}

// per-pillar max over the lanes of the window: segmented inclusive max-scan, then read the pillar's last lane
// (kBroadcast = false: only the pillar's LAST lane holds the result — enough for the last layer's store)
template <bool kBroadcast = true>
__device__ __forceinline__ void seg_max16(const Window &w, int lane, float (&a)[16]) {
  for (int d = 1; d < w.maxlen; d <<= 1) {
    const bool take = lane - d >= w.s0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float o = __shfl_up_sync(0xffffffffu, a[j], d);
      a[j] = take ? fmaxf(a[j], o) : a[j];
    }
  }
  if (kBroadcast) {
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = __shfl_sync(0xffffffffu, a[j], w.s1);
  }
}

// ---- canvas writers ----------------------------------------------------------------------------------------------
// The 14 writer warps of a CTA share its 8 sub-ranges: a writer claims the next RUN (4 strips = 512 cells) of a
// sub-range (shared-memory ticket, preferring "its own" sub-range so that claims stay near every sub-range's compute
// frontier), and
//   1. zero-fills the run in all C_out planes with bulk-async copies (TMA engine, cp.async.bulk shared -> global, 2 KB
//      per plane from a zero buffer in shared memory; one lane issues them, no registers or LSU slots involved) —
//      no dependency on the PFN, this is ~97 % of the bytes;
//   2. for every strip of the run that holds pillars: waits until both column halves of all of them are in `feats`
//      (progress counters of the sub-range), then per pillar reads the 512-byte feature row coalesced (lane l =
//      channels 4l..4l+3) and drops the values into their planes (4-byte stores into sectors zero-filled a moment
//      ago: they merge in L2).
// At most 14 runs per CTA are zero-filled ahead of their features, so the merge window is microseconds.
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void put_pillar(const CanvasArgs &cv, const float *feats, const int p, const int gcell,
                                           const int lane) {
  if (4 * lane >= cv.Cout) return;
  const float4 v = __ldcg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(p) * cv.Cout) + lane);
  const size_t G = static_cast<size_t>(cv.G);
  const int b = gcell / cv.G;
  float *o = cv.canvas + (static_cast<size_t>(b) * cv.Cout + 4 * lane) * G + (gcell - b * cv.G);
  o[0] = v.x;
  o[G] = v.y;
  o[2 * G] = v.z;
  o[3 * G] = v.w;
}

__device__ __forceinline__ void canvas_writer(const CanvasArgs &cv, const float *feats, const volatile int *s_prog,
                                              int *s_next, const uint32_t zero_smem, const int wid, const int lane,
                                              const int dbg) {
  const size_t G = static_cast<size_t>(cv.G);
  const int *sb = cv.sb + 8 * blockIdx.x, *pb = cv.pb + 8 * blockIdx.x;
  for (;;) {
    int sub = -1, s_first = 0, s_last = 0;
    for (int a = 0; a < 8 && sub < 0; ++a) {  // next run of the preferred sub-range, else of the following ones
      const int sidx = (wid + a) & 7;
      const int s0 = __ldg(sb + sidx), s1 = __ldg(sb + sidx + 1);
      if (*reinterpret_cast<volatile int *>(s_next + sidx) * kCvRun >= s1 - s0) continue;  // exhausted (cheap pre-check)
      int idx = 0;
      if (lane == 0) idx = atomicAdd(s_next + sidx, 1);
      idx = __shfl_sync(0xffffffffu, idx, 0);
      if (s0 + kCvRun * idx < s1) {
        sub = sidx;
        s_first = s0 + kCvRun * idx;
        s_last = min(s_first + kCvRun, s1);
      }
    }
    if (sub < 0) break;
    // 1. zeros: [c0, c1) global cells, split at a frame boundary if the run straddles one
    if (lane == 0) {
      const int c0 = s_first * kCvStrip, c1 = min(s_last * kCvStrip, cv.NC);
      const int b0 = c0 / cv.G;
      const int cm = min(c1, (b0 + 1) * cv.G);  // end of the part inside frame b0
      float *o0 = cv.canvas + static_cast<size_t>(b0) * cv.Cout * G + (c0 - b0 * cv.G);
      float *o1 = cv.canvas + static_cast<size_t>(b0 + 1) * cv.Cout * G;
      for (int ch = 0; ch < cv.Cout; ++ch) {
        bulk_s2g(o0 + ch * G, zero_smem, static_cast<uint32_t>(cm - c0) * 4u);
        if (c1 > cm) bulk_s2g(o1 + ch * G, zero_smem, static_cast<uint32_t>(c1 - cm) * 4u);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    // 2. features
    bool waited = false;
    for (int strip = s_first; strip < s_last; ++strip) {
      const int gc = strip * kCvStrip + 4 * lane;
      const int4 pid = (gc < cv.NC) ? __ldg(reinterpret_cast<const int4 *>(cv.table + gc)) : make_int4(-1, -1, -1, -1);
      int cnt = __reduce_add_sync(0xffffffffu, (pid.x >= 0) + (pid.y >= 0) + (pid.z >= 0) + (pid.w >= 0));
      if (dbg & 16) cnt = 0;
      if (cnt == 0) continue;
      const int need = __ldg(cv.spre + strip) + cnt - __ldg(pb + sub);
      unsigned idle = 0;
      while (min(s_prog[2 * sub], s_prog[2 * sub + 1]) < need) {
        __nanosleep(100);
        if (++idle > (1u << 24)) return;  // watchdog (seconds): never hang the device on a broken partition
      }
      __threadfence_block();  // acquire: the feature rows behind the progress counters
      if (!waited) {          // the zeros of this run have landed before any feature is dropped on them
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        __syncwarp();
        waited = true;
      }
      const int g0 = strip * kCvStrip;
      for (unsigned m = __ballot_sync(0xffffffffu, pid.x >= 0); m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        put_pillar(cv, feats, __shfl_sync(0xffffffffu, pid.x, src), g0 + 4 * src, lane);
      }
      for (unsigned m = __ballot_sync(0xffffffffu, pid.y >= 0); m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        put_pillar(cv, feats, __shfl_sync(0xffffffffu, pid.y, src), g0 + 4 * src + 1, lane);
      }
      for (unsigned m = __ballot_sync(0xffffffffu, pid.z >= 0); m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        put_pillar(cv, feats, __shfl_sync(0xffffffffu, pid.z, src), g0 + 4 * src + 2, lane);
      }
      for (unsigned m = __ballot_sync(0xffffffffu, pid.w >= 0); m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        put_pillar(cv, feats, __shfl_sync(0xffffffffu, pid.w, src), g0 + 4 * src + 3, lane);
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk writes done before the CTA exits
}

template <bool kCanvas>
__global__ void __launch_bounds__(kCanvas ? kCvThreads : kW2Threads, 1)
k_pfn_tcw2(const float *__restrict__ rows_src, const int *__restrict__ kept_idx, const int *__restrict__ num_points,
           const int *__restrict__ coors, const int *__restrict__ bounds8, float *__restrict__ feats,
           const __grid_constant__ Kargs k, const __grid_constant__ CanvasArgs cv) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kNT = kCanvas ? kCvThreads : kW2Threads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float *s_deco = reinterpret_cast<float *>(smem_raw + k.o_scr);  // [2 sets][128][20] private staging rows
  float *s_ss = reinterpret_cast<float *>(smem_raw + k.o_ss);     // [L][2][128]
  int *s_live = reinterpret_cast<int *>(smem_raw + k.o_tab);      // [2 sets][4] live windows of chunk c (ring)
  int *s_prog = s_live + 4 * kW2Sets;  // canvas: [8 sub-ranges][2 column halves] pillars whose features are in `feats`
  int *s_next = s_prog + 16;           // canvas: [8 sub-ranges] next strip ticket
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + k.o_bar);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + kW2NumBars);
  const uint32_t bar0 = smem_u32(s_bar);
  const uint32_t bar_w = bar0 + 8 * kW2BarW;
  const uint32_t smem_base = smem_u32(smem_raw);
  const int L = k.L;
  long long *s_ts = reinterpret_cast<long long *>(smem_raw + k.o_bar + 128);  // dbg: [8][24] timestamps (plan reserves them)
#define MBEV_TS(slot) do { if ((k.dbg & 8) && blockIdx.x == 0 && lane == 0 && c >= 2 && c < 10) s_ts[(c - 2) * 24 + (slot)] = clock64(); } while (0)

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
  if (kCanvas && tid < 24) s_prog[tid] = 0;  // progress counters and strip tickets
  if (kCanvas) {  // zero source of the writers' bulk copies (generic-proxy writes, read by the async proxy)
    for (int i = tid; i < kCvZeroBytes / 16; i += kNT)
      reinterpret_cast<float4 *>(smem_raw + k.o_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  if (warp == kW2EpiWarps) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  for (int l = 0; l < L; ++l) {
    for (int i = tid; i < k.U[l]; i += kNT) {
      s_ss[(2 * l) * MBEV_MAX_UNITS + i] = __ldg(k.scale[l] + i);
      s_ss[(2 * l + 1) * MBEV_MAX_UNITS + i] = __ldg(k.shift[l] + i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (kCanvas) {  // register budget per role (see kCvRegs*)
    if (warp < kW2EpiWarps) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kCvRegsEpi));
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCvRegsRest));
  }

  if (kCanvas && warp >= kCvWriterWarp0) {
    // =========================================== canvas writers ===============================================
    if (!(k.dbg & 32))
      canvas_writer(cv, feats, s_prog, s_next, smem_base + k.o_zero, warp - kCvWriterWarp0, lane, k.dbg);
  } else if (warp >= kW2EpiWarps) {
    // =========================================== MMA issuer of set `set` =====================================
    const int set = warp - kW2EpiWarps;
    const bool leader = lane == 0;
    if (leader && set == 0) {
      mbar_expect_tx(bar_w, k.w_bytes);  // weights: global image -> shared memory, resident, shared by both sets
      for (uint32_t off = 0; off < k.w_bytes; off += 32768u)
        bulk_g2s(smem_base + off, reinterpret_cast<const char *>(k.w_img) + off, min(32768u, k.w_bytes - off), bar_w);
    }
    __syncwarp();
    mbar_wait(bar_w, 0);
    const uint32_t bs = bar0 + 8 * (kW2BarSet + kW2BarsPerSet * set);
    const uint32_t t_ah = tmem + kW2SetCols * set, t_al = t_ah + 64, t_d = t_ah + 128;
    uint32_t par_x0 = 0, par_evx = 0, par_evm = 0;
    for (int c = 0;; ++c) {
      mbar_wait(bs + 8 * kW2X0, par_x0);
      par_x0 ^= 1u;
      if (*reinterpret_cast<volatile int *>(s_live + 4 * set + (c & 3)) == 0) break;
      tc_fence_after();
      for (int l = 0; l < L; ++l) {
        const int U = k.U[l];
        const uint32_t idesc = make_idesc(U);
        const uint32_t lbo = static_cast<uint32_t>(U) * 16u;
        const uint64_t dh0 = make_bdesc(smem_base + k.w_off[l][0], lbo, 128u);
        const uint64_t dl0 = make_bdesc(smem_base + k.w_off[l][1], lbo, 128u);
        const uint64_t dstep = static_cast<uint64_t>(lbo >> 3);  // one K-step (8 K) in the descriptor address field
        const int nks = (l == 0) ? (k.Kp[0] >> 3) : (k.U[l - 1] >> 3);
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
          if (leader) {
            const uint64_t koff = dstep * static_cast<uint64_t>(part * nks);
#pragma unroll 1
            for (int j = 0; j < nks; ++j) {  // al*wh, ah*wl, ah*wh : small terms first
              const uint64_t dh = dh0 + koff + dstep * j, dl = dl0 + koff + dstep * j;
              mma_tf32_ts(t_d, t_al + 8u * j, dh, idesc, acc);
              mma_tf32_ts(t_d, t_ah + 8u * j, dl, idesc, 1u);
              mma_tf32_ts(t_d, t_ah + 8u * j, dh, idesc, 1u);
              acc = 1;
            }
            tc_commit(bs + 8 * ((l > 0 && part == 0) ? kW2XC : kW2D));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================================== epilogue warps ===============================================
    const int set = warp >> 3, quad = warp & 3, h = (warp >> 2) & 1;
    const int row = (quad << 5) | lane;
    const uint32_t tl = tmem + (static_cast<uint32_t>(quad * 32) << 16) + kW2SetCols * set;  // lane quadrant, set columns
    const uint32_t t_ah = tl, t_al = tl + 64, t_d = tl + 128;
    const uint32_t bs = bar0 + 8 * (kW2BarSet + kW2BarsPerSet * set);
    const int sub = 8 * blockIdx.x + 4 * set + quad;
    const int *bnd = kCanvas ? cv.pb : bounds8;
    const int2 *ord = kCanvas ? cv.ord : nullptr;
    int cursor = __ldg(bnd + sub);
    const int pend = (kCanvas && (k.dbg & 64)) ? cursor : __ldg(bnd + sub + 1);  // dbg 64: writers only
    const int cursor0 = cursor;
    uint32_t par_d = 0, par_x0 = 0, par_xc = 0;
    float *xd = s_deco + (set * kRows + row) * kDecoPitch;
    int *live = s_live + 4 * set;

    for (int c = 0;; ++c) {
      if (warp == 0) MBEV_TS(0);
      const Window w = pack_window(num_points, ord, cursor, pend, k.T, lane);
      if (warp == 0) MBEV_TS(1);
      if (h == 0) build_x0(k, w, rows_src, kept_idx, coors, xd, t_ah, t_al);
      if (warp == 0) MBEV_TS(2);
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
      if (warp == 0) MBEV_TS(3);

      for (int l = 0; l < L; ++l) {
        const int U = k.U[l];
        const bool last = (l == L - 1);
        mbar_wait(bs + 8 * kW2D, par_d);
        par_d ^= 1u;
        tc_fence_after();
        if (warp == 0) MBEV_TS(4 + 4 * l);
        const int Uh = U >> 1;
        const int nbat = Uh >> 4;
        const float *sc = s_ss + (2 * l) * MBEV_MAX_UNITS + h * Uh, *sh = s_ss + (2 * l + 1) * MBEV_MAX_UNITS + h * Uh;
        if (last) {
#pragma unroll 1
          for (int b = 0; b < nbat; ++b) {
            const int col0 = h * Uh + 16 * b;
            float a[16];
            load_bn_relu(t_d + static_cast<uint32_t>(col0), sc + 16 * b, sh + 16 * b, a);
            seg_max16<false>(w, lane, a);
            if (w.inwin && lane == w.s1) {  // the last lane of each pillar holds its max: 16 columns, 64 contiguous bytes
              float4 *out = reinterpret_cast<float4 *>(feats + static_cast<size_t>(w.pil) * U + col0);
              out[0] = make_float4(a[0], a[1], a[2], a[3]);
              out[1] = make_float4(a[4], a[5], a[6], a[7]);
              out[2] = make_float4(a[8], a[9], a[10], a[11]);
              out[3] = make_float4(a[12], a[13], a[14], a[15]);
            }
          }
        } else {
          // non-last layers have U <= 64: at most two 16-column batches per warp, kept in registers across the
          // "x consumed" wait
          float a0[16], a1[16];
          const uint32_t c0 = static_cast<uint32_t>(h * Uh);
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
          if (warp == 0) MBEV_TS(5 + 4 * l);
          seg_max16(w, lane, a0);  // runs while the x-part MMAs do
          if (nbat > 1) seg_max16(w, lane, a1);
          if (warp == 0) MBEV_TS(6 + 4 * l);
          mbar_wait(bs + 8 * kW2XC, par_xc);  // the x-part MMAs have read A: its columns are free again
          par_xc ^= 1u;
          if (warp == 0) MBEV_TS(7 + 4 * l);
          tc_fence_after();
          split_store16(t_ah + c0, t_al + c0, a0);  // max half: K index = U + unit index, same A columns
          if (nbat > 1) split_store16(t_ah + c0 + 16, t_al + c0 + 16, a1);
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bs + 8 * kW2EVM);
        }
      }
      if (warp == 0) MBEV_TS(16);
      cursor += w.cnt;
      if (kCanvas) {  // release: this warp's feature stores, then the progress counter its writer polls
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          *reinterpret_cast<volatile int *>(s_prog + 2 * (4 * set + quad) + h) = cursor - cursor0;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kW2EpiWarps) tmem_dealloc(tmem, kTmemCols);
  if ((k.dbg & 8) && blockIdx.x == 0 && tid == 0) {
    for (int cc = 0; cc < 8; ++cc) {
      const long long *q = s_ts + cc * 24;
      printf("chunk %d: pack %lld x0 %lld rdv %lld | L0: Dw %lld ld+st %lld segmax %lld XCw %lld | L1: mst %lld Dw... %lld ld+st %lld segmax %lld XCw %lld | L2: mst+Dw %lld epi %lld | total %lld\n", cc + 2,
             q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3], q[5] - q[4], q[6] - q[5], q[7] - q[6],
             0LL, q[8] - q[7], q[9] - q[8], q[10] - q[9], q[11] - q[10], q[12] - q[11], q[16] - q[12], q[16] - q[0]);
    }
  }
#undef MBEV_TS
}

// shared-memory plan of k_pfn_tcw2 (returns false when the stack does not fit)
inline bool tcw2_plan(Kargs &k, bool canvas = false) {
  if (tcw_split(k) == 0) return false;
  for (int l = 0; l + 1 < k.L; ++l)
    if (k.U[l] > 64) return false;
  uint32_t o = (k.w_bytes + 127u) & ~127u;
  k.o_scr = o; o += kW2Sets * kRows * kDecoPitch * 4;
  k.o_ss = o; o += k.L * 2 * MBEV_MAX_UNITS * 4;
  k.o_tab = o; o += 4 * kW2Sets * 4 + 24 * 4;  // live ring + the canvas kernel's progress counters and strip tickets
  o = (o + 15u) & ~15u;
  k.o_bar = o; o += 128 + 8 * 24 * 8;  // barriers + TMEM slot, developer timestamps
  k.o_zero = 0;
  if (canvas) {
    o = (o + 127u) & ~127u;
    k.o_zero = o; o += kCvZeroBytes;
  }
  k.smem_bytes = static_cast<int>(o);
  return k.smem_bytes <= kSmemLimit;
}

// FULL forward: two-pipeline warp-local kernel when it applies, else the one-pipeline one, else (and for every
// STATS launch) the block-level kernel. MBEV_TC_KERNEL=tcw|tc forces the older kernels (developer knob).
inline int launch(const Plan &pl, const float *rows, const int32_t *kept_idx, const int32_t *num_points,
                  const int32_t *coors, float *feats, int stat_layer, cudaStream_t stream) {
  Kargs k = pl.k;
  k.stat_layer = stat_layer;
  static const int dbg = getenv("MBEV_TC_DBG") ? atoi(getenv("MBEV_TC_DBG")) : 0;
  static const char *force = getenv("MBEV_TC_KERNEL");
  k.dbg = dbg;
  static bool attr_done = false;
  if (!attr_done) {
    MBEV_CUDA(cudaFuncSetAttribute(k_pfn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    MBEV_CUDA(cudaFuncSetAttribute(k_pfn_tcw<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    MBEV_CUDA(cudaFuncSetAttribute(k_pfn_tcw<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    MBEV_CUDA(cudaFuncSetAttribute(k_pfn_tcw2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_done = true;
  }
  const bool allow_w = !(force && !strcmp(force, "tc"));
  const bool allow_w2 = allow_w && !(force && !strcmp(force, "tcw"));
  const int split = (stat_layer < 0 && allow_w) ? tcw_split(k) : 0;
  Kargs k2 = k;
  if (split && allow_w2 && tcw2_plan(k2)) {
    k_pfn_tcw2<false><<<pl.grid, kW2Threads, k2.smem_bytes, stream>>>(rows, kept_idx, num_points, coors, pl.bounds, feats,
                                                                      k2, CanvasArgs());
  } else if (split) {
    // k_pfn_tcw needs the weights, the [128][20] staging rows, scale/shift and the barriers — not the block
    // kernel's transposition scratch and tables
    uint32_t o = (k.w_bytes + 127u) & ~127u;
    k.o_scr = o; o += kRows * kDecoPitch * 4 + 8 * 24 * 8;  // staging rows + the developer timestamp area
    k.o_ss = o; o += k.L * 2 * MBEV_MAX_UNITS * 4;
    k.o_tab = o; o += 16;
    k.o_bar = o; o += 40 * 8 + 8;
    k.smem_bytes = static_cast<int>(o);
    if (split == 4)
      k_pfn_tcw<4><<<pl.grid, 17 * 32, k.smem_bytes, stream>>>(rows, kept_idx, num_points, coors, pl.bounds, feats, k);
    else
      k_pfn_tcw<2><<<pl.grid, 9 * 32, k.smem_bytes, stream>>>(rows, kept_idx, num_points, coors, pl.bounds, feats, k);
  } else {
    k_pfn_tc<<<pl.grid, kThreads, k.smem_bytes, stream>>>(rows, kept_idx, num_points, coors, pl.bounds, feats, k);
  }
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

}  // namespace tc
}  // namespace mbev
