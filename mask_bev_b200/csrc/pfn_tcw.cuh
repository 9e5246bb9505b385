// K2-TCW — the tcgen05 PFN forward with warp-local pillars (the FULL-forward kernel when T <= 32).
//
// Same math, operands and handshakes as k_pfn_tc (pfn_tc.cuh): A in tensor memory, 3xTF32, weights resident in
// shared memory, accumulators alternating D0 / D1. What changes is how rows are laid out, which removes every
// block barrier and the shared-memory transposition from the layer loop:
//
//   * a chunk is four WINDOWS of 32 rows, one per TMEM lane quadrant. Quadrant w of CTA b walks its own
//     contiguous pillar sub-range (row-balanced bounds, 4 per CTA) and packs whole pillars into its window: a
//     pillar never straddles a warp, so the per-pillar max is a segmented max over the lanes of one warp —
//     log2(longest pillar) shuffle steps, no shared memory.
//   * kSplit epilogue warps share a quadrant and split the columns of every layer; each recomputes the window's
//     packing on its own (one coalesced load + a warp scan), so the warps of a quadrant never synchronise with
//     each other. A warp talks to the MMA issuer only: one mbarrier arrival per finished 16-column batch
//     ("x and max K-parts of the next layer are in TMEM"), one wait per layer ("accumulator complete").
//   * once per chunk all warps meet on bar_x0 (the layer-0 input of every window is in TMEM and nobody still
//     reads the previous chunk's last accumulator); a live-window counter read after that barrier ends the loop
//     for everybody at the same chunk.
//   * cluster mean: each lane re-adds its pillar's points in slot order through shuffles (bit-identical to the
//     sequential slot-order sum upstream computes); decoration goes through a private shared-memory row only to keep the
//     registers statically indexed.
#pragma once
#include <cstdio>

#include "pfn_tc.cuh"

namespace mbev {
namespace tc {

constexpr int kWBarW = 0, kWBarD = 1, kWBarX0 = 3, kWBarEv = 4;
constexpr int kWMaxSplit = 4;
constexpr int kWNumBars = kWBarEv + 8 * kWMaxSplit;

// kSplit = epilogue warps per quadrant (column split of every layer): 4 when every U_l is a multiple of 64, else 2
template <int kSplit>
__global__ void __launch_bounds__((4 * kSplit + 1) * 32, 1)
k_pfn_tcw(const float *__restrict__ rows_src, const int *__restrict__ kept_idx, const int *__restrict__ num_points,
          const int *__restrict__ coors, const int *__restrict__ bounds4, float *__restrict__ feats,
          const __grid_constant__ Kargs k) {
  constexpr int kWEpiWarps = 4 * kSplit;
  constexpr int kWThreads = (kWEpiWarps + 1) * 32;  // + the MMA issuer warp
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = k.T;
  float *s_deco = reinterpret_cast<float *>(smem_raw + k.o_scr);  // [128][20] private rows (layer-0 input staging)
  float *s_ss = reinterpret_cast<float *>(smem_raw + k.o_ss);     // [L][2][128]
  int *s_live = reinterpret_cast<int *>(smem_raw + k.o_tab);      // [4] live windows of chunk c (ring)
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + k.o_bar);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + kWNumBars);
  const uint32_t bar0 = smem_u32(s_bar);
  const uint32_t bar_w = bar0 + 8 * kWBarW, bar_d = bar0 + 8 * kWBarD, bar_x0 = bar0 + 8 * kWBarX0,
                 bar_ev = bar0 + 8 * kWBarEv;
  const uint32_t smem_base = smem_u32(smem_raw);
  const int L = k.L;
  long long *s_ts = reinterpret_cast<long long *>(smem_raw + k.o_scr + 10240);  // [8][24] dbg timestamps
#define MBEV_TS(slot) do { if ((k.dbg & 8) && blockIdx.x == 0 && lane == 0 && c >= 2 && c < 10) s_ts[(c - 2) * 24 + (slot)] = clock64(); } while (0)

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_d, 1);
    mbar_init(bar_d + 8, 1);
    mbar_init(bar_x0, kWEpiWarps);
    for (int i = 0; i < 8 * kSplit; ++i) mbar_init(bar_ev + 8 * i, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 4) s_live[tid] = 0;
  __syncwarp();
  if (warp == kWEpiWarps) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  for (int l = 0; l < L; ++l) {
    for (int i = tid; i < k.U[l]; i += kWThreads) {
      s_ss[(2 * l) * MBEV_MAX_UNITS + i] = __ldg(k.scale[l] + i);
      s_ss[(2 * l + 1) * MBEV_MAX_UNITS + i] = __ldg(k.shift[l] + i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == kWEpiWarps) {
    // =========================================== MMA issuer ===================================================
    const bool leader = lane == 0 ;
    const bool issue = leader && !(k.dbg & 1);
    if (leader) {
      mbar_expect_tx(bar_w, k.w_bytes);  // weights: global image -> shared memory, resident for the whole kernel
      for (uint32_t off = 0; off < k.w_bytes; off += 32768u)
        bulk_g2s(smem_base + off, reinterpret_cast<const char *>(k.w_img) + off, min(32768u, k.w_bytes - off), bar_w);
    }
    __syncwarp();
    mbar_wait(bar_w, 0);
    uint32_t ev = 0, par_x0 = 0;
    for (int c = 0;; ++c) {
      mbar_wait(bar_x0, par_x0);
      par_x0 ^= 1u;
      if (*reinterpret_cast<volatile int *>(s_live + (c & 3)) == 0) break;
      tc_fence_after();
      MBEV_TS(12);
      for (int l = 0; l < L; ++l) {
        const int U = k.U[l];
        const uint32_t idesc = make_idesc(U);
        const uint32_t lbo = static_cast<uint32_t>(U) * 16u;
        const uint32_t d_col = tmem + kColD + ((l & 1) ? 128u : 0u);
        const uint64_t dh0 = make_bdesc(smem_base + k.w_off[l][0], lbo, 128u);
        const uint64_t dl0 = make_bdesc(smem_base + k.w_off[l][1], lbo, 128u);
        const uint64_t dstep = static_cast<uint64_t>(lbo >> 3);  // one K-step (8 K) in the descriptor address field
        uint32_t acc = 0;
        if (l == 0) {
          if (issue) {
            const int nks = k.Kp[0] >> 3;
#pragma unroll 1
            for (int j = 0; j < nks; ++j) {  // al*wh, ah*wl, ah*wh : small terms first
              mma_tf32_ts(d_col, tmem + kColAL + 8u * j, dh0 + dstep * j, idesc, acc);
              mma_tf32_ts(d_col, tmem + kColAH + 8u * j, dl0 + dstep * j, idesc, 1u);
              mma_tf32_ts(d_col, tmem + kColAH + 8u * j, dh0 + dstep * j, idesc, 1u);
              acc = 1;
            }
          }
        } else {
          const int Up = k.U[l - 1];
          const int Uh = Up / kSplit;    // columns of layer l-1 per epilogue warp
          const int nbat = Uh >> 4;      // 16-column batches (= events) per warp
          for (int b = 0; b < nbat; ++b) {
#pragma unroll
            for (int g = 0; g < kSplit; ++g) {
              mbar_wait(bar_ev + 8 * (8 * g + ((ev + b) & 7)), ((ev + b) >> 3) & 1);
              tc_fence_after();
              if (b == 0 && g == 0) MBEV_TS(16 + l);
              if (issue) {
                const uint32_t c0 = static_cast<uint32_t>(g * Uh + 16 * b);
#pragma unroll
                for (int part = 0; part < 2; ++part) {  // x K-part, then the replicated-max K-part
                  const uint32_t kcol = c0 + (part ? static_cast<uint32_t>(Up) : 0u);
                  const uint64_t dh = dh0 + dstep * (kcol >> 3), dl = dl0 + dstep * (kcol >> 3);
#pragma unroll
                  for (uint32_t j = 0; j < 2; ++j) {
                    mma_tf32_ts(d_col, tmem + kColAL + kcol + 8u * j, dh + dstep * j, idesc, acc);
                    mma_tf32_ts(d_col, tmem + kColAH + kcol + 8u * j, dl + dstep * j, idesc, 1u);
                    mma_tf32_ts(d_col, tmem + kColAH + kcol + 8u * j, dh + dstep * j, idesc, 1u);
                    acc = 1;
                  }
                }
              }
              __syncwarp();
            }
          }
          ev += nbat;
        }
        if (leader) tc_commit(bar_d + 8 * (l & 1));
        __syncwarp();
        MBEV_TS(13 + l);
      }
    }
  } else {
    // =========================================== epilogue warps ===============================================
    const int quad = warp & 3, h = warp >> 2;
    const int row = (quad << 5) | lane;
    const uint32_t tlane = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const int sub = 8 * blockIdx.x + 2 * quad;  // the partition has 8 sub-ranges per CTA: a quadrant takes two
    int cursor = __ldg(bounds4 + sub);
    const int pend = __ldg(bounds4 + sub + 2);
    const uint32_t my_ev = bar_ev + 64u * h;
    uint32_t ev = 0, par_d0 = 0, par_d1 = 0, par_x0 = 0;
    float *xd = s_deco + row * kDecoPitch;

    for (int c = 0;; ++c) {
      if (warp == 0) MBEV_TS(0);
      // ---- pack whole pillars cursor, cursor+1, ... into this quadrant's 32-row window -----------------------
      const bool cand = cursor + lane < pend;
      const int np = cand ? __ldg(num_points + cursor + lane) : 0;
      const int need = cand ? np + (np < T ? 1 : 0) : 0;
      int incl = need;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const int cnt = __popc(__ballot_sync(0xffffffffu, cand && incl <= 32));  // prefix-closed
      const int nrows = cnt ? __shfl_sync(0xffffffffu, incl, cnt - 1) : 0;
      const int excl = incl - need;
      // lane r = row r of the window: its pillar is the last i < cnt with excl_i <= r
      int pi = 0;
#pragma unroll
      for (int bit = 16; bit; bit >>= 1) {
        const int j = pi + bit;
        const int e = __shfl_sync(0xffffffffu, excl, j & 31);
        if (j < cnt && e <= lane) pi = j;
      }
      const bool inwin = lane < nrows;
      int s0 = __shfl_sync(0xffffffffu, excl, pi);
      int nd = __shfl_sync(0xffffffffu, need, pi);
      const int n = __shfl_sync(0xffffffffu, np, pi);
      if (!inwin) {  // window padding: a segment of its own
        s0 = lane;
        nd = 1;
      }
      const int s1 = s0 + nd - 1;
      const int t = lane - s0;
      const bool real = inwin && t < n;
      const int pil = cursor + pi;
      int maxlen = inwin ? nd : 1;
#pragma unroll
      for (int o = 16; o; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));

      if (warp == 0) MBEV_TS(1);
      // ---- layer-0 input of this window (warp h == 0 of the quadrant) ---------------------------------------
      if (h == 0) {
        float pv[MBEV_MAX_POINT_DIM];
#pragma unroll
        for (int j = 0; j < MBEV_MAX_POINT_DIM; ++j) pv[j] = 0.f;
        if (real && !(k.dbg & 4)) {
          const size_t slot = static_cast<size_t>(pil) * T + t;
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
        if (inwin) {
          const int4 cc = __ldg(reinterpret_cast<const int4 *>(coors) + pil);  // (b, z, y, x)
          // upstream: coors.type_as(features) * vx + x_offset — float32 multiply THEN add (no FMA contraction)
          cx = __fadd_rn(__fmul_rn(static_cast<float>(cc.w), k.vx), k.xo);
          cy = __fadd_rn(__fmul_rn(static_cast<float>(cc.z), k.vy), k.yo);
          cz = __fadd_rn(__fmul_rn(static_cast<float>(cc.y), k.vz), k.zo);
        }
        // cluster mean: the pillar's points re-added in slot order by every lane of the pillar, / num_points
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int tt = 0; tt < maxlen; ++tt) {
          const float vx = __shfl_sync(0xffffffffu, pv[0], (s0 + tt) & 31);
          const float vy = __shfl_sync(0xffffffffu, pv[1], (s0 + tt) & 31);
          const float vz = __shfl_sync(0xffffffffu, pv[2], (s0 + tt) & 31);
          if (tt < n) {
            sx = __fadd_rn(sx, vx);
            sy = __fadd_rn(sy, vy);
            sz = __fadd_rn(sz, vz);
          }
        }
#pragma unroll
        for (int d = 0; d < kK0Pad; ++d) xd[d] = 0.f;  // virtual rows, window padding and the K padding
        if (real) {
          const float fn = static_cast<float>(n);
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
          if (k.dist)
            xd[d++] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r0, r0), __fmul_rn(r1, r1)), __fmul_rn(r2, r2)));
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 v = *reinterpret_cast<const float4 *>(xd + 4 * j4);
          split_tf32(v.x, hi[4 * j4 + 0], lo[4 * j4 + 0]);
          split_tf32(v.y, hi[4 * j4 + 1], lo[4 * j4 + 1]);
          split_tf32(v.z, hi[4 * j4 + 2], lo[4 * j4 + 2]);
          split_tf32(v.w, hi[4 * j4 + 3], lo[4 * j4 + 3]);
        }
        tmem_st16(tlane + kColAH, hi);
        tmem_st16(tlane + kColAL, lo);
        tc_wait_st();
      }
      if (warp == 0) MBEV_TS(2);
      // ---- chunk rendezvous: every window's layer-0 input is in TMEM, nobody reads the previous chunk's D ----
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (warp == 0) s_live[(c + 2) & 3] = 0;  // everybody read this slot two rendezvous ago
        if (h == 0 && cnt > 0) atomicAdd(s_live + (c & 3), 1);
        mbar_arrive(bar_x0);
      }
      mbar_wait(bar_x0, par_x0);
      par_x0 ^= 1u;
      if (*reinterpret_cast<volatile int *>(s_live + (c & 3)) == 0) break;
      if (warp == 0) MBEV_TS(3);

      // ---- layers -------------------------------------------------------------------------------------------
      for (int l = 0; l < L; ++l) {
        const int U = k.U[l];
        const bool last = (l == L - 1);
        const uint32_t d_col = kColD + ((l & 1) ? 128u : 0u);
        if (l & 1) {
          mbar_wait(bar_d + 8, par_d1);
          par_d1 ^= 1u;
        } else {
          mbar_wait(bar_d, par_d0);
          par_d0 ^= 1u;
        }
        tc_fence_after();
        if (warp == 0) MBEV_TS(4 + 2 * l);
        const int Uh = U / kSplit;
        const int nbat = Uh >> 4;
#pragma unroll 1
        for (int b = 0; b < nbat; ++b) {
          const int col0 = h * Uh + 16 * b;
          float a[16];
          {
            uint32_t v[16];
            tmem_ld16(tlane + d_col + static_cast<uint32_t>(col0), v);
            tc_wait_ld();
            const float4 *sc4 = reinterpret_cast<const float4 *>(s_ss + (2 * l) * MBEV_MAX_UNITS + col0);
            const float4 *sh4 = reinterpret_cast<const float4 *>(s_ss + (2 * l + 1) * MBEV_MAX_UNITS + col0);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 sc = sc4[j4], sh = sh4[j4];
              a[4 * j4 + 0] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 0]), sc.x, sh.x), 0.f);
              a[4 * j4 + 1] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 1]), sc.y, sh.y), 0.f);
              a[4 * j4 + 2] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 2]), sc.z, sh.z), 0.f);
              a[4 * j4 + 3] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 3]), sc.w, sh.w), 0.f);
            }
          }
          if (!last) {  // x half of the next layer's input: K index = unit index
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split_tf32(a[j], hi[j], lo[j]);
            tmem_st16(tlane + kColAH + static_cast<uint32_t>(col0), hi);
            tmem_st16(tlane + kColAL + static_cast<uint32_t>(col0), lo);
          }
          // per-pillar max: segmented inclusive max-scan up the lanes (log2(longest pillar) steps), then every lane
          // reads its pillar's last lane. (redux.sync with per-pillar masks is serialised per mask: slower.)
          for (int d = 1; d < ((k.dbg & 2) ? 0 : maxlen); d <<= 1) {
            const bool take = lane - d >= s0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float o = __shfl_up_sync(0xffffffffu, a[j], d);
              a[j] = take ? fmaxf(a[j], o) : a[j];
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) a[j] = __shfl_sync(0xffffffffu, a[j], s1);
          if (last) {
            if (inwin && t == 0) {  // first lane of each pillar writes its 16 columns (64 contiguous bytes)
              float4 *out = reinterpret_cast<float4 *>(feats + static_cast<size_t>(pil) * U + col0);
              out[0] = make_float4(a[0], a[1], a[2], a[3]);
              out[1] = make_float4(a[4], a[5], a[6], a[7]);
              out[2] = make_float4(a[8], a[9], a[10], a[11]);
              out[3] = make_float4(a[12], a[13], a[14], a[15]);
            }
          } else {  // max half of the next layer's input: K index = U + unit index
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split_tf32(a[j], hi[j], lo[j]);
            tmem_st16(tlane + kColAH + static_cast<uint32_t>(U + col0), hi);
            tmem_st16(tlane + kColAL + static_cast<uint32_t>(U + col0), lo);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(my_ev + 8 * (ev & 7));
            ++ev;
          }
        }
        if (warp == 0) MBEV_TS(5 + 2 * l);
      }
      cursor += cnt;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWEpiWarps) tmem_dealloc(tmem, kTmemCols);
  if ((k.dbg & 8) && blockIdx.x == 0 && tid == 0) {
    for (int cc = 0; cc < 8; ++cc) {
      const long long *q = s_ts + cc * 24;
      printf("chunk %d epi: pack %lld x0 %lld rdv %lld | D0w %lld L0 %lld D1w %lld L1 %lld D2w %lld L2 %lld | next %lld || mma: x0->start %lld L0c %lld ev1 %lld L1c %lld ev2 %lld L2c %lld\n", cc + 2,
             q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3], q[5] - q[4], q[6] - q[5], q[7] - q[6], q[8] - q[7], q[9] - q[8],
             cc < 7 ? q[24] - q[9] : 0LL, q[12] - q[2], q[13] - q[12], q[17] - q[13], q[14] - q[17], q[18] - q[14], q[15] - q[18]);
    }
  }
#undef MBEV_TS
}

inline int tcw_split(const Kargs &k) {  // 0: not supported
  if (k.T > 32) return 0;
  // measured on kitti_b16 [128,128,128]: 2 warps per quadrant 1.03 ms, 4 warps 1.07 ms (with one chunk in flight
  // the layer chain, not issue slots, is the limit) -> 2; MBEV_TC_SPLIT=4 selects the other instantiation
  static const int want = getenv("MBEV_TC_SPLIT") ? atoi(getenv("MBEV_TC_SPLIT")) : 2;
  int split = want == 4 ? 4 : 2;
  for (int l = 0; l < k.L; ++l) {
    if (k.U[l] % 32) return 0;
    if (k.U[l] % 64) split = 2;
  }
  return split;
}

}  // namespace tc
}  // namespace mbev
