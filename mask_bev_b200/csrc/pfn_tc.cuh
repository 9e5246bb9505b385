// K2-TC — pillar feature net forward on the 5th-generation tensor cores (tcgen05 / TMEM, sm_100a).
//
// Same contract as k_pfn in pfn.cu (mmdet3d PillarFeatureNet.forward / PFNLayer.forward as called from
// mask_bev_encoders.py:119-120): compact rows = stored points + one weighted virtual row per
// padded pillar, every activation on chip, HBM traffic = gathered points in + (pillar, C_out) features out.
// What changes is where the Linear layers run:
//
//   * one CTA per SM, 128 threads = 128 TMEM lanes = one chunk of <= 128 rows (whole pillars, greedily packed
//     from the CTA's contiguous pillar range). Thread r owns row r from decoration to the last layer.
//   * layer input A (128 x K_l) lives in TENSOR MEMORY, written by its owning thread with tcgen05.st; the
//     weights B = nn.Linear.weight (U_l x K_l, K-major — the layout PyTorch already stores) stay resident in
//     shared memory for the whole kernel (one cp.async.bulk per CTA); D (128 x U_l fp32) accumulates in TMEM.
//     tcgen05.mma.cta_group::1.kind::tf32, A from TMEM, B through a no-swizzle K-major shared-memory descriptor.
//   * fp32 parity (1e-5) rules out plain TF32, so every product is split 3xTF32:  a = ah + al, w = wh + wl
//     (each part exactly representable in TF32, round-to-nearest), D = al*wh + ah*wl + ah*wh in the fp32
//     accumulator. The dropped al*wl term is ~2^-22 relative. 3 MMAs per K-step at 2048 MAC/clk/SM is still
//     ~9x the fp32 FMA pipe (128 MAC/clk/SM).
//   * the concat [x || max_p] is kept as upstream has it (K_l = 2 U_{l-1}): the per-pillar max is replicated
//     into the second K-half of each row of A, so the pillar term rides in the same accumulator and the
//     epilogue needs no per-pillar add.
//   * epilogue per 32 columns: tcgen05.ld -> BN scale/shift + ReLU in registers -> (x half of next A: split,
//     tcgen05.st) -> transpose through a 128x33 shared scratch -> one thread per column walks the rows of its
//     pillars (segmented max, warp-uniform control flow) and writes the max back over the rows -> each row
//     thread reads its pillar's max, splits it and stores the max half of the next A. Last layer: the column
//     threads store the pillar max straight to HBM (coalesced 128 B per warp).
//   * train mode: STATS launches accumulate sum / sum-of-squares of the pre-BN output of layer s in fp64 in the
//     column phase (fixed order: static pillar partition -> run-to-run identical), as k_pfn does.
//
// Supported when every U_l is a multiple of 32 and <= 128, every K_l <= 128, T + 1 <= 128 and the weight image
// fits shared memory ([128,128,128], [128,64,128], [64], ...). Other stacks run on the fp32-FMA kernel.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace mbev {
namespace tc {

constexpr int kEpiThreads = 256;            // 8 epilogue warps
constexpr int kThreads = kEpiThreads + 32;  // + the MMA issuer warp
constexpr int kRows = 128;      // rows per chunk = TMEM lanes = MMA M
constexpr int kPcap = 64;       // pillars per chunk (>= 128 / 2: a padded pillar has >= 2 rows)
constexpr int kScrPitch = 33;   // transposition scratch [128][33] floats
constexpr int kDecoPitch = 20;  // decorated layer-0 input staging [128][20] floats (float4-aligned, conflict-free)
constexpr int kPtPitch = 8;
constexpr int kK0Pad = 16;
// Layer-0 input columns in TENSOR MEMORY use a fixed "wide" order so that the decoration needs only statically
// indexed registers whatever the configuration: [0,3) xyz (legacy: the in-place centre offset), [3,8) the other raw
// point features, [8,11) cluster offset, [11,14) centre offset, [14] distance, [15] zero. k_prep_weights_tc puts the
// Linear weight column of upstream's dense order (raw | cluster | centre | distance) at its wide slot and zeros at
// the unused ones, so the products are the same numbers.
constexpr int kWideExtra = 3, kWideCluster = 8, kWideCentre = 11, kWideDist = 14;
constexpr uint32_t kColAH = 0, kColAL = 128, kColD = 256, kTmemCols = 512;
constexpr int kSmemLimit = 227 * 1024;

struct Kargs {
  int L;
  int K[MBEV_MAX_LAYERS];   // Linear.in_features
  int Kp[MBEV_MAX_LAYERS];  // padded to a multiple of 8 (one tf32 MMA K-step)
  int U[MBEV_MAX_LAYERS];
  uint32_t w_off[MBEV_MAX_LAYERS][2];  // byte offset of the hi / lo weight image of layer l (global image == smem image)
  uint32_t w_bytes;
  const float *w_img;
  const float *scale[MBEV_MAX_LAYERS];
  const float *shift[MBEV_MAX_LAYERS];
  int C, D0, T;
  int cluster, vcenter, dist, legacy, vcd;
  float vx, vy, vz, xo, yo, zo;
  int um;
  uint32_t o_scr, o_ss, o_tab, o_bar;  // byte offsets in dynamic shared memory
  int smem_bytes;
  int stat_layer;    // -1: full forward; s: accumulate the statistics of layer s and stop
  double *partials;  // (gridDim.x, 2, um)
  int bf16;          // 1: layers >= 1 run as single-pass bf16 MMAs (kind::f16), layer 0 stays 3xTF32 (k_pfn_tcw2 only)
};

// ---- row-balanced static partition ------------------------------------------------------------------------
// Pillars are numbered in order of first appearance, so the heavy pillars of a frame come first: an equal-count
// split leaves the CTAs 3x apart in rows. k_row_blocks sums the compact rows (n + [n < T]) of blocks of 32
// pillars, k_row_bounds prefix-sums the blocks and gives sub-range b (8 per CTA) the blocks whose prefix falls in
// [R b / G, R (b+1) / G). Static and deterministic (train-mode statistics stay run-to-run identical).
constexpr int kBlkPillars = 32;
constexpr int kRbBlocks = 64;  // blocks of 32 pillars per partition CTA (2048 pillars)

// per block of 32 pillars: compact rows; per CTA (64 blocks): their total
__global__ void __launch_bounds__(256)
k_row_blocks(const int *__restrict__ num_points, const int *__restrict__ num_pillars, const int T, const int nb,
             int *__restrict__ blocksum, int *__restrict__ ctot) {
  __shared__ int s_w[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, P = *num_pillars;
  static_assert(kBlkPillars == 32, "one pillar per lane");
  int wsum = 0;
#pragma unroll
  for (int i = 0; i < kRbBlocks / 8; ++i) {
    const int j = blockIdx.x * kRbBlocks + warp * (kRbBlocks / 8) + i;
    const int p = j * kBlkPillars + lane;
    int s = 0;
    if (j < nb && p < P) {
      const int n = __ldg(num_points + p);
      s = n + (n < T ? 1 : 0);
    }
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0 && j < nb) blocksum[j] = s;
    wsum += s;
  }
  if (lane == 0) s_w[warp] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_w[w];
    ctot[blockIdx.x] = t;
  }
}

// Two-level prefix (every CTA re-reduces the <= ~1 k CTA totals, then scans its own 64 blocks) and the bounds that
// fall inside this CTA: bound b = first block j whose exclusive row prefix reaches ceil(R b / G). The former
// single-CTA version of this kernel cost a fixed ~60 us per forward call.
__global__ void __launch_bounds__(256)
k_row_bounds(const int *__restrict__ blocksum, const int *__restrict__ ctot, const int nbA, const int nb,
             const int *__restrict__ num_pillars, const int *__restrict__ num_points, const int T, const int G,
             int *__restrict__ bounds) {
  __shared__ int s_red[2][8];
  __shared__ int s_pre[kRbBlocks + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, bid = blockIdx.x;
  int base = 0, tot = 0;
  for (int j = tid; j < nbA; j += 256) {
    const int v = __ldg(ctot + j);
    tot += v;
    if (j < bid) base += v;
  }
  base = __reduce_add_sync(0xffffffffu, base);
  tot = __reduce_add_sync(0xffffffffu, tot);
  if (lane == 0) {
    s_red[0][warp] = base;
    s_red[1][warp] = tot;
  }
  const int j0 = bid * kRbBlocks;
  const int nloc = min(kRbBlocks, nb - j0);
  if (warp == 0) {
    const int a = (2 * lane < nloc) ? __ldg(blocksum + j0 + 2 * lane) : 0;
    const int c = (2 * lane + 1 < nloc) ? __ldg(blocksum + j0 + 2 * lane + 1) : 0;
    int inc = a + c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    s_pre[2 * lane] = inc - a - c;
    s_pre[2 * lane + 1] = inc - c;
    if (lane == 31) s_pre[kRbBlocks] = inc;
  }
  __syncthreads();
  base = 0;
  tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    base += s_red[0][w];
    tot += s_red[1][w];
  }
  const long long R = tot;
  const int P = *num_pillars;
  const long long lo_pre = base, hi_pre = static_cast<long long>(base) + s_pre[nloc];
  for (int b = tid; b <= G; b += 256) {
    if (b == G) {
      if (bid == nbA - 1) bounds[G] = P;
      continue;
    }
    const long long target = (R * b + G - 1) / G;
    if (target <= 0) {
      if (bid == 0) bounds[b] = 0;
      continue;
    }
    if (!(lo_pre < target && target <= hi_pre)) continue;
    int lo = 1, hi = nloc;  // smallest local j in [1, nloc] with base + s_pre[j] >= target
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (base + s_pre[mid] >= target) hi = mid; else lo = mid + 1;
    }
    // the crossing lies inside block j0 + lo - 1: walk its pillars so that the bound is pillar-exact (with one
    // frame the 1184 sub-ranges hold ~24 pillars each, less than a block)
    long long acc = static_cast<long long>(base) + s_pre[lo - 1];
    int p = (j0 + lo - 1) * kBlkPillars;
    const int pe = min(p + kBlkPillars, P);
    while (p < pe && acc < target) {
      const int n = __ldg(num_points + p);
      acc += n + (n < T ? 1 : 0);
      ++p;
    }
    bounds[b] = p;
  }
}

// ---- kernel ---------------------------------------------------------------------------------------------------
// Warp roles: warps 0-7 = 256 epilogue threads in two groups of 128 (group h = warp>>2 owns columns
// 16h..16h+15 of every 32-column slab; inside a group thread `row` = 32*(warp&3)+lane owns TMEM lane `row`);
// warp 8 = MMA issuer (one elected lane issues, all lanes walk the handshakes).
// Handshakes: bar_x0 (256 arrivals): "layer-0 input of the next chunk is in TMEM" (or stop);
// slab rings (2 groups x 8 mbarriers, 128 arrivals): "16 more K-columns of the next layer's input are in TMEM";
// bar_d[2] (tcgen05.commit): "the accumulator of a layer is complete". Accumulators alternate D0 / D1 so the
// MMAs of layer l+1 start on the first finished K-columns while layer l's epilogue is still draining its D.
// The two groups only meet at chunk boundaries; in between each runs on its own named barrier, so one group's
// shared-memory column walk overlaps the other's TMEM traffic.
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 3, 256;" ::: "memory"); }
__device__ __forceinline__ void grp_sync(int h) { asm volatile("bar.sync %0, 128;" ::"r"(h + 1) : "memory"); }

constexpr int kBarW = 0, kBarD = 1, kBarX0 = 3, kBarSlab = 4, kNumBars = 20;

__global__ void __launch_bounds__(kThreads, 1)
k_pfn_tc(const float *__restrict__ rows_src, const int *__restrict__ kept_idx, const int *__restrict__ num_points,
         const int *__restrict__ coors, const int *__restrict__ bounds, float *__restrict__ feats,
         const __grid_constant__ Kargs k) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = k.T;
  float *s_scr = reinterpret_cast<float *>(smem_raw + k.o_scr);
  float *s_pts = s_scr;                      // [128][8]   dead once the layer-0 input is built
  float *s_deco = s_scr + kRows * kPtPitch;  // [128][20]
  float *s_ss = reinterpret_cast<float *>(smem_raw + k.o_ss);  // [L][2][128]
  int *s_n = reinterpret_cast<int *>(smem_raw + k.o_tab);      // [64]
  int *s_prow0 = s_n + kPcap;                                  // [65] (+3 pad)
  int *s_rinfo = s_prow0 + kPcap + 4;                          // [128]: first row of the row's pillar | pillar << 8
  int *s_rend = s_rinfo + kRows;                               // [128]: 1 = the row closes its pillar
  float *s_roww = reinterpret_cast<float *>(s_rend + kRows);   // [128]
  float *s_mean = s_roww + kRows;                              // [64][4]
  float *s_ctr = s_mean + kPcap * 4;                           // [64][4]
  int *s_misc = reinterpret_cast<int *>(s_ctr + kPcap * 4);    // [8]: scan carries, counts, [7] = continue flag
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + k.o_bar);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + kNumBars);
  const uint32_t bar0 = smem_u32(s_bar);
  const uint32_t bar_w = bar0 + 8 * kBarW, bar_d = bar0 + 8 * kBarD, bar_x0 = bar0 + 8 * kBarX0,
                 bar_slab = bar0 + 8 * kBarSlab;
  const uint32_t smem_base = smem_u32(smem_raw);
  const int last_layer = (k.stat_layer >= 0) ? k.stat_layer : k.L - 1;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_d, 1);
    mbar_init(bar_d + 8, 1);
    mbar_init(bar_x0, kEpiThreads);
    for (int i = 0; i < 16; ++i) mbar_init(bar_slab + 8 * i, kEpiThreads / 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == kEpiThreads / 32) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  for (int l = 0; l < k.L; ++l) {
    if (k.stat_layer < 0 || l < k.stat_layer) {
      for (int i = tid; i < k.U[l]; i += kThreads) {
        s_ss[(2 * l) * MBEV_MAX_UNITS + i] = __ldg(k.scale[l] + i);
        s_ss[(2 * l + 1) * MBEV_MAX_UNITS + i] = __ldg(k.shift[l] + i);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == kEpiThreads / 32) {
    // =========================================== MMA issuer ===================================================
    // every lane walks the handshakes (warp-uniform control flow); one elected lane issues copies, MMAs and commits,
    // with operands the compiler can prove warp-uniform (see k_pfn_tcw2: no per-MMA ELECT / R2UR waterfall)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    if (elect_one()) {
      mbar_expect_tx(bar_w, k.w_bytes);  // weights: global image -> shared memory, resident for the whole kernel
      for (uint32_t off = 0; off < k.w_bytes; off += 32768u)
        bulk_g2s(smem_base + off, reinterpret_cast<const char *>(k.w_img) + off, min(32768u, k.w_bytes - off), bar_w);
    }
    __syncwarp();
    mbar_wait(bar_w, 0);
    uint32_t ev = 0, par_x0 = 0;  // ev: events consumed per group ring (both rings advance together)
    for (;;) {
      mbar_wait(bar_x0, par_x0);  // layer-0 input of the next chunk, or the stop signal
      par_x0 ^= 1u;
      if (*reinterpret_cast<volatile int *>(s_misc + 7) == 0) break;
      tc_fence_after();
      for (int l = 0; l <= last_layer; ++l) {
        const int U = k.U[l];
        const uint32_t idesc = make_idesc(U);
        const uint32_t lbo = static_cast<uint32_t>(U) * 16u;
        const uint32_t d_col = tmem_u + kColD + ((l & 1) ? 128u : 0u);
        // descriptors advance by one K-step (two 16-byte K-chunks = 2 * lbo bytes) = (2 * lbo) >> 4 in the address field
        const uint64_t dh0 = make_bdesc(smem_base + k.w_off[l][0], lbo, 128u);
        const uint64_t dl0 = make_bdesc(smem_base + k.w_off[l][1], lbo, 128u);
        const uint32_t dstep = lbo >> 3;
        uint32_t acc = 0;
        if (l == 0) {
          if (elect_one()) {
            const int nks = k.Kp[0] >> 3;
#pragma unroll 1
            for (int j = 0; j < nks; ++j) {  // al*wh, ah*wl, ah*wh : small terms first
              mma_tf32_ts(d_col, tmem_u + kColAL + 8u * j, dh0 + dstep * j, idesc, acc);
              mma_tf32_ts(d_col, tmem_u + kColAH + 8u * j, dl0 + dstep * j, idesc, 1u);
              mma_tf32_ts(d_col, tmem_u + kColAH + 8u * j, dh0 + dstep * j, idesc, 1u);
              acc = 1;
            }
          }
        } else {
          // 16-column K-parts in the order the groups finish them: x(q) g0, x(q) g1, max(q) g0, max(q) g1
          const int Up = k.U[l - 1];
          const int nev = 2 * (Up >> 5);  // events per group ring in this layer
          for (int e = 0; e < nev; ++e) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              mbar_wait(bar_slab + 8 * (8 * g + ((ev + e) & 7)), ((ev + e) >> 3) & 1);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t kcol = static_cast<uint32_t>(((e & 1) ? Up : 0) + 32 * (e >> 1) + 16 * g);
                const uint64_t dh = dh0 + dstep * (kcol >> 3), dl = dl0 + dstep * (kcol >> 3);
#pragma unroll
                for (uint32_t j = 0; j < 2; ++j) {
                  mma_tf32_ts(d_col, tmem_u + kColAL + kcol + 8u * j, dh + dstep * j, idesc, acc);
                  mma_tf32_ts(d_col, tmem_u + kColAH + kcol + 8u * j, dl + dstep * j, idesc, 1u);
                  mma_tf32_ts(d_col, tmem_u + kColAH + kcol + 8u * j, dh + dstep * j, idesc, 1u);
                  acc = 1;
                }
              }
              __syncwarp();
            }
          }
          ev += nev;
        }
        if (elect_one()) tc_commit(bar_d + 8 * (l & 1));
        __syncwarp();
      }
    }
  } else {
    // =========================================== epilogue warps ===============================================
    const int row = ((warp & 3) << 5) | lane, h = warp >> 2;
    const uint32_t tlane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);  // this warp's TMEM lane quadrant
    const int p_begin = __ldg(bounds + 8 * blockIdx.x), p_end = __ldg(bounds + 8 * blockIdx.x + 8);  // 8 sub-ranges per CTA
    const uint32_t my_slab = bar_slab + 64u * h;
    uint32_t ev = 0, par_d0 = 0, par_d1 = 0;
    double st1[4] = {0.0, 0.0, 0.0, 0.0}, st2[4] = {0.0, 0.0, 0.0, 0.0};
    // column-walk role inside the group: 16 columns x 8 pillar ranges
    const int cw_c = lane & 15, cw_g = ((warp & 3) << 1) | (lane >> 4);

    for (int p = p_begin; p < p_end;) {
      epi_sync();  // previous chunk is done with the tables and the scratch
      // ---- pack whole pillars p, p+1, ... into <= 128 rows (<= 64 pillars) ----------------------------------
      const bool cand = tid < kPcap && p + tid < p_end;
      const int n_i = cand ? __ldg(num_points + p + tid) : 0;
      const int need = cand ? n_i + (n_i < T ? 1 : 0) : 0;
      int incl = need;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31 && warp < 2) s_misc[warp] = incl;
      epi_sync();
      if (warp == 1) incl += s_misc[0];
      const bool fits = cand && incl <= kRows;
      const unsigned bal = __ballot_sync(0xffffffffu, fits);
      if (lane == 0 && warp < 2) s_misc[4 + warp] = __popc(bal);
      if (fits) {
        s_n[tid] = n_i;
        s_prow0[tid + 1] = incl;
        const int r0 = incl - need;
        for (int t = 0; t < need; ++t) {
          s_rinfo[r0 + t] = r0 | (tid << 8);
          s_rend[r0 + t] = (t == need - 1);
        }
        const int4 c = __ldg(reinterpret_cast<const int4 *>(coors) + p + tid);  // (b, z, y, x)
        // upstream: coors.type_as(features) * vx + x_offset — float32 multiply THEN add (no FMA contraction)
        s_ctr[tid * 4 + 0] = __fadd_rn(__fmul_rn(static_cast<float>(c.w), k.vx), k.xo);
        s_ctr[tid * 4 + 1] = __fadd_rn(__fmul_rn(static_cast<float>(c.z), k.vy), k.yo);
        s_ctr[tid * 4 + 2] = __fadd_rn(__fmul_rn(static_cast<float>(c.y), k.vz), k.zo);
      }
      if (tid == 0) s_prow0[0] = 0;
      epi_sync();
      const int npil = s_misc[4] + s_misc[5];  // fits is prefix-closed: the count of leading pillars taken
      const int nrows = s_prow0[npil];
      // ---- gather this row's point (group 0) ----------------------------------------------------------------
      int pl = 0;
      bool real = false;
      if (h == 0) {
        int n = 0, t = 0;
        const bool inrange = row < nrows;
        if (inrange) {
          const int info = s_rinfo[row];
          pl = info >> 8;
          t = row - (info & 255);
          n = s_n[pl];
          real = t < n;
        }
        s_roww[row] = inrange ? (real ? 1.f : static_cast<float>(T - n)) : 0.f;
        if (real) {
          const size_t slot = static_cast<size_t>(p + pl) * T + t;
          const int src = kept_idx ? __ldg(kept_idx + slot) : static_cast<int>(slot);
          const float *pp = rows_src + static_cast<size_t>(src) * k.C;
          float *dst = s_pts + row * kPtPitch;
          if (k.C == 4) {
            *reinterpret_cast<float4 *>(dst) = __ldg(reinterpret_cast<const float4 *>(pp));
          } else {
            for (int c = 0; c < k.C; ++c) dst[c] = __ldg(pp + c);
          }
        }
      }
      epi_sync();
      if (tid < npil) {  // cluster mean: sum over the pillar's slots in slot order / num_points
        const int n = s_n[tid];
        const float *pp = s_pts + s_prow0[tid] * kPtPitch;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int t = 0; t < n; ++t) {
          sx = __fadd_rn(sx, pp[t * kPtPitch + 0]);
          sy = __fadd_rn(sy, pp[t * kPtPitch + 1]);
          sz = __fadd_rn(sz, pp[t * kPtPitch + 2]);
        }
        const float fn = static_cast<float>(n);
        s_mean[tid * 4 + 0] = __fdiv_rn(sx, fn);
        s_mean[tid * 4 + 1] = __fdiv_rn(sy, fn);
        s_mean[tid * 4 + 2] = __fdiv_rn(sz, fn);
      }
      epi_sync();
      // ---- decoration -> layer-0 input row (staged through shared memory so that the registers are statically
      //      indexed), split and stored to TMEM lanes ----------------------------------------------------------
      if (h == 0) {
        float *xd = s_deco + row * kDecoPitch;
#pragma unroll
        for (int d = 0; d < kK0Pad; ++d) xd[d] = 0.f;  // virtual rows, chunk padding and the K padding
        if (real) {
          const float *pp = s_pts + row * kPtPitch;
          const float x = pp[0], y = pp[1], z = pp[2];
          const float ex = __fsub_rn(x, s_ctr[pl * 4 + 0]), ey = __fsub_rn(y, s_ctr[pl * 4 + 1]),
                      ez = __fsub_rn(z, s_ctr[pl * 4 + 2]);
          const bool alias = k.vcenter && k.legacy;  // legacy: centre offset written in place over xyz
          const float r0 = alias ? ex : x, r1 = alias ? ey : y, r2 = (alias && k.vcd > 2) ? ez : z;  // 2-channel centre: z stays raw
          xd[0] = r0;
          xd[1] = r1;
          xd[2] = r2;
          for (int c = 3; c < k.C; ++c) xd[kWideExtra + c - 3] = pp[c];
          if (k.cluster) {
            xd[kWideCluster + 0] = __fsub_rn(x, s_mean[pl * 4 + 0]);
            xd[kWideCluster + 1] = __fsub_rn(y, s_mean[pl * 4 + 1]);
            xd[kWideCluster + 2] = __fsub_rn(z, s_mean[pl * 4 + 2]);
          }
          if (k.vcenter) {
            xd[kWideCentre + 0] = ex;
            xd[kWideCentre + 1] = ey;
            if (k.vcd > 2) xd[kWideCentre + 2] = ez;
          }
          if (k.dist)
            xd[kWideDist] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r0, r0), __fmul_rn(r1, r1)), __fmul_rn(r2, r2)));
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
      if (tid == 0) s_misc[7] = 1;
      tc_fence_before();
      mbar_arrive(bar_x0);  // (the staging area aliased with the scratch is next written after D0 is ready)

      // column-walk range of this thread: rows of pillars [npil*g/8, npil*(g+1)/8)
      const int cw_pa = (npil * cw_g) >> 3;
      const int cw_ra = s_prow0[cw_pa], cw_rb = s_prow0[(npil * (cw_g + 1)) >> 3];
      const int fr = row < nrows ? (s_rinfo[row] & 255) : row;  // first row of this row's pillar (padding rows: self)

      // ---- layers -------------------------------------------------------------------------------------------
      for (int l = 0; l <= last_layer; ++l) {
        const int U = k.U[l];
        const bool stats = (l == k.stat_layer);
        const bool last = (l == k.L - 1);
        const uint32_t d_col = kColD + ((l & 1) ? 128u : 0u);
        if (l & 1) {
          mbar_wait(bar_d + 8, par_d1);
          par_d1 ^= 1u;
        } else {
          mbar_wait(bar_d, par_d0);
          par_d0 ^= 1u;
        }
        tc_fence_after();
        const int nq = U >> 5;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < nq) {
            float a[16];
            {
              uint32_t v[16];
              tmem_ld16(tlane + d_col + 32u * q + 16u * h, v);
              tc_wait_ld();
              if (stats) {
#pragma unroll
                for (int j = 0; j < 16; ++j) a[j] = __uint_as_float(v[j]);
              } else {
                const float4 *sc4 = reinterpret_cast<const float4 *>(s_ss + (2 * l) * MBEV_MAX_UNITS + 32 * q + 16 * h);
                const float4 *sh4 = reinterpret_cast<const float4 *>(s_ss + (2 * l + 1) * MBEV_MAX_UNITS + 32 * q + 16 * h);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  const float4 sc = sc4[j4], sh = sh4[j4];
                  a[4 * j4 + 0] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 0]), sc.x, sh.x), 0.f);
                  a[4 * j4 + 1] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 1]), sc.y, sh.y), 0.f);
                  a[4 * j4 + 2] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 2]), sc.z, sh.z), 0.f);
                  a[4 * j4 + 3] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 3]), sc.w, sh.w), 0.f);
                }
              }
            }
            if (!last && !stats) {  // x half of the next layer's input: K index = unit index
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) split_tf32(a[j], hi[j], lo[j]);
              tmem_st16(tlane + kColAH + 32u * q + 16u * h, hi);
              tmem_st16(tlane + kColAL + 32u * q + 16u * h, lo);
              tc_wait_st();
              tc_fence_before();
              mbar_arrive(my_slab + 8 * (ev & 7));
              ++ev;
            }
            {
              float *dst = s_scr + row * kScrPitch + 16 * h;
#pragma unroll
              for (int j = 0; j < 16; ++j) dst[j] = a[j];
            }
            grp_sync(h);
            // column walk: this thread owns column 32q+16h+cw_c over rows [cw_ra, cw_rb)
            {
              float *col = s_scr + 16 * h + cw_c;
              if (stats) {
                double s1 = 0.0, s2 = 0.0;
                for (int r = cw_ra; r < cw_rb; ++r) {
                  const double y = static_cast<double>(col[r * kScrPitch]);
                  const double w = static_cast<double>(s_roww[r]);
                  s1 += w * y;
                  s2 += w * y * y;
                }
                st1[q] += s1;
                st2[q] += s2;
              } else if (last) {
                float mx = 0.f;  // post-ReLU values are >= 0
                float *out = feats + static_cast<size_t>(p + cw_pa) * U + 32 * q + 16 * h + cw_c;
#pragma unroll 4
                for (int r = cw_ra; r < cw_rb; ++r) {
                  mx = fmaxf(mx, col[r * kScrPitch]);
                  if (s_rend[r]) {
                    *out = mx;
                    out += U;
                    mx = 0.f;
                  }
                }
              } else {
                float mx = 0.f;
                int first = cw_ra;
#pragma unroll 4
                for (int r = cw_ra; r < cw_rb; ++r) {
                  mx = fmaxf(mx, col[r * kScrPitch]);
                  if (s_rend[r]) {
                    col[first * kScrPitch] = mx;  // parked in the pillar's first row
                    mx = 0.f;
                    first = r + 1;
                  }
                }
              }
            }
            grp_sync(h);
            if (!last && !stats) {  // max half of the next layer's input: K index = U + unit index
              const float *src = s_scr + fr * kScrPitch + 16 * h;
              float m[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) m[j] = src[j];
              grp_sync(h);  // scratch may be overwritten by the next slab from here on
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) split_tf32(m[j], hi[j], lo[j]);
              tmem_st16(tlane + kColAH + static_cast<uint32_t>(U) + 32u * q + 16u * h, hi);
              tmem_st16(tlane + kColAL + static_cast<uint32_t>(U) + 32u * q + 16u * h, lo);
              tc_wait_st();
              tc_fence_before();
              mbar_arrive(my_slab + 8 * (ev & 7));
              ++ev;
            }
          }
        }
      }
      p += npil;
    }
    // stop signal for the MMA issuer
    epi_sync();
    if (tid == 0) s_misc[7] = 0;
    mbar_arrive(bar_x0);

    if (k.stat_layer >= 0) {
      // deterministic CTA reduction: the 8 ranges' partials of each column are summed in range order
      double *red = reinterpret_cast<double *>(s_scr);  // [8][2][128] = 16 KB
      const int U = k.U[k.stat_layer];
      const int nq = U >> 5;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q < nq) {
          red[(cw_g * 2 + 0) * MBEV_MAX_UNITS + 32 * q + 16 * h + cw_c] = st1[q];
          red[(cw_g * 2 + 1) * MBEV_MAX_UNITS + 32 * q + 16 * h + cw_c] = st2[q];
        }
      }
      epi_sync();
      for (int u = tid; u < U; u += kEpiThreads) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          a += red[(g * 2 + 0) * MBEV_MAX_UNITS + u];
          b += red[(g * 2 + 1) * MBEV_MAX_UNITS + u];
        }
        k.partials[(static_cast<size_t>(blockIdx.x) * 2 + 0) * k.um + u] = a;
        k.partials[(static_cast<size_t>(blockIdx.x) * 2 + 1) * k.um + u] = b;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiThreads / 32) tmem_dealloc(tmem, kTmemCols);
}

// nn.Linear weights (U_l, K_l) -> hi / lo TF32 images in the UMMA K-major no-swizzle layout:
// 16-byte unit (kchunk = kk/4, u) at ((kchunk * U + u) * 4 + kk%4) floats; K zero-padded to Kp; layer 0 in the wide
// column order (kWide*).
struct PrepArgs {
  int L;
  int bf16;          // layers >= 1: one bf16 image at hi[l] (16-byte unit = 8 K elements), no lo image
  int map0[kK0Pad];  // wide layer-0 slot -> Linear.weight column of layer 0, -1 = unused slot (zero column)
  int K[MBEV_MAX_LAYERS], Kp[MBEV_MAX_LAYERS], U[MBEV_MAX_LAYERS];
  const float *w[MBEV_MAX_LAYERS];
  float *hi[MBEV_MAX_LAYERS], *lo[MBEV_MAX_LAYERS];
};

__global__ void k_prep_weights_tc(const PrepArgs a) {
  const int l = blockIdx.y;
  if (l >= a.L) return;
  const int K = a.K[l], Kp = a.Kp[l], U = a.U[l];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < U * Kp; i += gridDim.x * blockDim.x) {
    const int u = i / Kp, kk = i - u * Kp;
    const int col = (l == 0) ? a.map0[kk] : (kk < K ? kk : -1);
    const float v = col >= 0 ? __ldg(a.w[l] + static_cast<size_t>(u) * K + col) : 0.f;
    if (a.bf16 && l > 0) {  // UMMA K-major no-swizzle core matrix: 8 rows x 16 bytes = 8 bf16 along K
      uint16_t *img = reinterpret_cast<uint16_t *>(a.hi[l]);
      uint32_t b;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b) : "f"(0.f), "f"(v));
      img[(((kk >> 3) * U + u) << 3) + (kk & 7)] = static_cast<uint16_t>(b & 0xffffu);
      continue;
    }
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    const int idx = (((kk >> 2) * U + u) << 2) + (kk & 3);
    a.hi[l][idx] = __uint_as_float(hi);
    a.lo[l][idx] = __uint_as_float(lo);
  }
}

struct Plan {
  Kargs k;
  PrepArgs prep;
  size_t ws_bytes;
  int grid;
  int nb;         // blocks of kBlkPillars pillars covering the capacity
  int nbA;        // partition CTAs (kRbBlocks blocks each)
  int *blocksum;  // (nb)
  int *ctot;      // (nbA)
  int *bounds;    // (8 * grid + 1) first pillar of each sub-range (k_pfn_tcw2: one per set and TMEM lane quadrant)
  // k_pfn_tcw2 walks every sub-range sorted by row count (pfn_tcw2.cuh): order[i] = pillar at sorted position i of
  // its sub-range, np_sorted[i] = its num_points (scratch the kernel fills itself)
  int *order, *np_sorted;
  int64_t cap;
};

// MBEV_OK when the stack fits the tensor-core kernel, MBEV_ERR_UNSUPPORTED when it must run on the FMA kernel.
inline int make_plan(const MbevPfnParams *p, int C, int T, int64_t pillar_capacity, void *ws, Plan *out) {
  if (!p || p->num_layers < 1 || p->num_layers > MBEV_MAX_LAYERS) return MBEV_ERR_BAD_ARG;
  const bool bf16 = p->gemm_path == MBEV_GEMM_TCGEN05_BF16;
  if (C < 3 || C > MBEV_MAX_POINT_DIM || T < 1) return MBEV_ERR_UNSUPPORTED;
  if (T + 1 > kRows) return MBEV_ERR_UNSUPPORTED;  // a pillar and its virtual row must fit one chunk
  Kargs &k = out->k;
  k = Kargs();
  PrepArgs &pa = out->prep;
  pa = PrepArgs();
  k.L = pa.L = p->num_layers;
  k.bf16 = pa.bf16 = bf16 ? 1 : 0;
  k.C = C;
  k.T = T;
  k.cluster = p->with_cluster_center != 0;
  k.vcenter = p->with_voxel_center != 0;
  k.dist = p->with_distance != 0;
  k.legacy = p->legacy != 0;
  k.vcd = p->voxel_center_dims;
  if (k.vcenter && k.vcd != 2 && k.vcd != 3) return MBEV_ERR_BAD_ARG;
  k.D0 = C + (k.cluster ? 3 : 0) + (k.vcenter ? k.vcd : 0) + (k.dist ? 1 : 0);
  static_assert(kWideExtra + MBEV_MAX_POINT_DIM - 3 <= kWideCluster && kWideDist < kK0Pad, "wide layer-0 layout");
  for (int j = 0; j < kK0Pad; ++j) pa.map0[j] = -1;
  {
    int d = 0;
    for (int j = 0; j < 3; ++j) pa.map0[j] = d++;
    for (int j = 3; j < C; ++j) pa.map0[kWideExtra + j - 3] = d++;
    if (k.cluster) for (int j = 0; j < 3; ++j) pa.map0[kWideCluster + j] = d++;
    if (k.vcenter) for (int j = 0; j < k.vcd; ++j) pa.map0[kWideCentre + j] = d++;
    if (k.dist) pa.map0[kWideDist] = d++;
  }
  k.vx = p->vx; k.vy = p->vy; k.vz = p->vz;
  k.xo = p->x_offset; k.yo = p->y_offset; k.zo = p->z_offset;
  int um = 0;
  uint32_t off = 0;
  for (int l = 0; l < k.L; ++l) {
    const int U = p->units[l];
    if (U < 32 || (U & 31) || U > MBEV_MAX_UNITS) return MBEV_ERR_UNSUPPORTED;
    const int K = (l == 0) ? k.D0 : 2 * p->units[l - 1];
    if (p->in_dim[l] != K) return MBEV_ERR_BAD_ARG;
    const int Kp = (l == 0) ? kK0Pad : K;
    if (Kp > 128 || (Kp & 7)) return MBEV_ERR_UNSUPPORTED;
    k.U[l] = pa.U[l] = U;
    k.K[l] = pa.K[l] = K;
    k.Kp[l] = pa.Kp[l] = Kp;
    if (bf16 && l > 0) {  // one bf16 image
      if (Kp & 15) return MBEV_ERR_UNSUPPORTED;
      k.w_off[l][0] = k.w_off[l][1] = off;
      off += static_cast<uint32_t>(U) * Kp * 2u;
    } else {
      for (int h = 0; h < 2; ++h) {
        k.w_off[l][h] = off;
        off += static_cast<uint32_t>(U) * Kp * 4u;
      }
    }
    um = std::max(um, U);
  }
  k.um = um;
  k.w_bytes = off;
  // shared memory: [weights image][scratch 128x33][scale/shift][tables][barriers]
  uint32_t o = off;
  o = (o + 127u) & ~127u;
  k.o_scr = o; o += kRows * kScrPitch * 4;
  o = (o + 15u) & ~15u;
  k.o_ss = o; o += k.L * 2 * MBEV_MAX_UNITS * 4;
  k.o_tab = o; o += (kPcap + (kPcap + 4) + kRows + kRows + kRows + kPcap * 4 + kPcap * 4 + 8) * 4;
  o = (o + 15u) & ~15u;
  k.o_bar = o; o += kNumBars * 8 + 8;
  k.smem_bytes = static_cast<int>(o);
  if (k.smem_bytes > kSmemLimit) return MBEV_ERR_UNSUPPORTED;
  // workspace: weight image, per-CTA statistic partials
  Carver cw(ws);
  float *img = cw.take<float>(off / 4);
  k.w_img = img;
  for (int l = 0; l < k.L; ++l) {
    pa.hi[l] = img ? img + k.w_off[l][0] / 4 : nullptr;
    pa.lo[l] = img ? img + k.w_off[l][1] / 4 : nullptr;
  }
  out->grid = kNumSMs;
  k.partials = cw.take<double>(static_cast<size_t>(out->grid) * 2 * um);
  const int64_t nb = (std::max<int64_t>(pillar_capacity, 1) + kBlkPillars - 1) / kBlkPillars;
  if (nb > (1 << 24)) return MBEV_ERR_UNSUPPORTED;
  out->nb = static_cast<int>(nb);
  out->blocksum = cw.take<int>(out->nb);
  out->nbA = (out->nb + kRbBlocks - 1) / kRbBlocks;
  out->ctot = cw.take<int>(out->nbA);
  out->bounds = cw.take<int>(8 * out->grid + 1);
  out->cap = std::max<int64_t>(pillar_capacity, 1);
  out->order = cw.take<int>(static_cast<size_t>(out->cap));
  out->np_sorted = cw.take<int>(static_cast<size_t>(out->cap));
  out->ws_bytes = cw.off;
  return MBEV_OK;
}

inline int launch_prep(const MbevPfnParams *p, Plan &pl, const int32_t *num_points, const int32_t *num_pillars_dev,
                       cudaStream_t stream) {
  for (int l = 0; l < pl.k.L; ++l) {
    if (!p->weight[l]) return MBEV_ERR_BAD_ARG;
    pl.prep.w[l] = p->weight[l];
  }
  k_prep_weights_tc<<<dim3(16, pl.k.L), 256, 0, stream>>>(pl.prep);
  MBEV_CHECK_LAUNCH();
  k_row_blocks<<<pl.nbA, 256, 0, stream>>>(num_points, num_pillars_dev, pl.k.T, pl.nb, pl.blocksum, pl.ctot);
  MBEV_CHECK_LAUNCH();
  k_row_bounds<<<pl.nbA, 256, 0, stream>>>(pl.blocksum, pl.ctot, pl.nbA, pl.nb, num_pillars_dev, num_points, pl.k.T,
                                           8 * pl.grid, pl.bounds);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

}  // namespace tc
}  // namespace mbev
