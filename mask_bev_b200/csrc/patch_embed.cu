// F2 — Swin patch embedding that consumes PILLARS instead of the pseudo image.
//
// Replaces, for the consumer at /root/reference/mask_bev/models/networks/swin/swin.py:578-586, 745-746
//     x = LayerNorm([C, ny, nx], eps)(canvas)                  (mask_bev_encoders.py:75, 92)
//     x = PatchEmbed(x)  = Conv2d(C, E, kernel = stride = ps)(corner-padded x).flatten(2).transpose(1, 2) [-> LayerNorm(E)]
// (mmdet 3.x PatchEmbed as imported at swin.py:13; padding='corner', bias=True). The canvas (5.2 GB for 16 frames at
// 800 x 800 x 128) is never written: with xh = (x - mu_b) * rstd_b, x = 0 outside the pillars,
//     token[b, t, :] = P0[t, :] - mu_b rstd_b P1[t, :] + rstd_b * sum_{pillars p in patch t} W_{dy,dx} (f_p * lnw[:, y, x])
// where P0 = conv(ln_bias) + conv bias and P1 = conv(ln_weight) are PARAMETER-ONLY images (prepared once per weight
// update by the host side) and the sparse term is one (pillars x C) x (C x E) product per pillar, grouped by the
// pillar's position inside its patch (ps * ps classes, one E x C weight slice each):
//   k_pe_hist / k_pe_bases / k_pe_place : counting sort of the pillars by class into 128-row aligned slabs
//   k_pe_gemm   : tcgen05 3xTF32 (fp32 parity as in K2): A = f * lnw built in registers -> TMEM, B = the class's weight
//                 slice (UMMA K-major image, resident in shared memory while the CTA stays in the class), D -> Z[pillar]
//   k_pe_tokens : a warp per token sums the Z rows of its patch in cell order (fixed order: run-to-run identical),
//                 adds the parameter images, applies the patch LayerNorm and writes (B, Hp*Wp, E).
#include <algorithm>

#include "common.cuh"
#include "ln_stats.cuh"
#include "tc_ptx.cuh"

namespace mbev {
namespace {

using namespace tc;

constexpr int kPeRows = 128;           // rows per MMA = TMEM lanes
constexpr int kPeWorkers = 256;        // 8 warps: (TMEM lane quadrant) x (channel half)
constexpr int kPeThreads = kPeWorkers + 32;
constexpr int kPeMaxClasses = 64;      // ps <= 8
constexpr int kPeSmemLimit = 227 * 1024;
constexpr int kPlaceItems = 8, kPlaceThreads = 256;

__device__ __forceinline__ int pe_class(const int4 cc, const int ps) { return (cc.z % ps) * ps + (cc.w % ps); }  // (b, z, y, x)

__global__ void __launch_bounds__(256)
k_pe_hist(const int *__restrict__ coors, const int *__restrict__ num_pillars, const int ps, int *__restrict__ count) {
  __shared__ int s_c[kPeMaxClasses];
  if (threadIdx.x < kPeMaxClasses) s_c[threadIdx.x] = 0;
  __syncthreads();
  const int P = *num_pillars;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x)
    atomicAdd(&s_c[pe_class(__ldg(reinterpret_cast<const int4 *>(coors) + p), ps)], 1);
  __syncthreads();
  if (threadIdx.x < ps * ps && s_c[threadIdx.x]) atomicAdd(count + threadIdx.x, s_c[threadIdx.x]);
}

// base[k] = first slot of class k (128-aligned), base[ncls] = slots in use; cursors reset
__global__ void k_pe_bases(const int *__restrict__ count, const int ncls, int *__restrict__ base, int *__restrict__ cursor) {
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int k = 0; k < ncls; ++k) {
      base[k] = acc;
      acc += (count[k] + kPeRows - 1) / kPeRows * kPeRows;
    }
    base[ncls] = acc;
  }
  if (threadIdx.x < ncls) cursor[threadIdx.x] = 0;
}

// slot of every pillar: a CTA reserves one range per class (one global atomic per class and CTA), its pillars take
// the places inside by shared-memory atomics. The order inside a class is arbitrary — nothing downstream depends on it.
__global__ void __launch_bounds__(kPlaceThreads)
k_pe_place(const int *__restrict__ coors, const int *__restrict__ num_pillars, const int ps, const int *__restrict__ base,
           int *__restrict__ cursor, int *__restrict__ perm) {
  __shared__ int s_c[kPeMaxClasses], s_b[kPeMaxClasses];
  if (threadIdx.x < kPeMaxClasses) s_c[threadIdx.x] = 0;
  __syncthreads();
  const int P = *num_pillars;
  const int p0 = blockIdx.x * (kPlaceThreads * kPlaceItems);
  int cls[kPlaceItems], rk[kPlaceItems];
#pragma unroll
  for (int i = 0; i < kPlaceItems; ++i) {
    const int p = p0 + i * kPlaceThreads + threadIdx.x;
    cls[i] = -1;
    if (p < P) {
      cls[i] = pe_class(__ldg(reinterpret_cast<const int4 *>(coors) + p), ps);
      rk[i] = atomicAdd(&s_c[cls[i]], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < ps * ps) s_b[threadIdx.x] = s_c[threadIdx.x] ? base[threadIdx.x] + atomicAdd(cursor + threadIdx.x, s_c[threadIdx.x]) : 0;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kPlaceItems; ++i) {
    if (cls[i] < 0) continue;
    const int p = p0 + i * kPlaceThreads + threadIdx.x;
    const int slot = s_b[cls[i]] + rk[i];
    perm[slot] = p;
  }
}

// conv weight (E, C, ps, ps) -> per class k = dy * ps + dx the hi / lo TF32 images of W_k (E x C) in the UMMA K-major
// no-swizzle layout of pfn_tc.cuh: float (c, e) at ((c / 4) * E + e) * 4 + c % 4. img: [class][hi | lo][E * C]
__global__ void k_pe_prep_weights(const float *__restrict__ w, const int E, const int C, const int ps, float *__restrict__ img) {
  const int k = blockIdx.y;
  const int dy = k / ps, dx = k % ps;
  float *hi = img + static_cast<size_t>(k) * 2 * E * C, *lo = hi + static_cast<size_t>(E) * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E * C; i += gridDim.x * blockDim.x) {
    const int e = i / C, c = i - e * C;
    const float v = __ldg(w + ((static_cast<size_t>(e) * C + c) * ps + dy) * ps + dx);
    uint32_t h, l;
    split_tf32(v, h, l);
    const int idx = (((c >> 2) * E + e) << 2) + (c & 3);
    hi[idx] = __uint_as_float(h);
    lo[idx] = __uint_as_float(l);
  }
}

struct PeArgs {
  const float *feats;
  const int *coors, *perm, *base;
  const float *lnw_cl;  // (ny * nx, C)
  const float *w_img;
  float *Z;             // (pillar capacity, E): row = pillar id
  int C, E, nx, ncls;
  uint32_t img_bytes;   // E * C * 4: one hi or lo image
  uint32_t o_stage, o_bar;  // byte offsets in dynamic shared memory: per-warp staging tiles, barriers
  int smem_bytes;
};

// One chunk = 128 consecutive slots of one class. Worker warps build A = f * lnw (split to TF32 hi / lo) in tensor
// memory and drain D to Z, both through a swizzled staging tile so that global accesses are whole lines. The issuer
// warp keeps the class's weight slice in shared memory (reloaded by bulk copies when the CTA crosses a class border,
// 16 times per launch in total) and issues 3 MMAs per K-step with uniform operands. Phases of a chunk are serial
// (A and D fill the 512 TMEM columns); the kernel is bound by the gather / store bytes, not by the tensor pipe.
template <int kC>
__global__ void __launch_bounds__(kPeThreads, 1)
k_pe_gemm(const __grid_constant__ PeArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kHalf = kC / 2;  // channels per worker thread
  constexpr uint32_t kColAH = 0, kColAL = 128, kColD = 256;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + a.o_bar);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 3);
  const uint32_t bar_w = smem_u32(s_bar), bar_a = bar_w + 8, bar_d = bar_w + 16;
  const uint32_t smem_base = smem_u32(smem_raw);
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_a, kPeWorkers / 32);
    mbar_init(bar_d, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kPeWorkers / 32) tmem_alloc(smem_u32(s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const int E = a.E;
  const int nch = __ldg(a.base + a.ncls) / kPeRows;
  const int c_lo = static_cast<int>(static_cast<long long>(nch) * blockIdx.x / gridDim.x);
  const int c_hi = static_cast<int>(static_cast<long long>(nch) * (blockIdx.x + 1) / gridDim.x);

  if (warp == kPeWorkers / 32) {
    // =========================================== MMA issuer ===================================================
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = make_idesc(E);
    const uint32_t lbo = static_cast<uint32_t>(E) * 16u;
    const uint64_t dh0 = make_bdesc(smem_base, lbo, 128u);
    const uint64_t dl0 = make_bdesc(smem_base + a.img_bytes, lbo, 128u);
    const uint32_t dstep = lbo >> 3;
    uint32_t par_w = 0, par_a = 0, par_d = 0;
    int cur = -1;
    for (int c = c_lo; c < c_hi; ++c) {
      int cls = 0;
      while (cls + 1 < a.ncls && __ldg(a.base + cls + 1) <= c * kPeRows) ++cls;
      if (cls != cur) {  // previous chunk's MMAs have retired (bar_d below): the weight slab may be overwritten
        cur = cls;
        if (elect_one()) {
          const char *src = reinterpret_cast<const char *>(a.w_img) + static_cast<size_t>(cls) * 2 * a.img_bytes;
          const uint32_t bytes = 2 * a.img_bytes;
          mbar_expect_tx(bar_w, bytes);
          for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(smem_base + off, src + off, min(32768u, bytes - off), bar_w);
        }
        __syncwarp();
        mbar_wait(bar_w, par_w);
        par_w ^= 1u;
      }
      mbar_wait(bar_a, par_a);
      par_a ^= 1u;
      tc_fence_after();
      if (elect_one()) {
        uint32_t acc = 0;
#pragma unroll 2
        for (int j = 0; j < kC / 8; ++j) {  // al*wh, ah*wl, ah*wh : small terms first
          const uint64_t dh = dh0 + dstep * j, dl = dl0 + dstep * j;
          mma_tf32_ts(tmem_u + kColD, tmem_u + kColAL + 8u * j, dh, idesc, acc);
          mma_tf32_ts(tmem_u + kColD, tmem_u + kColAH + 8u * j, dl, idesc, 1u);
          mma_tf32_ts(tmem_u + kColD, tmem_u + kColAH + 8u * j, dh, idesc, 1u);
          acc = 1;
        }
        tc_commit(bar_d);
      }
      __syncwarp();
      mbar_wait(bar_d, par_d);
      par_d ^= 1u;
    }
  } else {
    // =========================================== workers =======================================================
    // Warp (quad, hh) owns rows [32 quad, 32 quad + 32) x channels [hh kHalf, (hh + 1) kHalf). Tensor memory wants one row
    // per lane, global memory wants a warp instruction to cover whole 128-byte lines: a thread-per-row float4 access
    // touches 32 lines per instruction (32 L1 wavefronts; the first form of this kernel spent ~14 k of its 30 k cycles
    // per chunk there). So rows travel through a per-warp 4 KB staging tile: global <-> tile with 8 lanes per row (4
    // rows x 128 contiguous bytes per instruction), tile <-> registers with a lane per row; 16-byte units are XOR-
    // swizzled by the row so that both sides are bank-conflict free.
    constexpr int kPW = kHalf >= 32 ? 32 : 16;      // channels per pass
    constexpr int kPasses = kHalf / kPW;
    constexpr int kLPR = kPW / 4;                   // lanes per row on the global side (16-byte pieces)
    constexpr int kRPI = 32 / kLPR;                 // rows per instruction
    constexpr int kIts = 32 / kRPI;                 // instructions per pass and tensor
    const int quad = warp & 3, hh = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t tl = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    float4 *stage = reinterpret_cast<float4 *>(smem_raw + a.o_stage) + warp * 256;  // 32 rows x 8 units of 16 bytes
    const int grow = lane / kLPR, piece = lane % kLPR;  // global side: row inside the instruction, 16-byte piece
    auto unit = [](int r, int j) { return r * 8 + (j ^ (r & 7)); };
    uint32_t par_d = 0;
    float4 g[kPasses][kIts];
    // products f * lnw of the chunk's rows, requested a whole chunk ahead (under the previous chunk's MMAs and drain)
    auto request = [&](int c) -> int {
      const int p = (c < c_hi) ? __ldg(a.perm + c * kPeRows + row) : -1;
      int cell = 0;
      if (p >= 0) {
        const int4 cc = __ldg(reinterpret_cast<const int4 *>(a.coors) + p);
        cell = cc.z * a.nx + cc.w;
      }
#pragma unroll
      for (int it = 0; it < kIts; ++it) {
        const int r = it * kRPI + grow;
        const int pr = __shfl_sync(0xffffffffu, p, r), cr = __shfl_sync(0xffffffffu, cell, r);
#pragma unroll
        for (int ps = 0; ps < kPasses; ++ps) {
          if (pr >= 0) {
            const int ch = hh * kHalf + ps * kPW + 4 * piece;
            const float4 x = __ldg(reinterpret_cast<const float4 *>(a.feats + static_cast<size_t>(pr) * kC + ch));
            const float4 y = __ldg(reinterpret_cast<const float4 *>(a.lnw_cl + static_cast<size_t>(cr) * kC + ch));
            g[ps][it] = make_float4(__fmul_rn(x.x, y.x), __fmul_rn(x.y, y.y), __fmul_rn(x.z, y.z), __fmul_rn(x.w, y.w));
          } else {
            g[ps][it] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      return p;
    };
    int p_next = request(c_lo);
    for (int c = c_lo; c < c_hi; ++c) {
      const int p = p_next;
#pragma unroll
      for (int ps = 0; ps < kPasses; ++ps) {
#pragma unroll
        for (int it = 0; it < kIts; ++it) stage[unit(it * kRPI + grow, piece)] = g[ps][it];
        __syncwarp();
#pragma unroll
        for (int jb = 0; jb < kPW / 16; ++jb) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 v = stage[unit(lane, 4 * jb + q)];
            split_tf32_alu(v.x, hi[4 * q + 0], lo[4 * q + 0]);
            split_tf32_alu(v.y, hi[4 * q + 1], lo[4 * q + 1]);
            split_tf32_alu(v.z, hi[4 * q + 2], lo[4 * q + 2]);
            split_tf32_alu(v.w, hi[4 * q + 3], lo[4 * q + 3]);
          }
          const uint32_t col = static_cast<uint32_t>(hh * kHalf + ps * kPW + 16 * jb);
          tmem_st16(tl + kColAH + col, hi);
          tmem_st16(tl + kColAL + col, lo);
        }
        __syncwarp();
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a);
      p_next = request(c + 1);
      mbar_wait(bar_d, par_d);
      par_d ^= 1u;
      tc_fence_after();
      // drain: D columns [hh E/2, (hh + 1) E/2) of the warp's 32 rows -> tile -> 128-byte row pieces of Z[pillar]
      const int Eh = E >> 1;
      for (int c0 = 0; c0 < Eh; c0 += 32) {
        const bool wide = c0 + 32 <= Eh;  // 32-column pass, or the 16-column tail (E / 2 is a multiple of 16)
        if (wide) {
          uint32_t v[32];
          tmem_ld32(tl + kColD + static_cast<uint32_t>(hh * Eh + c0), v);
          tc_wait_ld();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            stage[unit(lane, q)] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                               __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        } else {
          uint32_t v[16];
          tmem_ld16(tl + kColD + static_cast<uint32_t>(hh * Eh + c0), v);
          tc_wait_ld();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            stage[unit(lane, q)] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                               __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
        __syncwarp();
        const int lpr = wide ? 8 : 4, rpi = 32 / lpr;
        const int dr = lane / lpr, dp = lane % lpr;
        for (int it = 0; it < 32 / rpi; ++it) {
          const int r = it * rpi + dr;
          const int pr = __shfl_sync(0xffffffffu, p, r);
          const float4 v = stage[unit(r, dp)];
          if (pr >= 0) *reinterpret_cast<float4 *>(a.Z + static_cast<size_t>(pr) * E + hh * Eh + c0 + 4 * dp) = v;
        }
        __syncwarp();
      }
      tc_fence_before();  // D is read: the next chunk's MMAs (ordered after this warp's next arrival) may overwrite it
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kPeWorkers / 32) tmem_dealloc(tmem, 512);
}

// A warp per kTok consecutive tokens; lane e + 32 j holds embedding channel e + 32 j. The cell-table lookups of all
// the warp's tokens are requested together, then the Z rows (row = pillar id): the kernel is a chain of dependent
// gathers and would otherwise wait out one memory round trip per token and stage.
constexpr int kTok = 4;
template <int kJ>
__global__ void __launch_bounds__(256)
k_pe_tokens(const float *__restrict__ Z, const int *__restrict__ table, const float2 *__restrict__ stats,
            const float *__restrict__ P0, const float *__restrict__ P1, const float *__restrict__ norm_w,
            const float *__restrict__ norm_b, const float norm_eps, const int batch, const int ny, const int nx,
            const int ps, const int Hp, const int Wp, const int E, float *__restrict__ tokens) {
  const int lane = threadIdx.x & 31;
  // token position major, frame group minor: the parameter images P0 / P1 (61 MB at 200 x 200 x 192) are read once per
  // position and kTok frames, and the warps that share a position run next to each other (L2 hits) — frame-major order
  // re-streams both images from HBM for every frame
  const long long g = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int HW = Hp * Wp;
  const int nfg = (batch + kTok - 1) / kTok;
  if (g >= static_cast<long long>(HW) * nfg) return;
  const int tt = static_cast<int>(g / nfg), b0 = static_cast<int>(g - static_cast<long long>(tt) * nfg) * kTok;
  int bq[kTok], ttq[kTok];
#pragma unroll
  for (int i = 0; i < kTok; ++i) {
    bq[i] = min(b0 + i, batch - 1);
    ttq[i] = tt;
  }
  float acc[kTok][kJ];
#pragma unroll
  for (int i = 0; i < kTok; ++i)
#pragma unroll
    for (int j = 0; j < kJ; ++j) acc[i][j] = 0.f;
  const int ncell = ps * ps;
  for (int c0 = 0; c0 < ncell; c0 += 32) {  // cells of the patch in (dy, dx) order
    const int ci = c0 + lane;
    int pid[kTok];
#pragma unroll
    for (int i = 0; i < kTok; ++i) {
      pid[i] = -1;
      if (ci < ncell) {
        const int py = ttq[i] / Wp, px = ttq[i] - py * Wp;
        const int y = py * ps + ci / ps, x = px * ps + ci % ps;
        if (y < ny && x < nx) pid[i] = __ldg(table + static_cast<size_t>(bq[i]) * ny * nx + static_cast<size_t>(y) * nx + x);
      }
    }
#pragma unroll
    for (int i = 0; i < kTok; ++i) {
      unsigned m = __ballot_sync(0xffffffffu, pid[i] >= 0);
      while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        const int s = __shfl_sync(0xffffffffu, pid[i], l);
        const float *z = Z + static_cast<size_t>(s) * E + lane;
#pragma unroll
        for (int j = 0; j < kJ; ++j)
          if (lane + 32 * j < E) acc[i][j] = __fadd_rn(acc[i][j], __ldg(z + 32 * j));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kTok; ++i) {
    if (b0 + i >= batch) break;
    const float2 st = stats[bq[i]];
    const float nmr = -st.x * st.y;  // -mean * rstd
    float y[kJ];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const int e = lane + 32 * j;
      y[j] = 0.f;
      if (e < E) {
        const size_t o = static_cast<size_t>(ttq[i]) * E + e;
        y[j] = fmaf(st.y, acc[i][j], fmaf(nmr, __ldg(P1 + o), __ldg(P0 + o)));
        sum += y[j];
      }
    }
    if (norm_w) {  // nn.LayerNorm(E): biased variance, two passes
#pragma unroll
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / static_cast<float>(E);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < kJ; ++j)
        if (lane + 32 * j < E) {
          const float d = y[j] - mean;
          sq = fmaf(d, d, sq);
        }
#pragma unroll
      for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq / static_cast<float>(E) + norm_eps);
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const int e = lane + 32 * j;
        if (e < E) y[j] = fmaf((y[j] - mean) * rstd, __ldg(norm_w + e), __ldg(norm_b + e));
      }
    }
    float *out = tokens + (static_cast<size_t>(bq[i]) * HW + tt) * E + lane;
#pragma unroll
    for (int j = 0; j < kJ; ++j)
      if (lane + 32 * j < E) out[32 * j] = y[j];
  }
}

struct PeWs {
  int *count, *base, *cursor, *perm;
  float *Z;
  double2 *partial;
  size_t slots, bytes;
};

PeWs carve_pe(void *ws, int batch, int64_t cap, int ncls, int E) {
  Carver c(ws);
  PeWs w;
  const size_t P = static_cast<size_t>(cap > 0 ? cap : 1);
  w.slots = (P + static_cast<size_t>(ncls) * (kPeRows - 1) + kPeRows - 1) / kPeRows * kPeRows;
  w.count = c.take<int>(kPeMaxClasses);
  w.base = c.take<int>(kPeMaxClasses + 1);
  w.cursor = c.take<int>(kPeMaxClasses);
  w.perm = c.take<int>(w.slots);
  w.Z = c.take<float>(P * static_cast<size_t>(E));
  w.partial = c.take<double2>(static_cast<size_t>(batch) * kStatBlocks);
  w.bytes = c.off;
  return w;
}

bool pe_shape_ok(int batch, int C, int ny, int nx, int ps, int E) {
  if (batch < 1 || batch > MBEV_MAX_BATCH || ny < 1 || nx < 1 || ps < 1 || ps * ps > kPeMaxClasses) return false;
  if (C != 32 && C != 64 && C != 128) return false;            // A = hi + lo images of K = C columns in tensor memory
  if (E < 32 || E > 256 || (E % 32)) return false;             // two 16-column-batched halves; N of one tcgen05.mma
  if (2u * static_cast<uint32_t>(E) * C * 4u + (kPeWorkers / 32) * 4096u + 192u > static_cast<uint32_t>(kPeSmemLimit)) return false;  // weight slab + staging
  if (static_cast<int64_t>(ny) * nx * batch > 0x7fffffffLL) return false;
  return true;
}

}  // namespace
}  // namespace mbev

using namespace mbev;

extern "C" int mbev_patch_embed_supported(int batch, int C, int ny, int nx, int patch, int embed_dims) {
  return pe_shape_ok(batch, C, ny, nx, patch, embed_dims) ? 1 : 0;
}

extern "C" int mbev_patch_embed_workspace_bytes(int batch, int64_t pillar_capacity, int patch, int embed_dims, size_t *bytes) {
  if (!bytes || batch < 1 || batch > MBEV_MAX_BATCH || pillar_capacity < 0 || patch < 1 || patch * patch > kPeMaxClasses ||
      embed_dims < 1)
    return MBEV_ERR_BAD_ARG;
  *bytes = carve_pe(nullptr, batch, pillar_capacity, patch * patch, embed_dims).bytes;
  return MBEV_OK;
}

extern "C" int mbev_patch_embed_prepare_weights(const float *conv_weight, int embed_dims, int C, int patch, float *w_img,
                                                void *stream_) {
  if (!conv_weight || !w_img || embed_dims < 1 || C < 1 || patch < 1 || patch * patch > kPeMaxClasses) return MBEV_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  k_pe_prep_weights<<<dim3(std::max(1, (embed_dims * C + 255) / 256), patch * patch), 256, 0, stream>>>(conv_weight, embed_dims, C,
                                                                                                    patch, w_img);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

extern "C" int mbev_patch_embed_forward(const float *feats, const int32_t *coors, const int32_t *cell_table,
                                        const int32_t *pillar_base, int64_t pillar_capacity, int batch, int C, int ny,
                                        int nx, int patch, int embed_dims, const float *ln_weight_cl, float ln_eps,
                                        const float *w_img, const float *p0, const float *p1, const float *norm_weight,
                                        const float *norm_bias, float norm_eps, float *tokens, float *stats_out,
                                        void *workspace, size_t workspace_bytes, void *stream_) {
  if (!coors || !cell_table || !pillar_base || !ln_weight_cl || !w_img || !p0 || !p1 || !tokens || !stats_out || !workspace)
    return MBEV_ERR_BAD_ARG;
  if ((norm_weight == nullptr) != (norm_bias == nullptr) || !(ln_eps >= 0.f) || pillar_capacity < 0) return MBEV_ERR_BAD_ARG;
  if (!pe_shape_ok(batch, C, ny, nx, patch, embed_dims)) return MBEV_ERR_UNSUPPORTED;
  if (pillar_capacity > 0 && !feats) return MBEV_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(feats) | reinterpret_cast<uintptr_t>(ln_weight_cl) | reinterpret_cast<uintptr_t>(w_img)) & 15)
    return MBEV_ERR_UNSUPPORTED;
  const int ncls = patch * patch, E = embed_dims;
  const PeWs w = carve_pe(workspace, batch, pillar_capacity, ncls, E);
  if (workspace_bytes < w.bytes) return MBEV_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int *num_pillars = pillar_base + batch;
  const int Hp = (ny + patch - 1) / patch, Wp = (nx + patch - 1) / patch;

  // LayerNorm statistics of the (never written) canvas, from the pillar rows
  k_ln_partials<<<dim3(kStatBlocks, batch), kStatThreads, 0, stream>>>(feats, pillar_base, C, w.partial);
  MBEV_CHECK_LAUNCH();
  float2 *stats = reinterpret_cast<float2 *>(stats_out);
  k_ln_finalize<<<(batch + 127) / 128, 128, 0, stream>>>(w.partial, batch, static_cast<double>(C) * ny * nx,
                                                        static_cast<double>(ln_eps), stats);
  MBEV_CHECK_LAUNCH();

  // pillars -> class-sorted, 128-aligned slabs
  MBEV_CUDA(cudaMemsetAsync(w.count, 0, sizeof(int) * kPeMaxClasses, stream));
  MBEV_CUDA(cudaMemsetAsync(w.perm, 0xff, sizeof(int) * w.slots, stream));
  const int cap = static_cast<int>(std::max<int64_t>(pillar_capacity, 1));
  k_pe_hist<<<std::max(1, std::min((cap + 2047) / 2048, kNumSMs * 4)), 256, 0, stream>>>(coors, num_pillars, patch, w.count);
  MBEV_CHECK_LAUNCH();
  k_pe_bases<<<1, kPeMaxClasses, 0, stream>>>(w.count, ncls, w.base, w.cursor);
  MBEV_CHECK_LAUNCH();
  k_pe_place<<<(cap + kPlaceThreads * kPlaceItems - 1) / (kPlaceThreads * kPlaceItems), kPlaceThreads, 0, stream>>>(
      coors, num_pillars, patch, w.base, w.cursor, w.perm);
  MBEV_CHECK_LAUNCH();

  PeArgs a;
  a.feats = feats;
  a.coors = coors;
  a.perm = w.perm;
  a.base = w.base;
  a.lnw_cl = ln_weight_cl;
  a.w_img = w_img;
  a.Z = w.Z;
  a.C = C;
  a.E = E;
  a.nx = nx;
  a.ncls = ncls;
  a.img_bytes = static_cast<uint32_t>(E) * C * 4u;
  a.o_stage = (2u * a.img_bytes + 127u) & ~127u;
  a.o_bar = a.o_stage + (kPeWorkers / 32) * 4096u;
  a.smem_bytes = static_cast<int>(a.o_bar + 64);
  const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>(w.slots / kPeRows, kNumSMs)));
  if (C == 128) {
    MBEV_CUDA(cudaFuncSetAttribute(k_pe_gemm<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeSmemLimit));
    k_pe_gemm<128><<<grid, kPeThreads, a.smem_bytes, stream>>>(a);
  } else if (C == 64) {
    MBEV_CUDA(cudaFuncSetAttribute(k_pe_gemm<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeSmemLimit));
    k_pe_gemm<64><<<grid, kPeThreads, a.smem_bytes, stream>>>(a);
  } else {
    MBEV_CUDA(cudaFuncSetAttribute(k_pe_gemm<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeSmemLimit));
    k_pe_gemm<32><<<grid, kPeThreads, a.smem_bytes, stream>>>(a);
  }
  MBEV_CHECK_LAUNCH();

  const long long nwarps = static_cast<long long>(Hp) * Wp * ((batch + kTok - 1) / kTok);
  const int blocks = static_cast<int>((nwarps + 7) / 8);
  const int J = (E + 31) / 32;
#define MBEV_PE_TOK(JJ)                                                                                                  \
  k_pe_tokens<JJ><<<blocks, 256, 0, stream>>>(w.Z, cell_table, stats, p0, p1, norm_weight, norm_bias, norm_eps, batch, \
                                             ny, nx, patch, Hp, Wp, E, tokens)
  switch (J) {
    case 1: MBEV_PE_TOK(1); break;
    case 2: MBEV_PE_TOK(2); break;
    case 3: MBEV_PE_TOK(3); break;
    case 4: MBEV_PE_TOK(4); break;
    case 5: MBEV_PE_TOK(5); break;
    case 6: MBEV_PE_TOK(6); break;
    case 7: MBEV_PE_TOK(7); break;
    default: MBEV_PE_TOK(8); break;
  }
#undef MBEV_PE_TOK
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}
