// tcgen05 3xTF32 "rows x small matrix" GEMM shared by the pillar patch embedding (patch_embed.cu: gathered pillar rows
// times a position-class weight slice) and the PFN backward (pfn_bwd.cu: recompute Y = X W^T and dX = dY W over the compact
// row space). Z (rows, E) = A (rows, C) * B^T with B (E, C) given as hi / lo TF32 images in the UMMA K-major no-swizzle
// layout; C in {32, 64, 128}, E a multiple of 32 with 2 * E * C * 4 bytes + 32 KB of staging inside shared memory.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace mbev {
namespace {

using namespace tc;

constexpr int kPeRows = 128;           // rows per MMA = TMEM lanes
constexpr int kPeWorkers = 256;        // 8 warps: (TMEM lane quadrant) x (channel half)
constexpr int kPeThreads = kPeWorkers + 32;
constexpr int kPeSmemLimit = 227 * 1024;


// conv weight (E, C, ps, ps) -> per class k = dy * ps + dx the hi / lo TF32 images of W_k (E x C) in the UMMA K-major
// no-swizzle layout of pfn_tc.cuh: float (c, e) at ((c / 4) * E + e) * 4 + c % 4. img: [class][hi | lo][E * C]
// transposed = 1 (ps = 1): w is (C, E) row-major and the image is that of its transpose (dX = dY * W: B = W^T)
__global__ void k_pe_prep_weights(const float *__restrict__ w, const int E, const int C, const int ps, float *__restrict__ img,
                                  const int transposed = 0) {
  const int k = blockIdx.y;
  const int dy = k / ps, dx = k % ps;
  float *hi = img + static_cast<size_t>(k) * 2 * E * C, *lo = hi + static_cast<size_t>(E) * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E * C; i += gridDim.x * blockDim.x) {
    const int e = i / C, c = i - e * C;
    const float v = transposed ? __ldg(w + static_cast<size_t>(c) * E + e) : __ldg(w + ((static_cast<size_t>(e) * C + c) * ps + dy) * ps + dx);
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
  const int *num_rows;  // plain mode (kGather = false): device row count; A = feats (R, C) row-major, Z (R, E)
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
template <int kC, bool kGather>
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
  const int R = kGather ? 0 : *a.num_rows;  // plain rows: the row count lives on the device
  const int nch = kGather ? __ldg(a.base + a.ncls) / kPeRows : (R + kPeRows - 1) / kPeRows;
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
      if (kGather)
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
      int p = -1, cell = 0;
      if (c < c_hi) {
        if (kGather) {
          p = __ldg(a.perm + c * kPeRows + row);
          if (p >= 0) {
            const int4 cc = __ldg(reinterpret_cast<const int4 *>(a.coors) + p);
            cell = cc.z * a.nx + cc.w;
          }
        } else {
          p = c * kPeRows + row;
          if (p >= R) p = -1;
        }
      }
#pragma unroll
      for (int it = 0; it < kIts; ++it) {
        const int r = it * kRPI + grow;
        const int pr = __shfl_sync(0xffffffffu, p, r), cr = __shfl_sync(0xffffffffu, cell, r);
#pragma unroll
        for (int ps = 0; ps < kPasses; ++ps) {
          if (pr >= 0) {
            const int ch = hh * kHalf + ps * kPW + 4 * piece;
            // plain rows are written earlier in the same stream by other kernels: a coherent load, not the read-only path
            const float4 *xp = reinterpret_cast<const float4 *>(a.feats + static_cast<size_t>(pr) * kC + ch);
            const float4 x = kGather ? __ldg(xp) : *xp;
            if (kGather) {
              const float4 y = __ldg(reinterpret_cast<const float4 *>(a.lnw_cl + static_cast<size_t>(cr) * kC + ch));
              g[ps][it] = make_float4(__fmul_rn(x.x, y.x), __fmul_rn(x.y, y.y), __fmul_rn(x.z, y.z), __fmul_rn(x.w, y.w));
            } else {
              g[ps][it] = x;
            }
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


// shared-memory plan of k_pe_gemm for an (E x C) weight slice
inline bool pe_gemm_plan(PeArgs &a, int E, int C) {
  a.C = C;
  a.E = E;
  a.img_bytes = static_cast<uint32_t>(E) * C * 4u;
  a.o_stage = (2u * a.img_bytes + 127u) & ~127u;
  a.o_bar = a.o_stage + (kPeWorkers / 32) * 4096u;
  a.smem_bytes = static_cast<int>(a.o_bar + 64);
  return (C == 32 || C == 64 || C == 128) && E >= 32 && E <= 256 && (E % 32) == 0 && a.smem_bytes <= kPeSmemLimit;
}

// plain-rows launch: Z (R, E) = A (R, C) * B^T, R = *num_rows_dev (<= rows_cap)
inline int launch_rows_gemm(const float *A, const int *num_rows_dev, int64_t rows_cap, int C, int E, const float *w_img,
                            float *Z, cudaStream_t stream) {
  PeArgs a{};
  if (!pe_gemm_plan(a, E, C)) return MBEV_ERR_UNSUPPORTED;
  a.feats = A;
  a.num_rows = num_rows_dev;
  a.w_img = w_img;
  a.Z = Z;
  a.ncls = 1;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((rows_cap + kPeRows - 1) / kPeRows, kNumSMs)));
  if (C == 128) {
    MBEV_CUDA(cudaFuncSetAttribute(k_pe_gemm<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeSmemLimit));
    k_pe_gemm<128, false><<<grid, kPeThreads, a.smem_bytes, stream>>>(a);
  } else if (C == 64) {
    MBEV_CUDA(cudaFuncSetAttribute(k_pe_gemm<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeSmemLimit));
    k_pe_gemm<64, false><<<grid, kPeThreads, a.smem_bytes, stream>>>(a);
  } else {
    MBEV_CUDA(cudaFuncSetAttribute(k_pe_gemm<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeSmemLimit));
    k_pe_gemm<32, false><<<grid, kPeThreads, a.smem_bytes, stream>>>(a);
  }
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

}  // namespace
}  // namespace mbev
