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
#include "rows_gemm_tc.cuh"

namespace mbev {
namespace {

constexpr int kPeMaxClasses = 64;      // ps <= 8
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

  PeArgs a{};
  if (!pe_gemm_plan(a, E, C)) return MBEV_ERR_UNSUPPORTED;
  a.feats = feats;
  a.coors = coors;
  a.perm = w.perm;
  a.base = w.base;
  a.lnw_cl = ln_weight_cl;
  a.w_img = w_img;
  a.Z = w.Z;
  a.nx = nx;
  a.ncls = ncls;
  const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>(w.slots / kPeRows, kNumSMs)));
  if (C == 128) {
    MBEV_CUDA(cudaFuncSetAttribute(k_pe_gemm<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeSmemLimit));
    k_pe_gemm<128, true><<<grid, kPeThreads, a.smem_bytes, stream>>>(a);
  } else if (C == 64) {
    MBEV_CUDA(cudaFuncSetAttribute(k_pe_gemm<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeSmemLimit));
    k_pe_gemm<64, true><<<grid, kPeThreads, a.smem_bytes, stream>>>(a);
  } else {
    MBEV_CUDA(cudaFuncSetAttribute(k_pe_gemm<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeSmemLimit));
    k_pe_gemm<32, true><<<grid, kPeThreads, a.smem_bytes, stream>>>(a);
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
