// K1 — hard voxelisation of a batch of LiDAR frames into pillars (sm_100a).
//
// Replaces mask_bev_encoders.py:95-117 (per-frame strict range filter + mmcv.ops.Voxelization + batch
// index padding). Semantics: SURVEY.md A.1/A.2 — pillars in order of first appearance, slots in input
// order, first T points per pillar, first V pillars per frame.
//
// Upstream's deterministic CUDA path is an O(N^2) scan plus a <<<1,1>>> serial walk. Here:
//   k_assign : one thread per point (float4 load), range test + IEEE divide/floor, then ONE atomicExch per
//              point threads it onto a per-cell linked list (head table = the cell table, L2 resident).
//              Points are visited in reverse so lists come out roughly ascending.
//   k_rank   : each point walks its cell's list counting smaller row indices -> its input-order rank
//              (early exit once T smaller are seen); rank-0 points are pillar heads; warp ballot packs the
//              head flags into a bitmask.
//   k_scan_* : two-level popcount scan of the bitmask (small CTAs) -> pillar number of every head in appearance
//              order, per-frame pillar counts clipped to V, global pillar bases.
//   k_emit   : kept points write their slot; heads write coors / num_points and turn the cell table entry
//              into the global pillar id (the scatter's inverse map / occupancy mask).
// All ordering comes from row indices, never from atomic arrival order, so results are bit-reproducible.
#include <algorithm>

#include "common.cuh"

namespace mbev {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ bool point_cell(const float x, const float y, const float z, const GeoK &g, int &cell) {
  if (g.strict) {
    // mask_bev_encoders.py:113-117 — strict compares in float32; NaN fails
    if (!(g.lo[0] < x && x < g.hi[0] && g.lo[1] < y && y < g.hi[1] && g.lo[2] < z && z < g.hi[2])) return false;
  }
  // mmcv dynamic_voxelize: floor((p - lo) / vs) with float32 subtract and a true IEEE divide
  const float qx = floorf(__fdiv_rn(__fsub_rn(x, g.lo[0]), g.vs[0]));
  const float qy = floorf(__fdiv_rn(__fsub_rn(y, g.lo[1]), g.vs[1]));
  const float qz = floorf(__fdiv_rn(__fsub_rn(z, g.lo[2]), g.vs[2]));
  if (!(qx >= 0.f && qx < static_cast<float>(g.nx))) return false;
  if (!(qy >= 0.f && qy < static_cast<float>(g.ny))) return false;
  if (!(qz >= 0.f && qz < static_cast<float>(g.nz))) return false;
  cell = (static_cast<int>(qz) * g.ny + static_cast<int>(qy)) * g.nx + static_cast<int>(qx);
  return true;
}

template <int CSTATIC>
__global__ void __launch_bounds__(kThreads)
k_assign(const float *__restrict__ pts, const int total, const __grid_constant__ Frames fr,
         const __grid_constant__ GeoK g, int *__restrict__ head, int *__restrict__ next,
         int *__restrict__ cellid) {
  __shared__ int s_off[MBEV_MAX_BATCH + 1];
  for (int t = threadIdx.x; t <= fr.batch; t += kThreads) s_off[t] = fr.off[t];
  __syncthreads();
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t >= total) return;
  const int i = total - 1 - t;  // reverse visit order: LIFO lists end up ~ascending from the head
  float x, y, z;
  if (CSTATIC == 4) {
    const float4 p = __ldg(reinterpret_cast<const float4 *>(pts) + i);
    x = p.x; y = p.y; z = p.z;
  } else {
    const float *p = pts + static_cast<size_t>(i) * g.C;
    x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
  }
  int cell;
  if (!point_cell(x, y, z, g, cell)) {
    cellid[i] = -1;
    return;
  }
  const int f = frame_of(s_off, fr.batch, i);
  const int gc = f * g.cells + cell;
  cellid[i] = gc;
  next[i] = atomicExch(head + gc, i);
}

// ---- F4: augmentations in K1's load stage -------------------------------------------------------
// The point side of /root/reference/mask_bev/augmentations/semantic_kitti_mask_augmentations.py in the order the
// training configs list them (configs/training/semantic_kitti/01*.yml:34-49): RandomDropPoints (:152-162) -> Flip
// (:44-56) -> RandomRotate (:73-101) -> JitterPoints (:116-149). Per-frame decisions and angles are drawn on the host
// (MbevFrameAugment); the per-point randomness is either given (drop_u / noise: bit-exact replay of a numpy stream)
// or generated here from (seed, point row) with Philox4x32-10. A dropped point is simply not voxelised: input order
// of the others is what the rank walk uses, exactly as if the row had been removed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (static_cast<float>(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0, 1)

__global__ void __launch_bounds__(kThreads)
k_assign_aug(const float *__restrict__ pts, const int total, const __grid_constant__ Frames fr,
             const __grid_constant__ GeoK g, const MbevFrameAugment *__restrict__ fa, const float *__restrict__ drop_u,
             const double *__restrict__ noise, const unsigned long long seed, float *__restrict__ pts_out,
             int *__restrict__ head, int *__restrict__ next, int *__restrict__ cellid) {
  __shared__ int s_off[MBEV_MAX_BATCH + 1];
  for (int t = threadIdx.x; t <= fr.batch; t += kThreads) s_off[t] = fr.off[t];
  __syncthreads();
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t >= total) return;
  const int i = total - 1 - t;  // reverse visit order, as k_assign
  const int f = frame_of(s_off, fr.batch, i);
  const MbevFrameAugment a = fa[f];
  float v[MBEV_MAX_POINT_DIM];
  const float *p = pts + static_cast<size_t>(i) * g.C;
#pragma unroll
  for (int j = 0; j < MBEV_MAX_POINT_DIM; ++j) v[j] = j < g.C ? __ldg(p + j) : 0.f;
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  bool dropped = false;
  if (a.drop_prob > 0.f) {  // keep = u >= p  (:158)
    const float u = drop_u ? __ldg(drop_u + i) : u01(philox4x32_10(make_uint4(static_cast<uint32_t>(i), 0u, 0u, 0u), key).x);
    dropped = u < a.drop_prob;
  }
  if (a.flip_x) v[0] = -v[0];  // (:51, :54)
  if (a.flip_y) v[1] = -v[1];
  if (a.rotate) {  // float64 R @ (x, y, z, 1), stored back into the float32 cloud (:88-97)
    const double x = static_cast<double>(v[0]), y = static_cast<double>(v[1]);
    // (+ 0.0: the matrix rows' 0 * z + 0 * 1 terms turn a -0 result into +0, as numpy's product does)
    v[0] = static_cast<float>(__dadd_rn(__dadd_rn(__dmul_rn(a.cos_t, x), __dmul_rn(-a.sin_t, y)), 0.0));
    v[1] = static_cast<float>(__dadd_rn(__dadd_rn(__dmul_rn(a.sin_t, x), __dmul_rn(a.cos_t, y)), 0.0));
    v[2] = static_cast<float>(__dadd_rn(static_cast<double>(v[2]), 0.0));
  }
  if (a.jitter) {  // cloud += noise (float64 sum rounded to float32), intensity clipped to [0, 1] (:134-148)
    if (noise) {
      const double *n = noise + static_cast<size_t>(i) * g.C;
#pragma unroll
      for (int j = 0; j < MBEV_MAX_POINT_DIM; ++j)
        if (j < g.C) v[j] = static_cast<float>(__dadd_rn(static_cast<double>(v[j]), __ldg(n + j)));
    } else {
      const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(i), 1u, 0u, 0u), key);
      const float m0 = sqrtf(-2.f * logf(u01(r.x))), m1 = sqrtf(-2.f * logf(u01(r.z)));
      float z4[4];
      sincospif(2.f * u01(r.y), &z4[1], &z4[0]);
      sincospif(2.f * u01(r.w), &z4[3], &z4[2]);
      z4[0] *= m0; z4[1] *= m0; z4[2] *= m1; z4[3] *= m1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j >= g.C) break;
        float n = z4[j] * a.jitter_std[j];
        if (a.jitter_max[j] > 0.f) n = fminf(fmaxf(n, -a.jitter_max[j]), a.jitter_max[j]);
        v[j] += n;
      }
    }
    if (g.C > 3) v[3] = fminf(fmaxf(v[3], 0.f), 1.f);
  }
  float *o = pts_out + static_cast<size_t>(i) * g.C;
#pragma unroll
  for (int j = 0; j < MBEV_MAX_POINT_DIM; ++j)
    if (j < g.C) o[j] = v[j];
  int cell;
  if (dropped || !point_cell(v[0], v[1], v[2], g, cell)) {
    cellid[i] = -1;
    return;
  }
  const int gc = f * g.cells + cell;
  cellid[i] = gc;
  next[i] = atomicExch(head + gc, i);
}

__global__ void __launch_bounds__(kThreads)
k_rank(const int total, const int T, const int *__restrict__ head, const int *__restrict__ next,
       const int *__restrict__ cellid, int *__restrict__ rank, int *__restrict__ aux,
       unsigned *__restrict__ firstbits) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  bool is_first = false;
  if (i < total) {
    const int gc = cellid[i];
    int r = -1;
    if (gc >= 0) {
      int j = head[gc];
      int smaller = 0, first = i, n = 0;
      while (j >= 0) {
        ++n;
        if (j < i) {
          ++smaller;
          first = min(first, j);
          if (smaller >= T) break;
        }
        j = next[j];
      }
      if (smaller < T) {  // walked the whole list: `first` is the pillar head, n the cell population
        r = smaller;
        is_first = (r == 0);
        aux[i] = is_first ? n : first;
      }
    }
    rank[i] = r;
  }
  const unsigned b = __ballot_sync(0xffffffffu, is_first);
  if ((threadIdx.x & 31) == 0 && i < total) firstbits[i >> 5] = b;
}

// Single-CTA form (small inputs: one launch beats three). words = ceil(total/32). wprefix[w] = number of head flags in words < w.
__global__ void __launch_bounds__(1024)
k_scan(const unsigned *__restrict__ bits, const int words, const int total, const __grid_constant__ Frames fr,
       const int V, int *__restrict__ wprefix, int *__restrict__ fprefix, int *__restrict__ pillar_base) {
  __shared__ int s_warp[32];
  __shared__ int s_raw[MBEV_MAX_BATCH + 1];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // each warp owns a contiguous span of words and walks it 32 words at a time, so every load and every wprefix
  // store is one coalesced 128-byte access (a thread-contiguous split made this kernel 53 us of load latency)
  const int per = ((words + 31) / 32 + 31) / 32 * 32;  // words per warp, multiple of 32
  const int w0 = min(words, warp * per), w1 = min(words, w0 + per);
  int s = 0;
#pragma unroll 4
  for (int w = w0 + lane; w < w1; w += 32) s += __popc(__ldg(bits + w));
#pragma unroll
  for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0) s_warp[warp] = s;
  __syncthreads();
  if (tid < 32) {
    int x = s_warp[tid];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, x, d);
      if (tid >= d) x += u;
    }
    s_warp[tid] = x;  // inclusive over warps
    if (tid == 31) s_total = x;
  }
  __syncthreads();
  int run = warp ? s_warp[warp - 1] : 0;  // head flags before this warp's span
#pragma unroll 2
  for (int wb = w0; wb < w1; wb += 32) {
    const int w = wb + lane;
    const int c = (w < w1) ? __popc(__ldg(bits + w)) : 0;
    int v = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += u;
    }
    if (w < w1) wprefix[w] = run + v - c;
    run += __shfl_sync(0xffffffffu, v, 31);
  }
  __syncthreads();  // wprefix (global) written by this CTA is visible to it after the barrier
  for (int f = tid; f <= fr.batch; f += 1024) {
    const int i = fr.off[f];
    int p;
    if (i >= total) p = s_total;
    else p = wprefix[i >> 5] + __popc(bits[i >> 5] & ((1u << (i & 31)) - 1u));
    s_raw[f] = p;
    fprefix[f] = p;
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    pillar_base[0] = 0;
    for (int f = 0; f < fr.batch; ++f) {
      acc += min(s_raw[f + 1] - s_raw[f], V);
      pillar_base[f + 1] = acc;
    }
  }
}

// Head-flag scan, words = ceil(total/32), wprefix[w] = number of head flags in words < w. Three small launches of
// 256-thread CTAs instead of one 1024-thread CTA: (1) per-CTA popcount totals, (2) every CTA re-reduces the totals
// before it and scans its own 2048 words, (3) per-frame pillar counts and bases. Besides the shorter critical path
// (28 -> ~16 us on kitti_b16) the small CTAs fit next to other kernels' CTAs in the pipelined entry. (Measured: K1
// still overlaps K3 only, not K2's persistent CTAs — also with the K2 shared-memory carve-out requested for these
// kernels, which just costs them L1: 0.106 -> 0.129 ms.)
constexpr int kScanWords = 2048;  // 8 warps x 8 coalesced groups of 32 words

__device__ __forceinline__ int scan_warp_sum(const unsigned *__restrict__ bits, const int w0, const int words, const int lane) {
  int s = 0;
#pragma unroll
  for (int g = 0; g < kScanWords / 256; ++g) {
    const int w = w0 + 32 * g + lane;
    s += (w < words) ? __popc(__ldg(bits + w)) : 0;
  }
  return __reduce_add_sync(0xffffffffu, s);
}

__global__ void __launch_bounds__(kThreads)
k_scan_part(const unsigned *__restrict__ bits, const int words, int *__restrict__ ctot) {
  __shared__ int s_w[kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = scan_warp_sum(bits, blockIdx.x * kScanWords + warp * (kScanWords / 8), words, lane);
  if (lane == 0) s_w[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += s_w[w];
    ctot[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kThreads)
k_scan_final(const unsigned *__restrict__ bits, const int words, const int *__restrict__ ctot,
             int *__restrict__ wprefix) {
  __shared__ int s_w[kThreads / 32], s_b[kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base = 0;
  for (int j = threadIdx.x; j < static_cast<int>(blockIdx.x); j += kThreads) base += __ldg(ctot + j);
  base = __reduce_add_sync(0xffffffffu, base);
  const int w0 = blockIdx.x * kScanWords + warp * (kScanWords / 8);
  const int mine = scan_warp_sum(bits, w0, words, lane);
  if (lane == 0) {
    s_b[warp] = base;
    s_w[warp] = mine;
  }
  __syncthreads();
  int run = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    run += s_b[w];
    if (w < warp) run += s_w[w];
  }
#pragma unroll
  for (int g = 0; g < kScanWords / 256; ++g) {
    const int w = w0 + 32 * g + lane;
    const int c = (w < words) ? __popc(__ldg(bits + w)) : 0;
    int v = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += u;
    }
    if (w < words) wprefix[w] = run + v - c;
    run += __shfl_sync(0xffffffffu, v, 31);
  }
}

__global__ void __launch_bounds__(kThreads)
k_scan_frames(const unsigned *__restrict__ bits, const int *__restrict__ wprefix, const int *__restrict__ ctot,
              const int nparts, const int total, const __grid_constant__ Frames fr, const int V,
              int *__restrict__ fprefix, int *__restrict__ pillar_base) {
  __shared__ int s_raw[MBEV_MAX_BATCH + 1];
  __shared__ int s_w[kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int t = 0;
  for (int j = tid; j < nparts; j += kThreads) t += __ldg(ctot + j);
  t = __reduce_add_sync(0xffffffffu, t);
  if (lane == 0) s_w[warp] = t;
  __syncthreads();
  int s_total = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) s_total += s_w[w];
  for (int f = tid; f <= fr.batch; f += kThreads) {
    const int i = fr.off[f];
    int p;
    if (i >= total) p = s_total;
    else p = wprefix[i >> 5] + __popc(bits[i >> 5] & ((1u << (i & 31)) - 1u));
    s_raw[f] = p;
    fprefix[f] = p;
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    pillar_base[0] = 0;
    for (int f = 0; f < fr.batch; ++f) {
      acc += min(s_raw[f + 1] - s_raw[f], V);
      pillar_base[f + 1] = acc;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
k_emit(const int total, const __grid_constant__ Frames fr, const __grid_constant__ GeoK g,
       const int *__restrict__ cellid, const int *__restrict__ rank, const int *__restrict__ aux,
       const unsigned *__restrict__ bits, const int *__restrict__ wprefix, const int *__restrict__ fprefix,
       const int *__restrict__ pillar_base, int *__restrict__ table, int *__restrict__ coors,
       int *__restrict__ num_points, int *__restrict__ kept_idx) {
  __shared__ int s_off[MBEV_MAX_BATCH + 1];
  for (int t = threadIdx.x; t <= fr.batch; t += kThreads) s_off[t] = fr.off[t];
  __syncthreads();
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= total) return;
  const int r = rank[i];
  if (r < 0) return;
  const int a = aux[i];
  const int first = (r == 0) ? i : a;
  const int f = frame_of(s_off, fr.batch, i);
  const int local = wprefix[first >> 5] + __popc(bits[first >> 5] & ((1u << (first & 31)) - 1u)) - fprefix[f];
  const int gc = cellid[i];
  if (local >= g.V) {  // pillar never created (max_voxels); every point of it is dropped
    if (r == 0) table[gc] = -1;
    return;
  }
  const int pid = pillar_base[f] + local;
  kept_idx[static_cast<size_t>(pid) * g.T + r] = i;
  if (r == 0) {
    const int c = gc - f * g.cells;
    const int x = c % g.nx, yz = c / g.nx;
    reinterpret_cast<int4 *>(coors)[pid] = make_int4(f, yz / g.ny, yz % g.ny, x);
    num_points[pid] = min(a, g.T);
    table[gc] = pid;
  }
}

// (P, T, C) zero-padded voxel tensor as mmcv returns it.
__global__ void __launch_bounds__(kThreads)
k_gather_voxels(const float *__restrict__ pts, const int *__restrict__ kept_idx, const int *__restrict__ num_points,
                const int *__restrict__ num_pillars, const int T, const int C, float *__restrict__ voxels) {
  const long long slots = static_cast<long long>(*num_pillars) * T;
  for (long long s = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; s < slots;
       s += static_cast<long long>(gridDim.x) * kThreads) {
    const int p = static_cast<int>(s / T), t = static_cast<int>(s - static_cast<long long>(p) * T);
    float *dst = voxels + s * C;
    if (t < num_points[p]) {
      const float *src = pts + static_cast<size_t>(kept_idx[s]) * C;
      for (int k = 0; k < C; ++k) dst[k] = __ldg(src + k);
    } else {
      for (int k = 0; k < C; ++k) dst[k] = 0.f;
    }
  }
}

struct VoxWs {
  int *next, *cellid, *rank, *aux, *wprefix, *fprefix, *ctot;
  unsigned *bits;
  size_t bytes;
};

VoxWs carve(void *ws, int batch, int64_t total) {
  Carver c(ws);
  VoxWs w;
  const size_t n = static_cast<size_t>(total > 0 ? total : 1);
  const size_t words = (n + 31) / 32;
  w.next = c.take<int>(n);
  w.cellid = c.take<int>(n);
  w.rank = c.take<int>(n);
  w.aux = c.take<int>(n);
  w.bits = c.take<unsigned>(words);
  w.wprefix = c.take<int>(words);
  w.fprefix = c.take<int>(static_cast<size_t>(batch) + 1);
  w.ctot = c.take<int>((words + kScanWords - 1) / kScanWords + 1);
  w.bytes = c.off;
  return w;
}

int check_geo(const MbevGeometry *geo, int batch) {
  if (!geo) return MBEV_ERR_BAD_ARG;
  if (batch < 1 || batch > MBEV_MAX_BATCH) return MBEV_ERR_BAD_ARG;
  if (geo->num_feats < 3 || geo->num_feats > MBEV_MAX_POINT_DIM) return MBEV_ERR_UNSUPPORTED;
  if (geo->max_points < 1 || geo->max_voxels < 1) return MBEV_ERR_BAD_ARG;
  for (int j = 0; j < 3; ++j)
    if (geo->grid[j] < 1 || !(geo->voxel[j] > 0.f)) return MBEV_ERR_BAD_ARG;
  const int64_t cells = static_cast<int64_t>(geo->grid[0]) * geo->grid[1] * geo->grid[2];
  if (cells * batch > 0x7fffffffLL) return MBEV_ERR_UNSUPPORTED;
  return MBEV_OK;
}

}  // namespace
}  // namespace mbev

using namespace mbev;

extern "C" int64_t mbev_pillar_capacity(const int64_t *frame_offsets_host, int batch, const MbevGeometry *geo) {
  if (!frame_offsets_host || !geo || batch < 1) return -1;
  const int64_t cells = static_cast<int64_t>(geo->grid[0]) * geo->grid[1] * geo->grid[2];
  const int64_t per_frame = std::min<int64_t>(geo->max_voxels, cells);
  int64_t cap = 0;
  for (int f = 0; f < batch; ++f)
    cap += std::min<int64_t>(std::max<int64_t>(frame_offsets_host[f + 1] - frame_offsets_host[f], 0), per_frame);
  return cap;
}

extern "C" int mbev_voxelize_workspace_bytes(const MbevGeometry *geo, int batch, int64_t total_points,
                                             size_t *bytes) {
  if (!bytes || total_points < 0) return MBEV_ERR_BAD_ARG;
  const int st = check_geo(geo, batch);
  if (st) return st;
  if (total_points > 0x7fffffffLL) return MBEV_ERR_UNSUPPORTED;
  *bytes = carve(nullptr, batch, total_points).bytes;
  return MBEV_OK;
}

static int voxelize_impl(const float *points, const int64_t *frame_offsets_host, int batch, const MbevGeometry *geo,
                         int32_t *cell_table, int32_t *coors, int32_t *num_points, int32_t *kept_idx,
                         int32_t *pillar_base, int64_t pillar_capacity, void *workspace, size_t workspace_bytes,
                         void *stream_, const MbevAugment *aug) {
  const int st = check_geo(geo, batch);
  if (st) return st;
  if (!frame_offsets_host || !cell_table || !coors || !num_points || !kept_idx || !pillar_base || !workspace)
    return MBEV_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Frames fr;
  fr.batch = batch;
  if (frame_offsets_host[0] != 0) return MBEV_ERR_BAD_ARG;
  for (int f = 0; f <= batch; ++f) {
    if (frame_offsets_host[f] > 0x7fffffffLL || (f && frame_offsets_host[f] < frame_offsets_host[f - 1]))
      return MBEV_ERR_BAD_ARG;
    fr.off[f] = static_cast<int32_t>(frame_offsets_host[f]);
  }
  const int total = fr.off[batch];
  if (total > 0 && !points) return MBEV_ERR_BAD_ARG;
  const GeoK g = make_geok(*geo);
  // a frame cannot hold more pillars than it has points, than V, or than cells: capacity = sum over frames
  // (the bound the Python host side sizes its buffers with, functional.pillar_capacity)
  if (pillar_capacity < mbev_pillar_capacity(frame_offsets_host, batch, geo)) return MBEV_ERR_BAD_ARG;
  const VoxWs w = carve(workspace, batch, total);
  if (workspace_bytes < w.bytes) return MBEV_ERR_WORKSPACE;

  MBEV_CUDA(cudaMemsetAsync(cell_table, 0xff, sizeof(int32_t) * static_cast<size_t>(batch) * g.cells, stream));
  const int words = (total + 31) / 32;
  if (total > 0) {
    const int blocks = (total + kThreads - 1) / kThreads;
    if (aug)
      k_assign_aug<<<blocks, kThreads, 0, stream>>>(points, total, fr, g, aug->frames, aug->drop_u, aug->noise, aug->seed,
                                                   aug->points_out, cell_table, w.next, w.cellid);
    else if (g.C == 4 && (reinterpret_cast<uintptr_t>(points) & 15) == 0)
      k_assign<4><<<blocks, kThreads, 0, stream>>>(points, total, fr, g, cell_table, w.next, w.cellid);
    else
      k_assign<0><<<blocks, kThreads, 0, stream>>>(points, total, fr, g, cell_table, w.next, w.cellid);
    MBEV_CHECK_LAUNCH();
    k_rank<<<blocks, kThreads, 0, stream>>>(total, g.T, cell_table, w.next, w.cellid, w.rank, w.aux, w.bits);
    MBEV_CHECK_LAUNCH();
  }
  const int nparts = (words + kScanWords - 1) / kScanWords;
  if (nparts <= 8) {  // up to ~500 k points: the single-CTA scan is one launch and short
    k_scan<<<1, 1024, 0, stream>>>(w.bits, words, total, fr, g.V, w.wprefix, w.fprefix, pillar_base);
    MBEV_CHECK_LAUNCH();
  } else {
    k_scan_part<<<nparts, kThreads, 0, stream>>>(w.bits, words, w.ctot);
    MBEV_CHECK_LAUNCH();
    k_scan_final<<<nparts, kThreads, 0, stream>>>(w.bits, words, w.ctot, w.wprefix);
    MBEV_CHECK_LAUNCH();
    k_scan_frames<<<1, kThreads, 0, stream>>>(w.bits, w.wprefix, w.ctot, nparts, total, fr, g.V, w.fprefix, pillar_base);
    MBEV_CHECK_LAUNCH();
  }
  if (total > 0) {
    const int blocks = (total + kThreads - 1) / kThreads;
    k_emit<<<blocks, kThreads, 0, stream>>>(total, fr, g, w.cellid, w.rank, w.aux, w.bits, w.wprefix, w.fprefix,
                                            pillar_base, cell_table, coors, num_points, kept_idx);
    MBEV_CHECK_LAUNCH();
  }
  return MBEV_OK;
}

extern "C" int mbev_voxelize(const float *points, const int64_t *frame_offsets_host, int batch,
                             const MbevGeometry *geo, int32_t *cell_table, int32_t *coors, int32_t *num_points,
                             int32_t *kept_idx, int32_t *pillar_base, int64_t pillar_capacity, void *workspace,
                             size_t workspace_bytes, void *stream_) {
  return voxelize_impl(points, frame_offsets_host, batch, geo, cell_table, coors, num_points, kept_idx, pillar_base,
                       pillar_capacity, workspace, workspace_bytes, stream_, nullptr);
}

extern "C" int mbev_voxelize_augmented(const float *points, const int64_t *frame_offsets_host, int batch,
                                       const MbevGeometry *geo, const MbevAugment *aug, int32_t *cell_table,
                                       int32_t *coors, int32_t *num_points, int32_t *kept_idx, int32_t *pillar_base,
                                       int64_t pillar_capacity, void *workspace, size_t workspace_bytes, void *stream_) {
  if (!aug || !aug->frames || !aug->points_out) return MBEV_ERR_BAD_ARG;
  return voxelize_impl(points, frame_offsets_host, batch, geo, cell_table, coors, num_points, kept_idx, pillar_base,
                       pillar_capacity, workspace, workspace_bytes, stream_, aug);
}

extern "C" int mbev_gather_voxels(const float *points, const int32_t *kept_idx, const int32_t *num_points,
                                  const int32_t *num_pillars_dev, int64_t pillar_capacity, int T, int C,
                                  float *voxels, void *stream_) {
  if (!kept_idx || !num_points || !num_pillars_dev || !voxels || T < 1 || C < 1) return MBEV_ERR_BAD_ARG;
  if (pillar_capacity <= 0) return MBEV_OK;
  if (!points) return MBEV_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long slots = static_cast<long long>(pillar_capacity) * T;
  const int blocks = static_cast<int>(std::min<long long>((slots + kThreads - 1) / kThreads, kNumSMs * 16));
  k_gather_voxels<<<blocks, kThreads, 0, stream>>>(points, kept_idx, num_points, num_pillars_dev, T, C, voxels);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}
