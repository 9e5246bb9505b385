// K3 / K3' — pillar features <-> dense BEV canvas (sm_100a).
//
// Replaces mmdet3d PointPillarsScatter.forward_batch (mask_bev_encoders.py:122-123), which per frame does
// memset + boolean select (host sync) + transposed index_put + stack: the canvas is written ~3x and read
// once. Here the canvas (B, C, ny*nx) fp32 NCHW is written exactly once by a persistent streaming kernel
// that walks the cell table (cell -> pillar id, -1 empty): every 16-byte store carries either zeros or
// features, so DRAM traffic = canvas bytes + table + the occupied feature rows (HBM-bound, write-only).
// The backward is the gather dfeats[p,:] = dcanvas[b,:,y,x] driven by the same table.
#include <cstdlib>

#include "common.cuh"

namespace mbev {
namespace {

constexpr int kCells = 128;    // cells per tile: one warp-wide float4 store covers a whole tile row (512 B)
constexpr int kThreads = 256;  // 8 warps; warp w owns channels w, w+8, ...

// smem: feature rows of the occupied cells of this tile, [kCells][C+1] (odd pitch: conflict-free column reads)
__global__ void __launch_bounds__(kThreads)
k_scatter(const float *__restrict__ feats, const int *__restrict__ table, const int C, const int G,
          const int tiles_per_frame, const int num_tiles, float *__restrict__ canvas) {
  extern __shared__ float s_rows[];
  __shared__ int s_pid[kCells];
  __shared__ int s_any;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pitch = C + 1;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_frame;
    const int g0 = (tile - b * tiles_per_frame) * kCells;
    __syncthreads();  // previous iteration's readers are done with s_pid / s_rows
    if (tid == 0) s_any = 0;
    __syncthreads();
    if (tid < kCells) {
      const int g = g0 + tid;
      const int pid = (g < G) ? __ldg(table + static_cast<size_t>(b) * G + g) : -1;
      s_pid[tid] = pid;
      if (pid >= 0) s_any = 1;
    }
    __syncthreads();
    const bool any = s_any != 0;
    if (any) {
      // stage occupied rows: one warp per cell, coalesced 128-bit reads of the (C) row
      for (int c = warp; c < kCells; c += kThreads / 32) {
        const int pid = s_pid[c];
        if (pid < 0) continue;
        const float *src = feats + static_cast<size_t>(pid) * C;
        float *dst = s_rows + c * pitch;
        for (int k = lane; k < C; k += 32) dst[k] = __ldg(src + k);
      }
      __syncthreads();
    }
    // lane owns cells 4*lane .. 4*lane+3 of the tile
    const int c0 = lane * 4;
    const int p0 = s_pid[c0], p1 = s_pid[c0 + 1], p2 = s_pid[c0 + 2], p3 = s_pid[c0 + 3];
    const bool mine = (p0 >= 0) | (p1 >= 0) | (p2 >= 0) | (p3 >= 0);
    const int g = g0 + c0;
    float *out = canvas + (static_cast<size_t>(b) * C) * G + g;
    if (g + 3 < G) {
      for (int ch = warp; ch < C; ch += kThreads / 32) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mine) {
          if (p0 >= 0) v.x = s_rows[(c0 + 0) * pitch + ch];
          if (p1 >= 0) v.y = s_rows[(c0 + 1) * pitch + ch];
          if (p2 >= 0) v.z = s_rows[(c0 + 2) * pitch + ch];
          if (p3 >= 0) v.w = s_rows[(c0 + 3) * pitch + ch];
        }
        st_global_v4_stream(out + static_cast<size_t>(ch) * G, v);
      }
    } else if (g < G) {  // ragged tail of the frame
      for (int ch = warp; ch < C; ch += kThreads / 32) {
        for (int k = 0; k < 4 && g + k < G; ++k) {
          const int p = s_pid[c0 + k];
          out[static_cast<size_t>(ch) * G + k] = (p >= 0) ? s_rows[(c0 + k) * pitch + ch] : 0.f;
        }
      }
    }
  }
}

// Barrier-free form of the one-pass scatter: a WARP owns a run of 512 cells (4 x 128) and writes all C planes of
// it. No shared memory and no block barrier: the 16 pillar ids of a lane stay in registers, feature values of
// occupied cells come straight from L1/L2 (a pillar's C floats are one 512-byte row, re-read channel by channel
// by the same lane), and per plane the warp emits 4 consecutive 512-byte stores = 2 KB contiguous. ncu on
// k_scatter showed no store throttling at 78 % of the HBM peak — it waits on its own three barriers per tile and
// on the table load — so this version removes them.
constexpr int kS2Cells = 512;

// `csplit` > 1 (small batches): a task is (run, channel chunk) so that one frame still fills the machine.
// Persistent form (grid-stride over tasks): kept selectable (MBEV_SCATTER=1) next to k_scatter_run, which replaced it.
__global__ void __launch_bounds__(kThreads)
k_scatter_warp(const float *__restrict__ feats, const int *__restrict__ table, const int C, const int G,
               const int tiles_per_frame, const int num_tiles, const int csplit, float *__restrict__ canvas) {
  const int lane = threadIdx.x & 31;
  const int nw = gridDim.x * (kThreads / 32);
  const int cper = (C + csplit - 1) / csplit;
  for (int task = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); task < num_tiles * csplit; task += nw) {
    const int tile = num_tiles - 1 - task / csplit;
    const int ch0 = (task % csplit) * cper, ch1 = min(C, ch0 + cper);
    const int b = tile / tiles_per_frame;
    const int g0 = (tile - b * tiles_per_frame) * kS2Cells + 4 * lane;
    int4 pid[4];
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g = g0 + 128 * k;
      pid[k] = (g < G) ? __ldg(reinterpret_cast<const int4 *>(table + static_cast<size_t>(b) * G + g))
                       : make_int4(-1, -1, -1, -1);
      any |= (pid[k].x & pid[k].y & pid[k].z & pid[k].w) >= 0;
    }
    float *out = canvas + (static_cast<size_t>(b) * C) * G + g0;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!__any_sync(0xffffffffu, any)) {  // a run without pillars: pure zero stream
      for (int ch = ch0; ch < ch1; ++ch) {
        float *o = out + static_cast<size_t>(ch) * G;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (g0 + 128 * k < G) st_global_v4_stream(o + 128 * k, z);
      }
      continue;
    }
    // compose every 16-byte store, one plane at a time; the feature values of plane ch+1 are requested before the
    // stores of plane ch are issued, so the store stream does not stall behind its own loads
    auto load_plane = [&](int ch, float4 (&v)[4]) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = z;
        if (any) {
          if (pid[k].x >= 0) v[k].x = __ldg(feats + static_cast<size_t>(pid[k].x) * C + ch);
          if (pid[k].y >= 0) v[k].y = __ldg(feats + static_cast<size_t>(pid[k].y) * C + ch);
          if (pid[k].z >= 0) v[k].z = __ldg(feats + static_cast<size_t>(pid[k].z) * C + ch);
          if (pid[k].w >= 0) v[k].w = __ldg(feats + static_cast<size_t>(pid[k].w) * C + ch);
        }
      }
    };
    float4 nxt[4];
    load_plane(ch0, nxt);
    for (int ch = ch0; ch < ch1; ++ch) {
      float4 cur[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
      if (ch + 1 < ch1) load_plane(ch + 1, nxt);
      float *o = out + static_cast<size_t>(ch) * G;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (g0 + 128 * k < G) st_global_v4_stream_nc(o + 128 * k, cur[k]);
    }
  }
}

// Lean form of k_scatter_warp with the run length as a template parameter (KK x 128 cells): only the zero-stream
// fast path and the composing loop with load-ahead. MBEV_SCATTER=5 runs KK = 2 (half the registers per warp, more
// resident warps, 1 KB instead of 2 KB contiguous per plane and warp), MBEV_SCATTER=6 runs KK = 4.
template <int KK>
__global__ void __launch_bounds__(kThreads)
k_scatter_run(const float *__restrict__ feats, const int *__restrict__ table, const int C, const int G,
              const int tiles_per_frame, const int num_tiles, const int csplit, float *__restrict__ canvas) {
  const int lane = threadIdx.x & 31;
  const int nw = gridDim.x * (kThreads / 32);
  const int cper = (C + csplit - 1) / csplit;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int task = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); task < num_tiles * csplit; task += nw) {
    const int tile = num_tiles - 1 - task / csplit;
    const int ch0 = (task % csplit) * cper, ch1 = min(C, ch0 + cper);
    const int b = tile / tiles_per_frame;
    const int g0 = (tile - b * tiles_per_frame) * (128 * KK) + 4 * lane;
    int4 pid[KK];
    bool any = false;
#pragma unroll
    for (int k = 0; k < KK; ++k) {
      const int g = g0 + 128 * k;
      pid[k] = (g < G) ? __ldg(reinterpret_cast<const int4 *>(table + static_cast<size_t>(b) * G + g))
                       : make_int4(-1, -1, -1, -1);
      any |= (pid[k].x & pid[k].y & pid[k].z & pid[k].w) >= 0;
    }
    float *out = canvas + (static_cast<size_t>(b) * C) * G + g0;
    if (!__any_sync(0xffffffffu, any)) {
      for (int ch = ch0; ch < ch1; ++ch) {
        float *o = out + static_cast<size_t>(ch) * G;
#pragma unroll
        for (int k = 0; k < KK; ++k)
          if (g0 + 128 * k < G) st_global_v4_stream(o + 128 * k, z);
      }
      continue;
    }
    auto load_plane = [&](int ch, float4 (&v)[KK]) {
#pragma unroll
      for (int k = 0; k < KK; ++k) {
        v[k] = z;
        if (any) {
          if (pid[k].x >= 0) v[k].x = __ldg(feats + static_cast<size_t>(pid[k].x) * C + ch);
          if (pid[k].y >= 0) v[k].y = __ldg(feats + static_cast<size_t>(pid[k].y) * C + ch);
          if (pid[k].z >= 0) v[k].z = __ldg(feats + static_cast<size_t>(pid[k].z) * C + ch);
          if (pid[k].w >= 0) v[k].w = __ldg(feats + static_cast<size_t>(pid[k].w) * C + ch);
        }
      }
    };
    float4 nxt[KK];
    load_plane(ch0, nxt);
    for (int ch = ch0; ch < ch1; ++ch) {
      float4 cur[KK];
#pragma unroll
      for (int k = 0; k < KK; ++k) cur[k] = nxt[k];
      if (ch + 1 < ch1) load_plane(ch + 1, nxt);
      float *o = out + static_cast<size_t>(ch) * G;
#pragma unroll
      for (int k = 0; k < KK; ++k)
        if (g0 + 128 * k < G) st_global_v4_stream_nc(o + 128 * k, cur[k]);
    }
  }
}

// Warp-local staging form of the one-pass scatter (MBEV_SCATTER=4). k_scatter_warp composes every 16-byte store
// from four predicated loads and spends ~40 instructions per plane on a run that holds ~17 pillars (60 % issue
// utilisation; zeros alone stream at 7.0 TB/s with the same access pattern). Here the warp keeps one plane of its
// run (512 floats) in shared memory, all zeros: per plane, lane i drops the value of pillar i at its cell (one load
// + one shared store per PILLAR, not per cell), the warp reads its 4 x 16 bytes back and streams them out, and lane
// i zeroes its cell again. The (cell, pillar) list of the run is built once with ballots. No block barrier.
__global__ void __launch_bounds__(kThreads)
k_scatter_stage(const float *__restrict__ feats, const int *__restrict__ table, const int C, const int G,
                const int tiles_per_frame, const int num_tiles, const int csplit, float *__restrict__ canvas) {
  __shared__ __align__(16) float s_plane[kThreads / 32][kS2Cells];
  __shared__ int s_pid[kThreads / 32][kS2Cells];
  __shared__ unsigned short s_off[kThreads / 32][kS2Cells];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = gridDim.x * (kThreads / 32);
  const int cper = (C + csplit - 1) / csplit;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float *plane = s_plane[warp];
  for (int i = lane; i < kS2Cells; i += 32) plane[i] = 0.f;
  __syncwarp();
  for (int task = blockIdx.x * (kThreads / 32) + warp; task < num_tiles * csplit; task += nw) {
    const int tile = num_tiles - 1 - task / csplit;  // last frame first (its feature rows are the freshest in L2)
    const int ch0 = (task % csplit) * cper, ch1 = min(C, ch0 + cper);
    const int b = tile / tiles_per_frame;
    const int g0 = (tile - b * tiles_per_frame) * kS2Cells + 4 * lane;
    int np = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g = g0 + 128 * k;
      const int4 p4 = (g < G) ? __ldg(reinterpret_cast<const int4 *>(table + static_cast<size_t>(b) * G + g))
                              : make_int4(-1, -1, -1, -1);
      const int pk[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned m = __ballot_sync(0xffffffffu, pk[j] >= 0);
        if (pk[j] >= 0) {
          const int pos = np + __popc(m & ((1u << lane) - 1u));
          s_pid[warp][pos] = pk[j];
          s_off[warp][pos] = static_cast<unsigned short>(128 * k + 4 * lane + j);
        }
        np += __popc(m);
      }
    }
    __syncwarp();
    float *out = canvas + (static_cast<size_t>(b) * C) * G + g0;
    if (np == 0) {  // a run without pillars: pure zero stream
      for (int ch = ch0; ch < ch1; ++ch) {
        float *o = out + static_cast<size_t>(ch) * G;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (g0 + 128 * k < G) st_global_v4_stream(o + 128 * k, z);
      }
      continue;
    }
    // the first 32 pillars of the run live in registers (the usual case: ~17 per run), the rest in the list
    const bool mine = lane < np;
    const float *row0 = feats + static_cast<size_t>(mine ? s_pid[warp][lane] : 0) * C;
    const int off0 = mine ? s_off[warp][lane] : 0;
    float nxt = mine ? __ldg(row0 + ch0) : 0.f;
    for (int ch = ch0; ch < ch1; ++ch) {
      if (mine) plane[off0] = nxt;
      for (int i = 32 + lane; i < np; i += 32) plane[s_off[warp][i]] = __ldg(feats + static_cast<size_t>(s_pid[warp][i]) * C + ch);
      if (mine && ch + 1 < ch1) nxt = __ldg(row0 + ch + 1);
      __syncwarp();
      float *o = out + static_cast<size_t>(ch) * G;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 v = *reinterpret_cast<const float4 *>(plane + 128 * k + 4 * lane);
        if (g0 + 128 * k < G) st_global_v4_stream_nc(o + 128 * k, v);
      }
      __syncwarp();
      if (mine) plane[off0] = 0.f;
      for (int i = 32 + lane; i < np; i += 32) plane[s_off[warp][i]] = 0.f;
    }
    __syncwarp();
  }
}

// bf16 canvas (BASELINE config 4 / north star "1e-2 in bf16"): the same one-pass walk, every value rounded to
// nearest-even bf16 on the way out, 8 bytes per lane and store — the canvas bytes, i.e. K3's roofline, halve.
__device__ __forceinline__ void st_global_v2_stream_nc(void *p, uint32_t a, uint32_t b) {
  asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__global__ void __launch_bounds__(kThreads)
k_scatter_warp_bf16(const float *__restrict__ feats, const int *__restrict__ table, const int C, const int G,
                    const int tiles_per_frame, const int num_tiles, const int csplit, uint16_t *__restrict__ canvas) {
  const int lane = threadIdx.x & 31;
  const int nw = gridDim.x * (kThreads / 32);
  const int cper = (C + csplit - 1) / csplit;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int task = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); task < num_tiles * csplit; task += nw) {
    const int tile = task / csplit;
    const int ch0 = (task - tile * csplit) * cper, ch1 = min(C, ch0 + cper);
    const int b = tile / tiles_per_frame;
    const int g0 = (tile - b * tiles_per_frame) * kS2Cells + 4 * lane;
    int4 pid[4];
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g = g0 + 128 * k;
      pid[k] = (g < G) ? __ldg(reinterpret_cast<const int4 *>(table + static_cast<size_t>(b) * G + g))
                       : make_int4(-1, -1, -1, -1);
      any |= (pid[k].x & pid[k].y & pid[k].z & pid[k].w) >= 0;
    }
    uint16_t *out = canvas + (static_cast<size_t>(b) * C) * G + g0;
    if (!__any_sync(0xffffffffu, any)) {
      for (int ch = ch0; ch < ch1; ++ch) {
        uint16_t *o = out + static_cast<size_t>(ch) * G;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (g0 + 128 * k < G) st_global_v2_stream_nc(o + 128 * k, 0u, 0u);
      }
      continue;
    }
    auto load_plane = [&](int ch, float4 (&v)[4]) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = z;
        if (any) {
          if (pid[k].x >= 0) v[k].x = __ldg(feats + static_cast<size_t>(pid[k].x) * C + ch);
          if (pid[k].y >= 0) v[k].y = __ldg(feats + static_cast<size_t>(pid[k].y) * C + ch);
          if (pid[k].z >= 0) v[k].z = __ldg(feats + static_cast<size_t>(pid[k].z) * C + ch);
          if (pid[k].w >= 0) v[k].w = __ldg(feats + static_cast<size_t>(pid[k].w) * C + ch);
        }
      }
    };
    float4 nxt[4];
    load_plane(ch0, nxt);
    for (int ch = ch0; ch < ch1; ++ch) {
      float4 cur[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
      if (ch + 1 < ch1) load_plane(ch + 1, nxt);
      uint16_t *o = out + static_cast<size_t>(ch) * G;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (g0 + 128 * k < G)
          st_global_v2_stream_nc(o + 128 * k, pack_bf16x2(cur[k].x, cur[k].y), pack_bf16x2(cur[k].z, cur[k].w));
    }
  }
}

// One-pass scatter with the zeros and the features on separate instruction streams (the default when planes are
// 32-byte aligned). k_scatter_warp composes every 16-byte store from four predicated loads — ~10 instructions per
// store, 45 % issue utilisation, and the store stream stalls behind the feature loads; a plain memset of the same
// canvas runs at 7.4 TB/s, k_scatter_warp at 5.6. Here a warp owns a run of 512 cells as before, but
//   zero pass   : per plane 4 x 512-byte stores, lanes whose 32-byte sector holds a pillar are predicated off
//                 (4 loop-invariant predicates) — ~2 instructions per store, no loads in the loop;
//   sector pass : per occupied sector (8 cells, lanes 2j / 2j+1 of one k) the warp turns into "lane = channel quad":
//                 the <= 8 pillar rows are read coalesced (512 B each), transposed in registers and written as whole
//                 sectors, 32 planes per store instruction.
// Every byte is still written exactly once and every sector whole (its two halves by adjacent instructions).
__global__ void __launch_bounds__(kThreads, 4)
k_scatter_holes(const float *__restrict__ feats, const int *__restrict__ table, const int C, const int G,
                const int tiles_per_frame, const int num_tiles, const int csplit, float *__restrict__ canvas) {
  const int lane = threadIdx.x & 31;
  const int nw = gridDim.x * (kThreads / 32);
  const int cper = C / csplit;  // multiple of 4
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int task = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); task < num_tiles * csplit; task += nw) {
    const int tile = task / csplit;
    const int ch0 = (task - tile * csplit) * cper, ch1 = ch0 + cper;
    const int b = tile / tiles_per_frame;
    const int g0 = (tile - b * tiles_per_frame) * kS2Cells + 4 * lane;
    int4 pid[4];
    bool hole[4];  // my sector (this lane's 16 bytes + its pair lane's) holds a pillar: the sector pass writes it
    unsigned smask[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g = g0 + 128 * k;
      pid[k] = (g < G) ? __ldg(reinterpret_cast<const int4 *>(table + static_cast<size_t>(b) * G + g))
                       : make_int4(-1, -1, -1, -1);
      const bool occ = (pid[k].x & pid[k].y & pid[k].z & pid[k].w) >= 0;
      const unsigned m = __ballot_sync(0xffffffffu, occ);
      const unsigned sect = (m | (m >> 1)) & 0x55555555u;  // bit 2j: sector j (lanes 2j, 2j+1) is occupied
      smask[k] = sect;
      hole[k] = ((sect >> (lane & ~1)) & 1u) != 0 || g >= G;
    }
    float *out = canvas + (static_cast<size_t>(b) * C) * G + g0;
    for (int ch = ch0; ch < ch1; ++ch) {
      float *o = out + static_cast<size_t>(ch) * G;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (!hole[k]) st_global_v4_stream(o + 128 * k, z);
    }
    // sector pass: lane = channel quad 4*lane .. 4*lane+3 (when inside this task's channel chunk)
    const int cq = 4 * lane;
    const bool cact = cq >= ch0 && cq < ch1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      for (unsigned m = smask[k]; m; m &= m - 1) {
        const int src = __ffs(m) - 1;  // even lane; the sector is cells 4*src .. 4*src+7 of this k
        float *so = canvas + (static_cast<size_t>(b) * C + cq) * G + (g0 - 4 * lane) + 128 * k + 4 * src;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int p0 = __shfl_sync(0xffffffffu, pid[k].x, src + half), p1 = __shfl_sync(0xffffffffu, pid[k].y, src + half);
          const int p2 = __shfl_sync(0xffffffffu, pid[k].z, src + half), p3 = __shfl_sync(0xffffffffu, pid[k].w, src + half);
          if (!cact) continue;
          float4 v0 = z, v1 = z, v2 = z, v3 = z;
          if (p0 >= 0) v0 = __ldg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(p0) * C + cq));
          if (p1 >= 0) v1 = __ldg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(p1) * C + cq));
          if (p2 >= 0) v2 = __ldg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(p2) * C + cq));
          if (p3 >= 0) v3 = __ldg(reinterpret_cast<const float4 *>(feats + static_cast<size_t>(p3) * C + cq));
          float *ho = so + 4 * half;
          st_global_v4_stream(ho, make_float4(v0.x, v1.x, v2.x, v3.x));
          st_global_v4_stream(ho + static_cast<size_t>(G), make_float4(v0.y, v1.y, v2.y, v3.y));
          st_global_v4_stream(ho + 2 * static_cast<size_t>(G), make_float4(v0.z, v1.z, v2.z, v3.z));
          st_global_v4_stream(ho + 3 * static_cast<size_t>(G), make_float4(v0.w, v1.w, v2.w, v3.w));
        }
      }
    }
  }
}

// ---- two-kernel form used by the fused path: K3a needs only the cell table, so it streams the zeros of the
// canvas while K2 (compute-bound, no DRAM traffic) runs on the main stream; K3b then writes the sectors that hold
// at least one pillar. The unit is the 32-byte DRAM sector = 8 consecutive cells of one channel plane: every
// sector is written exactly once, by exactly one of the two kernels, as a whole (no partial-sector read-modify-
// write), so DRAM traffic stays canvas bytes + table + feature rows. Needs G % 8 == 0 and a 32-byte aligned canvas.
constexpr int kFillCells = 128;  // cells per tile: lane owns 4 cells (16 B), a lane pair owns one sector per plane

__global__ void __launch_bounds__(kThreads)
k_fill_empty(const int *__restrict__ table, const int C, const int G, const int tiles_per_frame,
             const int num_tiles, float *__restrict__ canvas) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_frame;
    const int g = (tile - b * tiles_per_frame) * kFillCells + lane * 4;
    int occ = 0;  // bit 0: some cell of my half-sector holds a pillar
    if (g < G) {
      const int4 a = __ldg(reinterpret_cast<const int4 *>(table + static_cast<size_t>(b) * G + g));
      occ = ((a.x & a.y & a.z & a.w) >= 0) ? 1 : 0;
    }
    occ |= __shfl_xor_sync(0xffffffffu, occ, 1);  // the other half of the 32-byte sector
    if (g >= G || occ) continue;                   // K3b owns sectors with a pillar
    float *out = canvas + (static_cast<size_t>(b) * C) * G + g;
    for (int ch = warp; ch < C; ch += nwarp) st_global_v4_stream(out + static_cast<size_t>(ch) * G, z);
  }
}

// one warp per pillar; the pillar in the lowest occupied cell of a sector writes the sector for all channels:
// lane pair (2j, 2j+1) = the two 16-byte halves of channel (16 i + j)'s sector, so every store instruction
// covers 16 whole sectors
__global__ void __launch_bounds__(kThreads)
k_scatter_sectors(const float *__restrict__ feats, const int *__restrict__ coors, const int *__restrict__ num_pillars,
                  const int *__restrict__ table, const int batch, const int C, const int ny, const int nx,
                  float *__restrict__ canvas) {
  const int lane = threadIdx.x & 31;
  const int P = *num_pillars;
  const int G = ny * nx;
  const int nwarps = gridDim.x * (kThreads / 32);
  const int half = lane & 1, j = lane >> 1;
  for (int p = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); p < P; p += nwarps) {
    const int4 c = __ldg(reinterpret_cast<const int4 *>(coors) + p);  // (b, z, y, x)
    if (c.x < 0 || c.x >= batch || c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) continue;
    const int g = c.z * nx + c.w;
    const int gs = g & ~7;
    const int mine = lane < 8 ? __ldg(table + static_cast<size_t>(c.x) * G + gs + lane) : -1;
    const unsigned occ = __ballot_sync(0xffffffffu, mine >= 0) & 0xffu;
    if ((__ffs(occ) - 1) != (g & 7)) continue;  // warp-uniform: another pillar of this sector writes it
    int pid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) pid[k] = __shfl_sync(0xffffffffu, mine, 4 * half + k);
    float *out = canvas + (static_cast<size_t>(c.x) * C) * G + gs + 4 * half;
    for (int ch = j; ch < C; ch += 16) {
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = pid[k] >= 0 ? __ldg(feats + static_cast<size_t>(pid[k]) * C + ch) : 0.f;
      st_global_v4_stream(out + static_cast<size_t>(ch) * G, make_float4(v[0], v[1], v[2], v[3]));
    }
  }
}

// Generic fallback when G is not a multiple of 4 (plane rows are then not 16-byte aligned).
__global__ void __launch_bounds__(kThreads)
k_scatter_scalar(const float *__restrict__ feats, const int *__restrict__ table, const int C, const int G,
                 const long long total, float *__restrict__ canvas) {
  for (long long e = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * kThreads) {
    const int g = static_cast<int>(e % G);
    const long long bc = e / G;
    const int ch = static_cast<int>(bc % C);
    const int b = static_cast<int>(bc / C);
    const int pid = __ldg(table + static_cast<size_t>(b) * G + g);
    canvas[e] = pid >= 0 ? __ldg(feats + static_cast<size_t>(pid) * C + ch) : 0.f;
  }
}

// K3' : one warp per occupied cell, lanes over channels.
__global__ void __launch_bounds__(kThreads)
k_gather_bwd(const float *__restrict__ dcanvas, const int *__restrict__ table, const int C, const int G,
             const int tiles_per_frame, const int num_tiles, float *__restrict__ dfeats) {
  __shared__ int s_pid[kCells];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_frame;
    const int g0 = (tile - b * tiles_per_frame) * kCells;
    __syncthreads();
    if (tid < kCells) {
      const int g = g0 + tid;
      s_pid[tid] = (g < G) ? __ldg(table + static_cast<size_t>(b) * G + g) : -1;
    }
    __syncthreads();
    for (int c = warp; c < kCells; c += kThreads / 32) {
      const int pid = s_pid[c];
      if (pid < 0) continue;
      const float *src = dcanvas + (static_cast<size_t>(b) * C) * G + (g0 + c);
      float *dst = dfeats + static_cast<size_t>(pid) * C;
      for (int ch = lane; ch < C; ch += 32) dst[ch] = __ldg(src + static_cast<size_t>(ch) * G);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
k_build_table(const int *__restrict__ coors, const int *__restrict__ num_pillars, const int batch, const int ny,
              const int nx, int *__restrict__ table) {
  const int P = *num_pillars;
  for (int p = blockIdx.x * kThreads + threadIdx.x; p < P; p += gridDim.x * kThreads) {
    const int4 c = reinterpret_cast<const int4 *>(coors)[p];  // (b, z, y, x)
    if (c.x < 0 || c.x >= batch || c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) continue;
    table[(static_cast<size_t>(c.x) * ny + c.z) * nx + c.w] = p;
  }
}

}  // namespace
}  // namespace mbev

using namespace mbev;

extern "C" int mbev_build_cell_table(const int32_t *coors, const int32_t *num_pillars_dev, int64_t pillar_capacity,
                                     int batch, int ny, int nx, int32_t *cell_table, void *stream_) {
  if (!cell_table || !num_pillars_dev || batch < 1 || ny < 1 || nx < 1) return MBEV_ERR_BAD_ARG;
  if (static_cast<int64_t>(batch) * ny * nx > 0x7fffffffLL) return MBEV_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MBEV_CUDA(cudaMemsetAsync(cell_table, 0xff, sizeof(int32_t) * static_cast<size_t>(batch) * ny * nx, stream));
  if (pillar_capacity <= 0) return MBEV_OK;
  if (!coors) return MBEV_ERR_BAD_ARG;
  const int blocks = static_cast<int>(std::min<int64_t>((pillar_capacity + kThreads - 1) / kThreads, kNumSMs * 8));
  k_build_table<<<blocks, kThreads, 0, stream>>>(coors, num_pillars_dev, batch, ny, nx, cell_table);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

extern "C" int mbev_scatter_forward(const float *feats, const int32_t *cell_table, int batch, int c_out, int ny,
                                    int nx, float *canvas, void *stream_) {
  if (!cell_table || !canvas || batch < 1 || c_out < 1 || ny < 1 || nx < 1) return MBEV_ERR_BAD_ARG;
  const int64_t G64 = static_cast<int64_t>(ny) * nx;
  if (G64 * batch > 0x7fffffffLL) return MBEV_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int G = static_cast<int>(G64);
  const size_t smem = sizeof(float) * kCells * (static_cast<size_t>(c_out) + 1);
  // 5 (default): k_scatter_run<2>, non-persistent; 6: the same with 512-cell runs; 1: k_scatter_warp (persistent,
  // 512-cell runs); 2: k_scatter_holes; 4: k_scatter_stage; 0: tile kernel with block barriers
  // (2 measured 2.3 ms against 0.97 ms for 1 on kitti_b16: single 32-byte sectors written apart from their line cost
  // DRAM read-modify-writes; kept selectable as a documented negative result.)
  static const int variant = getenv("MBEV_SCATTER") ? atoi(getenv("MBEV_SCATTER")) : 5;
  static const int s2_ctas_env = getenv("MBEV_SCATTER_CTAS") ? atoi(getenv("MBEV_SCATTER_CTAS")) : 0;
  static const int s2_ctas = getenv("MBEV_SCATTER_CTAS") ? atoi(getenv("MBEV_SCATTER_CTAS")) : 6;
  if (variant >= 1 && (G & 3) == 0 && (reinterpret_cast<uintptr_t>(canvas) & 15) == 0) {
    const int tiles_per_frame = (G + kS2Cells - 1) / kS2Cells;
    const int num_tiles = tiles_per_frame * batch;
    const int want_warps = kNumSMs * s2_ctas * (kThreads / 32);
    int csplit = 1;  // split the channels of a run over several warps until every resident warp has a task
    while (csplit < 16 && num_tiles * csplit < want_warps && c_out % (8 * csplit) == 0) csplit *= 2;
    const int tasks = num_tiles * csplit;
    const int blocks = std::min((tasks + kThreads / 32 - 1) / (kThreads / 32), kNumSMs * s2_ctas);
    // the hole kernel needs whole 32-byte sectors per lane pair (planes 32-byte aligned) and one channel quad per lane
    const bool holes = variant == 2 && (G & 7) == 0 && (reinterpret_cast<uintptr_t>(canvas) & 31) == 0 &&
                       (c_out & 3) == 0 && c_out <= 128;
    if (variant == 5 || variant == 6) {
      // default (5): 256-cell runs, one CTA per 8 tasks and NO grid-stride loop — the hardware CTA scheduler balances
      // the unequal runs better than a persistent grid did (0.94 -> 0.87 ms on kitti_b16); 6: 512-cell runs
      const int kk = variant == 5 ? 2 : 4;
      const int tpf = (G + 128 * kk - 1) / (128 * kk);
      const int nt = tpf * batch;
      int cs = 1;  // small batches: split the channels of a run over several warps
      while (cs < 16 && nt * cs < want_warps && c_out % (8 * cs) == 0) cs *= 2;
      const int64_t tasks5 = static_cast<int64_t>(nt) * cs;
      const int full = static_cast<int>((tasks5 + kThreads / 32 - 1) / (kThreads / 32));
      const int blk = s2_ctas_env > 0 ? std::min(full, kNumSMs * s2_ctas_env) : full;
      if (variant == 5) k_scatter_run<2><<<blk, kThreads, 0, stream>>>(feats, cell_table, c_out, G, tpf, nt, cs, canvas);
      else k_scatter_run<4><<<blk, kThreads, 0, stream>>>(feats, cell_table, c_out, G, tpf, nt, cs, canvas);
    } else if (variant == 4)  // 40 KB of shared memory per CTA: 5 CTAs per SM
      k_scatter_stage<<<std::min(blocks, kNumSMs * 5), kThreads, 0, stream>>>(feats, cell_table, c_out, G, tiles_per_frame,
                                                                             num_tiles, csplit, canvas);
    else if (holes)
      k_scatter_holes<<<blocks, kThreads, 0, stream>>>(feats, cell_table, c_out, G, tiles_per_frame, num_tiles, csplit,
                                                       canvas);
    else
      k_scatter_warp<<<blocks, kThreads, 0, stream>>>(feats, cell_table, c_out, G, tiles_per_frame, num_tiles, csplit,
                                                      canvas);
  } else if ((G & 3) == 0 && (reinterpret_cast<uintptr_t>(canvas) & 15) == 0 && smem <= 200 * 1024) {
    const int tiles_per_frame = (G + kCells - 1) / kCells;
    const int num_tiles = tiles_per_frame * batch;
    static bool attr_done = false;  // idempotent; a benign race sets the same value twice
    if (!attr_done) {
      MBEV_CUDA(cudaFuncSetAttribute(k_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_done = true;
    }
    int per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
    const int blocks = std::min(num_tiles, kNumSMs * per_sm);
    k_scatter<<<blocks, kThreads, smem, stream>>>(feats, cell_table, c_out, G, tiles_per_frame, num_tiles, canvas);
  } else {
    const long long total = static_cast<long long>(batch) * c_out * G;
    const int blocks = static_cast<int>(std::min<long long>((total + kThreads - 1) / kThreads, kNumSMs * 16));
    k_scatter_scalar<<<blocks, kThreads, 0, stream>>>(feats, cell_table, c_out, G, total, canvas);
  }
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

extern "C" int mbev_scatter_forward_bf16(const float *feats, const int32_t *cell_table, int batch, int c_out, int ny,
                                         int nx, void *canvas_bf16, void *stream_) {
  if (!cell_table || !canvas_bf16 || batch < 1 || c_out < 1 || ny < 1 || nx < 1) return MBEV_ERR_BAD_ARG;
  const int64_t G64 = static_cast<int64_t>(ny) * nx;
  if (G64 * batch > 0x7fffffffLL) return MBEV_ERR_UNSUPPORTED;
  if ((G64 & 3) || (reinterpret_cast<uintptr_t>(canvas_bf16) & 7)) return MBEV_ERR_UNSUPPORTED;
  const int G = static_cast<int>(G64);
  const int tiles_per_frame = (G + kS2Cells - 1) / kS2Cells;
  const int num_tiles = tiles_per_frame * batch;
  const int want_warps = kNumSMs * 6 * (kThreads / 32);
  int csplit = 1;
  while (csplit < 16 && num_tiles * csplit < want_warps && c_out % (8 * csplit) == 0) csplit *= 2;
  const int tasks = num_tiles * csplit;
  const int blocks = (tasks + kThreads / 32 - 1) / (kThreads / 32);  // no grid-stride: the CTA scheduler balances
  k_scatter_warp_bf16<<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      feats, cell_table, c_out, G, tiles_per_frame, num_tiles, csplit, static_cast<uint16_t *>(canvas_bf16));
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

extern "C" int mbev_scatter_split_supported(int ny, int nx, const float *canvas) {
  const int64_t G64 = static_cast<int64_t>(ny) * nx;
  return (G64 % 8 == 0) && ((reinterpret_cast<uintptr_t>(canvas) & 31) == 0) ? 1 : 0;
}

extern "C" int mbev_scatter_fill_empty(const int32_t *cell_table, int batch, int c_out, int ny, int nx,
                                       float *canvas, void *stream_) {
  if (!cell_table || !canvas || batch < 1 || c_out < 1 || ny < 1 || nx < 1) return MBEV_ERR_BAD_ARG;
  const int64_t G64 = static_cast<int64_t>(ny) * nx;
  if (G64 * batch > 0x7fffffffLL || !mbev_scatter_split_supported(ny, nx, canvas)) return MBEV_ERR_UNSUPPORTED;
  const int G = static_cast<int>(G64);
  const int tiles_per_frame = (G + kFillCells - 1) / kFillCells;
  const int num_tiles = tiles_per_frame * batch;
  // a streaming writer needs few warps per SM to saturate HBM; keep its footprint small so that it can share the
  // SMs with K2 without taking its issue slots
  static bool attr_done = false;
  if (!attr_done) {
    // same shared-memory / L1 split as the 200+ KB K2 kernels, so that the SM does not have to drain to switch
    // configuration and the two kernels can be co-resident
    MBEV_CUDA(cudaFuncSetAttribute(k_fill_empty, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
    attr_done = true;
  }
  static const int fill_ctas = getenv("MBEV_FILL_CTAS") ? atoi(getenv("MBEV_FILL_CTAS")) : 2;
  static const int fill_thr = getenv("MBEV_FILL_THREADS") ? atoi(getenv("MBEV_FILL_THREADS")) : kThreads;
  const int blocks = std::min(num_tiles, kNumSMs * fill_ctas);
  k_fill_empty<<<blocks, fill_thr, 0, static_cast<cudaStream_t>(stream_)>>>(cell_table, c_out, G, tiles_per_frame,
                                                                           num_tiles, canvas);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

extern "C" int mbev_scatter_occupied(const float *feats, const int32_t *coors, const int32_t *num_pillars_dev,
                                     int64_t pillar_capacity, const int32_t *cell_table, int batch, int c_out,
                                     int ny, int nx, float *canvas, void *stream_) {
  if (!cell_table || !canvas || !num_pillars_dev || batch < 1 || c_out < 1 || ny < 1 || nx < 1) return MBEV_ERR_BAD_ARG;
  if (pillar_capacity <= 0) return MBEV_OK;
  if (!feats || !coors) return MBEV_ERR_BAD_ARG;
  const int64_t G64 = static_cast<int64_t>(ny) * nx;
  if (G64 * batch > 0x7fffffffLL || !mbev_scatter_split_supported(ny, nx, canvas)) return MBEV_ERR_UNSUPPORTED;
  const int64_t want = (pillar_capacity + kThreads / 32 - 1) / (kThreads / 32);
  const int blocks = static_cast<int>(std::min<int64_t>(want, kNumSMs * 16));
  k_scatter_sectors<<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      feats, coors, num_pillars_dev, cell_table, batch, c_out, ny, nx, canvas);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}

extern "C" int mbev_scatter_backward(const float *dcanvas, const int32_t *cell_table, int batch, int c_out, int ny,
                                     int nx, float *dfeats, void *stream_) {
  if (!dcanvas || !cell_table || !dfeats || batch < 1 || c_out < 1 || ny < 1 || nx < 1) return MBEV_ERR_BAD_ARG;
  const int64_t G64 = static_cast<int64_t>(ny) * nx;
  if (G64 * batch > 0x7fffffffLL) return MBEV_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int G = static_cast<int>(G64);
  const int tiles_per_frame = (G + kCells - 1) / kCells;
  const int num_tiles = tiles_per_frame * batch;
  const int blocks = std::min(num_tiles, kNumSMs * 8);
  k_gather_bwd<<<blocks, kThreads, 0, stream>>>(dcanvas, cell_table, c_out, G, tiles_per_frame, num_tiles, dfeats);
  MBEV_CHECK_LAUNCH();
  return MBEV_OK;
}
