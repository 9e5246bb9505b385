// Shared host/device helpers for the mask_bev_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/mask_bev_b200.h"

namespace mbev {

extern std::atomic<int64_t> g_launches;  // defined in api.cu

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace.
struct Carver {
  char *base;
  size_t off = 0;
  explicit Carver(void *p) : base(static_cast<char *>(p)) {}
  template <typename T>
  T *take(size_t n) {
    T *p = reinterpret_cast<T *>(base ? base + off : nullptr);
    off = align_up(off + n * sizeof(T));
    return p;
  }
};

#define MBEV_CHECK_LAUNCH()                                  \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return static_cast<int>(e__);    \
    ::mbev::g_launches.fetch_add(1, std::memory_order_relaxed); \
  } while (0)

#define MBEV_CUDA(call)                                      \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return static_cast<int>(e__);    \
  } while (0)

// Frame row offsets passed by value (kernel parameter space; read in place via __grid_constant__).
struct Frames {
  int32_t batch;
  int32_t off[MBEV_MAX_BATCH + 1];
};

// Device-side copy of MbevGeometry with derived fields.
struct GeoK {
  float lo[3], hi[3], vs[3];
  int32_t nx, ny, nz;
  int32_t cells;  // nx*ny*nz per frame
  int32_t T, V, C, strict;
};

inline GeoK make_geok(const MbevGeometry &g) {
  GeoK k;
  for (int j = 0; j < 3; ++j) {
    k.lo[j] = g.range[j];
    k.hi[j] = g.range[3 + j];
    k.vs[j] = g.voxel[j];
  }
  k.nx = g.grid[0];
  k.ny = g.grid[1];
  k.nz = g.grid[2];
  k.cells = g.grid[0] * g.grid[1] * g.grid[2];
  k.T = g.max_points;
  k.V = g.max_voxels;
  k.C = g.num_feats;
  k.strict = g.strict_filter;
  return k;
}

#ifdef __CUDACC__
// frame of global point row i: largest f with off[f] <= i  (s_off: batch+1 entries, non-decreasing)
__device__ __forceinline__ int frame_of(const int *s_off, int batch, int i) {
  int lo = 0, hi = batch;  // invariant: off[lo] <= i < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s_off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void st_global_v4_stream(float *p, float4 v) {
  // streaming store: written once, never re-read by this kernel
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// same store without the compiler-level memory clobber: loads of later iterations may be scheduled above it
// (use only where the stored range cannot alias anything the kernel reads)
__device__ __forceinline__ void st_global_v4_stream_nc(float *p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
#endif

}  // namespace mbev
