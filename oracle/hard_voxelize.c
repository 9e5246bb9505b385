/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C restatement of the CPU hard-voxelization loop that MaskBEV reaches through
 *   /root/reference/mask_bev/models/encoders/mask_bev_encoders.py:69  (Voxelization(...) built)
 *   /root/reference/mask_bev/models/encoders/mask_bev_encoders.py:100 (self._voxel_layer(res))
 * The arithmetic lives in the un-vendored dependency mmcv==2.0.0 (Dockerfile:25):
 *   mmcv/ops/csrc/pytorch/cpu/voxelization.cpp  dynamic_voxelize_forward_cpu_kernel +
 *   hard_voxelize_forward_cpu_kernel, restated from its published algorithm.
 * PARITY UNPINNED: the reference ships no golden vectors for this path and mmcv cannot be
 * installed here; this restatement is pinned only by the hand-derived vectors in
 * tests/golden/  and by property tests.
 *
 * Also holds the range filter of mask_bev_encoders.py:113-117 (strict compares in float32).
 *
 * Build: see oracle/Makefile  ->  oracle/_build/liboracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* mask_bev_encoders.py:113-117 — keep point iff lo < v < hi on x,y,z (strict, float32).
 * Writes the source index of every kept point, in input order. Returns the kept count. */
int64_t mbev_oracle_filter_in_range(const float *pts, int64_t n, int c, const float range[6],
                                    int64_t *keep_idx) {
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float *p = pts + i * c;
    int ok = 1;
    for (int j = 0; j < 3; ++j) ok = ok && (range[j] < p[j]) && (p[j] < range[3 + j]);
    if (ok) keep_idx[m++] = i;
  }
  return m;
}

/* grid_size_j = round((hi_j - lo_j) / vs_j) in float32 (mmcv/ops/voxelize.py Voxelization.__init__). */
void mbev_oracle_grid_size(const float range[6], const float vs[3], int grid[3]) {
  for (int j = 0; j < 3; ++j) grid[j] = (int)roundf((range[3 + j] - range[j]) / vs[j]);
}

/*
 * hard_voxelize_forward (CPU), NDim = 3.
 *   pts        (n, c) float32, already range-filtered by the caller (or not: out-of-grid points are skipped)
 *   voxels     (max_voxels, T, c) float32, caller-zeroed
 *   coors      (max_voxels, 3) int32 (z, y, x), caller-zeroed
 *   num_points (max_voxels) int32, caller-zeroed
 *   kept_idx   (max_voxels, T) int64 or NULL: source row of every stored point, -1 padded (caller-filled)
 * Returns voxel_num.
 */
int64_t mbev_oracle_hard_voxelize(const float *pts, int64_t n, int c, const float vs[3],
                                  const float range[6], int T, int64_t max_voxels, float *voxels,
                                  int32_t *coors, int32_t *num_points, int64_t *kept_idx) {
  int grid[3];
  mbev_oracle_grid_size(range, vs, grid);
  const int64_t cells = (int64_t)grid[0] * grid[1] * grid[2];
  int32_t *table = (int32_t *)malloc(sizeof(int32_t) * (size_t)cells);
  if (!table) return -1;
  memset(table, 0xff, sizeof(int32_t) * (size_t)cells); /* -1 */
  int64_t voxel_num = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float *p = pts + i * c;
    int cc[3];
    int failed = 0;
    for (int j = 0; j < 3; ++j) {
      /* float32 subtract, float32 divide, floor */
      const float d = p[j] - range[j];
      const float q = d / vs[j];
      const int v = (int)floorf(q);
      if (!(q == q)) { failed = 1; break; } /* NaN: (int)NaN is UB in C; upstream drops via c<0||c>=g on x86 INT_MIN */
      if (v < 0 || v >= grid[j]) { failed = 1; break; }
      cc[j] = v;
    }
    if (failed) continue;
    const int64_t cell = ((int64_t)cc[2] * grid[1] + cc[1]) * grid[0] + cc[0];
    int32_t vid = table[cell];
    if (vid == -1) {
      if (max_voxels != -1 && voxel_num >= max_voxels) continue;
      vid = (int32_t)voxel_num++;
      table[cell] = vid;
      coors[vid * 3 + 0] = cc[2];
      coors[vid * 3 + 1] = cc[1];
      coors[vid * 3 + 2] = cc[0];
    }
    const int num = num_points[vid];
    if (T == -1 || num < T) {
      memcpy(voxels + ((int64_t)vid * T + num) * c, p, sizeof(float) * (size_t)c);
      if (kept_idx) kept_idx[(int64_t)vid * T + num] = i;
      num_points[vid] = num + 1;
    }
  }
  free(table);
  return voxel_num;
}

/* PointPillarsScatter.forward_batch restated: canvas (B, C, ny*nx) caller-zeroed. */
void mbev_oracle_scatter(const float *feat, const int32_t *coors4, int64_t p, int c, int batch, int ny,
                         int nx, float *canvas) {
  const int64_t g = (int64_t)ny * nx;
  for (int64_t i = 0; i < p; ++i) {
    const int b = coors4[i * 4 + 0];
    if (b < 0 || b >= batch) continue;
    const int64_t idx = (int64_t)coors4[i * 4 + 2] * nx + coors4[i * 4 + 3];
    for (int k = 0; k < c; ++k) canvas[((int64_t)b * c + k) * g + idx] = feat[i * c + k];
  }
}
