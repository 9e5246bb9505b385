/*
 * mask_bev_b200 — C ABI of the B200-native point-cloud -> BEV front end.
 *
 * Drop-in boundary for MaskBEV's encoder path
 *   /root/reference/mask_bev/models/encoders/mask_bev_encoders.py:69-75 (construction),
 *   :95-111 (voxelize), :113-117 (range filter), :119-120 (encode), :122-123 (middle_encode).
 * The reference binds this path through two compiled/third-party interfaces that are NOT in its tree:
 *   mmcv==2.0.0   mmcv.ops.Voxelization -> ext_module.hard_voxelize_forward      (Dockerfile:25)
 *   mmdet3d==1.1.0 PillarFeatureNet / PFNLayer / PointPillarsScatter (torch ops)  (Dockerfile:28)
 * Each entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in _host. The caller (PyTorch) owns every
 *     buffer, including the workspace; the library never allocates or frees device memory, keeps no global
 *     state and reads no environment variable. All work is enqueued on `stream`; no entry point synchronises.
 *   - Return value: 0 ok; negative = MbevStatus (bad argument / unsupported shape / workspace too small);
 *     positive = cudaError_t of a failed launch. No exceptions cross this boundary.
 *   - `stream` is a cudaStream_t passed as void* so that this header needs no CUDA include.
 */
#ifndef MASK_BEV_B200_H_
#define MASK_BEV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MBEV_API __attribute__((visibility("default")))
#else
#define MBEV_API
#endif

#define MBEV_ABI_VERSION 11
#define MBEV_MAX_BATCH 128  /* frames per call */
#define MBEV_MAX_LAYERS 4   /* PFN layers */
#define MBEV_MAX_UNITS 128  /* widest PFNLayer.units supported by the fused kernel */
#define MBEV_MAX_POINT_DIM 8

typedef enum MbevStatus {
  MBEV_OK = 0,
  MBEV_ERR_BAD_ARG = -1,
  MBEV_ERR_UNSUPPORTED = -2,
  MBEV_ERR_WORKSPACE = -3,
  MBEV_ERR_NO_DEVICE = -4
} MbevStatus;

/* Voxel-layer geometry: the arguments of mmcv.ops.Voxelization(voxel_size, point_cloud_range,
 * max_num_points, max_voxels, deterministic=True) as built at mask_bev_encoders.py:67-69, all float32
 * exactly as `torch.tensor(voxel_size)` / `torch.tensor(coors_range)` hand them to the compiled op. */
typedef struct MbevGeometry {
  float range[6];       /* x0,y0,z0,x1,y1,z1 */
  float voxel[3];       /* vx,vy,vz */
  int32_t grid[3];      /* nx,ny,nz = round((hi-lo)/vs) in float32 */
  int32_t max_points;   /* T */
  int32_t max_voxels;   /* V (per frame) */
  int32_t num_feats;    /* C: floats per point row */
  int32_t strict_filter; /* 1: apply mask_bev_encoders.py:113-117 (lo < v < hi, strict) before the voxel test */
} MbevGeometry;

/* One PFN stack: mmdet3d PillarFeatureNet(in_channels, feat_channels, with_distance, with_cluster_center,
 * with_voxel_center, voxel_size, point_cloud_range, norm_cfg=BN1d(eps 1e-3, momentum 0.01), mode='max',
 * legacy=True) as built at mask_bev_encoders.py:70-72. */
typedef struct MbevPfnParams {
  int32_t num_layers;
  int32_t in_dim[MBEV_MAX_LAYERS];   /* in_0 = C+3+vcd+dist ; in_l = 2*units_{l-1} */
  int32_t units[MBEV_MAX_LAYERS];    /* PFNLayer.units */
  const float *weight[MBEV_MAX_LAYERS]; /* pfn_layers.l.linear.weight  (units_l, in_l) row-major */
  const float *scale[MBEV_MAX_LAYERS];  /* eval: gamma/sqrt(running_var+eps) ; train: written by the library */
  const float *shift[MBEV_MAX_LAYERS];  /* eval: beta - running_mean*scale   ; train: written by the library */
  int32_t with_cluster_center, with_voxel_center, with_distance, legacy;
  int32_t voxel_center_dims;         /* 3 (mmdet3d 1.1.0) or 2 (mmdet3d 0.x fossil, mask_bev_encoders.py:165-166) */
  float vx, vy, vz, x_offset, y_offset, z_offset; /* float32(vx), float32(vx/2 + x0) ... */
  int32_t gemm_path; /* forward Linear layers: MBEV_GEMM_AUTO (tcgen05 3xTF32 when the stack fits it, else fp32 FMA),
                        MBEV_GEMM_FMA, MBEV_GEMM_TCGEN05 (MBEV_ERR_UNSUPPORTED if the stack does not fit),
                        MBEV_GEMM_TCGEN05_BF16 (layers >= 1 as single-pass bf16 tensor-core MMAs with fp32 accumulation,
                        layer 0 — raw coordinates in — stays 3xTF32; eval mode, T <= 32, every units % 32 == 0 and
                        <= 64 for non-last layers; 1e-2 tolerance class, BASELINE.json north star "1e-2 in bf16") */
} MbevPfnParams;

enum { MBEV_GEMM_AUTO = 0, MBEV_GEMM_FMA = 1, MBEV_GEMM_TCGEN05 = 2, MBEV_GEMM_TCGEN05_BF16 = 3 };

/* Version / capability probes (host only, no GPU needed). */
MBEV_API int mbev_abi_version(void);
MBEV_API const char *mbev_build_info(void);
MBEV_API const char *mbev_status_string(int status);

/* ------------------------------------------------------------------------------------------------
 * K1  hard voxelisation of a batch of frames.
 * Replaces: per-frame `_filter_in_range` + `mmcv.ops.Voxelization.forward` + batch-index padding
 *           (mask_bev_encoders.py:95-111; mmcv hard_voxelize_forward, deterministic=True).
 *   points            (total_points, C) float32, frames concatenated
 *   frame_offsets_host (batch+1) int64 HOST array, row offsets of each frame in `points`
 *   cell_table        (batch * nz*ny*nx) int32 OUT: global pillar id of each cell, -1 = empty
 *                     (the occupancy mask and the scatter's inverse map)
 *   coors             (pillar_capacity, 4) int32 OUT (b, z, y, x)
 *   num_points        (pillar_capacity) int32 OUT, 1..T
 *   kept_idx          (pillar_capacity, T) int32 OUT: row in `points` of slot t; slots >= num_points undefined
 *   pillar_base       (batch+1) int32 OUT: first global pillar id of each frame; [batch] = total P
 * Pillars are numbered frame by frame, inside a frame in order of first appearance; slots in input order.
 * pillar_capacity must be >= mbev_pillar_capacity(): sum over frames of min(points of the frame, V, cells) — a
 * frame cannot hold more pillars than it has points, than max_voxels, or than cells (host only, no GPU needed;
 * -1 on a bad argument).
 * ---------------------------------------------------------------------------------------------- */
MBEV_API int64_t mbev_pillar_capacity(const int64_t *frame_offsets_host, int batch, const MbevGeometry *geo);
MBEV_API int mbev_voxelize_workspace_bytes(const MbevGeometry *geo, int batch, int64_t total_points, size_t *bytes);
MBEV_API int mbev_voxelize(const float *points, const int64_t *frame_offsets_host, int batch, const MbevGeometry *geo,
                  int32_t *cell_table, int32_t *coors, int32_t *num_points, int32_t *kept_idx,
                  int32_t *pillar_base, int64_t pillar_capacity, void *workspace, size_t workspace_bytes,
                  void *stream);

/* F4: the point side of the reference's augmentations applied in K1's load stage, in the order the
 * training configs list them (configs/training/semantic_kitti/01*.yml:34-49; classes in
 * mask_bev/augmentations/semantic_kitti_mask_augmentations.py): RandomDropPoints (:152-162) -> Flip (:44-56) ->
 * RandomRotate (:73-101, float64 like numpy's R @ p) -> JitterPoints (:116-149, intensity clipped to [0, 1]).
 * Per-frame decisions are drawn by the host (one MbevFrameAugment per frame, DEVICE array); per-point randomness is
 * replayed from drop_u / noise when given (bit-exact against a numpy stream) or generated in the kernel from
 * (seed, point row). points_out (total_points, C) receives the augmented cloud: kept_idx indexes IT, so the PFN must
 * gather from points_out. A dropped point keeps its row (and is simply not voxelised). */
typedef struct MbevFrameAugment {
  double cos_t, sin_t;     /* rotation about z; used when rotate != 0 */
  float drop_prob;         /* per-point drop probability of this frame, 0 = keep all */
  float jitter_std[4];     /* generator mode: std of the x, y, z, intensity noise */
  float jitter_max[4];     /* generator mode: clip of the noise, <= 0 = none */
  int32_t flip_x, flip_y;  /* x -> -x, y -> -y */
  int32_t rotate;
  int32_t jitter;          /* add noise (given or generated), then clip intensity */
} MbevFrameAugment;
typedef struct MbevAugment {
  const MbevFrameAugment *frames; /* DEVICE, batch entries */
  const float *drop_u;            /* DEVICE (total_points) uniforms in [0, 1): dropped iff u < drop_prob; NULL = generated */
  const double *noise;            /* DEVICE (total_points, C) additive noise, scaled and clipped; NULL = generated */
  uint64_t seed;
  float *points_out;              /* DEVICE (total_points, C) */
} MbevAugment;
MBEV_API int mbev_voxelize_augmented(const float *points, const int64_t *frame_offsets_host, int batch,
                                     const MbevGeometry *geo, const MbevAugment *aug, int32_t *cell_table,
                                     int32_t *coors, int32_t *num_points, int32_t *kept_idx, int32_t *pillar_base,
                                     int64_t pillar_capacity, void *workspace, size_t workspace_bytes, void *stream);


/* Materialise the zero-padded (P, T, C) voxel tensor that mmcv's op returns (voxels_out) and, optionally,
 * the -1 padded kept-index matrix rebased to frame-local rows. num_pillars_dev: device int32 (total P). */
MBEV_API int mbev_gather_voxels(const float *points, const int32_t *kept_idx, const int32_t *num_points,
                       const int32_t *num_pillars_dev, int64_t pillar_capacity, int T, int C, float *voxels,
                       void *stream);

/* ------------------------------------------------------------------------------------------------
 * K2  pillar feature net forward (decorate -> L x (Linear, BN, ReLU, max over T [, concat]) ).
 * Replaces: mmdet3d PillarFeatureNet.forward / PFNLayer.forward (mask_bev_encoders.py:119-120).
 *   rows       float32 row store: either `points` (sparse mode, kept_idx != NULL) or the padded voxel tensor
 *              (P,T,C) (dense mode, kept_idx == NULL; row of slot (p,t) is p*T+t)
 *   num_pillars_dev  device int32: number of pillars to process (no host sync on the fused path)
 *   feats      (pillar_capacity, units[L-1]) float32 OUT
 * Eval mode uses params->scale/shift (BatchNorm folded with running statistics).
 * Train mode (mbev_pfn_forward_train) computes batch statistics over all P*T slots exactly as BatchNorm1d
 * does on the padded tensor, writes the folded scale/shift it used into `scale_shift_out`
 * (L x 2 x MBEV_MAX_UNITS floats) and mean / biased variance into `batch_stats_out` (same shape) so the
 * host can update running statistics with momentum 0.01 and the unbiased variance.
 * Two device implementations share this contract (params->gemm_path): the Linear layers either run on the
 * tcgen05 tensor cores with every product split 3xTF32 (fp32-accurate: hi*hi + hi*lo + lo*hi in an fp32 TMEM
 * accumulator), or on the fp32 FMA pipe. mbev_pfn_path() reports which one a call would take.
 * ---------------------------------------------------------------------------------------------- */
MBEV_API int mbev_pfn_workspace_bytes(const MbevPfnParams *params, int T, int64_t pillar_capacity, int train,
                             size_t *bytes);
MBEV_API int mbev_pfn_path(const MbevPfnParams *params, int T); /* MBEV_GEMM_FMA / _TCGEN05 / _TCGEN05_BF16, or < 0 */
MBEV_API int mbev_pfn_forward(const float *rows, int C, const int32_t *kept_idx, const int32_t *num_points,
                     const int32_t *coors, const int32_t *num_pillars_dev, int64_t pillar_capacity, int T,
                     const MbevPfnParams *params, float *feats, void *workspace, size_t workspace_bytes,
                     void *stream);
MBEV_API int mbev_pfn_forward_train(const float *rows, int C, const int32_t *kept_idx, const int32_t *num_points,
                           const int32_t *coors, const int32_t *num_pillars_dev, int64_t pillar_capacity, int T,
                           const MbevPfnParams *params, const float *const *gamma, const float *const *beta,
                           float eps, float *feats, float *scale_shift_out, float *batch_stats_out,
                           void *workspace, size_t workspace_bytes, void *stream);
/* Backward of the train- or eval-mode forward w.r.t. the parameters (raw points carry no gradient,
 * SURVEY.md §3.4; the reference gets this from autograd over the dense op sequence).
 *   rows_capacity_hint  upper bound of the compact row count R = sum_p (n_p + [n_p < T]) used to size the
 *                       workspace (e.g. total_points + pillar_capacity); <= 0 means pillar_capacity * (T+1)
 *   scale_shift  (L,2,MBEV_MAX_UNITS) folded scale/shift the forward used (forward_train output, or the eval fold)
 *   batch_stats  (L,2,MBEV_MAX_UNITS) mean / variance the forward normalised with (batch or running)
 *   train        1: BatchNorm batch-statistics backward (mean/var depend on the rows); 0: frozen statistics
 *   dfeats       (pillar_capacity, units[L-1])
 * Outputs (HOST arrays of L device pointers): dweight[l] (units_l, in_l), dgamma[l], dbeta[l] (units_l).
 * Reductions use a fixed order: results are run-to-run identical. */
MBEV_API int mbev_pfn_backward_workspace_bytes(const MbevPfnParams *params, int T, int64_t pillar_capacity,
                                               int64_t rows_capacity_hint, size_t *bytes);
MBEV_API int mbev_pfn_backward(const float *rows, int C, const int32_t *kept_idx, const int32_t *num_points,
                               const int32_t *coors, const int32_t *num_pillars_dev, int64_t pillar_capacity, int T,
                               int64_t rows_capacity_hint, const MbevPfnParams *params, const float *scale_shift,
                               const float *batch_stats, float eps, int train, const float *dfeats,
                               float *const *dweight, float *const *dgamma, float *const *dbeta, void *workspace,
                               size_t workspace_bytes, void *stream);

/* A TRAINING STEP's pair (train-mode BatchNorm): the forward runs once in compact row space (fp32 FMA products, the
 * rows BatchNorm's backward differentiates) and LEAVES its activations in `workspace`; the backward consumes them
 * instead of recomputing the forward. Replaces the same reference lines as mbev_pfn_forward_train + mbev_pfn_backward
 * (PFNLayer.forward in train mode and its autograd, mask_bev_encoders.py:119-120); outputs and tolerances are the same.
 *   workspace  mbev_pfn_backward_workspace_bytes(params, T, pillar_capacity, rows_capacity_hint) bytes, caller-owned;
 *              the caller must keep it untouched between the two calls and pass the SAME pillar_capacity, T,
 *              rows_capacity_hint and params to both (the layout is a function of those). */
MBEV_API int mbev_pfn_forward_train_rows(const float *rows, int C, const int32_t *kept_idx, const int32_t *num_points,
                                         const int32_t *coors, const int32_t *num_pillars_dev, int64_t pillar_capacity,
                                         int T, int64_t rows_capacity_hint, const MbevPfnParams *params,
                                         const float *const *gamma, const float *const *beta, float eps, float *feats,
                                         float *scale_shift_out, float *batch_stats_out, void *workspace,
                                         size_t workspace_bytes, void *stream);
MBEV_API int mbev_pfn_backward_rows(const int32_t *num_pillars_dev, int64_t pillar_capacity, int C, int T,
                                    int64_t rows_capacity_hint, const MbevPfnParams *params, float eps,
                                    const float *dfeats, float *const *dweight, float *const *dgamma,
                                    float *const *dbeta, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * K3  scatter to the dense BEV canvas, one streaming pass (zeros and features written exactly once).
 * Replaces: mmdet3d PointPillarsScatter.forward_batch (mask_bev_encoders.py:122-123).
 *   cell_table (batch * ny*nx) int32: pillar id per cell or -1 (from mbev_voxelize or mbev_build_cell_table)
 *   canvas     (batch, C_out, ny, nx) float32 OUT, NCHW contiguous (the reference's layout)
 * K3' backward: dfeats[p, :] = dcanvas[b, :, y, x].
 * ---------------------------------------------------------------------------------------------- */
MBEV_API int mbev_build_cell_table(const int32_t *coors, const int32_t *num_pillars_dev, int64_t pillar_capacity,
                          int batch, int ny, int nx, int32_t *cell_table, void *stream);
MBEV_API int mbev_scatter_forward(const float *feats, const int32_t *cell_table, int batch, int c_out, int ny, int nx,
                         float *canvas, void *stream);
/* The same scatter with the stores handed to the TMA engine (cp.async.bulk shared -> global): a persistent grid of
 * ctas_per_sm x 148 CTAs of 128 threads / <= 64 registers / 17 KB of shared memory. With ctas_per_sm = 1 such a CTA fits
 * on an SM next to K2's persistent CTA, so this scatter can run UNDER the pillar feature net of the next batch
 * (mbev_encode_batch_pipelined); stand-alone use wants ctas_per_sm >= 4. Bit-identical canvas. Needs C_out % 4 == 0,
 * ny*nx % 4 == 0, 16-byte aligned feats / canvas (probe: mbev_scatter_stream_supported; MBEV_ERR_UNSUPPORTED otherwise). */
MBEV_API int mbev_scatter_stream_supported(int c_out, int ny, int nx, const float *canvas);
MBEV_API int mbev_scatter_forward_stream(const float *feats, const int32_t *cell_table, int batch, int c_out, int ny,
                                         int nx, float *canvas, int ctas_per_sm, void *stream);
/* Same scatter with a bfloat16 canvas (BASELINE.json config 4; north star tolerance 1e-2 in bf16): every value is
 * the fp32 feature rounded to nearest-even bf16, so the result equals the fp32 canvas cast to bf16 bit for bit.
 * canvas_bf16: (batch, C_out, ny, nx) bfloat16; needs ny*nx % 4 == 0 and an 8-byte aligned canvas. Forward only. */
MBEV_API int mbev_scatter_forward_bf16(const float *feats, const int32_t *cell_table, int batch, int c_out, int ny,
                                       int nx, void *canvas_bf16, void *stream);
/* Channels-last canvas (north star item 3): canvas_nhwc is (batch, ny, nx, C_out) float32 — i.e. the (batch, C_out, ny,
 * nx) tensor in torch.channels_last memory format. One pass over the cell table; every cell is one contiguous
 * 4*C_out-byte piece (the pillar's feature row or zeros), every byte written once. Needs C_out % 4 == 0 and 16-byte
 * aligned feats / canvas. Backward: dfeats[p, :] = dcanvas_nhwc[b, y, x, :] (rows >= the pillar count and rows whose
 * cell the table gives to another pillar get zeros). */
MBEV_API int mbev_scatter_forward_nhwc(const float *feats, const int32_t *cell_table, int batch, int c_out, int ny,
                                       int nx, float *canvas_nhwc, void *stream);
MBEV_API int mbev_scatter_backward(const float *dcanvas, const int32_t *cell_table, int batch, int c_out, int ny,
                          int nx, float *dfeats, void *stream);
MBEV_API int mbev_scatter_backward_nhwc(const float *dcanvas_nhwc, const int32_t *cell_table, const int32_t *coors,
                                        const int32_t *num_pillars_dev, int64_t rows, int batch, int c_out, int ny,
                                        int nx, float *dfeats, void *stream);

/* ------------------------------------------------------------------------------------------------
 * K3+LN  scatter fused with the LayerNorm that follows it (forward and backward) — SURVEY.md §8 row f1.
 * Replaces: `self._layer_norm(self.middle_encode(...))` with nn.LayerNorm([C, ny, nx], eps=1e-3)
 *           (mask_bev_encoders.py:75, 91-92): per frame, normalisation over all C*ny*nx elements, element-wise affine.
 * The statistics come from the pillar features alone (all other cells are exact zeros; fp64, fixed order), then one
 * streaming pass writes ((x - mean_b) * rstd_b) * weight + bias: weight and bias are read once per BATCH, the canvas
 * is written once.
 *   pillar_base (batch+1) device int32 from mbev_voxelize (pillars of frame b are [pillar_base[b], pillar_base[b+1]))
 *   ln_weight, ln_bias (C, ny, nx) float32;  out (batch, C, ny, nx) float32;  stats_out (batch, 2) = mean, rstd
 * Needs ny*nx % 4 == 0 and 16-byte aligned out / weight / bias (probe: mbev_scatter_layernorm_supported).
 * walk: the schedule of the streaming pass, same arithmetic per element (bit-identical results):
 *   MBEV_LN_WALK_RUNS    a warp owns (256-cell run, channel chunk, ONE frame); weight / bias re-read from L2 per frame
 *   MBEV_LN_WALK_FRAMES  a warp owns (128 cells, 4 channels) and walks the frames with weight / bias in registers
 *                        (needs C_out % 4 == 0 and 16-byte aligned feats)
 * ---------------------------------------------------------------------------------------------- */
enum { MBEV_LN_WALK_RUNS = 0, MBEV_LN_WALK_FRAMES = 1 };
MBEV_API int mbev_scatter_layernorm_supported(int batch, int c_out, int ny, int nx, const float *out,
                                              const float *ln_weight, const float *ln_bias);
MBEV_API int mbev_scatter_layernorm_workspace_bytes(int batch, size_t *bytes);
MBEV_API int mbev_scatter_layernorm_forward(const float *feats, const int32_t *cell_table, const int32_t *pillar_base,
                                            int batch, int c_out, int ny, int nx, const float *ln_weight,
                                            const float *ln_bias, float eps, int walk, float *out,
                                            float *stats_out, void *workspace, size_t workspace_bytes, void *stream);

/* Backward of the above (autograd of mask_bev_encoders.py:91-92; the reference has no code for it, torch derives it).
 * With xh = (x - mean_b) * rstd_b, g = dout * weight, M = C*ny*nx:
 *   dbias = sum_b dout;  dweight = sum_b dout * xh  (dense);  dfeats[p,:] = rstd_b * (g - mean(g) - xh * mean(g*xh))
 * at the pillar's cell. dout is read once, dweight / dbias are written once; the per-frame sums are fp64 and
 * fixed-order (run-to-run identical).
 *   dout (batch, C, ny, nx); feats (pillar_capacity, C) = the forward's input; coors (pillar_capacity, 4) (b,z,y,x);
 *   num_pillars_dev = device pillar count (pillar_base + batch); stats (batch, 2) = the forward's stats_out;
 *   dfeats (pillar_capacity, C): rows of pillars that are not in cell_table and rows >= the pillar count get zeros;
 *   dweight, dbias (C, ny, nx).
 * Needs C % 4 == 0, ny*nx % 4 == 0 (probe: mbev_scatter_layernorm_backward_supported) and 16-byte aligned pointers
 * (MBEV_ERR_UNSUPPORTED otherwise). */
MBEV_API int mbev_scatter_layernorm_backward_supported(int batch, int c_out, int ny, int nx);
MBEV_API int mbev_scatter_layernorm_backward_workspace_bytes(int batch, int c_out, int ny, int nx, size_t *bytes);
MBEV_API int mbev_scatter_layernorm_backward(const float *dout, const float *feats, const int32_t *cell_table,
                                             const int32_t *coors, const int32_t *num_pillars_dev,
                                             int64_t pillar_capacity, int batch, int c_out, int ny, int nx,
                                             const float *ln_weight, const float *stats, float *dfeats,
                                             float *dweight, float *dbias, void *workspace, size_t workspace_bytes,
                                             void *stream);

/* ------------------------------------------------------------------------------------------------
 * F2 — Swin patch embedding on pillars. Replaces, for the first consumer of the pseudo image
 * (/root/reference/mask_bev/models/networks/swin/swin.py:578-586 construction, :745-746 call; mmdet PatchEmbed,
 * swin.py:13), the chain  nn.LayerNorm([C,ny,nx], eps) (mask_bev_encoders.py:75, 92)  ->  corner padding ->
 * Conv2d(C, E, kernel = stride = patch, bias) -> flatten(2).transpose(1, 2) [-> nn.LayerNorm(E)]
 * WITHOUT materialising the canvas: tokens (batch, Hp*Wp, E), Hp = ceil(ny / patch), Wp = ceil(nx / patch).
 *   feats (pillar_capacity, C), coors (pillar_capacity, 4) (b,z,y,x), cell_table (batch, ny*nx), pillar_base
 *   (batch + 1) as K1 / K2 produce them;
 *   ln_weight_cl (ny*nx, C): the LayerNorm weight in channels-last order;
 *   w_img: mbev_patch_embed_prepare_weights(conv weight (E, C, patch, patch)) -> patch*patch*2*E*C floats;
 *   p0, p1 (Hp*Wp, E): parameter-only images conv(pad(ln_bias)) + conv bias and conv(pad(ln_weight)), channels-last;
 *   norm_weight / norm_bias (E) or both NULL (patch_norm = False); stats_out (batch, 2) = mean, rstd of the canvas.
 * Supported (probe): C in {32, 64, 128}, E a multiple of 32 <= 256 with 2*E*C*4 bytes of weights fitting shared
 * memory, patch <= 8. fp32 parity through 3xTF32 tcgen05 MMAs; results are run-to-run identical. */
MBEV_API int mbev_patch_embed_supported(int batch, int C, int ny, int nx, int patch, int embed_dims);
MBEV_API int mbev_patch_embed_workspace_bytes(int batch, int64_t pillar_capacity, int patch, int embed_dims,
                                              size_t *bytes);
MBEV_API int mbev_patch_embed_prepare_weights(const float *conv_weight, int embed_dims, int C, int patch, float *w_img,
                                              void *stream);
MBEV_API int mbev_patch_embed_forward(const float *feats, const int32_t *coors, const int32_t *cell_table,
                                      const int32_t *pillar_base, int64_t pillar_capacity, int batch, int C, int ny,
                                      int nx, int patch, int embed_dims, const float *ln_weight_cl, float ln_eps,
                                      const float *w_img, const float *p0, const float *p1, const float *norm_weight,
                                      const float *norm_bias, float norm_eps, float *tokens, float *stats_out,
                                      void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Fused batch entry (additive): K1 -> K2(eval) -> K3 on one stream, no host sync.
 * Equivalent to MaskBevEncoder.forward (mask_bev_encoders.py:77-91) without the trailing LayerNorm.
 * ---------------------------------------------------------------------------------------------- */
MBEV_API int mbev_encode_batch_workspace_bytes(const MbevGeometry *geo, const MbevPfnParams *params, int batch,
                                      int64_t total_points, int64_t pillar_capacity, size_t *bytes);
MBEV_API int mbev_encode_batch(const float *points, const int64_t *frame_offsets_host, int batch,
                      const MbevGeometry *geo, const MbevPfnParams *params, int32_t *cell_table,
                      int32_t *coors, int32_t *num_points, int32_t *kept_idx, int32_t *pillar_base,
                      int64_t pillar_capacity, float *feats, float *canvas, void *workspace,
                      size_t workspace_bytes, void *stream);

/* Same, with HOST input: copies `points_host` (pinned or pageable) to `points_dev` on `stream` first. */
MBEV_API int mbev_encode_batch_host(const float *points_host, float *points_dev, const int64_t *frame_offsets_host,
                           int batch, const MbevGeometry *geo, const MbevPfnParams *params,
                           int32_t *cell_table, int32_t *coors, int32_t *num_points, int32_t *kept_idx,
                           int32_t *pillar_base, int64_t pillar_capacity, float *feats, float *canvas,
                           void *workspace, size_t workspace_bytes, void *stream);

/* Three-stage pipeline for a stream of batches — the call bench.py times (`value` with points_host = NULL: points
 * already resident in points_dev; `e2e` with pinned host points).
 *   stage 1 on `prep_stream`: wait `ev_consumed` (the batch that used THIS buffer set two calls ago has left K3), copy
 *           `points_host` to `points_dev` when points_host != NULL, K1; record `ev_ready`.
 *   stage 2 on `pfn_stream` : wait `ev_ready`, K2 into `feats`; record `ev_feats`.
 *   stage 3 on `stream`     : wait `ev_feats`, K3 into `canvas`; record `ev_consumed`. The canvas is ordered by the
 *           caller's `stream` like any other output.
 * The caller alternates TWO buffer sets (points_dev when copying, cell_table, coors, num_points, kept_idx, pillar_base,
 * vox_workspace, feats, and the three events): K1 (+ the H2D copy) of batch i+2, K2 of batch i+1 and K3 of batch i are
 * then in flight together. K2 is bound by the tensor / epilogue pipes and moves almost no DRAM bytes, K3 is a pure HBM
 * write stream; with scatter_ctas_per_sm = 1 the TMA-engine scatter (mbev_scatter_forward_stream) shares every SM with
 * K2's persistent CTA. scatter_ctas_per_sm = 0 selects mbev_scatter_forward. `canvas` and `workspace`
 * (mbev_pfn_workspace_bytes) stay single: stage 2 is ordered by pfn_stream, stage 3 by stream. vox_workspace:
 * mbev_voxelize_workspace_bytes. Events come from mbev_event_create (cudaEventDisableTiming). Results are bit-identical
 * to mbev_encode_batch. */
MBEV_API int mbev_event_create(void **event);
MBEV_API int mbev_event_destroy(void *event);
MBEV_API int mbev_encode_batch_pipelined(const float *points_host, float *points_dev,
                                         const int64_t *frame_offsets_host, int batch, const MbevGeometry *geo,
                                         const MbevPfnParams *params, int32_t *cell_table, int32_t *coors,
                                         int32_t *num_points, int32_t *kept_idx, int32_t *pillar_base,
                                         int64_t pillar_capacity, float *feats, float *canvas, void *vox_workspace,
                                         size_t vox_workspace_bytes, void *workspace, size_t workspace_bytes,
                                         int scatter_ctas_per_sm, void *stream, void *prep_stream, void *pfn_stream,
                                         void *ev_ready, void *ev_feats, void *ev_consumed);

/* Launch counter: number of library kernels enqueued by this process since load (for bench.py's
 * `gpu_launches`). Thread-safe. */
MBEV_API int64_t mbev_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MASK_BEV_B200_H_ */
