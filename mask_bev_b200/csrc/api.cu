// C-ABI glue: version probes, launch counter, and the fused batch entries K1 -> K2(eval) -> K3.
#include <algorithm>

#include "common.cuh"

namespace mbev {
std::atomic<int64_t> g_launches{0};
}

using namespace mbev;

extern "C" int mbev_abi_version(void) { return MBEV_ABI_VERSION; }

extern "C" const char *mbev_build_info(void) {
  return "mask_bev_b200 sm_100a nvcc " __DATE__ " " __TIME__;
}

extern "C" const char *mbev_status_string(int status) {
  switch (status) {
    case MBEV_OK: return "ok";
    case MBEV_ERR_BAD_ARG: return "bad argument";
    case MBEV_ERR_UNSUPPORTED: return "unsupported shape or configuration";
    case MBEV_ERR_WORKSPACE: return "workspace too small";
    case MBEV_ERR_NO_DEVICE: return "no CUDA device";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "unknown status";
}

extern "C" int64_t mbev_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

namespace {
struct FusedWs {
  size_t vox_off, vox_bytes, pfn_off, pfn_bytes, total;
};

int fused_ws(const MbevGeometry *geo, const MbevPfnParams *params, int batch, int64_t total_points,
             int64_t pillar_capacity, FusedWs *w) {
  size_t vb = 0, pb = 0;
  int st = mbev_voxelize_workspace_bytes(geo, batch, total_points, &vb);
  if (st) return st;
  st = mbev_pfn_workspace_bytes(params, geo->max_points, pillar_capacity, 0, &pb);
  if (st) return st;
  w->vox_off = 0;
  w->vox_bytes = vb;
  w->pfn_off = align_up(vb);
  w->pfn_bytes = pb;
  w->total = w->pfn_off + align_up(pb);
  return MBEV_OK;
}
}  // namespace

extern "C" int mbev_encode_batch_workspace_bytes(const MbevGeometry *geo, const MbevPfnParams *params, int batch,
                                                 int64_t total_points, int64_t pillar_capacity, size_t *bytes) {
  if (!geo || !params || !bytes) return MBEV_ERR_BAD_ARG;
  FusedWs w;
  const int st = fused_ws(geo, params, batch, total_points, pillar_capacity, &w);
  if (st) return st;
  *bytes = w.total;
  return MBEV_OK;
}

extern "C" int mbev_encode_batch(const float *points, const int64_t *frame_offsets_host, int batch,
                                 const MbevGeometry *geo, const MbevPfnParams *params, int32_t *cell_table,
                                 int32_t *coors, int32_t *num_points, int32_t *kept_idx, int32_t *pillar_base,
                                 int64_t pillar_capacity, float *feats, float *canvas, void *workspace,
                                 size_t workspace_bytes, void *stream) {
  if (!geo || !params || !frame_offsets_host || !workspace || !feats || !canvas) return MBEV_ERR_BAD_ARG;
  if (batch < 1 || batch > MBEV_MAX_BATCH) return MBEV_ERR_BAD_ARG;
  if (geo->grid[2] != 1) return MBEV_ERR_UNSUPPORTED;  // pillars: one cell along z (mask_bev_module.py:62)
  FusedWs w;
  int st = fused_ws(geo, params, batch, frame_offsets_host[batch], pillar_capacity, &w);
  if (st) return st;
  if (workspace_bytes < w.total) return MBEV_ERR_WORKSPACE;
  char *ws = static_cast<char *>(workspace);
  st = mbev_voxelize(points, frame_offsets_host, batch, geo, cell_table, coors, num_points, kept_idx, pillar_base,
                     pillar_capacity, ws + w.vox_off, w.vox_bytes, stream);
  if (st) return st;
  st = mbev_pfn_forward(points, geo->num_feats, kept_idx, num_points, coors, pillar_base + batch, pillar_capacity,
                        geo->max_points, params, feats, ws + w.pfn_off, w.pfn_bytes, stream);
  if (st) return st;
  return mbev_scatter_forward(feats, cell_table, batch, params->units[params->num_layers - 1], geo->grid[1],
                              geo->grid[0], canvas, stream);
}

extern "C" int mbev_encode_batch_host(const float *points_host, float *points_dev, const int64_t *frame_offsets_host,
                                      int batch, const MbevGeometry *geo, const MbevPfnParams *params,
                                      int32_t *cell_table, int32_t *coors, int32_t *num_points, int32_t *kept_idx,
                                      int32_t *pillar_base, int64_t pillar_capacity, float *feats, float *canvas,
                                      void *workspace, size_t workspace_bytes, void *stream) {
  if (!geo || !frame_offsets_host || batch < 1 || batch > MBEV_MAX_BATCH) return MBEV_ERR_BAD_ARG;
  const int64_t total = frame_offsets_host[batch];
  if (total > 0) {
    if (!points_host || !points_dev) return MBEV_ERR_BAD_ARG;
    MBEV_CUDA(cudaMemcpyAsync(points_dev, points_host, sizeof(float) * static_cast<size_t>(total) * geo->num_feats,
                              cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  }
  return mbev_encode_batch(points_dev, frame_offsets_host, batch, geo, params, cell_table, coors, num_points,
                           kept_idx, pillar_base, pillar_capacity, feats, canvas, workspace, workspace_bytes, stream);
}

extern "C" int mbev_event_create(void **event) {
  if (!event) return MBEV_ERR_BAD_ARG;
  cudaEvent_t e = nullptr;
  MBEV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *event = e;
  return MBEV_OK;
}

extern "C" int mbev_event_destroy(void *event) {
  if (!event) return MBEV_OK;
  MBEV_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
  return MBEV_OK;
}

// ---- three-stage pipeline for a stream of batches ---------------------------------------------------------------
//   prep_stream : [H2D +] K1 of batch i+2      (latency / L2-bound integer work, short kernels)
//   pfn_stream  : K2 of batch i+1              (tensor / epilogue pipes, ~no DRAM traffic, one 576-thread CTA per SM)
//   stream      : K3 of batch i                (pure HBM write stream; the TMA-engine form needs one 128-thread CTA per SM)
// K2 and K3 sit on different rooflines and, with scatter_ctas_per_sm = 1, fit on the same SM at the same time.
extern "C" int mbev_encode_batch_pipelined(const float *points_host, float *points_dev,
                                           const int64_t *frame_offsets_host, int batch, const MbevGeometry *geo,
                                           const MbevPfnParams *params, int32_t *cell_table, int32_t *coors,
                                           int32_t *num_points, int32_t *kept_idx, int32_t *pillar_base,
                                           int64_t pillar_capacity, float *feats, float *canvas, void *vox_workspace,
                                           size_t vox_workspace_bytes, void *workspace, size_t workspace_bytes,
                                           int scatter_ctas_per_sm, void *stream, void *prep_stream, void *pfn_stream,
                                           void *ev_ready, void *ev_feats, void *ev_consumed) {
  if (!geo || !params || !frame_offsets_host || batch < 1 || batch > MBEV_MAX_BATCH) return MBEV_ERR_BAD_ARG;
  if (!prep_stream || !pfn_stream || !ev_ready || !ev_feats || !ev_consumed || prep_stream == stream ||
      pfn_stream == stream || prep_stream == pfn_stream || !feats || !canvas || !workspace || scatter_ctas_per_sm < 0)
    return MBEV_ERR_BAD_ARG;
  if (geo->grid[2] != 1) return MBEV_ERR_UNSUPPORTED;
  cudaStream_t ps = static_cast<cudaStream_t>(prep_stream), fs = static_cast<cudaStream_t>(pfn_stream),
               ms = static_cast<cudaStream_t>(stream);
  cudaEvent_t e_ready = static_cast<cudaEvent_t>(ev_ready), e_feats = static_cast<cudaEvent_t>(ev_feats),
              e_consumed = static_cast<cudaEvent_t>(ev_consumed);
  const int64_t total = frame_offsets_host[batch];
  const int c_out = params->units[params->num_layers - 1], ny = geo->grid[1], nx = geo->grid[0];
  size_t pb = 0;
  int st = mbev_pfn_workspace_bytes(params, geo->max_points, pillar_capacity, 0, &pb);
  if (st) return st;
  if (workspace_bytes < pb) return MBEV_ERR_WORKSPACE;
  // stage 1 (prep stream): the batch that used THIS buffer set two calls ago has left K3 (a never-recorded event
  // counts as complete)
  MBEV_CUDA(cudaStreamWaitEvent(ps, e_consumed, 0));
  if (points_host && total > 0) {
    if (!points_dev) return MBEV_ERR_BAD_ARG;
    MBEV_CUDA(cudaMemcpyAsync(points_dev, points_host, sizeof(float) * static_cast<size_t>(total) * geo->num_feats,
                              cudaMemcpyHostToDevice, ps));
  }
  st = mbev_voxelize(points_dev, frame_offsets_host, batch, geo, cell_table, coors, num_points, kept_idx, pillar_base,
                     pillar_capacity, vox_workspace, vox_workspace_bytes, prep_stream);
  MBEV_CUDA(cudaEventRecord(e_ready, ps));
  // stage 2 (pfn stream): K2 into this set's feats
  MBEV_CUDA(cudaStreamWaitEvent(fs, e_ready, 0));
  if (!st)
    st = mbev_pfn_forward(points_dev, geo->num_feats, kept_idx, num_points, coors, pillar_base + batch, pillar_capacity,
                          geo->max_points, params, feats, workspace, workspace_bytes, pfn_stream);
  MBEV_CUDA(cudaEventRecord(e_feats, fs));
  // stage 3 (the caller's stream): K3; the canvas is ordered by `stream` like any other output
  MBEV_CUDA(cudaStreamWaitEvent(ms, e_feats, 0));
  if (!st) {
    if (scatter_ctas_per_sm > 0 && mbev_scatter_stream_supported(c_out, ny, nx, canvas))
      st = mbev_scatter_forward_stream(feats, cell_table, batch, c_out, ny, nx, canvas, scatter_ctas_per_sm, stream);
    else
      st = mbev_scatter_forward(feats, cell_table, batch, c_out, ny, nx, canvas, stream);
  }
  MBEV_CUDA(cudaEventRecord(e_consumed, ms));  // also on error, so that the prep stream never waits forever
  return st;
}
