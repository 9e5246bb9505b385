// K2' placeholder while the forward path is validated on hardware; replaced by the real backward.
#include "common.cuh"
extern "C" int mbev_pfn_backward_workspace_bytes(const MbevPfnParams *, int, int64_t, size_t *bytes) {
  if (bytes) *bytes = 16;
  return MBEV_OK;
}
extern "C" int mbev_pfn_backward(const float *, int, const int32_t *, const int32_t *, const int32_t *, const int32_t *,
                                 int64_t, int, const MbevPfnParams *, const float *const *, const float *, const float *,
                                 float, int, const float *, float *const *, float *const *, float *const *, void *,
                                 size_t, void *) {
  return MBEV_ERR_UNSUPPORTED;
}
