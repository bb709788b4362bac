// placeholder: tcgen05 path not built yet
#include "ls_internal.cuh"
int lsf_init(ls_handle*, cudaStream_t) { return 1; }
void lsf_destroy(ls_handle*) {}
int lsf_available(const ls_handle*) { return 0; }
int lsf_step(ls_handle* h, int, const ls_step_params*, int, const float*, const float*, const float*, const float*,
             int64_t, int64_t, int64_t, const float*, float*, float*, cudaStream_t) {
  return ls_fail(h, LS_EUNSUPPORTED, "tcgen05 path not built");
}
