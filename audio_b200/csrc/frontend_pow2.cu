// Register-FFT fast path for power-of-two n_fft (placeholder: not implemented yet, the generic
// kernel serves every size).
#include "common.cuh"

namespace b200a {

size_t pow2_workspace_extra(const b200a_frontend_desc*) { return 0; }

int pow2_prepare(const b200a_frontend_desc*, void*, size_t, cudaStream_t) { return B200A_OK; }

int frontend_run_pow2(const b200a_frontend_desc*, const void*, int, const float*, int64_t, int64_t, int64_t, int64_t,
                      float*, float*, int64_t, cudaStream_t) {
  return B200A_EUNSUPPORTED;
}

}  // namespace b200a
