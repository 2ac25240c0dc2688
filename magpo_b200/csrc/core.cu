// Library-level entry points: version string and CUDA error reporting.
#include <string.h>

#include "common.cuh"

namespace magpo {

static thread_local char g_err[512] = "";

void set_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), file, line);
}

}  // namespace magpo

extern "C" {
const char* magpo_version(void) { return "magpo_b200 0.1 (sm_100a)"; }
const char* magpo_last_cuda_error(void) { return magpo::g_err; }
}
