// mqdet_b200 — error plumbing shared by all C-ABI entry points.
#include "common.cuh"
#include "../../include/mqdet_b200.h"
#include <stdarg.h>

namespace mqdet {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return MQDET_ERR_CUDA;
  }
  return MQDET_OK;
}

}  // namespace mqdet

extern "C" const char* mqdet_last_error(void) { return mqdet::g_err; }
extern "C" int mqdet_version(void) { return 100; }
