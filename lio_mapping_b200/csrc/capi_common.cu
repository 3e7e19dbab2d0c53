// lio_mapping_b200 — library-wide C-ABI helpers (error text, device probe).
#include "common.cuh"
#include <cstring>

static thread_local char g_err[512] = "";

void lio_set_last_error(const char *file, int line, const char *msg) {
  snprintf(g_err, sizeof(g_err), "%s:%d: %s", file, line, msg ? msg : "");
}

extern "C" const char *lio_last_error(void) { return g_err; }
extern "C" int lio_version(void) { return 100; }
extern "C" int lio_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    lio_set_last_error(__FILE__, __LINE__, cudaGetErrorString(e));
    cudaGetLastError();
    return 0;
  }
  return n;
}
