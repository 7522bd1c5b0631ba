// Library-level entry points: version, error string, device sync.
#include "st_common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void st_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

ST_API int st_version(void) { return 100; }

ST_API const char* st_last_error(void) { return g_err; }

ST_API int st_device_sync(void) {
  ST_CUDA_CALL(cudaDeviceSynchronize());
  return ST_OK;
}
