// Error reporting, version and device probe of the hsv C-ABI.
#include <stdarg.h>
#include <string.h>
#include "hsv_common.cuh"

namespace hsv {
int g_pdl = 1;
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace hsv

extern "C" int hsv_version(void) { return HSV_VERSION; }

// bring-up switch (not part of the drop-in contract): 0 disables programmatic dependent launch
extern "C" int hsv_set_pdl(int on) {
  hsv::g_pdl = on ? 1 : 0;
  return HSV_OK;
}

extern "C" const char *hsv_last_error(void) { return hsv::g_err; }

extern "C" int hsv_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

extern "C" int64_t hsv_blk16_rows(int64_t L) { return hsv::blk16_rows(L); }
