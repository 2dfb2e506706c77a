// libg2g: version, error string and device check of the C ABI (include/g2g.h).
#include "g2g_common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void g2g_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int g2g_version(void) { return G2G_VERSION; }
extern "C" const char* g2g_last_error(void) { return g_err; }

extern "C" int g2g_check_device(void) {
  int dev = 0, major = 0;
  G2G_CUDA_TRY(cudaGetDevice(&dev));
  G2G_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    g2g_set_error("libg2g is built for sm_100a (B200) only; device %d has compute capability %d.x", dev, major);
    return G2G_ERR_UNSUPPORTED;
  }
  return G2G_OK;
}
