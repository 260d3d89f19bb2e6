// Error reporting, version and device probe of libsir.
#include "sir_common.cuh"

namespace sir {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
}  // namespace sir

extern "C" const char* sir_last_error(void) { return sir::g_err; }
extern "C" int sir_abi_version(void) { return 3; }

extern "C" int sir_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  SIR_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p{};
  SIR_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (p.major != 10) {
    sir::set_error("libsir is built for sm_100a only; device %d is sm_%d%d", dev, p.major, p.minor);
    return SIR_E_DEVICE;
  }
  return SIR_OK;
}
