// Library runtime: error reporting, scratch arena, launch accounting.
#include "common.cuh"

#include <stdarg.h>
#include <mutex>

namespace ngb {

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launches{0};

static std::mutex g_mu;
static float* g_scratch = nullptr;
static int64_t g_scratch_bytes = 0;

void set_error(const std::string& msg) { t_last_error = msg; }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_last_error = buf;
  return code;
}

float* scratch_ptr() { return g_scratch; }
int64_t scratch_bytes() { return g_scratch_bytes; }

int ensure_init() {
  if (g_scratch) return NGB_OK;
  return ngb_init(0);
}

}  // namespace ngb

using namespace ngb;

extern "C" int ngb_abi_version(void) { return NGB_ABI_VERSION; }

extern "C" const char* ngb_last_error(void) { return t_last_error.c_str(); }

extern "C" int64_t ngb_launch_count(void) { return g_launches.load(); }

extern "C" int ngb_init(int64_t scratch_bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (scratch_bytes <= 0) scratch_bytes = 64ll << 20;
  if (g_scratch && g_scratch_bytes >= scratch_bytes) return NGB_OK;
  int dev = 0;
  NGB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  NGB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    return fail(NGB_ERR_UNSUPPORTED, "libngb200 is built for sm_100a (B200) only; device %d is sm_%d%d", dev,
                prop.major, prop.minor);
  }
  if (g_scratch) {
    NGB_CUDA(cudaFree(g_scratch));
    g_scratch = nullptr;
  }
  NGB_CUDA(cudaMalloc(&g_scratch, scratch_bytes));
  g_scratch_bytes = scratch_bytes;
  return NGB_OK;
}

extern "C" int ngb_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_scratch) {
    cudaFree(g_scratch);
    g_scratch = nullptr;
    g_scratch_bytes = 0;
  }
  return NGB_OK;
}
