// Library runtime: error reporting, scratch arena, launch accounting.
#include "common.cuh"

#include <stdarg.h>
#include <mutex>

namespace ngb {

int gemm_set_host_word(unsigned long long* dev_alias);
int conv_set_host_word(unsigned long long* dev_alias);
static unsigned long long* g_pipe_host = nullptr;       // host-mapped word the tcgen05 kernels write when a pipeline wait gives up

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launches{0};

static std::mutex g_mu;
// One scratch arena per LANE.  A lane is a stream the host runs concurrently with others (ngb_set_lane): kernels of
// different lanes may overlap on the GPU, so they must not share split-reduction partials or staged operand copies.
// Lane 0 is allocated by ngb_init(); lanes 1..kLanes-1 on first use (outside graph capture: see ngb_set_lane).
static float* g_scratch[kLanes] = {nullptr, nullptr, nullptr, nullptr};
static int64_t g_scratch_bytes = 0;
static thread_local int t_lane = 0;
static thread_local int t_corun = 0;

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

float* scratch_ptr() { return g_scratch[t_lane]; }
int64_t scratch_bytes() { return g_scratch_bytes; }
int current_lane() { return t_lane; }
int corun_hint() { return t_corun; }

int ensure_init() {
  if (g_scratch[t_lane]) return NGB_OK;
  if (t_lane != 0) return fail(NGB_ERR_SCRATCH, "lane %d has no scratch arena (ngb_set_lane allocates it)", t_lane);
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
  if (g_scratch[0] && g_scratch_bytes >= scratch_bytes) return NGB_OK;
  int dev = 0;
  NGB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  NGB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    return fail(NGB_ERR_UNSUPPORTED, "libngb200 is built for sm_100a (B200) only; device %d is sm_%d%d", dev,
                prop.major, prop.minor);
  }
  for (int l = 0; l < kLanes; ++l) {
    const bool had = g_scratch[l] != nullptr;
    if (had) {
      NGB_CUDA(cudaFree(g_scratch[l]));
      g_scratch[l] = nullptr;
    }
    if (l == 0 || had) NGB_CUDA(cudaMalloc(&g_scratch[l], scratch_bytes));
  }
  g_scratch_bytes = scratch_bytes;
  if (!g_pipe_host) {
    unsigned long long* dev_alias = nullptr;
    NGB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_pipe_host), sizeof(unsigned long long), cudaHostAllocMapped));
    *g_pipe_host = 0;
    NGB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dev_alias), g_pipe_host, 0));
    if (gemm_set_host_word(dev_alias) || conv_set_host_word(dev_alias))
      return fail(NGB_ERR_CUDA, "ngb_init: could not publish the pipeline status word");
  }
  return NGB_OK;
}

// Selects the lane (0..3) whose scratch arena / conv workspace the calling thread's subsequent launches use; returns the
// previous lane, or -1 on error.  The first selection of a lane allocates its arena (cudaMalloc: do it before capturing).
extern "C" int ngb_set_lane(int lane) {
  if (lane < 0 || lane >= kLanes) {
    fail(NGB_ERR_INVALID, "ngb_set_lane: lane must be in [0, %d)", kLanes);
    return -1;
  }
  const int prev = t_lane;
  if (!g_scratch[lane]) {
    if (!g_scratch[0] && ngb_init(0) != NGB_OK) return -1;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_scratch[lane] && cudaMalloc(&g_scratch[lane], g_scratch_bytes) != cudaSuccess) {
      fail(NGB_ERR_CUDA, "ngb_set_lane: cudaMalloc of the lane's scratch arena failed");
      return -1;
    }
  }
  t_lane = lane;
  return prev;
}

namespace ngb {
unsigned long long gemm_take_timeout();
unsigned long long conv_take_timeout();
}

extern "C" int ngb_xgpu_status(void);

// Non-blocking health check of everything that can fail asynchronously on the device: a tcgen05 pipeline wait that gave
// up (the kernel left early: its output is incomplete) and the data-parallel peer exchange (ngb_xgpu_status).  Reads
// host-mapped words only -- cheap enough to poll every few training steps; NGB_ERR_CUDA + ngb_last_error() on failure.
extern "C" int ngb_runtime_status(void) {
  if (g_pipe_host) {
    const unsigned long long v = *reinterpret_cast<volatile unsigned long long*>(g_pipe_host);
    if (v)
      return fail(NGB_ERR_CUDA, "a tcgen05 pipeline wait timed out (tag %llu, block %llu, thread %llu): the kernel left early and its "
                  "output is incomplete", (v >> 48) & 0x7fff, (v >> 16) & 0xffffffffull, v & 0xffff);
  }
  return ngb_xgpu_status();
}

// Hint for persistent kernels: 1 = the calling thread's next launches are expected to run BESIDE another lane's kernels
// (leave part of every SM free so the other lane's kernels can become resident), 0 (default) = they have the GPU to
// themselves.  Returns the previous value.
extern "C" int ngb_set_corun(int corun) {
  const int prev = t_corun;
  t_corun = corun ? 1 : 0;
  return prev;
}

// A producer / MMA / epilogue warp that waits on an mbarrier for more than a second records where and leaves the kernel
// instead of hanging the GPU.  Returns 0 if no tcgen05 pipeline timed out since the last call, else NGB_ERR_CUDA with the
// location in ngb_last_error().  Synchronises the device (debug / test use).
extern "C" int ngb_debug_check_pipelines(void) {
  NGB_CUDA(cudaDeviceSynchronize());
  const unsigned long long a = gemm_take_timeout(), b = conv_take_timeout();
  if (a | b) {
    if (g_pipe_host) *g_pipe_host = 0;
    const unsigned long long v = a ? a : b;
    return fail(NGB_ERR_CUDA, "tcgen05 pipeline timeout in %s kernel: wait tag %llu, block %llu, thread %llu", a ? "gemm" : "conv",
                (v >> 48) & 0x7fff, (v >> 16) & 0xffffffffull, v & 0xffff);
  }
  return NGB_OK;
}

extern "C" int ngb_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (int l = 0; l < kLanes; ++l) {
    if (g_scratch[l]) cudaFree(g_scratch[l]);
    g_scratch[l] = nullptr;
  }
  g_scratch_bytes = 0;
  return NGB_OK;
}
