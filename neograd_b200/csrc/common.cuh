// Shared host/device helpers for libngb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>

#include "../../include/ngb200.h"

namespace ngb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const std::string& msg);
int fail(int code, const char* fmt, ...);

// Per-device scratch arena (split-K partials, reduction partials). Allocated by ngb_init().
constexpr int kLanes = 4;     // concurrent host-driven streams with private scratch (ngb_set_lane)
float* scratch_ptr();          // arena of the calling thread's current lane
int64_t scratch_bytes();
int current_lane();
int corun_hint();               // ngb_set_corun: leave room on the SMs for another lane's kernels
int ensure_init();

extern std::atomic<int64_t> g_launches;
extern std::atomic<int> g_default_engine;   // NGB_GEMM_AUTO or NGB_GEMM_SIMT (ngb_set_default_engine)
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define NGB_CHECK_ARG(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return ::ngb::fail(NGB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define NGB_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return ::ngb::fail(NGB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                                      \
  } while (0)

#define NGB_LAUNCH_CHECK()                                                                               \
  do {                                                                                                   \
    cudaError_t _e = cudaPeekAtLastError();                                                              \
    if (_e != cudaSuccess)                                                                               \
      return ::ngb::fail(NGB_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),       \
                         __FILE__, __LINE__);                                                            \
    ::ngb::count_launch();                                                                               \
  } while (0)

// out[k] (+)= sum_j part[j*K + k]; deterministic second pass of every split reduction (elementwise.cu).
int launch_sum_partials(const float* part, int splits, int64_t K, float* out, int accumulate, cudaStream_t st);
int launch_sum_partials2(const float* part, int splits, int64_t K1, float* out1, int acc1, int64_t K2, float* out2, int acc2,
                         cudaStream_t st);

inline cudaStream_t as_stream(ngb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Grid size for a grid-stride kernel: enough CTAs to fill the chip a few times, never more than the work.
inline int grid_for(int64_t work_items, int threads, int ctas_per_sm = 8) {
  int64_t need = ceil_div<int64_t>(work_items, threads);
  int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Division by a launch-constant divisor without the ~40-instruction runtime divide (index decomposition in the
// bandwidth-bound kernels was instruction-bound before this).  Exact for every 32-bit unsigned x.
struct FastDiv {
  uint32_t d, m, s1, s2;
  __host__ __device__ FastDiv() : d(1), m(0), s1(0), s2(0) {}
  __host__ __device__ explicit FastDiv(uint32_t div) : d(div) {
    uint32_t l = 0;
    while ((uint64_t(1) << l) < div) ++l;
    m = (uint32_t)(((uint64_t(1) << 32) * ((uint64_t(1) << l) - div)) / div) + 1;
    s1 = l < 1 ? l : 1;
    s2 = l > 0 ? l - 1 : 0;
  }
  __device__ __forceinline__ uint32_t div(uint32_t x) const {
    const uint32_t t = __umulhi(m, x);
    return (t + ((x - t) >> s1)) >> s2;
  }
  __device__ __forceinline__ void divmod(uint32_t x, uint32_t& q, uint32_t& r) const {
    q = div(x);
    r = x - q * d;
  }
};

// ---- device helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0 (and broadcast to all when `broadcast`). smem: >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* smem, bool broadcast = false) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect smem reuse across calls
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = lane < nw ? smem[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) smem[0] = v;
  }
  if (broadcast) {
    __syncthreads();
    v = smem[0];
  }
  return v;
}

// Streaming 128-bit accesses (read-once data: keep it out of L1).
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

}  // namespace ngb
