// Cross-GPU gradient exchange over NVLink / NVSwitch PEER MEMORY (data-parallel training): every rank's flat
// gradient arena lives in symmetric memory that all GPUs of the node can load from, so the gradient "allreduce" is a
// kernel in which each GPU sums the W peer buffers itself (fixed rank order: every replica gets bit-identical sums),
// bracketed by two flag barriers written with system-scope stores straight into the peers' memory.  For the 178 KB of
// LeNet's gradients an NCCL allreduce is pure latency (~35 us inside the step); this exchange is three launches of a few
// microseconds and is captured in the step's CUDA graph like any other kernel.
#include "common.cuh"

namespace ngb {

static __device__ unsigned long long g_xgpu_timeout = 0;

__device__ __forceinline__ uint64_t gtimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// flags[r] -> rank r's flag words (one per peer).  Rank `rank` publishes its epoch to every peer and waits until every
// peer has published the same epoch here.  One block, thread r handles peer r.  The wait is bounded (2 s): a missing
// peer surfaces as an error (ngb_xgpu_check), never as a hung GPU.
__global__ void xgpu_barrier_kernel(uint32_t* const* __restrict__ flags, int rank, int world, uint32_t* epoch) {
  __shared__ uint32_t e_sh;
  if (threadIdx.x == 0) {
    e_sh = *epoch + 1;
    *epoch = e_sh;
  }
  __syncthreads();
  const uint32_t e = e_sh;
  const int r = threadIdx.x;
  if (r < world) {
    __threadfence_system();   // this GPU's gradient writes (earlier kernels) are visible system-wide before the flag is
    uint32_t* remote = flags[r] + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(e) : "memory");
    const uint32_t* mine = flags[rank] + r;
    const uint64_t t0 = gtimer_ns();
    uint32_t v, spins = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - e) >= 0) break;
      if ((++spins & 0xFF) == 0 && gtimer_ns() - t0 > 2000000000ull) {
        atomicCAS(&g_xgpu_timeout, 0ull, (1ull << 63) | ((unsigned long long)rank << 32) | (unsigned long long)r);
        break;
      }
    }
  }
}

// out[i] = sum_r peers[r][i] in rank order (128-bit system-scope loads: peer data must not come from a stale L1 line)
__global__ void __launch_bounds__(256) xgpu_reduce_sum_kernel(const float* const* __restrict__ peers, float* __restrict__ out,
                                                              int64_t n4, int64_t n, int world) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      float4 v;
      asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(peers[r] + 4 * i));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(out + 4 * i) = s;
  }
  // tail (n not a multiple of 4)
  for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) {
      float v;
      asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(peers[r] + i));
      s += v;
    }
    out[i] = s;
  }
}

}  // namespace ngb

using namespace ngb;

extern "C" int ngb_xgpu_barrier(const void* flag_ptrs_dev, int rank, int world, uint32_t* epoch_dev, ngb_stream_t stream) {
  NGB_CHECK_ARG(flag_ptrs_dev && epoch_dev && world >= 1 && world <= 32 && rank >= 0 && rank < world, "ngb_xgpu_barrier: bad arguments");
  xgpu_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<uint32_t* const*>(flag_ptrs_dev), rank, world, epoch_dev);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_xgpu_reduce_sum(const void* peer_ptrs_dev, float* out, int64_t n, int world, ngb_stream_t stream) {
  NGB_CHECK_ARG(peer_ptrs_dev && out && n >= 0 && world >= 1 && world <= 32, "ngb_xgpu_reduce_sum: bad arguments");
  if (n == 0) return NGB_OK;
  NGB_CHECK_ARG(aligned16(out), "ngb_xgpu_reduce_sum: out must be 16-byte aligned");
  const int64_t n4 = n / 4;
  xgpu_reduce_sum_kernel<<<grid_for(n4 > 0 ? n4 : 1, 256, 2), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float* const*>(peer_ptrs_dev), out, n4, n, world);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

// 0 if no cross-GPU barrier timed out since the last call, else NGB_ERR_CUDA (synchronises the device; debug / tests)
extern "C" int ngb_xgpu_check(void) {
  NGB_CUDA(cudaDeviceSynchronize());
  unsigned long long v = 0, z = 0;
  NGB_CUDA(cudaMemcpyFromSymbol(&v, g_xgpu_timeout, sizeof(v)));
  if (v) {
    NGB_CUDA(cudaMemcpyToSymbol(g_xgpu_timeout, &z, sizeof(z)));
    return fail(NGB_ERR_CUDA, "cross-GPU barrier timed out: rank %llu never heard from rank %llu", (v >> 32) & 0x7fffffffull,
                v & 0xffffffffull);
  }
  return NGB_OK;
}
