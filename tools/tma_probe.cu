// Development probe: what does a 4-D TMA box {8 (i0), 4 (i1), 32 (c), 1 (n)} with the 128B/32B-atom swizzle look like in
// shared memory, and does it complete (transaction bytes) when part of the box is out of bounds?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I neograd_b200/csrc tools/tma_probe.cu -L neograd_b200/lib -lngb200 -o tools/bin/tma_probe
#include "tc_common.cuh"
#include <vector>
using namespace ngb;
using namespace ngb::tc;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__global__ void probe(const __grid_constant__ CUtensorMap tm, float* dump, int rank, int bytes, int c0, int c1, int c2, int c3, int* status) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  float* s = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = -777.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    fence_proxy_async();
    mbar_arrive_expect_tx(bar, bytes);
    if (rank == 4) tma_load_4d(smem, &tm, bar, c0, c1, c2, c3);
    else if (rank == 3) tma_load_3d(smem, &tm, bar, c0, c1, c2);
    else tma_load_2d(smem, &tm, bar, c0, c1);
    // bounded wait without trap
    int ok = 0;
    for (int i = 0; i < 2000000; ++i) if (mbar_try_wait(bar, 0)) { ok = 1; break; }
    *status = ok;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) dump[i] = s[i];
}

int main() {
  setenv("NGB_TMAP_FLOAT32", "1", 1);
  const int I0 = 64, I1 = 8, C = 32, N = 2;
  std::vector<float> h(I0 * I1 * C * N);
  for (int n = 0; n < N; ++n) for (int c = 0; c < C; ++c) for (int i1 = 0; i1 < I1; ++i1) for (int i0 = 0; i0 < I0; ++i0)
    h[((n * C + c) * I1 + i1) * I0 + i0] = float(n * 100000 + c * 1000 + i1 * 64 + i0);
  float *d, *dump; int* st;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&dump, 4096); cudaMalloc(&st, 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  struct V { int rank; uint32_t box[4]; int sw; int coords[4]; };
  const V vs[] = {
    {4, {32, 1, 32, 1}, 0, {0, 0, 0, 0}}, {4, {32, 1, 32, 1}, 2, {0, 0, 0, 0}}, {4, {32, 1, 32, 1}, 2, {-1, -1, 0, 0}},
    {4, {32, 1, 32, 1}, 2, {40, 7, 0, 1}}, {4, {32, 1, 32, 1}, 1, {0, 0, 0, 0}}, {3, {32, 1, 32, 1}, 2, {0, 0, 0, 0}},
  };
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Enc enc = (Enc)get_encode_tiled();
  for (const V& v : vs) {
    CUtensorMap tm;
    cuuint64_t dims[4] = {I0, I1, C, N}, str[3] = {I0 * 4, I0 * I1 * 4, I0 * I1 * C * 4};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle sw = v.sw == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : v.sw == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, v.rank, d, dims, str, v.box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int bytes = 4; for (int i = 0; i < v.rank; ++i) bytes *= v.box[i];
    printf("== rank %d box {%u,%u,%u,%u} swizzle %d coords (%d,%d,%d,%d): encode=%d ", v.rank, v.box[0], v.box[1], v.box[2], v.box[3], v.sw,
           v.coords[0], v.coords[1], v.coords[2], v.coords[3], (int)r);
    if (r) { printf("\n"); continue; }
    cudaMemset(st, 0, 4);
    probe<<<1, 128, 8192 + 1024 + 64>>>(tm, dump, v.rank, bytes, v.coords[0], v.coords[1], v.coords[2], v.coords[3], st);
    cudaError_t e = cudaDeviceSynchronize();
    int ok = -1; std::vector<float> o(1024);
    cudaMemcpy(&ok, st, 4, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), dump, 4096, cudaMemcpyDeviceToHost);
    printf("err=%s completed=%d\n", cudaGetErrorString(e), ok);
    if (e != cudaSuccess) break;
    for (int r2 = 0; r2 < 5; ++r2) { printf("  row%d:", r2); for (int k = 0; k < 32; ++k) printf(" %g", o[r2 * 32 + k]); printf("\n"); }
  }
  return 0;
}
