// Development probe for the tcgen05 / TMA primitives in neograd_b200/csrc/tc_common.cuh: one CTA, one 128 x 64 x 32 tile.
// Dumps the shared-memory image TMA produced and the TMEM accumulator so descriptor / layout mistakes are visible.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I neograd_b200/csrc tools/tc_probe.cu \
//        -L neograd_b200/lib -lngb200 -o gpurun_out/tc_probe   (run with LD_LIBRARY_PATH=neograd_b200/lib)
#include "tc_common.cuh"
#include <vector>
#include <cmath>
using namespace ngb;
using namespace ngb::tc;

constexpr int M = 128, N = 64, K = 32;

struct DescParams { uint32_t lbo, sbo, layout, kstep; };

template <int A_MN, int B_MN>
__global__ void probe(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, float* dump_a,
                      float* dump_b, float* out, uint32_t* info, DescParams dp) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;
  uint8_t* sb = smem + M * K * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + N * K * 4);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    info[0] = tmem;
    info[1] = smem_u32(sa);
    mbar_arrive_expect_tx(&bars[0], (M + N) * K * 4);
    if (A_MN) {
      for (int c = 0; c < 4; ++c) tma_load_2d(sa + c * 4096, &ta, &bars[0], c * 32, 0);
    } else {
      tma_load_2d(sa, &ta, &bars[0], 0, 0);
    }
    if (B_MN) {
      tma_load_2d(sb, &tb, &bars[0], 0, 0);
      tma_load_2d(sb + 4096, &tb, &bars[0], 32, 0);
    } else {
      tma_load_2d(sb, &tb, &bars[0], 0, 0);
    }
  }
  mbar_wait(&bars[0], 0);
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) dump_a[i] = reinterpret_cast<float*>(sa)[i];
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) dump_b[i] = reinterpret_cast<float*>(sb)[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_tf32(M, N, A_MN, B_MN);
    const uint64_t mn = make_smem_desc_base(dp.lbo, dp.sbo, dp.layout);
    const uint64_t ad = A_MN ? mn : make_smem_desc_base(16, 1024);
    const uint64_t bd = B_MN ? mn : make_smem_desc_base(16, 1024);
    info[2] = idesc;
    for (int k = 0; k < 4; ++k) {
      uint64_t a = smem_desc(ad, smem_u32(sa) + k * (A_MN ? dp.kstep : 32));
      uint64_t b = smem_desc(bd, smem_u32(sb) + k * (B_MN ? dp.kstep : 32));
      if (k == 0) { info[4] = (uint32_t)a; info[5] = (uint32_t)(a >> 32); info[6] = (uint32_t)b; info[7] = (uint32_t)(b >> 32); }
      umma_tf32(tmem, a, b, idesc, k > 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

template <int A_MN, int B_MN>
int run(DescParams dp, bool atom32b) {
  std::vector<float> A(M * K), B(N * K), C(M * N);
  auto aval = [](int m, int k) { return float((m * 7 + k * 3) % 11 - 5); };
  auto bval = [](int n, int k) { return float((n * 5 + k * 2) % 9 - 4); };
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { if (A_MN) A[k * M + m] = aval(m, k); else A[m * K + k] = aval(m, k); }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { if (B_MN) B[k * N + n] = bval(n, k); else B[n * K + k] = bval(n, k); }
  float *dA, *dB, *dumpa, *dumpb, *dout; uint32_t* dinfo;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dumpa, A.size() * 4));
  CK(cudaMalloc(&dumpb, B.size() * 4)); CK(cudaMalloc(&dout, C.size() * 4)); CK(cudaMalloc(&dinfo, 64));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, C.size() * 4));
  CUtensorMap ta, tb;
  uint64_t dims[2], str[1]; uint32_t box[2];
  if (A_MN) { dims[0] = M; dims[1] = K; str[0] = M * 4; box[0] = 32; box[1] = 32; }
  else { dims[0] = K; dims[1] = M; str[0] = K * 4; box[0] = 32; box[1] = 128; }
  if (make_tmap_f32(&ta, dA, 2, dims, str, box, A_MN && atom32b)) { printf("tmap a failed: %s\n", ngb_last_error()); return 1; }
  if (B_MN) { dims[0] = N; dims[1] = K; str[0] = N * 4; box[0] = 32; box[1] = 32; }
  else { dims[0] = K; dims[1] = N; str[0] = K * 4; box[0] = 32; box[1] = 64; }
  if (make_tmap_f32(&tb, dB, 2, dims, str, box, B_MN && atom32b)) { printf("tmap b failed: %s\n", ngb_last_error()); return 1; }
  const int smem = (M + N) * K * 4 + 1024 + 256;
  CK(cudaFuncSetAttribute(probe<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe<A_MN, B_MN><<<1, 128, smem>>>(ta, tb, dumpa, dumpb, dout, dinfo, dp);
  CK(cudaDeviceSynchronize());
  std::vector<float> db(B.size());
  uint32_t info[16];
  CK(cudaMemcpy(db.data(), dumpb, db.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(C.data(), dout, C.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(info, dinfo, 64, cudaMemcpyDeviceToHost));
  int bad = 0, nz = 0; double maxerr = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0; for (int k = 0; k < K; ++k) ref += double(aval(m, k)) * bval(n, k);
    double err = fabs(ref - C[m * N + n]);
    if (!(err <= 1e-3)) ++bad;
    if (err > maxerr) maxerr = err;
    if (C[m * N + n] != 0) ++nz;
  }
  printf("A_MN=%d B_MN=%d atom32b=%d lbo=%u sbo=%u layout=%u kstep=%u : D mismatches %d / %d (max err %g, nonzero %d) %s\n", A_MN, B_MN,
         (int)atom32b, dp.lbo, dp.sbo, dp.layout, dp.kstep, bad, M * N, maxerr, nz, bad ? "" : "  <== OK");
  if (B_MN && bad && dp.lbo == 4096 && dp.sbo == 512) {
    printf("   smem B rows 0..3 (first 32 floats each):\n");
    for (int r = 0; r < 4; ++r) { printf("   r%d:", r); for (int e = 0; e < 32; ++e) printf(" %g", db[r * 32 + e]); printf("\n"); }
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dumpa); cudaFree(dumpb); cudaFree(dout); cudaFree(dinfo);
  return 0;
}

int main() {
  run<0, 0>({16, 1024, 2, 32}, false);
  const DescParams cands[] = {{4096, 512, 1, 1024}, {512, 4096, 1, 1024}, {4096, 1024, 1, 1024}, {4096, 256, 1, 1024},
                              {4096, 512, 1, 512}, {4096, 1024, 2, 1024}};
  for (const auto& c : cands) {
    run<0, 1>(c, true);
    if (c.layout == 2) run<0, 1>(c, false);
  }
  for (const auto& c : cands) run<1, 0>(c, true);
  run<1, 1>(cands[0], true);
  return 0;
}
