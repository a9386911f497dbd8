// fp32 SIMT GEMM for the shapes that do not qualify for the tcgen05 path (skinny K/N, unaligned leading
// dimensions) and as the exact-fp32 engine.  64x64x16 CTA tile, 4x4 register tile per thread, split-K with a
// deterministic second pass when the output has too few tiles to fill 148 SMs (weight gradients: reduction
// over the batch).   Reference arithmetic: np.dot at neograd/autograd/ops/basics.py:194, 206, 207.
#include "tc_common.cuh"

namespace ngb {

constexpr int BM = 64, BN = 64, BK = 16;

// same fused epilogue as the tcgen05 kernel (tc_common.cuh: Epilogue), in exact fp32
__device__ __forceinline__ float simt_act(int act, float alpha, float v) {
  if (act == NGB_RELU) return fmaxf(v, 0.f);
  if (act == NGB_LEAKY_RELU) return v >= 0.f ? v : alpha * v;
  if (act == NGB_SIGMOID) return 1.f / (1.f + expf(-v));
  if (act == NGB_TANH) return tanhf(v);
  return v;
}
__device__ __forceinline__ float simt_act_grad(int act, float alpha, float z) {
  if (act == NGB_RELU) return z >= 0.f ? 1.f : 0.f;
  if (act == NGB_LEAKY_RELU) return z >= 0.f ? 1.f : alpha;
  if (act == NGB_SIGMOID) { const float t = expf(-fabsf(z)), d = 1.f + t; return t / (d * d); }            // = s(1-s), no cancellation
  if (act == NGB_TANH) { const float t = expf(-2.f * fabsf(z)), d = 1.f + t; return 4.f * t / (d * d); }  // = 1 - tanh^2
  return 1.f;
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_simt_kernel(int64_t M, int64_t N, int64_t K, const float* __restrict__ A,
                                                        int64_t lda, const float* __restrict__ B, int64_t ldb,
                                                        float* __restrict__ C, int64_t ldc, int accumulate,
                                                        int64_t k_per_split, float* __restrict__ partials,
                                                        const tc::Epilogue e, float* __restrict__ cs_part) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = min(K, kbeg + k_per_split);
  const int tm = (t / 16) * 4, tn = (t % 16) * 4;
  float acc[4][4] = {};
  float bsum[4] = {};                                    // column sums of B (bias gradient): first row of tiles, row-group 0
  const bool do_colsum = e.colsum != nullptr && blockIdx.y == 0 && tm == 0;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk, mm;
      if (TA) { mm = t % 64; kk = t / 64 + 4 * j; } else { kk = t % 16; mm = t / 16 + 16 * j; }
      const int64_t gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < kend) v = TA ? A[gk * lda + gm] : A[gm * lda + gk];
      As[kk][mm] = v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk, nn;
      if (TB) { kk = t % 16; nn = t / 16 + 16 * j; } else { nn = t % 64; kk = t / 64 + 4 * j; }
      const int64_t gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < N && gk < kend) v = TB ? B[gn * ldb + gk] : B[gk * ldb + gn];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][tm]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tn]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (do_colsum) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bsum[j] += bv[j];
      }
    }
    __syncthreads();
  }

  if (do_colsum) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gn = n0 + tn + j;
      if (gn < N) {
        if (partials) cs_part[(int64_t)blockIdx.z * N + gn] = bsum[j];
        else e.colsum[gn] = e.colsum_accumulate ? e.colsum[gn] + bsum[j] : bsum[j];
      }
    }
  }
  if (partials) {  // split-K: partial tile -> scratch[z][M*N] (dense, ld = N)
    float* P = partials + (int64_t)blockIdx.z * M * N;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t gm = m0 + tm + i, gn = n0 + tn + j;
        if (gm < M && gn < N) P[gm * N + gn] = acc[i][j];
      }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t gm = m0 + tm + i, gn = n0 + tn + j;
        if (gm < M && gn < N) {
          float v = acc[i][j];
          if (e.bias) v += e.bias[gn];
          if (e.mask_z) v *= simt_act_grad(e.act, e.alpha, e.mask_z[gm * e.ldm + gn]);
          if (e.store_c) {
            float* c = C + gm * ldc + gn;
            *c = accumulate ? *c + v : v;
          }
          if (e.C2) e.C2[gm * e.ldc2 + gn] = simt_act(e.act, e.alpha, v);
        }
      }
  }
}

// out[m*ldc+n] (+)= sum_z part[z][m*N+n];  cs_out[n] (+)= sum_z cs_part[z][n]   (one deterministic pass for both)
__global__ void __launch_bounds__(256) gemm_splitk_reduce2_kernel(const float* __restrict__ part, int splits, int64_t M, int64_t N,
                                                                  float* __restrict__ C, int64_t ldc, int accumulate,
                                                                  const float* __restrict__ cs_part, float* __restrict__ cs_out,
                                                                  int cs_accumulate) {
  const int64_t MN = M * N, total = MN + (cs_out ? N : 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    if (i < MN) {
      for (int z = 0; z < splits; ++z) s += part[(int64_t)z * MN + i];
      const int64_t m = i / N, n = i - m * N;
      float* c = C + m * ldc + n;
      *c = accumulate ? *c + s : s;
    } else {
      const int64_t n = i - MN;
      for (int z = 0; z < splits; ++z) s += cs_part[(int64_t)z * N + n];
      cs_out[n] = cs_accumulate ? cs_out[n] + s : s;
    }
  }
}

int launch_splitk_reduce2(const float* part, int splits, int64_t M, int64_t N, float* C, int64_t ldc, int accumulate,
                          const float* cs_part, float* cs_out, int cs_accumulate, cudaStream_t st) {
  gemm_splitk_reduce2_kernel<<<grid_for(M * N + N, 256), 256, 0, st>>>(part, splits, M, N, C, ldc, accumulate, cs_part, cs_out,
                                                                      cs_accumulate);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

// out[m*ldc+n] (+)= sum_z part[z][m*N+n]
__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t M,
                                                                 int64_t N, float* __restrict__ C, int64_t ldc,
                                                                 int accumulate) {
  const int64_t MN = M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < MN; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(int64_t)z * MN + i];
    const int64_t m = i / N, n = i - m * N;
    float* c = C + m * ldc + n;
    *c = accumulate ? *c + s : s;
  }
}

int launch_splitk_reduce(const float* part, int splits, int64_t M, int64_t N, float* C, int64_t ldc, int accumulate,
                         cudaStream_t st) {
  gemm_splitk_reduce_kernel<<<grid_for(M * N, 256), 256, 0, st>>>(part, splits, M, N, C, ldc, accumulate);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

int gemm_simt(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
              int64_t ldb, float* C, int64_t ldc, int accumulate, cudaStream_t st, const tc::Epilogue* epi) {
  tc::Epilogue e;
  if (epi) e = *epi;
  if (!C) e.store_c = 0;
  const bool fused = e.bias || e.mask_z || e.C2;       // needs the complete sums: no split-K
  const int64_t tiles = ceil_div<int64_t>(M, BM) * ceil_div<int64_t>(N, BN);
  int64_t splits = 1;
  if (tiles < kNumSMs && K >= 512 && !fused) {
    splits = ceil_div<int64_t>(2 * kNumSMs, tiles);
    const int64_t max_by_k = K / 128;
    if (splits > max_by_k) splits = max_by_k;
    if (splits > 1) {
      if (ensure_init() != NGB_OK) return NGB_ERR_CUDA;
      const int64_t cap = scratch_bytes() / ((M * N + N) * (int64_t)sizeof(float));
      if (splits > cap) splits = cap;
    }
    if (splits < 1) splits = 1;
  }
  int64_t kps = ceil_div<int64_t>(ceil_div<int64_t>(K, splits), BK) * BK;
  if (K == 0) kps = BK;
  splits = K == 0 ? 1 : ceil_div<int64_t>(K, kps);
  dim3 grid((unsigned)ceil_div<int64_t>(N, BN), (unsigned)ceil_div<int64_t>(M, BM), (unsigned)splits);
  float* part = splits > 1 ? scratch_ptr() : nullptr;
  float* cs_part = part ? part + splits * M * N : nullptr;
#define NGB_SIMT_LAUNCH(TA, TB) \
  gemm_simt_kernel<TA, TB><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kps, part, e, cs_part)
  if (!transA && !transB) NGB_SIMT_LAUNCH(false, false);
  else if (!transA && transB) NGB_SIMT_LAUNCH(false, true);
  else if (transA && !transB) NGB_SIMT_LAUNCH(true, false);
  else NGB_SIMT_LAUNCH(true, true);
#undef NGB_SIMT_LAUNCH
  NGB_LAUNCH_CHECK();
  if (splits > 1) {
    if (e.colsum) return launch_splitk_reduce2(part, (int)splits, M, N, C, ldc, accumulate, cs_part, e.colsum, e.colsum_accumulate, st);
    return launch_splitk_reduce(part, (int)splits, M, N, C, ldc, accumulate, st);
  }
  return NGB_OK;
}

}  // namespace ngb
