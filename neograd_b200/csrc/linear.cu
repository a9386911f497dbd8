// Linear layers as ONE kernel per pass (neograd/nn/layers/misc.py:63 `dot(inputs, W) + b`, followed by an activation
// layer, nn/activations.py:20 / 205 / 55):
//   forward   z = x.W + b   and   a = act(z)          (both written by the GEMM's epilogue)
//   data-grad dx (+)= (ug . W^T) * act'(z_prev)       (basics.py:206 fused with the previous activation's backward, activations.py:31 / 216)
//   wgt-grad  dW (+)= x^T . ug ,  db (+)= column sums of ug     (basics.py:207 + the bias `unbroadcast`, autograd/utils.py:75)
//   head      z = x.W + b ; loss = SoftmaxCE(z, t) ; dz = (softmax(z) - t) / rows    (nn/loss.py:153-170)
// Three engines behind each entry point:
//   * tcgen05 TF32 GEMM with the fused epilogue (gemm_tcgen05.cu) for tensor-core sized layers,
//   * the fp32 SIMT GEMM with the same epilogue (gemm_simt.cu) for exact-fp32 mode and odd shapes,
//   * the small-N kernels below for layers with at most 32 outputs (classifier heads, the GAN discriminator's 256 -> 1):
//     a 64-wide GEMM tile would waste 84-98 % of its columns there, and the work is a few MFLOP -- these are latency
//     kernels: one pass over x, W resident in shared memory, deterministic fixed-order reductions (last block sums the
//     per-block partials, no atomics on data).
#include "tc_common.cuh"

namespace ngb {

int gemm_simt(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
              int64_t ldb, float* C, int64_t ldc, int accumulate, cudaStream_t st, const tc::Epilogue* epi);

namespace {

__device__ __forceinline__ float lin_act(int act, float alpha, float v) {
  if (act == NGB_RELU) return fmaxf(v, 0.f);
  if (act == NGB_LEAKY_RELU) return v >= 0.f ? v : alpha * v;
  if (act == NGB_SIGMOID) return 1.f / (1.f + expf(-v));
  if (act == NGB_TANH) return tanhf(v);
  return v;
}
__device__ __forceinline__ float lin_act_grad(int act, float alpha, float z) {
  if (act == NGB_RELU) return z >= 0.f ? 1.f : 0.f;
  if (act == NGB_LEAKY_RELU) return z >= 0.f ? 1.f : alpha;
  if (act == NGB_SIGMOID) { const float t = expf(-fabsf(z)), d = 1.f + t; return t / (d * d); }            // = s(1-s), no cancellation
  if (act == NGB_TANH) { const float t = expf(-2.f * fabsf(z)), d = 1.f + t; return 4.f * t / (d * d); }  // = 1 - tanh^2
  return 1.f;
}

// Per-lane slots for the "last block reduces" pattern: [0] = blocks-done counter (left at 0 by the kernel that used it).
__device__ unsigned int g_done_counter[kLanes * 8];

// W (K x N, row-major) -> shared memory [K][NP] zero-padded to NT columns; NP odd => rows of consecutive k sit on
// different banks for any fixed column.
template <int NT>
__device__ __forceinline__ void stage_w(float* Ws, const float* __restrict__ W, int K, int N) {
  constexpr int NP = NT | 1;
  for (int i = threadIdx.x; i < K * NP; i += blockDim.x) {
    const int k = i / NP, n = i - k * NP;
    Ws[i] = n < N ? W[(int64_t)k * N + n] : 0.f;
  }
}

struct HeadArgs {
  const float* t;       // targets [M, N]
  float inv_n, eps;
  float* dz;            // (softmax(z) - t) * inv_n
  float* loss;          // scalar
  float* partials;      // [gridDim.x]
  unsigned int* done;
};

// One warp per row: lanes split K (coalesced read of the row), NT accumulators per lane, butterfly-reduced so that every
// lane ends with all N outputs.  HEAD: softmax cross-entropy of the row against t, its gradient, and the loss reduction.
template <int NT, bool HEAD>
__global__ void __launch_bounds__(256) linear_smalln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                                const float* __restrict__ b, float* __restrict__ z,
                                                                float* __restrict__ a, int act, float alpha, int64_t M, int N,
                                                                int K, HeadArgs h) {
  extern __shared__ float smem_lin[];
  constexpr int NP = NT | 1;
  float* Ws = smem_lin;
  __shared__ float red[32];
  __shared__ bool is_last;
  stage_w<NT>(Ws, W, K, N);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  float bias_mine = (b != nullptr && lane < N) ? b[lane] : 0.f;
  float bias_all[NT];                                    // (HEAD only; dead code otherwise)
#pragma unroll
  for (int n = 0; n < NT; ++n) bias_all[n] = (HEAD && b != nullptr && n < N) ? b[n] : 0.f;
  float loss_acc = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * nwarp + warp; row < M; row += (int64_t)gridDim.x * nwarp) {
    const float* xr = x + row * K;
    float acc[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float xv = xr[k];
      const float* wr = Ws + k * NP;
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[n] = fmaf(xv, wr[n], acc[n]);
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n] = warp_sum(acc[n]);
    // lane n keeps output n
    float mine = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) mine = lane == n ? acc[n] : mine;
    mine += bias_mine;
    if (lane < N) {
      if (z) z[row * N + lane] = mine;
      if (a) a[row * N + lane] = lin_act(act, alpha, mine);
    }
    if (HEAD) {
      // softmax over the N logits (nn/activations.py:161-174), loss.py:155 and :170.  Every lane holds all logits, so each
      // evaluates max and sum SEQUENTIALLY over the classes -- the summation order of softmax_ce_{fwd,bwd}_kernel
      // (loss.cu), which makes this head's gradient bit-identical to the unfused kernels'.
      float mx = -INFINITY;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        if (n < N) mx = fmaxf(mx, acc[n] + bias_all[n]);
      float sum = 0.f;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        if (n < N) sum += expf(acc[n] + bias_all[n] - mx);
      const float inv = 1.f / sum;
      if (lane < N) {
        const float p = expf(mine - mx) * inv;
        const float tv = h.t[row * N + lane];
        if (tv != 0.f) loss_acc += tv * logf(p + h.eps);
        h.dz[row * N + lane] = h.inv_n * (p - tv);
      }
    }
  }
  if (HEAD) {
    const float s = block_sum(loss_acc, red);
    if (threadIdx.x == 0) {
      h.partials[blockIdx.x] = s;
      __threadfence();
      is_last = atomicAdd(h.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {      // fixed-shape sum of the per-block partials (<= 1024 of them, four per thread): deterministic
      __threadfence();
      const int nbk = (int)gridDim.x, t = threadIdx.x;
      const volatile float* pp = h.partials;     // written by other blocks of this launch
      const float t0 = t < nbk ? pp[t] : 0.f, t1 = t + 256 < nbk ? pp[t + 256] : 0.f;
      const float t2 = t + 512 < nbk ? pp[t + 512] : 0.f, t3 = t + 768 < nbk ? pp[t + 768] : 0.f;
      const float tot = block_sum((t0 + t1) + (t2 + t3), red);
      if (t == 0) {
        *h.loss = -h.inv_n * tot;      // loss.py:156: (-1/num_examples) * entropy
        *h.done = 0;
      }
    }
  }
}

// dx[m][k] (+)= (sum_n ug[m][n] W[k][n]) * act'(zprev[m][k]): one thread per element, k fastest (coalesced), W in smem.
template <int NT>
__global__ void __launch_bounds__(256) linear_smalln_dgrad_kernel(const float* __restrict__ ug, const float* __restrict__ W,
                                                                  const float* __restrict__ zprev, float* __restrict__ dx,
                                                                  int act, float alpha, int64_t M, int N, int K,
                                                                  int accumulate, FastDiv dK) {
  extern __shared__ float smem_lin[];
  constexpr int NP = NT | 1;
  float* Ws = smem_lin;
  stage_w<NT>(Ws, W, K, N);
  __syncthreads();
  const int64_t total = M * K;
  const bool small = total < (int64_t(1) << 32);       // (the usual case: one multiply-shift instead of a 64-bit divide per element)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = small ? (int64_t)dK.div((uint32_t)i) : i / K;
    const int k = (int)(i - m * K);
    const float* ur = ug + m * N;
    const float* wr = Ws + k * NP;
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) s = fmaf(n < N ? ur[n] : 0.f, wr[n], s);
    if (zprev) s *= lin_act_grad(act, alpha, zprev[i]);
    dx[i] = accumulate ? dx[i] + s : s;
  }
}

// dW[k][n] = sum_m x[m][k] ug[m][n], db[n] = sum_m ug[m][n]: a reduction over the batch with only K x N outputs.  Block = KB
// adjacent k (one 16 / 32-byte piece of every x row, so the strided walk down the batch still uses whole sectors), all N
// outputs of each; the last block sums ug itself (the bias gradient).  Threads stride over the rows with KB x NT register
// accumulators, then one fixed-shape tree (warp shuffles, then the 8 warp results in order) gives the block's outputs: no
// partial buffers, no second pass, bit-reproducible.
template <int NT, int KB>
__global__ void __launch_bounds__(256) linear_smalln_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ ug,
                                                                  float* __restrict__ dW, float* __restrict__ db, int64_t M, int N,
                                                                  int K, int acc_w, int acc_b) {
  __shared__ float red[8][KB * NT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_bias = blockIdx.x == gridDim.x - 1;
  const int k0 = blockIdx.x * KB;
  const bool xvec = (KB % 4 == 0) && (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  float acc[KB][NT];
#pragma unroll
  for (int j = 0; j < KB; ++j)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[j][n] = 0.f;
#pragma unroll 4
  for (int64_t m = tid; m < M; m += 256) {
    float u[NT];
    const float* ur = ug + m * N;
#pragma unroll
    for (int n = 0; n < NT; ++n) u[n] = n < N ? ur[n] : 0.f;
    if (is_bias) {
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[0][n] += u[n];
    } else {
      float xv[KB];
      const float* xr = x + m * K + k0;
      if (xvec && k0 + KB <= K) {
#pragma unroll
        for (int j = 0; j < KB; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(xr + j);
          xv[j] = t.x; xv[j + 1] = t.y; xv[j + 2] = t.z; xv[j + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < KB; ++j) xv[j] = k0 + j < K ? xr[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < KB; ++j)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[j][n] = fmaf(xv[j], u[n], acc[j][n]);
    }
  }
#pragma unroll
  for (int j = 0; j < KB; ++j)
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const float w = warp_sum(acc[j][n]);
      if (lane == 0) red[warp][j * NT + n] = w;
    }
  __syncthreads();
  if (tid < KB * NT) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][tid];
    const int j = tid / NT, n = tid - j * NT;
    if (n < N) {
      if (is_bias) {
        if (j == 0 && db) db[n] = acc_b ? db[n] + s : s;
      } else if (k0 + j < K) {
        float* d = dW + (int64_t)(k0 + j) * N + n;
        *d = acc_w ? *d + s : s;
      }
    }
  }
}

// The same gradients for n_in <= 256 and a batch worth spreading over the chip: a block owns a SLAB of rows and a thread one
// column k of 256 / K rows in flight, so consecutive threads read consecutive words of x (rows are K contiguous floats) and
// the ug row is a broadcast; sums over the rows in flight in shared memory (fixed order), one row of K*N + N partials per
// block, summed across blocks in a fixed order by the second pass (launch_sum_partials2).  The kernel above walks the whole
// batch with K / KB + 1 blocks (33 for the GAN's 256 -> 1 head: 18 - 25 us for an 8 MB read); this one takes 4 + 3 us.
template <int NT>
__global__ void __launch_bounds__(256) linear_smalln_wgrad_slab_kernel(const float* __restrict__ x, const float* __restrict__ ug,
                                                                       float* __restrict__ part, int64_t M, int N, int K,
                                                                       int rows_per_block, int R) {
  __shared__ float red[256 * NT];
  __shared__ float redb[256];            // [R][NT] bias partials (R * NT <= 256 is checked by the host)
  const int tid = threadIdx.x;
  const int r = tid / K, k = tid - r * K;
  const bool live = r < R;
  float acc[NT], accb[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) acc[n] = accb[n] = 0.f;
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t m1 = m0 + rows_per_block < M ? m0 + rows_per_block : M;
  if (live) {
#pragma unroll 4
    for (int64_t m = m0 + r; m < m1; m += R) {
      const float xv = x[m * K + k];
      const float* ur = ug + m * N;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const float u = n < N ? ur[n] : 0.f;
        acc[n] = fmaf(xv, u, acc[n]);
        accb[n] += u;                     // (used by the k == 0 thread of the row only)
      }
    }
  }
#pragma unroll
  for (int n = 0; n < NT; ++n) red[tid * NT + n] = live ? acc[n] : 0.f;
  if (live && k == 0) {
#pragma unroll
    for (int n = 0; n < NT; ++n) redb[r * NT + n] = accb[n];
  }
  __syncthreads();
  float* prow = part + (int64_t)blockIdx.x * (K * N + N);
  for (int e = tid; e < K * N + N; e += 256) {
    float sum = 0.f;
    if (e < K * N) {
      const int kk = e / N, n = e - kk * N;
      for (int rr = 0; rr < R; ++rr) sum += red[(rr * K + kk) * NT + n];
    } else {
      const int n = e - K * N;
      for (int rr = 0; rr < R; ++rr) sum += redb[rr * NT + n];
    }
    prow[e] = sum;
  }
}

constexpr int kSmallNMax = 32;
constexpr int64_t kSmallSmemMax = 96 * 1024;

inline int nt_for(int64_t N) { return N <= 1 ? 1 : (N <= 4 ? 4 : (N <= 16 ? 16 : 32)); }
inline int64_t small_smem(int64_t N, int64_t K) { return K * (nt_for(N) | 1) * (int64_t)sizeof(float); }
inline bool small_ok(int64_t N, int64_t K) { return N >= 1 && N <= kSmallNMax && K >= 1 && small_smem(N, K) <= kSmallSmemMax; }

unsigned int* done_counter() {
  static unsigned int* p = nullptr;     // (looked up once: cudaGetSymbolAddress is a slow runtime call)
  if (!p) cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_done_counter);
  return p + current_lane() * 8;
}

template <typename Kern>
int set_smem(Kern kern, int64_t bytes) {
  if (bytes > 48 * 1024) NGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmemMax));
  return NGB_OK;
}

template <int NT, bool HEAD>
int launch_small_fwd(const float* x, const float* W, const float* b, float* z, float* a, int act, float alpha, int64_t M, int64_t N,
                     int64_t K, const HeadArgs& h, cudaStream_t st) {
  auto kern = linear_smalln_fwd_kernel<NT, HEAD>;
  const int64_t smem = small_smem(N, K);
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  const int threads = 256, rows_per_cta = threads / 32;
  int64_t grid = ceil_div<int64_t>(M, rows_per_cta);
  const int64_t cap = 4 * kNumSMs;                           // rows are independent: many warps in flight hide the row loads
  if (grid > cap) grid = cap;
  kern<<<(int)grid, threads, smem, st>>>(x, W, b, z, a, act, alpha, M, (int)N, (int)K, h);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

template <bool HEAD>
int small_fwd(const float* x, const float* W, const float* b, float* z, float* a, int act, float alpha, int64_t M, int64_t N, int64_t K,
              const HeadArgs& h, cudaStream_t st) {
  switch (nt_for(N)) {
    case 1: return launch_small_fwd<1, HEAD>(x, W, b, z, a, act, alpha, M, N, K, h, st);
    case 4: return launch_small_fwd<4, HEAD>(x, W, b, z, a, act, alpha, M, N, K, h, st);
    case 16: return launch_small_fwd<16, HEAD>(x, W, b, z, a, act, alpha, M, N, K, h, st);
    default: return launch_small_fwd<32, HEAD>(x, W, b, z, a, act, alpha, M, N, K, h, st);
  }
}

template <int NT>
int launch_small_dgrad(const float* ug, const float* W, const float* zprev, float* dx, int act, float alpha, int64_t M, int64_t N,
                       int64_t K, int accumulate, cudaStream_t st) {
  auto kern = linear_smalln_dgrad_kernel<NT>;
  const int64_t smem = small_smem(N, K);
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  kern<<<grid_for(M * K, 256, 4), 256, smem, st>>>(ug, W, zprev, dx, act, alpha, M, (int)N, (int)K, accumulate, FastDiv((uint32_t)K));
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

template <int NT>
int launch_small_wgrad(const float* x, const float* ug, float* dW, float* db, int64_t M, int64_t N, int64_t K, int acc_w, int acc_b,
                       cudaStream_t st) {
  const int R = K <= 256 ? (int)(256 / K) : 0;              // rows in flight per block of the slab kernel
  // (x of at least 4 MB: below that the single-pass kernel is as fast and one launch fewer; weight gradient alone: likewise)
  if (db != nullptr && K <= 256 && R * NT <= 256 && M * K >= (int64_t(1) << 20)) {
    int rc = ensure_init();
    if (rc) return rc;
    const int64_t row_floats = K * N + N;
    int64_t grid = ceil_div<int64_t>(M, (int64_t)R * 4);    // at least four rows per thread
    if (grid > 2 * kNumSMs) grid = 2 * kNumSMs;
    const int64_t cap = scratch_bytes() / (int64_t)sizeof(float) / row_floats;
    if (grid > cap) grid = cap;
    if (grid >= 1) {
      const int rows_per_block = (int)ceil_div<int64_t>(M, grid);
      grid = ceil_div<int64_t>(M, rows_per_block);
      float* part = scratch_ptr();
      linear_smalln_wgrad_slab_kernel<NT><<<(int)grid, 256, 0, st>>>(x, ug, part, M, (int)N, (int)K, rows_per_block, R);
      NGB_LAUNCH_CHECK();
      return launch_sum_partials2(part, (int)grid, K * N, dW, acc_w, N, db, acc_b, st);
    }
  }
  constexpr int KB = NT <= 4 ? 8 : (NT <= 16 ? 4 : 2);       // KB x NT <= 64 accumulators per thread
  const int blocks = (int)ceil_div<int64_t>(K, KB) + 1;      // + the bias block
  linear_smalln_wgrad_kernel<NT, KB><<<blocks, 256, 0, st>>>(x, ug, dW, db, M, (int)N, (int)K, acc_w, acc_b);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

bool use_tc(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, const void* A, const void* B) {
  return g_default_engine.load() == NGB_GEMM_AUTO && tc::gemm_tc_eligible(transA, transB, M, N, K, lda, ldb, N) && aligned16(A) &&
         aligned16(B);
}

}  // namespace
}  // namespace ngb

using namespace ngb;

extern "C" int ngb_linear_fwd(const float* x, const float* W, const float* b, float* z, float* a, int act, float alpha,
                              int64_t batch, int64_t n_in, int64_t n_out, ngb_stream_t stream) {
  NGB_CHECK_ARG(batch >= 0 && n_in >= 1 && n_out >= 1, "ngb_linear_fwd: bad shape");
  NGB_CHECK_ARG(z != nullptr || a != nullptr, "ngb_linear_fwd: no output requested");
  NGB_CHECK_ARG(act >= -1 && act <= NGB_TANH && act != NGB_EXP && act != NGB_LOG, "ngb_linear_fwd: unsupported activation %d", act);
  if (batch == 0) return NGB_OK;
  cudaStream_t st = as_stream(stream);
  if (small_ok(n_out, n_in)) return small_fwd<false>(x, W, b, z, a, act, alpha, batch, n_out, n_in, HeadArgs{}, st);
  tc::Epilogue e;
  e.bias = b; e.act = act; e.alpha = alpha; e.C2 = a; e.ldc2 = n_out; e.store_c = z ? 1 : 0;
  if (use_tc(0, 0, batch, n_out, n_in, n_in, n_out, x, W))
    return tc::gemm_tc(0, 0, batch, n_out, n_in, x, n_in, W, n_out, z, n_out, 0, st, &e);
  return gemm_simt(0, 0, batch, n_out, n_in, x, n_in, W, n_out, z, n_out, 0, st, &e);
}

extern "C" int ngb_linear_dgrad(const float* ug, const float* W, const float* zprev, float* dx, int act, float alpha,
                                int64_t batch, int64_t n_in, int64_t n_out, int accumulate, ngb_stream_t stream) {
  NGB_CHECK_ARG(batch >= 0 && n_in >= 1 && n_out >= 1, "ngb_linear_dgrad: bad shape");
  NGB_CHECK_ARG(act >= -1 && act <= NGB_TANH && act != NGB_EXP && act != NGB_LOG, "ngb_linear_dgrad: unsupported activation %d", act);
  if (batch == 0) return NGB_OK;
  cudaStream_t st = as_stream(stream);
  if (small_ok(n_out, n_in)) {
    switch (nt_for(n_out)) {
      case 1: return launch_small_dgrad<1>(ug, W, zprev, dx, act, alpha, batch, n_out, n_in, accumulate, st);
      case 4: return launch_small_dgrad<4>(ug, W, zprev, dx, act, alpha, batch, n_out, n_in, accumulate, st);
      case 16: return launch_small_dgrad<16>(ug, W, zprev, dx, act, alpha, batch, n_out, n_in, accumulate, st);
      default: return launch_small_dgrad<32>(ug, W, zprev, dx, act, alpha, batch, n_out, n_in, accumulate, st);
    }
  }
  tc::Epilogue e;
  e.act = act; e.alpha = alpha; e.mask_z = zprev; e.ldm = n_in;
  // dx[batch, n_in] = ug[batch, n_out] . W[n_in, n_out]^T
  if (use_tc(0, 1, batch, n_in, n_out, n_out, n_out, ug, W))
    return tc::gemm_tc(0, 1, batch, n_in, n_out, ug, n_out, W, n_out, dx, n_in, accumulate, st, &e);
  return gemm_simt(0, 1, batch, n_in, n_out, ug, n_out, W, n_out, dx, n_in, accumulate, st, &e);
}

extern "C" int ngb_linear_wgrad(const float* x, const float* ug, float* dW, float* db, int64_t batch, int64_t n_in, int64_t n_out,
                                int accumulate_w, int accumulate_b, ngb_stream_t stream) {
  NGB_CHECK_ARG(batch >= 0 && n_in >= 1 && n_out >= 1, "ngb_linear_wgrad: bad shape");
  NGB_CHECK_ARG(dW != nullptr, "ngb_linear_wgrad: dW is NULL");
  cudaStream_t st = as_stream(stream);
  if (batch == 0) {
    if (!accumulate_w) NGB_CUDA(cudaMemsetAsync(dW, 0, n_in * n_out * sizeof(float), st));
    if (db && !accumulate_b) NGB_CUDA(cudaMemsetAsync(db, 0, n_out * sizeof(float), st));
    return NGB_OK;
  }
  if (n_out <= kSmallNMax && n_in * nt_for(n_out) <= 4096) {
    switch (nt_for(n_out)) {
      case 1: return launch_small_wgrad<1>(x, ug, dW, db, batch, n_out, n_in, accumulate_w, accumulate_b, st);
      case 4: return launch_small_wgrad<4>(x, ug, dW, db, batch, n_out, n_in, accumulate_w, accumulate_b, st);
      case 16: return launch_small_wgrad<16>(x, ug, dW, db, batch, n_out, n_in, accumulate_w, accumulate_b, st);
      default: return launch_small_wgrad<32>(x, ug, dW, db, batch, n_out, n_in, accumulate_w, accumulate_b, st);
    }
  }
  tc::Epilogue e;
  e.colsum = db; e.colsum_accumulate = accumulate_b;
  // dW[n_in, n_out] = x[batch, n_in]^T . ug[batch, n_out]
  if (use_tc(1, 0, n_in, n_out, batch, n_in, n_out, x, ug))
    return tc::gemm_tc(1, 0, n_in, n_out, batch, x, n_in, ug, n_out, dW, n_out, accumulate_w, st, &e);
  return gemm_simt(1, 0, n_in, n_out, batch, x, n_in, ug, n_out, dW, n_out, accumulate_w, st, &e);
}

extern "C" int ngb_linear_softmax_ce_serves(int64_t batch, int64_t n_in, int64_t n_out) {
  return (batch >= 1 && small_ok(n_out, n_in)) ? 1 : 0;
}

extern "C" int ngb_linear_softmax_ce(const float* x, const float* W, const float* b, const float* targets, float* z, float* loss,
                                     float* dz, int64_t batch, int64_t n_in, int64_t n_out, float inv_num_examples, float epsilon,
                                     ngb_stream_t stream) {
  NGB_CHECK_ARG(batch >= 1 && n_in >= 1 && n_out >= 1, "ngb_linear_softmax_ce: bad shape");
  NGB_CHECK_ARG(targets && loss && dz, "ngb_linear_softmax_ce: NULL argument");
  if (!small_ok(n_out, n_in)) return fail(NGB_ERR_UNSUPPORTED, "ngb_linear_softmax_ce: serves layers with at most 32 outputs");
  int rc = ensure_init();
  if (rc) return rc;
  HeadArgs h;
  h.t = targets; h.inv_n = inv_num_examples; h.eps = epsilon; h.dz = dz; h.loss = loss;
  h.partials = scratch_ptr(); h.done = done_counter();
  return small_fwd<true>(x, W, b, z, nullptr, -1, 0.f, batch, n_out, n_in, h, as_stream(stream));
}
