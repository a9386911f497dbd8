// SoftmaxCE forward / backward (neograd/nn/loss.py:124-172, softmax at nn/activations.py:161-174).
// logits and targets are viewed as (outer, L, inner) with the softmax along L, which covers any `axis` of a
// C-contiguous tensor.  One thread owns one (outer, inner) softmax row; rows are short (L = 10 classes), so the
// kernel is a single coalesced pass when inner == 1 is small and L fits in registers' reach via L1.
// The loss is reduced warp-shuffle -> shared memory -> per-CTA partial -> deterministic second pass.
#include "common.cuh"

namespace ngb {

__global__ void __launch_bounds__(256) softmax_ce_fwd_kernel(const float* __restrict__ z, const float* __restrict__ t,
                                                             int64_t outer, int64_t L, int64_t inner, float epsilon,
                                                             float* __restrict__ partial) {
  __shared__ float red[32];
  const int64_t rows = outer * inner;
  float acc = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = r / inner, in = r - o * inner;
    const float* zr = z + o * L * inner + in;
    const float* tr = t + o * L * inner + in;
    float mx = -INFINITY;
    for (int64_t l = 0; l < L; ++l) mx = fmaxf(mx, zr[l * inner]);
    float sum = 0.f;
    for (int64_t l = 0; l < L; ++l) sum += expf(zr[l * inner] - mx);
    const float inv = 1.f / sum;
    for (int64_t l = 0; l < L; ++l) {
      const float tv = tr[l * inner];
      if (tv != 0.f) acc += tv * logf(expf(zr[l * inner] - mx) * inv + epsilon);   // loss.py:155
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void softmax_ce_finish_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ loss) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) *loss = scale * s;   // loss.py:156: (-1/num_examples) * entropy
}

__global__ void __launch_bounds__(256) softmax_ce_bwd_kernel(const float* __restrict__ z, const float* __restrict__ t,
                                                             const float* __restrict__ ug, int64_t outer, int64_t L,
                                                             int64_t inner, float inv_n, float* __restrict__ dz,
                                                             int accumulate) {
  const int64_t rows = outer * inner;
  const float scale = (ug ? ug[0] : 1.f) * inv_n;   // loss.py:170: (ug / num_examples)
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = r / inner, in = r - o * inner;
    const int64_t base = o * L * inner + in;
    float mx = -INFINITY;
    for (int64_t l = 0; l < L; ++l) mx = fmaxf(mx, z[base + l * inner]);
    float sum = 0.f;
    for (int64_t l = 0; l < L; ++l) sum += expf(z[base + l * inner] - mx);
    const float inv = 1.f / sum;
    for (int64_t l = 0; l < L; ++l) {
      const int64_t a = base + l * inner;
      const float g = scale * (expf(z[a] - mx) * inv - t[a]);
      dz[a] = accumulate ? dz[a] + g : g;
    }
  }
}

// Softmax layer (activations.py:105-180): one thread per (outer, inner) slice.
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ x, int64_t outer, int64_t L, int64_t inner,
                                                          float* __restrict__ y) {
  const int64_t rows = outer * inner;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = r / inner, in = r - o * inner;
    const int64_t base = o * L * inner + in;
    float mx = -INFINITY;
    for (int64_t l = 0; l < L; ++l) mx = fmaxf(mx, x[base + l * inner]);
    float sum = 0.f;
    for (int64_t l = 0; l < L; ++l) sum += expf(x[base + l * inner] - mx);
    const float inv = 1.f / sum;
    for (int64_t l = 0; l < L; ++l) y[base + l * inner] = expf(x[base + l * inner] - mx) * inv;
  }
}

// dx_i = s_i * (ug_i - sum_j ug_j s_j): the product of the softmax Jacobian with ug (activations.py:131-158).
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ x, const float* __restrict__ ug,
                                                          int64_t outer, int64_t L, int64_t inner, float* __restrict__ dx,
                                                          int accumulate) {
  const int64_t rows = outer * inner;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = r / inner, in = r - o * inner;
    const int64_t base = o * L * inner + in;
    float mx = -INFINITY;
    for (int64_t l = 0; l < L; ++l) mx = fmaxf(mx, x[base + l * inner]);
    float sum = 0.f;
    for (int64_t l = 0; l < L; ++l) sum += expf(x[base + l * inner] - mx);
    const float inv = 1.f / sum;
    float dotp = 0.f;
    for (int64_t l = 0; l < L; ++l) dotp += ug[base + l * inner] * (expf(x[base + l * inner] - mx) * inv);
    for (int64_t l = 0; l < L; ++l) {
      const int64_t a = base + l * inner;
      const float g = (expf(x[a] - mx) * inv) * (ug[a] - dotp);
      dx[a] = accumulate ? dx[a] + g : g;
    }
  }
}

// BCE as ONE pass (nn/loss.py:69-85 builds it from eleven elementwise ops): entropy = sum(o * log(t + eps) +
// (1 - o) * log(1 - t + eps)), cost = (-1/n) * entropy -- with the reference's (swapped) argument roles kept as they are
// (SURVEY A.7).  Same fp32 operations per element as the composite; the reduction is per-block partials summed in a fixed
// order by the last block to finish.
__device__ unsigned int g_bce_done[kLanes];

__global__ void __launch_bounds__(256) bce_fwd_kernel(const float* __restrict__ o, const float* __restrict__ t, int64_t n, float eps,
                                                      float scale, float* __restrict__ partial, float* __restrict__ loss,
                                                      unsigned int* done) {
  __shared__ float red[32];
  __shared__ bool is_last;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float ov = o[i], tv = t[i];
    acc += ov * logf(tv + eps) + (1.f - ov) * logf((1.f - tv) + eps);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = acc;
    __threadfence();
    is_last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    const volatile float* pp = partial;
    const int nb = (int)gridDim.x, th = threadIdx.x;          // (grid <= 1024)
    const float a0 = th < nb ? pp[th] : 0.f, a1 = th + 256 < nb ? pp[th + 256] : 0.f;
    const float a2 = th + 512 < nb ? pp[th + 512] : 0.f, a3 = th + 768 < nb ? pp[th + 768] : 0.f;
    const float tot = block_sum((a0 + a1) + (a2 + a3), red);
    if (th == 0) {
      *loss = scale * tot;
      *done = 0;
    }
  }
}

// d cost / d o = scale * ug * (log(t + eps) - log(1 - t + eps));  d cost / d t = scale * ug * (o / (t + eps) - (1 - o) / (1 - t + eps))
__global__ void __launch_bounds__(256) bce_bwd_kernel(const float* __restrict__ o, const float* __restrict__ t,
                                                      const float* __restrict__ ug, int64_t n, float eps, float scale,
                                                      float* __restrict__ d_o, int acc_o, float* __restrict__ d_t, int acc_t) {
  const float g = scale * (ug ? ug[0] : 1.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float ov = o[i], tv = t[i];
    if (d_o) {
      const float v = g * (logf(tv + eps) - logf((1.f - tv) + eps));
      d_o[i] = acc_o ? d_o[i] + v : v;
    }
    if (d_t) {
      const float v = g * (ov / (tv + eps) - (1.f - ov) / ((1.f - tv) + eps));
      d_t[i] = acc_t ? d_t[i] + v : v;
    }
  }
}

}  // namespace ngb

using namespace ngb;

extern "C" int ngb_bce_fwd(const float* outputs, const float* targets, int64_t n, float inv_num_examples, float epsilon, float* loss,
                           ngb_stream_t stream) {
  NGB_CHECK_ARG(n >= 1 && outputs && targets && loss, "ngb_bce_fwd: bad arguments");
  int rc = ensure_init();
  if (rc) return rc;
  static unsigned int* done = nullptr;   // (looked up once)
  if (!done) NGB_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&done), g_bce_done));
  int grid = grid_for(n, 256, 2);
  if (grid > 1024) grid = 1024;
  bce_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(outputs, targets, n, epsilon, -inv_num_examples, scratch_ptr(), loss,
                                                      done + current_lane());
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_bce_bwd(const float* outputs, const float* targets, const float* ug, int64_t n, float inv_num_examples,
                           float epsilon, float* d_outputs, int accumulate_o, float* d_targets, int accumulate_t,
                           ngb_stream_t stream) {
  NGB_CHECK_ARG(n >= 1 && outputs && targets && (d_outputs || d_targets), "ngb_bce_bwd: bad arguments");
  bce_bwd_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(outputs, targets, ug, n, epsilon, -inv_num_examples, d_outputs,
                                                                 accumulate_o, d_targets, accumulate_t);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_softmax_fwd(const float* x, int64_t outer, int64_t L, int64_t inner, float* y, ngb_stream_t stream) {
  NGB_CHECK_ARG(outer >= 0 && L >= 1 && inner >= 0, "ngb_softmax_fwd: bad shape");
  if (outer * inner == 0) return NGB_OK;
  softmax_fwd_kernel<<<grid_for(outer * inner, 256), 256, 0, as_stream(stream)>>>(x, outer, L, inner, y);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_softmax_bwd(const float* x, const float* ug, int64_t outer, int64_t L, int64_t inner, float* dx,
                               int accumulate, ngb_stream_t stream) {
  NGB_CHECK_ARG(outer >= 0 && L >= 1 && inner >= 0, "ngb_softmax_bwd: bad shape");
  if (outer * inner == 0) return NGB_OK;
  softmax_bwd_kernel<<<grid_for(outer * inner, 256), 256, 0, as_stream(stream)>>>(x, ug, outer, L, inner, dx, accumulate);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_softmax_ce_fwd(const float* logits, const float* targets, int64_t outer, int64_t L, int64_t inner,
                                  float inv_num_examples, float epsilon, float* loss, ngb_stream_t stream) {
  NGB_CHECK_ARG(outer >= 1 && L >= 1 && inner >= 1, "ngb_softmax_ce_fwd: empty input");
  int rc = ensure_init();
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const int grid = grid_for(outer * inner, 256, 2);
  softmax_ce_fwd_kernel<<<grid, 256, 0, st>>>(logits, targets, outer, L, inner, epsilon, scratch_ptr());
  NGB_LAUNCH_CHECK();
  softmax_ce_finish_kernel<<<1, 32, 0, st>>>(scratch_ptr(), grid, -inv_num_examples, loss);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_softmax_ce_bwd(const float* logits, const float* targets, const float* ug, int64_t outer, int64_t L,
                                  int64_t inner, float inv_num_examples, float* dlogits, int accumulate,
                                  ngb_stream_t stream) {
  NGB_CHECK_ARG(outer >= 1 && L >= 1 && inner >= 1, "ngb_softmax_ce_bwd: empty input");
  softmax_ce_bwd_kernel<<<grid_for(outer * inner, 256), 256, 0, as_stream(stream)>>>(logits, targets, ug, outer, L, inner,
                                                                                    inv_num_examples, dlogits, accumulate);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}
