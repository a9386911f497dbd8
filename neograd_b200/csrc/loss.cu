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

}  // namespace ngb

using namespace ngb;

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
