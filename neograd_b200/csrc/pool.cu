// MaxPool2D / MaxPool3D forward (+ uint8 argmax) and backward.  HBM-bound.
// Reference: neograd/autograd/ops/conv.py:384-402 / 483-501 (forward), 404-431 / 503-530 (backward).
//
// Semantics reproduced exactly:
//   * the window is taken over the ZERO-padded input, so padding participates in the max with value 0;
//   * ties -> first element in row-major window order (np.argmax);
//   * outputs are stored in neograd order: position (p,q) at flat index q*Ho + p;
//   * backward ASSIGNS whole window blocks in fragment order (W outer, H inner): an input cell ends up with the
//     value written by the LAST window that covers it (ug if the cell is that window's argmax, else 0).
//     Written here as a gather (one thread per input cell, no atomics, deterministic); uncovered cells get 0.
#include "common.cuh"

namespace ngb {

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          uint8_t* __restrict__ idx, int64_t outer, int H, int W, int Ho,
                                                          int Wo, int KH, int KW, int pad, int stride) {
  const int64_t hw = (int64_t)Ho * Wo, total = outer * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = i / hw;
    const int f = (int)(i - o * hw);
    const int q = f / Ho, p = f - q * Ho;
    const float* xi = x + o * (int64_t)H * W;
    float best = 0.f;
    int arg = 0;
    bool first = true;
    for (int kh = 0; kh < KH; ++kh) {
      const int h = p * stride + kh - pad;
      for (int kw = 0; kw < KW; ++kw) {
        const int w = q * stride + kw - pad;
        const float v = (h >= 0 && h < H && w >= 0 && w < W) ? __ldg(xi + (int64_t)h * W + w) : 0.f;
        if (first || v > best) { best = v; arg = kh * KW + kw; first = false; }
      }
    }
    y[i] = best;
    if (idx) idx[i] = (uint8_t)arg;
  }
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ ug, const uint8_t* __restrict__ idx,
                                                          float* __restrict__ dx, int64_t outer, int H, int W, int Ho,
                                                          int Wo, int KH, int KW, int pad, int stride, int accumulate) {
  const int64_t hw_in = (int64_t)H * W, total = outer * hw_in, hw_out = (int64_t)Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = i / hw_in;
    const int r = (int)(i - o * hw_in);
    const int h = r / W, w = r - h * W;
    const int hp = h + pad, wp = w + pad;
    int q = wp / stride, p = hp / stride;
    if (q > Wo - 1) q = Wo - 1;
    if (p > Ho - 1) p = Ho - 1;
    float g = 0.f;
    if (q * stride + KW > wp && p * stride + KH > hp) {  // covered by window (p,q): the last one in fragment order
      const int64_t f = o * hw_out + (int64_t)q * Ho + p;
      const int local = (hp - p * stride) * KW + (wp - q * stride);
      if ((int)idx[f] == local) g = ug[f];
    }
    dx[i] = accumulate ? dx[i] + g : g;
  }
}

}  // namespace ngb

using namespace ngb;

static int pool_geometry(int64_t H, int64_t W, int64_t KH, int64_t KW, int64_t pad, int64_t stride, int64_t* Ho, int64_t* Wo) {
  NGB_CHECK_ARG(stride >= 1, "Stride must be greater than or equal to 1");
  NGB_CHECK_ARG(pad >= 0, "Padding must be greater than or equal to zero");
  NGB_CHECK_ARG(KH >= 1 && KW >= 1 && KH * KW <= 256, "pool window must have 1..256 cells (uint8 argmax)");
  NGB_CHECK_ARG(H + 2 * pad >= KH && W + 2 * pad >= KW, "pool window larger than the padded input");
  *Ho = (H + 2 * pad - KH) / stride + 1;
  *Wo = (W + 2 * pad - KW) / stride + 1;
  return NGB_OK;
}

extern "C" int ngb_maxpool_fwd(const float* x, float* y, uint8_t* idx, int64_t outer, int64_t H, int64_t W, int64_t KH,
                               int64_t KW, int64_t pad, int64_t stride, ngb_stream_t stream) {
  int64_t Ho, Wo;
  int rc = pool_geometry(H, W, KH, KW, pad, stride, &Ho, &Wo);
  if (rc) return rc;
  const int64_t total = outer * Ho * Wo;
  if (total == 0) return NGB_OK;
  maxpool_fwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(x, y, idx, outer, (int)H, (int)W, (int)Ho, (int)Wo,
                                                                          (int)KH, (int)KW, (int)pad, (int)stride);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_maxpool_bwd(const float* ug, const uint8_t* idx, float* dx, int64_t outer, int64_t H, int64_t W,
                               int64_t KH, int64_t KW, int64_t pad, int64_t stride, int accumulate, ngb_stream_t stream) {
  int64_t Ho, Wo;
  int rc = pool_geometry(H, W, KH, KW, pad, stride, &Ho, &Wo);
  if (rc) return rc;
  NGB_CHECK_ARG(idx != nullptr, "ngb_maxpool_bwd needs the argmax indices saved by ngb_maxpool_fwd");
  const int64_t total = outer * H * W;
  if (total == 0) return NGB_OK;
  maxpool_bwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(ug, idx, dx, outer, (int)H, (int)W, (int)Ho, (int)Wo,
                                                                          (int)KH, (int)KW, (int)pad, (int)stride, accumulate);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}
