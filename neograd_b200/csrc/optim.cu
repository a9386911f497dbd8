// Fused optimizer steps over a flat fp32 parameter arena (one launch for the whole model).
// Reference: neograd/nn/optim.py -- GD 42-47, Momentum 63-92, RMSProp 112-140, Adam 168-197.
// HBM-bound: Adam moves 28 B/element (read p,g,m,v; write p,m,v), SGD 12 B/element; 128-bit accesses, grid-stride,
// grid a multiple of 148 SMs.  `grad_scale` folds the data-parallel 1/world averaging into the same pass.
#include "common.cuh"

namespace ngb {

struct AdamP { float lr, b1, b2, eps, ibc1, ibc2, gs; };

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamP& a) {
  g *= a.gs;
  m = a.b1 * m + (1.f - a.b1) * g;          // optim.py:196
  v = a.b2 * v + (1.f - a.b2) * g * g;      // optim.py:197
  const float mh = m * a.ibc1;              // optim.py:177
  const float vh = v * a.ibc2;              // optim.py:178
  p -= a.lr * (mh / (sqrtf(vh) + a.eps));   // optim.py:179
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, AdamP a, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nt) {
      float4 pp = *reinterpret_cast<float4*>(p + 4 * i), gg = ld_stream4(g + 4 * i);
      float4 mm = *reinterpret_cast<float4*>(m + 4 * i), vv = *reinterpret_cast<float4*>(v + 4 * i);
      adam1(pp.x, gg.x, mm.x, vv.x, a); adam1(pp.y, gg.y, mm.y, vv.y, a);
      adam1(pp.z, gg.z, mm.z, vv.z, a); adam1(pp.w, gg.w, mm.w, vv.w, a);
      *reinterpret_cast<float4*>(p + 4 * i) = pp;
      *reinterpret_cast<float4*>(m + 4 * i) = mm;
      *reinterpret_cast<float4*>(v + 4 * i) = vv;
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nt) adam1(p[i], g[i], m[i], v[i], a);
}

__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g, int64_t n, float lr,
                                                  float gs, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nt) {
      float4 pp = *reinterpret_cast<float4*>(p + 4 * i), gg = ld_stream4(g + 4 * i);
      pp.x -= lr * (gg.x * gs); pp.y -= lr * (gg.y * gs); pp.z -= lr * (gg.z * gs); pp.w -= lr * (gg.w * gs);
      *reinterpret_cast<float4*>(p + 4 * i) = pp;
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nt) p[i] -= lr * (g[i] * gs);   // optim.py:47
}

__global__ void __launch_bounds__(256) momentum_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, int64_t n, float lr, float beta, float gs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float mm = beta * m[i] + (1.f - beta) * (g[i] * gs);   // optim.py:92
    m[i] = mm;
    p[i] -= lr * mm;                                             // optim.py:70
  }
}

__global__ void __launch_bounds__(256) rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ r, int64_t n, float lr, float beta, float eps,
                                                      float gs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gg = g[i] * gs;
    const float rr = beta * r[i] + (1.f - beta) * gg * gg;       // optim.py:140
    r[i] = rr;
    p[i] -= lr * (gg / (sqrtf(rr) + eps));                       // optim.py:118
  }
}

}  // namespace ngb

using namespace ngb;

extern "C" int ngb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                             float epsilon, float inv_bias_corr1, float inv_bias_corr2, float grad_scale,
                             ngb_stream_t stream) {
  if (n == 0) return NGB_OK;
  AdamP a{lr, beta1, beta2, epsilon, inv_bias_corr1, inv_bias_corr2, grad_scale};
  const int vec = aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v);
  adam_kernel<<<grid_for(ceil_div<int64_t>(n, 4), 256), 256, 0, as_stream(stream)>>>(p, g, m, v, n, a, vec);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_sgd_step(float* p, const float* g, int64_t n, float lr, float grad_scale, ngb_stream_t stream) {
  if (n == 0) return NGB_OK;
  sgd_kernel<<<grid_for(ceil_div<int64_t>(n, 4), 256), 256, 0, as_stream(stream)>>>(p, g, n, lr, grad_scale,
                                                                                   aligned16(p) && aligned16(g));
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_momentum_step(float* p, const float* g, float* m, int64_t n, float lr, float beta, float grad_scale,
                                 ngb_stream_t stream) {
  if (n == 0) return NGB_OK;
  momentum_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(p, g, m, n, lr, beta, grad_scale);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_rmsprop_step(float* p, const float* g, float* r, int64_t n, float lr, float beta, float epsilon,
                                float grad_scale, ngb_stream_t stream) {
  if (n == 0) return NGB_OK;
  rmsprop_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(p, g, r, n, lr, beta, epsilon, grad_scale);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}
