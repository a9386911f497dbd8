// Fused optimizer steps over a flat fp32 parameter arena (one launch for the whole model).
// Reference: neograd/nn/optim.py -- GD 42-47, Momentum 63-92, RMSProp 112-140, Adam 168-197.
// HBM-bound: Adam moves 28 B/element (read p,g,m,v; write p,m,v), SGD 12 B/element; 128-bit accesses, grid-stride,
// grid a multiple of 148 SMs.  `grad_scale` folds the data-parallel 1/world averaging into the same pass.
// Hyper-parameters arrive as doubles; 1-beta and the bias corrections are formed in double and rounded once.
#include "common.cuh"

namespace ngb {

struct AdamP { float lr, b1, omb1, b2, omb2, eps, ibc1, ibc2, gs; };

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamP& a) {
  g *= a.gs;
  m = a.b1 * m + a.omb1 * g;                // optim.py:196
  v = a.b2 * v + a.omb2 * (g * g);          // optim.py:197
  const float mh = m * a.ibc1;              // optim.py:177
  const float vh = v * a.ibc2;              // optim.py:178
  p -= a.lr * (mh / (sqrtf(vh) + a.eps));   // optim.py:179
}

// T_DEV: the iteration count lives in device memory: t_dev[0] = steps taken so far, t_dev[1] = blocks-done counter (0 between
// launches).  tick = 1: this launch is step t_dev[0] + 1 and records it when its last block finishes (every block has read
// the old value by then); tick = 2: same step number, but further launches of the step follow (frozen parameters split the
// arena into several ranges), nothing is recorded; tick = 0: use t_dev[0] as is.
template <bool T_DEV>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, AdamP a, int32_t* __restrict__ t_dev,
                                                   double b1d, double b2d, int vec_ok, int tick) {
  int32_t step = 0;
  if (T_DEV) {
    step = *reinterpret_cast<volatile int32_t*>(t_dev) + (tick ? 1 : 0);
    const double t = (double)step;
    a.ibc1 = (float)(1.0 / (1.0 - pow(b1d, t)));
    a.ibc2 = (float)(1.0 / (1.0 - pow(b2d, t)));
  }
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
  if (T_DEV && tick == 1) {                                        // optim.py:169, recorded by the last block to finish
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      unsigned int* cnt = reinterpret_cast<unsigned int*>(t_dev + 1);
      if (atomicAdd(cnt, 1u) == gridDim.x - 1) {
        t_dev[0] = step;
        *cnt = 0;
      }
    }
  }
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
                                                       float* __restrict__ m, int64_t n, float lr, float beta, float omb,
                                                       float gs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float mm = beta * m[i] + omb * (g[i] * gs);            // optim.py:92
    m[i] = mm;
    p[i] -= lr * mm;                                             // optim.py:70
  }
}

__global__ void __launch_bounds__(256) rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ r, int64_t n, float lr, float beta, float omb,
                                                      float eps, float gs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gg = g[i] * gs;
    const float rr = beta * r[i] + omb * (gg * gg);              // optim.py:140
    r[i] = rr;
    p[i] -= lr * (gg / (sqrtf(rr) + eps));                       // optim.py:118
  }
}

static AdamP make_adam(double lr, double b1, double b2, double eps, double t, double gs) {
  AdamP a;
  a.lr = (float)lr; a.b1 = (float)b1; a.omb1 = (float)(1.0 - b1); a.b2 = (float)b2; a.omb2 = (float)(1.0 - b2);
  a.eps = (float)eps; a.gs = (float)gs;
  a.ibc1 = t > 0 ? (float)(1.0 / (1.0 - pow(b1, t))) : 1.f;
  a.ibc2 = t > 0 ? (float)(1.0 / (1.0 - pow(b2, t))) : 1.f;
  return a;
}

}  // namespace ngb

using namespace ngb;

extern "C" int ngb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                             double epsilon, int64_t t, double grad_scale, ngb_stream_t stream) {
  NGB_CHECK_ARG(t >= 1, "ngb_adam_step: t is the 1-based iteration count");
  if (n == 0) return NGB_OK;
  const AdamP a = make_adam(lr, beta1, beta2, epsilon, (double)t, grad_scale);
  const int vec = aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v);
  adam_kernel<false><<<grid_for(ceil_div<int64_t>(n, 4), 256), 256, 0, as_stream(stream)>>>(p, g, m, v, n, a, nullptr, 0, 0, vec, 0);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1,
                                 double beta2, double epsilon, int32_t* t_dev, int tick, double grad_scale,
                                 ngb_stream_t stream) {
  NGB_CHECK_ARG(t_dev != nullptr, "ngb_adam_step_dev: t_dev is NULL");
  NGB_CHECK_ARG(tick >= 0 && tick <= 2, "ngb_adam_step_dev: tick must be 0, 1 or 2");
  cudaStream_t st = as_stream(stream);
  const AdamP a = make_adam(lr, beta1, beta2, epsilon, 0, grad_scale);
  const int vec = aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v);
  // (n == 0 with tick == 1 still launches one block: the step must be counted)
  if (n == 0 && tick != 1) return NGB_OK;
  adam_kernel<true><<<grid_for(ceil_div<int64_t>(n > 0 ? n : 1, 4), 256), 256, 0, st>>>(p, g, m, v, n, a, t_dev, beta1, beta2, vec, tick);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_sgd_step(float* p, const float* g, int64_t n, double lr, double grad_scale, ngb_stream_t stream) {
  if (n == 0) return NGB_OK;
  sgd_kernel<<<grid_for(ceil_div<int64_t>(n, 4), 256), 256, 0, as_stream(stream)>>>(p, g, n, (float)lr, (float)grad_scale,
                                                                                   aligned16(p) && aligned16(g));
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_momentum_step(float* p, const float* g, float* m, int64_t n, double lr, double beta, double grad_scale,
                                 ngb_stream_t stream) {
  if (n == 0) return NGB_OK;
  momentum_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(p, g, m, n, (float)lr, (float)beta, (float)(1.0 - beta),
                                                                  (float)grad_scale);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_rmsprop_step(float* p, const float* g, float* r, int64_t n, double lr, double beta, double epsilon,
                                double grad_scale, ngb_stream_t stream) {
  if (n == 0) return NGB_OK;
  rmsprop_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(p, g, r, n, (float)lr, (float)beta, (float)(1.0 - beta),
                                                                 (float)epsilon, (float)grad_scale);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}
