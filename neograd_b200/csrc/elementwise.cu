// Bandwidth-bound elementwise kernels: broadcasting binary ops (forward and the fused
// local-grad x upstream-grad + unbroadcast-reduction + accumulate backward), unary ops, axis sums,
// fill / axpy / transpose.   Reference arithmetic: neograd/autograd/ops/basics.py, autograd/utils.py:34-78.
//
// Design for B200: every kernel is HBM-bound, so the rules are (1) 128-bit coalesced accesses on the contiguous
// axis, (2) grids sized as a multiple of 148 SMs with grid-stride loops, (3) reductions go warp-shuffle ->
// shared memory -> (optional) deterministic second pass over per-CTA partials in the scratch arena; no atomics,
// so results are bit-reproducible run to run (needed for 1/2/4/8-GPU parity).
#include "common.cuh"

namespace ngb {

constexpr int MAXD = 6;
enum { OP_IDENT = 100 };

struct BDims {
  int rank;
  int64_t size[MAXD];
  int64_t sa[MAXD];
  int64_t sb[MAXD];
  bool red[MAXD];
};

// ------------------------------------------------------------------------------------------------ device math
__device__ __forceinline__ float binary_apply(int op, float a, float b) {
  switch (op) {
    case NGB_ADD: return a + b;
    case NGB_SUB: return a - b;
    case NGB_MUL: return a * b;
    case NGB_DIV: return a / b;
    case NGB_POW: return powf(a, b);
    default: return b;  // NGB_SECOND
  }
}

// grad of (a op b) wrt operand `side`, times ug.  basics.py:32-33, 76-77, 119-120, 163-164, 325-327
__device__ __forceinline__ float binary_grad(int op, int side, float a, float b, float ug) {
  switch (op) {
    case NGB_ADD: return ug;
    case NGB_SUB: return side == 0 ? ug : -ug;
    case NGB_MUL: return side == 0 ? b * ug : a * ug;
    case NGB_DIV: return side == 0 ? (1.f / b) * ug : ((-a) / (b * b)) * ug;
    case NGB_POW: return side == 0 ? (powf(a, b - 1.f) * b) * ug : (powf(a, b) * logf(a)) * ug;
    case NGB_SECOND: return side == 0 ? 0.f : ug;
    default: return ug;  // OP_IDENT
  }
}

__device__ __forceinline__ float unary_apply(int op, float alpha, float x) {
  switch (op) {
    case NGB_EXP: return expf(x);
    case NGB_LOG: return logf(x);
    case NGB_RELU: return fmaxf(0.f, x);
    case NGB_LEAKY_RELU: return x >= 0.f ? x : alpha * x;
    case NGB_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return tanhf(x);
  }
}

__device__ __forceinline__ float unary_grad(int op, float alpha, float x, float ug) {
  switch (op) {
    case NGB_EXP: return expf(x) * ug;
    case NGB_LOG: return (1.f / x) * ug;
    case NGB_RELU: return x >= 0.f ? ug : 0.f;              // activations.py:31 -- grad is 1 at x == 0
    case NGB_LEAKY_RELU: return x >= 0.f ? ug : alpha * ug;  // activations.py:216
    case NGB_SIGMOID: {
      // s(1 - s) (activations.py:62-63) written as e^-|x| / (1 + e^-|x|)^2: the same number, but `1 - s` cancels in fp32 once
      // the unit saturates (relative error 6e-8 / (1 - s)), which the float64 reference does not suffer from
      const float t = expf(-fabsf(x)), d = 1.f + t;
      return (t / (d * d)) * ug;
    }
    default: {
      // 1 - tanh^2 (activations.py:94-95) = 4 e^-2|x| / (1 + e^-2|x|)^2, for the same reason
      const float t = expf(-2.f * fabsf(x)), d = 1.f + t;
      return (4.f * t / (d * d)) * ug;
    }
  }
}

// ------------------------------------------------------------------------------------------------ flat kernels
// AM/BM: 0 = pointer (contiguous, same length as out), 1 = scalar immediate.
template <int AM, int BM>
__global__ void __launch_bounds__(256) ew_binary_flat_kernel(int op, const float* __restrict__ a, float as,
                                                             const float* __restrict__ b, float bs,
                                                             float* __restrict__ out, int64_t n, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nthreads) {
      float4 va, vb, r;
      if (AM == 0) va = ld_stream4(a + 4 * i); else va = make_float4(as, as, as, as);
      if (BM == 0) vb = ld_stream4(b + 4 * i); else vb = make_float4(bs, bs, bs, bs);
      r.x = binary_apply(op, va.x, vb.x);
      r.y = binary_apply(op, va.y, vb.y);
      r.z = binary_apply(op, va.z, vb.z);
      r.w = binary_apply(op, va.w, vb.w);
      st_stream4(out + 4 * i, r);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nthreads) {
    float x = AM == 0 ? a[i] : as;
    float y = BM == 0 ? b[i] : bs;
    out[i] = binary_apply(op, x, y);
  }
}

// Generic broadcast forward: one thread per output element, coordinates by div/mod on the collapsed dims.
// The innermost collapsed dim is walked by consecutive threads, so accesses stay coalesced (stride 1) or
// broadcast (stride 0) for every NumPy-broadcastable pair.
__global__ void __launch_bounds__(256) ew_binary_generic_kernel(int op, BDims d, const float* __restrict__ a, float as,
                                                                const float* __restrict__ b, float bs,
                                                                float* __restrict__ out, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t rem = i, oa = 0, ob = 0;
#pragma unroll
    for (int k = MAXD - 1; k >= 0; --k) {
      if (k < d.rank) {
        int64_t c = rem % d.size[k];
        rem /= d.size[k];
        oa += c * d.sa[k];
        ob += c * d.sb[k];
      }
    }
    float x = a ? a[oa] : as;
    float y = b ? b[ob] : bs;
    out[i] = binary_apply(op, x, y);
  }
}

// Row-broadcast fast path (bias add): out[r, c] = a[r*sa0 + c*sa1] op b[r*sb0 + c*sb1], inner strides in {0,1},
// inner size % 4 == 0, everything 16B aligned: 128-bit accesses.
__global__ void __launch_bounds__(256) ew_binary_rows4_kernel(int op, const float* __restrict__ a, int64_t sa0, int sa1,
                                                              const float* __restrict__ b, int64_t sb0, int sb1,
                                                              float* __restrict__ out, int64_t rows, int64_t cols4) {
  const int64_t total = rows * cols4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols4, c4 = i - r * cols4;
    float4 va, vb, o;
    if (sa1) va = ld_stream4(a + r * sa0 + 4 * c4);
    else { float s = __ldg(a + r * sa0); va = make_float4(s, s, s, s); }
    if (sb1) vb = __ldg(reinterpret_cast<const float4*>(b + r * sb0 + 4 * c4));
    else { float s = __ldg(b + r * sb0); vb = make_float4(s, s, s, s); }
    o.x = binary_apply(op, va.x, vb.x);
    o.y = binary_apply(op, va.y, vb.y);
    o.z = binary_apply(op, va.z, vb.z);
    o.w = binary_apply(op, va.w, vb.w);
    st_stream4(out + 4 * i, o);
  }
}

// ------------------------------------------------------------------------------------------------ backward kernels
// No reduction: grad[i] (+)= f(a[oa], b[ob], ug[i]).
__global__ void __launch_bounds__(256) ew_bwd_elem_kernel(int op, int side, BDims d, const float* __restrict__ ug,
                                                          const float* __restrict__ a, float as,
                                                          const float* __restrict__ b, float bs,
                                                          float* __restrict__ grad, int accumulate, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t rem = i, oa = 0, ob = 0;
#pragma unroll
    for (int k = MAXD - 1; k >= 0; --k) {
      if (k < d.rank) {
        int64_t c = rem % d.size[k];
        rem /= d.size[k];
        oa += c * d.sa[k];
        ob += c * d.sb[k];
      }
    }
    float g = binary_grad(op, side, a ? a[oa] : as, b ? b[ob] : bs, ug[i]);
    grad[i] = accumulate ? grad[i] + g : g;
  }
}

// Column reduction, pattern [R][K] (reduced rows, kept columns): the bias-gradient shape (N,out)->(1,out).
// CTA = 32 columns x 8 row-lanes; gridDim.y splits R; partial sums go to `part[split][K]` (or straight to grad
// when there is a single split).  Loads are coalesced along K.
__global__ void __launch_bounds__(256) ew_bwd_colreduce_kernel(int op, int side, const float* __restrict__ ug,
                                                               const float* __restrict__ a, float as, int64_t sa0, int64_t sa1,
                                                               const float* __restrict__ b, float bs, int64_t sb0, int64_t sb1,
                                                               float* __restrict__ dst, int accumulate, int64_t R, int64_t K,
                                                               int64_t rows_per_split) {
  __shared__ float tile[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t col = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r1 = min(R, r0 + rows_per_split);
  float acc = 0.f;
  if (col < K) {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float av = a ? a[r * sa0 + col * sa1] : as;
      float bv = b ? b[r * sb0 + col * sb1] : bs;
      acc += binary_grad(op, side, av, bv, ug[r * K + col]);
    }
  }
  tile[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < K) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += tile[j][tx];
    float* o = dst + (int64_t)blockIdx.y * K + col;
    *o = (accumulate && gridDim.y == 1) ? *o + s : s;
  }
}

// Column sum of ug itself (the local gradient of add / sub / identity is +-1): the bias-gradient hot case.  128-bit loads,
// 4 independent loads in flight per thread, CTA = 128 columns x 8 row lanes.
__global__ void __launch_bounds__(256) colsum4_kernel(const float* __restrict__ ug, float* __restrict__ dst, int accumulate,
                                                      int64_t R, int64_t K4, int64_t rows_per_split, float scale) {
  __shared__ float4 tile[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c4 = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r1 = min(R, r0 + rows_per_split);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 < K4) {
    const float* base = ug + 4 * c4;
    int64_t r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      const float4 a = ld_stream4(base + r * K4 * 4), b = ld_stream4(base + (r + 8) * K4 * 4);
      const float4 c = ld_stream4(base + (r + 16) * K4 * 4), d = ld_stream4(base + (r + 24) * K4 * 4);
      acc.x += (a.x + b.x) + (c.x + d.x); acc.y += (a.y + b.y) + (c.y + d.y);
      acc.z += (a.z + b.z) + (c.z + d.z); acc.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; r < r1; r += 8) {
      const float4 a = ld_stream4(base + r * K4 * 4);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  tile[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c4 < K4) {
    float4 s4 = tile[0][tx];
#pragma unroll
    for (int j = 1; j < 8; ++j) { const float4 t = tile[j][tx]; s4.x += t.x; s4.y += t.y; s4.z += t.z; s4.w += t.w; }
    s4.x *= scale; s4.y *= scale; s4.z *= scale; s4.w *= scale;
    float4* o = reinterpret_cast<float4*>(dst + ((int64_t)blockIdx.y * K4 + c4) * 4);
    if (accumulate && gridDim.y == 1) { const float4 old = *o; s4.x += old.x; s4.y += old.y; s4.z += old.z; s4.w += old.w; }
    *o = s4;
  }
}

// Full reduction, pattern [R]: every CTA grid-strides the whole range and leaves one partial.
__global__ void __launch_bounds__(256) ew_bwd_fullreduce_kernel(int op, int side, const float* __restrict__ ug,
                                                                const float* __restrict__ a, float as, int64_t sa0,
                                                                const float* __restrict__ b, float bs, int64_t sb0,
                                                                float* __restrict__ dst, int accumulate, int64_t R) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x)
    acc += binary_grad(op, side, a ? a[r * sa0] : as, b ? b[r * sb0] : bs, ug[r]);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) dst[blockIdx.x] = (accumulate && gridDim.x == 1) ? dst[blockIdx.x] + acc : acc;
}

// Sum `splits` partial rows of length K into out (deterministic second pass).  CTA = 32 columns x 8 split lanes: lane j
// adds splits j, j+8, ... and the eight lane sums are combined in a fixed order, so the result does not depend on timing.
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ part, int splits, int64_t K,
                                                           float* __restrict__ out, int accumulate,
    float* __restrict__ out2 = nullptr, int64_t K1 = 0, int accumulate2 = 0) {   // out2: columns >= K1 go to out2[k - K1]
  __shared__ float tile[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int64_t k0 = (int64_t)blockIdx.x * 32; k0 < K; k0 += (int64_t)gridDim.x * 32) {
    const int64_t k = k0 + tx;
    float s = 0.f;
    if (k < K) {
      int j = ty;
      for (; j + 24 < splits; j += 32) {
        const float a = part[(int64_t)j * K + k], b = part[(int64_t)(j + 8) * K + k];
        const float c = part[(int64_t)(j + 16) * K + k], d = part[(int64_t)(j + 24) * K + k];
        s += (a + b) + (c + d);
      }
      for (; j < splits; j += 8) s += part[(int64_t)j * K + k];
    }
    tile[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && k < K) {
      float r = tile[0][tx];
#pragma unroll
      for (int j = 1; j < 8; ++j) r += tile[j][tx];
      float* dst = (out2 && k >= K1) ? out2 + (k - K1) : out + k;
      const int ac = (out2 && k >= K1) ? accumulate2 : accumulate;
      *dst = ac ? *dst + r : r;
    }
    __syncthreads();
  }
}

// Same sum for MANY partials of a SHORT row (the per-CTA partials of the persistent conv weight-gradient kernels: ~600 x 150):
// 8 columns x 32 split lanes per CTA, so a thread walks splits/32 rows instead of splits/8 -- the kernel is a chain of
// dependent L2 round trips, its time is the length of that chain.  Lane sums are combined in a fixed order.
__global__ void __launch_bounds__(256) sum_partials_tall_kernel(const float* __restrict__ part, int splits, int64_t K,
                                                                float* __restrict__ out, int accumulate,
    float* __restrict__ out2 = nullptr, int64_t K1 = 0, int accumulate2 = 0) {   // out2: columns >= K1 go to out2[k - K1]
  __shared__ float tile[32][9];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  for (int64_t k0 = (int64_t)blockIdx.x * 8; k0 < K; k0 += (int64_t)gridDim.x * 8) {
    const int64_t k = k0 + tx;
    float s = 0.f;
    if (k < K) {
      int j = ty;
      for (; j + 96 < splits; j += 128) {
        const float a = part[(int64_t)j * K + k], b = part[(int64_t)(j + 32) * K + k];
        const float c = part[(int64_t)(j + 64) * K + k], d = part[(int64_t)(j + 96) * K + k];
        s += (a + b) + (c + d);
      }
      for (; j < splits; j += 32) s += part[(int64_t)j * K + k];
    }
    tile[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && k < K) {
      float r = tile[0][tx];
#pragma unroll
      for (int j = 1; j < 32; ++j) r += tile[j][tx];
      float* dst = (out2 && k >= K1) ? out2 + (k - K1) : out + k;
      const int ac = (out2 && k >= K1) ? accumulate2 : accumulate;
      *dst = ac ? *dst + r : r;
    }
    __syncthreads();
  }
}

// Row reduction, pattern [K][R]: one warp per kept row, lanes stride the contiguous reduced axis.
__global__ void __launch_bounds__(256) ew_bwd_rowreduce_kernel(int op, int side, const float* __restrict__ ug,
                                                               const float* __restrict__ a, float as, int64_t sa0, int64_t sa1,
                                                               const float* __restrict__ b, float bs, int64_t sb0, int64_t sb1,
                                                               float* __restrict__ grad, int accumulate, int64_t K, int64_t R) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp; k < K; k += nwarps) {
    float acc = 0.f;
    for (int64_t r = lane; r < R; r += 32) {
      float av = a ? a[k * sa0 + r * sa1] : as;
      float bv = b ? b[k * sb0 + r * sb1] : bs;
      acc += binary_grad(op, side, av, bv, ug[k * R + r]);
    }
    acc = warp_sum(acc);
    if (lane == 0) grad[k] = accumulate ? grad[k] + acc : acc;
  }
}

// General fallback: one CTA per kept element, threads walk the reduced index space.
__global__ void __launch_bounds__(256) ew_bwd_general_kernel(int op, int side, BDims d, const float* __restrict__ ug,
                                                             const float* __restrict__ a, float as,
                                                             const float* __restrict__ b, float bs,
                                                             float* __restrict__ grad, int accumulate, int64_t Ktot, int64_t Rtot) {
  __shared__ float red[32];
  // contiguous strides of the broadcast shape
  int64_t su[MAXD];
  {
    int64_t s = 1;
    for (int k = d.rank - 1; k >= 0; --k) { su[k] = s; s *= d.size[k]; }
  }
  for (int64_t o = blockIdx.x; o < Ktot; o += gridDim.x) {
    // decompose o over kept dims
    int64_t base_u = 0, base_a = 0, base_b = 0, rem = o;
    for (int k = d.rank - 1; k >= 0; --k) {
      if (!d.red[k]) {
        int64_t c = rem % d.size[k];
        rem /= d.size[k];
        base_u += c * su[k]; base_a += c * d.sa[k]; base_b += c * d.sb[k];
      }
    }
    float acc = 0.f;
    for (int64_t r = threadIdx.x; r < Rtot; r += blockDim.x) {
      int64_t ou = base_u, oa = base_a, ob = base_b, rr = r;
      for (int k = d.rank - 1; k >= 0; --k) {
        if (d.red[k]) {
          int64_t c = rr % d.size[k];
          rr /= d.size[k];
          ou += c * su[k]; oa += c * d.sa[k]; ob += c * d.sb[k];
        }
      }
      acc += binary_grad(op, side, a ? a[oa] : as, b ? b[ob] : bs, ug[ou]);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) grad[o] = accumulate ? grad[o] + acc : acc;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ unary / misc kernels
__global__ void __launch_bounds__(256) ew_unary_fwd_kernel(int op, float alpha, const float* __restrict__ x,
                                                           float* __restrict__ out, int64_t n, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nt) {
      float4 v = ld_stream4(x + 4 * i), r;
      r.x = unary_apply(op, alpha, v.x); r.y = unary_apply(op, alpha, v.y);
      r.z = unary_apply(op, alpha, v.z); r.w = unary_apply(op, alpha, v.w);
      st_stream4(out + 4 * i, r);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nt) out[i] = unary_apply(op, alpha, x[i]);
}

__global__ void __launch_bounds__(256) ew_unary_bwd_kernel(int op, float alpha, const float* __restrict__ x,
                                                           const float* __restrict__ ug, float* __restrict__ out,
                                                           int64_t n, int accumulate, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nt) {
      float4 v = ld_stream4(x + 4 * i), g = ld_stream4(ug + 4 * i), r;
      r.x = unary_grad(op, alpha, v.x, g.x); r.y = unary_grad(op, alpha, v.y, g.y);
      r.z = unary_grad(op, alpha, v.z, g.z); r.w = unary_grad(op, alpha, v.w, g.w);
      if (accumulate) {
        float4 o = *reinterpret_cast<const float4*>(out + 4 * i);
        r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
      }
      st_stream4(out + 4 * i, r);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nt) {
    float g = unary_grad(op, alpha, x[i], ug[i]);
    out[i] = accumulate ? out[i] + g : g;
  }
}

__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ dst, float v, int64_t n, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4 v4 = make_float4(v, v, v, v);
    for (int64_t i = tid; i < n4; i += nt) st_stream4(dst + 4 * i, v4);
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nt) dst[i] = v;
}

__global__ void __launch_bounds__(256) axpy_kernel(float* __restrict__ dst, const float* __restrict__ src, float alpha,
                                                   int64_t n, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nt) {
      float4 s = ld_stream4(src + 4 * i);
      float4 d = *reinterpret_cast<const float4*>(dst + 4 * i);
      d.x += alpha * s.x; d.y += alpha * s.y; d.z += alpha * s.z; d.w += alpha * s.w;
      *reinterpret_cast<float4*>(dst + 4 * i) = d;
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nt) dst[i] += alpha * src[i];
}

// 32x32 tiled transpose through padded shared memory (both sides coalesced).
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                        int64_t rows, int64_t cols) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t tiles_c = ceil_div<int64_t>(cols, 32), tiles_r = ceil_div<int64_t>(rows, 32);
  for (int64_t t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const int64_t tr = t / tiles_c, tc = t - tr * tiles_c;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      int64_t r = tr * 32 + ty + j, c = tc * 32 + tx;
      if (r < rows && c < cols) tile[ty + j][tx] = src[r * cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      int64_t c = tc * 32 + ty + j, r = tr * 32 + tx;
      if (r < rows && c < cols) dst[c * rows + r] = tile[tx][ty + j];
    }
    __syncthreads();
  }
}

// General axis permutation: one thread per output element; consecutive threads write consecutive addresses.
struct PermDims { int rank; int64_t out_size[MAXD]; int64_t src_stride[MAXD]; };
__global__ void __launch_bounds__(256) permute_kernel(const float* __restrict__ src, float* __restrict__ dst, PermDims d,
                                                      int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t rem = i, off = 0;
#pragma unroll
    for (int k = MAXD - 1; k >= 0; --k) {
      if (k < d.rank) {
        const int64_t c = rem % d.out_size[k];
        rem /= d.out_size[k];
        off += c * d.src_stride[k];
      }
    }
    dst[i] = __ldg(src + off);
  }
}

// ------------------------------------------------------------------------------------------------ host side
// Builds the (collapsed) broadcast description.  Returns false on incompatible shapes.
static bool build_dims(const int64_t* a_shape, int a_rank, const int64_t* b_shape, int b_rank, int side /* -1: fwd */,
                       BDims* out, int64_t* total) {
  int rank = a_rank > b_rank ? a_rank : b_rank;
  if (rank > 16) return false;
  int64_t size[16], sa[16], sb[16];
  bool red[16];
  int64_t ca = 1, cb = 1;
  for (int k = rank - 1; k >= 0; --k) {
    int ka = k - (rank - a_rank), kb = k - (rank - b_rank);
    int64_t da = ka >= 0 ? a_shape[ka] : 1, db = kb >= 0 ? b_shape[kb] : 1;
    if (da != db && da != 1 && db != 1) return false;
    int64_t s = da == 1 ? db : da;
    size[k] = s;
    sa[k] = (da == 1 && s != 1) ? 0 : ca;
    sb[k] = (db == 1 && s != 1) ? 0 : cb;
    if (da == 1) sa[k] = 0;   // size-1 dims never advance
    if (db == 1) sb[k] = 0;
    // utils.py:66-70: an axis is summed when the operand's dim differs from the broadcast dim (missing counts)
    int64_t dop = side == 0 ? (ka >= 0 ? da : -1) : (kb >= 0 ? db : -1);
    red[k] = side >= 0 && dop != s;
    ca *= da;
    cb *= db;
  }
  // drop size-1 dims, then merge neighbours with compatible strides and equal reduce flag
  int n = 0;
  int64_t csize[16], csa[16], csb[16];
  bool cred[16];
  int64_t tot = 1;
  for (int k = 0; k < rank; ++k) {
    tot *= size[k];
    if (size[k] == 1) continue;
    if (n > 0 && cred[n - 1] == red[k] && csa[n - 1] == sa[k] * size[k] && csb[n - 1] == sb[k] * size[k]) {
      csize[n - 1] *= size[k];
      csa[n - 1] = sa[k];
      csb[n - 1] = sb[k];
    } else if (n > 0 && cred[n - 1] == red[k] && csa[n - 1] == 0 && sa[k] == 0 && csb[n - 1] == sb[k] * size[k]) {
      csize[n - 1] *= size[k];
      csb[n - 1] = sb[k];
    } else if (n > 0 && cred[n - 1] == red[k] && csb[n - 1] == 0 && sb[k] == 0 && csa[n - 1] == sa[k] * size[k]) {
      csize[n - 1] *= size[k];
      csa[n - 1] = sa[k];
    } else if (n > 0 && cred[n - 1] == red[k] && csa[n - 1] == 0 && sa[k] == 0 && csb[n - 1] == 0 && sb[k] == 0) {
      csize[n - 1] *= size[k];
    } else {
      csize[n] = size[k]; csa[n] = sa[k]; csb[n] = sb[k]; cred[n] = red[k];
      ++n;
    }
  }
  if (n > MAXD) return false;
  out->rank = n;
  for (int k = 0; k < MAXD; ++k) {
    out->size[k] = k < n ? csize[k] : 1;
    out->sa[k] = k < n ? csa[k] : 0;
    out->sb[k] = k < n ? csb[k] : 0;
    out->red[k] = k < n ? cred[k] : false;
  }
  *total = tot;
  return true;
}

static int reduce_dispatch(int op, int side, const BDims& d, int64_t total, const float* ug, const float* a, float as,
                           const float* b, float bs, float* grad, int accumulate, cudaStream_t st) {
  if (total == 0) return NGB_OK;
  int64_t K = 1, R = 1;
  int nred = 0;
  for (int k = 0; k < d.rank; ++k) {
    if (d.red[k]) { R *= d.size[k]; ++nred; } else K *= d.size[k];
  }
  if (nred == 0) {
    ew_bwd_elem_kernel<<<grid_for(total, 256), 256, 0, st>>>(op, side, d, ug, a, as, b, bs, grad, accumulate, total);
    NGB_LAUNCH_CHECK();
    return NGB_OK;
  }
  if (d.rank == 1 && d.red[0]) {
    int grid = grid_for(R, 256 * 8, 2);
    float* dst = grad;
    if (grid > 1) {
      NGB_CHECK_ARG(ensure_init() == NGB_OK, "scratch init failed");
      dst = scratch_ptr();
    }
    ew_bwd_fullreduce_kernel<<<grid, 256, 0, st>>>(op, side, ug, a, as, d.sa[0], b, bs, d.sb[0], dst, accumulate, R);
    NGB_LAUNCH_CHECK();
    if (grid > 1) {
      sum_partials_kernel<<<1, 256, 0, st>>>(scratch_ptr(), grid, 1, grad, accumulate);
      NGB_LAUNCH_CHECK();
    }
    return NGB_OK;
  }
  const bool col_pat = d.rank == 2 && d.red[0] && !d.red[1];
  const bool row_pat = d.rank == 2 && !d.red[0] && d.red[1];
  if (col_pat) {
    const int64_t Kc = d.size[1], Rc = d.size[0];
    const int64_t sa0 = d.sa[0], sa1 = d.sa[1], sb0 = d.sb[0], sb1 = d.sb[1];
    // local gradient == +-ug (add, sub, identity): vectorised column sum
    const bool plain = op == NGB_ADD || op == NGB_SUB || op == OP_IDENT || (op == NGB_SECOND && side == 1);
    if (plain && Kc % 4 == 0 && aligned16(ug) && aligned16(grad)) {
      const float scale = (op == NGB_SUB && side == 1) ? -1.f : 1.f;
      const int64_t K4 = Kc / 4, cb = ceil_div<int64_t>(K4, 32);
      int64_t splits = ceil_div<int64_t>(4 * kNumSMs, cb);
      const int64_t max_splits = ceil_div<int64_t>(Rc, 128);
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
      if (splits > 1) {
        NGB_CHECK_ARG(ensure_init() == NGB_OK, "scratch init failed");
        if (splits * Kc * (int64_t)sizeof(float) > scratch_bytes()) splits = scratch_bytes() / (Kc * sizeof(float));
        if (splits < 1) splits = 1;
      }
      const int64_t rps = ceil_div<int64_t>(Rc, splits);
      splits = ceil_div<int64_t>(Rc, rps);
      dim3 grid((unsigned)cb, (unsigned)splits);
      colsum4_kernel<<<grid, 256, 0, st>>>(ug, splits > 1 ? scratch_ptr() : grad, accumulate, Rc, K4, rps, scale);
      NGB_LAUNCH_CHECK();
      if (splits > 1) {
        sum_partials_kernel<<<grid_for(Kc * 8, 256), 256, 0, st>>>(scratch_ptr(), (int)splits, Kc, grad, accumulate);
        NGB_LAUNCH_CHECK();
      }
      return NGB_OK;
    }
    const int64_t colblocks = ceil_div<int64_t>(Kc, 32);
    int64_t splits = ceil_div<int64_t>(2 * kNumSMs, colblocks);
    const int64_t max_splits = ceil_div<int64_t>(Rc, 64);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 1) {
      NGB_CHECK_ARG(ensure_init() == NGB_OK, "scratch init failed");
      if (splits * Kc * (int64_t)sizeof(float) > scratch_bytes()) splits = scratch_bytes() / (Kc * sizeof(float));
      if (splits < 1) splits = 1;
    }
    const int64_t rps = ceil_div<int64_t>(Rc, splits);
    splits = ceil_div<int64_t>(Rc, rps);
    dim3 grid((unsigned)colblocks, (unsigned)splits);
    float* dst = splits > 1 ? scratch_ptr() : grad;
    ew_bwd_colreduce_kernel<<<grid, 256, 0, st>>>(op, side, ug, a, as, sa0, sa1, b, bs, sb0, sb1, dst, accumulate, Rc, Kc, rps);
    NGB_LAUNCH_CHECK();
    if (splits > 1) {
      sum_partials_kernel<<<grid_for(Kc, 256), 256, 0, st>>>(scratch_ptr(), (int)splits, Kc, grad, accumulate);
      NGB_LAUNCH_CHECK();
    }
    return NGB_OK;
  }
  if (row_pat) {
    const int64_t Kr = d.size[0], Rr = d.size[1];
    ew_bwd_rowreduce_kernel<<<grid_for(Kr * 32, 256), 256, 0, st>>>(op, side, ug, a, as, d.sa[0], d.sa[1], b, bs, d.sb[0],
                                                                  d.sb[1], grad, accumulate, Kr, Rr);
    NGB_LAUNCH_CHECK();
    return NGB_OK;
  }
  int grid = (int)(K < (int64_t)kNumSMs * 8 ? K : (int64_t)kNumSMs * 8);
  ew_bwd_general_kernel<<<grid, 256, 0, st>>>(op, side, d, ug, a, as, b, bs, grad, accumulate, K, R);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

// rows of K1 + K2 columns: the first K1 sums go to out1, the rest to out2 (weight + bias gradient partials in one launch)
int launch_sum_partials2(const float* part, int splits, int64_t K1, float* out1, int acc1, int64_t K2, float* out2, int acc2,
                         cudaStream_t st) {
  const int64_t K = K1 + K2;
  if (splits >= 128 && K <= 8192)
    sum_partials_tall_kernel<<<grid_for(K * 32, 256), 256, 0, st>>>(part, splits, K, out1, acc1, out2, K1, acc2);
  else
    sum_partials_kernel<<<grid_for(K * 8, 256), 256, 0, st>>>(part, splits, K, out1, acc1, out2, K1, acc2);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

int launch_sum_partials(const float* part, int splits, int64_t K, float* out, int accumulate, cudaStream_t st) {
  if (splits >= 128 && K <= 8192)
    sum_partials_tall_kernel<<<grid_for(K * 32, 256), 256, 0, st>>>(part, splits, K, out, accumulate);
  else
    sum_partials_kernel<<<grid_for(K * 8, 256), 256, 0, st>>>(part, splits, K, out, accumulate);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

}  // namespace ngb

using namespace ngb;

extern "C" int ngb_ew_binary_fwd(int op, const float* a, float a_scalar, const int64_t* a_shape, int a_rank,
                                 const float* b, float b_scalar, const int64_t* b_shape, int b_rank, float* out,
                                 ngb_stream_t stream) {
  NGB_CHECK_ARG(op >= NGB_ADD && op <= NGB_SECOND, "ngb_ew_binary_fwd: bad op %d", op);
  NGB_CHECK_ARG(a_rank >= 0 && b_rank >= 0, "negative rank");
  BDims d;
  int64_t total = 0;
  NGB_CHECK_ARG(build_dims(a_shape, a_rank, b_shape, b_rank, -1, &d, &total),
                "operands could not be broadcast together");
  if (total == 0) return NGB_OK;
  cudaStream_t st = as_stream(stream);
  int64_t na = 1, nb = 1;
  for (int k = 0; k < a_rank; ++k) na *= a_shape[k];
  for (int k = 0; k < b_rank; ++k) nb *= b_shape[k];
  const bool a_full = a && na == total, b_full = b && nb == total;
  const bool a_sc = !a, b_sc = !b;
  if ((a_full || a_sc) && (b_full || b_sc)) {
    const int vec = aligned16(out) && (a_sc || aligned16(a)) && (b_sc || aligned16(b));
    const int grid = grid_for(ceil_div<int64_t>(total, 4), 256);
    if (!a_sc && !b_sc) ew_binary_flat_kernel<0, 0><<<grid, 256, 0, st>>>(op, a, 0.f, b, 0.f, out, total, vec);
    else if (a_sc && !b_sc) ew_binary_flat_kernel<1, 0><<<grid, 256, 0, st>>>(op, nullptr, a_scalar, b, 0.f, out, total, vec);
    else if (!a_sc && b_sc) ew_binary_flat_kernel<0, 1><<<grid, 256, 0, st>>>(op, a, 0.f, nullptr, b_scalar, out, total, vec);
    else ew_binary_flat_kernel<1, 1><<<grid, 256, 0, st>>>(op, nullptr, a_scalar, nullptr, b_scalar, out, total, vec);
    NGB_LAUNCH_CHECK();
    return NGB_OK;
  }
  if (a && b && d.rank == 2 && (d.sa[1] == 0 || d.sa[1] == 1) && (d.sb[1] == 0 || d.sb[1] == 1) && d.size[1] % 4 == 0 &&
      d.sa[0] % 4 == 0 && d.sb[0] % 4 == 0 && aligned16(a) && aligned16(b) && aligned16(out)) {
    const int64_t cols4 = d.size[1] / 4;
    ew_binary_rows4_kernel<<<grid_for(d.size[0] * cols4, 256), 256, 0, st>>>(op, a, d.sa[0], (int)d.sa[1], b, d.sb[0],
                                                                           (int)d.sb[1], out, d.size[0], cols4);
    NGB_LAUNCH_CHECK();
    return NGB_OK;
  }
  ew_binary_generic_kernel<<<grid_for(total, 256), 256, 0, st>>>(op, d, a, a_scalar, b, b_scalar, out, total);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_ew_binary_bwd(int op, int side, const float* ug, const float* a, float a_scalar,
                                 const int64_t* a_shape, int a_rank, const float* b, float b_scalar,
                                 const int64_t* b_shape, int b_rank, float* grad, int accumulate, ngb_stream_t stream) {
  NGB_CHECK_ARG(op >= NGB_ADD && op <= NGB_SECOND, "ngb_ew_binary_bwd: bad op %d", op);
  NGB_CHECK_ARG(side == 0 || side == 1, "side must be 0 or 1");
  BDims d;
  int64_t total = 0;
  NGB_CHECK_ARG(build_dims(a_shape, a_rank, b_shape, b_rank, side, &d, &total), "operands could not be broadcast together");
  return reduce_dispatch(op, side, d, total, ug, a, a_scalar, b, b_scalar, grad, accumulate, as_stream(stream));
}

extern "C" int ngb_reduce_sum(const float* x, const int64_t* shape, int rank, uint32_t axis_mask, float* out,
                              int accumulate, ngb_stream_t stream) {
  NGB_CHECK_ARG(rank >= 0 && rank <= 16, "rank out of range");
  // Express as the backward of IDENT for an operand whose reduced axes have size 1.
  int64_t op_shape[16];
  for (int k = 0; k < rank; ++k) op_shape[k] = (axis_mask >> k) & 1u ? 1 : shape[k];
  BDims d;
  int64_t total = 0;
  // side 0 operand = op_shape, operand b = full shape (only used to define the broadcast shape)
  NGB_CHECK_ARG(build_dims(op_shape, rank, shape, rank, 0, &d, &total), "bad shapes");
  // axes of size 1 that are being "reduced" are no-ops and were dropped by build_dims
  return reduce_dispatch(OP_IDENT, 0, d, total, x, nullptr, 0.f, nullptr, 0.f, out, accumulate, as_stream(stream));
}

extern "C" int ngb_ew_unary_fwd(int op, float alpha, const float* x, float* out, int64_t n, ngb_stream_t stream) {
  NGB_CHECK_ARG(op >= NGB_EXP && op <= NGB_TANH, "ngb_ew_unary_fwd: bad op %d", op);
  if (n == 0) return NGB_OK;
  const int vec = aligned16(x) && aligned16(out);
  ew_unary_fwd_kernel<<<grid_for(ceil_div<int64_t>(n, 4), 256), 256, 0, as_stream(stream)>>>(op, alpha, x, out, n, vec);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_ew_unary_bwd(int op, float alpha, const float* x, const float* ug, float* out, int64_t n,
                                int accumulate, ngb_stream_t stream) {
  NGB_CHECK_ARG(op >= NGB_EXP && op <= NGB_TANH, "ngb_ew_unary_bwd: bad op %d", op);
  if (n == 0) return NGB_OK;
  const int vec = aligned16(x) && aligned16(out) && aligned16(ug);
  ew_unary_bwd_kernel<<<grid_for(ceil_div<int64_t>(n, 4), 256), 256, 0, as_stream(stream)>>>(op, alpha, x, ug, out, n,
                                                                                            accumulate, vec);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_fill(float* dst, float value, int64_t n, ngb_stream_t stream) {
  if (n == 0) return NGB_OK;
  fill_kernel<<<grid_for(ceil_div<int64_t>(n, 4), 256), 256, 0, as_stream(stream)>>>(dst, value, n, aligned16(dst));
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_axpy(float* dst, const float* src, float alpha, int64_t n, ngb_stream_t stream) {
  if (n == 0) return NGB_OK;
  axpy_kernel<<<grid_for(ceil_div<int64_t>(n, 4), 256), 256, 0, as_stream(stream)>>>(dst, src, alpha, n,
                                                                                    aligned16(dst) && aligned16(src));
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_copy(float* dst, const float* src, int64_t n, ngb_stream_t stream) {
  NGB_CHECK_ARG(n >= 0, "ngb_copy: negative length");
  if (n == 0 || dst == src) return NGB_OK;
  NGB_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)));
  return NGB_OK;
}

extern "C" int ngb_permute(const float* src, float* dst, const int64_t* shape, const int* perm, int rank, ngb_stream_t stream) {
  NGB_CHECK_ARG(rank >= 1 && rank <= MAXD, "ngb_permute: rank must be 1..%d", MAXD);
  int64_t stride[MAXD], total = 1;
  for (int k = rank - 1; k >= 0; --k) { stride[k] = total; total *= shape[k]; }
  PermDims d;
  d.rank = rank;
  bool seen[MAXD] = {};
  for (int k = 0; k < MAXD; ++k) { d.out_size[k] = 1; d.src_stride[k] = 0; }
  for (int k = 0; k < rank; ++k) {
    NGB_CHECK_ARG(perm[k] >= 0 && perm[k] < rank && !seen[perm[k]], "ngb_permute: perm is not a permutation");
    seen[perm[k]] = true;
    d.out_size[k] = shape[perm[k]];
    d.src_stride[k] = stride[perm[k]];
  }
  if (total == 0) return NGB_OK;
  permute_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(src, dst, d, total);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_transpose2d(const float* src, float* dst, int64_t rows, int64_t cols, ngb_stream_t stream) {
  if (rows * cols == 0) return NGB_OK;
  int64_t tiles = ceil_div<int64_t>(rows, 32) * ceil_div<int64_t>(cols, 32);
  int grid = (int)(tiles < (int64_t)kNumSMs * 8 ? tiles : (int64_t)kNumSMs * 8);
  transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, dst, rows, cols);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}
