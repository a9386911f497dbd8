// Input pipeline on the device: what the reference's notebooks do to a batch in NumPy before
// `ng.tensor(...)` -- examples/digits.ipynb cell 5 (per-sample standardisation), cell 6 (`np.eye(10)[y]` one-hot),
// examples/gans/basic.ipynb (`np.random.randn` noise) -- as bandwidth kernels, so a batch crosses PCIe in its source format
// (uint8 pixels, integer class ids: 4x fewer bytes than fp32 + one-hot rows) and is expanded in HBM.
#include "common.cuh"

namespace ngb {

// dst[i] = src[i] * scale + shift; 16 pixels per thread (one 128-bit load, four 128-bit stores)
__global__ void __launch_bounds__(256) u8_affine_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t n,
                                                        float scale, float shift, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n16 = n >> 4;
    for (int64_t i = tid; i < n16; i += nt) {
      const uint4 q = *reinterpret_cast<const uint4*>(src + 16 * i);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 o;
        o.x = (float)(w[j] & 0xff) * scale + shift;
        o.y = (float)((w[j] >> 8) & 0xff) * scale + shift;
        o.z = (float)((w[j] >> 16) & 0xff) * scale + shift;
        o.w = (float)(w[j] >> 24) * scale + shift;
        st_stream4(dst + 16 * i + 4 * j, o);
      }
    }
    done = n16 << 4;
  }
  for (int64_t i = done + tid; i < n; i += nt) dst[i] = (float)src[i] * scale + shift;
}

// digits.ipynb cell 5: (x - mean(x, axis=1)) / std(x, axis=1) (population std), one warp per row.  U8: the source is uint8.
template <bool U8>
__global__ void __launch_bounds__(256) standardize_rows_kernel(const void* __restrict__ src_, float* __restrict__ dst, int64_t rows,
                                                               int cols) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    const uint8_t* s8 = reinterpret_cast<const uint8_t*>(src_) + r * cols;
    const float* s32 = reinterpret_cast<const float*>(src_) + r * cols;
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) sum += U8 ? (float)s8[c] : s32[c];
    const float mean = warp_sum(sum) / (float)cols;
    float ss = 0.f;
    for (int c = lane; c < cols; c += 32) { const float d = (U8 ? (float)s8[c] : s32[c]) - mean; ss += d * d; }
    const float inv = 1.f / sqrtf(warp_sum(ss) / (float)cols);
    for (int c = lane; c < cols; c += 32) dst[r * cols + c] = ((U8 ? (float)s8[c] : s32[c]) - mean) * inv;
  }
}

// digits.ipynb cell 6: np.eye(classes)[labels]
__global__ void __launch_bounds__(256) one_hot_kernel(const int32_t* __restrict__ labels, float* __restrict__ dst, int64_t rows,
                                                      int classes) {
  const int64_t n = rows * classes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / classes;
    dst[i] = (int)(i - r * classes) == labels[r] ? 1.f : 0.f;
  }
}

// Counter-based normal generator: Philox4x32-10 keyed by `seed`, counter = (element group, offset) -> 4 uniforms -> two
// Box-Muller pairs.  Stateless: element i of call `offset` is the same whatever the launch geometry.
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(256) randn_kernel(float* __restrict__ dst, int64_t n, uint64_t seed, const uint32_t* __restrict__ offset_dev,
                                                    uint32_t offset_host) {
  const uint32_t off = offset_dev ? *offset_dev : offset_host;
  const int64_t n4 = (n + 3) >> 2;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (int64_t)gridDim.x * blockDim.x) {
    uint32_t u[4];
    philox4x32((uint32_t)g, (uint32_t)(g >> 32), off, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), u);
    float z[4];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const float u1 = ((float)u[2 * p] + 1.f) * 2.3283064365386963e-10f;      // (0, 1]
      const float u2 = (float)u[2 * p + 1] * 2.3283064365386963e-10f;          // [0, 1)
      const float rad = sqrtf(-2.f * __logf(u1));
      float sn, cs;
      __sincosf(6.283185307179586f * u2, &sn, &cs);
      z[2 * p] = rad * cs;
      z[2 * p + 1] = rad * sn;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * g + j < n) dst[4 * g + j] = z[j];
  }
}

// advances a device-resident call counter (so a captured graph draws fresh noise on every replay)
__global__ void bump_kernel(uint32_t* c) { *c += 1; }

}  // namespace ngb

using namespace ngb;

extern "C" int ngb_prep_u8_affine(const uint8_t* src, float* dst, int64_t n, float scale, float shift, ngb_stream_t stream) {
  NGB_CHECK_ARG(n >= 0, "ngb_prep_u8_affine: negative size");
  if (n == 0) return NGB_OK;
  const int vec = aligned16(src) && aligned16(dst);
  u8_affine_kernel<<<grid_for(ceil_div<int64_t>(n, 16), 256), 256, 0, as_stream(stream)>>>(src, dst, n, scale, shift, vec);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_prep_standardize_rows(const void* src, int src_is_u8, float* dst, int64_t rows, int64_t cols, ngb_stream_t stream) {
  NGB_CHECK_ARG(rows >= 0 && cols >= 1 && cols <= 0x7fffffff, "ngb_prep_standardize_rows: bad shape");
  if (rows == 0) return NGB_OK;
  const int grid = grid_for(rows * 32, 256);
  if (src_is_u8) standardize_rows_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(src, dst, rows, (int)cols);
  else standardize_rows_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(src, dst, rows, (int)cols);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_prep_one_hot(const int32_t* labels, float* dst, int64_t rows, int64_t classes, ngb_stream_t stream) {
  NGB_CHECK_ARG(rows >= 0 && classes >= 1 && classes <= 0x7fffffff, "ngb_prep_one_hot: bad shape");
  if (rows == 0) return NGB_OK;
  one_hot_kernel<<<grid_for(rows * classes, 256), 256, 0, as_stream(stream)>>>(labels, dst, rows, (int)classes);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_randn(float* dst, int64_t n, uint64_t seed, uint32_t* offset_dev, uint32_t offset, ngb_stream_t stream) {
  NGB_CHECK_ARG(n >= 0, "ngb_randn: negative size");
  if (n == 0) return NGB_OK;
  randn_kernel<<<grid_for(ceil_div<int64_t>(n, 4), 256), 256, 0, as_stream(stream)>>>(dst, n, seed, offset_dev, offset);
  NGB_LAUNCH_CHECK();
  if (offset_dev) {
    bump_kernel<<<1, 1, 0, as_stream(stream)>>>(offset_dev);
    NGB_LAUNCH_CHECK();
  }
  return NGB_OK;
}
