// Development probe for conv_igemm_tcgen05.cu: includes the kernel translation unit, runs one small forward conv and
// checks it against a host loop; reports pipeline timeouts (which barrier) instead of a poisoned context.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I neograd_b200/csrc tools/igemm_probe.cu -L neograd_b200/lib -lngb200 -o tools/bin/igemm_probe
#include "../neograd_b200/csrc/conv_igemm_tcgen05.cu"
#include <vector>
#include <cmath>
using namespace ngb;

int run(int N, int C, int H, int W, int O, int KH, int KW, int pad) {
  const int Ho = H + 2 * pad - KH + 1, Wo = W + 2 * pad - KW + 1;
  std::vector<float> x((size_t)N * C * H * W), w((size_t)O * C * KH * KW), b(O), y((size_t)N * O * Ho * Wo);
  for (size_t i = 0; i < x.size(); ++i) x[i] = float(int((i * 7 + i / 13) % 7) - 3);
  for (size_t i = 0; i < w.size(); ++i) w[i] = float(int((i * 5 + i / 11) % 5) - 2);
  for (int i = 0; i < O; ++i) b[i] = float(i % 3);
  float *dx, *dw, *db, *dy;
  cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dw, w.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dy, y.size() * 4);
  cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dy, 0xff, y.size() * 4);
  int rc = conv_igemm_fprop(dx, dw, db, dy, N, C, H, W, O, KH, KW, pad, 0);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long to = tc::take_mbar_timeout();
  printf("conv %dx%dx%dx%d O=%d k=%dx%d pad=%d: rc=%d (%s) cuda=%s timeout=%llx (tag %llu block %llu thread %llu)\n", N, C, H, W, O, KH, KW,
         pad, rc, rc ? ngb_last_error() : "", cudaGetErrorString(e), to, (to >> 48) & 0x7fff, (to >> 16) & 0xffffffff, to & 0xffff);
  if (e != cudaSuccess) return 1;
  cudaMemcpy(y.data(), dy, y.size() * 4, cudaMemcpyDeviceToHost);
  size_t bad = 0; double maxerr = 0;
  for (int n = 0; n < N; ++n) for (int o = 0; o < O; ++o) for (int p = 0; p < Ho; ++p) for (int q = 0; q < Wo; ++q) {
    double acc = b[o];
    for (int c = 0; c < C; ++c) for (int kh = 0; kh < KH; ++kh) for (int kw = 0; kw < KW; ++kw) {
      const int h = p + kh - pad, ww = q + kw - pad;
      if (h >= 0 && h < H && ww >= 0 && ww < W) acc += double(x[((size_t)(n * C + c) * H + h) * W + ww]) * w[((size_t)(o * C + c) * KH + kh) * KW + kw];
    }
    const float got = y[((size_t)(n * O + o) * Wo + q) * Ho + p];   // neograd order: (p,q) at q*Ho + p
    const double err = fabs(got - acc);
    if (!(err <= 1e-3)) { if (bad < 5) printf("  mismatch n%d o%d p%d q%d: got %g want %g\n", n, o, p, q, got, acc); ++bad; }
    if (err > maxerr) maxerr = err;
  }
  printf("  mismatches %zu / %zu (max err %g)\n", bad, y.size(), maxerr);
  cudaFree(dx); cudaFree(dw); cudaFree(db); cudaFree(dy);
  return 0;
}

int main() {
  if (ngb_init(0)) { printf("init failed: %s\n", ngb_last_error()); return 1; }
  if (run(2, 32, 32, 32, 32, 3, 3, 1)) return 1;
  if (run(2, 32, 16, 16, 64, 3, 3, 1)) return 1;
  if (run(1, 64, 20, 24, 96, 3, 3, 1)) return 1;
  if (run(1, 32, 36, 40, 128, 5, 5, 2)) return 1;
  return 0;
}
