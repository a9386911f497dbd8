// Direct (CUDA-core) convolution kernels: the general path for every geometry (any C, O, rectangular filters,
// padding, stride) and the hot path for the small-channel layers of the LeNet / digits configs, where
// K = C*KH*KW is far below what the tensor-core implicit GEMM needs.  The multi-channel, stride-1 layers of the
// C5 sweep go to conv_igemm_tcgen05.cu instead (see the dispatcher at the bottom).
// Reference: neograd/autograd/ops/conv.py:245-269 (fprop), 299-309 (dgrad), 311-321 (wgrad), 323-327 (bgrad);
// single-channel Conv2D (conv.py:131-203) is the C = O = 1 case.
//
// Layout: x (N,C,H,W) natural; y / ug (N,O,Ho,Wo) in neograd order (position (p,q) at q*Ho+p), i.e. H is the
// fastest output axis while W is the fastest input axis.  Every kernel stages its input patch in shared memory
// transposed, so both the global reads (along W) and the global writes (along p) are coalesced.
#include "common.cuh"

namespace ngb {

int conv_igemm_eligible(int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW, int64_t pad,
                        int64_t stride);
int conv_igemm_fprop(const float* x, const float* w, const float* b, float* y, int64_t N, int64_t C, int64_t H, int64_t W,
                     int64_t O, int64_t KH, int64_t KW, int64_t pad, cudaStream_t st);
int conv_igemm_dgrad(const float* ug, const float* w, float* dx, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O,
                     int64_t KH, int64_t KW, int64_t pad, int accumulate, cudaStream_t st);
int conv_igemm_wgrad(const float* ug, const float* x, float* dw, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O,
                     int64_t KH, int64_t KW, int64_t pad, int accumulate, cudaStream_t st);

// conv_small_mma.cu: TF32 mma.sync kernels for small-channel layers (fprop mode 0 / dgrad mode 1, wgrad)
int conv_mma_gather(const float* in, const float* w, const float* bias, float* out, int N, int C, int H, int W, int O, int KH,
                    int KW, int pad, int stride, int Ho, int Wo, int mode, int accumulate, cudaStream_t st, bool* done,
                    uint8_t* pool_idx = nullptr);
int wgrad_layout_for_tests(int C, int H, int W, int KH, int KW, int pad, int stride, int* out);
bool conv_sparse_serves(int C, int H, int W, int O, int KH, int KW, int pad, int Ho, int Wo);
bool conv_fused_fwd_serves(int C, int H, int W, int O, int KH, int KW, int pad, int stride, int Ho, int Wo);
int conv_fused_relu_pool_fwd(const float* x, const float* w, const float* bias, float* y, uint8_t* idx, int N, int C, int H, int W,
                             int O, int KH, int KW, int pad, int stride, int Ho, int Wo, cudaStream_t st, bool* done);
int conv_sparse_wbgrad_pooled(const float* ug, const float* ypool, const uint8_t* pidx, const float* z, const float* x, float* dw,
                              float* db, int N, int C, int H, int W, int O, int KH, int KW, int pad, int stride, int Ho, int Wo,
                              int acc_w, int acc_b, cudaStream_t st, bool* done);
int conv_mma_wgrad(const float* ug, const float* x, float* dw, int N, int C, int H, int W, int O, int KH, int KW, int pad,
                   int stride, int Ho, int Wo, int accumulate, cudaStream_t st, bool* done, const uint8_t* pidx = nullptr,
                   const float* ypool = nullptr, const float* zmask = nullptr);

int conv_mma_dgrad_rows(const float* ug, const float* w, float* dx, int N, int C, int H, int W, int O, int KH, int KW, int pad,
                        int stride, int Ho, int Wo, int accumulate, cudaStream_t st, bool* done, const uint8_t* pidx = nullptr,
                        const float* ypool = nullptr, const float* zmask = nullptr);

constexpr int OT = 8;  // output channels per thread (register tile)

struct ConvGeo {
  int N, C, H, W, O, KH, KW, pad, stride, Ho, Wo;
};

// ---------------------------------------------------------------------------------------------- fprop
// CTA = IMGS images x (TP x TQ) output positions x OT output channels; 256 threads, one position each.
// smem: patchT[img][c][pw][ph+1] (input patch, transposed so that p is contiguous) and wts[c][kh][kw][OT].
__global__ void __launch_bounds__(256) conv_fprop_direct_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ b, float* __restrict__ y,
                                                                ConvGeo g, int TP, int TQ, int IMGS, int CC) {
  extern __shared__ float sm[];
  const int PH = (TP - 1) * g.stride + g.KH, PW = (TQ - 1) * g.stride + g.KW;
  const int PHp = PH | 1;
  float* patch = sm;                                        // IMGS*CC*PW*PHp
  float* wts = sm + (((size_t)IMGS * CC * PW * PHp + 3) & ~size_t(3));   // CC*KH*KW*OT, 16B aligned
  const int tiles_p = ceil_div(g.Ho, TP), tiles_q = ceil_div(g.Wo, TQ);
  const int o_tiles = ceil_div(g.O, OT);
  int bid = blockIdx.x;
  const int tq = bid % tiles_q; bid /= tiles_q;
  const int tp = bid % tiles_p; bid /= tiles_p;
  const int ot = bid % o_tiles; bid /= o_tiles;
  const int n0 = bid * IMGS;
  const int p0 = tp * TP, q0 = tq * TQ, o0 = ot * OT;

  const int t = threadIdx.x;
  const int per_img = TP * TQ;
  const int img = t / per_img, rem = t - img * per_img;
  const int lq = rem / TP, lp = rem - lq * TP;
  const bool active = img < IMGS && (n0 + img) < g.N && (p0 + lp) < g.Ho && (q0 + lq) < g.Wo;

  float acc[OT];
#pragma unroll
  for (int o = 0; o < OT; ++o) acc[o] = 0.f;

  const int h_base = p0 * g.stride - g.pad, w_base = q0 * g.stride - g.pad;
  const int kk = g.KH * g.KW;

  for (int c0 = 0; c0 < g.C; c0 += CC) {
    const int cc = min(CC, g.C - c0);
    __syncthreads();
    // input patch: coalesced along W, stored transposed
    const int patch_elems = IMGS * cc * PH * PW;
    for (int e = t; e < patch_elems; e += 256) {
      int r = e;
      const int pw = r % PW; r /= PW;
      const int ph = r % PH; r /= PH;
      const int c = r % cc; r /= cc;
      const int im = r;
      const int n = n0 + im, h = h_base + ph, ww = w_base + pw;
      float v = 0.f;
      if (n < g.N && h >= 0 && h < g.H && ww >= 0 && ww < g.W)
        v = __ldg(x + (((int64_t)n * g.C + c0 + c) * g.H + h) * g.W + ww);
      patch[((size_t)(im * CC + c) * PW + pw) * PHp + ph] = v;
    }
    // weights for this (o-tile, c-chunk): wts[c][kh*KW+kw][o]
    const int w_elems = cc * kk * OT;
    for (int e = t; e < w_elems; e += 256) {
      const int o = e % OT;
      const int r = e / OT;            // c*kk + tap
      const int c = r / kk, tap = r - c * kk;
      float v = 0.f;
      if (o0 + o < g.O) v = __ldg(w + ((int64_t)(o0 + o) * g.C + c0 + c) * kk + tap);
      wts[e] = v;
    }
    __syncthreads();
    if (active) {
      for (int c = 0; c < cc; ++c) {
        const float* pc = patch + (size_t)(img * CC + c) * PW * PHp;
        const float* wc = wts + (size_t)c * kk * OT;
        for (int kh = 0; kh < g.KH; ++kh) {
          for (int kw = 0; kw < g.KW; ++kw) {
            const float xv = pc[(size_t)(lq * g.stride + kw) * PHp + lp * g.stride + kh];
            const float4 w0 = *reinterpret_cast<const float4*>(wc + (kh * g.KW + kw) * OT);
            const float4 w1 = *reinterpret_cast<const float4*>(wc + (kh * g.KW + kw) * OT + 4);
            acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
            acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
            acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
            acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
          }
        }
      }
    }
  }
  if (active) {
    const int n = n0 + img, p = p0 + lp, q = q0 + lq;
#pragma unroll
    for (int o = 0; o < OT; ++o) {
      if (o0 + o < g.O)
        y[((int64_t)n * g.O + o0 + o) * g.Ho * g.Wo + (int64_t)q * g.Ho + p] = acc[o] + __ldg(b + o0 + o);
    }
  }
}

// ---------------------------------------------------------------------------------------------- dgrad
// dx[n,c,h,w] = sum_{o,kh,kw} ug[n,o,p,q] * w[o,c,kh,kw] with p*s + kh - pad = h, q*s + kw - pad = w.
// CTA = one image x (TH x TW) input cells x OT input channels ("c tile"); the ug patch that can reach the tile is
// staged in smem as ugp[o][q][p] (p contiguous, as in global memory), weights as wts[o][kh][kw][ct].
__global__ void __launch_bounds__(256) conv_dgrad_direct_kernel(const float* __restrict__ ug, const float* __restrict__ w,
                                                                float* __restrict__ dx, ConvGeo g, int TH, int TW, int OC,
                                                                int accumulate) {
  extern __shared__ float sm[];
  const int tiles_h = ceil_div(g.H, TH), tiles_w = ceil_div(g.W, TW), c_tiles = ceil_div(g.C, OT);
  int bid = blockIdx.x;
  const int tw = bid % tiles_w; bid /= tiles_w;
  const int th = bid % tiles_h; bid /= tiles_h;
  const int ct = bid % c_tiles; bid /= c_tiles;
  const int n = bid;
  const int h0 = th * TH, w0 = tw * TW, c0 = ct * OT;
  // output positions that can touch rows [h0, h0+TH): p*s <= h+pad <= p*s+KH-1
  const int s = g.stride;
  int plo = (h0 + g.pad - (g.KH - 1) + s - 1) / s; if (h0 + g.pad - (g.KH - 1) < 0) plo = 0;
  int phi = (h0 + TH - 1 + g.pad) / s; if (phi > g.Ho - 1) phi = g.Ho - 1;
  int qlo = (w0 + g.pad - (g.KW - 1) + s - 1) / s; if (w0 + g.pad - (g.KW - 1) < 0) qlo = 0;
  int qhi = (w0 + TW - 1 + g.pad) / s; if (qhi > g.Wo - 1) qhi = g.Wo - 1;
  const int PP = max(phi - plo + 1, 0), PQ = max(qhi - qlo + 1, 0);
  const int PPmax = (TH + g.KH - 2) / s + 2, PQmax = (TW + g.KW - 2) / s + 2;  // smem carve-up is geometry-independent
  const int PPp = PPmax | 1;
  float* ugp = sm;                                   // OC * PQmax * PPp
  float* wts = sm + (((size_t)OC * PQmax * PPp + 3) & ~size_t(3));   // OC * KH*KW * OT, 16B aligned
  const int kk = g.KH * g.KW;

  const int t = threadIdx.x;
  const int lh = t / TW, lw = t - lh * TW;
  const int h = h0 + lh, ww = w0 + lw;
  const bool active = lh < TH && h < g.H && ww < g.W;
  float acc[OT];
#pragma unroll
  for (int c = 0; c < OT; ++c) acc[c] = 0.f;

  for (int ob = 0; ob < g.O; ob += OC) {
    const int oc = min(OC, g.O - ob);
    __syncthreads();
    const int pe = oc * PQ * PP;
    for (int e = t; e < pe; e += 256) {
      int r = e;
      const int lp = r % PP; r /= PP;
      const int lq = r % PQ; r /= PQ;
      const int o = r;
      ugp[((size_t)o * PQmax + lq) * PPp + lp] =
          __ldg(ug + ((int64_t)n * g.O + ob + o) * g.Ho * g.Wo + (int64_t)(qlo + lq) * g.Ho + plo + lp);
    }
    const int we = oc * kk * OT;
    for (int e = t; e < we; e += 256) {
      const int c = e % OT;
      const int r = e / OT;
      const int o = r / kk, tap = r - o * kk;
      float v = 0.f;
      if (c0 + c < g.C) v = __ldg(w + ((int64_t)(ob + o) * g.C + c0 + c) * kk + tap);
      wts[e] = v;
    }
    __syncthreads();
    if (active) {
      for (int kh = 0; kh < g.KH; ++kh) {
        const int hn = h + g.pad - kh;
        if (hn < 0 || hn % s) continue;
        const int p = hn / s;
        if (p < plo || p > phi) continue;
        for (int kw = 0; kw < g.KW; ++kw) {
          const int wn = ww + g.pad - kw;
          if (wn < 0 || wn % s) continue;
          const int q = wn / s;
          if (q < qlo || q > qhi) continue;
          const int tap = kh * g.KW + kw;
          for (int o = 0; o < oc; ++o) {
            const float gv = ugp[((size_t)o * PQmax + (q - qlo)) * PPp + (p - plo)];
            const float4 wa = *reinterpret_cast<const float4*>(wts + ((size_t)o * kk + tap) * OT);
            const float4 wb = *reinterpret_cast<const float4*>(wts + ((size_t)o * kk + tap) * OT + 4);
            acc[0] = fmaf(gv, wa.x, acc[0]); acc[1] = fmaf(gv, wa.y, acc[1]);
            acc[2] = fmaf(gv, wa.z, acc[2]); acc[3] = fmaf(gv, wa.w, acc[3]);
            acc[4] = fmaf(gv, wb.x, acc[4]); acc[5] = fmaf(gv, wb.y, acc[5]);
            acc[6] = fmaf(gv, wb.z, acc[6]); acc[7] = fmaf(gv, wb.w, acc[7]);
          }
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int c = 0; c < OT; ++c) {
      if (c0 + c < g.C) {
        float* d = dx + (((int64_t)n * g.C + c0 + c) * g.H + h) * g.W + ww;
        *d = accumulate ? *d + acc[c] : acc[c];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- wgrad
// dw[o,c,kh,kw] = sum_{n,p,q} ug[n,o,p,q] * xpad[n,c,p*s+kh,q*s+kw].  One thread per filter element, gridDim.y
// splits the batch; per-split partials are summed by a deterministic second pass.
__global__ void __launch_bounds__(256) conv_wgrad_direct_kernel(const float* __restrict__ ug, const float* __restrict__ x,
                                                                float* __restrict__ part, ConvGeo g, int imgs_per_split) {
  const int kk = g.KH * g.KW;
  const int64_t nw = (int64_t)g.O * g.C * kk;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nw) return;
  const int tap = (int)(e % kk);
  const int c = (int)((e / kk) % g.C);
  const int o = (int)(e / ((int64_t)kk * g.C));
  const int kh = tap / g.KW, kw = tap - kh * g.KW;
  const int nb = blockIdx.y * imgs_per_split, ne = min(g.N, nb + imgs_per_split);
  // valid output range for this tap (input index must fall inside the unpadded image)
  const int s = g.stride;
  int p_lo = 0, p_hi = g.Ho - 1, q_lo = 0, q_hi = g.Wo - 1;
  while (p_lo <= p_hi && p_lo * s + kh - g.pad < 0) ++p_lo;
  while (p_hi >= p_lo && p_hi * s + kh - g.pad >= g.H) --p_hi;
  while (q_lo <= q_hi && q_lo * s + kw - g.pad < 0) ++q_lo;
  while (q_hi >= q_lo && q_hi * s + kw - g.pad >= g.W) --q_hi;
  float acc = 0.f;
  for (int n = nb; n < ne; ++n) {
    const float* gn = ug + ((int64_t)n * g.O + o) * g.Ho * g.Wo;
    const float* xn = x + ((int64_t)n * g.C + c) * g.H * g.W;
    for (int q = q_lo; q <= q_hi; ++q) {
      const float* gq = gn + (int64_t)q * g.Ho;
      const float* xq = xn + (q * s + kw - g.pad);
      for (int p = p_lo; p <= p_hi; ++p) acc = fmaf(__ldg(gq + p), __ldg(xq + (int64_t)(p * s + kh - g.pad) * g.W), acc);
    }
  }
  part[(int64_t)blockIdx.y * nw + e] = acc;
}



// ---------------------------------------------------------------------------------------------- fprop / dgrad, small layers
// Register-tiled kernels for the LeNet / digits layers (stride 1, a few channels, images of a few KB).  A CTA stages G
// images (zero-padded) and the whole filter bank in smem; each thread owns 4 output channels x 8 consecutive outputs along
// the OUTPUT's contiguous axis and slides an 8+K-1 window of the input through registers, so the inner step is one
// broadcast 128-bit weight load per tap for 32 FMAs (the generic kernels above issue 3 smem loads per 8 FMAs and are
// LSU-bound).  Forward: outputs run along p (neograd order), window slides along kh.  Data gradient: outputs run along
// w, window slides along kw over the un-scrambled upstream gradient.
constexpr int RT = 8;   // outputs per thread along the contiguous axis
// CT = output channels per thread: 4 for multi-channel layers, 1 for neograd's single-filter Conv2D (conv.py:131-152),
// where a 4-wide channel tile would spend three quarters of its FMAs and weight loads on padding

struct SmallGeo {
  int N, Cin, Cout;      // input / output channels of THIS pass (dgrad: Cin = O, Cout = C)
  int A, Bx;             // input plane: A rows x Bx contiguous columns AS STAGED (zero padded)
  int JA, JB;            // output plane extents: JB contiguous
  int KA, KB;            // filter extents along the staged input's row / column axes
  int G, tpi;            // images per CTA pass, threads per image
  int jb_blocks, co_groups;
  int in_plane, in_rows, in_cols, pad_a, pad_b;   // raw input plane (elements), its extents and the padding applied in smem
  int transpose_in;      // dgrad: the raw input is ug in neograd order [o][q][p]; staged as [o][p][q]
};

// out[n][co][ja][jb] (jb contiguous).  fprop: (ja, jb) = (q, p); staged input rows = w, cols = h?  -- see the two wrappers:
// the kernel itself only knows: out[ja][jb] = sum_{ci, ka, kb} in_s[ci][ja + ka][jb + kb] * wts[ci][ka][kb][co]
// where in_s is the staged plane with its CONTIGUOUS axis along jb.
template <int CT>
__global__ void __launch_bounds__(256) conv_small_kernel(const float* __restrict__ in, const float* __restrict__ wts_g,
                                                         const float* __restrict__ bias, float* __restrict__ out, SmallGeo g,
                                                         int accumulate, FastDiv d_cols, FastDiv d_plane, FastDiv d_tpi,
                                                         FastDiv d_jbb, FastDiv d_ja) {
  extern __shared__ float sm[];
  const int plane_s = g.A * g.Bx + 1;                       // staged plane stride (odd-ish: spreads banks)
  const int wsize = g.Cin * g.KA * g.KB * g.co_groups * CT;
  float* sw = sm;                                           // [ci][ka][kb][co (padded to groups of 4)]
  float* sx = sm + ((wsize + 3) & ~3);                      // [G][ci][A][Bx]
  const int t = threadIdx.x;
  for (int e = t; e < wsize; e += 256) sw[e] = __ldg(wts_g + e);
  const int64_t passes = ceil_div<int64_t>(g.N, g.G);
  for (int64_t pass = blockIdx.x; pass < passes; pass += gridDim.x) {
    const int n0 = (int)(pass * g.G);
    const int ng = min(g.G, g.N - n0);
    __syncthreads();
    // ---- stage: zero fill (padding), then the raw planes
    const int stage_elems = ng * g.Cin * plane_s;
    if (g.pad_a | g.pad_b) {
      for (int e = t; e < stage_elems; e += 256) sx[e] = 0.f;
      __syncthreads();
    }
    const int raw = ng * g.Cin * g.in_plane;
    const float* src = in + (int64_t)n0 * g.Cin * g.in_plane;
    // (the staging is paid in issue slots: 128-bit loads and one index decomposition per four elements when rows allow)
    const bool vec4 = (g.in_cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
    if (vec4) {
#pragma unroll 2
      for (int e4 = t; e4 < (raw >> 2); e4 += 256) {
        uint32_t pl, f, r, c;
        d_plane.divmod((uint32_t)e4 * 4u, pl, f);          // pl = image*Cin + ci
        d_cols.divmod(f, r, c);                              // raw plane is [in_rows][in_cols], cols contiguous
        const float4 v = ld_stream4(src + 4 * (int64_t)e4);
        if (g.transpose_in) {                                // [q][p] -> staged [p][q]
          float* d = sx + pl * plane_s + (c + g.pad_a) * g.Bx + r + g.pad_b;
          d[0] = v.x; d[g.Bx] = v.y; d[2 * g.Bx] = v.z; d[3 * g.Bx] = v.w;
        } else {
          float* d = sx + pl * plane_s + (r + g.pad_a) * g.Bx + c + g.pad_b;
          d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
      }
    } else {
#pragma unroll 4
      for (int e = t; e < raw; e += 256) {
        uint32_t pl, f, r, c;
        d_plane.divmod(e, pl, f);          // pl = image*Cin + ci
        d_cols.divmod(f, r, c);            // raw plane is [in_rows][in_cols], cols contiguous
        const float v = __ldg(src + e);
        if (g.transpose_in) sx[pl * plane_s + (c + g.pad_a) * g.Bx + r + g.pad_b] = v;   // [q][p] -> staged [p][q]
        else sx[pl * plane_s + (r + g.pad_a) * g.Bx + c + g.pad_b] = v;
      }
    }
    __syncthreads();
    // ---- compute: thread <-> (image, co group, ja, jb block)
    uint32_t img, rem;
    d_tpi.divmod(t, img, rem);
    if ((int)img < ng) {
      uint32_t r2, jbb, cg, ja;
      d_jbb.divmod(rem, r2, jbb);
      d_ja.divmod(r2, cg, ja);
      const int jb0 = jbb * RT;
      float acc[CT][RT];
#pragma unroll
      for (int c = 0; c < CT; ++c)
#pragma unroll
        for (int i = 0; i < RT; ++i) acc[c][i] = 0.f;
      const float* xi = sx + (size_t)img * g.Cin * plane_s;
      for (int ci = 0; ci < g.Cin; ++ci) {
        for (int ka = 0; ka < g.KA; ++ka) {
          const float* row = xi + ci * plane_s + (ja + ka) * g.Bx + jb0;
          const float* wr = sw + ((ci * g.KA + ka) * g.KB) * g.co_groups * CT + cg * CT;
          float win[RT];
#pragma unroll
          for (int i = 0; i < RT - 1; ++i) win[i + 1] = row[i];
          for (int kb = 0; kb < g.KB; ++kb) {
#pragma unroll
            for (int i = 0; i < RT - 1; ++i) win[i] = win[i + 1];
            win[RT - 1] = row[kb + RT - 1];
            if (CT == 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(wr + kb * g.co_groups * CT);
#pragma unroll
              for (int i = 0; i < RT; ++i) {
                acc[0][i] = fmaf(win[i], w4.x, acc[0][i]);
                acc[1 % CT][i] = fmaf(win[i], w4.y, acc[1 % CT][i]);
                acc[2 % CT][i] = fmaf(win[i], w4.z, acc[2 % CT][i]);
                acc[3 % CT][i] = fmaf(win[i], w4.w, acc[3 % CT][i]);
              }
            } else {
              const float w1 = wr[kb * g.co_groups * CT];
#pragma unroll
              for (int i = 0; i < RT; ++i) acc[0][i] = fmaf(win[i], w1, acc[0][i]);
            }
          }
        }
      }
      // ---- store: RT contiguous outputs per channel
      const int n = n0 + img;
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const int co = cg * CT + c;
        if (co < g.Cout) {
          const float bv = bias ? __ldg(bias + co) : 0.f;
          float* o = out + (((int64_t)n * g.Cout + co) * g.JA + ja) * g.JB + jb0;
#pragma unroll
          for (int i = 0; i < RT; ++i) {
            if (jb0 + i < g.JB) o[i] = accumulate ? o[i] + acc[c][i] + bv : acc[c][i] + bv;
          }
        }
      }
    }
  }
}

// weights -> [ci][ka][kb][co padded]:  fprop  ci = c, (ka, kb) = (kw, kh), co = o          (out (ja,jb) = (q,p): rows w, cols h)
//                                      dgrad  ci = o, (ka, kb) = (KH-1-kh, KW-1-kw), co = c (out (ja,jb) = (h,w): rows p, cols q)
__global__ void __launch_bounds__(256) pack_small_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int O, int C,
                                                                 int KH, int KW, int dgrad, int cop) {
  const int Cin = dgrad ? O : C, KA = dgrad ? KH : KW, KB = dgrad ? KW : KH;
  const int total = Cin * KA * KB * cop;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % cop;
    const int kb = (i / cop) % KB;
    const int ka = (i / (cop * KB)) % KA;
    const int ci = i / (cop * KB * KA);
    float v = 0.f;
    if (!dgrad) {
      if (co < O) v = w[(((int64_t)co * C + ci) * KH + kb) * KW + ka];            // o = co, c = ci, kh = kb, kw = ka
    } else {
      if (co < C) v = w[(((int64_t)ci * C + co) * KH + (KH - 1 - ka)) * KW + (KW - 1 - kb)];   // o = ci, c = co
    }
    wp[i] = v;
  }
}

// fprop: staged input plane is x transposed?  No: fprop needs the window to slide along p (= h), so the staged plane keeps
// h CONTIGUOUS: x[c][h][w] is staged as [w][h] (transpose_in), rows = w (ja = q, taps kw), cols = h (jb = p, taps kh).
// dgrad: ug [o][q][p] (p contiguous) must slide along q (= w): staged as [p][q] (transpose_in), rows = p (ja = h), cols = q.
static int conv_small(const float* in, const float* w, const float* bias, float* out, const ConvGeo& cg, int dgrad, int accumulate,
                      cudaStream_t st, bool* done) {
  *done = false;
  if (cg.stride != 1) return NGB_OK;
  SmallGeo g;
  g.N = cg.N;
  int padA, padB;
  if (!dgrad) {
    if (cg.pad > cg.KH - 1 && cg.pad > cg.KW - 1) { /* fine: zero halo is staged */ }
    g.Cin = cg.C; g.Cout = cg.O;
    g.in_rows = cg.H; g.in_cols = cg.W;                 // raw x plane [h][w]
    g.KA = cg.KW; g.KB = cg.KH;
    padA = cg.pad; padB = cg.pad;
    g.A = cg.W + 2 * padA; g.Bx = cg.H + 2 * padB;      // staged [w][h]
    g.JA = cg.Wo; g.JB = cg.Ho;
  } else {
    if (cg.pad > cg.KH - 1 || cg.pad > cg.KW - 1) return NGB_OK;
    g.Cin = cg.O; g.Cout = cg.C;
    g.in_rows = cg.Wo; g.in_cols = cg.Ho;               // raw ug plane [q][p]
    g.KA = cg.KH; g.KB = cg.KW;
    padA = cg.KH - 1 - cg.pad; padB = cg.KW - 1 - cg.pad;
    g.A = cg.Ho + 2 * padA; g.Bx = cg.Wo + 2 * padB;    // staged [p][q]
    g.JA = cg.H; g.JB = cg.W;
  }
  g.transpose_in = 1;
  g.pad_a = padA; g.pad_b = padB;
  g.in_plane = g.in_rows * g.in_cols;
  const int CT = g.Cout == 1 ? 1 : 4;
  g.co_groups = ceil_div(g.Cout, CT);
  g.jb_blocks = ceil_div(g.JB, RT);
  // the sliding window reads RT + KB - 1 columns from jb0: keep that inside the staged row
  const int need_cols = g.jb_blocks * RT + g.KB - 1;
  if (need_cols > g.Bx) g.Bx = need_cols;
  g.tpi = g.co_groups * g.JA * g.jb_blocks;
  if (g.tpi > 256 || g.Cin > 64) return NGB_OK;
  g.G = 256 / g.tpi;
  const int plane_s = g.A * g.Bx + 1;
  const int wsize = g.Cin * g.KA * g.KB * g.co_groups * CT;
  while (g.G > 1 && ((size_t)((wsize + 3) & ~3) + (size_t)g.G * g.Cin * plane_s) * 4 > 44 * 1024) --g.G;
  if (g.G > g.N) g.G = g.N > 0 ? g.N : 1;
  const size_t smem = ((size_t)((wsize + 3) & ~3) + (size_t)g.G * g.Cin * plane_s) * 4;
  if (smem > 46 * 1024 || (int64_t)wsize * 4 > (1 << 20)) return NGB_OK;
  int rc = ensure_init();
  if (rc) return rc;
  // packed weights live at the END of the scratch arena (the front is used by split reductions)
  float* wp = scratch_ptr() + (scratch_bytes() / 4 - (1 << 18));
  pack_small_weights_kernel<<<ceil_div(wsize, 256), 256, 0, st>>>(w, wp, cg.O, cg.C, cg.KH, cg.KW, dgrad, g.co_groups * CT);
  NGB_LAUNCH_CHECK();
  const int64_t passes = ceil_div<int64_t>(g.N, g.G);
  const int grid = (int)(passes < (int64_t)kNumSMs * 4 ? passes : (int64_t)kNumSMs * 4);
  if (CT == 1)
    conv_small_kernel<1><<<grid, 256, smem, st>>>(in, wp, bias, out, g, accumulate, FastDiv((uint32_t)g.in_cols),
                                                 FastDiv((uint32_t)g.in_plane), FastDiv((uint32_t)g.tpi),
                                                 FastDiv((uint32_t)g.jb_blocks), FastDiv((uint32_t)g.JA));
  else
    conv_small_kernel<4><<<grid, 256, smem, st>>>(in, wp, bias, out, g, accumulate, FastDiv((uint32_t)g.in_cols),
                                                 FastDiv((uint32_t)g.in_plane), FastDiv((uint32_t)g.tpi),
                                                 FastDiv((uint32_t)g.jb_blocks), FastDiv((uint32_t)g.JA));
  NGB_LAUNCH_CHECK();
  *done = true;
  return NGB_OK;
}


// ---------------------------------------------------------------------------------------------- dgrad, small layers
// dx[n][c][h][w] = sum_{o,kh,kw} ug[n][o][p = h+pad-kh][q = w+pad-kw] * w[o][c][kh][kw]   (conv.py:299-309)
// Thread <-> (image, h, block of DW consecutive w) and ALL input channels c (<= 8) in registers.  ug is staged
// un-scrambled as [o][p][q] with a zero halo of KW-1 columns only; rows p outside the image are skipped (no work on the
// vertical halo), and a DW+KW-1 register window slides along kw.  Weights sit in smem as [o][kh][kw][8 c] (two broadcast
// 128-bit loads per tap for 8*DW FMAs).
constexpr int DW = 4;
template <int KW>
__global__ void __launch_bounds__(256) conv_dgrad_small_kernel(const float* __restrict__ ug, const float* __restrict__ wts_g,
                                                               float* __restrict__ dx, ConvGeo g, int G, int tpi, int wblocks,
                                                               int accumulate, FastDiv d_howo, FastDiv d_ho, FastDiv d_tpi,
                                                               FastDiv d_wb) {
  extern __shared__ float sm[];
  const int halo = KW - 1;
  const int Bq = g.Wo + 2 * halo;                      // staged row: [halo | q = 0..Wo-1 | halo]
  const int plane_s = g.Ho * Bq + 1;
  const int wsize = g.O * g.KH * KW * 8;
  float* sw = sm;                                      // [o][kh][kw][8]
  float* su = sm + wsize;                              // [G][o][Ho][Bq]
  const int t = threadIdx.x;
  for (int e = t; e < wsize; e += 256) sw[e] = __ldg(wts_g + e);
  const int HoWo = g.Ho * g.Wo;
  const int64_t passes = ceil_div<int64_t>(g.N, G);
  for (int64_t pass = blockIdx.x; pass < passes; pass += gridDim.x) {
    const int n0 = (int)(pass * G);
    const int ng = min(G, g.N - n0);
    __syncthreads();
    for (int e = t; e < ng * g.O * plane_s; e += 256) su[e] = 0.f;
    __syncthreads();
    const float* src = ug + (int64_t)n0 * g.O * HoWo;
    const int raw = ng * g.O * HoWo;
#pragma unroll 4
    for (int e = t; e < raw; e += 256) {                 // global [o][q][p] -> smem [o][p][halo + q]
      uint32_t pl, f, q, p;
      d_howo.divmod(e, pl, f);
      d_ho.divmod(f, q, p);
      su[pl * plane_s + p * Bq + halo + q] = __ldg(src + e);
    }
    __syncthreads();
    uint32_t img, rem;
    d_tpi.divmod(t, img, rem);
    if ((int)img < ng) {
      uint32_t h, wb;
      d_wb.divmod(rem, h, wb);
      const int w0 = wb * DW;
      float acc[8][DW];
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int i = 0; i < DW; ++i) acc[c][i] = 0.f;
      // dx[h][w0+i] needs ug[p][q = w0 + i + pad - kw]; with j = KW-1-kw the staged column is (w0 + pad) + i + j
      const int col0 = w0 + g.pad;
      int kh_lo = (int)h + g.pad - (g.Ho - 1); if (kh_lo < 0) kh_lo = 0;
      int kh_hi = (int)h + g.pad; if (kh_hi > g.KH - 1) kh_hi = g.KH - 1;
      for (int o = 0; o < g.O; ++o) {
        const float* up = su + ((size_t)img * g.O + o) * plane_s;
        for (int kh = kh_lo; kh <= kh_hi; ++kh) {
          const int p = (int)h + g.pad - kh;
          const float* row = up + p * Bq + col0;
          const float* wr = sw + ((o * g.KH + kh) * KW) * 8;
          float win[DW + KW - 1];
#pragma unroll
          for (int i = 0; i < DW + KW - 1; ++i) win[i] = row[i];
#pragma unroll
          for (int j = 0; j < KW; ++j) {                  // j = KW-1-kw
            const float4 wa = *reinterpret_cast<const float4*>(wr + (KW - 1 - j) * 8);
            const float4 wb4 = *reinterpret_cast<const float4*>(wr + (KW - 1 - j) * 8 + 4);
#pragma unroll
            for (int i = 0; i < DW; ++i) {
              const float v = win[i + j];
              acc[0][i] = fmaf(v, wa.x, acc[0][i]); acc[1][i] = fmaf(v, wa.y, acc[1][i]);
              acc[2][i] = fmaf(v, wa.z, acc[2][i]); acc[3][i] = fmaf(v, wa.w, acc[3][i]);
              acc[4][i] = fmaf(v, wb4.x, acc[4][i]); acc[5][i] = fmaf(v, wb4.y, acc[5][i]);
              acc[6][i] = fmaf(v, wb4.z, acc[6][i]); acc[7][i] = fmaf(v, wb4.w, acc[7][i]);
            }
          }
        }
      }
      const int n = n0 + img;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < g.C) {
          float* d = dx + (((int64_t)n * g.C + c) * g.H + h) * g.W + w0;
#pragma unroll
          for (int i = 0; i < DW; ++i)
            if (w0 + i < g.W) d[i] = accumulate ? d[i] + acc[c][i] : acc[c][i];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) pack_dgrad_small_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int O,
                                                                       int C, int KH, int KW) {
  const int total = O * KH * KW * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i & 7;
    const int tap = (i >> 3) % (KH * KW);
    const int o = (i >> 3) / (KH * KW);
    wp[i] = c < C ? w[((int64_t)o * C + c) * KH * KW + tap] : 0.f;
  }
}

static int conv_dgrad_small(const float* ug, const float* w, float* dx, const ConvGeo& g, int accumulate, cudaStream_t st, bool* done) {
  *done = false;
  if (g.stride != 1 || g.C > 8 || g.O > 64 || (g.KW != 3 && g.KW != 5)) return NGB_OK;
  if (g.pad > g.KW - 1) return NGB_OK;                  // the staged column halo is KW-1 wide
  const int wblocks = ceil_div(g.W, DW);
  const int tpi = g.H * wblocks;
  if (tpi > 256) return NGB_OK;
  // every window read must stay inside the staged row: col0 + DW + KW - 2 <= Bq - 1 for the last block
  const int Bq = g.Wo + 2 * (g.KW - 1);
  if ((wblocks - 1) * DW + g.pad + DW + g.KW - 2 > Bq - 1) return NGB_OK;
  int G = 256 / tpi;
  const int plane_s = g.Ho * Bq + 1;
  const int wsize = g.O * g.KH * g.KW * 8;
  while (G > 1 && ((size_t)wsize + (size_t)G * g.O * plane_s) * 4 > 44 * 1024) --G;
  if (G > g.N) G = g.N > 0 ? g.N : 1;
  const size_t smem = ((size_t)wsize + (size_t)G * g.O * plane_s) * 4;
  if (smem > 46 * 1024 || wsize * 4 > (1 << 20)) return NGB_OK;
  int rc = ensure_init();
  if (rc) return rc;
  float* wp = scratch_ptr() + (scratch_bytes() / 4 - (1 << 18));
  pack_dgrad_small_weights_kernel<<<ceil_div(wsize, 256), 256, 0, st>>>(w, wp, g.O, g.C, g.KH, g.KW);
  NGB_LAUNCH_CHECK();
  const int64_t passes = ceil_div<int64_t>(g.N, G);
  const int grid = (int)(passes < (int64_t)kNumSMs * 4 ? passes : (int64_t)kNumSMs * 4);
  const FastDiv dhowo((uint32_t)(g.Ho * g.Wo)), dho((uint32_t)g.Ho), dtpi((uint32_t)tpi), dwb((uint32_t)wblocks);
  if (g.KW == 3)
    conv_dgrad_small_kernel<3><<<grid, 256, smem, st>>>(ug, wp, dx, g, G, tpi, wblocks, accumulate, dhowo, dho, dtpi, dwb);
  else
    conv_dgrad_small_kernel<5><<<grid, 256, smem, st>>>(ug, wp, dx, g, G, tpi, wblocks, accumulate, dhowo, dho, dtpi, dwb);
  NGB_LAUNCH_CHECK();
  *done = true;
  return NGB_OK;
}

// ---------------------------------------------------------------------------------------------- wgrad, small layers
// The LeNet / digits layers (C*KH*O <= 512 filter rows, images of a few KB).  A CTA walks its share of the batch one image
// at a time: x[n] and ug[n] (un-scrambled to [o][p][q]) are staged in smem, thread <-> (filter row (o,c,kh), slice of
// output rows p) keeps the KW taps of its filter row in registers and slides a KW-wide window of x along q, so the
// inner step is 2 smem loads for KW FMAs.  Slices are reduced through smem at the end; per-CTA partials are summed by
// the deterministic second pass.  (The thread-per-filter-element kernel below ran at 0.7 % of its roofline here.)
template <int KW>
__global__ void __launch_bounds__(512) conv_wgrad_small_kernel(const float* __restrict__ ug, const float* __restrict__ x,
                                                               float* __restrict__ part, ConvGeo g, int rows_t, int slices,
                                                               int imgs_per_cta, FastDiv d_hw, FastDiv d_howo, FastDiv d_ho) {
  extern __shared__ float sm[];
  const int HW = g.H * g.W, HoWo = g.Ho * g.Wo;
  const int xs = HW + 1, us = HoWo + 1;            // padded plane strides (bank spread)
  float* sx = sm;                                   // [C][xs]
  float* su = sm + g.C * xs;                        // [O][us], natural (p,q) order
  const int t = threadIdx.x;
  const int row = t % rows_t, slice = t / rows_t;   // row = (o*C + c)*KH + kh
  const bool active = slice < slices;
  const int kh = row % g.KH;
  const int c = (row / g.KH) % g.C;
  const int o = row / (g.KH * g.C);
  float acc[KW];
#pragma unroll
  for (int j = 0; j < KW; ++j) acc[j] = 0.f;
  const int n0 = blockIdx.x * imgs_per_cta, n1 = min(g.N, n0 + imgs_per_cta);
  for (int n = n0; n < n1; ++n) {
    __syncthreads();
    const float* xn = x + (int64_t)n * g.C * HW;
#pragma unroll 4
    for (int e = t; e < g.C * HW; e += blockDim.x) {
      uint32_t cc, f;
      d_hw.divmod(e, cc, f);
      sx[cc * xs + f] = __ldg(xn + e);
    }
    const float* un = ug + (int64_t)n * g.O * HoWo;
#pragma unroll 4
    for (int e = t; e < g.O * HoWo; e += blockDim.x) {   // global (o, q, p) -> smem (o, p, q)
      uint32_t oo, f, q, p;
      d_howo.divmod(e, oo, f);
      d_ho.divmod(f, q, p);
      su[oo * us + p * g.Wo + q] = __ldg(un + e);
    }
    __syncthreads();
    if (active) {
      const float* xc = sx + c * xs;
      const float* uo = su + o * us;
      for (int p = slice; p < g.Ho; p += slices) {
        const int h = p * g.stride + kh - g.pad;
        if (h < 0 || h >= g.H) continue;
        const float* xr = xc + h * g.W;
        const float* ur = uo + p * g.Wo;
        if (g.stride == 1) {
          float win[KW];                                 // win[j] = x[h][q + j - pad]
#pragma unroll
          for (int j = 0; j < KW - 1; ++j) { const int w = j - g.pad; win[j + 1] = (w >= 0 && w < g.W) ? xr[w] : 0.f; }
          for (int q = 0; q < g.Wo; ++q) {
#pragma unroll
            for (int j = 0; j < KW - 1; ++j) win[j] = win[j + 1];
            const int w = q + KW - 1 - g.pad;
            win[KW - 1] = (w >= 0 && w < g.W) ? xr[w] : 0.f;
            const float gv = ur[q];
#pragma unroll
            for (int j = 0; j < KW; ++j) acc[j] = fmaf(gv, win[j], acc[j]);
          }
        } else {
          for (int q = 0; q < g.Wo; ++q) {
            const float gv = ur[q];
#pragma unroll
            for (int j = 0; j < KW; ++j) {
              const int w = q * g.stride + j - g.pad;
              if (w >= 0 && w < g.W) acc[j] = fmaf(gv, xr[w], acc[j]);
            }
          }
        }
      }
    }
  }
  // reduce the slices (fixed order -> deterministic) and emit this CTA's partial
  __syncthreads();
  float* red = sm;                                  // [slices][rows_t][KW]
  if (active) {
#pragma unroll
    for (int j = 0; j < KW; ++j) red[(slice * rows_t + row) * KW + j] = acc[j];
  }
  __syncthreads();
  for (int e = t; e < rows_t * KW; e += blockDim.x) {
    float s2 = 0.f;
    for (int sl = 0; sl < slices; ++sl) s2 += red[sl * rows_t * KW + e];
    part[(int64_t)blockIdx.x * rows_t * KW + e] = s2;   // e = ((o*C + c)*KH + kh)*KW + kw: dw's own layout
  }
}

// returns true when the small-layer kernel was launched
static int conv_wgrad_small(const float* ug, const float* x, float* dw, const ConvGeo& g, int accumulate, cudaStream_t st, bool* done) {
  *done = false;
  const int rows_t = g.O * g.C * g.KH;
  if (rows_t > 512 || (g.KW != 2 && g.KW != 3 && g.KW != 5 && g.KW != 7)) return NGB_OK;
  const int threads = rows_t > 256 ? 512 : 256;
  const int slices = threads / rows_t;
  const size_t smem_stage = ((size_t)g.C * (g.H * g.W + 1) + (size_t)g.O * (g.Ho * g.Wo + 1)) * 4;
  const size_t smem_red = (size_t)slices * rows_t * g.KW * 4;
  const size_t smem = smem_stage > smem_red ? smem_stage : smem_red;
  if (smem > 46 * 1024) return NGB_OK;
  int rc = ensure_init();
  if (rc) return rc;
  const int64_t nw = (int64_t)rows_t * g.KW;
  int64_t ctas = (int64_t)kNumSMs * (smem > 20 * 1024 ? 2 : 4);
  if (ctas > g.N) ctas = g.N;
  if (ctas * nw * 4 > scratch_bytes()) ctas = scratch_bytes() / (nw * 4);
  if (ctas < 1) return NGB_OK;
  const int ipc = (int)ceil_div<int64_t>(g.N, ctas);
  ctas = ceil_div<int64_t>(g.N, ipc);
  const FastDiv dhw((uint32_t)(g.H * g.W)), dhowo((uint32_t)(g.Ho * g.Wo)), dho((uint32_t)g.Ho);
  switch (g.KW) {
    case 2: conv_wgrad_small_kernel<2><<<(unsigned)ctas, threads, smem, st>>>(ug, x, scratch_ptr(), g, rows_t, slices, ipc, dhw, dhowo, dho); break;
    case 3: conv_wgrad_small_kernel<3><<<(unsigned)ctas, threads, smem, st>>>(ug, x, scratch_ptr(), g, rows_t, slices, ipc, dhw, dhowo, dho); break;
    case 5: conv_wgrad_small_kernel<5><<<(unsigned)ctas, threads, smem, st>>>(ug, x, scratch_ptr(), g, rows_t, slices, ipc, dhw, dhowo, dho); break;
    default: conv_wgrad_small_kernel<7><<<(unsigned)ctas, threads, smem, st>>>(ug, x, scratch_ptr(), g, rows_t, slices, ipc, dhw, dhowo, dho); break;
  }
  NGB_LAUNCH_CHECK();
  *done = true;
  return launch_sum_partials(scratch_ptr(), (int)ctas, nw, dw, accumulate, st);
}

// ---------------------------------------------------------------------------------------------- bgrad
// db[o] = sum_{n,f} ug[n,o,f].  grid = (O, splits over N); every image contributes one contiguous HoWo run: 128-bit loads,
// four in flight per thread, no 64-bit divides.
// MASKED: ug is the pooled gradient of a Conv -> ReLU -> MaxPool(2x2, s2) block and the sum runs over ug * relu'(z at the
// window's argmax) -- y (the pool's output) > 0 means mask 1, otherwise the mask is z_first >= 0 (see maxpool_relu_mask_kernel);
// the bias gradient then needs neither the full-resolution gradient nor the compact masked copy.
// FLAGGED: the mask comes from bit 2 of the pool's argmax bytes instead (ngb_maxpool_relu_fwd): ug and one byte per element.
template <bool MASKED, bool FLAGGED = false>
__global__ void __launch_bounds__(256) conv_bgrad_kernel(const float* __restrict__ ug, float* __restrict__ part, int64_t N,
                                                         int64_t O, int64_t HW, int64_t imgs_per_split, FastDiv d_hw4, int vec,
                                                         const float* __restrict__ y, const float* __restrict__ z, int Hp, int Wo,
                                                         FastDiv d_hp, const uint8_t* __restrict__ idx = nullptr) {
  __shared__ float red[32];
  const int64_t o = blockIdx.x;
  const int64_t nb = (int64_t)blockIdx.y * imgs_per_split, ne = min(N, nb + imgs_per_split);
  float acc = 0.f;
  auto masked = [&](float gv, float yv, int64_t plane, uint32_t f) {
    if (!(yv > 0.f)) {
      uint32_t q, p;
      d_hp.divmod(f, q, p);                         // pooled position (p', q') sits at q'*Hp + p'
      const float zv = __ldg(z + plane * (4 * HW) + (int64_t)(2 * p) * Wo + 2 * q);
      gv = zv >= 0.f ? gv : 0.f;
    }
    return gv;
  };
  if (vec) {
    const uint32_t hw4 = (uint32_t)(HW >> 2);
    const uint32_t span4 = (uint32_t)((ne - nb) * hw4);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (uint32_t i = threadIdx.x; i < span4; i += 256) {
      uint32_t n, f4;
      d_hw4.divmod(i, n, f4);
      const int64_t plane = (nb + n) * O + o;
      float4 v = ld_stream4(ug + plane * HW + 4 * (int64_t)f4);
      if (FLAGGED) {
        const uint32_t a4 = __ldg(reinterpret_cast<const uint32_t*>(idx + plane * HW + 4 * (int64_t)f4));
        v.x = (a4 & 0x4u) ? v.x : 0.f; v.y = (a4 & 0x400u) ? v.y : 0.f;
        v.z = (a4 & 0x40000u) ? v.z : 0.f; v.w = (a4 & 0x4000000u) ? v.w : 0.f;
      } else if (MASKED) {
        const float4 yv = ld_stream4(y + plane * HW + 4 * (int64_t)f4);
        v.x = masked(v.x, yv.x, plane, 4 * f4); v.y = masked(v.y, yv.y, plane, 4 * f4 + 1);
        v.z = masked(v.z, yv.z, plane, 4 * f4 + 2); v.w = masked(v.w, yv.w, plane, 4 * f4 + 3);
      }
      a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
    }
    acc = (a0 + a1) + (a2 + a3);
  } else {
    const int64_t span = (ne - nb) * HW;
    for (int64_t i = threadIdx.x; i < span; i += blockDim.x) {
      const int64_t n = nb + i / HW, f = i % HW;
      float v = ug[(n * O + o) * HW + f];
      if (FLAGGED) v = (idx[(n * O + o) * HW + f] & 4) ? v : 0.f;
      else if (MASKED) v = masked(v, y[(n * O + o) * HW + f], n * O + o, (uint32_t)f);
      acc += v;
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) part[(int64_t)blockIdx.y * O + o] = acc;
}

static int make_geo(ConvGeo* g, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW, int64_t pad,
                    int64_t stride) {
  NGB_CHECK_ARG(stride >= 1, "Stride must be greater than or equal to 1");
  NGB_CHECK_ARG(pad >= 0, "Padding must be greater than or equal to zero");
  NGB_CHECK_ARG(N >= 0 && C >= 1 && O >= 1 && KH >= 1 && KW >= 1 && H >= 1 && W >= 1, "conv: bad dimensions");
  NGB_CHECK_ARG(H + 2 * pad >= KH && W + 2 * pad >= KW, "conv: filter larger than the padded input");
  NGB_CHECK_ARG(N * C * H * W < (int64_t(1) << 40), "conv: tensor too large");
  g->N = (int)N; g->C = (int)C; g->H = (int)H; g->W = (int)W; g->O = (int)O; g->KH = (int)KH; g->KW = (int)KW;
  g->pad = (int)pad; g->stride = (int)stride;
  g->Ho = (int)((H + 2 * pad - KH) / stride + 1);
  g->Wo = (int)((W + 2 * pad - KW) / stride + 1);
  return NGB_OK;
}

constexpr int kDirectSmem = 96 * 1024;

}  // namespace ngb

using namespace ngb;

extern "C" int ngb_conv_fprop(const float* x, const float* w, const float* b, float* y, int64_t N, int64_t C, int64_t H,
                              int64_t W, int64_t O, int64_t KH, int64_t KW, int64_t pad, int64_t stride, ngb_stream_t stream) {
  ConvGeo g;
  int rc = make_geo(&g, N, C, H, W, O, KH, KW, pad, stride);
  if (rc) return rc;
  if (N == 0) return NGB_OK;
  cudaStream_t st = as_stream(stream);
  if (conv_igemm_eligible(N, C, H, W, O, KH, KW, pad, stride) && aligned16(x) && aligned16(y) && aligned16(w))
    return conv_igemm_fprop(x, w, b, y, N, C, H, W, O, KH, KW, pad, st);
  {
    bool done = false;
    rc = conv_mma_gather(x, w, b, y, g.N, g.C, g.H, g.W, g.O, g.KH, g.KW, g.pad, g.stride, g.Ho, g.Wo, 0, 0, st, &done);
    if (rc || done) return rc;
    rc = conv_small(x, w, b, y, g, 0, 0, st, &done);
    if (rc || done) return rc;
  }
  // tile shape: p (fastest output axis) first
  int TP = g.Ho < 16 ? g.Ho : 16;
  int TQ = 256 / TP; if (TQ > g.Wo) TQ = g.Wo;
  int IMGS = 256 / (TP * TQ); if (IMGS < 1) IMGS = 1; if (IMGS > g.N) IMGS = g.N;
  const int PH = (TP - 1) * g.stride + g.KH, PW = (TQ - 1) * g.stride + g.KW, PHp = PH | 1;
  const int64_t per_c = (int64_t)IMGS * PW * PHp + (int64_t)g.KH * g.KW * OT;
  int CC = (int)((kDirectSmem / 4 - 4) / per_c);
  if (CC < 1) return fail(NGB_ERR_UNSUPPORTED, "conv fprop: filter/stride too large for the direct kernel tile");
  if (CC > g.C) CC = g.C;
  if (CC > 16) CC = 16;
  const size_t smem = (size_t)CC * per_c * 4 + 16;
  static bool attr = false;
  if (!attr) { NGB_CUDA(cudaFuncSetAttribute(conv_fprop_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDirectSmem)); attr = true; }
  const int64_t blocks = (int64_t)ceil_div(g.N, IMGS) * ceil_div(g.O, OT) * ceil_div(g.Ho, TP) * ceil_div(g.Wo, TQ);
  NGB_CHECK_ARG(blocks < (int64_t(1) << 31), "conv fprop: grid too large");
  conv_fprop_direct_kernel<<<(unsigned)blocks, 256, smem, st>>>(x, w, b, y, g, TP, TQ, IMGS, CC);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_conv_dgrad(const float* ug, const float* w, float* dx, int64_t N, int64_t C, int64_t H, int64_t W,
                              int64_t O, int64_t KH, int64_t KW, int64_t pad, int64_t stride, int accumulate,
                              ngb_stream_t stream) {
  ConvGeo g;
  int rc = make_geo(&g, N, C, H, W, O, KH, KW, pad, stride);
  if (rc) return rc;
  if (N == 0) return NGB_OK;
  cudaStream_t st = as_stream(stream);
  if (conv_igemm_eligible(N, C, H, W, O, KH, KW, pad, stride) && aligned16(ug) && aligned16(dx) && aligned16(w))
    return conv_igemm_dgrad(ug, w, dx, N, C, H, W, O, KH, KW, pad, accumulate, st);
  // (conv_small's dgrad mode is correct but measured slower than tap skipping on LeNet's 8x8 -> 12x12 layer -- the
  //  full-correlation zero halo makes it do 2.25x the useful work -- so small layers use conv_dgrad_small instead.)
  {
    bool done = false;
    rc = conv_mma_dgrad_rows(ug, w, dx, g.N, g.C, g.H, g.W, g.O, g.KH, g.KW, g.pad, g.stride, g.Ho, g.Wo, accumulate, st, &done);
    if (rc || done) return rc;
    rc = conv_mma_gather(ug, w, nullptr, dx, g.N, g.C, g.H, g.W, g.O, g.KH, g.KW, g.pad, g.stride, g.Ho, g.Wo, 1, accumulate, st, &done);
    if (rc || done) return rc;
    rc = conv_dgrad_small(ug, w, dx, g, accumulate, st, &done);
    if (rc || done) return rc;
  }
  int TW = g.W < 32 ? g.W : 32;
  int TH = 256 / TW; if (TH > g.H) TH = g.H;
  const int PPmax = (TH + g.KH - 2) / g.stride + 2, PQmax = (TW + g.KW - 2) / g.stride + 2, PPp = PPmax | 1;
  const int64_t per_o = (int64_t)PQmax * PPp + (int64_t)g.KH * g.KW * OT;
  int OC = (int)((kDirectSmem / 4 - 4) / per_o);
  if (OC < 1) return fail(NGB_ERR_UNSUPPORTED, "conv dgrad: filter too large for the direct kernel tile");
  if (OC > g.O) OC = g.O;
  if (OC > 32) OC = 32;
  const size_t smem = (size_t)OC * per_o * 4 + 16;
  static bool attr = false;
  if (!attr) { NGB_CUDA(cudaFuncSetAttribute(conv_dgrad_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDirectSmem)); attr = true; }
  const int64_t blocks = (int64_t)g.N * ceil_div(g.C, OT) * ceil_div(g.H, TH) * ceil_div(g.W, TW);
  NGB_CHECK_ARG(blocks < (int64_t(1) << 31), "conv dgrad: grid too large");
  conv_dgrad_direct_kernel<<<(unsigned)blocks, 256, smem, st>>>(ug, w, dx, g, TH, TW, OC, accumulate);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_conv_wgrad(const float* ug, const float* x, float* dw, int64_t N, int64_t C, int64_t H, int64_t W,
                              int64_t O, int64_t KH, int64_t KW, int64_t pad, int64_t stride, int accumulate,
                              ngb_stream_t stream) {
  ConvGeo g;
  int rc = make_geo(&g, N, C, H, W, O, KH, KW, pad, stride);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const int64_t nw = O * C * KH * KW;
  if (N == 0) {
    if (!accumulate) return ngb_fill(dw, 0.f, nw, stream);
    return NGB_OK;
  }
  if (conv_igemm_eligible(N, C, H, W, O, KH, KW, pad, stride) && g.Ho % 4 == 0 && aligned16(ug) && aligned16(x) && aligned16(dw))
    return conv_igemm_wgrad(ug, x, dw, N, C, H, W, O, KH, KW, pad, accumulate, st);
  {
    bool done = false;
    rc = conv_mma_wgrad(ug, x, dw, g.N, g.C, g.H, g.W, g.O, g.KH, g.KW, g.pad, g.stride, g.Ho, g.Wo, accumulate, st, &done);
    if (rc || done) return rc;
    rc = conv_wgrad_small(ug, x, dw, g, accumulate, st, &done);
    if (rc || done) return rc;
  }
  rc = ensure_init();
  if (rc) return rc;
  const int64_t xblocks = ceil_div<int64_t>(nw, 256);
  int64_t splits = ceil_div<int64_t>(4 * kNumSMs, xblocks);
  if (splits > N) splits = N;
  const int64_t cap = scratch_bytes() / (nw * (int64_t)sizeof(float));
  if (splits > cap) splits = cap;
  if (splits < 1) return fail(NGB_ERR_SCRATCH, "conv wgrad: scratch arena too small for %lld filter elements", (long long)nw);
  const int64_t ips = ceil_div<int64_t>(N, splits);
  splits = ceil_div<int64_t>(N, ips);
  dim3 grid((unsigned)xblocks, (unsigned)splits);
  conv_wgrad_direct_kernel<<<grid, 256, 0, st>>>(ug, x, scratch_ptr(), g, (int)ips);
  NGB_LAUNCH_CHECK();
  return launch_sum_partials(scratch_ptr(), (int)splits, nw, dw, accumulate, st);
}

// dw (+)= wgrad of a convolution that is followed by ReLU and a 2x2 / stride 2 MaxPool, straight from the POOLED upstream
// gradient: ug_pool (gradient of the pool's output), y_pool (the pool's output), idx (its argmax indices) and z (the conv
// output = ReLU input).  The kernel applies the ReLU mask and rebuilds the full-resolution gradient (three quarters zeros)
// in shared memory only.  Small-channel tensor-core path only (NGB_ERR_UNSUPPORTED otherwise: the caller materialises the
// gradient with ngb_maxpool_relu_bwd and calls ngb_conv_wgrad).
extern "C" int ngb_conv_wgrad_pooled(const float* ug_pool, const float* y_pool, const uint8_t* idx, const float* z, const float* x,
                                     float* dw, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW,
                                     int64_t pad, int64_t stride, int accumulate, ngb_stream_t stream) {
  ConvGeo g;
  int rc = make_geo(&g, N, C, H, W, O, KH, KW, pad, stride);
  if (rc) return rc;
  NGB_CHECK_ARG(idx != nullptr, "ngb_conv_wgrad_pooled needs the argmax bytes of ngb_maxpool_relu_fwd");
  if (N == 0) return fail(NGB_ERR_UNSUPPORTED, "ngb_conv_wgrad_pooled: empty batch");
  bool done = false;
  rc = conv_sparse_wbgrad_pooled(ug_pool, y_pool, idx, z, x, dw, nullptr, g.N, g.C, g.H, g.W, g.O, g.KH, g.KW, g.pad, g.stride, g.Ho,
                                 g.Wo, accumulate, 0, as_stream(stream), &done);
  if (rc || done) return rc;
  rc = conv_mma_wgrad(ug_pool, x, dw, g.N, g.C, g.H, g.W, g.O, g.KH, g.KW, g.pad, g.stride, g.Ho, g.Wo, accumulate, as_stream(stream),
                      &done, idx, y_pool, z);
  if (rc) return rc;
  if (!done) return fail(NGB_ERR_UNSUPPORTED, "ngb_conv_wgrad_pooled: geometry not served by the small-channel tensor-core kernel");
  return NGB_OK;
}

static int conv_bgrad_impl(const float* ug, float* db, int64_t N, int64_t O, int64_t HoWo, int accumulate, const float* y,
                           const float* z, int64_t Hp, int64_t Wo_full, ngb_stream_t stream, const uint8_t* idx) {
  NGB_CHECK_ARG(N >= 0 && O >= 1 && HoWo >= 0, "conv bgrad: bad dimensions");
  cudaStream_t st = as_stream(stream);
  if (N == 0 || HoWo == 0) {
    if (!accumulate) return ngb_fill(db, 0.f, O, stream);
    return NGB_OK;
  }
  int rc = ensure_init();
  if (rc) return rc;
  int64_t splits = ceil_div<int64_t>(8 * kNumSMs, O);
  if (splits > N) splits = N;
  // keep at least ~4K elements per CTA
  const int64_t min_imgs = ceil_div<int64_t>(4096, HoWo);
  if (splits > ceil_div<int64_t>(N, min_imgs)) splits = ceil_div<int64_t>(N, min_imgs);
  if (splits < 1) splits = 1;
  const int64_t ips = ceil_div<int64_t>(N, splits);
  splits = ceil_div<int64_t>(N, ips);
  dim3 grid((unsigned)O, (unsigned)splits);
  const int vec = (HoWo % 4 == 0) && aligned16(ug) && (y == nullptr || aligned16(y)) && (ips * (HoWo / 4) < (int64_t(1) << 32)) &&
                  (idx == nullptr || (reinterpret_cast<uintptr_t>(idx) & 3) == 0);
  const FastDiv dhw4((uint32_t)(vec ? HoWo / 4 : 1)), dhp((uint32_t)(Hp > 0 ? Hp : 1));
  if (idx)
    conv_bgrad_kernel<false, true><<<grid, 256, 0, st>>>(ug, scratch_ptr(), N, O, HoWo, ips, dhw4, vec, nullptr, nullptr, 1, 1, dhp, idx);
  else if (y)
    conv_bgrad_kernel<true><<<grid, 256, 0, st>>>(ug, scratch_ptr(), N, O, HoWo, ips, dhw4, vec, y, z, (int)Hp, (int)Wo_full, dhp);
  else
    conv_bgrad_kernel<false><<<grid, 256, 0, st>>>(ug, scratch_ptr(), N, O, HoWo, ips, dhw4, vec, nullptr, nullptr, 1, 1, dhp);
  NGB_LAUNCH_CHECK();
  return launch_sum_partials(scratch_ptr(), (int)splits, O, db, accumulate, st);
}

extern "C" int ngb_conv_bgrad(const float* ug, float* db, int64_t N, int64_t O, int64_t HoWo, int accumulate,
                              ngb_stream_t stream) {
  return conv_bgrad_impl(ug, db, N, O, HoWo, accumulate, nullptr, nullptr, 0, 0, stream, nullptr);
}

static int conv_bgrad_impl(const float* ug, float* db, int64_t N, int64_t O, int64_t HoWo, int accumulate, const float* y,
                           const float* z, int64_t Hp, int64_t Wo_full, ngb_stream_t stream, const uint8_t* idx = nullptr);

// Weight AND bias gradient (conv.py:311-327) of conv -> ReLU -> MaxPool(2x2, s2) from the pooled upstream gradient in one call:
// first layers with few input channels get both from one pass over the pooled gradient (sparse fp32 kernel); other geometries
// run ngb_conv_wgrad_pooled + ngb_conv_bgrad_pooled.  db may be NULL.
extern "C" int ngb_conv_wbgrad_pooled(const float* ug_pool, const float* y_pool, const uint8_t* idx, const float* z, const float* x,
                                      float* dw, float* db, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH,
                                      int64_t KW, int64_t pad, int64_t stride, int accumulate_w, int accumulate_b,
                                      ngb_stream_t stream) {
  ConvGeo g;
  int rc = make_geo(&g, N, C, H, W, O, KH, KW, pad, stride);
  if (rc) return rc;
  NGB_CHECK_ARG(idx != nullptr, "ngb_conv_wbgrad_pooled needs the argmax bytes of ngb_maxpool_relu_fwd");
  if (N == 0) return fail(NGB_ERR_UNSUPPORTED, "ngb_conv_wbgrad_pooled: empty batch");
  if ((g.Ho & 1) || (g.Wo & 1)) return fail(NGB_ERR_UNSUPPORTED, "ngb_conv_wbgrad_pooled: odd conv output plane");
  bool done = false;
  rc = conv_sparse_wbgrad_pooled(ug_pool, y_pool, idx, z, x, dw, db, g.N, g.C, g.H, g.W, g.O, g.KH, g.KW, g.pad, g.stride, g.Ho, g.Wo,
                                 accumulate_w, accumulate_b, as_stream(stream), &done);
  if (rc || done) return rc;
  rc = conv_mma_wgrad(ug_pool, x, dw, g.N, g.C, g.H, g.W, g.O, g.KH, g.KW, g.pad, g.stride, g.Ho, g.Wo, accumulate_w, as_stream(stream),
                      &done, idx, y_pool, z);
  if (rc) return rc;
  if (!done) return fail(NGB_ERR_UNSUPPORTED, "ngb_conv_wbgrad_pooled: geometry not served by the pooled weight-gradient kernels");
  if (db) return conv_bgrad_impl(ug_pool, db, g.N, g.O, (g.Ho / 2) * (g.Wo / 2), accumulate_b, nullptr, nullptr, 0, 0, stream, idx);
  return NGB_OK;
}

// y_pool / idx = maxpool2x2s2(relu(conv(x, w) + b)) in one kernel; the convolution's output is never written (first-layer
// geometries only: NGB_ERR_UNSUPPORTED -> ngb_conv_fprop + ngb_maxpool_relu_fwd).  idx carries the ReLU flag in bit 2.
extern "C" int ngb_conv_relu_pool_fwd(const float* x, const float* w, const float* b, float* y_pool, uint8_t* idx, int64_t N,
                                      int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW, int64_t pad,
                                      int64_t stride, ngb_stream_t stream) {
  ConvGeo g;
  int rc = make_geo(&g, N, C, H, W, O, KH, KW, pad, stride);
  if (rc) return rc;
  NGB_CHECK_ARG(idx != nullptr && y_pool != nullptr, "ngb_conv_relu_pool_fwd: NULL output");
  if (N == 0) return fail(NGB_ERR_UNSUPPORTED, "ngb_conv_relu_pool_fwd: empty batch");
  bool done = false;
  rc = conv_fused_relu_pool_fwd(x, w, b, y_pool, idx, g.N, g.C, g.H, g.W, g.O, g.KH, g.KW, g.pad, g.stride, g.Ho, g.Wo,
                                as_stream(stream), &done);
  if (rc) return rc;
  if (!done) return fail(NGB_ERR_UNSUPPORTED, "ngb_conv_relu_pool_fwd: geometry not served by the fused first-layer kernel");
  return NGB_OK;
}

extern "C" int ngb_conv_relu_pool_fwd_serves(int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW,
                                             int64_t pad, int64_t stride) {
  if (stride < 1 || pad < 0 || N < 1 || C < 1 || O < 1 || KH < 1 || KW < 1 || H + 2 * pad < KH || W + 2 * pad < KW ||
      N * C * H * W >= (int64_t(1) << 40))
    return 0;
  const int Ho = (int)((H + 2 * pad - KH) / stride + 1), Wo = (int)((W + 2 * pad - KW) / stride + 1);
  return conv_fused_fwd_serves((int)C, (int)H, (int)W, (int)O, (int)KH, (int)KW, (int)pad, (int)stride, Ho, Wo) ? 1 : 0;
}

// Debug aid (host only, no GPU needed): the shared-memory layout the small-channel weight-gradient kernel would use for a
// layer -- out[0..5] = row stride, column stride, plane pitch, taps-kh-fastest flag, wavefronts per gather load x 1000 for
// that layout and for the natural one.
extern "C" int ngb_debug_wgrad_layout(int64_t C, int64_t H, int64_t W, int64_t KH, int64_t KW, int64_t pad, int64_t stride,
                                      int* out) {
  NGB_CHECK_ARG(out != nullptr && C >= 1 && KH >= 1 && KW >= 1 && stride >= 1 && pad >= 0 && H + 2 * pad >= KH && W + 2 * pad >= KW &&
                    C * KH * KW <= 4096 && H <= 4096 && W <= 4096,
                "ngb_debug_wgrad_layout: bad geometry");
  return wgrad_layout_for_tests((int)C, (int)H, (int)W, (int)KH, (int)KW, (int)pad, (int)stride, out);
}

// 1 when ngb_conv_wbgrad_pooled serves this geometry with ONE kernel (callers that would otherwise overlap the weight and
// the bias gradient on two streams ask first), 0 otherwise.
extern "C" int ngb_conv_wbgrad_pooled_is_fused(int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW,
                                               int64_t pad, int64_t stride) {
  if (stride < 1 || pad < 0 || N < 1 || C < 1 || O < 1 || KH < 1 || KW < 1 || H + 2 * pad < KH || W + 2 * pad < KW ||
      N * C * H * W >= (int64_t(1) << 40))
    return 0;
  const int Ho = (int)((H + 2 * pad - KH) / stride + 1), Wo = (int)((W + 2 * pad - KW) / stride + 1);
  return conv_sparse_serves((int)C, (int)H, (int)W, (int)O, (int)KH, (int)KW, (int)pad, Ho, Wo) ? 1 : 0;
}

// dx (+)= data gradient (conv.py:299-309) of a convolution followed by ReLU and a 2x2 / stride 2 MaxPool, straight from the
// pooled upstream gradient (same operands as ngb_conv_wgrad_pooled).  Row-decomposed tensor-core kernel only.
extern "C" int ngb_conv_dgrad_pooled(const float* ug_pool, const float* y_pool, const uint8_t* idx, const float* z, const float* w,
                                     float* dx, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW,
                                     int64_t pad, int64_t stride, int accumulate, ngb_stream_t stream) {
  ConvGeo g;
  int rc = make_geo(&g, N, C, H, W, O, KH, KW, pad, stride);
  if (rc) return rc;
  NGB_CHECK_ARG(idx != nullptr, "ngb_conv_dgrad_pooled needs the argmax bytes of ngb_maxpool_relu_fwd");
  if (N == 0) return fail(NGB_ERR_UNSUPPORTED, "ngb_conv_dgrad_pooled: empty batch");
  bool done = false;
  rc = conv_mma_dgrad_rows(ug_pool, w, dx, g.N, g.C, g.H, g.W, g.O, g.KH, g.KW, g.pad, g.stride, g.Ho, g.Wo, accumulate,
                           as_stream(stream), &done, idx, y_pool, z);
  if (rc) return rc;
  if (!done) return fail(NGB_ERR_UNSUPPORTED, "ngb_conv_dgrad_pooled: geometry not served by the row-decomposed tensor-core kernel");
  return NGB_OK;
}

// db (+)= sum over n and pooled positions of ug_pool where bit 2 of the argmax byte (ngb_maxpool_relu_fwd: relu'(z at the
// argmax)) is set: the bias gradient from ug_pool and one byte per element.  HpWp = pooled positions per (n, o) plane.
extern "C" int ngb_conv_bgrad_pooled_idx(const float* ug_pool, const uint8_t* idx, float* db, int64_t N, int64_t O, int64_t HpWp,
                                         int accumulate, ngb_stream_t stream) {
  NGB_CHECK_ARG(idx != nullptr, "ngb_conv_bgrad_pooled_idx needs the argmax bytes of ngb_maxpool_relu_fwd");
  return conv_bgrad_impl(ug_pool, db, N, O, HpWp, accumulate, nullptr, nullptr, 0, 0, stream, idx);
}

// db (+)= bias gradient of a convolution followed by ReLU and a 2x2 / stride 2 MaxPool, straight from the pooled upstream
// gradient: sum over n and pooled positions of ug_pool * relu'(z at the argmax).  Ho x Wo = the conv output plane (even).
extern "C" int ngb_conv_bgrad_pooled(const float* ug_pool, const float* y_pool, const float* z, float* db, int64_t N, int64_t O,
                                     int64_t Ho, int64_t Wo, int accumulate, ngb_stream_t stream) {
  NGB_CHECK_ARG(y_pool != nullptr && z != nullptr && Ho >= 2 && Wo >= 2 && Ho % 2 == 0 && Wo % 2 == 0,
                "ngb_conv_bgrad_pooled: needs the pool output, the conv output and an even output plane");
  return conv_bgrad_impl(ug_pool, db, N, O, (Ho / 2) * (Wo / 2), accumulate, y_pool, z, Ho / 2, Wo, stream, nullptr);
}
