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
//
// Memory layout problem: the input is contiguous along W while the output is contiguous along p (the H direction),
// so a thread-per-output kernel reads with stride W (measured 21 % / 13 % of HBM peak for fwd / bwd).  Both kernels
// therefore stage a tile through shared memory: coalesced 128-bit global reads -> smem -> transposed access in smem ->
// coalesced global writes.  A work item is (group of G planes, strip of TP output rows p, all q); G is chosen so that
// small planes (LeNet: 24x24, 8x8) still give every CTA ~16 KB to move.
#include "common.cuh"

namespace ngb {

// 2x2 pools: bit 2 of an argmax byte written by ngb_maxpool_relu_fwd says whether the ReLU gradient at the argmax is 1
constexpr int kPoolReluFlag = 4;

struct PoolGeo {
  int H, W, Ho, Wo, KH, KW, pad, stride;
  int TP;       // output rows per strip
  int G;        // planes per work item
  int rows;     // input rows staged per strip: (TP-1)*stride + KH
  int Wp;       // padded smem row length for the forward input tile (W + 2*pad, made odd)
  int strips;   // ceil(Ho / TP)
  FastDiv d_w4, d_w, d_rows, d_tp, d_wo, d_stride;
};

// ---------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(256) maxpool_fwd_tiled_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                uint8_t* __restrict__ idx, int64_t outer, PoolGeo g) {
  extern __shared__ float sm[];   // [G][rows][Wp], padded coordinates (column c <-> w = c - pad)
  const int64_t groups = ceil_div<int64_t>(outer, g.G);
  const int64_t items = groups * g.strips;
  const int plane_sm = g.rows * g.Wp;
  const int Wpad = g.W + 2 * g.pad;
  const bool vec = (g.W % 4 == 0) && ((((int64_t)g.H * g.W) % 4) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t grp = item / g.strips;
    const int strip = (int)(item - grp * g.strips);
    const int64_t o0 = grp * g.G;
    const int ng = (int)((int64_t)g.G < outer - o0 ? (int64_t)g.G : outer - o0);
    const int p0 = strip * g.TP;
    const int tp = min(g.TP, g.Ho - p0);
    const int hrow0 = p0 * g.stride;                       // first padded row staged
    const int nrows = (tp - 1) * g.stride + g.KH;
    __syncthreads();                                       // previous item's smem reads are done
    // ---- stage: zero halo + coalesced copy of the valid rows
    if (g.pad > 0) {
      for (int e = threadIdx.x; e < ng * plane_sm; e += 256) sm[e] = 0.f;
      __syncthreads();
    }
    if (vec) {
      const int w4 = g.W >> 2;
      const int total = ng * g.rows * w4;
#pragma unroll 4
      for (int e = threadIdx.x; e < total; e += 256) {
        uint32_t R, c4, gi, r;
        g.d_w4.divmod(e, R, c4);
        g.d_rows.divmod(R, gi, r);
        const int h = hrow0 + (int)r - g.pad;
        if ((int)r < nrows && h >= 0 && h < g.H) {
          const float4 v = ld_stream4(x + ((o0 + gi) * g.H + h) * (int64_t)g.W + 4 * c4);
          float* d = sm + gi * plane_sm + r * g.Wp + g.pad + 4 * c4;
          d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
      }
    } else {
      const int total = ng * g.rows * g.W;
#pragma unroll 4
      for (int e = threadIdx.x; e < total; e += 256) {
        uint32_t R, c, gi, r;
        g.d_w.divmod(e, R, c);
        g.d_rows.divmod(R, gi, r);
        const int h = hrow0 + (int)r - g.pad;
        if ((int)r < nrows && h >= 0 && h < g.H) sm[gi * plane_sm + r * g.Wp + g.pad + c] = __ldg(x + ((o0 + gi) * g.H + h) * (int64_t)g.W + c);
      }
    }
    __syncthreads();
    // ---- compute: thread <-> (plane, q, p) with p fastest (the output's contiguous axis)
    const int total_out = ng * g.Wo * g.TP;
    for (int e = threadIdx.x; e < total_out; e += 256) {
      uint32_t rest, ulp, gi, uq;
      g.d_tp.divmod(e, rest, ulp);
      g.d_wo.divmod(rest, gi, uq);
      const int lp = (int)ulp, q = (int)uq;
      if (lp >= tp) continue;
      const float* base = sm + gi * plane_sm + (lp * g.stride) * g.Wp + q * g.stride;
      float best = base[0];
      int arg = 0;
      for (int kh = 0; kh < g.KH; ++kh) {
        for (int kw = 0; kw < g.KW; ++kw) {
          const float v = base[kh * g.Wp + kw];
          if (v > best) { best = v; arg = kh * g.KW + kw; }   // strict > keeps the FIRST maximum (np.argmax)
        }
      }
      const int64_t f = (o0 + gi) * (int64_t)g.Ho * g.Wo + (int64_t)q * g.Ho + p0 + lp;
      y[f] = best;
      if (idx) idx[f] = (uint8_t)arg;
    }
    (void)Wpad;
  }
}

// ---------------------------------------------------------------------------------------------- backward
// Stage ug and idx of the strip in smem as [G][Wo][TPp] (p contiguous, as in global memory), then one thread per input
// cell (w fastest) looks up the LAST window that covers it.
__global__ void __launch_bounds__(256) maxpool_bwd_tiled_kernel(const float* __restrict__ ug, const uint8_t* __restrict__ idx,
                                                                float* __restrict__ dx, int64_t outer, PoolGeo g,
                                                                int accumulate) {
  extern __shared__ float sm[];
  // strip of input rows [h0, h1) <-> output rows that can be "the last window covering" those rows
  const int TPp = g.TP | 1;
  const int imask = (g.KH * g.KW <= 4) ? 3 : 255;                     // 2x2 pools: drop ngb_maxpool_relu_fwd's flag bit
  float* sug = sm;                                                    // [G][Wo][TPp]
  uint8_t* sidx = reinterpret_cast<uint8_t*>(sm + g.G * g.Wo * TPp);   // [G][Wo][TPp]
  const int64_t groups = ceil_div<int64_t>(outer, g.G);
  // input strips: rows whose last covering window lies in [p0, p0+TP): h+pad in [p0*s, (p0+TP)*s) (last strip: to H)
  const int64_t items = groups * g.strips;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t grp = item / g.strips;
    const int strip = (int)(item - grp * g.strips);
    const int64_t o0 = grp * g.G;
    const int ng = (int)((int64_t)g.G < outer - o0 ? (int64_t)g.G : outer - o0);
    const int p0 = strip * g.TP;
    const int tp = min(g.TP, g.Ho - p0);
    __syncthreads();
    const int total_o = ng * g.Wo * g.TP;
#pragma unroll 4
    for (int e = threadIdx.x; e < total_o; e += 256) {
      uint32_t rest, ulp, gi, uq;
      g.d_tp.divmod(e, rest, ulp);
      g.d_wo.divmod(rest, gi, uq);
      const int lp = (int)ulp, q = (int)uq;
      if (lp >= tp) continue;
      const int64_t f = (o0 + gi) * (int64_t)g.Ho * g.Wo + (int64_t)q * g.Ho + p0 + lp;
      sug[(gi * g.Wo + q) * TPp + lp] = __ldg(ug + f);
      sidx[(gi * g.Wo + q) * TPp + lp] = idx[f] & imask;
    }
    __syncthreads();
    // padded rows hp = h + pad handled by this strip
    int hp_lo = p0 * g.stride, hp_hi = (p0 + tp) * g.stride;
    if (strip == 0) hp_lo = 0;
    if (strip == g.strips - 1) hp_hi = g.H + 2 * g.pad;
    int h_lo = max(hp_lo - g.pad, 0), h_hi = min(hp_hi - g.pad, g.H);
    const int nrows = max(h_hi - h_lo, 0);
    const int per_plane = nrows * g.W;
    for (int gi = 0; gi < ng; ++gi)
    for (int e = threadIdx.x; e < per_plane; e += 256) {
      uint32_t ur, uw;
      g.d_w.divmod(e, ur, uw);
      const int w = (int)uw, r = (int)ur;
      const int h = h_lo + r;
      const int hp = h + g.pad, wp = w + g.pad;
      int q = (int)g.d_stride.div(wp), p = (int)g.d_stride.div(hp);
      if (q > g.Wo - 1) q = g.Wo - 1;
      if (p > g.Ho - 1) p = g.Ho - 1;
      float gv = 0.f;
      if (q * g.stride + g.KW > wp && p * g.stride + g.KH > hp) {   // covered by window (p,q): the last in fragment order
        const int local = (hp - p * g.stride) * g.KW + (wp - q * g.stride);
        const int s = (gi * g.Wo + q) * TPp + (p - p0);
        if ((int)sidx[s] == local) gv = sug[s];
      }
      float* d = dx + ((o0 + gi) * g.H + h) * (int64_t)g.W + w;
      *d = accumulate ? *d + gv : gv;
    }
  }
}

// ---------------------------------------------------------------------------------------------- 2x2 / stride 2 / pad 0
// The geometry every config uses (LeNet, the C5 pool sweep).  Only the OUTPUT tile (4x smaller than the input) goes
// through shared memory: threads read two 128-bit row segments (coalesced along W), reduce two windows each, park the
// results in smem as [q][p], and the tile is written back with p contiguous.
struct Pool2Geo {
  int H, W, Ho, Wo, TP, G, strips, w4;
  FastDiv d_w4, d_tp, d_wo;
};

__device__ __forceinline__ void max4(float a, float b, float c, float d, float& best, int& arg) {
  best = a; arg = 0;                        // window order (kh,kw): (0,0) (0,1) (1,0) (1,1); strict > keeps the first max
  if (b > best) { best = b; arg = 1; }
  if (c > best) { best = c; arg = 2; }
  if (d > best) { best = d; arg = 3; }
}

// RELU: the input is max(0, x) of the given buffer (the ReLU in front of the pool is applied on load, its output is
// never written: ngb_maxpool_relu_fwd).
template <bool RELU>
__global__ void __launch_bounds__(256) maxpool2x2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                             uint8_t* __restrict__ idx, int64_t outer, Pool2Geo g) {
  extern __shared__ float sm[];
  const int TPp = g.TP | 1;
  float* sy = sm;                                                       // [G][Wo][TPp]
  uint8_t* si = reinterpret_cast<uint8_t*>(sm + g.G * g.Wo * TPp);       // [G][Wo][TPp]
  const int64_t items = ceil_div<int64_t>(outer, g.G) * g.strips;
  const int wq = g.Wo >> 1;                                             // float4 columns that hold two complete windows
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t grp = item / g.strips;
    const int strip = (int)(item - grp * g.strips);
    const int64_t o0 = grp * g.G;
    const int ng = (int)((int64_t)g.G < outer - o0 ? (int64_t)g.G : outer - o0);
    const int p0 = strip * g.TP;
    const int tp = min(g.TP, g.Ho - p0);
    __syncthreads();
    const int total = ng * g.TP * g.w4;
#pragma unroll 2
    for (int e = threadIdx.x; e < total; e += 256) {
      uint32_t rest, c4, gi, lp;
      g.d_w4.divmod(e, rest, c4);
      g.d_tp.divmod(rest, gi, lp);
      if ((int)lp >= tp) continue;
      const float* r0 = x + ((o0 + gi) * g.H + 2 * (p0 + (int)lp)) * (int64_t)g.W + 4 * c4;
      float4 a = ld_stream4(r0), b = ld_stream4(r0 + g.W);
      const float first0 = a.x, first1 = a.z;      // first cell of each window before the clamp
      if (RELU) {   // same expression as the elementwise ReLU kernel (elementwise.cu): fmaxf(0, x)
        a.x = fmaxf(0.f, a.x); a.y = fmaxf(0.f, a.y); a.z = fmaxf(0.f, a.z); a.w = fmaxf(0.f, a.w);
        b.x = fmaxf(0.f, b.x); b.y = fmaxf(0.f, b.y); b.z = fmaxf(0.f, b.z); b.w = fmaxf(0.f, b.w);
      }
      float v0, v1;
      int a0, a1;
      max4(a.x, a.y, b.x, b.y, v0, a0);
      max4(a.z, a.w, b.z, b.w, v1, a1);
      if (RELU) {
        // bit 2 of the index byte = relu'(x at the argmax) (activations.py:31: x >= 0).  A positive maximum passed the ReLU;
        // a zero maximum means every cell was clamped, the argmax is the first cell and the gradient flows iff that cell
        // is exactly zero or above.  The pooled backward kernels read ug and this byte and nothing else.
        if (v0 > 0.f || first0 >= 0.f) a0 |= kPoolReluFlag;
        if (v1 > 0.f || first1 >= 0.f) a1 |= kPoolReluFlag;
      }
      const int q = 2 * (int)c4;
      const int s0 = ((int)gi * g.Wo + q) * TPp + (int)lp;
      if (q < g.Wo) { sy[s0] = v0; si[s0] = (uint8_t)a0; }
      if (q + 1 < g.Wo) { sy[s0 + TPp] = v1; si[s0 + TPp] = (uint8_t)a1; }
    }
    __syncthreads();
    const int total_out = ng * g.Wo * g.TP;
#pragma unroll 2
    for (int e = threadIdx.x; e < total_out; e += 256) {
      uint32_t rest, lp, gi, q;
      g.d_tp.divmod(e, rest, lp);
      g.d_wo.divmod(rest, gi, q);
      if ((int)lp >= tp) continue;
      const int s0 = ((int)gi * g.Wo + (int)q) * TPp + (int)lp;
      const int64_t f = (o0 + gi) * (int64_t)g.Ho * g.Wo + (int64_t)q * g.Ho + p0 + (int)lp;
      y[f] = sy[s0];
      if (idx) idx[f] = si[s0];
    }
    (void)wq;
  }
}

// MASK: the result is additionally multiplied by the ReLU gradient mask (xm >= 0) of the buffer the ReLU in front of
// the pool read: one kernel for maxpool backward + relu backward (ngb_maxpool_relu_bwd).
template <bool MASK>
__global__ void __launch_bounds__(256) maxpool2x2_bwd_kernel(const float* __restrict__ ug, const uint8_t* __restrict__ idx,
                                                             float* __restrict__ dx, int64_t outer, Pool2Geo g, int accumulate,
                                                             const float* __restrict__ xm) {
  extern __shared__ float sm[];
  const int TPp = g.TP | 1;
  float* sg = sm;
  uint8_t* si = reinterpret_cast<uint8_t*>(sm + g.G * g.Wo * TPp);
  const int64_t items = ceil_div<int64_t>(outer, g.G) * g.strips;
  // 128-bit staging of ug / 32-bit staging of idx along p when every strip start is a multiple of 4
  const bool vec = (g.Ho % 4 == 0) && (g.TP % 4 == 0) && ((reinterpret_cast<uintptr_t>(ug) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(idx) & 3) == 0);
  const int tp4 = g.TP >> 2;
  const FastDiv d_tp4((uint32_t)(tp4 > 0 ? tp4 : 1));
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t grp = item / g.strips;
    const int strip = (int)(item - grp * g.strips);
    const int64_t o0 = grp * g.G;
    const int ng = (int)((int64_t)g.G < outer - o0 ? (int64_t)g.G : outer - o0);
    const int p0 = strip * g.TP;
    const int tp = min(g.TP, g.Ho - p0);
    __syncthreads();
    if (vec) {
      const int total_o = ng * g.Wo * tp4;
#pragma unroll 4
      for (int e = threadIdx.x; e < total_o; e += 256) {
        uint32_t rest, l4, gi, q;
        d_tp4.divmod(e, rest, l4);
        g.d_wo.divmod(rest, gi, q);
        const int lp = 4 * (int)l4;
        if (lp < tp) {
          const int64_t f = (o0 + gi) * (int64_t)g.Ho * g.Wo + (int64_t)q * g.Ho + p0 + lp;
          const float4 v = ld_stream4(ug + f);
          const uchar4 a = *reinterpret_cast<const uchar4*>(idx + f);
          const int s0 = ((int)gi * g.Wo + (int)q) * TPp + lp;
          sg[s0] = v.x; sg[s0 + 1] = v.y; sg[s0 + 2] = v.z; sg[s0 + 3] = v.w;
          si[s0] = a.x & 3; si[s0 + 1] = a.y & 3; si[s0 + 2] = a.z & 3; si[s0 + 3] = a.w & 3;   // bit 2: ReLU flag
        }
      }
    } else {
      const int total_o = ng * g.Wo * g.TP;
#pragma unroll 4
      for (int e = threadIdx.x; e < total_o; e += 256) {
        uint32_t rest, lp, gi, q;
        g.d_tp.divmod(e, rest, lp);
        g.d_wo.divmod(rest, gi, q);
        if ((int)lp < tp) {
          const int64_t f = (o0 + gi) * (int64_t)g.Ho * g.Wo + (int64_t)q * g.Ho + p0 + (int)lp;
          const int s0 = ((int)gi * g.Wo + (int)q) * TPp + (int)lp;
          sg[s0] = __ldg(ug + f);
          si[s0] = idx[f] & 3;                   // 2x2 windows: bit 2 is ngb_maxpool_relu_fwd's ReLU flag
        }
      }
    }
    __syncthreads();
    // each thread owns a 2-row x 4-column block of dx = two complete windows: one smem lookup, two 128-bit stores
    const int total_w = ng * g.TP * g.w4;
#pragma unroll 2
    for (int e = threadIdx.x; e < total_w; e += 256) {
      uint32_t rest, c4, gi, lp;
      g.d_w4.divmod(e, rest, c4);
      g.d_tp.divmod(rest, gi, lp);
      if ((int)lp >= tp) continue;
      const int q = 2 * (int)c4;
      const int s0 = ((int)gi * g.Wo + q) * TPp + (int)lp;
      float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
      if (q < g.Wo) {
        const float gv = sg[s0];
        const int a = si[s0];
        r0.x = a == 0 ? gv : 0.f; r0.y = a == 1 ? gv : 0.f; r1.x = a == 2 ? gv : 0.f; r1.y = a == 3 ? gv : 0.f;
      }
      if (q + 1 < g.Wo) {
        const float gv = sg[s0 + TPp];
        const int a = si[s0 + TPp];
        r0.z = a == 0 ? gv : 0.f; r0.w = a == 1 ? gv : 0.f; r1.z = a == 2 ? gv : 0.f; r1.w = a == 3 ? gv : 0.f;
      }
      const int64_t off0 = ((o0 + gi) * g.H + 2 * (p0 + (int)lp)) * (int64_t)g.W + 4 * c4;
      float* d0 = dx + off0;
      float* d1 = d0 + g.W;
      if (MASK) {   // relu backward: gradient passes where the relu input was >= 0 (activations.py:31)
        const float4 m0 = ld_stream4(xm + off0), m1 = ld_stream4(xm + off0 + g.W);
        r0.x = m0.x >= 0.f ? r0.x : 0.f; r0.y = m0.y >= 0.f ? r0.y : 0.f; r0.z = m0.z >= 0.f ? r0.z : 0.f; r0.w = m0.w >= 0.f ? r0.w : 0.f;
        r1.x = m1.x >= 0.f ? r1.x : 0.f; r1.y = m1.y >= 0.f ? r1.y : 0.f; r1.z = m1.z >= 0.f ? r1.z : 0.f; r1.w = m1.w >= 0.f ? r1.w : 0.f;
      }
      if (accumulate) {
        const float4 o0v = *reinterpret_cast<const float4*>(d0), o1v = *reinterpret_cast<const float4*>(d1);
        r0.x += o0v.x; r0.y += o0v.y; r0.z += o0v.z; r0.w += o0v.w;
        r1.x += o1v.x; r1.y += o1v.y; r1.z += o1v.z; r1.w += o1v.w;
      }
      st_stream4(d0, r0);
      st_stream4(d1, r1);
    }
    // an odd H leaves one row no window covers: it receives zero gradient (the last strip owns it)
    if ((g.H & 1) && strip == g.strips - 1 && !accumulate) {
      for (int e = threadIdx.x; e < ng * g.w4; e += 256) {
        uint32_t gi, c4;
        g.d_w4.divmod(e, gi, c4);
        st_stream4(dx + ((o0 + gi) * g.H + g.H - 1) * (int64_t)g.W + 4 * c4, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
  }
}

// true when the specialised 2x2 kernels apply; fills g
static bool make_pool2_geo(Pool2Geo* g, int64_t outer, int64_t H, int64_t W, int64_t KH, int64_t KW, int64_t pad, int64_t stride,
                           int64_t Ho, int64_t Wo, const void* a, const void* b) {
  if (KH != 2 || KW != 2 || stride != 2 || pad != 0 || W % 4 != 0 || (H * W) % 4 != 0 || !aligned16(a) || !aligned16(b)) return false;
  g->H = (int)H; g->W = (int)W; g->Ho = (int)Ho; g->Wo = (int)Wo; g->w4 = (int)(W / 4);
  const int target = 20 * 1024;
  int TP = (int)Ho;
  while ((int64_t)Wo * (TP | 1) * 5 > target && TP > 8) TP = ((TP / 2) + 7) & ~7;
  if ((int64_t)Wo * (TP | 1) * 5 > 40 * 1024) return false;
  g->TP = TP;
  g->strips = (int)((Ho + TP - 1) / TP);
  int64_t G = target / ((int64_t)Wo * (TP | 1) * 5);
  if (G < 1) G = 1;
  if (G > 64) G = 64;
  if (G > outer) G = outer;
  g->G = (int)G;
  g->d_w4 = FastDiv((uint32_t)g->w4);
  g->d_tp = FastDiv((uint32_t)TP);
  g->d_wo = FastDiv((uint32_t)Wo);
  return true;
}

// ---------------------------------------------------------------------------------------------- fallback (any geometry)
// One thread per output (fwd) / input cell (bwd), straight from global memory: used when a plane row is too wide for the
// staged tile.
__global__ void __launch_bounds__(256) maxpool_fwd_simple_kernel(const float* __restrict__ x, float* __restrict__ y,
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

__global__ void __launch_bounds__(256) maxpool_bwd_simple_kernel(const float* __restrict__ ug, const uint8_t* __restrict__ idx,
                                                                 float* __restrict__ dx, int64_t outer, int H, int W, int Ho,
                                                                 int Wo, int KH, int KW, int pad, int stride, int accumulate) {
  const int64_t hw_in = (int64_t)H * W, total = outer * hw_in, hw_out = (int64_t)Ho * Wo;
  const int imask = (KH * KW <= 4) ? 3 : 255;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = i / hw_in;
    const int r = (int)(i - o * hw_in);
    const int h = r / W, w = r - h * W;
    const int hp = h + pad, wp = w + pad;
    int q = wp / stride, p = hp / stride;
    if (q > Wo - 1) q = Wo - 1;
    if (p > Ho - 1) p = Ho - 1;
    float g = 0.f;
    if (q * stride + KW > wp && p * stride + KH > hp) {
      const int64_t f = o * hw_out + (int64_t)q * Ho + p;
      const int local = (hp - p * stride) * KW + (wp - q * stride);
      if ((int)(idx[f] & imask) == local) g = ug[f];
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
  NGB_CHECK_ARG(H < (1 << 20) && W < (1 << 20), "pool plane too large");
  *Ho = (H + 2 * pad - KH) / stride + 1;
  *Wo = (W + 2 * pad - KW) / stride + 1;
  return NGB_OK;
}

constexpr int kPoolSmem = 48 * 1024;   // per CTA: 4 CTAs/SM keep enough loads in flight

static int make_pool_geo(PoolGeo* g, int64_t outer, int64_t H, int64_t W, int64_t KH, int64_t KW, int64_t pad, int64_t stride,
                         int64_t Ho, int64_t Wo, bool backward) {
  g->H = (int)H; g->W = (int)W; g->Ho = (int)Ho; g->Wo = (int)Wo; g->KH = (int)KH; g->KW = (int)KW;
  g->pad = (int)pad; g->stride = (int)stride;
  const int target = 24 * 1024;
  int Wp = (int)(W + 2 * pad) | 1;
  g->Wp = Wp;
  // strip height: as many output rows as fit the budget (a multiple of 8 when splitting, so writes stay 32B-sector sized)
  int TP = (int)Ho;
  for (;;) {
    const int64_t rows = (int64_t)(TP - 1) * stride + KH;
    const int64_t bytes = backward ? (int64_t)Wo * (TP | 1) * 5 : rows * Wp * 4;
    if (bytes <= target || TP <= 8) {
      if (bytes > kPoolSmem) return NGB_ERR_UNSUPPORTED;   // caller falls back to the simple kernel
      break;
    }
    TP = (TP / 2 + 7) & ~7;
    if (TP < 8) TP = 8;
  }
  g->TP = TP;
  g->rows = (TP - 1) * (int)stride + (int)KH;
  g->strips = (int)((Ho + TP - 1) / TP);
  const int64_t per_plane = backward ? (int64_t)Wo * (TP | 1) * 5 : (int64_t)g->rows * Wp * 4;
  int64_t G = target / (per_plane > 0 ? per_plane : 1);
  if (G < 1) G = 1;
  if (G > 64) G = 64;
  if (G > outer) G = outer;
  g->G = (int)G;
  g->d_w4 = FastDiv((uint32_t)(W >= 4 ? W / 4 : 1));
  g->d_w = FastDiv((uint32_t)W);
  g->d_rows = FastDiv((uint32_t)g->rows);
  g->d_tp = FastDiv((uint32_t)TP);
  g->d_wo = FastDiv((uint32_t)Wo);
  g->d_stride = FastDiv((uint32_t)stride);
  return NGB_OK;
}

// y = maxpool(relu(x)) and its argmax over relu(x), without materialising relu(x).  Only the 2x2 / stride 2 / pad 0
// geometry on 16-byte aligned rows (the specialised kernel); anything else returns NGB_ERR_UNSUPPORTED and the caller
// runs ReLU and the pool separately.
extern "C" int ngb_maxpool_relu_fwd(const float* x, float* y, uint8_t* idx, int64_t outer, int64_t H, int64_t W, int64_t KH,
                                    int64_t KW, int64_t pad, int64_t stride, ngb_stream_t stream) {
  int64_t Ho, Wo;
  int rc = pool_geometry(H, W, KH, KW, pad, stride, &Ho, &Wo);
  if (rc) return rc;
  Pool2Geo g2;
  if (outer * Ho * Wo == 0 || !make_pool2_geo(&g2, outer, H, W, KH, KW, pad, stride, Ho, Wo, x, x))
    return fail(NGB_ERR_UNSUPPORTED, "maxpool+relu fusion: only 2x2 / stride 2 / pad 0 on 16-byte aligned rows");
  const size_t smem2 = (size_t)g2.G * g2.Wo * (g2.TP | 1) * 5 + 16;
  const int64_t items2 = ceil_div<int64_t>(outer, g2.G) * g2.strips;
  const int grid2 = (int)(items2 < (int64_t)kNumSMs * 8 ? items2 : (int64_t)kNumSMs * 8);
  maxpool2x2_fwd_kernel<true><<<grid2, 256, smem2, as_stream(stream)>>>(x, y, idx, outer, g2);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

// dx (+)= (xm >= 0) * maxpool_backward(ug, idx): the ReLU backward of the layer in front of the pool applied in the same
// pass (xm = the buffer that ReLU read).  Same geometry restriction as ngb_maxpool_relu_fwd.
extern "C" int ngb_maxpool_relu_bwd(const float* ug, const uint8_t* idx, const float* xm, float* dx, int64_t outer, int64_t H,
                                    int64_t W, int64_t KH, int64_t KW, int64_t pad, int64_t stride, int accumulate,
                                    ngb_stream_t stream) {
  int64_t Ho, Wo;
  int rc = pool_geometry(H, W, KH, KW, pad, stride, &Ho, &Wo);
  if (rc) return rc;
  NGB_CHECK_ARG(idx != nullptr && xm != nullptr, "ngb_maxpool_relu_bwd needs the argmax indices and the relu input");
  Pool2Geo g2;
  if (outer * H * W == 0 || (H & 1) || !aligned16(xm) || !make_pool2_geo(&g2, outer, H, W, KH, KW, pad, stride, Ho, Wo, dx, dx))
    return fail(NGB_ERR_UNSUPPORTED, "maxpool+relu fusion: only 2x2 / stride 2 / pad 0, even H, 16-byte aligned rows");
  const size_t smem2 = (size_t)g2.G * g2.Wo * (g2.TP | 1) * 5 + 16;
  const int64_t items2 = ceil_div<int64_t>(outer, g2.G) * g2.strips;
  const int grid2 = (int)(items2 < (int64_t)kNumSMs * 8 ? items2 : (int64_t)kNumSMs * 8);
  maxpool2x2_bwd_kernel<true><<<grid2, 256, smem2, as_stream(stream)>>>(ug, idx, dx, outer, g2, accumulate, xm);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

// gm[e] = ug[e] * relu'(x at the window's argmax), per POOLED element (2x2 / stride 2): the ReLU+pool backward pair kept
// in its compact, pooled form for consumers that scatter it themselves (ngb_conv_wgrad_pooled) or only need its sum (bias
// gradient).  y > 0 means the argmax cell passed the ReLU with a positive value (mask 1); y == 0 means every cell of the
// window was clamped, the argmax is the first cell, and the mask is x_first >= 0 (activations.py:31) -- only those windows
// read x at all.
__global__ void __launch_bounds__(256) maxpool_relu_mask_kernel(const float* __restrict__ ug, const float* __restrict__ y,
                                                                const float* __restrict__ x, float* __restrict__ gm, int64_t total,
                                                                int H, int W, FastDiv d_plane, FastDiv d_hp) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const float g = ug[e];
    float out = g;
    if (!(y[e] > 0.f)) {
      uint32_t plane, fe, q, p;
      d_plane.divmod((uint32_t)e, plane, fe);       // pooled maps are stored with position (p', q') at q'*Hp + p'
      d_hp.divmod(fe, q, p);
      const float xv = __ldg(x + ((int64_t)plane * H + 2 * p) * W + 2 * q);
      out = xv >= 0.f ? g : 0.f;
    }
    gm[e] = out;
  }
}

extern "C" int ngb_maxpool_relu_mask(const float* ug, const float* y, const float* x, float* gm, int64_t outer, int64_t H,
                                     int64_t W, ngb_stream_t stream) {
  NGB_CHECK_ARG(outer >= 0 && H >= 2 && W >= 2 && (H % 2) == 0 && (W % 2) == 0, "ngb_maxpool_relu_mask: 2x2 / stride 2 pooling of even planes only");
  const int64_t total = outer * (H / 2) * (W / 2);
  if (total == 0) return NGB_OK;
  NGB_CHECK_ARG(total < (int64_t(1) << 32), "ngb_maxpool_relu_mask: tensor too large");
  maxpool_relu_mask_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(ug, y, x, gm, total, (int)H, (int)W,
                                                                               FastDiv((uint32_t)((H / 2) * (W / 2))), FastDiv((uint32_t)(H / 2)));
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_maxpool_fwd(const float* x, float* y, uint8_t* idx, int64_t outer, int64_t H, int64_t W, int64_t KH,
                               int64_t KW, int64_t pad, int64_t stride, ngb_stream_t stream) {
  int64_t Ho, Wo;
  int rc = pool_geometry(H, W, KH, KW, pad, stride, &Ho, &Wo);
  if (rc) return rc;
  if (outer * Ho * Wo == 0) return NGB_OK;
  Pool2Geo g2;
  if (make_pool2_geo(&g2, outer, H, W, KH, KW, pad, stride, Ho, Wo, x, x)) {
    const size_t smem2 = (size_t)g2.G * g2.Wo * (g2.TP | 1) * 5 + 16;
    const int64_t items2 = ceil_div<int64_t>(outer, g2.G) * g2.strips;
    const int grid2 = (int)(items2 < (int64_t)kNumSMs * 8 ? items2 : (int64_t)kNumSMs * 8);
    maxpool2x2_fwd_kernel<false><<<grid2, 256, smem2, as_stream(stream)>>>(x, y, idx, outer, g2);
    NGB_LAUNCH_CHECK();
    return NGB_OK;
  }
  PoolGeo g;
  rc = make_pool_geo(&g, outer, H, W, KH, KW, pad, stride, Ho, Wo, false);
  if (rc) {
    const int64_t total = outer * Ho * Wo;
    maxpool_fwd_simple_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(x, y, idx, outer, (int)H, (int)W, (int)Ho,
                                                                                   (int)Wo, (int)KH, (int)KW, (int)pad, (int)stride);
    NGB_LAUNCH_CHECK();
    return NGB_OK;
  }
  const size_t smem = (size_t)g.G * g.rows * g.Wp * 4;
  const int64_t items = ceil_div<int64_t>(outer, g.G) * g.strips;
  const int grid = (int)(items < (int64_t)kNumSMs * 8 ? items : (int64_t)kNumSMs * 8);
  maxpool_fwd_tiled_kernel<<<grid, 256, smem, as_stream(stream)>>>(x, y, idx, outer, g);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

extern "C" int ngb_maxpool_bwd(const float* ug, const uint8_t* idx, float* dx, int64_t outer, int64_t H, int64_t W,
                               int64_t KH, int64_t KW, int64_t pad, int64_t stride, int accumulate, ngb_stream_t stream) {
  int64_t Ho, Wo;
  int rc = pool_geometry(H, W, KH, KW, pad, stride, &Ho, &Wo);
  if (rc) return rc;
  NGB_CHECK_ARG(idx != nullptr, "ngb_maxpool_bwd needs the argmax indices saved by ngb_maxpool_fwd");
  if (outer * H * W == 0) return NGB_OK;
  Pool2Geo g2;
  if (make_pool2_geo(&g2, outer, H, W, KH, KW, pad, stride, Ho, Wo, dx, dx)) {
    const size_t smem2 = (size_t)g2.G * g2.Wo * (g2.TP | 1) * 5 + 16;
    const int64_t items2 = ceil_div<int64_t>(outer, g2.G) * g2.strips;
    const int grid2 = (int)(items2 < (int64_t)kNumSMs * 8 ? items2 : (int64_t)kNumSMs * 8);
    maxpool2x2_bwd_kernel<false><<<grid2, 256, smem2, as_stream(stream)>>>(ug, idx, dx, outer, g2, accumulate, nullptr);
    NGB_LAUNCH_CHECK();
    return NGB_OK;
  }
  PoolGeo g;
  rc = make_pool_geo(&g, outer, H, W, KH, KW, pad, stride, Ho, Wo, true);
  if (rc) {
    const int64_t total = outer * H * W;
    maxpool_bwd_simple_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(ug, idx, dx, outer, (int)H, (int)W, (int)Ho,
                                                                                   (int)Wo, (int)KH, (int)KW, (int)pad, (int)stride,
                                                                                   accumulate);
    NGB_LAUNCH_CHECK();
    return NGB_OK;
  }
  const size_t smem = (size_t)g.G * g.Wo * (g.TP | 1) * 5 + 16;
  const int64_t items = ceil_div<int64_t>(outer, g.G) * g.strips;
  const int grid = (int)(items < (int64_t)kNumSMs * 8 ? items : (int64_t)kNumSMs * 8);
  maxpool_bwd_tiled_kernel<<<grid, 256, smem, as_stream(stream)>>>(ug, idx, dx, outer, g, accumulate);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}
