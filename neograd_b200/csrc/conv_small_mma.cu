// Tensor-core convolution kernels for SMALL-CHANNEL layers (LeNet / digits family: C or O below the 32-channel
// granularity the tcgen05 implicit GEMM in conv_igemm_tcgen05.cu needs).
// Reference: neograd/autograd/ops/conv.py:245-269 (fprop), 299-309 (dgrad), 311-321 (wgrad).
//
// Why not tcgen05 here: with C = 1..8 input channels a filter tap contributes fewer than 8 K-elements, and a UMMA
// shared-memory descriptor can only describe operands whose rows are 16-byte chunks at a constant stride -- an im2col
// view of a 6-channel 12x12 image is not such an operand without first materialising it (25x the bytes).  The warp-level
// mma.sync.m16n8k8 TF32 instruction takes its fragments from registers, so every lane can GATHER its elements straight
// from the staged image:      A[f][k] = img_s[ aoff[f] + koff[k] ]
// (offset of output position f plus offset of tap/channel k -- the im2col matrix is never built).  These layers are
// memory/LSU bound (150..400 MACs per output), the tensor pipe only has to stop being the limiter: the CUDA-core
// kernels in conv_direct.cu ran at 8-10 TFLOP/s = 4 % of the HBM roofline of the layer.
//
// All kernels: a persistent CTA stages G whole images (TF32-rounded, zero halo for padding) in shared memory and walks
// its share of the batch.  Exact-fp32 mode (ngb_set_default_engine(NGB_GEMM_SIMT)) never comes here.
#include <array>
#include <map>
#include <mutex>

#include "common.cuh"

#include <type_traits>

namespace ngb {

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_global_pred(float* p, float v, bool pred) {   // @pred st.global.f32 [p], v
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q st.global.f32 [%0], %1;\n\t}" ::"l"(p),
      "f"(v), "r"((int)pred)
      : "memory");
}
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {     // compile-time loop: the index can be a template argument
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}
template <int IMM>
__device__ __forceinline__ uint32_t lds32_imm(uint32_t addr) {   // [reg + compile-time byte offset]: no address arithmetic
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

// Persistent CTAs own a CONTIGUOUS range of images (balanced to +-1 image over the grid) and walk it G images at a time:
// with pass-granular round robin, 512 passes over 444 resident CTAs meant two waves at 58 % efficiency.
__device__ __forceinline__ void cta_image_range(int N, int& lo, int& cnt) {
  const int per = N / (int)gridDim.x, rem = N % (int)gridDim.x;
  const int b = (int)blockIdx.x;
  lo = b * per + (b < rem ? b : rem);
  cnt = per + (b < rem ? 1 : 0);
}

// D(16x8) += A(16x8, row) * B(8x8, col); lane = 4*g + t:
//   a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  b0 (k=t, n=g) b1 (k=t+4, n=g);  d0 (g, 2t) d1 (g, 2t+1) d2 (g+8, 2t) d3 (g+8, 2t+1)
// The k labels are only names: both operands use the same permutation of the reduction index.
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------------ fprop / dgrad
// out[n][co][f] = bias[co] + sum_k img_s[n][aoff[f] + koff[k]] * Wp[k][co]
//   fprop: f = q*Ho + p (neograd order), k = (c,kh,kw);      img = x natural [c][h+pad][w+pad]
//   dgrad: f = h*W + w,               k = (kh,kw,o padded); img = ug un-scrambled [o][p+KH-1-pad][q+KW-1-pad]
struct GatherGeo {
  int N, Cin, Cout;
  int raw_rows, raw_cols, raw_plane;   // raw input plane as stored in HBM (cols contiguous)
  int transpose;                       // staged plane = raw plane transposed (dgrad: [q][p] -> [p][q])
  int halo_r, halo_c;                  // zero halo of the staged plane
  int spitch, splane, img_stride;      // staged row pitch, plane pitch, image pitch (floats)
  int F, F16, MTI;                     // outputs per (image, channel); padded to 16; m-tiles per image
  int fdiv, a_mod, a_div;              // aoff(f) = (f % fdiv) * a_mod + (f / fdiv) * a_div
  int K, Kp;                           // reduction length, padded to 8
  int NP, NPp;                         // Cout padded to 8; smem pitch of the packed weights (NP + 4)
  int mode;                            // 0 fprop, 1 dgrad
  int contig;                          // staged images == raw images (no halo, pitch == row length)
  int C, O, KH, KW, Op8;
  int G;                               // images per pass
};

// POOL (forward only, 16-position tiles that are exactly TWO rows of the output plane as the pool sees it, i.e. the plane's
// contiguous extent is 8): bias + ReLU + 2x2 / stride-2 max pool in the epilogue -- the convolution's output is never
// written.  A tile then holds four complete windows; window pw = (rows 2pw, 2pw+1, 2pw+8, 2pw+9) sits in the lane pair
// gq = 2pw, 2pw+1, which swap their two values with one shuffle each.  Values, argmax bytes and the ReLU flag (bit 2) are
// those of ngb_conv_fprop followed by ngb_maxpool_relu_fwd (pool.cu: first maximum in window order, flag = max > 0 or
// first cell >= 0).
template <int MT, int NT, bool ACC, bool POOL = false>
__global__ void __launch_bounds__(256, 4) conv_gather_mma_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              uint8_t* __restrict__ pool_idx,
                                                              GatherGeo g, FastDiv d_plane, FastDiv d_cols,
                                                              FastDiv d_fdiv, FastDiv d_mti, FastDiv d_npp, FastDiv d_op8,
                                                              FastDiv d_kw, FastDiv d_kk) {
  extern __shared__ __align__(16) uint32_t smem_u[];
  uint32_t* wp = smem_u;                         // [Kp/8][4 t][NPp][2]: (k = 8ks+t, k = 8ks+t+4) adjacent -> one LDS.64
  int* kpair = reinterpret_cast<int*>(wp + g.Kp * g.NPp);   // [Kp/8][4 t][2]
  int* aoff = kpair + g.Kp;                      // [F16]
  uint32_t* xs = reinterpret_cast<uint32_t*>(aoff + g.F16);  // [G][img_stride]  (16-byte aligned: all sizes are multiples of 4)
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const int KK = g.KH * g.KW;

  // ---- prologue: tables, packed weights, zero halo (once: staging only ever rewrites the interior)
  for (int e = tid; e < g.Kp * g.NPp; e += nthr) {
    const int half = e & 1;
    uint32_t r, n;
    d_npp.divmod((uint32_t)e >> 1, r, n);
    const int k = ((int)r >> 2) * 8 + ((int)r & 3) + 4 * half;
    float v = 0.f;
    if (g.mode == 0) {
      if (k < g.K && (int)n < g.O) v = __ldg(w + (int64_t)n * g.K + k);
    } else {
      uint32_t tap, o;
      d_op8.divmod((uint32_t)k, tap, o);
      if ((int)o < g.O && (int)n < g.C) v = __ldg(w + ((int64_t)o * g.C + n) * KK + tap);
    }
    wp[e] = to_tf32(v);
  }
  for (int k = tid; k < g.Kp; k += nthr) {
    int off = 0;
    if (g.mode == 0) {
      if (k < g.K) {
        uint32_t c, tap, kh, kw;
        d_kk.divmod((uint32_t)k, c, tap);
        d_kw.divmod(tap, kh, kw);
        off = (int)c * g.splane + (int)kh * g.spitch + (int)kw;
      }
    } else {
      uint32_t tap, o, kh, kw;
      d_op8.divmod((uint32_t)k, tap, o);
      d_kw.divmod(tap, kh, kw);
      if ((int)o < g.O) off = (int)o * g.splane + (g.KH - 1 - (int)kh) * g.spitch + (g.KW - 1 - (int)kw);
    }
    kpair[(k >> 3) * 8 + (k & 3) * 2 + ((k >> 2) & 1)] = off * 4;   // byte offset
  }
  for (int f = tid; f < g.F16; f += nthr) {
    int off = 0;
    if (f < g.F) {
      uint32_t hi, lo;
      d_fdiv.divmod((uint32_t)f, hi, lo);
      off = (int)lo * g.a_mod + (int)hi * g.a_div;
    }
    aoff[f] = off * 4;   // byte offset
  }
  if (g.halo_r | g.halo_c) {
    for (int e = tid; e < g.G * g.img_stride; e += nthr) xs[e] = 0u;
  }

  // per-thread constants of the fragment loads and of the epilogue
  const uint32_t xs_sa = smem_u32(xs);
  const uint32_t kp_sa = smem_u32(kpair) + t * 8;
  const uint32_t wp_sa = smem_u32(wp) + (uint32_t)(t * g.NPp + gq) * 8u;
  const uint32_t wstep = 32u * (uint32_t)g.NPp;          // one k-step of packed weights: 4 x NPp x 2 words
  const bool full_tiles = (g.F16 == g.F);
  int coff[NT][2];
  bool cvalid[NT][2];
  float bv[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int b2 = 0; b2 < 2; ++b2) {
      const int co = j * 8 + 2 * t + b2;
      cvalid[j][b2] = co < g.Cout;
      coff[j][b2] = co * g.F;
      bv[j][b2] = (bias != nullptr && co < g.Cout) ? __ldg(bias + co) : 0.f;
    }

  const int raw_img = g.Cin * g.raw_plane;
  int img_lo, img_cnt;
  cta_image_range(g.N, img_lo, img_cnt);
  for (int n0 = img_lo; n0 < img_lo + img_cnt; n0 += g.G) {
    const int ng = min(g.G, img_lo + img_cnt - n0);
    __syncthreads();   // previous pass's fragments are consumed (first pass: the prologue's zero fill is complete)
    // ---- stage ng images, rounded to TF32
    const float* src = in + (int64_t)n0 * raw_img;
    const int raw = ng * raw_img;
    if (g.contig) {
      // no halo, pitch == row length: the staged images ARE the raw images -> straight 128-bit copy + TF32 rounding
      const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
      const int raw4 = vec ? (raw >> 2) : 0;
#pragma unroll 4
      for (int e4 = tid; e4 < raw4; e4 += nthr) {
        const float4 v = ld_stream4(src + 4 * (int64_t)e4);
        *reinterpret_cast<uint4*>(xs + 4 * e4) = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      }
      for (int e = raw4 * 4 + tid; e < raw; e += nthr) xs[e] = to_tf32(__ldg(src + e));
    } else if (g.transpose) {
      // consecutive lanes walk the raw contiguous axis = the staged ROW axis: pitch = 4 (mod 8) keeps the stores spread
#pragma unroll 4
      for (int e = tid; e < raw; e += nthr) {
        uint32_t pl, f, r, c;
        d_plane.divmod((uint32_t)e, pl, f);
        d_cols.divmod(f, r, c);
        xs[pl * g.splane + (c + g.halo_r) * g.spitch + r + g.halo_c] = to_tf32(__ldg(src + e));
      }
    } else {
      const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
      const int raw4 = vec ? (raw >> 2) : 0;
#pragma unroll 2
      for (int e4 = tid; e4 < raw4; e4 += nthr) {
        const float4 v = ld_stream4(src + 4 * (int64_t)e4);
        uint32_t pl, f, r, c;
        d_plane.divmod((uint32_t)e4 * 4u, pl, f);
        d_cols.divmod(f, r, c);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xs[pl * g.splane + (r + g.halo_r) * g.spitch + c + g.halo_c] = to_tf32(vv[j]);
          if (++c == (uint32_t)g.raw_cols) {
            c = 0;
            if (++r == (uint32_t)g.raw_rows) { r = 0; ++pl; }
          }
        }
      }
      for (int e = raw4 * 4 + tid; e < raw; e += nthr) {
        uint32_t pl, f, r, c;
        d_plane.divmod((uint32_t)e, pl, f);
        d_cols.divmod(f, r, c);
        xs[pl * g.splane + (r + g.halo_r) * g.spitch + c + g.halo_c] = to_tf32(__ldg(src + e));
      }
    }
    __syncthreads();

    // ---- compute: a warp owns MT m-tiles (16 consecutive outputs of one image each) x all NT channel tiles.
    // All shared-memory operands are addressed by 32-bit BYTE addresses (tables hold byte offsets): one add per load.
    const int units = ng * g.MTI;
    float* const obase = POOL ? out : out + (int64_t)n0 * g.Cout * g.F;
    for (int u0 = warp * MT; u0 < units; u0 += nwarps * MT) {
      uint32_t A0[MT], A1[MT];
      int eo[MT], tile_f[MT];                   // element offset of (image, f0 + g) inside this pass's output block; f0
      {
        uint32_t img, tile;
        d_mti.divmod((uint32_t)u0, img, tile);
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const int f0 = (int)tile * 16;
          const uint32_t ib = xs_sa + (uint32_t)((int)img * g.img_stride) * 4u;
          A0[i] = ib + (uint32_t)aoff[f0 + gq];
          A1[i] = ib + (uint32_t)aoff[f0 + gq + 8];
          eo[i] = (u0 + i < units) ? (int)img * g.Cout * g.F + f0 + gq : -1;
          tile_f[i] = f0;
          if (++tile == (uint32_t)g.MTI) { tile = 0; ++img; }
          if (u0 + i + 1 >= units) { img = 0; tile = 0; }   // idle tail tiles recompute tile 0 (never stored)
        }
      }
      float acc[MT][NT][4];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

      const int ksteps = g.Kp >> 3;
      uint32_t kp = kp_sa, wq = wp_sa;
#pragma unroll 2
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint2 ko = lds64(kp);
        uint2 b[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) b[j] = lds64(wq + j * 64);
        kp += 32;
        wq += wstep;
        uint32_t a[MT][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) {          // all gathers first: 4*MT loads in flight before the first MMA needs one
          a[i][0] = lds32(A0[i] + ko.x); a[i][1] = lds32(A1[i] + ko.x);
          a[i][2] = lds32(A0[i] + ko.y); a[i][3] = lds32(A1[i] + ko.y);
        }
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) mma_tf32(acc[i][j], a[i][0], a[i][1], a[i][2], a[i][3], b[j].x, b[j].y);
      }

      if (POOL) {
        // pooled plane of one (image, channel): [Wp = F/32... ][Hp]: window (ph = tile within the image, pw = gq / 2) is stored
        // at flat index pw * Hp + ph (neograd order of the pool's own output), Hp = MTI
        const int Fp = g.F >> 2;
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const bool tile_ok = eo[i] >= 0;
          uint32_t img, tile;
          d_mti.divmod((uint32_t)(u0 + i), img, tile);
          const int64_t pbase = ((int64_t)(n0 + (int)img) * g.Cout) * Fp + (gq >> 1) * g.MTI + (int)tile;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) {
              const float v0 = acc[i][j][b2] + bv[j][b2], v1 = acc[i][j][2 + b2] + bv[j][b2];     // rows gq, gq + 8
              const float p0 = __shfl_xor_sync(0xffffffffu, v0, 4), p1 = __shfl_xor_sync(0xffffffffu, v1, 4);
              if ((gq & 1) == 0 && tile_ok && cvalid[j][b2]) {
                // window cells in the pool's row-major order: a = (row 2pw), b = (2pw+1), c = (2pw+8), d = (2pw+9)
                const float first = v0;
                const float a = fmaxf(0.f, v0), b = fmaxf(0.f, p0), c = fmaxf(0.f, v1), d = fmaxf(0.f, p1);
                float best = a;
                int arg = 0;
                if (b > best) { best = b; arg = 1; }
                if (c > best) { best = c; arg = 2; }
                if (d > best) { best = d; arg = 3; }
                if (best > 0.f || first >= 0.f) arg |= 4;      // kPoolReluFlag (pool.cu)
                const int64_t o = pbase + (int64_t)(j * 8 + 2 * t + b2) * Fp;
                out[o] = best;
                pool_idx[o] = (uint8_t)arg;
              }
            }
          }
        }
        continue;
      }
      // ---- epilogue: lanes of equal t hold 8 consecutive outputs of one channel -> 32-byte store segments.
      // One 64-bit pointer per (tile, channel column); the stores are PREDICATED (no branch per element).
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const bool tile_ok = eo[i] >= 0;
        const bool r0 = tile_ok && (full_tiles || tile_f[i] + gq < g.F), r1 = tile_ok && (full_tiles || tile_f[i] + gq + 8 < g.F);
        float* const pt = obase + (tile_ok ? eo[i] : 0);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
          for (int b2 = 0; b2 < 2; ++b2) {
            float* const pc = pt + coff[j][b2];
            const float bias_v = bv[j][b2];
            if (ACC) {
              if (cvalid[j][b2] && r0) pc[0] += acc[i][j][b2] + bias_v;
              if (cvalid[j][b2] && r1) pc[8] += acc[i][j][2 + b2] + bias_v;
            } else {
              st_global_pred(pc, acc[i][j][b2] + bias_v, cvalid[j][b2] && r0);
              st_global_pred(pc + 8, acc[i][j][2 + b2] + bias_v, cvalid[j][b2] && r1);
            }
          }
        }
      }
    }
  }
}

static inline int round_up_mod(int v, int m, int r) {   // smallest v' >= v with v' % m == r
  int d = (r - v % m + m) % m;
  return v + d;
}

// Images per pass and grid size of a persistent kernel.  Every CTA owns ceil(N/grid) or floor(N/grid) images; that range
// is walked in ceil(range/gmax) passes of (nearly) equal size G, so no CTA ends with a one-image straggler pass.  A
// smaller G shrinks the shared-memory footprint, which can raise occupancy and hence the grid: iterate to a fixed point.
template <typename Kernel>
static int balance_images(Kernel kernel, int threads, size_t fixed, size_t per_img, int N, int gmin, int gmax, bool corun, int* G_out,
                          int* grid_out) {
  int G = gmax < 1 ? 1 : gmax, grid = 1;
  for (int it = 0; it < 4; ++it) {
    int occ = 1;
    NGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, fixed + (size_t)G * per_img));
    if (occ < 1) occ = 1;
    if (corun && occ > 1) occ = (occ + 1) / 2;   // ngb_set_corun: leave half of every SM to the other lane's kernels
    grid = N < kNumSMs * occ ? N : kNumSMs * occ;
    const int ipc = ceil_div(N, grid);
    int Gn = ceil_div(ipc, ceil_div(ipc, gmax));
    if (Gn < gmin) Gn = gmin;
    if (Gn >= G) break;
    G = Gn;
  }
  *G_out = G;
  *grid_out = grid;
  return NGB_OK;
}

template <int MT, int NT, bool ACC, bool POOL = false>
static int launch_gather(const float* in, const float* w, const float* bias, float* out, GatherGeo g, int gmax, size_t fixed,
                         size_t per_img, cudaStream_t st, uint8_t* pool_idx = nullptr) {
  static bool attr = false;
  if (!attr) {
    NGB_CUDA(cudaFuncSetAttribute(conv_gather_mma_kernel<MT, NT, ACC, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    // maximum shared-memory carveout: a kernel of another lane can only become co-resident on an SM whose L1 / shared split
    // already has room for it -- with the default (smallest fitting) split the second kernel waits for the first to drain
    NGB_CUDA(cudaFuncSetAttribute(conv_gather_mma_kernel<MT, NT, ACC, POOL>, cudaFuncAttributePreferredSharedMemoryCarveout, getenv("NGB_CARVEOUT_DEFAULT") ? cudaSharedmemCarveoutDefault : cudaSharedmemCarveoutMaxShared));
    attr = true;
  }
  int grid = 1;
  int rc = balance_images(conv_gather_mma_kernel<MT, NT, ACC, POOL>, 256, fixed, per_img, g.N, 1, gmax, corun_hint() != 0, &g.G, &grid);
  if (rc) return rc;
  // 8 warps at <= 64 registers: four CTAs (32 warps) per SM.  (9-12 warp blocks tile some pass sizes without an idle
  // tail, but their register footprint drops the SM to two CTAs, which costs more than the tail.)
  const int nwarps = 8;
  const size_t smem = fixed + (size_t)g.G * per_img;
  conv_gather_mma_kernel<MT, NT, ACC, POOL><<<grid, nwarps * 32, smem, st>>>(
      in, w, bias, out, pool_idx, g, FastDiv((uint32_t)g.raw_plane), FastDiv((uint32_t)g.raw_cols), FastDiv((uint32_t)g.fdiv),
      FastDiv((uint32_t)g.MTI), FastDiv((uint32_t)g.NPp), FastDiv((uint32_t)g.Op8), FastDiv((uint32_t)g.KW),
      FastDiv((uint32_t)(g.KH * g.KW)));
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

// mode 0 = fprop (in = x, out = y), mode 1 = dgrad (in = ug, out = dx).  *done tells whether the kernel was launched.
// pool_idx != nullptr (forward only): `out` is the 2x2-pooled map and pool_idx its argmax bytes (POOL epilogue).
int conv_mma_gather(const float* in, const float* w, const float* bias, float* out, int N, int C, int H, int W, int O, int KH,
                    int KW, int pad, int stride, int Ho, int Wo, int mode, int accumulate, cudaStream_t st, bool* done,
                    uint8_t* pool_idx) {
  *done = false;
  if (g_default_engine.load() == NGB_GEMM_SIMT) return NGB_OK;
  if (pool_idx && (mode != 0 || accumulate || Ho != 8 || Wo != 8)) return NGB_OK;     // a 16-position tile = two rows of the 8 x 8 plane
  GatherGeo g;
  g.N = N; g.C = C; g.O = O; g.KH = KH; g.KW = KW; g.mode = mode;
  g.Op8 = (O + 7) & ~7;
  int srows;
  if (mode == 0) {
    if (O < 4 || O > 32) return NGB_OK;                   // one channel tile is 8 wide: below 4 the CUDA-core kernels win
    g.Cin = C; g.Cout = O;
    g.raw_rows = H; g.raw_cols = W; g.transpose = 0;
    g.halo_r = pad; g.halo_c = pad;
    srows = H + 2 * pad;
    g.spitch = round_up_mod(W + 2 * pad, 8, 4);
    g.splane = srows * g.spitch;
    g.F = Ho * Wo; g.fdiv = Ho; g.a_mod = stride * g.spitch; g.a_div = stride;
    g.K = C * KH * KW;
  } else {
    if (stride != 1 || pad > KH - 1 || pad > KW - 1) return NGB_OK;
    if (C < 4 || C > 32) return NGB_OK;
    g.Cin = O; g.Cout = C;
    g.raw_rows = Wo; g.raw_cols = Ho; g.transpose = 1;   // raw ug plane is [q][p]
    g.halo_r = KH - 1 - pad; g.halo_c = KW - 1 - pad;
    srows = Ho + 2 * g.halo_r;
    g.spitch = round_up_mod(Wo + 2 * g.halo_c, 8, 4);
    g.splane = round_up_mod(srows * g.spitch, 32, 8);    // lanes t walk the planes: 8 banks apart
    g.F = H * W; g.fdiv = W; g.a_mod = 1; g.a_div = g.spitch;
    g.K = KH * KW * g.Op8;
  }
  g.raw_plane = g.raw_rows * g.raw_cols;
  g.contig = (!g.transpose && g.halo_r == 0 && g.halo_c == 0 && g.spitch == g.raw_cols && g.splane == g.raw_plane) ? 1 : 0;
  g.Kp = (g.K + 7) & ~7;
  g.NP = (g.Cout + 7) & ~7;
  g.NPp = g.NP + 4;
  g.MTI = ceil_div(g.F, 16);
  g.F16 = g.MTI * 16;
  g.img_stride = g.Cin * g.splane;   // splane is a multiple of 4
  const int NT = g.NP / 8;
  if (NT != 1 && NT != 2 && NT != 4) return NGB_OK;
  const size_t fixed = ((size_t)g.Kp * g.NPp + g.Kp + g.F16) * 4;
  const size_t per_img = (size_t)g.img_stride * 4;
  const size_t budget = 72 * 1024;
  if (fixed + per_img > 96 * 1024 || g.Kp > 4096) return NGB_OK;
  int G = 8;
  while (G > 1 && fixed + G * per_img > budget) --G;
  if (G > N) G = N;
  // m-tiles per warp: 4 (2 when there are four channel tiles), unless an image has too few tiles to feed 8 warps
  const int MTmax = NT == 4 ? 2 : 4;
  const int best_mt = (MTmax == 4 && G * g.MTI < 32) ? 2 : MTmax;
  int rc = ensure_init();
  if (rc) return rc;
#define NGB_GATHER(MT_, NT_)                                                                              \
  rc = pool_idx ? launch_gather<MT_, NT_, false, true>(in, w, bias, out, g, G, fixed, per_img, st, pool_idx) \
     : accumulate ? launch_gather<MT_, NT_, true>(in, w, bias, out, g, G, fixed, per_img, st)             \
                  : launch_gather<MT_, NT_, false>(in, w, bias, out, g, G, fixed, per_img, st)
  if (NT == 1) { if (best_mt == 4) NGB_GATHER(4, 1); else NGB_GATHER(2, 1); }
  else if (NT == 2) { if (best_mt == 4) NGB_GATHER(4, 2); else NGB_GATHER(2, 2); }
  else NGB_GATHER(2, 4);
#undef NGB_GATHER
  if (rc) return rc;
  *done = true;
  return NGB_OK;
}

// ------------------------------------------------------------------------------------------------ dgrad, row-decomposed
// dx[n][c][h][w] = sum_{o,kh,kw} ug[n][o][p = h+pad-kh][q = w+pad-kw] * W[o][c][kh][kw]        (conv.py:299-309, stride 1)
//
// The gather kernel above needs 4 shared-memory loads per MMA for this pass: its N dimension is C (6 -> 8), so a gathered
// A fragment is used once, and the full-correlation zero halo makes it multiply 2.25x the useful taps.  Here the sum is
// split as   T[(c,kh)][(p,w)] = sum_{o,kw} W[o][c][kh][kw] * ug[o][p][w+pad-kw]      (one GEMM per image)
//            dx[c][h][w]      = sum_kh T[(c,kh)][(h+pad-kh, w)]                        (shift-add over kh)
// so that M = (c, kh) is FULL (two channels x 8 kh slots per 16-row tile), the A operand is the filter bank --
// K = KW*O (80) values per row, held in registers for the whole kernel -- and the only gathered operand is B:
// 2 loads per MMA, no vertical halo, no table lookups inside the k loop (all k offsets are warp-uniform).
// A warp owns (channel pair j, image slot): it walks the image's position tiles, and adds every finished 16x8 tile into
// the image's dx accumulator in shared memory at row offset kh (only this warp touches channels 2j, 2j+1 of the slot:
// plain read-modify-write, deterministic).  The accumulators are then written to HBM with coalesced 128-bit stores.
struct RowsGeo {
  int N, C, O, H, W, Ho, Wo, KH, pad;
  int hl, P, plane, uimg;      // staged ug: left halo, row pitch, plane pitch, image pitch (floats)
  int NPOS, NT8;               // Ho*W positions per image, tiles of 8
  int HWp, dimg;               // dx accumulator plane pitch (even) and image pitch
  int MTC, IP;                 // m-tiles of (c,kh) rows, images in flight per CTA
  int pair_ok;                 // 64-bit read-modify-write possible (W and HWp even)
  int pooled, Hp, Fp;          // ug is a 2x2-pooled gradient (+ pool output, argmax indices, conv output): masked and scattered on the way in
};

template <int KW, int OB>
__global__ void __launch_bounds__(384) conv_dgrad_rows_mma_kernel(const float* __restrict__ ug, const float* __restrict__ w,
                                                                  const uint8_t* __restrict__ pidx, const float* __restrict__ ypool,
                                                                  const float* __restrict__ zmask, FastDiv d_fp, FastDiv d_hp,
                                                                  float* __restrict__ dx, RowsGeo g, int accumulate,
                                                                  FastDiv d_w, FastDiv d_howo, FastDiv d_ho, FastDiv d_hw,
                                                                  FastDiv d_o, FastDiv d_c, FastDiv d_howo4, FastDiv d_ho4) {
  extern __shared__ __align__(16) uint32_t smem_u[];
  constexpr int KS = KW * OB;
  int* boff = reinterpret_cast<int*>(smem_u);               // [NT8*8] byte offset of position n in a staged plane
  int* prow = boff + g.NT8 * 8;                             // [NT8*8] p of position n (for the h range check)
  uint32_t* us = reinterpret_cast<uint32_t*>(prow + g.NT8 * 8);   // [IP][O8][plane]
  float* dxs = reinterpret_cast<float*>(us + g.IP * g.uimg);      // [IP][C][HWp]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const int j = warp % g.MTC, slot = warp / g.MTC;
  const int KK = g.KH * KW;

  // ---- stationary A fragments: rows are the (c, kh) pairs packed densely, r = c*KH + kh (tile j holds rows 16j..16j+15),
  // k-step ks = kw*OB + ob, labels t / t+4 -> o = 8ob + t (+4)
  const int rowa = 16 * j + gq, rowb = rowa + 8;             // this lane's two rows (fragment rows g and g+8)
  const int ca = rowa / g.KH, kha = rowa - ca * g.KH, cb = rowb / g.KH, khb = rowb - cb * g.KH;
  const bool oka = ca < g.C, okb = cb < g.C;
  uint32_t a[KS][4];
#pragma unroll
  for (int kw = 0; kw < KW; ++kw)
#pragma unroll
    for (int ob = 0; ob < OB; ++ob) {
      const int ks = kw * OB + ob;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = (e & 1) ? cb : ca, kh = (e & 1) ? khb : kha;
        const int o = ob * 8 + t + ((e & 2) ? 4 : 0);
        float v = 0.f;
        if (c < g.C && o < g.O) v = __ldg(w + ((int64_t)o * g.C + c) * KK + kh * KW + kw);
        a[ks][e] = to_tf32(v);
      }
    }
  for (int n = tid; n < g.NT8 * 8; n += nthr) {
    int off = 0, p = 0;
    if (n < g.NPOS) {
      uint32_t pp, ww;
      d_w.divmod((uint32_t)n, pp, ww);
      p = (int)pp;
      off = ((int)pp * g.P + (int)ww) * 4;
    }
    boff[n] = off;
    prow[n] = p;
  }
  // zero once: halo columns, padded channel planes (never rewritten) and the dx accumulators
  for (int e = tid; e < g.IP * (g.uimg + g.dimg); e += nthr) us[e] = 0u;
  __syncthreads();

  const uint32_t us_sa = smem_u32(us) + (uint32_t)(slot * g.uimg + t * g.plane) * 4u;
  const uint32_t bo_sa = smem_u32(boff) + (uint32_t)gq * 4u;
  const uint32_t dx_sa = smem_u32(dxs) + (uint32_t)(slot * g.dimg) * 4u;
  const uint32_t plane_b = (uint32_t)g.plane * 4u;
  const int HoWo = g.Ho * g.Wo;
  int img_lo, img_cnt;
  cta_image_range(g.N, img_lo, img_cnt);
  for (int n0 = img_lo; n0 < img_lo + img_cnt; n0 += g.IP) {
    const int ng = min(g.IP, img_lo + img_cnt - n0);
    if (g.pooled == 2) {
      // ---- POOLED gradient, square plane, four pooled rows per 128-bit load: the thread that owns a window writes its whole
      // 2x2 cell block (masked value at the argmax, zeros elsewhere) -- no zeroing pass, no barrier, one index decomposition
      // per four windows.  Stored cell (r, c) = (2p' + kh, 2q' + kw) is conv position (p, q) = (c, r): staged row c, column r.
      const int64_t base = (int64_t)n0 * g.O * g.Fp;
      const int raw4 = (ng * g.O * g.Fp) >> 2;
#pragma unroll 2
      for (int e4 = tid; e4 < raw4; e4 += nthr) {
        uint32_t pl, fe, qq, pp, img, o;
        d_fp.divmod((uint32_t)e4 * 4u, pl, fe);
        d_hp.divmod(fe, qq, pp);
        d_o.divmod(pl, img, o);
        const float4 gv = ld_stream4(ug + base + 4 * (int64_t)e4);
        const uint32_t a4 = __ldg(reinterpret_cast<const uint32_t*>(pidx + base + 4 * (int64_t)e4));
        uint32_t* row0 = us + img * g.uimg + o * g.plane + (2 * qq) * g.P + g.hl + 2 * pp;     // staged row 2q', column 2p'
        const float gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
          const uint32_t ab = (a4 >> (8 * j2)) & 0xFFu, a = ab & 3u;
          const uint32_t tv = to_tf32((ab & 4u) ? gs[j2] : 0.f);
          // argmax cell a = kh*2 + kw sits at staged (row 2q' + kw, column 2p' + kh)
          *reinterpret_cast<uint2*>(row0 + 2 * j2) = make_uint2(a == 0u ? tv : 0u, a == 2u ? tv : 0u);
          *reinterpret_cast<uint2*>(row0 + g.P + 2 * j2) = make_uint2(a == 1u ? tv : 0u, a == 3u ? tv : 0u);
        }
      }
    } else if (g.pooled) {
      // ---- ug arrives as the POOLED gradient of a ReLU + 2x2/s2 MaxPool behind this convolution: zero the interiors, then
      // every pooled element is masked (relu') and dropped at its window's argmax cell; the full gradient never exists in HBM
      for (int e = tid; e < ng * g.O * g.Ho; e += nthr) {          // interior rows [hl, hl + Wo) of every (img, o, p)
        uint32_t pl, p, img, o;
        d_ho.divmod((uint32_t)e, pl, p);
        d_o.divmod(pl, img, o);
        uint32_t* row = us + img * g.uimg + o * g.plane + p * g.P + g.hl;
        for (int q = 0; q < g.Wo; ++q) row[q] = 0u;
      }
      __syncthreads();
      const int64_t base = (int64_t)n0 * g.O * g.Fp;
      const int raw = ng * g.O * g.Fp;
#pragma unroll 2
      for (int e = tid; e < raw; e += nthr) {
        uint32_t pl, fe, qq, pp, img, o;
        d_fp.divmod((uint32_t)e, pl, fe);                 // pl = img*O + o; pooled position (p', q') sits at q'*Hp + p'
        d_hp.divmod(fe, qq, pp);
        d_o.divmod(pl, img, o);
        const uint32_t ab = pidx[base + e];                // bits 0-1: argmax cell, bit 2: relu'(z at the argmax)
        const float gv = (ab & 4u) ? __ldg(ug + base + e) : 0.f;
        const uint32_t a = ab & 3u;
        const uint32_t fx = (2 * pp + (a >> 1)) * (uint32_t)g.Wo + 2 * qq + (a & 1);   // flat index in the stored (Ho, Wo) plane
        uint32_t q, p;
        d_ho.divmod(fx, q, p);                            // ... which the convolution reads as position (p, q) = q*Ho + p
        us[img * g.uimg + o * g.plane + p * g.P + g.hl + q] = to_tf32(gv);
      }
    } else
    // ---- stage: global [o][q][p] (p contiguous) -> smem [o][p][hl + q].  The staging costs issue slots, not bandwidth
    // (it was as many instructions as the MMA phase): 128-bit loads along p, one index decomposition per four elements.
    {
      const float* src = ug + (int64_t)n0 * g.O * HoWo;
      const int raw = ng * g.O * HoWo;
      if ((g.Ho & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const int pl_stride_pad = (g.uimg - g.O * g.plane);       // gap between images in the staged copy (padded planes)
#pragma unroll 2
        for (int e4 = tid; e4 < (raw >> 2); e4 += nthr) {
          const float4 v = ld_stream4(src + 4 * (int64_t)e4);
          uint32_t pl, f4, q, p4, img, o;
          d_howo4.divmod((uint32_t)e4, pl, f4);                     // pl = img*O + o
          d_ho4.divmod(f4, q, p4);
          d_o.divmod(pl, img, o);
          uint32_t* d = us + pl * g.plane + img * pl_stride_pad + (p4 * 4) * g.P + g.hl + q;
          d[0] = to_tf32(v.x); d[g.P] = to_tf32(v.y); d[2 * g.P] = to_tf32(v.z); d[3 * g.P] = to_tf32(v.w);
        }
      } else {
#pragma unroll 4
        for (int e = tid; e < raw; e += nthr) {
          uint32_t pl, f, q, p;
          d_howo.divmod((uint32_t)e, pl, f);                 // pl = img*O + o
          d_ho.divmod(f, q, p);
          uint32_t img, o;
          d_o.divmod(pl, img, o);
          us[img * g.uimg + o * g.plane + p * g.P + g.hl + q] = to_tf32(__ldg(src + e));
        }
      }
    }
    __syncthreads();

    if (slot < ng) {
      // Every m-tile (warp j) adds into ITS OWN copy of the image's dx accumulator, so no two warps ever touch the same
      // word (a channel's kh rows may straddle two tiles); the copies are summed when the image is written out.
      float* const dx_j = dxs + (slot * g.MTC + j) * g.C * g.HWp;
      float* const dxa = dx_j + ca * g.HWp + (kha - g.pad) * g.W + 2 * t;    // + n0: element (c, h = p + kh - pad, w)
      float* const dxb = dx_j + cb * g.HWp + (khb - g.pad) * g.W + 2 * t;
      // no padding, whole tiles, even W: every (row, position) of a tile lands inside the plane, 8-byte aligned
      const bool fast = g.pad == 0 && (g.NPOS & 7) == 0 && g.pair_ok;
      uint32_t base[OB][2];                                  // planes of o = 8ob + t and o + 4 (byte addresses)
#pragma unroll
      for (int ob = 0; ob < OB; ++ob) {
        base[ob][0] = us_sa + (uint32_t)(ob * 8) * plane_b;
        base[ob][1] = base[ob][0] + 4u * plane_b;
      }
      auto shift_add = [&](const float (&acc)[4], int nt) {
        // T[(c,kh)][n] -> dxs[c][(p + kh - pad)*W + w] = dxs[c][(kh - pad)*W + n]
        if (fast) {
          if (oka) { float2* q2 = reinterpret_cast<float2*>(dxa + nt * 8); float2 o2 = *q2; o2.x += acc[0]; o2.y += acc[1]; *q2 = o2; }
          if (okb) { float2* q2 = reinterpret_cast<float2*>(dxb + nt * 8); float2 o2 = *q2; o2.x += acc[2]; o2.y += acc[3]; *q2 = o2; }
        } else {
          const int nb = nt * 8 + 2 * t;
          const int p0 = prow[nb], p1 = prow[nb + 1];
          const bool n0ok = nb < g.NPOS, n1ok = nb + 1 < g.NPOS;
          if (oka) {
            const int h0 = p0 + kha - g.pad, h1 = p1 + kha - g.pad;
            if (n0ok && h0 >= 0 && h0 < g.H) dxa[nt * 8] += acc[0];
            if (n1ok && h1 >= 0 && h1 < g.H) dxa[nt * 8 + 1] += acc[1];
          }
          if (okb) {
            const int h0 = p0 + khb - g.pad, h1 = p1 + khb - g.pad;
            if (n0ok && h0 >= 0 && h0 < g.H) dxb[nt * 8] += acc[2];
            if (n1ok && h1 >= 0 && h1 < g.H) dxb[nt * 8 + 1] += acc[3];
          }
        }
        __syncwarp();   // the next tile's read-modify-write may hit addresses other lanes just wrote
      };
      // two position tiles in flight: independent accumulator chains hide the MMA latency
      for (int nt = 0; nt < g.NT8; nt += 2) {
        const bool two = nt + 1 < g.NT8;
        const uint32_t bo0 = lds32(bo_sa + (uint32_t)nt * 32u);
        const uint32_t bo1 = two ? lds32(bo_sa + (uint32_t)(nt + 1) * 32u) : bo0;
        uint32_t p0[OB][2], p1[OB][2];
#pragma unroll
        for (int ob = 0; ob < OB; ++ob) {
          p0[ob][0] = base[ob][0] + bo0; p0[ob][1] = base[ob][1] + bo0;
          p1[ob][0] = base[ob][0] + bo1; p1[ob][1] = base[ob][1] + bo1;
        }
        float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
        static_for<0, KW>([&](auto kwc) {
          constexpr int kw = decltype(kwc)::value;
          constexpr int IMM = (KW - 1 - kw) * 4;               // column shift of this tap: an immediate of the load
#pragma unroll
          for (int ob = 0; ob < OB; ++ob) {
            const uint32_t b00 = lds32_imm<IMM>(p0[ob][0]), b01 = lds32_imm<IMM>(p0[ob][1]);
            const uint32_t b10 = lds32_imm<IMM>(p1[ob][0]), b11 = lds32_imm<IMM>(p1[ob][1]);
            const uint32_t(&ak)[4] = a[kw * OB + ob];
            mma_tf32(acc0, ak[0], ak[1], ak[2], ak[3], b00, b01);
            mma_tf32(acc1, ak[0], ak[1], ak[2], ak[3], b10, b11);
          }
        });
        shift_add(acc0, nt);
        if (two) shift_add(acc1, nt + 1);
      }
    }
    __syncthreads();
    // ---- write the accumulators out (coalesced; the per-m-tile copies are summed in a fixed order) and clear them
    {
      const int HW = g.H * g.W;
      float* dst = dx + (int64_t)n0 * g.C * HW;
      const int total = ng * g.C * HW;
      const int copy = g.C * g.HWp;                          // one m-tile's copy of an image
      if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        for (int e4 = tid; e4 < (total >> 2); e4 += nthr) {
          uint32_t pl, f, img, c;
          d_hw.divmod((uint32_t)e4 * 4u, pl, f);            // pl = img*C + c
          d_c.divmod(pl, img, c);
          float* sp = dxs + (img * g.MTC) * copy + c * g.HWp + f;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int m = 0; m < g.MTC; ++m) {
            float4* s4 = reinterpret_cast<float4*>(sp + m * copy);
            const float4 u = *s4;
            *s4 = make_float4(0.f, 0.f, 0.f, 0.f);
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
          }
          float4* gp = reinterpret_cast<float4*>(dst + 4 * (int64_t)e4);
          if (accumulate) { const float4 o4 = *gp; v.x += o4.x; v.y += o4.y; v.z += o4.z; v.w += o4.w; }
          *gp = v;
        }
      } else {
        for (int e = tid; e < total; e += nthr) {
          uint32_t pl, f, img, c;
          d_hw.divmod((uint32_t)e, pl, f);
          d_c.divmod(pl, img, c);
          float* sp = dxs + (img * g.MTC) * copy + c * g.HWp + f;
          float v = 0.f;
          for (int m = 0; m < g.MTC; ++m) { v += sp[m * copy]; sp[m * copy] = 0.f; }
          dst[e] = accumulate ? dst[e] + v : v;
        }
      }
    }
    // (the next pass's staging writes `us`, which nobody reads any more; the __syncthreads after it orders dxs too)
  }
}

template <int KW, int OB>
static int launch_dgrad_rows(const float* ug, const float* w, const uint8_t* pidx, const float* ypool, const float* zmask, float* dx,
                             const RowsGeo& g, int accumulate, size_t smem, cudaStream_t st) {
  // (images per pass = image slots of the block, fixed by its warp count: only the grid is balanced here)
  static bool attr = false;
  if (!attr) {
    NGB_CUDA(cudaFuncSetAttribute(conv_dgrad_rows_mma_kernel<KW, OB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    // maximum shared-memory carveout: a kernel of another lane can only become co-resident on an SM whose L1 / shared split
    // already has room for it -- with the default (smallest fitting) split the second kernel waits for the first to drain
    NGB_CUDA(cudaFuncSetAttribute(conv_dgrad_rows_mma_kernel<KW, OB>, cudaFuncAttributePreferredSharedMemoryCarveout, getenv("NGB_CARVEOUT_DEFAULT") ? cudaSharedmemCarveoutDefault : cudaSharedmemCarveoutMaxShared));
    attr = true;
  }
  const int threads = g.MTC * g.IP * 32;
  int occ = 1;
  NGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_dgrad_rows_mma_kernel<KW, OB>, threads, smem));
  if (occ < 1) occ = 1;
  if (corun_hint() && occ > 1) occ = (occ + 1) / 2;
  const int64_t passes = ceil_div<int64_t>(g.N, g.IP);
  const int grid = (int)(passes < (int64_t)kNumSMs * occ ? passes : (int64_t)kNumSMs * occ);
  conv_dgrad_rows_mma_kernel<KW, OB><<<grid, threads, smem, st>>>(ug, w, pidx, ypool, zmask, FastDiv((uint32_t)(g.Fp > 0 ? g.Fp : 1)),
                                                                   FastDiv((uint32_t)(g.Hp > 0 ? g.Hp : 1)), dx, g, accumulate, FastDiv((uint32_t)g.W),
                                                                   FastDiv((uint32_t)(g.Ho * g.Wo)), FastDiv((uint32_t)g.Ho),
                                                                   FastDiv((uint32_t)(g.H * g.W)), FastDiv((uint32_t)g.O), FastDiv((uint32_t)g.C),
                                                                   FastDiv((uint32_t)((g.Ho * g.Wo) / 4 > 0 ? (g.Ho * g.Wo) / 4 : 1)),
                                                                   FastDiv((uint32_t)(g.Ho / 4 > 0 ? g.Ho / 4 : 1)));
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

int conv_mma_dgrad_rows(const float* ug, const float* w, float* dx, int N, int C, int H, int W, int O, int KH, int KW, int pad,
                        int stride, int Ho, int Wo, int accumulate, cudaStream_t st, bool* done, const uint8_t* pidx,
                        const float* ypool, const float* zmask) {
  *done = false;
  if (g_default_engine.load() == NGB_GEMM_SIMT) return NGB_OK;
  if (pidx && ((Ho & 1) || (Wo & 1))) return NGB_OK;
  if (stride != 1 || pad > KW - 1 || pad > KH - 1 || KH > 8 || C < 3 || C > 8 || W < 8) return NGB_OK;
  const int OB = (O + 7) / 8;
  if (!((KW == 5 || KW == 3) && (OB == 1 || OB == 2)) && !(KW == 3 && OB == 4)) return NGB_OK;
  RowsGeo g;
  g.N = N; g.C = C; g.O = O; g.H = H; g.W = W; g.Ho = Ho; g.Wo = Wo; g.KH = KH; g.pad = pad;
  g.pooled = pidx ? 1 : 0; g.Hp = Ho / 2; g.Fp = (Ho / 2) * (Wo / 2);
  g.hl = KW - 1 - pad;
  // whole-window staging (kernel comment): square plane, four pooled rows per 128-bit load, 8-byte aligned cell pairs
  if (pidx && Ho == Wo && (g.Hp & 3) == 0 && (g.hl & 1) == 0 && ((uintptr_t)ug & 15u) == 0 && ((uintptr_t)pidx & 3u) == 0 &&
      !getenv("NGB_DGRAD_NO_CELL"))
    g.pooled = 2;
  g.P = round_up_mod(Wo + 2 * g.hl, 8, 4);
  g.plane = round_up_mod(Ho * g.P, 32, 8);
  g.uimg = OB * 8 * g.plane;
  g.NPOS = Ho * W;
  g.NT8 = ceil_div(g.NPOS, 8);
  g.HWp = (H * W + 3) & ~3;
  g.MTC = ceil_div(C * KH, 16);                  // (c, kh) rows packed densely into 16-row tiles
  g.dimg = g.MTC * C * g.HWp;                    // one accumulator copy per m-tile
  g.pair_ok = (W % 2 == 0) ? 1 : 0;
  // the last position tile reads up to 7 positions past NPOS at offset 0 (valid); every real read stays inside the plane:
  // max col = W - 1 + KW - 1 = Wo + 2*hl - 1 < P
  const size_t fixed = (size_t)g.NT8 * 8 * 2 * 4;
  const size_t per_img = ((size_t)g.uimg + g.dimg) * 4;
  int IP = 12 / g.MTC;                                   // up to 12 warps
  while (IP > 1 && fixed + IP * per_img > 64 * 1024) --IP;
  if (IP > N) IP = N;
  if (fixed + IP * per_img > 96 * 1024) return NGB_OK;
  if (IP >= 3 && fixed + 3 * per_img <= 40 * 1024) IP = 3;   // smaller CTAs, more of them per SM
  g.IP = IP;
  const size_t smem = fixed + (size_t)IP * per_img;
  int rc = ensure_init();
  if (rc) return rc;
#define NGB_ROWS(KW_, OB_) rc = launch_dgrad_rows<KW_, OB_>(ug, w, pidx, ypool, zmask, dx, g, accumulate, smem, st)
  if (KW == 5 && OB == 1) NGB_ROWS(5, 1);
  else if (KW == 5 && OB == 2) NGB_ROWS(5, 2);
  else if (KW == 3 && OB == 1) NGB_ROWS(3, 1);
  else if (KW == 3 && OB == 2) NGB_ROWS(3, 2);
  else NGB_ROWS(3, 4);
#undef NGB_ROWS
  if (rc) return rc;
  *done = true;
  return NGB_OK;
}

// ------------------------------------------------------------------------------------------------ wgrad
// dw[o][m] = sum_n sum_f ug[n][o][f] * xs[n][moff[m] + poff[f]],   m = (c,kh,kw) (dw's own inner layout), f = q*Ho + p.
// GEMM view: D[m][o], reduction over the output positions of every image.  The 8 positions of a k-step are consecutive
// f; lane label t stands for position f0+2t and label t+4 for f0+2t+1, so the ug fragment is ONE 64-bit load and the
// x fragment rows sit 2 staged rows (= 8 banks mod 32) apart.  Warps = MS row splits x KS position splits; partial
// tiles are summed across KS in a fixed order, per-CTA partials by the deterministic second pass.
struct WgradGeo {
  int N, C, O, KH, KW, Ho;
  int raw_rows, raw_cols, raw_plane, halo;   // x plane
  int sr, sc, splane, ximg;                  // staged x: element (c, r, col) at c*splane + r*sr + col*sc (layout search below)
  int khfast;                                // A rows of one channel ordered kw*KH + kh instead of kh*KW + kw
  int F, F8, KSI, Up, NP, uimg;              // staged ug: [NP][Up] per image, zero tail
  int CKK, Mrows;                            // filter elements per output channel; rows covered by the warps (MS*MTW*16)
  int a_mod, a_div;                          // poff(f) = (f % Ho) * a_mod + (f / Ho) * a_div
  int MS, KS, G;
  int xcontig;                               // staged x == raw x (no halo, pitch == row length)
  int pooled, Wo, Fp, Hp;                    // ug is a 2x2-pooled gradient + argmax indices: scattered into the staged tile
};

// PF: register-prefetched staging for the pooled form (host-checked: xcontig, 16-byte aligned operands, Hp % 4 == 0, one group's
// x <= kPFX and its pooled gradient <= kPFU 128-bit units per thread).  The global loads of group i+1 are issued before the MMA
// phase of group i and land in registers while the tensor pipe works; the pooled windows are written as whole 2x2 cells
// (value at the argmax, zeros elsewhere), so the tile needs no separate zeroing pass -- two barriers per group, and no global
// latency between them.  (ncu, non-PF: 36 % / 22 % issue utilisation, warps parked on long-scoreboard and barrier stalls.)
constexpr int kPFX = 4, kPFU = 2;
template <int MTW, int NT, bool PF>
__global__ void __launch_bounds__(256) conv_wgrad_mma_kernel(const float* __restrict__ ug, const float* __restrict__ x,
                                                             const uint8_t* __restrict__ pidx, const float* __restrict__ ypool,
                                                             const float* __restrict__ zmask, float* __restrict__ part, WgradGeo g,
                                                             FastDiv d_fp, FastDiv d_hp, FastDiv d_plane,
                                                             FastDiv d_cols, FastDiv d_ho, FastDiv d_ksi, FastDiv d_kk,
                                                             FastDiv d_kw, FastDiv d_f, FastDiv d_ckk, FastDiv d_o,
                                                             FastDiv d_fp4, FastDiv d_wp, FastDiv d_kh) {
  extern __shared__ __align__(16) uint32_t smem_u[];
  int* moff = reinterpret_cast<int*>(smem_u);          // [Mrows]
  int* poff = moff + g.Mrows;                          // [F8]
  uint32_t* xs = reinterpret_cast<uint32_t*>(poff + g.F8);   // [G][ximg]
  uint32_t* us = xs + g.G * g.ximg;                    // [G][uimg]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const int ms = warp % g.MS, kq = warp / g.MS;

  for (int m = tid; m < g.Mrows; m += nthr) {
    // rows past the last filter element repeat it: their lanes then read words another lane reads anyway (a broadcast, not
    // a bank conflict); their accumulators are never written out
    uint32_t c, tap, kh, kw;
    d_kk.divmod((uint32_t)min(m, g.CKK - 1), c, tap);
    if (g.khfast) d_kh.divmod(tap, kw, kh); else d_kw.divmod(tap, kh, kw);
    moff[m] = ((int)c * g.splane + (int)kh * g.sr + (int)kw * g.sc) * 4;   // byte offset
  }
  for (int f = tid; f < g.F8; f += nthr) {
    int off = 0;
    if (f < g.F) {
      uint32_t q, p;
      d_ho.divmod((uint32_t)f, q, p);
      off = (int)p * g.a_mod + (int)q * g.a_div;
    }
    poff[f] = off * 4;   // byte offset
  }
  // zero once: x halo, ug tails (f >= F) and padded channels (o >= O) are never rewritten by the staging below
  for (int e = tid; e < g.G * (g.ximg + g.uimg); e += nthr) xs[e] = 0u;

  float acc[MTW][NT][4];
#pragma unroll
  for (int i = 0; i < MTW; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

  __syncthreads();
  int M0[MTW], M1[MTW];
#pragma unroll
  for (int i = 0; i < MTW; ++i) {
    M0[i] = moff[(ms * MTW + i) * 16 + gq];
    M1[i] = moff[(ms * MTW + i) * 16 + gq + 8];
  }

  const uint32_t xs_sa = smem_u32(xs), po_sa = smem_u32(poff);
  const uint32_t us_sa = smem_u32(us) + (uint32_t)(gq * g.Up) * 4u;       // this lane's channel row of the ug fragment
  const uint32_t ximg_b = (uint32_t)g.ximg * 4u, uimg_b = (uint32_t)g.uimg * 4u, up8_b = (uint32_t)g.Up * 32u;
  const int xraw_img = g.C * g.raw_plane, uraw_img = g.O * g.F;
  int img_lo, img_cnt;
  cta_image_range(g.N, img_lo, img_cnt);
  // ---- PF state: one group's global data in flight
  float4 pfx[kPFX], pfg[kPFU];
  uint32_t pfa[kPFU];
  const int pad_planes_pf = g.NP - g.O;
  auto prefetch = [&](int n0, int ng) {
    const float* src = x + (int64_t)n0 * xraw_img;
    const int raw4 = (ng * xraw_img) >> 2;
#pragma unroll
    for (int i = 0; i < kPFX; ++i) {
      const int e4 = tid + i * nthr;
      if (e4 < raw4) pfx[i] = ld_stream4(src + 4 * (int64_t)e4);
    }
    const int64_t base = (int64_t)n0 * g.O * g.Fp;
    const int raw4u = (ng * g.O * g.Fp) >> 2;
#pragma unroll
    for (int i = 0; i < kPFU; ++i) {
      const int e4 = tid + i * nthr;
      if (e4 < raw4u) {
        // unit = four pooled positions p' = 4*pg .. 4*pg+3 of column q' (pooled (p', q') sits at q'*Hp + p').  Consecutive
        // lanes take consecutive q': their scattered 64-bit stores then sit side by side (with p' fastest they were
        // 8*Wo words apart = the same banks: 3-way conflicts in the ncu source view)
        uint32_t pl, u, pg, q;
        d_fp4.divmod((uint32_t)e4, pl, u);             // pl = img*O + o
        d_wp.divmod(u, pg, q);
        const uint32_t p = 4u * pg;
        const int64_t off = base + (int64_t)pl * g.Fp + q * (uint32_t)g.Hp + p;
        pfg[i] = ld_stream4(ug + off);
        pfa[i] = __ldg(reinterpret_cast<const uint32_t*>(pidx + off));   // per byte: argmax cell | relu'(z at the argmax) << 2
      }
    }
  };
  auto commit = [&](int ng) {
    const int raw4 = (ng * xraw_img) >> 2;
#pragma unroll
    for (int i = 0; i < kPFX; ++i) {
      const int e4 = tid + i * nthr;
      if (e4 < raw4) {
        const float4 v = pfx[i];
        if (g.xcontig) {
          *reinterpret_cast<uint4*>(xs + 4 * e4) = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        } else {   // four elements of one row (raw_cols % 4 == 0) into the searched layout
          uint32_t pl, f, r, c;
          d_plane.divmod((uint32_t)e4 * 4u, pl, f);
          d_cols.divmod(f, r, c);
          uint32_t* d = xs + pl * g.splane + (r + g.halo) * g.sr + (c + g.halo) * g.sc;
          d[0] = to_tf32(v.x); d[g.sc] = to_tf32(v.y); d[2 * g.sc] = to_tf32(v.z); d[3 * g.sc] = to_tf32(v.w);
        }
      }
    }
    const int raw4u = (ng * g.O * g.Fp) >> 2;
#pragma unroll
    for (int i = 0; i < kPFU; ++i) {
      const int e4 = tid + i * nthr;
      if (e4 < raw4u) {
        uint32_t pl, u, pg, q;
        d_fp4.divmod((uint32_t)e4, pl, u);
        d_wp.divmod(u, pg, q);
        const uint32_t p = 4u * pg;
        const uint32_t img = d_o.div(pl);
        uint32_t* row = us + (pl + img * pad_planes_pf) * g.Up + 2 * q + (2 * p) * (uint32_t)g.Wo;
        const float gs[4] = {pfg[i].x, pfg[i].y, pfg[i].z, pfg[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t ab = (pfa[i] >> (8 * j)) & 0xFFu, a = ab & 3u;
          const uint32_t tv = to_tf32((ab & 4u) ? gs[j] : 0.f);
          // the window's 2x2 cells: row kh = a >> 1, column kw = a & 1 holds the value
          *reinterpret_cast<uint2*>(row + (2 * j) * g.Wo) = make_uint2(a == 0u ? tv : 0u, a == 1u ? tv : 0u);
          *reinterpret_cast<uint2*>(row + (2 * j + 1) * g.Wo) = make_uint2(a == 2u ? tv : 0u, a == 3u ? tv : 0u);
        }
      }
    }
  };
  if constexpr (PF) {
    if (img_cnt > 0) prefetch(img_lo, min(g.G, img_cnt));
  }
  for (int n0 = img_lo; n0 < img_lo + img_cnt; n0 += g.G) {
    const int ng = min(g.G, img_lo + img_cnt - n0);
    __syncthreads();
    if constexpr (PF) {
      commit(ng);
    } else {
    if (g.xcontig) {  // ---- x: staged image == raw image
      const float* src = x + (int64_t)n0 * xraw_img;
      const int raw = ng * xraw_img;
      const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
      const int raw4 = vec ? (raw >> 2) : 0;
#pragma unroll 4
      for (int e4 = tid; e4 < raw4; e4 += nthr) {
        const float4 v = ld_stream4(src + 4 * (int64_t)e4);
        *reinterpret_cast<uint4*>(xs + 4 * e4) = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      }
      for (int e = raw4 * 4 + tid; e < raw; e += nthr) xs[e] = to_tf32(__ldg(src + e));
    } else {  // ---- x: natural layout + halo
      const float* src = x + (int64_t)n0 * xraw_img;
      const int raw = ng * xraw_img;
      const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
      const int raw4 = vec ? (raw >> 2) : 0;
#pragma unroll 2
      for (int e4 = tid; e4 < raw4; e4 += nthr) {
        const float4 v = ld_stream4(src + 4 * (int64_t)e4);
        uint32_t pl, f, r, c;
        d_plane.divmod((uint32_t)e4 * 4u, pl, f);
        d_cols.divmod(f, r, c);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xs[pl * g.splane + (r + g.halo) * g.sr + (c + g.halo) * g.sc] = to_tf32(vv[j]);
          if (++c == (uint32_t)g.raw_cols) {
            c = 0;
            if (++r == (uint32_t)g.raw_rows) { r = 0; ++pl; }
          }
        }
      }
      for (int e = raw4 * 4 + tid; e < raw; e += nthr) {
        uint32_t pl, f, r, c;
        d_plane.divmod((uint32_t)e, pl, f);
        d_cols.divmod(f, r, c);
        xs[pl * g.splane + (r + g.halo) * g.sr + (c + g.halo) * g.sc] = to_tf32(__ldg(src + e));
      }
    }
    if (g.pooled) {
      // ---- ug arrives POOLED (2x2 / stride 2, already masked by the ReLU gradient) with the argmax index of every window:
      // the full-resolution gradient (three quarters zeros) is rebuilt right here in shared memory and never exists in HBM
      const int rows4 = ng * g.uimg / 4;
      for (int e4 = tid; e4 < rows4; e4 += nthr) reinterpret_cast<uint4*>(us)[e4] = make_uint4(0u, 0u, 0u, 0u);
      __syncthreads();
      const int64_t base = (int64_t)n0 * g.O * g.Fp;
      const int raw = ng * g.O * g.Fp;
      const int pad_planes = g.NP - g.O;
      // value of pooled element e: ug[e] * relu'(z at the window's argmax) = bit 2 of the argmax byte (ngb_maxpool_relu_fwd)
      if ((g.Hp & 3) == 0 && ((uintptr_t)(ug + base) & 15u) == 0 && ((uintptr_t)(pidx + base) & 3u) == 0) {
        // four pooled positions along p' per thread: one index decomposition, 128-bit loads of ug / y, 32-bit load of idx
#pragma unroll 2
        for (int e4 = tid; e4 < (raw >> 2); e4 += nthr) {
          uint32_t pl, fe, q, p;
          d_fp.divmod((uint32_t)e4 * 4u, pl, fe);        // pl = img*O + o; pooled position (p', q') sits at q'*Hp + p'
          d_hp.divmod(fe, q, p);
          const uint32_t img = d_o.div(pl);
          const float4 gv = ld_stream4(ug + base + 4 * (int64_t)e4);
          const uint32_t a4 = *reinterpret_cast<const uint32_t*>(pidx + base + 4 * (int64_t)e4);
          uint32_t* row = us + (pl + img * pad_planes) * g.Up + 2 * q;
          const float gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t ab = (a4 >> (8 * j)) & 0xFFu, a = ab & 3u;
            row[(2 * (p + j) + (a >> 1)) * (uint32_t)g.Wo + (a & 1)] = to_tf32((ab & 4u) ? gs[j] : 0.f);
          }
        }
      } else {
#pragma unroll 4
        for (int e = tid; e < raw; e += nthr) {
          uint32_t pl, fe, q, p;
          d_fp.divmod((uint32_t)e, pl, fe);
          d_hp.divmod(fe, q, p);
          const uint32_t img = d_o.div(pl);
          const uint32_t ab = pidx[base + e], a = ab & 3u;   // argmax inside the window, row-major (kh, kw)
          const uint32_t f = (2 * p + (a >> 1)) * (uint32_t)g.Wo + 2 * q + (a & 1);
          us[(pl + img * pad_planes) * g.Up + f] = to_tf32((ab & 4u) ? __ldg(ug + base + e) : 0.f);
        }
      }
    } else
    {  // ---- ug: [o][f] as stored (neograd order), plane pitch Up
      const float* src = ug + (int64_t)n0 * uraw_img;
      const int raw = ng * uraw_img;
      const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && (g.F % 4 == 0);
      const int pad_planes = g.NP - g.O;             // zero planes between images in the staged copy
      if (vec) {
#pragma unroll 4
        for (int e4 = tid; e4 < (raw >> 2); e4 += nthr) {
          const float4 v = ld_stream4(src + 4 * (int64_t)e4);
          uint32_t pl, f;
          d_f.divmod((uint32_t)e4 * 4u, pl, f);      // pl = img*O + o
          const uint32_t img = d_o.div(pl);
          uint4 tv = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
          *reinterpret_cast<uint4*>(us + (pl + img * pad_planes) * g.Up + f) = tv;
        }
      } else {
        for (int e = tid; e < raw; e += nthr) {
          uint32_t pl, f;
          d_f.divmod((uint32_t)e, pl, f);
          const uint32_t img = d_o.div(pl);
          us[(pl + img * pad_planes) * g.Up + f] = to_tf32(__ldg(src + e));
        }
      }
    }
    }  // !PF
    __syncthreads();
    if constexpr (PF) {
      const int nn = n0 + g.G;
      if (nn < img_lo + img_cnt) prefetch(nn, min(g.G, img_lo + img_cnt - nn));
    }

    // warp kq takes k-steps kq, kq+KS, ... of EVERY image (no index division in the loop; per-image bases are hoisted:
    // the kernel is issue-bound, every integer instruction in this loop costs as much as a load)
    for (int img = 0; img < ng; ++img) {
      uint32_t xa0[MTW], xa1[MTW];
#pragma unroll
      for (int i = 0; i < MTW; ++i) {
        xa0[i] = xs_sa + (uint32_t)img * ximg_b + (uint32_t)M0[i];
        xa1[i] = xs_sa + (uint32_t)img * ximg_b + (uint32_t)M1[i];
      }
      uint32_t pq = po_sa + (uint32_t)kq * 32u + (uint32_t)t * 8u;            // &poff[8*ks + 2t], ks = kq
      uint32_t ub = us_sa + (uint32_t)img * uimg_b + (uint32_t)kq * 32u + (uint32_t)t * 8u;
      const uint32_t kstep_b = (uint32_t)g.KS * 32u;
#pragma unroll 2
      for (int ks = kq; ks < g.KSI; ks += g.KS) {
        const uint2 pp = lds64(pq);
        uint2 b[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) b[j] = lds64(ub + (uint32_t)j * up8_b);
        pq += kstep_b;
        ub += kstep_b;
        uint32_t a[MTW][4];
#pragma unroll
        for (int i = 0; i < MTW; ++i) {
          a[i][0] = lds32(xa0[i] + pp.x); a[i][1] = lds32(xa1[i] + pp.x);
          a[i][2] = lds32(xa0[i] + pp.y); a[i][3] = lds32(xa1[i] + pp.y);
        }
#pragma unroll
        for (int i = 0; i < MTW; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) mma_tf32(acc[i][j], a[i][0], a[i][1], a[i][2], a[i][3], b[j].x, b[j].y);
      }
    }
  }

  // ---- reduce the KS position splits in a fixed order, then emit this CTA's partial in dw's layout
  __syncthreads();
  float* red = reinterpret_cast<float*>(xs);          // [Mrows][NT*8]
  constexpr int NR = NT * 8;
  for (int r = 0; r < g.KS; ++r) {
    if (kq == r) {
#pragma unroll
      for (int i = 0; i < MTW; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int m = (ms * MTW + i) * 16 + gq + ((e & 2) ? 8 : 0);
            const int n = j * 8 + 2 * t + (e & 1);
            float* p = red + m * NR + n;
            *p = (r == 0) ? acc[i][j][e] : *p + acc[i][j][e];
          }
    }
    __syncthreads();
  }
  const int nw = g.O * g.CKK;
  for (int e = tid; e < nw; e += nthr) {
    uint32_t o, m;
    d_ckk.divmod((uint32_t)e, o, m);
    if (g.khfast) {   // A row of filter element (c, kh, kw)
      uint32_t c, tap, kh, kw;
      d_kk.divmod(m, c, tap);
      d_kw.divmod(tap, kh, kw);
      m = c * (uint32_t)(g.KH * g.KW) + kw * (uint32_t)g.KH + kh;
    }
    part[(int64_t)blockIdx.x * nw + e] = red[m * NR + o];
  }
}

// Layout of the staged x tile.  Lane (gq, t) of an A-fragment load reads word moff[row gq] + poff[position t]: 8 filter
// elements x 4 output positions of one k-step.  Words that coincide (filter tap + position shift reach the same pixel) are
// broadcasts; DIFFERENT words in one bank serialise (the natural pitch gave 2 wavefronts on every gather load: ncu source
// view, profiles/).  The row / column strides, the plane pitch and the order of the taps in the A rows are free, so the host
// searches them for the fewest wavefronts per load, then the smallest tile; the result is cached per geometry.
struct XLayout { int sr, sc, splane, khfast; float wavefronts; };

static float layout_cost(int C, int KH, int KW, int Ho, int F, int stride, int Mrows, int sr, int sc, int splane, int khfast) {
  const int CKK = C * KH * KW, KSI = (F + 7) / 8;
  int nks = (Ho % 8 == 0) ? 2 : 2 * ((Ho + 7) / 8) + 1;   // k-steps never straddle a column when Ho % 8 == 0: the pattern repeats
  if (nks > KSI) nks = KSI;
  long tot = 0, n = 0;
  int mo[8], po[4];
  for (int ks = 0; ks < nks; ++ks)
    for (int par = 0; par < 2; ++par) {
      for (int t = 0; t < 4; ++t) {
        int f = 8 * ks + 2 * t + par;
        if (f >= F) f = F - 1;
        po[t] = ((f % Ho) * sr + (f / Ho) * sc) * stride;
      }
      for (int rg = 0; rg < Mrows / 8; ++rg) {
        for (int gq = 0; gq < 8; ++gq) {
          int m = rg * 8 + gq;
          if (m > CKK - 1) m = CKK - 1;
          const int c = m / (KH * KW), tap = m % (KH * KW);
          const int kh = khfast ? tap % KH : tap / KW, kw = khfast ? tap / KH : tap % KW;
          mo[gq] = c * splane + kh * sr + kw * sc;
        }
        int words[32][4], cnt[32] = {0};
        int worst = 1;
        for (int gq = 0; gq < 8; ++gq)
          for (int t = 0; t < 4; ++t) {
            const int w = mo[gq] + po[t], b = w & 31;
            bool seen = false;
            for (int k = 0; k < cnt[b] && k < 4; ++k) seen |= (words[b][k] == w);
            if (!seen) {
              if (cnt[b] < 4) words[b][cnt[b]] = w;
              ++cnt[b];
              if (cnt[b] > worst) worst = cnt[b];
            }
          }
        tot += worst;
        ++n;
      }
    }
  return (float)tot / (float)n;
}

static XLayout search_x_layout(int C, int H, int W, int KH, int KW, int pad, int stride, int Ho, int Wo, int Mrows) {
  static std::mutex mu;
  static std::map<std::array<int, 10>, XLayout> cache;
  const std::array<int, 10> key = {C, H, W, KH, KW, pad, stride, Ho, Wo, Mrows};
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  const int R = H + 2 * pad, Cc = W + 2 * pad;
  XLayout best{Cc, 1, R * Cc, 0, 1e9f};
  long best_size = 0;
  for (int tr = 0; tr < 2; ++tr)            // tr: columns strided, rows contiguous
    for (int S = (tr ? R : Cc); S < (tr ? R : Cc) + 32; ++S) {
      const int foot = S * (tr ? Cc : R);
      for (int P = foot; P < foot + (C > 1 ? 32 : 4); ++P) {
        if ((C * P) & 3) continue;             // images stay 16-byte multiples: plane index img*C + c addresses the group
        for (int kf = 0; kf < 2; ++kf) {
          const int sr = tr ? 1 : S, sc = tr ? S : 1;
          const float c = layout_cost(C, KH, KW, Ho, Ho * Wo, stride, Mrows, sr, sc, P, kf);
          const long size = (long)C * P;
          // 2 % of a wavefront is not worth a bigger tile; otherwise fewer wavefronts first
          if (c < best.wavefronts - 0.02f || (c < best.wavefronts + 0.02f && size < best_size)) {
            best = XLayout{sr, sc, P, kf, c};
            best_size = size;
          }
        }
        if (C == 1) break;
      }
    }
  cache[key] = best;
  return best;
}

// Host-only view of the search for tests: out = {row stride, column stride, plane pitch, taps kh-fastest, wavefronts per
// A-gather load x 1000} for the weight-gradient tile of the given layer.
int wgrad_layout_for_tests(int C, int H, int W, int KH, int KW, int pad, int stride, int* out) {
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  const int mt_total = ceil_div(C * KH * KW, 16), MS = ceil_div(mt_total, 5), MTW = ceil_div(mt_total, MS);
  const int Mrows = MS * MTW * 16;
  const XLayout L = search_x_layout(C, H, W, KH, KW, pad, stride, Ho, Wo, Mrows);
  out[0] = L.sr; out[1] = L.sc; out[2] = L.splane; out[3] = L.khfast; out[4] = (int)(L.wavefronts * 1000.f + 0.5f);
  const float natural = layout_cost(C, KH, KW, Ho, Ho * Wo, stride, Mrows, round_up_mod(W + 2 * pad, 8, 4), 1,
                                    (H + 2 * pad) * round_up_mod(W + 2 * pad, 8, 4), 0);
  out[5] = (int)(natural * 1000.f + 0.5f);
  return NGB_OK;
}

template <int MTW, int NT, bool PF>
static int launch_wgrad(const float* ug, const float* x, const uint8_t* pidx, const float* ypool, const float* zmask, float* part,
                        WgradGeo g, int nwarps, int gmin, int gmax, size_t fixed, size_t per_img, int* grid_out, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    NGB_CUDA(cudaFuncSetAttribute(conv_wgrad_mma_kernel<MTW, NT, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    // maximum shared-memory carveout: a kernel of another lane can only become co-resident on an SM whose L1 / shared split
    // already has room for it -- with the default (smallest fitting) split the second kernel waits for the first to drain
    NGB_CUDA(cudaFuncSetAttribute(conv_wgrad_mma_kernel<MTW, NT, PF>, cudaFuncAttributePreferredSharedMemoryCarveout, getenv("NGB_CARVEOUT_DEFAULT") ? cudaSharedmemCarveoutDefault : cudaSharedmemCarveoutMaxShared));
    attr = true;
  }
  // Beside the data-gradient chain (ngb_set_corun) a persistent grid at full occupancy would hold every register of
  // every SM until it ends and stall the chain's next kernel: balance_images then leaves half of each SM free.
  int grid_i = 1;
  int rc = balance_images(conv_wgrad_mma_kernel<MTW, NT, PF>, nwarps * 32, fixed, per_img, g.N, gmin, gmax, corun_hint() != 0, &g.G,
                          &grid_i);
  if (rc) return rc;
  const size_t smem = fixed + (size_t)g.G * per_img;
  const int64_t nwts = (int64_t)g.O * g.CKK;
  int64_t grid = grid_i;
  if (grid * nwts * 4 > scratch_bytes()) grid = scratch_bytes() / (nwts * 4);
  if (grid < 1) return fail(NGB_ERR_SCRATCH, "conv wgrad: scratch arena too small for %lld filter elements", (long long)nwts);
  *grid_out = (int)grid;
  conv_wgrad_mma_kernel<MTW, NT, PF><<<(unsigned)grid, nwarps * 32, smem, st>>>(
      ug, x, pidx, ypool, zmask, part, g, FastDiv((uint32_t)(g.Fp > 0 ? g.Fp : 1)), FastDiv((uint32_t)(g.Hp > 0 ? g.Hp : 1)),
      FastDiv((uint32_t)g.raw_plane), FastDiv((uint32_t)g.raw_cols), FastDiv((uint32_t)g.Ho),
      FastDiv((uint32_t)g.KSI), FastDiv((uint32_t)(g.KH * g.KW)), FastDiv((uint32_t)g.KW), FastDiv((uint32_t)g.F),
      FastDiv((uint32_t)g.CKK), FastDiv((uint32_t)g.O), FastDiv((uint32_t)(g.Fp >= 4 ? g.Fp / 4 : 1)),
      FastDiv((uint32_t)(g.Wo >= 2 ? g.Wo / 2 : 1)), FastDiv((uint32_t)g.KH));
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

// pidx != null: `ug` is the 2x2 / stride 2 POOLED gradient [N][O][Ho/2 * Wo/2] (pooled position (p', q') at q'*(Ho/2) + p'),
// pidx the argmax index of every pooling window, ypool the pool's output and zmask the conv output (= ReLU input); the kernel
// applies the ReLU mask and scatters the gradient into the full-resolution tile itself.
int conv_mma_wgrad(const float* ug, const float* x, float* dw, int N, int C, int H, int W, int O, int KH, int KW, int pad,
                   int stride, int Ho, int Wo, int accumulate, cudaStream_t st, bool* done, const uint8_t* pidx,
                   const float* ypool, const float* zmask) {
  *done = false;
  if (g_default_engine.load() == NGB_GEMM_SIMT) return NGB_OK;
  if (O < 4 || O > 16) return NGB_OK;
  if (pidx && ((Ho & 1) || (Wo & 1))) return NGB_OK;
  WgradGeo g;
  g.pooled = pidx ? 1 : 0; g.Wo = Wo; g.Hp = Ho / 2; g.Fp = (Ho / 2) * (Wo / 2);
  g.N = N; g.C = C; g.O = O; g.KH = KH; g.KW = KW; g.Ho = Ho;
  g.raw_rows = H; g.raw_cols = W; g.raw_plane = H * W; g.halo = pad;
  const int mt_total = ceil_div(C * KH * KW, 16);
  g.MS = ceil_div(mt_total, 5);
  if (g.MS > 8) return NGB_OK;
  const int MTW = ceil_div(mt_total, g.MS);
  g.Mrows = g.MS * MTW * 16;
  if (getenv("NGB_WGRAD_NATURAL_LAYOUT")) {   // experiments: the pre-search layout
    g.sr = round_up_mod(W + 2 * pad, 8, 4); g.sc = 1; g.splane = (H + 2 * pad) * g.sr; g.khfast = 0;
  } else {
    const XLayout L = search_x_layout(C, H, W, KH, KW, pad, stride, Ho, Wo, g.Mrows);
    g.sr = L.sr; g.sc = L.sc; g.splane = L.splane; g.khfast = L.khfast;
  }
  g.ximg = C * g.splane;               // multiple of 4 (search constraint)
  g.xcontig = (pad == 0 && g.sc == 1 && g.sr == W && g.splane == H * W) ? 1 : 0;
  g.F = Ho * Wo; g.F8 = (g.F + 7) & ~7; g.KSI = g.F8 / 8;
  g.Up = round_up_mod(g.F8, 32, 8);
  g.NP = (O + 7) & ~7;
  g.uimg = g.NP * g.Up;
  g.CKK = C * KH * KW;
  g.a_mod = stride * g.sr; g.a_div = stride * g.sc;
  const int NT = g.NP / 8;
  g.KS = 8 / g.MS; if (g.KS < 1) g.KS = 1;
  const int nwarps = g.MS * g.KS;
  const size_t fixed = ((size_t)g.Mrows + g.F8) * 4;
  const size_t per_img = ((size_t)g.ximg + g.uimg) * 4;
  const size_t red_bytes = (size_t)g.Mrows * NT * 8 * 4;
  if (fixed + per_img > 96 * 1024) return NGB_OK;
  int G = 8;
  while (G > 1 && fixed + G * per_img > 48 * 1024) --G;
  if (G > N) G = N;
  const int gmin = (int)ceil_div<size_t>(red_bytes, per_img);   // the reduction tile reuses the staging area
  if (G < gmin) return NGB_OK;
  // register-prefetched staging (kernel comment): whole groups must fit the per-thread prefetch registers
  bool pf = false;
  if (pidx && (W & 3) == 0 && (g.Hp & 3) == 0 && !getenv("NGB_WGRAD_NO_PF") &&
      (((uintptr_t)x | (uintptr_t)ug) & 15u) == 0 && ((uintptr_t)pidx & 3u) == 0) {
    const int nthr = nwarps * 32;
    const int gx = kPFX * nthr * 4 / (C * H * W), gu = kPFU * nthr * 4 / (O * g.Fp);
    const int gp = gx < gu ? gx : gu;
    if (gp >= gmin && gp >= 1) {
      pf = true;
      if (G > gp) G = gp;
    }
  }
  if (const char* e = getenv("NGB_WGRAD_G")) {   // experiments: cap the images per group
    const int ge = atoi(e);
    if (ge >= gmin && ge >= 1 && ge < G) G = ge;
  }
  int rc = ensure_init();
  if (rc) return rc;
  const int64_t nw = (int64_t)O * g.CKK;
  if (nw * 4 * kNumSMs * 4 > scratch_bytes()) return NGB_OK;
  int grid = 0;
  float* part = scratch_ptr();
#define NGB_WGRAD(M_, N_)                                                                                                  \
  rc = pf ? launch_wgrad<M_, N_, true>(ug, x, pidx, ypool, zmask, part, g, nwarps, gmin, G, fixed, per_img, &grid, st)     \
          : launch_wgrad<M_, N_, false>(ug, x, pidx, ypool, zmask, part, g, nwarps, gmin, G, fixed, per_img, &grid, st)
  if (NT == 1) {
    switch (MTW) { case 1: NGB_WGRAD(1, 1); break; case 2: NGB_WGRAD(2, 1); break; case 3: NGB_WGRAD(3, 1); break;
                   case 4: NGB_WGRAD(4, 1); break; default: NGB_WGRAD(5, 1); break; }
  } else {
    switch (MTW) { case 1: NGB_WGRAD(1, 2); break; case 2: NGB_WGRAD(2, 2); break; case 3: NGB_WGRAD(3, 2); break;
                   case 4: NGB_WGRAD(4, 2); break; default: NGB_WGRAD(5, 2); break; }
  }
#undef NGB_WGRAD
  if (rc) return rc;
  *done = true;
  return launch_sum_partials(part, grid, nw, dw, accumulate, st);
}

// ------------------------------------------------------------------------------------------------ sparse pooled wgrad + bgrad
// First-layer weight AND bias gradient of conv -> ReLU -> MaxPool(2x2, s2) straight from the pooled upstream gradient, on the
// CUDA cores in full fp32.  Of the four cells of a pooling window only the argmax cell carries gradient, so
//   dw[o][c][kh][kw] = sum over windows of gm * x[c][pc*stride + kh][qc*stride + kw],   db[o] = sum of gm,
// gm = ug_pool * relu'(z at the argmax) (bit 2 of the argmax byte, written by ngb_maxpool_relu_fwd), (pc, qc) = the argmax
// cell's conv position: C*KH*KW FMAs per (window, o) -- a quarter of the dense product, no scattered tile, no TF32 rounding,
// and two coalesced loads (gradient, argmax byte) per item.  With few input channels the filter fits the registers of ONE
// thread: every lane keeps all C*KH*KW (+1 bias) sums of one output channel.
// Lane = (o, window): the O lanes of a window read x patches that differ only by the argmax offset (<= 4 distinct words per
// load, mostly broadcasts) and a warp spans 32/O neighbouring windows.  (One window column per lane -- every lane its own
// patch -- cost 2.9 shared-memory wavefronts per load and ran as slow as the tensor-core kernel, which pads 25 x 6 to 32 x 8,
// gathers every A fragment and spends as many instructions staging the scattered gradient as on the MMAs.)
// x images arrive through cp.async into a double buffer; one barrier per group of images.
struct SparseGeo {
  int N, O, H, W, Ho, Wo, Hp, Wp, Fp, stride;
  int G, wpw;          // images per group, windows per warp iteration (32 / O)
  int ximg;            // C*H*W (global)
  int S, simg;         // staged row pitch and image size (C*H*S): S puts the two argmax rows of a window on disjoint banks
  int w4;              // 16-byte chunks per x row (W / 4), 0: rows are not 16-byte multiples, the image is copied flat (S == W)
  int square;          // Ho == Wo: conv position of the stored cell (r, c) is (c, r) without a division
  int d_il, d_q, d_p;  // one warp-iteration step (kSparseWarps * wpw windows) as (images, pooled columns, pooled rows)
};

constexpr int kSparseWarps = 12;

template <int C, int KH, int KW>
__global__ void __launch_bounds__(kSparseWarps * 32, 3) conv_wbgrad_pooled_sparse_kernel(const float* __restrict__ ug,
                                                                                    const uint8_t* __restrict__ pidx,
                                                                                    const float* __restrict__ x,
                                                                                    float* __restrict__ part, float* __restrict__ partb,
                                                                                    SparseGeo g, FastDiv d_fp, FastDiv d_hp,
                                                                                    FastDiv d_ho, FastDiv d_wpw, FastDiv d_w4) {
  constexpr int CKK = C * KH * KW;
  extern __shared__ __align__(16) float smem_f[];
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  uint32_t o, wl;
  d_wpw.divmod((uint32_t)lane, o, wl);
  const bool active = (int)o < g.O;              // 32 % O spare lanes idle
  const int gimg = g.G * g.simg;                 // floats per buffer
  float* bufs = smem_f;                          // [2][G][simg]

  float acc[CKK], accb = 0.f;
#pragma unroll
  for (int i = 0; i < CKK; ++i) acc[i] = 0.f;

  int img_lo, img_cnt;
  cta_image_range(g.N, img_lo, img_cnt);
  auto stage = [&](int n0, int ng, int b) {      // 16-byte cp.async: images are 16-byte multiples (host-checked)
    const float* src = x + (int64_t)n0 * g.ximg;
    const uint32_t dst = smem_u32(bufs + b * gimg);
    const int n16 = (ng * g.ximg) >> 2;
    for (int e = tid; e < n16; e += nthr) {
      uint32_t d = (uint32_t)e * 16u;
      if (g.w4) {                                // row r of the group (images are C*H rows back to back), chunk k
        uint32_t r, k;
        d_w4.divmod((uint32_t)e, r, k);
        d = (r * (uint32_t)g.S + 4u * k) * 4u;
      }
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + d), "l"(src + 4 * (int64_t)e) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (img_cnt > 0) stage(img_lo, min(g.G, img_cnt), 0);
  int cur = 0;
  const int wstep = kSparseWarps * g.wpw;
  const uint32_t ofp = o * (uint32_t)g.Fp, img_fp = (uint32_t)(g.O * g.Fp);
  for (int n0 = img_lo; n0 < img_lo + img_cnt; n0 += g.G) {
    const int ng = min(g.G, img_lo + img_cnt - n0);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                             // buffer `cur` is complete; everyone is done reading the other one
    const int nn = n0 + g.G;
    if (nn < img_lo + img_cnt) stage(nn, min(g.G, img_lo + img_cnt - nn), cur ^ 1);
    const float* xb = bufs + cur * gimg;
    const int nwin = active ? ng * g.Fp : 0;     // windows in the group (every one has O lanes)
    const float* ugb = ug + (int64_t)n0 * img_fp;
    const uint8_t* idb = pidx + (int64_t)n0 * img_fp;
    // operands of one item, loaded one iteration ahead
    // (image, pooled column q', pooled row p') of this lane's window: decoded once per group, then stepped
    int w = warp * g.wpw + (int)wl;
    uint32_t il, fe0, qq, pp;
    d_fp.divmod((uint32_t)w, il, fe0);
    d_hp.divmod(fe0, qq, pp);
    float gv = 0.f, gv_n = 0.f;
    uint32_t av = 0, av_n = 0;
    auto fetch = [&](bool ok, uint32_t i, uint32_t q, uint32_t p_, float& gout, uint32_t& aout) {
      if (ok) {
        const uint32_t off = i * img_fp + ofp + q * (uint32_t)g.Hp + p_;
        gout = __ldg(ugb + off);
        aout = __ldg(idb + off);
      }
    };
    fetch(w < nwin, il, qq, pp, gv, av);
    for (; w < nwin; w += wstep) {
      // the next window of this lane
      uint32_t pn = pp + (uint32_t)g.d_p, qn = qq + (uint32_t)g.d_q, in_ = il + (uint32_t)g.d_il;
      if (pn >= (uint32_t)g.Hp) { pn -= (uint32_t)g.Hp; ++qn; }
      if (qn >= (uint32_t)g.Wp) { qn -= (uint32_t)g.Wp; ++in_; }
      fetch(w + wstep < nwin, in_, qn, pn, gv_n, av_n);
      const float gm = (av & 4u) ? gv : 0.f;
      // argmax cell in the conv output plane as stored: (r, c) = (2p' + kh_a, 2q' + kw_a), flat r*Wo + c -> the convolution
      // produced it at position (flat % Ho, flat / Ho)
      const uint32_t r = 2 * pp + ((av >> 1) & 1u), cc = 2 * qq + (av & 1u);
      uint32_t pc, qc;
      if (g.square) { pc = cc; qc = r; } else d_ho.divmod(r * (uint32_t)g.Wo + cc, qc, pc);
      const float* xw = xb + il * (uint32_t)g.simg + (pc * (uint32_t)g.stride) * (uint32_t)g.S + qc * (uint32_t)g.stride;
      pp = pn; qq = qn; il = in_; gv = gv_n; av = av_n;
      accb += gm;
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int kh = 0; kh < KH; ++kh) {
          const float* xr = xw + c * g.H * g.S + kh * g.S;
#pragma unroll
          for (int kw = 0; kw < KW; ++kw) acc[(c * KH + kh) * KW + kw] = fmaf(gm, xr[kw], acc[(c * KH + kh) * KW + kw]);
        }
    }
    cur ^= 1;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");

  // ---- every lane's sums to shared memory, then one thread per (o, filter element) adds warps and lanes in a fixed order
  __syncthreads();
  float* red = smem_f;                           // [threads][CKK + 1]
#pragma unroll
  for (int i = 0; i < CKK; ++i) red[tid * (CKK + 1) + i] = acc[i];
  red[tid * (CKK + 1) + CKK] = accb;
  __syncthreads();
  for (int e = tid; e < g.O * (CKK + 1); e += nthr) {
    const int oo = e / (CKK + 1), k = e % (CKK + 1);
    float v = 0.f;
    for (int wi = 0; wi < kSparseWarps; ++wi)
      for (int l = 0; l < g.wpw; ++l) v += red[(wi * 32 + oo * g.wpw + l) * (CKK + 1) + k];
    // one row of O*CKK weight sums followed by O bias sums per CTA
    if (k < CKK) part[(int64_t)blockIdx.x * (g.O * (CKK + 1)) + oo * CKK + k] = v;
    else part[(int64_t)blockIdx.x * (g.O * (CKK + 1)) + g.O * CKK + oo] = v;
  }
}

template <int C, int KH, int KW>
static int launch_sparse(const float* ug, const uint8_t* pidx, const float* x, float* dw, float* db, SparseGeo g, int acc_w,
                         int acc_b, cudaStream_t st) {
  constexpr int CKK = C * KH * KW;
  auto kern = conv_wbgrad_pooled_sparse_kernel<C, KH, KW>;
  static bool attr = false;
  if (!attr) {
    NGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    NGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, getenv("NGB_CARVEOUT_DEFAULT") ? cudaSharedmemCarveoutDefault : cudaSharedmemCarveoutMaxShared));
    attr = true;
  }
  const int threads = kSparseWarps * 32;
  const size_t per_img = (size_t)2 * g.simg * 4;              // double buffer
  const size_t red_bytes = (size_t)threads * (CKK + 1) * 4;   // the reduction reuses the buffers
  int gmax = 6;
  if (const char* e = getenv("NGB_SPARSE_G")) gmax = atoi(e) > 0 ? atoi(e) : gmax;
  const int gmin = (int)ceil_div<size_t>(red_bytes, per_img);
  if (gmax < gmin) gmax = gmin;
  if (gmax * per_img > 96 * 1024) return -1;
  int grid = 1;
  int rc = balance_images(kern, threads, 0, per_img, g.N, gmin, gmax, corun_hint() != 0, &g.G, &grid);
  if (rc) return rc;
  const int64_t nw = (int64_t)g.O * CKK;
  if (((int64_t)grid * (nw + g.O) + g.O) * 4 > scratch_bytes()) return -1;
  float* part = scratch_ptr();
  float* partb = nullptr;
  const int wstep = kSparseWarps * g.wpw;                     // windows per CTA iteration, as (images, columns, rows)
  g.d_il = wstep / g.Fp;
  g.d_q = (wstep % g.Fp) / g.Hp;
  g.d_p = (wstep % g.Fp) % g.Hp;
  kern<<<grid, threads, (size_t)g.G * per_img, st>>>(ug, pidx, x, part, partb, g, FastDiv((uint32_t)g.Fp), FastDiv((uint32_t)g.Hp),
                                                     FastDiv((uint32_t)g.Ho), FastDiv((uint32_t)g.wpw),
                                                     FastDiv((uint32_t)(g.w4 > 0 ? g.w4 : 1)));
  NGB_LAUNCH_CHECK();
  // the bias columns are summed into db, or dropped (out2 == nullptr keeps them inside the row: sum them into the scratch row)
  return launch_sum_partials2(part, grid, nw, dw, acc_w, g.O, db ? db : part + (int64_t)grid * (nw + g.O), db ? acc_b : 0, st);
}

bool conv_sparse_serves(int C, int H, int W, int O, int KH, int KW, int pad, int Ho, int Wo) {
  if (getenv("NGB_NO_SPARSE_WGRAD")) return false;
  if (pad != 0 || (Ho & 1) || (Wo & 1) || O < 1 || O > 12 || ((C * H * W) & 3)) return false;
  return (C == 1 && KH == 5 && KW == 5) || (C == 1 && KH == 3 && KW == 3) || (C == 3 && KH == 3 && KW == 3);
}

// *done = false: geometry not served (the caller uses the tensor-core kernel / the separate bias kernel).
int conv_sparse_wbgrad_pooled(const float* ug, const float* ypool, const uint8_t* pidx, const float* z, const float* x, float* dw,
                              float* db, int N, int C, int H, int W, int O, int KH, int KW, int pad, int stride, int Ho, int Wo,
                              int acc_w, int acc_b, cudaStream_t st, bool* done) {
  *done = false;
  if (N < 1 || !conv_sparse_serves(C, H, W, O, KH, KW, pad, Ho, Wo) || ((uintptr_t)x & 15u)) return NGB_OK;
  (void)ypool; (void)z;                // the ReLU mask travels in bit 2 of the argmax bytes
  SparseGeo g;
  g.N = N; g.O = O; g.H = H; g.W = W; g.Ho = Ho; g.Wo = Wo; g.Hp = Ho / 2; g.Fp = (Ho / 2) * (Wo / 2);
  g.Wp = Wo / 2;
  g.stride = stride; g.ximg = C * H * W; g.G = 1; g.wpw = 32 / O; g.square = (Ho == Wo) ? 1 : 0;
  // Staged row pitch.  The lanes of a warp read words 2*wl*stride + {0, stride} (argmax column) + {0, S*stride} (argmax row), wl <
  // wpw: the two rows of a window land on disjoint banks when S*stride mod 32 lies in [2*wpw*stride, 32 - 2*wpw*stride].
  g.S = W; g.w4 = 0;
  if ((W & 3) == 0) {
    g.w4 = W / 4;
    const int span = 2 * g.wpw * stride;
    for (int Sx = W; Sx < W + 32; Sx += 4) {
      const int m = (Sx * stride) & 31;
      if (m >= span && m <= 32 - span) { g.S = Sx; break; }
    }
  }
  g.simg = C * H * g.S;
  int rc = ensure_init();
  if (rc) return rc;
  rc = -1;
  if (C == 1 && KH == 5 && KW == 5) rc = launch_sparse<1, 5, 5>(ug, pidx, x, dw, db, g, acc_w, acc_b, st);
  else if (C == 1 && KH == 3 && KW == 3) rc = launch_sparse<1, 3, 3>(ug, pidx, x, dw, db, g, acc_w, acc_b, st);
  else if (C == 3 && KH == 3 && KW == 3) rc = launch_sparse<3, 3, 3>(ug, pidx, x, dw, db, g, acc_w, acc_b, st);
  else return NGB_OK;
  if (rc == -1) return NGB_OK;
  if (rc) return rc;
  *done = true;
  return NGB_OK;
}

// ------------------------------------------------------------------------------------------------ conv + bias + ReLU + MaxPool(2x2, s2) forward
// First-layer block in ONE kernel: y_pool / idx = maxpool2x2(relu(conv(x, w) + b)) (conv.py:266-297, activations.py:20,
// conv.py:384-402).  The convolution's own output (4x the pooled map; for LeNet-B conv1 57 MB written and read back by the
// pool) never exists: the backward kernels of such a block read the pooled gradient and the argmax bytes only.  CUDA cores,
// full fp32: a thread owns one pooling window of every output channel -- conv positions p in 2q'..2q'+1 (image rows), q in
// 2p'..2p'+1 (image columns) -- i.e. 4*O accumulators fed from a 6 x 6 input patch in registers (64-bit shared-memory loads,
// conflict-free for the searched row pitch) and the filter bank in shared memory (128-bit broadcast loads): 25 taps x 4
// positions x O FMAs per 18 + 50 loads.  Bias, ReLU, the window maximum (first maximum wins, as maxpool2x2_fwd_kernel) and
// the ReLU-gradient flag (bit 2 of the argmax byte) finish in registers.  Square outputs with Ho % 4 == 0, pad 0, stride 1.
struct FusedFwdGeo {
  int N, H, W, Ho, Hp, Fp;
  int G, S, simg, ximg, tpi;     // images per group, staged row pitch / image size, raw image size, thread tiles per image
};

template <int C, int KH, int KW, int O>
__global__ void __launch_bounds__(256, 3) conv_relu_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, float* __restrict__ y,
                                                                   uint8_t* __restrict__ idx, FusedFwdGeo g, FastDiv d_tpi,
                                                                   FastDiv d_hp, FastDiv d_w2) {
  constexpr int CKK = C * KH * KW, OP = (O + 3) & ~3, PR = 2 + KH - 1, PC = 2 + KW - 1;
  static_assert((PC & 1) == 0, "odd filter widths only: the patch rows are read as 64-bit pairs");
  extern __shared__ __align__(16) float smem_f[];
  float* ws = smem_f;                            // [CKK][OP]
  float* bs = ws + CKK * OP;                     // [OP]
  float* bufs = bs + OP;                         // [2][G][simg]   (CKK*OP + OP is a multiple of 4)
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int e = tid; e < CKK * OP; e += nthr) {
    const int k = e / OP, o = e % OP;
    ws[e] = o < O ? __ldg(w + (int64_t)o * CKK + k) : 0.f;
  }
  for (int e = tid; e < OP; e += nthr) bs[e] = e < O ? __ldg(bias + e) : 0.f;

  const int gimg = g.G * g.simg;
  int img_lo, img_cnt;
  cta_image_range(g.N, img_lo, img_cnt);
  auto stage = [&](int n0, int ng, int b) {      // 8-byte cp.async: the staged pitch is even, not a multiple of 4
    const float* src = x + (int64_t)n0 * g.ximg;
    const uint32_t dst = smem_u32(bufs + b * gimg);
    const int n8 = (ng * g.ximg) >> 1;
    for (int e = tid; e < n8; e += nthr) {
      uint32_t r, k;
      d_w2.divmod((uint32_t)e, r, k);            // row r of the group (images are C*H rows back to back), pair k
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (r * (uint32_t)g.S + 2u * k) * 4u), "l"(src + 2 * (int64_t)e)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (img_cnt > 0) stage(img_lo, min(g.G, img_cnt), 0);
  int cur = 0;
  for (int n0 = img_lo; n0 < img_lo + img_cnt; n0 += g.G) {
    const int ng = min(g.G, img_lo + img_cnt - n0);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                             // buffer `cur` (and, the first time, the filter bank) is complete
    const int nn = n0 + g.G;
    if (nn < img_lo + img_cnt) stage(nn, min(g.G, img_lo + img_cnt - nn), cur ^ 1);
    const float* xb = bufs + cur * gimg;
    for (int ti = tid; ti < ng * g.tpi; ti += nthr) {
      uint32_t il, rem, qp, pq;
      d_tpi.divmod((uint32_t)ti, il, rem);
      d_hp.divmod(rem, qp, pq);                  // window (p' = pq, q' = qp): p' fastest over lanes
      float acc[2][2][O];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2)
#pragma unroll
          for (int o = 0; o < O; ++o) acc[a][b2][o] = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float* xp = xb + il * (uint32_t)g.simg + (c * g.H + 2 * qp) * (uint32_t)g.S + 2 * pq;
        // one patch row at a time (6 values live instead of 36): row r feeds output row a through filter row kh = r - a
#pragma unroll
        for (int r = 0; r < PR; ++r) {
          float row[PC];
#pragma unroll
          for (int k = 0; k < PC; k += 2) {
            const float2 t2 = *reinterpret_cast<const float2*>(xp + r * g.S + k);
            row[k] = t2.x; row[k + 1] = t2.y;
          }
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int kh = r - a;
            if (kh < 0 || kh >= KH) continue;
#pragma unroll
            for (int kw = 0; kw < KW; ++kw) {
              // keep the scheduler from hoisting every filter load of the (fully unrolled) tile to the top: it then needs
              // ~230 registers and spills; with the loads pinned per tap pair the tile fits 2-3 CTAs per SM
              if ((kw & 1) == 0) asm volatile("" ::: "memory");
              float wv[OP];
#pragma unroll
              for (int o4 = 0; o4 < OP; o4 += 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(ws + ((c * KH + kh) * KW + kw) * OP + o4);
                wv[o4] = t4.x; wv[o4 + 1] = t4.y; wv[o4 + 2] = t4.z; wv[o4 + 3] = t4.w;
              }
#pragma unroll
              for (int b2 = 0; b2 < 2; ++b2)
#pragma unroll
                for (int o = 0; o < O; ++o) acc[a][b2][o] = fmaf(row[b2 + kw], wv[o], acc[a][b2][o]);
            }
          }
        }
      }
      // stored cell (kh2, kw2) of the window is conv position (p, q) = (2q' + kw2, 2p' + kh2) = acc[kw2][kh2]
      const int64_t f = ((int64_t)(n0 + (int)il) * O) * g.Fp + (int64_t)qp * g.Hp + pq;
#pragma unroll
      for (int o = 0; o < O; ++o) {
        const float b0 = bs[o];
        const float z0 = acc[0][0][o] + b0, z1 = acc[1][0][o] + b0, z2 = acc[0][1][o] + b0, z3 = acc[1][1][o] + b0;
        float best = fmaxf(0.f, z0);
        int arg = 0;
        const float r1 = fmaxf(0.f, z1), r2 = fmaxf(0.f, z2), r3 = fmaxf(0.f, z3);
        if (r1 > best) { best = r1; arg = 1; }
        if (r2 > best) { best = r2; arg = 2; }
        if (r3 > best) { best = r3; arg = 3; }
        if (best > 0.f || z0 >= 0.f) arg |= 4;     // relu'(z at the argmax), as maxpool2x2_fwd_kernel<true>
        y[f + (int64_t)o * g.Fp] = best;
        idx[f + (int64_t)o * g.Fp] = (uint8_t)arg;
      }
    }
    cur ^= 1;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
}

template <int C, int KH, int KW, int O>
static int launch_fused_fwd(const float* x, const float* w, const float* bias, float* y, uint8_t* idx, FusedFwdGeo g, cudaStream_t st) {
  constexpr int CKK = C * KH * KW, OP = (O + 3) & ~3;
  auto kern = conv_relu_pool_fwd_kernel<C, KH, KW, O>;
  static bool attr = false;
  if (!attr) {
    NGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    NGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, getenv("NGB_CARVEOUT_DEFAULT") ? cudaSharedmemCarveoutDefault : cudaSharedmemCarveoutMaxShared));
    attr = true;
  }
  const size_t fixed = (size_t)(CKK * OP + OP) * 4;
  const size_t per_img = (size_t)2 * g.simg * 4;
  // images per group: fill the 256 threads' tile slots as evenly as possible
  int best_g = 1;
  double best_eff = 0;
  for (int G = 1; G <= 12 && fixed + G * per_img <= 72 * 1024; ++G) {
    const int tiles = G * g.tpi;
    const double eff = (double)tiles / (double)(ceil_div(tiles, 256) * 256);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_g = G; }
  }
  if (fixed + best_g * per_img > 96 * 1024) return -1;
  int grid = 1;
  int rc = balance_images(kern, 256, fixed, per_img, g.N, best_g, best_g, corun_hint() != 0, &g.G, &grid);
  if (rc) return rc;
  kern<<<grid, 256, fixed + (size_t)g.G * per_img, st>>>(x, w, bias, y, idx, g, FastDiv((uint32_t)g.tpi), FastDiv((uint32_t)g.Hp),
                                                        FastDiv((uint32_t)(g.W / 2)));
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

// small-channel layers whose output plane is 8 x (even): the tensor-core gather kernel with the pooling epilogue
static bool gather_pool_serves(int C, int H, int W, int O, int KH, int KW, int pad, int stride, int Ho, int Wo) {
  if (g_default_engine.load() == NGB_GEMM_SIMT || getenv("NGB_NO_FUSED_CONV2_POOL")) return false;
  if (Ho != 8 || Wo != 8 || O < 4 || O > 32 || C < 2) return false;
  const int NP = (O + 7) & ~7;
  if (NP != 8 && NP != 16 && NP != 32) return false;
  const int spitch = W + 2 * pad + 8, K = C * KH * KW;            // (upper bounds of the staged pitch / table sizes)
  const size_t fixed = ((size_t)((K + 7) & ~7) * (NP + 4) + ((K + 7) & ~7) + (size_t)Ho * Wo) * 4;
  return fixed + (size_t)C * (H + 2 * pad) * spitch * 4 <= 96 * 1024 && K <= 4096;
}

bool conv_fused_fwd_serves(int C, int H, int W, int O, int KH, int KW, int pad, int stride, int Ho, int Wo) {
  if (getenv("NGB_NO_FUSED_CONV_POOL")) return false;
  if (gather_pool_serves(C, H, W, O, KH, KW, pad, stride, Ho, Wo)) return true;
  if (pad != 0 || stride != 1 || Ho != Wo || (Ho & 1) || (W & 1)) return false;
  return (C == 1 && KH == 5 && KW == 5 && O == 6) || (C == 1 && KH == 3 && KW == 3 && (O == 4 || O == 8));
}

// *done = false: geometry not served (the caller runs the convolution and the fused ReLU + pool kernel separately).
int conv_fused_relu_pool_fwd(const float* x, const float* w, const float* bias, float* y, uint8_t* idx, int N, int C, int H, int W,
                             int O, int KH, int KW, int pad, int stride, int Ho, int Wo, cudaStream_t st, bool* done) {
  *done = false;
  if (N >= 1 && !getenv("NGB_NO_FUSED_CONV_POOL") && gather_pool_serves(C, H, W, O, KH, KW, pad, stride, Ho, Wo))
    return conv_mma_gather(x, w, bias, y, N, C, H, W, O, KH, KW, pad, stride, Ho, Wo, 0, 0, st, done, idx);
  if (N < 1 || !conv_fused_fwd_serves(C, H, W, O, KH, KW, pad, stride, Ho, Wo) || ((uintptr_t)x & 7u)) return NGB_OK;
  FusedFwdGeo g;
  g.N = N; g.H = H; g.W = W; g.Ho = Ho; g.Hp = Ho / 2; g.Fp = (Ho / 2) * (Wo / 2);
  g.ximg = C * H * W; g.tpi = g.Fp; g.G = 1;
  // staged row pitch: even (64-bit patch loads); lanes walk p' first, then q', and read pair (S*q' + p') mod 16 of the 16
  // bank pairs: S = Hp (mod 16) lets a half-warp cover all of them once
  g.S = W + (W & 1);
  for (int Sx = W + (W & 1); Sx < W + 16; Sx += 2)
    if ((Sx & 15) == (g.Hp & 15)) { g.S = Sx; break; }
  g.simg = C * H * g.S;
  int rc = ensure_init();
  if (rc) return rc;
  rc = -1;
  if (C == 1 && KH == 5 && KW == 5 && O == 6) rc = launch_fused_fwd<1, 5, 5, 6>(x, w, bias, y, idx, g, st);
  else if (C == 1 && KH == 3 && KW == 3 && O == 4) rc = launch_fused_fwd<1, 3, 3, 4>(x, w, bias, y, idx, g, st);
  else if (C == 1 && KH == 3 && KW == 3 && O == 8) rc = launch_fused_fwd<1, 3, 3, 8>(x, w, bias, y, idx, g, st);
  if (rc == -1) return NGB_OK;
  if (rc) return rc;
  *done = true;
  return NGB_OK;
}

}  // namespace ngb
