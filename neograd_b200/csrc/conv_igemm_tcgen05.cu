// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 / TMEM), operands fed by TMA.
// Covers the multi-channel, stride-1 layers of the C5 sweep (Conv3D, neograd/autograd/ops/conv.py:245-269 forward,
// 299-309 data gradient, 311-321 weight gradient); everything else stays on the direct kernels (conv_direct.cu).
//
// Forward and data-gradient are the same "transposing" convolution in neograd's layouts:
//     out[n, co, j1, j0] = bias[co] + sum_{t1,t0,ci} in[n, ci, i1 = j0 + t1 - pad1, i0 = j1 + t0 - pad0] * Wp[t1*T0+t0][co][ci]
// where i0 / j0 are the CONTIGUOUS axes of in / out; the output's fast axis runs along the input's slow spatial axis:
//   forward : in = x  (i0 = w, i1 = h),  out = y  in neograd order (j0 = p along h, j1 = q along w; conv.py:39-47 vs 149-151)
//   dgrad   : in = ug (i0 = p, i1 = q),  out = dx natural layout   (j0 = w, j1 = h), filter flipped, pad' = K-1-pad
//
// Why a staging pass: a filter tap shifts the input by one ELEMENT along its contiguous axis, and a TMA box must start
// on a 16-byte boundary in global memory (tools/tma_probe.cu: a swizzled box at inner coordinate -1 faults).  So the
// input is first re-laid out channels-innermost, S[n][i0][i1][c] (one read + one write of the input, rounding to TF32
// on the way), after which every tap is a shift on OUTER tensor-map dimensions and padding is TMA's zero fill.
//
// GEMM view per work item: D[128 positions, BN channels] += A[128, K] * B[K, BN],  K = taps x Cin.
//   * A (K-major, 128B swizzle): ONE 4-D box {32 c, 32 along i1, 4 along i0, 1 image} per K-block = (tap, 32 channels):
//     row r = di0*32 + di1, so TMEM lane (r mod 32) walks the output's contiguous axis j0 and the epilogue stores are
//     128-byte coalesced with no shuffle.
//   * B (K-major): weights pre-packed once per call into Wp[tap][co][ci] (scratch arena), boxes {32 ci, BN co}.
//   * accumulators in TMEM (2 x BN columns, double buffered); epilogue warps tcgen05.ld -> (+bias / +=) -> global.
// Weight gradient: dw[o,c,tap] = sum_positions ug[pos, o] * x[pos + tap, c] with BOTH operands read from their staged
// (channels-innermost = MN-major) form, K = positions, split over CTAs with a deterministic second pass.
// Warp roles are those of gemm_tcgen05.cu: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue.
#include "tc_common.cuh"

#include <mutex>

namespace ngb {

using namespace tc;

int launch_splitk_reduce(const float* part, int splits, int64_t M, int64_t N, float* C, int64_t ldc, int accumulate,
                         cudaStream_t st);

namespace {

constexpr int kThreads = 192;
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;       // channels per K-block
constexpr int UMMA_K = 8;
constexpr int TILE_J0 = 32;       // output extent along its contiguous axis (= i1 direction): one 128-byte line per warp store
constexpr int TILE_J1 = 4;        // output extent along j1 (= i0 direction): one TMEM lane quadrant each

// TG = taps along i0 that share ONE A box per stage.  The box covers TILE_J1 + TG - 1 rows of i0; tap t0 of the group is
// the 128-row window starting t0 * 32 rows (= t0 * 4 KB, a multiple of the 1 KB swizzle atom) into it, so the A operand
// is fetched once per (t1, channel block) instead of once per tap.  The kernel is bound by the L2 -> SMEM fill rate
// (~60 B/clk/SM against 128-190 B/clk the MMAs of a 64/128-column tile could consume), so bytes per MMA are what counts.
template <int BLOCK_N, int TG>
struct Cfg {
  static constexpr int kABytes = (TILE_J1 + TG - 1) * TILE_J0 * BLOCK_K * 4;   // 16 KB for TG = 1, 24 KB for TG = 3
  static constexpr int kBTap = BLOCK_N * BLOCK_K * 4;
  static constexpr int kBBytes = TG * kBTap;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (218 * 1024) / kStageBytes > 8 ? 8 : (218 * 1024) / kStageBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static_assert(kStages >= 2, "pipeline needs two stages");
};

struct IgemmArgs {
  int N, Cin, Cout;
  int I0, I1, J0, J1;        // input / output spatial extents (contiguous axis first)
  int T0, T1, pad0, pad1;    // taps along i0 / i1 and the padding on those axes
  int tiles_j0, tiles_j1, tiles_co;
  int cblocks;               // Cin / 32
  float* out;
  const float* bias;         // may be null
  int accumulate;
};

// ---------------------------------------------------------------------------------------------- staging
// in[n][c][i1][i0] -> S[n][i0][i1][c] (TF32-rounded).  32 x 32 (c, i0) tiles through padded smem: reads coalesced
// along i0, writes coalesced along c.
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(256) stage_channels_last_kernel(const float* __restrict__ in, float* __restrict__ S, int N, int C,
                                                                  int I1, int I0, FastDiv d_t0, FastDiv d_tc, FastDiv d_i1) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int t0n = (I0 + 31) >> 5, tcn = (C + 31) >> 5;
  const int64_t items = (int64_t)N * I1 * tcn * t0n;
  for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
    uint32_t r, b0, bc, n, i1;
    d_t0.divmod((uint32_t)it, r, b0);   // items < 2^32 is checked by the caller
    d_tc.divmod(r, r, bc);
    d_i1.divmod(r, n, i1);
    const int c0 = bc * 32, i00 = b0 * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int c = c0 + ty + j, i0 = i00 + tx;
      tile[ty + j][tx] = (c < C && i0 < I0) ? __ldg(in + (((int64_t)n * C + c) * I1 + i1) * I0 + i0) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int i0 = i00 + ty + j, c = c0 + tx;
      if (i0 < I0 && c < C) S[(((int64_t)n * I0 + i0) * I1 + i1) * C + c] = to_tf32(tile[tx][ty + j]);
    }
    __syncthreads();
  }
}

// Same re-layout with 64 x 64 (c, i0) tiles and 128-bit global accesses on both sides (I0 % 4 == 0, C % 64 == 0):
// 16 KB per CTA iteration and four independent 128-bit loads in flight per thread.
__global__ void __launch_bounds__(256) stage_channels_last_v4_kernel(const float* __restrict__ in, float* __restrict__ S, int N, int C,
                                                                     int I1, int I0, FastDiv d_t0, FastDiv d_tc, FastDiv d_i1) {
  __shared__ float tile[64][65];
  const int l16 = threadIdx.x & 15, r16 = threadIdx.x >> 4;   // 16 float4 columns x 16 rows per pass
  const int t0n = (I0 + 63) >> 6, tcn = C >> 6;
  const int64_t items = (int64_t)N * I1 * tcn * t0n;
  for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
    uint32_t r, b0, bc, n, i1;
    d_t0.divmod((uint32_t)it, r, b0);
    d_tc.divmod(r, r, bc);
    d_i1.divmod(r, n, i1);
    const int c0 = bc * 64, i00 = b0 * 64;
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + r16 + 16 * k, i0 = i00 + 4 * l16;
      v[k] = i0 < I0 ? ld_stream4(in + (((int64_t)n * C + c) * I1 + i1) * I0 + i0) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float* d = &tile[r16 + 16 * k][4 * l16];
      d[0] = v[k].x; d[1] = v[k].y; d[2] = v[k].z; d[3] = v[k].w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i0l = r16 + 16 * k, i0 = i00 + i0l;
      if (i0 < I0) {
        float4 o;
        o.x = to_tf32(tile[4 * l16 + 0][i0l]); o.y = to_tf32(tile[4 * l16 + 1][i0l]);
        o.z = to_tf32(tile[4 * l16 + 2][i0l]); o.w = to_tf32(tile[4 * l16 + 3][i0l]);
        *reinterpret_cast<float4*>(S + (((int64_t)n * I0 + i0) * I1 + i1) * C + c0 + 4 * l16) = o;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- fprop / dgrad kernel
// CL = 2: the two CTAs of a cluster work on two position tiles of the SAME channel tile in lock step; each fetches half
// of every weight box and multicasts it to both, which halves the B bytes each SM pulls from L2.
template <int BLOCK_N, int TG, int CL>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w, const IgemmArgs g) {
  using C_ = Cfg<BLOCK_N, TG>;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr int kStages = C_::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * C_::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * C_::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_in);
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C_::kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();      // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kblocks = g.T1 * (g.T0 / TG) * g.cblocks;      // stages per tile (host guarantees T0 % TG == 0)
  const int tiles_sp = g.tiles_j0 * g.tiles_j1;
  const int64_t m_tiles = (int64_t)g.N * tiles_sp;
  // One loop index for every role.  CL = 1: w = work item (co fastest so CTAs that run together share the A tile in L2).
  // CL = 2: w = (pair of position tiles, co tile); this CTA takes tile 2*pair + rank.  A cluster whose second tile does
  // not exist still runs the pipeline on a clamped tile (barriers stay in lock step) and skips the stores.
  const int64_t num_work = (CL > 1 ? (m_tiles + 1) / 2 : m_tiles) * g.tiles_co;
  const int64_t w_first = CL > 1 ? (int64_t)(blockIdx.x / CL) : (int64_t)blockIdx.x;
  const int64_t w_step = CL > 1 ? (int64_t)(gridDim.x / CL) : (int64_t)gridDim.x;
  auto decode = [&](int64_t w, int& n, int& j0b, int& j1b, int& co0) -> bool {
    const int ct = (int)(w % g.tiles_co);
    int64_t r = w / g.tiles_co;
    bool valid = true;
    if (CL > 1) {
      r = 2 * r + crank;
      if (r >= m_tiles) { r = m_tiles - 1; valid = false; }
    }
    const int sp = (int)(r % tiles_sp);
    n = (int)(r / tiles_sp);
    j0b = (sp % g.tiles_j0) * TILE_J0;
    j1b = (sp / g.tiles_j0) * TILE_J1;
    co0 = ct * BLOCK_N;
    return valid;
  };

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t w = w_first; w < num_work; w += w_step) {
        int n, j0b, j1b, co0;
        decode(w, n, j0b, j1b, co0);
        for (int t1 = 0; t1 < g.T1; ++t1) {
          for (int t0 = 0; t0 < g.T0; t0 += TG) {
            for (int cb = 0; cb < g.cblocks; ++cb, ++it) {
              const int s = it % kStages;
              const uint32_t ph = (it / kStages) & 1;
              mbar_wait(&empty[s], ph ^ 1, 1);
              mbar_arrive_expect_tx(&full[s], C_::kStageBytes);
              // staged input dims {c, i1, i0, n}: the tap is a shift of the two spatial (outer) coordinates; the box
              // spans the i0 rows of all TG taps of the group
              tma_load_4d(smem_a + s * C_::kABytes, &tmap_in, &full[s], cb * BLOCK_K, j0b + t1 - g.pad1, j1b + t0 - g.pad0, n);
#pragma unroll
              for (int u = 0; u < TG; ++u) {
                if (CL > 1) {   // this CTA's half of the weight box goes to both CTAs
                  constexpr int kHalfRows = BLOCK_N / 2;
                  tma_load_2d_mc(smem_b + s * C_::kBBytes + u * C_::kBTap + crank * (kHalfRows * BLOCK_K * 4), &tmap_w, &full[s],
                                 cb * BLOCK_K, (t1 * g.T0 + t0 + u) * g.Cout + co0 + (int)crank * kHalfRows, (uint16_t)0x3);
                } else {
                  tma_load_2d(smem_b + s * C_::kBBytes + u * C_::kBTap, &tmap_w, &full[s], cb * BLOCK_K,
                              (t1 * g.T0 + t0 + u) * g.Cout + co0);
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M, BLOCK_N, 0, 0);
      constexpr uint64_t desc_base = make_smem_desc_base(16, 1024);   // K-major, 128B swizzle (both operands)
      uint32_t it = 0, tile_it = 0;
      for (int64_t w = w_first; w < num_work; w += w_step, ++tile_it) {
        const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&full[s], ph, 3);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + s * C_::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + s * C_::kBBytes);
#pragma unroll
          for (int u = 0; u < TG; ++u) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              umma_tf32(d_tmem, smem_desc(desc_base, a_addr + u * (TILE_J0 * BLOCK_K * 4) + k * UMMA_K * 4),
                        smem_desc(desc_base, b_addr + u * C_::kBTap + k * UMMA_K * 4), idesc, (kb > 0 || u > 0 || k > 0) ? 1u : 0u);
            }
          }
          if (CL > 1) umma_commit_mc(&empty[s], (uint16_t)0x3);   // the stage is free once BOTH CTAs have consumed it
          else umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int quad = warp & 3;                // TMEM lane quadrant == offset along j1; lane == offset along j0
    uint32_t tile_it = 0;
    const int64_t plane = (int64_t)g.J0 * g.J1;
    for (int64_t w = w_first; w < num_work; w += w_step, ++tile_it) {
      int n, j0b, j1b, co0;
      const bool tile_valid = decode(w, n, j0b, j1b, co0);
      const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_ph, 4);
      tc_fence_after();
      const int j0 = j0b + lane, j1 = j1b + quad;
      const bool ok = tile_valid && j0 < g.J0 && j1 < g.J1;
      float* obase = g.out + ((int64_t)n * g.Cout + co0) * plane + (int64_t)j1 * g.J0 + j0;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + (uint32_t(quad * 32) << 16) + acc * BLOCK_N + c0, v);
        tmem_ld_wait();
        if (c0 + 32 >= BLOCK_N) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        // bias of the chunk's 32 channels: one coalesced load, broadcast by shuffle (was 32 loads per thread)
        float bv = 0.f;
        if (g.bias && co0 + c0 + lane < g.Cout) bv = __ldg(g.bias + co0 + c0 + lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bv, j);
        if (ok) {
          if (g.accumulate) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (co0 + c0 + j < g.Cout) { float* o = obase + (int64_t)(c0 + j) * plane; *o += v[j]; }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (co0 + c0 + j < g.Cout) obase[(int64_t)(c0 + j) * plane] = v[j];
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();      // no CTA leaves while its peer can still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C_::kTmemCols);
  }
}

// Wp[tap][co][ci]: forward  tap = kh*KW + kw,                         co = o, ci = c
//                  dgrad    tap = t1*T0 + t0 with kh = KH-1-t0, kw = KW-1-t1 (T0 = KH, T1 = KW), co = c, ci = o
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int O, int C,
                                                           int KH, int KW, int dgrad) {
  const int64_t total = (int64_t)O * C * KH * KW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i enumerates the destination so writes are coalesced
    if (!dgrad) {
      const int c = (int)(i % C);
      const int o = (int)((i / C) % O);
      const int tap = (int)(i / ((int64_t)C * O));
      const int kh = tap / KW, kw = tap % KW;
      wp[i] = to_tf32(__ldg(w + (((int64_t)o * C + c) * KH + kh) * KW + kw));
    } else {
      const int o = (int)(i % O);
      const int c = (int)((i / O) % C);
      const int tap = (int)(i / ((int64_t)C * O));
      const int t1 = tap / KH, t0 = tap % KH;
      const int kh = KH - 1 - t0, kw = KW - 1 - t1;
      wp[i] = to_tf32(__ldg(w + (((int64_t)o * C + c) * KH + kh) * KW + kw));
    }
  }
}

// ---------------------------------------------------------------------------------------------- workspace
// Staged copies of activations can be GBs (C5: 4.3 GB), so they live in a grow-on-demand buffer of their own rather
// than in the fixed scratch arena.  Growing allocates, so it must not happen inside a CUDA-graph capture: callers that
// capture call ngb_workspace_reserve() first.
std::mutex g_ws_mu;
float* g_ws_lane[kLanes] = {nullptr, nullptr, nullptr, nullptr};   // one per lane (ngb_set_lane): lanes overlap on the GPU
int64_t g_ws_bytes_lane[kLanes] = {0, 0, 0, 0};

int workspace(int64_t bytes, float** out, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  float*& g_ws = g_ws_lane[current_lane()];
  int64_t& g_ws_bytes = g_ws_bytes_lane[current_lane()];
  if (bytes > g_ws_bytes) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone)
      return fail(NGB_ERR_SCRATCH, "conv workspace (%lld bytes) must be reserved with ngb_workspace_reserve() before graph capture",
                  (long long)bytes);
    if (g_ws) {
      NGB_CUDA(cudaDeviceSynchronize());
      NGB_CUDA(cudaFree(g_ws));
      g_ws = nullptr;
      g_ws_bytes = 0;
    }
    const int64_t want = bytes + (bytes >> 3);
    NGB_CUDA(cudaMalloc(&g_ws, want));
    g_ws_bytes = want;
  }
  *out = g_ws;
  return NGB_OK;
}

int stage(const float* in, float* S, int64_t N, int64_t C, int64_t I1, int64_t I0, cudaStream_t st) {
  if (C % 64 == 0 && I0 % 4 == 0 && aligned16(in) && aligned16(S)) {
    const int64_t t0v = ceil_div<int64_t>(I0, 64), tcv = C / 64;
    const int64_t itemsv = N * I1 * tcv * t0v;
    if (itemsv < (int64_t(1) << 32)) {
      const int gridv = (int)(itemsv < (int64_t)kNumSMs * 8 ? itemsv : (int64_t)kNumSMs * 8);
      stage_channels_last_v4_kernel<<<gridv, 256, 0, st>>>(in, S, (int)N, (int)C, (int)I1, (int)I0, FastDiv((uint32_t)t0v),
                                                          FastDiv((uint32_t)tcv), FastDiv((uint32_t)I1));
      NGB_LAUNCH_CHECK();
      return NGB_OK;
    }
  }
  const int64_t t0n = ceil_div<int64_t>(I0, 32), tcn = ceil_div<int64_t>(C, 32);
  const int64_t items = N * I1 * tcn * t0n;
  if (items >= (int64_t(1) << 32)) return fail(NGB_ERR_UNSUPPORTED, "conv igemm: tensor too large to stage");
  const int grid = (int)(items < (int64_t)kNumSMs * 16 ? items : (int64_t)kNumSMs * 16);
  stage_channels_last_kernel<<<grid, 256, 0, st>>>(in, S, (int)N, (int)C, (int)I1, (int)I0, FastDiv((uint32_t)t0n),
                                                  FastDiv((uint32_t)tcn), FastDiv((uint32_t)I1));
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

template <int BN, int TG, int CL>
int launch_igemm(const CUtensorMap& tin, const CUtensorMap& tw, const IgemmArgs& g, int64_t work, cudaStream_t st) {
  using C_ = Cfg<BN, TG>;
  auto kern = conv_igemm_kernel<BN, TG, CL>;
  static bool attr_set = false;
  static int max_clusters = 0;
  if (!attr_set) {
    NGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::kSmemBytes));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C_::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int grid;
  if (CL > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_clusters == 0) {   // clusters that can be resident at once (GPC boundaries can leave an SM unpaired)
      cfg.gridDim = dim3(kNumSMs / CL * CL);
      NGB_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
      if (max_clusters < 1) max_clusters = 1;
    }
    const int64_t pairs = (work + CL - 1) / CL;   // upper bound on cluster work items per channel tile set
    const int64_t clusters = pairs < max_clusters ? pairs : max_clusters;
    grid = (int)clusters * CL;
  } else {
    grid = (int)(work < kNumSMs ? work : kNumSMs);
  }
  cfg.gridDim = dim3(grid);
  NGB_CUDA(cudaLaunchKernelEx(&cfg, kern, tin, tw, g));
  count_launch();
  return NGB_OK;
}

// The shared host path.  `in` is [N][Cin][I1][I0]; out is [N][Cout][J1][J0].
int run_igemm(const float* in, const float* w, const float* bias, float* out, int64_t N, int64_t Cin, int64_t Cout, int64_t I0,
              int64_t I1, int64_t T0, int64_t T1, int64_t pad0, int64_t pad1, int64_t O, int64_t C, int64_t KH, int64_t KW,
              int dgrad, int accumulate, cudaStream_t st) {
  int rc = ensure_init();
  if (rc) return rc;
  const int64_t wp_elems = O * C * KH * KW;
  if (wp_elems * 4 > scratch_bytes()) return fail(NGB_ERR_SCRATCH, "conv igemm: packed weights do not fit the scratch arena");
  float* wp = scratch_ptr();
  pack_weights_kernel<<<grid_for(wp_elems, 256), 256, 0, st>>>(w, wp, (int)O, (int)C, (int)KH, (int)KW, dgrad);
  NGB_LAUNCH_CHECK();

  float* S = nullptr;
  rc = workspace(N * Cin * I0 * I1 * 4, &S, st);
  if (rc) return rc;
  rc = stage(in, S, N, Cin, I1, I0, st);
  if (rc) return rc;

  IgemmArgs g;
  g.N = (int)N; g.Cin = (int)Cin; g.Cout = (int)Cout;
  g.I0 = (int)I0; g.I1 = (int)I1;
  g.J0 = (int)(I1 + 2 * pad1 - T1 + 1);
  g.J1 = (int)(I0 + 2 * pad0 - T0 + 1);
  g.T0 = (int)T0; g.T1 = (int)T1; g.pad0 = (int)pad0; g.pad1 = (int)pad1;
  g.tiles_j0 = ceil_div(g.J0, TILE_J0);
  g.tiles_j1 = ceil_div(g.J1, TILE_J1);
  const int BN = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
  g.tiles_co = (int)ceil_div<int64_t>(Cout, BN);
  g.cblocks = (int)(Cin / BLOCK_K);
  g.out = out; g.bias = bias; g.accumulate = accumulate;

  // taps along i0 sharing one A box: 3 for 3-tap filters on the narrow tiles (their stage still fits 3-4 times)
  // taps along i0 sharing one A box: 3 for 3-tap filters on the narrow tiles (a 256-column stage would not fit twice)
  const int TG = (T0 == 3 && BN <= 128) ? 3 : 1;
  // CL = 2 (weight boxes multicast over a 2-CTA cluster) is implemented and parity-tested, but measured no faster on
  // B200: the L2 -> SM fabric delivers a copy per destination SM, so the bytes each SM receives are unchanged.
  constexpr int CL = 1;
  CUtensorMap tin, tw;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)I1, (uint64_t)I0, (uint64_t)N};
    const uint64_t str[3] = {(uint64_t)Cin * 4, (uint64_t)Cin * I1 * 4, (uint64_t)Cin * I1 * I0 * 4};
    const uint32_t box[4] = {BLOCK_K, TILE_J0, (uint32_t)(TILE_J1 + TG - 1), 1};
    rc = make_tmap_f32(&tin, S, 4, dims, str, box, false);
    if (rc) return rc;
    const uint64_t wdims[2] = {(uint64_t)Cin, (uint64_t)(T0 * T1 * Cout)};
    const uint64_t wstr[1] = {(uint64_t)Cin * 4};
    const uint32_t wbox[2] = {BLOCK_K, (uint32_t)(BN / CL)};     // CL = 2: each CTA of the pair fetches half of the rows
    rc = make_tmap_f32(&tw, wp, 2, wdims, wstr, wbox, false);
    if (rc) return rc;
  }
  const int64_t work = (int64_t)N * g.tiles_j0 * g.tiles_j1 * g.tiles_co;
  if (BN == 64) rc = TG == 3 ? launch_igemm<64, 3, CL>(tin, tw, g, work, st) : launch_igemm<64, 1, CL>(tin, tw, g, work, st);
  else if (BN == 128) rc = TG == 3 ? launch_igemm<128, 3, CL>(tin, tw, g, work, st) : launch_igemm<128, 1, CL>(tin, tw, g, work, st);
  else rc = launch_igemm<256, 1, CL>(tin, tw, g, work, st);
  return rc;
}


// ---------------------------------------------------------------------------------------------- weight gradient
// dw[o,c,kh,kw] = sum_{n,p,q} ug[n,o,p,q] * x[n,c,p+kh-pad,q+kw-pad]   (conv.py:311-321)
// GEMM per tap: D_tap[o, c] = sum_pos ug[o, pos] * Xs[pos + tap, c] with K = positions.  ug is read IN PLACE (its p axis
// is contiguous, i.e. K-major); x comes from its channels-last staged copy (MN-major).  A K-block is 32 consecutive p
// at one (n, q): the ug tile is loaded once (one 4-D box) and multiplied against G shifted x tiles (one 5-D box per
// tap of the CTA's tap group, each tap with its own TMEM accumulator).
// Work item = (128-wide o tile, BN-wide c tile, tap group, K split); partials go to the scratch arena and a second pass
// sums the splits deterministically into dw's (o,c,kh,kw) layout.
struct WgradArgs {
  int N, O, C, Ho, Wo, KH, KW, pad;
  int o_tiles, c_tiles, tap_groups, splits;
  int qblocks;               // ceil(Ho / 32): K-blocks of 32 consecutive p per (n, q)
  int64_t kblocks;           // N * Wo * qblocks
  int64_t kb_per_split;
  float* partial;            // [split][tap][O][C]
};

template <int BLOCK_N, int G>
struct WCfg {
  static constexpr int kABytes = 128 * 32 * 4;              // 4 chunks [32 pos][32 o]
  static constexpr int kBBytes = G * BLOCK_N * 32 * 4;      // G taps x (BN/32) chunks [32 pos][32 c]
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 6 ? 6 : (200 * 1024) / kStageBytes;
  static constexpr int kTmemCols = G * BLOCK_N < 32 ? 32 : G * BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static_assert(G * BLOCK_N <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM budget / power of two");
};

template <int BLOCK_N, int G>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_igemm_kernel(const __grid_constant__ CUtensorMap tmap_ug, const __grid_constant__ CUtensorMap tmap_x, const WgradArgs g) {
  using C_ = WCfg<BLOCK_N, G>;
  constexpr int kStages = C_::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * C_::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * C_::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_ug);
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C_::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int taps = g.KH * g.KW;
  const int64_t num_work = (int64_t)g.o_tiles * g.c_tiles * g.tap_groups * g.splits;

  // tap group fastest: CTAs that run side by side stream the SAME K range (the same ug / x positions) for different
  // taps, so the re-reads across tap groups are L2 hits instead of HBM traffic
  auto decode = [&](int64_t w, int& o0, int& c0, int& tap0, int& ntap, int& split) {
    const int tg = (int)(w % g.tap_groups);
    int64_t r = w / g.tap_groups;
    c0 = (int)(r % g.c_tiles) * BLOCK_N;
    r /= g.c_tiles;
    o0 = (int)(r % g.o_tiles) * 128;
    split = (int)(r / g.o_tiles);
    tap0 = tg * G;
    ntap = min(G, taps - tap0);
  };

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
        int o0, c0, tap0, ntap, split;
        decode(w, o0, c0, tap0, ntap, split);
        const int64_t kb0 = (int64_t)split * g.kb_per_split;
        const int64_t kb1 = min(g.kblocks, kb0 + g.kb_per_split);
        // (n, q, pb) of the first K-block; advanced incrementally (no 64-bit divides in the issue loop)
        int pb = (int)(kb0 % g.qblocks);
        int64_t r0 = kb0 / g.qblocks;
        int q = (int)(r0 % g.Wo), n = (int)(r0 / g.Wo);
        for (int64_t kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty[s], ph ^ 1, 11);
          mbar_arrive_expect_tx(&full[s], C_::kABytes + ntap * BLOCK_N * 128);
          uint8_t* sa = smem_a + s * C_::kABytes;
          uint8_t* sb = smem_b + s * C_::kBBytes;
          // A straight from ug [n][o][q][p] (p contiguous = K-major): box {32 p, 1 q, 128 o, 1}: no staging pass for ug
          tma_load_4d(sa, &tmap_ug, &full[s], pb * 32, q, o0, n);
          // staged x viewed as {32 c_in, H, W, N, C/32 c_out}: one box {32, 32 (h), 1 (w), 1, BN/32} per tap, rows ordered by h
          for (int t = 0; t < ntap; ++t) {
            const int tap = tap0 + t, kh = tap / g.KW, kw = tap - kh * g.KW;
            tma_load_5d(sb + t * (BLOCK_N / 32) * 4096, &tmap_x, &full[s], 0, pb * 32 + kh - g.pad, q + kw - g.pad, n, c0 / 32);
          }
          if (++pb == g.qblocks) { pb = 0; if (++q == g.Wo) { q = 0; ++n; } }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint64_t adesc_base = make_smem_desc_base(16, 1024);                          // ug: K-major, 128B swizzle
      constexpr uint64_t desc_base = make_smem_desc_base(4096, 512, kLayoutSW128Base32B);   // x: MN-major, 32B-atom swizzle
      uint32_t it = 0, tile_it = 0;
      for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x, ++tile_it) {
        int o0, c0, tap0, ntap, split;
        decode(w, o0, c0, tap0, ntap, split);
        const int64_t kb0 = (int64_t)split * g.kb_per_split;
        const int64_t kb1 = min(g.kblocks, kb0 + g.kb_per_split);
        mbar_wait(tmem_empty, (tile_it & 1) ^ 1, 12);
        tc_fence_after();
        const uint32_t idesc_n = make_idesc_tf32(128, ntap * BLOCK_N, 0, 1);
        for (int64_t kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&full[s], ph, 13);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + s * C_::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + s * C_::kBBytes);
          // the ntap x-tiles sit back to back as [tap][c chunk][32 pos][32 c] = ONE MN-major B operand with N = ntap*BN
          // columns, so a single MMA per k-step covers every tap of the group (A is read once, not once per tap)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_tf32(tmem_base, smem_desc(adesc_base, a_addr + k * UMMA_K * 4), smem_desc(desc_base, b_addr + k * 1024), idesc_n,
                      (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(tmem_full);
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue: TMEM lane = o, column = c
    const int quad = warp & 3;
    uint32_t tile_it = 0;
    for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x, ++tile_it) {
      int o0, c0, tap0, ntap, split;
      decode(w, o0, c0, tap0, ntap, split);
      const int64_t kb0 = (int64_t)split * g.kb_per_split;
      const bool has_k = kb0 < g.kblocks;
      mbar_wait(tmem_full, tile_it & 1, 14);
      tc_fence_after();
      const int o = o0 + quad * 32 + lane;
      for (int t = 0; t < ntap; ++t) {
        float* prow = g.partial + (((int64_t)split * taps + tap0 + t) * g.O + o) * g.C + c0;
#pragma unroll 1
        for (int cc = 0; cc < BLOCK_N; cc += 32) {
          float v[32];
          tmem_ld32(tmem_base + (uint32_t(quad * 32) << 16) + t * BLOCK_N + cc, v);
          tmem_ld_wait();
          if (o < g.O) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (c0 + cc + j < g.C) {   // C % 32 == 0: whole float4 groups
                const float4 o4 = has_k ? make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(prow + cc + j) = o4;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C_::kTmemCols);
  }
}

// dw[o][c][tap] (+)= sum_split partial[split][tap][o][c]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int splits, int taps,
                                                           int O, int C, int accumulate) {
  const int64_t total = (int64_t)taps * O * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(int64_t)z * total + i];
    const int c = (int)(i % C);
    const int o = (int)((i / C) % O);
    const int tap = (int)(i / ((int64_t)C * O));
    float* d = dw + ((int64_t)o * C + c) * taps + tap;
    *d = accumulate ? *d + s : s;
  }
}

template <int BN, int G>
int launch_wgrad(const CUtensorMap& tu, const CUtensorMap& tx, const WgradArgs& g, int grid, cudaStream_t st) {
  using C_ = WCfg<BN, G>;
  auto kern = conv_wgrad_igemm_kernel<BN, G>;
  static bool attr_set = false;
  if (!attr_set) {
    NGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::kSmemBytes));
    attr_set = true;
  }
  kern<<<grid, kThreads, C_::kSmemBytes, st>>>(tu, tx, g);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

}  // namespace

int conv_igemm_eligible(int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW, int64_t pad,
                        int64_t stride) {
  if (g_default_engine.load() == NGB_GEMM_SIMT) return 0;
  if (stride != 1) return 0;
  if (C % 32 || O % 32) return 0;                       // K-blocks of 32 channels for both passes
  if (pad > KH - 1 || pad > KW - 1) return 0;            // dgrad's pad' = K-1-pad must stay >= 0
  const int64_t Ho = H + 2 * pad - KH + 1, Wo = W + 2 * pad - KW + 1;
  if (Ho < 1 || Wo < 1) return 0;
  if (N * Ho * Wo < 4096) return 0;                      // tiny problems: the direct kernel is launch-bound anyway
  if (N > 0x7fffffff || (int64_t)C * H * W > 0x7fffffff) return 0;
  return 1;
}

int conv_igemm_fprop(const float* x, const float* w, const float* b, float* y, int64_t N, int64_t C, int64_t H, int64_t W,
                     int64_t O, int64_t KH, int64_t KW, int64_t pad, cudaStream_t st) {
  // in = x: i0 = w (taps kw), i1 = h (taps kh)
  return run_igemm(x, w, b, y, N, C, O, /*I0=*/W, /*I1=*/H, /*T0=*/KW, /*T1=*/KH, pad, pad, O, C, KH, KW, 0, 0, st);
}

int conv_igemm_dgrad(const float* ug, const float* w, float* dx, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O,
                     int64_t KH, int64_t KW, int64_t pad, int accumulate, cudaStream_t st) {
  const int64_t Ho = H + 2 * pad - KH + 1, Wo = W + 2 * pad - KW + 1;
  // in = ug (neograd order): i0 = p (taps along kh, flipped), i1 = q (taps along kw, flipped)
  return run_igemm(ug, w, nullptr, dx, N, O, C, /*I0=*/Ho, /*I1=*/Wo, /*T0=*/KH, /*T1=*/KW, KH - 1 - pad, KW - 1 - pad, O, C, KH,
                   KW, 1, accumulate, st);
}

int conv_igemm_wgrad(const float* ug, const float* x, float* dw, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O,
                     int64_t KH, int64_t KW, int64_t pad, int accumulate, cudaStream_t st) {
  int rc = ensure_init();
  if (rc) return rc;
  const int64_t Ho = H + 2 * pad - KH + 1, Wo = W + 2 * pad - KW + 1;
  const int taps = (int)(KH * KW);
  // x staged channels-last as Xs[n][w][h][c]; ug is consumed in place (its p axis is contiguous = K-major)
  float* Xs = nullptr;
  rc = workspace(N * C * H * W * 4, &Xs, st);
  if (rc) return rc;
  rc = stage(x, Xs, N, C, /*I1=*/H, /*I0=*/W, st);       // x is [n][c][h][w]:  i0 = w, i1 = h
  if (rc) return rc;

  WgradArgs g;
  g.N = (int)N; g.O = (int)O; g.C = (int)C; g.Ho = (int)Ho; g.Wo = (int)Wo; g.KH = (int)KH; g.KW = (int)KW; g.pad = (int)pad;
  const int BN = C <= 64 ? 64 : (C <= 128 ? 128 : 256);
  const int G = 256 / BN;
  g.o_tiles = (int)ceil_div<int64_t>(O, 128);
  g.c_tiles = (int)ceil_div<int64_t>(C, BN);
  g.tap_groups = ceil_div(taps, G);
  g.qblocks = (int)ceil_div<int64_t>(Ho, 32);
  g.kblocks = N * Wo * g.qblocks;
  const int64_t items = (int64_t)g.o_tiles * g.c_tiles * g.tap_groups;
  int64_t splits = ceil_div<int64_t>(2 * kNumSMs, items);
  const int64_t per_split_bytes = (int64_t)taps * O * C * 4;
  if (splits * per_split_bytes > scratch_bytes()) splits = scratch_bytes() / per_split_bytes;
  if (splits > g.kblocks) splits = g.kblocks;
  if (splits < 1) return fail(NGB_ERR_SCRATCH, "conv wgrad igemm: scratch arena too small");
  g.kb_per_split = ceil_div<int64_t>(g.kblocks, splits);
  g.splits = (int)ceil_div<int64_t>(g.kblocks, g.kb_per_split);
  g.partial = scratch_ptr();

  CUtensorMap tu, tx;
  {
    const uint64_t ud[4] = {(uint64_t)Ho, (uint64_t)Wo, (uint64_t)O, (uint64_t)N};
    const uint64_t us[3] = {(uint64_t)Ho * 4, (uint64_t)Ho * Wo * 4, (uint64_t)Ho * Wo * O * 4};
    const uint32_t ub[4] = {32, 1, 128, 1};
    rc = make_tmap_f32(&tu, ug, 4, ud, us, ub, false);
    if (rc) return rc;
    // channel dim split into (32 inner, C/32 outer) so that one 5-D box delivers BN/32 [32 pos][32 ch] chunks back to back
    const uint64_t xd[5] = {32, (uint64_t)H, (uint64_t)W, (uint64_t)N, (uint64_t)(C / 32)};
    const uint64_t xs[4] = {(uint64_t)C * 4, (uint64_t)C * H * 4, (uint64_t)C * H * W * 4, 128};
    const uint32_t xb[5] = {32, 32, 1, 1, (uint32_t)(BN / 32)};
    rc = make_tmap_f32(&tx, Xs, 5, xd, xs, xb, true);
    if (rc) return rc;
  }
  const int64_t work = items * g.splits;
  const int grid = (int)(work < kNumSMs ? work : kNumSMs);
  if (BN == 64) rc = launch_wgrad<64, 4>(tu, tx, g, grid, st);
  else if (BN == 128) rc = launch_wgrad<128, 2>(tu, tx, g, grid, st);
  else rc = launch_wgrad<256, 1>(tu, tx, g, grid, st);
  if (rc) return rc;
  wgrad_reduce_kernel<<<grid_for((int64_t)taps * O * C, 256), 256, 0, st>>>(g.partial, dw, g.splits, taps, (int)O, (int)C, accumulate);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

unsigned long long conv_take_timeout() { return tc::take_mbar_timeout(); }
int conv_set_host_word(unsigned long long* dev_alias) { return tc::set_mbar_host_word(dev_alias); }

}  // namespace ngb

using namespace ngb;

extern "C" int ngb_workspace_reserve(int64_t bytes) {
  float* p = nullptr;
  return workspace(bytes, &p, nullptr);
}
