// TF32 GEMM on the 5th-generation tensor cores: TMA -> 128B-swizzled shared-memory ring -> tcgen05.mma (one issuing
// thread) -> fp32 accumulators in TMEM (double buffered) -> tcgen05.ld epilogue.  Persistent, warp specialised:
//   warp 0    TMA producer          (empty[] -> TMA loads -> full[])
//   warp 1    MMA issuer + TMEM owner (full[] -> tcgen05.mma -> commit empty[] / tmem_full[])
//   warps 2-5 epilogue              (tmem_full[] -> tcgen05.ld -> global stores -> tmem_empty[])
// Both operand majors are supported so that Dot forward (A K-major, B MN-major), data-grad (both K-major) and
// weight-grad (both MN-major) run on row-major neograd tensors without a transpose pass:
//   basics.py:194  C = A.B     -> (A: K-major, B[K,N]: MN-major)
//   basics.py:206  dA = ug.B^T -> (ug: K-major, B as [N_mma=K, K_mma=N] : K-major)
//   basics.py:207  dB = A^T.ug -> (A^T: MN-major, ug: MN-major), reduction over the batch -> split-K
// Work items are (output tile, K-split); split partials go to the scratch arena and are summed by a second,
// deterministic pass (no atomics).
#include "tc_common.cuh"

#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace ngb {

int gemm_simt(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
              int64_t ldb, float* C, int64_t ldc, int accumulate, cudaStream_t st, const tc::Epilogue* epi = nullptr);
int launch_splitk_reduce(const float* part, int splits, int64_t M, int64_t N, float* C, int64_t ldc, int accumulate,
                         cudaStream_t st);
int launch_splitk_reduce2(const float* part, int splits, int64_t M, int64_t N, float* C, int64_t ldc, int accumulate,
                          const float* cs_part, float* cs_out, int cs_accumulate, cudaStream_t st);

namespace tc {

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_f32(CUtensorMap* out, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, bool atom32b, bool plain_f32_arg) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(NGB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  static const bool plain_env = getenv("NGB_TMAP_FLOAT32") != nullptr;  // debug switch: no TF32 rounding in the TMA unit
  const bool plain_f32 = plain_env || plain_f32_arg;
  cuuint64_t gdims[5], gstr[4];
  cuuint32_t bdims[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; bdims[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, plain_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, (cuuint32_t)rank,
                   const_cast<float*>(base), gdims, gstr, bdims, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(NGB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return NGB_OK;
}

// ---------------------------------------------------------------------------------------------- kernel
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;           // 32 fp32 = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;             // tf32: 32 bytes per instruction
constexpr int kThreads = 192;
constexpr uint32_t kSpinNone = 0;

// COLSUM: the weight-gradient GEMM of a Linear layer also produces the bias gradient -- the column sums of its B operand
// (the upstream gradient) -- by multiplying every B stage with an all-ones A tile into a second accumulator (every row of
// which is the column sum).  Costs one more UMMA per k-step on the CTAs of the first row of output tiles, 16 KB of shared
// memory for the ones tile and twice the TMEM columns.
// CTAS = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) computes a 256 x BLOCK_N tile; each CTA stages its own 128 rows
// of A and HALF of the B tile (BLOCK_N / 2 columns), the leader issues the M = 256 MMAs, each CTA's TMEM receives its 128
// rows of the accumulator and each CTA runs its own epilogue.
template <int BLOCK_N, bool COLSUM = false, int CTAS = 1>
struct Cfg {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 4;
  static constexpr int kBBytes = (BLOCK_N / CTAS) * BLOCK_K * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOnesBytes = COLSUM ? kABytes : 0;
  static constexpr int kRing = 192 * 1024 - kOnesBytes;
  static constexpr int kStages = kRing / kStageBytes > 8 ? 8 : kRing / kStageBytes;
  static constexpr int kAccCols = COLSUM ? 2 * BLOCK_N : BLOCK_N;        // TMEM columns per accumulator stage
  static constexpr int kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;  // 2 accumulator stages (power of two >= 32)
  static constexpr int kEpiBytes = 2 * 128 * 128;                        // two [128 rows][32 cols] staging tiles for TMA stores
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kOnesBytes + 1024 /*align*/ + 256 /*barriers*/;
};

struct GemmArgs {
  int M, N;            // output extent
  int k_blocks;        // ceil(K / BLOCK_K)
  int tiles_m, tiles_n;
  int splits;          // K splits per output tile
  int kb_per_split;
  float* C;            // final output (splits == 1) ...
  int64_t ldc;
  float* partial;      // ... or split partials [split][M*N] (ld = N)
  int accumulate;
  int a_3d, b_3d;      // MN-major operand described by a 3-D map {32, K, MN/32}: one TMA box per stage instead of MN/32
  int tma_store;       // epilogue goes registers -> swizzled smem -> TMA store (coalesced 128B lines); needs ldc % 4 == 0
  // ---- fused epilogue (Linear layers; all optional)
  const float* bias;   // [N]: added to every output row before anything else (layers/misc.py:63)
  int act;             // NGB_EPI_*: activation applied to (acc + bias) for the SECOND output C2 (activations.py:20, 205, 55)
  float alpha;         // LeakyReLU slope
  float* C2;           // second output [M, N] (ld = ldc2): act(acc + bias); C itself may be null then (tma_store maps: tmap_c2)
  int64_t ldc2;
  int store_c;         // write C (0 when only the activation is wanted)
  const float* mask_z; // [M, N] (ld = ldm): output is multiplied by act'(mask_z) (activations.py:31, 216) -- data-grad of a Linear
  int64_t ldm;         //   whose input was act(mask_z)
  float* colsum;       // COLSUM kernels: [N] column sums of the B operand (bias gradient), or the split partials [split][N]
  int colsum_accumulate;
};

// Fused-epilogue activations on a 32-column chunk held in registers.  The activation id is uniform over the launch, so it
// is switched on OUTSIDE the element loop: only the selected formula is executed (an epilogue warp is alone on its scheduler,
// nothing hides the latency of evaluating every activation and selecting).  Ids = ngb_unary_op; anything else = identity.
__device__ __forceinline__ void epi_act32(int act, float alpha, float (&v)[32]) {
  switch (act) {
    case NGB_RELU:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      break;
    case NGB_LEAKY_RELU:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = v[j] >= 0.f ? v[j] : alpha * v[j];
      break;
    case NGB_SIGMOID:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
      break;
    case NGB_TANH:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
      break;
    default: break;
  }
}
// v[j] *= d act(z[j]) / dz from the pre-activation z (the reference's masks: relu / leaky use z >= 0, activations.py:31, 216)
__device__ __forceinline__ void epi_mask4(int act, float alpha, float* v, const float4& z) {
  const float zz[4] = {z.x, z.y, z.z, z.w};
  switch (act) {
    case NGB_RELU:
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = zz[i] >= 0.f ? v[i] : 0.f;
      break;
    case NGB_LEAKY_RELU:
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = zz[i] >= 0.f ? v[i] : alpha * v[i];
      break;
    case NGB_SIGMOID:
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float t = __expf(-fabsf(zz[i])), d = 1.f + t; v[i] *= t / (d * d); }   // = s(1-s), no cancellation
      break;
    case NGB_TANH:
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float t = __expf(-2.f * fabsf(zz[i])), d = 1.f + t; v[i] *= 4.f * t / (d * d); }
      break;
    default: break;
  }
}

// A_MN / B_MN: operand is MN-major in global memory (1) or K-major (0).
template <int BLOCK_N, int A_MN, int B_MN, bool COLSUM = false, int CTAS = 1>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_c2, const GemmArgs g) {
  using C_ = Cfg<BLOCK_N, COLSUM, CTAS>;
  constexpr bool PAIR = CTAS == 2;
  static_assert(!(PAIR && COLSUM), "the ones-tile column sum is a single-CTA feature");
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs)
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;     // persistent work unit: a CTA, or a CTA pair
  const int num_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int TILE_M = BLOCK_M * CTAS;
  constexpr int kStages = C_::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * C_::kABytes;
  uint8_t* smem_epi = smem + kStages * C_::kStageBytes;
  uint8_t* smem_ones = smem_epi + C_::kEpiBytes;      // COLSUM: [128 x 32] tile of 1.0f (any operand layout of all-ones is all-ones)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * C_::kStageBytes + C_::kEpiBytes + C_::kOnesBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4 * CTAS); }   // (pair: both CTAs' epilogue warps)
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_2cta(tmem_slot, C_::kTmemCols);
    else tmem_alloc(tmem_slot, C_::kTmemCols);
  }
  if (COLSUM) {
    for (int i = threadIdx.x; i < C_::kOnesBytes / 16; i += kThreads)
      reinterpret_cast<float4*>(smem_ones)[i] = make_float4(1.f, 1.f, 1.f, 1.f);
    fence_proxy_async();       // generic-proxy writes -> visible to the tensor core's async-proxy reads
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();      // the peer's barriers exist before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_work = g.tiles_m * g.tiles_n * g.splits;     // (pair: tiles_m counts 256-row tiles)

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = unit; w < num_work; w += num_units) {
        const int split = w % g.splits, tile = w / g.splits;
        // this CTA's 128 rows of A, and (pair) its half of the B tile's columns
        const int m0 = (tile % g.tiles_m) * TILE_M + (int)cta_rank * BLOCK_M;
        const int n0 = (tile / g.tiles_m) * BLOCK_N + (int)cta_rank * (BLOCK_N / CTAS);
        constexpr int BN_LOAD = BLOCK_N / CTAS;
        const int kb0 = split * g.kb_per_split;
        const int kb1 = min(g.k_blocks, kb0 + g.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          // (pair: the leader's barrier counts the bytes of both CTAs' loads)
          if (cta_rank == 0) mbar_arrive_expect_tx(&full[s], C_::kStageBytes * CTAS);
          uint8_t* sa = smem_a + s * C_::kABytes;
          uint8_t* sb = smem_b + s * C_::kBBytes;
          const int k0 = kb * BLOCK_K;
          if (A_MN) {
            // global dims {M, K}: one {32 x 32} box per 32-wide M chunk; chunk c lands at sa + c*4096
            if (g.a_3d) {
              if (PAIR) tma_load_3d_2cta(sa, &tmap_a, &full[s], 0, k0, m0 / 32);
              else tma_load_3d(sa, &tmap_a, &full[s], 0, k0, m0 / 32);
            } else {
#pragma unroll
              for (int c = 0; c < BLOCK_M / 32; ++c) {
                if (PAIR) tma_load_2d_2cta(sa + c * 4096, &tmap_a, &full[s], m0 + c * 32, k0);
                else tma_load_2d(sa + c * 4096, &tmap_a, &full[s], m0 + c * 32, k0);
              }
            }
          } else {
            if (PAIR) tma_load_2d_2cta(sa, &tmap_a, &full[s], k0, m0);
            else tma_load_2d(sa, &tmap_a, &full[s], k0, m0);  // global dims {K, M}: box {32, 128}
          }
          if (B_MN) {
            if (g.b_3d) {
              if (PAIR) tma_load_3d_2cta(sb, &tmap_b, &full[s], 0, k0, n0 / 32);
              else tma_load_3d(sb, &tmap_b, &full[s], 0, k0, n0 / 32);
            } else {
#pragma unroll
              for (int c = 0; c < BN_LOAD / 32; ++c) {
                if (PAIR) tma_load_2d_2cta(sb + c * 4096, &tmap_b, &full[s], n0 + c * 32, k0);
                else tma_load_2d(sb + c * 4096, &tmap_b, &full[s], n0 + c * 32, k0);
              }
            }
          } else {
            if (PAIR) tma_load_2d_2cta(sb, &tmap_b, &full[s], k0, n0);
            else tma_load_2d(sb, &tmap_b, &full[s], k0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TILE_M, BLOCK_N, A_MN, B_MN);
      // K-major SW128: 8-row groups 1024 B apart (SBO); LBO unused.  MN-major tf32 must use the 32-byte-atom variant of
      // the 128B swizzle (tools/tc_probe.cu measured every other encoding to give zeros or garbage): 32-element MN
      // chunks 4096 B apart (LBO), 4-row K groups 512 B apart (SBO); one UMMA_K = 8 K-rows = 1024 B.
      constexpr uint64_t adesc_base = A_MN ? make_smem_desc_base(4096, 512, kLayoutSW128Base32B) : make_smem_desc_base(16, 1024);
      constexpr uint64_t bdesc_base = B_MN ? make_smem_desc_base(4096, 512, kLayoutSW128Base32B) : make_smem_desc_base(16, 1024);
      constexpr uint32_t a_kstep = A_MN ? 1024 : UMMA_K * 4;  // bytes to advance the start address per UMMA_K
      constexpr uint32_t b_kstep = B_MN ? 1024 : UMMA_K * 4;
      uint32_t it = 0, tile_it = 0;
      for (int w = unit; w < num_work; w += num_units, ++tile_it) {
        const int split = w % g.splits;
        const int kb0 = split * g.kb_per_split;
        const int kb1 = min(g.k_blocks, kb0 + g.kb_per_split);
        const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C_::kAccCols;
        const bool do_colsum = COLSUM && ((w / g.splits) % g.tiles_m) == 0;   // first row of output tiles only
        const uint32_t ones_addr = smem_u32(smem_ones);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + s * C_::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + s * C_::kBBytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if (PAIR)
              umma_tf32_2cta(d_tmem, smem_desc(adesc_base, a_addr + k * a_kstep), smem_desc(bdesc_base, b_addr + k * b_kstep),
                             idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else
              umma_tf32(d_tmem, smem_desc(adesc_base, a_addr + k * a_kstep), smem_desc(bdesc_base, b_addr + k * b_kstep),
                        idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            if (COLSUM && do_colsum)
              umma_tf32(d_tmem + BLOCK_N, smem_desc(adesc_base, ones_addr + k * a_kstep),
                        smem_desc(bdesc_base, b_addr + k * b_kstep), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (PAIR) umma_commit_2cta(&empty[s], 0b11);   // frees the slot in BOTH CTAs once these MMAs have read it
          else umma_commit(&empty[s]);                    // frees the smem slot once these MMAs have read it
        }
        if (PAIR) umma_commit_2cta(&tmem_full[acc], 0b11);
        else umma_commit(&tmem_full[acc]);  // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int et = (warp - 2) * 32 + lane;
    uint32_t tile_it = 0, chunk_it = 0;
    const bool vec_ok = g.partial ? (g.N % 4 == 0) : ((g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0));
    for (int w = unit; w < num_work; w += num_units, ++tile_it) {
      const int split = w % g.splits, tile = w / g.splits;
      const int m0 = (tile % g.tiles_m) * TILE_M + (int)cta_rank * BLOCK_M, n0 = (tile / g.tiles_m) * BLOCK_N;
      const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      const int lrow = quad * 32 + lane;
      const int row = m0 + lrow;
      float* out_row;
      int64_t ld;
      if (g.partial) { ld = g.N; out_row = g.partial + (int64_t)split * g.M * g.N + (int64_t)row * ld; }
      else { ld = g.ldc; out_row = g.C + (int64_t)row * ld; }
      const bool accum = !g.partial && g.accumulate;
      const bool tile_colsum = COLSUM && (tile % g.tiles_m) == 0;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + (uint32_t(quad * 32) << 16) + acc * C_::kAccCols + c0, v);
        tmem_ld_wait();
        if (COLSUM && tile_colsum && quad == 0) {
          // second accumulator: every row holds the column sums of the B operand; lane 0 (row 0) writes them
          float cs[32];
          tmem_ld32(tmem_base + acc * C_::kAccCols + BLOCK_N + c0, cs);
          tmem_ld_wait();
          if (lane == 0) {
            float* dst = g.partial ? g.colsum + (int64_t)split * g.N : g.colsum;
            const bool cacc = !g.partial && g.colsum_accumulate;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + c0 + j;
              if (n < g.N) dst[n] = cacc ? dst[n] + cs[j] : cs[j];
            }
          }
        }
        if (c0 + 32 >= BLOCK_N) {  // last chunk read: hand the accumulator back before the (slow) global stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_remote(&tmem_empty[acc], 0);    // the leader's MMA warp waits for both CTAs' epilogues
            else mbar_arrive(&tmem_empty[acc]);
          }
        }
        const int nbase = n0 + c0;
        if (nbase >= g.N) continue;           // (uniform across the CTA)
        // ---- fused epilogue on the registers: + bias, * act'(z_prev)
        if (g.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += (nbase + j < g.N) ? __ldg(g.bias + nbase + j) : 0.f;
        }
        // act'(z_prev) mask: on the TMA-store path it is applied to the staged tile with COALESCED loads of z_prev (below);
        // a thread reading the 128 bytes of its own row here costs a 32-line gather per load instruction
        const bool mask_in_smem = g.mask_z && !g.C2 && g.tma_store && (g.ldm % 4 == 0) && nbase + 32 <= g.N &&
                                  ((reinterpret_cast<uintptr_t>(g.mask_z) & 15) == 0);
        if (g.mask_z && !mask_in_smem && row < g.M) {
          const float* zr = g.mask_z + (int64_t)row * g.ldm + nbase;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 z4;
            z4.x = nbase + j < g.N ? zr[j] : 0.f;
            z4.y = nbase + j + 1 < g.N ? zr[j + 1] : 0.f;
            z4.z = nbase + j + 2 < g.N ? zr[j + 2] : 0.f;
            z4.w = nbase + j + 3 < g.N ? zr[j + 3] : 0.f;
            epi_mask4(g.act, g.alpha, v + j, z4);
          }
        }
        // ---- up to two outputs: C (the values themselves) and C2 = act(values)
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
          if (which == 0 && !g.store_c) continue;
          if (which == 1) {
            if (!g.C2) break;
            epi_act32(g.act, g.alpha, v);
          }
          if (g.tma_store) {
            // registers -> smem tile [128 rows][32 cols] in the 128B-swizzle the store tensor map expects -> one TMA store
            // (whole 128-byte lines, M / N tails clipped by the tensor map).  Two tiles alternate so the store of chunk i
            // overlaps the TMEM reads of chunk i+1.
            uint8_t* tile_s = smem_epi + (chunk_it & 1) * (128 * 128);
            if (et == 0) tma_store_wait_read<1>();          // the store that last read this tile has drained it
            asm volatile("bar.sync 1, 128;" ::: "memory");
            uint8_t* rp = tile_s + lrow * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(rp + ((j ^ (lrow & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (mask_in_smem && which == 0) {
              // second pass over the staged tile, thread <-> (row, 16-byte column group) so that a warp reads four whole
              // 128-byte lines of z_prev per instruction
              asm volatile("bar.sync 1, 128;" ::: "memory");
              float4 z4[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int idx = i * 128 + et, r = idx >> 3, c16 = idx & 7;
                z4[i] = (m0 + r < g.M) ? *reinterpret_cast<const float4*>(g.mask_z + (int64_t)(m0 + r) * g.ldm + nbase + c16 * 4)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int idx = i * 128 + et, r = idx >> 3, c16 = idx & 7;
                float4* sp = reinterpret_cast<float4*>(tile_s + r * 128 + ((c16 ^ (r & 7)) << 4));
                float4 t4 = *sp;
                float tv[4] = {t4.x, t4.y, t4.z, t4.w};
                epi_mask4(g.act, g.alpha, tv, z4[i]);
                *sp = make_float4(tv[0], tv[1], tv[2], tv[3]);
              }
            }
            fence_proxy_async();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) {
              if (which == 1) tma_store_2d(&tmap_c2, tile_s, nbase, m0, false);
              else if (g.partial) tma_store_3d(&tmap_c, tile_s, nbase, m0, split);
              else tma_store_2d(&tmap_c, tile_s, nbase, m0, accum);
              tma_store_commit();
            }
            ++chunk_it;
          } else if (row < g.M) {
            float* orow = which == 1 ? g.C2 + (int64_t)row * g.ldc2 : out_row;
            const bool acc_here = which == 0 && accum;
            const bool v_ok = which == 1 ? ((g.ldc2 % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C2) & 15) == 0)) : vec_ok;
            if (v_ok && nbase + 32 <= g.N) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                float4* p = reinterpret_cast<float4*>(orow + nbase + j);
                if (acc_here) { float4 old = *p; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                *p = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (nbase + j < g.N) orow[nbase + j] = acc_here ? orow[nbase + j] + v[j] : v[j];
              }
            }
          }
        }
      }
    }
    if (g.tma_store && et == 0) tma_store_wait_all();   // global writes complete before the kernel (and the split-K pass) moves on
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();      // the leader's MMAs read the peer's shared memory: nobody leaves early
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2cta(tmem_base, C_::kTmemCols);
    else tmem_dealloc(tmem_base, C_::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------- host launch
template <int BLOCK_N, int A_MN, int B_MN, bool COLSUM, int CTAS = 1>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc_, const CUtensorMap& tc2, const GemmArgs& g,
                  int grid, cudaStream_t st) {
  using C_ = Cfg<BLOCK_N, COLSUM, CTAS>;
  auto kern = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN, COLSUM, CTAS>;
  static bool attr_set = false;
  if (!attr_set) {
    NGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::kSmemBytes));
    attr_set = true;
  }
  if (CTAS == 1) {
    kern<<<grid, kThreads, C_::kSmemBytes, st>>>(ta, tb, tc_, tc2, g);
  } else {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = C_::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NGB_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc_, tc2, g));
  }
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

template <int BLOCK_N>
static int launch_major(int a_mn, int b_mn, bool colsum, int ctas, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc_,
                        const CUtensorMap& tc2, const GemmArgs& g, int grid, cudaStream_t st) {
  if constexpr (BLOCK_N <= 128) {
    if (colsum) return launch<BLOCK_N, 1, 1, true>(ta, tb, tc_, tc2, g, grid, st);   // (only the TN weight-gradient form)
  }
  if constexpr (BLOCK_N >= 128) {
    if (ctas == 2) {
      if (!a_mn && !b_mn) return launch<BLOCK_N, 0, 0, false, 2>(ta, tb, tc_, tc2, g, grid, st);
      if (!a_mn && b_mn) return launch<BLOCK_N, 0, 1, false, 2>(ta, tb, tc_, tc2, g, grid, st);
      if (a_mn && !b_mn) return launch<BLOCK_N, 1, 0, false, 2>(ta, tb, tc_, tc2, g, grid, st);
      return launch<BLOCK_N, 1, 1, false, 2>(ta, tb, tc_, tc2, g, grid, st);
    }
  }
  if (!a_mn && !b_mn) return launch<BLOCK_N, 0, 0, false>(ta, tb, tc_, tc2, g, grid, st);
  if (!a_mn && b_mn) return launch<BLOCK_N, 0, 1, false>(ta, tb, tc_, tc2, g, grid, st);
  if (a_mn && !b_mn) return launch<BLOCK_N, 1, 0, false>(ta, tb, tc_, tc2, g, grid, st);
  return launch<BLOCK_N, 1, 1, false>(ta, tb, tc_, tc2, g, grid, st);
}

bool gemm_tc_eligible(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc) {
  (void)transA; (void)transB; (void)ldc;
  if (M < 64 || N < 32 || K < 32) return false;
  if (M * N * K < (int64_t(1) << 22)) return false;          // tiny problems: launch-bound, SIMT is as fast
  if (lda % 4 || ldb % 4) return false;                        // TMA: global strides are multiples of 16 bytes
  if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff) return false;
  return true;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

int gemm_tc(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
            int64_t ldb, float* C, int64_t ldc, int accumulate, cudaStream_t st, const Epilogue* epi) {
  if (!aligned16(A) || !aligned16(B)) return fail(NGB_ERR_UNSUPPORTED, "tcgen05 GEMM needs 16-byte aligned operands");
  const int a_mn = transA ? 1 : 0;  // A stored [K,M] (M contiguous) -> MN-major
  const int b_mn = transB ? 0 : 1;  // B stored [K,N] (N contiguous) -> MN-major; stored [N,K] -> K-major
  Epilogue e;
  if (epi) e = *epi;
  const bool colsum = e.colsum != nullptr;
  const bool fused = e.bias || e.mask_z || e.C2;     // needs the complete sums in the epilogue: no split-K
  if (colsum && !(a_mn && b_mn)) return fail(NGB_ERR_UNSUPPORTED, "tcgen05 GEMM: column sums only in the TN (weight-gradient) form");

  // Tile width and K-splits from a cost model fitted to a sweep on B200 (profiles/r02_gemm_tiles.md).  At these sizes the
  // kernel is bound by the shared-memory fill: one k-block brings (128 + BN) x 128 B through TMA at ~75 GB/s per SM (the
  // chip-wide L2 -> SM cap of ~6300 B/clk), a work item costs ~2.5 us of pipeline fill + epilogue on top, a CTA runs
  // ceil(items / 148) items in a row, and K-splits pay a second pass over (splits + 1) x M x N floats at ~3 TB/s.
  const int k_blocks = (int)ceil_div<int64_t>(K, BLOCK_K);
  const bool can_split = !fused && k_blocks >= 8;
  // CTA pairs (cta_group::2): 256-row tiles, each CTA stages 128 rows of A and HALF of the B tile -> (128 + BN/2) x 128 B
  // per k-block and SM instead of (128 + BN) x 128 B.  Work units are then the 74 pairs.
  static const int pair_env = env_int("NGB_GEMM_2CTA", -1);
  int BN = 64, best_splits = 1, ctas = 1;
  {
    double best = 1e30;
    for (int cg = 1; cg <= 2; ++cg) {
      // Measured (profiles/r02_gemm_pair.md): the pair kernels win where a CTA streams several tiles back to back (4096^3:
      // 211 -> 187 us) and lose 2-10 % on one-tile-per-CTA problems, whose time is the load -> MMA -> epilogue latency chain
      // of a single tile rather than the shared-memory fill; so they are taken from 3 rounds of 128 x 256 tiles up.
      // NGB_GEMM_2CTA = 0 / 1 forces them off / on (wherever they exist).
      const bool pair_ok = !colsum && M > BLOCK_M && N > 64;
      const bool big = ceil_div<int64_t>(M, BLOCK_M) * ceil_div<int64_t>(N, 256) >= 3 * (int64_t)kNumSMs;
      if (cg == 2 && (!pair_ok || pair_env == 0 || (pair_env != 1 && !big))) break;
      if (cg == 1 && pair_ok && (pair_env == 1 || (pair_env != 0 && big))) continue;
      const int64_t tm_c = ceil_div<int64_t>(M, (int64_t)BLOCK_M * cg);
      const int units = kNumSMs / cg;
      for (int c = cg == 2 ? 128 : 64; c <= 256; c *= 2) {
        if (colsum && c > 128) break;                     // two accumulators per stage: 4 x BN TMEM columns <= 512
        if (c > 64 && c / 2 >= N) break;                  // (a tile more than twice as wide as the problem)
        const int64_t tiles_c = tm_c * ceil_div<int64_t>(N, c);
        const int max_s = can_split ? (k_blocks / 2 < 64 ? k_blocks / 2 : 64) : 1;
        for (int sp = 1; sp <= max_s; ++sp) {
          const int kb = (int)ceil_div<int64_t>(k_blocks, sp);
          const int sp_eff = (int)ceil_div<int64_t>(k_blocks, kb);
          const double rounds = (double)ceil_div<int64_t>(tiles_c * sp_eff, units);
          double t = rounds * (2.5 + kb * (128.0 + (double)c / cg) * 128.0 / 75e3);
          if (sp_eff > 1) t += 2.5 + (sp_eff + 1.0) * (double)M * (double)N * 4.0 / 3.0e6;
          if (t < best) { best = t; BN = c; best_splits = sp_eff; ctas = cg; }
        }
      }
    }
    const int f = env_int("NGB_GEMM_BN", 0);          // tuning / experiments
    if (f == 64 || f == 128 || (f == 256 && !colsum)) BN = f;
    if (ctas == 2 && BN < 128) BN = 128;
  }
  const int64_t tm = ceil_div<int64_t>(M, (int64_t)BLOCK_M * ctas);

  CUtensorMap ta, tb;
  // An MN-major operand whose MN extent is a multiple of 32 is described as {32, K, MN/32} (outer stride 128 B) so that a
  // single 3-D box brings the whole [MN/32 chunks][32 k][32 mn] stage; otherwise one 2-D box per chunk (zero-filled tails).
  const int a_3d = a_mn && (M % 32 == 0), b_3d = b_mn && (N % 32 == 0);
  {
    uint64_t dims[3], str[2];
    uint32_t box[3];
    int rc;
    if (a_3d) {
      dims[0] = 32; dims[1] = K; dims[2] = M / 32; str[0] = (uint64_t)lda * 4; str[1] = 128;
      box[0] = 32; box[1] = BLOCK_K; box[2] = BLOCK_M / 32;
      rc = make_tmap_f32(&ta, A, 3, dims, str, box, true);
    } else {
      if (a_mn) { dims[0] = M; dims[1] = K; box[0] = 32; box[1] = BLOCK_K; }
      else { dims[0] = K; dims[1] = M; box[0] = BLOCK_K; box[1] = BLOCK_M; }
      str[0] = (uint64_t)lda * 4;
      rc = make_tmap_f32(&ta, A, 2, dims, str, box, a_mn);
    }
    if (rc) return rc;
    if (b_3d) {
      dims[0] = 32; dims[1] = K; dims[2] = N / 32; str[0] = (uint64_t)ldb * 4; str[1] = 128;
      box[0] = 32; box[1] = BLOCK_K; box[2] = (uint32_t)(BN / ctas / 32);      // (pair: each CTA loads half of the tile's columns)
      rc = make_tmap_f32(&tb, B, 3, dims, str, box, true);
    } else {
      if (b_mn) { dims[0] = N; dims[1] = K; box[0] = 32; box[1] = BLOCK_K; }
      else { dims[0] = K; dims[1] = N; box[0] = BLOCK_K; box[1] = (uint32_t)(BN / ctas); }
      str[0] = (uint64_t)ldb * 4;
      rc = make_tmap_f32(&tb, B, 2, dims, str, box, b_mn);
    }
    if (rc) return rc;
  }

  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = (int)M; g.N = (int)N;
  g.k_blocks = k_blocks;
  g.tiles_m = (int)tm;
  g.tiles_n = (int)ceil_div<int64_t>(N, BN);
  const int tiles = g.tiles_m * g.tiles_n;
  int splits = can_split ? best_splits : 1;
  {
    const int f = env_int("NGB_GEMM_SPLITS", 0);
    if (f > 0 && can_split) splits = f < g.k_blocks ? f : g.k_blocks;
    if (splits > 1) {
      if (ensure_init() != NGB_OK) return NGB_ERR_CUDA;
      const int64_t cap = scratch_bytes() / ((M * N + N) * (int64_t)sizeof(float));
      if (splits > cap) splits = (int)cap;
    }
    if (splits < 1) splits = 1;
  }
  g.kb_per_split = (int)ceil_div<int64_t>(g.k_blocks, splits);
  g.splits = (int)ceil_div<int64_t>(g.k_blocks, g.kb_per_split);
  g.C = C; g.ldc = ldc;
  g.partial = g.splits > 1 ? scratch_ptr() : nullptr;
  g.accumulate = accumulate;
  g.a_3d = a_3d; g.b_3d = b_3d;
  g.bias = e.bias; g.act = e.act; g.alpha = e.alpha; g.C2 = e.C2; g.ldc2 = e.ldc2; g.store_c = (C != nullptr && e.store_c) ? 1 : 0;
  g.mask_z = e.mask_z; g.ldm = e.ldm;
  float* cs_part = g.partial ? g.partial + (int64_t)g.splits * M * N : nullptr;
  g.colsum = colsum ? (g.partial ? cs_part : e.colsum) : nullptr;
  g.colsum_accumulate = e.colsum_accumulate;
  const int work = tiles * g.splits;
  const int units_avail = kNumSMs / ctas;
  const int grid = (work < units_avail ? work : units_avail) * ctas;

  // output tensor maps for the TMA-store epilogue: C itself, or the split partials as {N, M, splits}; C2 likewise
  CUtensorMap tcm, tcm2;
  memset(&tcm, 0, sizeof(tcm));
  memset(&tcm2, 0, sizeof(tcm2));
  g.tma_store = 0;
  {
    const bool to_partial = g.splits > 1;
    const float* obase = to_partial ? g.partial : C;
    const int64_t old = to_partial ? N : ldc;
    static const bool no_tma_store = getenv("NGB_NO_TMA_STORE") != nullptr;
    bool ok = !no_tma_store;
    if (g.store_c) ok = ok && old % 4 == 0 && aligned16(obase);
    if (g.C2) ok = ok && g.ldc2 % 4 == 0 && aligned16(g.C2);
    if (ok && g.store_c) {
      uint64_t dims[3] = {(uint64_t)N, (uint64_t)M, (uint64_t)g.splits}, str[2] = {(uint64_t)old * 4, (uint64_t)M * N * 4};
      uint32_t box[3] = {32, 128, 1};
      ok = make_tmap_f32(&tcm, obase, to_partial ? 3 : 2, dims, str, box, false, /*plain_f32=*/true) == NGB_OK;
    }
    if (ok && g.C2) {
      uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, str[1] = {(uint64_t)g.ldc2 * 4};
      uint32_t box[2] = {32, 128};
      ok = make_tmap_f32(&tcm2, g.C2, 2, dims, str, box, false, /*plain_f32=*/true) == NGB_OK;
    }
    g.tma_store = ok ? 1 : 0;
  }
  int rc;
  if (BN == 64) rc = launch_major<64>(a_mn, b_mn, colsum, 1, ta, tb, tcm, tcm2, g, grid, st);
  else if (BN == 128) rc = launch_major<128>(a_mn, b_mn, colsum, ctas, ta, tb, tcm, tcm2, g, grid, st);
  else rc = launch_major<256>(a_mn, b_mn, colsum, ctas, ta, tb, tcm, tcm2, g, grid, st);
  if (rc) return rc;
  if (g.splits > 1) {
    if (!colsum) return launch_splitk_reduce(g.partial, g.splits, M, N, C, ldc, accumulate, st);
    return launch_splitk_reduce2(g.partial, g.splits, M, N, C, ldc, accumulate, cs_part, e.colsum, e.colsum_accumulate, st);
  }
  return NGB_OK;
}

}  // namespace tc
}  // namespace ngb

using namespace ngb;

namespace ngb {
std::atomic<int> g_default_engine{NGB_GEMM_AUTO};
unsigned long long gemm_take_timeout() { return tc::take_mbar_timeout(); }
int gemm_set_host_word(unsigned long long* dev_alias) { return tc::set_mbar_host_word(dev_alias); }
}

extern "C" int ngb_set_default_engine(int engine) {
  if (engine != NGB_GEMM_AUTO && engine != NGB_GEMM_SIMT) return g_default_engine.load();
  return g_default_engine.exchange(engine);
}

extern "C" int ngb_gemm_query(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                              int64_t ldc) {
  if (g_default_engine.load() == NGB_GEMM_SIMT) return NGB_GEMM_SIMT;
  return tc::gemm_tc_eligible(transA, transB, M, N, K, lda, ldb, ldc) ? NGB_GEMM_TCGEN05 : NGB_GEMM_SIMT;
}

extern "C" int ngb_gemm(int engine, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                        const float* B, int64_t ldb, float* C, int64_t ldc, int accumulate, ngb_stream_t stream) {
  NGB_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "ngb_gemm: negative dimension");
  NGB_CHECK_ARG(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "ngb_gemm: leading dimension too small");
  cudaStream_t st = as_stream(stream);
  if (M == 0 || N == 0) return NGB_OK;
  if (K == 0) {
    if (!accumulate) {
      for (int64_t m = 0; m < M; ++m) NGB_CUDA(cudaMemsetAsync(C + m * ldc, 0, N * sizeof(float), st));
    }
    return NGB_OK;
  }
  bool use_tc = false;
  if (engine == NGB_GEMM_AUTO) engine = g_default_engine.load();
  if (engine == NGB_GEMM_TCGEN05) {
    if (lda % 4 || ldb % 4 || K < 8) return fail(NGB_ERR_UNSUPPORTED, "shape not eligible for the tcgen05 GEMM");
    use_tc = true;
  } else if (engine == NGB_GEMM_AUTO) {
    use_tc = tc::gemm_tc_eligible(transA, transB, M, N, K, lda, ldb, ldc) && aligned16(A) && aligned16(B);
  }
  if (use_tc) return tc::gemm_tc(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, accumulate, st);
  return gemm_simt(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, accumulate, st);
}
