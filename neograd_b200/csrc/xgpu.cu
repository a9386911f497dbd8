// Data-parallel gradient exchange FUSED WITH the optimizer update, over NVLink / NVSwitch peer memory (SURVEY 8e;
// reference: the update it wraps is neograd/nn/optim.py:42-47, 63-92, 112-140, 168-197).
//
// One kernel per optimizer step does what used to be barrier -> reduce -> barrier -> tick -> update (5 launches):
//   1. PUSH   every block copies its chunk of this GPU's flat gradient range into slot [rank] of EVERY peer's receive
//             buffer with 128-bit stores that travel over NVLink (the receive buffers live in symmetric memory);
//   2. SIGNAL a system-scope release store of the step's epoch into the peer's flag word [rank][block];
//   3. WAIT   for the flag words [src][block] of all peers (acquire loads of LOCAL memory -- no remote polling);
//   4. REDUCE the world slots of the chunk in RANK ORDER (own gradient read in place), so every replica computes the
//             bit-identical sum, scale by grad_scale (1/world), and
//   5. UPDATE parameters and optimizer state of the chunk (Adam / GD / Momentum / RMSProp) in the same pass.
// Properties
//   * no grid-wide barrier and no circular wait: pushes never wait for anything, so a block can always make progress once
//     its peers' same-numbered blocks have been scheduled -- blocks need not be co-resident;
//   * receive slots are double buffered by the PARITY OF THE EPOCH, which lives in device memory and is advanced by the
//     kernel itself: a captured CUDA graph replays with fixed addresses and still alternates.  A peer can only be pushing
//     epoch e+2 into the slot this GPU reads at epoch e after it has completed epoch e+1, which needed this GPU's pushes
//     of e+1, which this GPU issues after finishing e (stream order): one synchronisation per step suffices, and there is
//     no trailing barrier;
//   * a peer that never shows up does NOT let the step proceed: after `timeout_ns` the kernel sets a sticky failure word
//     (device + host-mapped), skips the update, and every later exchange is a no-op; the host polls the mapped word
//     (ngb_xgpu_status, no synchronisation) and raises.
#include "common.cuh"

namespace ngb {

// sticky failure word: 0 = healthy; else (1<<63) | rank << 32 | missing peer
static __device__ unsigned long long g_xgpu_failed = 0;
static unsigned long long* g_host_status = nullptr;        // host-mapped mirror (cudaHostAllocMapped)
static unsigned long long* g_host_status_dev = nullptr;    // its device alias

__device__ __forceinline__ uint64_t gtimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// data written by a peer over NVLink: never from a (possibly stale) L1 line
__device__ __forceinline__ float4 ld_sys4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_sys1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

enum { XOPT_ADAM = 0, XOPT_SGD = 1, XOPT_MOMENTUM = 2, XOPT_RMSPROP = 3 };

struct XOpt {
  int kind;
  float lr, b1, omb1, b2, omb2, eps, gs;   // (Momentum / RMSProp use b1 / omb1 as beta / 1 - beta)
  double b1d, b2d;
  int tick;                                // Adam: 1 = this launch is step t+1 and records it, 2 = step t+1 not recorded, 0 = t
};

struct XArgs {
  float* const* recv;        // [world] device pointers: peer r's receive buffer  [2 parities][world slots][slot_stride]
  uint32_t* const* flags;    // [world] device pointers: peer r's flag words       [world][max_blocks]
  int rank, world, max_blocks;
  int64_t slot_stride;       // floats per slot (the arena size, padded)
  int64_t off, n;            // the range [off, off + n) of the arena this launch exchanges and updates
  const float* g;            // local gradient arena
  float *p, *s1, *s2;        // parameter arena, state arenas (Adam m / v; Momentum m; RMSProp r)
  uint32_t* epoch;           // [0] = exchanges completed, [1] = blocks-done counter (0 between launches)
  int32_t* t_dev;            // Adam: [0] = steps taken; nullptr otherwise
  unsigned long long timeout_ns;
  unsigned long long* host_status;
};

__device__ __forceinline__ void xupdate(const XOpt& o, float ibc1, float ibc2, float& p, float g, float& a, float& b) {
  g *= o.gs;
  switch (o.kind) {
    case XOPT_ADAM: {
      a = o.b1 * a + o.omb1 * g;                  // optim.py:196
      b = o.b2 * b + o.omb2 * (g * g);            // optim.py:197
      p -= o.lr * ((a * ibc1) / (sqrtf(b * ibc2) + o.eps));   // optim.py:177-179
      break;
    }
    case XOPT_SGD: p -= o.lr * g; break;          // optim.py:47
    case XOPT_MOMENTUM: a = o.b1 * a + o.omb1 * g; p -= o.lr * a; break;                              // optim.py:92, 70
    default: a = o.b1 * a + o.omb1 * (g * g); p -= o.lr * (g / (sqrtf(a) + o.eps)); break;          // optim.py:140, 118
  }
}

__global__ void __launch_bounds__(256) xgpu_exchange_update_kernel(const XArgs x, const XOpt o) {
  __shared__ int ok_sh;
  const int tid = threadIdx.x, nthr = blockDim.x, b = blockIdx.x, nb = gridDim.x;
  // sticky failure: nothing is exchanged or updated any more (the host raises at its next poll)
  if (*reinterpret_cast<volatile unsigned long long*>(&g_xgpu_failed) != 0) return;
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(x.epoch) + 1;
  const int par = (int)(e & 1u);
  // chunk of the range owned by this block (multiples of 4 floats; the last block takes the tail)
  const int64_t n4 = (x.n + 3) >> 2;
  const int64_t per = (n4 + nb - 1) / nb;
  const int64_t c0 = min(n4, (int64_t)b * per) * 4, c1 = min(x.n, min(n4, (int64_t)(b + 1) * per) * 4);
  const float* gl = x.g + x.off;
  const bool vec = ((x.off & 3) == 0) && ((reinterpret_cast<uintptr_t>(x.g) & 15) == 0);

  // ---- 1. push this chunk into slot [rank] of every peer (remote stores), 2. signal
  for (int d = 1; d < x.world; ++d) {
    const int r = (x.rank + d) % x.world;                 // staggered so the ranks do not all hit the same peer at once
    float* dst = x.recv[r] + ((int64_t)par * x.world + x.rank) * x.slot_stride + x.off;
    if (vec) {
      for (int64_t i = c0 + 4 * tid; i + 3 < c1; i += 4 * nthr)
        *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(gl + i);
      const int64_t tail = c0 + ((c1 - c0) & ~int64_t(3));
      for (int64_t i = tail + tid; i < c1; i += nthr) dst[i] = gl[i];
    } else {
      for (int64_t i = c0 + tid; i < c1; i += nthr) dst[i] = gl[i];
    }
  }
  __syncthreads();
  if (tid == 0) __threadfence_system();                   // (cumulative: covers the block's remote stores ordered by the barrier)
  __syncthreads();
  if (tid > 0 && tid < x.world) {
    const int r = (x.rank + tid) % x.world;
    st_release_sys(x.flags[r] + (int64_t)x.rank * x.max_blocks + b, e);
  }

  // ---- 3. wait for every peer's push of this chunk (local flag words)
  if (tid == 0) ok_sh = 1;
  __syncthreads();
  if (tid > 0 && tid < x.world) {
    const int r = (x.rank + tid) % x.world;
    const uint32_t* mine = x.flags[x.rank] + (int64_t)r * x.max_blocks + b;
    const uint64_t t0 = gtimer_ns();
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
      if ((++spins & 0x3FF) == 0) {
        if (*reinterpret_cast<volatile unsigned long long*>(&g_xgpu_failed) != 0 || gtimer_ns() - t0 > x.timeout_ns) {
          const unsigned long long code = (1ull << 63) | ((unsigned long long)x.rank << 32) | (unsigned long long)r;
          atomicCAS(&g_xgpu_failed, 0ull, code);
          if (x.host_status) *reinterpret_cast<volatile unsigned long long*>(x.host_status) = code;
          __threadfence_system();
          ok_sh = 0;
          break;
        }
      }
    }
  }
  __syncthreads();
  if (!ok_sh) return;                                     // the update is NOT applied on a failed exchange

  // ---- 4. reduce in rank order + 5. update
  float ibc1 = 1.f, ibc2 = 1.f;
  int32_t step = 0;
  if (o.kind == XOPT_ADAM) {
    step = *reinterpret_cast<volatile int32_t*>(x.t_dev) + (o.tick ? 1 : 0);
    ibc1 = (float)(1.0 / (1.0 - pow(o.b1d, (double)step)));
    ibc2 = (float)(1.0 / (1.0 - pow(o.b2d, (double)step)));
  }
  const float* mine_base = x.recv[x.rank] + (int64_t)par * x.world * x.slot_stride + x.off;
  float* pp = x.p + x.off;
  float* s1 = x.s1 ? x.s1 + x.off : nullptr;
  float* s2 = x.s2 ? x.s2 + x.off : nullptr;
  if (vec) {
    for (int64_t i = c0 + 4 * tid; i + 3 < c1; i += 4 * nthr) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < x.world; ++r) {
        const float4 v = r == x.rank ? *reinterpret_cast<const float4*>(gl + i) : ld_sys4(mine_base + (int64_t)r * x.slot_stride + i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      float4 pv = *reinterpret_cast<float4*>(pp + i);
      float4 a = s1 ? *reinterpret_cast<float4*>(s1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 bb = s2 ? *reinterpret_cast<float4*>(s2 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      xupdate(o, ibc1, ibc2, pv.x, s.x, a.x, bb.x); xupdate(o, ibc1, ibc2, pv.y, s.y, a.y, bb.y);
      xupdate(o, ibc1, ibc2, pv.z, s.z, a.z, bb.z); xupdate(o, ibc1, ibc2, pv.w, s.w, a.w, bb.w);
      *reinterpret_cast<float4*>(pp + i) = pv;
      if (s1) *reinterpret_cast<float4*>(s1 + i) = a;
      if (s2) *reinterpret_cast<float4*>(s2 + i) = bb;
    }
  }
  {
    const int64_t tail = vec ? c0 + ((c1 - c0) & ~int64_t(3)) : c0;
    for (int64_t i = tail + tid; i < c1; i += nthr) {
      float s = 0.f;
      for (int r = 0; r < x.world; ++r) s += r == x.rank ? gl[i] : ld_sys1(mine_base + (int64_t)r * x.slot_stride + i);
      float pv = pp[i], a = s1 ? s1[i] : 0.f, bb = s2 ? s2[i] : 0.f;
      xupdate(o, ibc1, ibc2, pv, s, a, bb);
      pp[i] = pv;
      if (s1) s1[i] = a;
      if (s2) s2[i] = bb;
    }
  }
  // ---- the last block to finish records the new epoch (and Adam's step count): every block has read the old values
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(x.epoch + 1, 1u) == (unsigned)nb - 1) {
      x.epoch[0] = e;
      x.epoch[1] = 0;
      if (o.kind == XOPT_ADAM && o.tick == 1) x.t_dev[0] = step;
    }
  }
}

// ---- low-latency variant for small arenas ---------------------------------------------------------------------------
// The kernel above costs a system-scope fence (every remote store of the block acknowledged over NVLink) plus a flag round
// trip before anybody can read: ~25 us of pure latency, measured, against 6 us for the update itself -- for LeNet's 178 KB of
// gradients that IS the multi-GPU overhead.  Here every float travels as ONE 8-byte word {value, epoch} (a naturally aligned
// 64-bit store is single-copy atomic, also across NVLink): the receiver polls the words it needs until their epoch half
// matches, so there is no fence, no flag array and no block-level handshake -- the latency is one NVLink store.  Twice the
// bytes, which is why it serves ranges up to kLLMaxFloats only; slots are parity double-buffered exactly as above (a word
// of epoch e can only be overwritten by e+2).  One float4 group per thread; the sum runs over the ranks in order as above.
constexpr int64_t kLLMaxFloats = (int64_t)kNumSMs * 256 * 4;

__device__ __forceinline__ void st_ll(unsigned long long* p, float v, uint32_t e) {
  const unsigned long long w = ((unsigned long long)e << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_ll(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}

__global__ void __launch_bounds__(256) xgpu_exchange_update_ll_kernel(const XArgs x, const XOpt o) {
  const int tid = threadIdx.x, nb = gridDim.x;
  if (*reinterpret_cast<volatile unsigned long long*>(&g_xgpu_failed) != 0) return;
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(x.epoch) + 1;
  const int par = (int)(e & 1u);
  const int64_t i0 = ((int64_t)blockIdx.x * 256 + tid) * 4;           // this thread's (up to) four elements of the range
  const int cnt = i0 >= x.n ? 0 : (x.n - i0 < 4 ? (int)(x.n - i0) : 4);
  const float* gl = x.g + x.off + i0;
  float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < cnt) g[j] = gl[j];

  // ---- push {value, epoch} words into slot [rank] of every peer
  for (int d = 1; d < x.world; ++d) {
    const int r = (x.rank + d) % x.world;
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(x.recv[r] + ((int64_t)par * x.world + x.rank) * x.slot_stride) + x.off + i0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < cnt) st_ll(dst + j, g[j], e);
  }

  // ---- gather the peers' words of the same elements; every word validates itself.  All loads of a round are issued before
  // any is looked at (28 words at 8 GPUs: polled one after the other they cost 28 L2 round trips, ~20 us), then the words
  // that have not arrived yet are asked for again.  The sum runs over the ranks in order.
  constexpr int PB = 8;                                    // peers per polling batch
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  bool ok = true;
  const uint64_t t0 = gtimer_ns();
  for (int base = 0; base < x.world && ok; base += PB) {
    unsigned long long w[PB][4];
    uint32_t pending = 0;
#pragma unroll
    for (int p = 0; p < PB; ++p)
      if (base + p < x.world && base + p != x.rank && cnt > 0) pending |= 1u << p;
    uint32_t spins = 0;
    while (pending) {
#pragma unroll
      for (int p = 0; p < PB; ++p) {
        if (!((pending >> p) & 1u)) continue;
        const unsigned long long* src =
            reinterpret_cast<const unsigned long long*>(x.recv[x.rank] + ((int64_t)par * x.world + base + p) * x.slot_stride) + x.off + i0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < cnt) w[p][j] = ld_ll(src + j);
      }
#pragma unroll
      for (int p = 0; p < PB; ++p) {
        if (!((pending >> p) & 1u)) continue;
        bool valid = true;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < cnt) valid = valid && (uint32_t)(w[p][j] >> 32) == e;
        if (valid) pending &= ~(1u << p);
      }
      if (pending && (++spins & 0xFF) == 0 &&
          (*reinterpret_cast<volatile unsigned long long*>(&g_xgpu_failed) != 0 || gtimer_ns() - t0 > x.timeout_ns)) {
        const int r = base + (__ffs(pending) - 1);
        const unsigned long long code = (1ull << 63) | ((unsigned long long)x.rank << 32) | (unsigned long long)r;
        atomicCAS(&g_xgpu_failed, 0ull, code);
        if (x.host_status) *reinterpret_cast<volatile unsigned long long*>(x.host_status) = code;
        __threadfence_system();
        ok = false;
        break;
      }
    }
    if (!ok) break;
#pragma unroll
    for (int p = 0; p < PB; ++p) {
      const int r = base + p;
      if (r >= x.world) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j >= cnt) continue;
        s[j] += r == x.rank ? g[j] : __uint_as_float((uint32_t)w[p][j]);
      }
    }
  }
  if (!__syncthreads_and(ok ? 1 : 0)) return;              // a block applies its update only if every word it needs arrived

  // ---- update
  if (cnt > 0) {
    float ibc1 = 1.f, ibc2 = 1.f;
    if (o.kind == XOPT_ADAM) {
      const int32_t step = *reinterpret_cast<volatile int32_t*>(x.t_dev) + (o.tick ? 1 : 0);
      ibc1 = (float)(1.0 / (1.0 - pow(o.b1d, (double)step)));
      ibc2 = (float)(1.0 / (1.0 - pow(o.b2d, (double)step)));
    }
    float* pp = x.p + x.off + i0;
    float* s1 = x.s1 ? x.s1 + x.off + i0 : nullptr;
    float* s2 = x.s2 ? x.s2 + x.off + i0 : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= cnt) continue;
      float pv = pp[j], a = s1 ? s1[j] : 0.f, bb = s2 ? s2[j] : 0.f;
      xupdate(o, ibc1, ibc2, pv, s[j], a, bb);
      pp[j] = pv;
      if (s1) s1[j] = a;
      if (s2) s2[j] = bb;
    }
  }
  // ---- the last block to finish records the new epoch (and Adam's step count): every block has read the old values
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(x.epoch + 1, 1u) == (unsigned)nb - 1) {
      x.epoch[0] = e;
      x.epoch[1] = 0;
      if (o.kind == XOPT_ADAM && o.tick == 1) x.t_dev[0] = *reinterpret_cast<volatile int32_t*>(x.t_dev) + 1;
    }
  }
}

}  // namespace ngb

using namespace ngb;

static int ensure_host_status() {
  if (g_host_status) return NGB_OK;
  NGB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_host_status), sizeof(unsigned long long), cudaHostAllocMapped));
  *g_host_status = 0;
  NGB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_host_status_dev), g_host_status, 0));
  return NGB_OK;
}

// Number of blocks ngb_xgpu_exchange_update launches for a range of n floats (identical on every rank: the flag words are
// indexed by block).  At most `max_blocks` (the flag array's extent).
extern "C" int ngb_xgpu_blocks(int64_t n, int max_blocks) {
  int64_t nb = (n + 4095) / 4096;          // >= 16 KB per block: tiny arenas (LeNet: 178 KB) stay latency-bound, not flag-bound
  if (nb < 1) nb = 1;
  if (nb > kNumSMs) nb = kNumSMs;
  if (nb > max_blocks) nb = max_blocks;
  return (int)nb;
}

extern "C" int ngb_xgpu_exchange_update(const void* recv_ptrs_dev, const void* flag_ptrs_dev, int rank, int world, int max_blocks,
                                        int64_t slot_stride, int64_t off, int64_t n, const float* grad, float* params, float* state1,
                                        float* state2, uint32_t* epoch_dev, int kind, double lr, double beta1, double beta2,
                                        double epsilon, int32_t* t_dev, int tick, double grad_scale, double timeout_s,
                                        ngb_stream_t stream) {
  NGB_CHECK_ARG(recv_ptrs_dev && flag_ptrs_dev && epoch_dev && grad && params, "ngb_xgpu_exchange_update: NULL argument");
  NGB_CHECK_ARG(world >= 2 && world <= 32 && rank >= 0 && rank < world && max_blocks >= 1, "ngb_xgpu_exchange_update: bad rank / world");
  NGB_CHECK_ARG(off >= 0 && n >= 0 && off + n <= slot_stride, "ngb_xgpu_exchange_update: range outside the arena");
  NGB_CHECK_ARG(kind >= XOPT_ADAM && kind <= XOPT_RMSPROP, "ngb_xgpu_exchange_update: unknown optimizer %d", kind);
  NGB_CHECK_ARG(kind != XOPT_ADAM || (t_dev && state1 && state2), "ngb_xgpu_exchange_update: Adam needs t_dev and both state arenas");
  NGB_CHECK_ARG((kind != XOPT_MOMENTUM && kind != XOPT_RMSPROP) || state1, "ngb_xgpu_exchange_update: missing state arena");
  int rc = ensure_host_status();
  if (rc) return rc;
  XArgs x;
  x.recv = reinterpret_cast<float* const*>(recv_ptrs_dev);
  x.flags = reinterpret_cast<uint32_t* const*>(flag_ptrs_dev);
  x.rank = rank; x.world = world; x.max_blocks = max_blocks; x.slot_stride = slot_stride; x.off = off; x.n = n;
  x.g = grad; x.p = params; x.s1 = state1; x.s2 = state2; x.epoch = epoch_dev; x.t_dev = t_dev;
  x.timeout_ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 60.0) * 1e9);
  x.host_status = g_host_status_dev;
  XOpt o;
  o.kind = kind; o.lr = (float)lr; o.b1 = (float)beta1; o.omb1 = (float)(1.0 - beta1); o.b2 = (float)beta2;
  o.omb2 = (float)(1.0 - beta2); o.eps = (float)epsilon; o.gs = (float)grad_scale; o.b1d = beta1; o.b2d = beta2; o.tick = tick;
  // small ranges in a receive buffer whose slots have room for {value, epoch} words: the low-latency protocol (same on every
  // rank: the decision depends on the arguments only)
  if (n >= 1 && n <= kLLMaxFloats && 2 * (off + n) <= slot_stride && getenv("NGB_XGPU_NO_LL") == nullptr) {
    const int nb = (int)((((n + 3) >> 2) + 255) / 256);
    xgpu_exchange_update_ll_kernel<<<nb, 256, 0, as_stream(stream)>>>(x, o);
    NGB_LAUNCH_CHECK();
    return NGB_OK;
  }
  const int nb = ngb_xgpu_blocks(n, max_blocks);
  xgpu_exchange_update_kernel<<<nb, 256, 0, as_stream(stream)>>>(x, o);
  NGB_LAUNCH_CHECK();
  return NGB_OK;
}

// Non-blocking health check of the peer exchange: 0 = healthy, else NGB_ERR_CUDA with the (rank, missing peer) in
// ngb_last_error().  Reads a host-mapped word the kernel writes on failure: no synchronisation, cheap enough to poll.
extern "C" int ngb_xgpu_status(void) {
  if (!g_host_status) return NGB_OK;
  const unsigned long long v = *reinterpret_cast<volatile unsigned long long*>(g_host_status);
  if (v)
    return fail(NGB_ERR_CUDA, "peer-memory gradient exchange failed: rank %llu timed out waiting for rank %llu; no further updates are applied",
                (v >> 32) & 0x7fffffffull, v & 0xffffffffull);
  return NGB_OK;
}

// Synchronising variant (tests / tools): also reads the device-side word.
extern "C" int ngb_xgpu_check(void) {
  NGB_CUDA(cudaDeviceSynchronize());
  unsigned long long v = 0;
  NGB_CUDA(cudaMemcpyFromSymbol(&v, g_xgpu_failed, sizeof(v)));
  if (v)
    return fail(NGB_ERR_CUDA, "peer-memory gradient exchange failed: rank %llu timed out waiting for rank %llu", (v >> 32) & 0x7fffffffull,
                v & 0xffffffffull);
  return NGB_OK;
}

// Clears the sticky failure (tests that provoke a timeout on purpose).
extern "C" int ngb_xgpu_reset(void) {
  NGB_CUDA(cudaDeviceSynchronize());
  unsigned long long z = 0;
  NGB_CUDA(cudaMemcpyToSymbol(g_xgpu_failed, &z, sizeof(z)));
  if (g_host_status) *g_host_status = 0;
  return NGB_OK;
}
